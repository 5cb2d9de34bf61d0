// tsm_diff_kernels.cuh - S8 revision-pair churn (docs/SPEC.md section 8).  Placeholder: filled in below.
#pragma once
#include "tsm_device.cuh"
