#!/bin/bash
# Time every build variant under tosem-2021-replication_b200/build_variants/ (one gpurun call):
#   gpurun -- 'bash tools/run_variants.sh > gpurun_out/variants.txt'
cd "$(dirname "$0")/.."
for so in tosem-2021-replication_b200/build_variants/lib_*.so; do
  TOSEMSCAN_LIB=$PWD/$so timeout 120 python tools/variant_bench.py 100000 0 2>&1 | tail -1
done
