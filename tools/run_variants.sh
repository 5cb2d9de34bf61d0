#!/bin/bash
# Time every build variant under tosem-2021-replication_b200/build_variants/ (one gpurun call):
#   gpurun -- 'bash tools/run_variants.sh [law] > gpurun_out/variants.txt'     law 0 = C2 (default), 1 = C4
cd "$(dirname "$0")/.."
for law in ${@:-0}; do
for so in tosem-2021-replication_b200/build_variants/lib_*.so; do
  TOSEMSCAN_LIB=$PWD/$so timeout 120 python tools/variant_bench.py 100000 $law 2>&1 | tail -1
done
done
