python tools/fuzz_sweep.py 2000 24 2>&1 | tail -3
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "newline_storms or pattern_ends or edge or chunk_edge" 2>&1 | tail -3
for law in 0 1; do
 TSM_SCAN_IMPL=1 python tools/variant_bench.py 100000 $law
 python tools/variant_bench.py 100000 $law
 TOSEMSCAN_LIB=$PWD/tosem-2021-replication_b200/build_variants/lib_w11c2rw2.so python tools/variant_bench.py 100000 $law
done
