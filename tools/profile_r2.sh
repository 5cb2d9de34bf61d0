#!/bin/bash
# Round-2 measurement pass (one gpurun call): gpurun --timeout 1700 -- 'bash tools/profile_r2.sh <tag>'
cd "$(dirname "$0")/.."
tag=${1:-r2}
o=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $o/${tag}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_scan|k_classify' -s 4 -c 2 -o $o/prof_${tag}_c2 python tools/variant_bench.py 100000 0 > $o/${tag}_ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_scan|k_classify' -s 4 -c 2 -o $o/prof_${tag}_c4 python tools/variant_bench.py 100000 1 > $o/${tag}_ncu_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_scan|k_diff_small|k_myers|k_gather|k_xscan_apply' -s 8 -c 8 -o $o/prof_${tag}_c5 python tools/diff_prof.py 50000 2 > $o/${tag}_ncu_c5.log 2>&1
{ for t in memcheck racecheck synccheck; do echo "== compute-sanitizer --tool $t python tools/sanitize_smoke.py"; timeout 600 compute-sanitizer --tool $t python tools/sanitize_smoke.py 2>&1 | tail -6; done; } > $o/${tag}_sanitizer.txt
python tools/fuzz_sweep.py 3000 60 > $o/${tag}_fuzz.txt 2>&1
python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err
python bench.py --config C4 --steps 10 --warmup 3 > $o/${tag}_bench_c4.json 2>> $o/${tag}_bench.err
python bench.py --config C5 --steps 10 --warmup 3 > $o/${tag}_bench_c5.json 2>> $o/${tag}_bench.err
python bench.py --config C3 --steps 5 --warmup 3 > $o/${tag}_bench_c3.json 2>> $o/${tag}_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > $o/${tag}_bench_reference.json 2>> $o/${tag}_bench.err
python bench.py --impl reference --config C5 --steps 3 --warmup 1 > $o/${tag}_bench_reference_c5.json 2>> $o/${tag}_bench.err
tail -3 $o/${tag}_fuzz.txt; cat $o/${tag}_sanitizer.txt | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|==" ; cut -c1-400 $o/${tag}_bench.json
