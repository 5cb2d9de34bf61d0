#!/bin/bash
# Short 8-GPU pass (C2 weak + C5 strong + the multi-GPU tests): gpurun --gpus 8 --timeout 900 -- 'bash tools/bench_n8_short.sh <tag>'
cd "$(dirname "$0")/.."
tag=${1:-r2}
o=gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 "${@:3}" > $o/${tag}_$2.json 2> $o/${tag}_$2.err; }
run 29611 bench_n8
run 29612 bench_c5_n8 --config C5 --steps 10 --warmup 3
timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_cli.py tests/test_history.py -m gpu -x -q 2>&1 | tail -4 > $o/${tag}_t_n8.log
cat $o/${tag}_t_n8.log; for f in bench_n8 bench_c5_n8; do cut -c1-260 $o/${tag}_$f.json; echo; done
