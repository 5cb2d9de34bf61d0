#!/bin/bash
# 8-GPU lines of the round: gpurun --gpus 8 --timeout 1200 -- 'bash tools/bench_n8.sh <tag>'
cd "$(dirname "$0")/.."
tag=${1:-r2}
o=gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 "${@:3}" > $o/${tag}_$2.json 2> $o/${tag}_$2.err; tail -c 300 $o/${tag}_$2.err | grep -v OMP_NUM | tail -3; }
nvidia-smi topo -m > $o/${tag}_topo.txt 2>&1
run 29601 bench_n8
run 29602 bench_c3_n8 --config C3 --steps 10 --warmup 3
run 29603 bench_c4_n8 --config C4 --steps 10 --warmup 3
run 29604 bench_c5_n8 --config C5 --steps 10 --warmup 3
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -4 > $o/${tag}_t_n8.log
NCCL_DEBUG=INFO timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29605 bench.py --gpus 8 --steps 3 --warmup 3 --scale 0.1 2>&1 | grep -E "NVLS|nranks|Connected all|Channel 00/" | head -8 > $o/${tag}_nccl_info.txt
cat $o/${tag}_t_n8.log; for f in bench_n8 bench_c3_n8 bench_c4_n8 bench_c5_n8; do cut -c1-260 $o/${tag}_$f.json; echo; done
