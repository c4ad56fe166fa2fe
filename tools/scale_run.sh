#!/bin/bash
# Multi-GPU measurements of one box (run under `gpurun --gpus 8`): strong scaling of the headline bench at
# N = 8 / 4 (B = 1) and N = 8 with B = 2 (15 windows per rank), BASELINE config 4 (data generation, E = 1024)
# and config 5 (data-parallel training, global batch 256).  Outputs: gpurun_out/<tag>_*.json
tag=${1:-scale}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
run 8 29601 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/${tag}_bench_8gpu.json 2> gpurun_out/${tag}_bench_8gpu.err
tail -c 900 gpurun_out/${tag}_bench_8gpu.json; tail -3 gpurun_out/${tag}_bench_8gpu.err
run 4 29602 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/${tag}_bench_4gpu.json 2> gpurun_out/${tag}_bench_4gpu.err
head -c 300 gpurun_out/${tag}_bench_4gpu.json; echo
run 8 29603 bench.py --gpus 8 --steps 10 --warmup 3 --batch 2 > gpurun_out/${tag}_bench_8gpu_b2.json 2> gpurun_out/${tag}_bench_8gpu_b2.err
head -c 300 gpurun_out/${tag}_bench_8gpu_b2.json; echo
run 8 29604 tools/generate_kolmogorov.py --out /tmp/kolmo_cfg4 --members 1024 > gpurun_out/${tag}_generate_8gpu.log 2>&1
tail -2 gpurun_out/${tag}_generate_8gpu.log
run 8 29605 tools/train_bench.py 10 > gpurun_out/${tag}_train_8gpu.json 2> gpurun_out/${tag}_train_8gpu.err
tail -1 gpurun_out/${tag}_train_8gpu.json
SDAB_MODE=bf16 run 8 29606 tools/train_bench.py 10 > gpurun_out/${tag}_train_8gpu_bf16.json 2> gpurun_out/${tag}_train_8gpu_bf16.err
tail -1 gpurun_out/${tag}_train_8gpu_bf16.json
SDAB_TRAIN_DDP=1 run 8 29607 tools/train_bench.py 10 > gpurun_out/${tag}_train_8gpu_ddp.json 2> gpurun_out/${tag}_train_8gpu_ddp.err
tail -1 gpurun_out/${tag}_train_8gpu_ddp.json
