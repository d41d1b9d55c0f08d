#!/bin/bash
tag=s5x
mkdir -p gpurun_out
nvidia-smi -L
timeout 1700 python -m pytest tests -m gpu -x -q -s > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/${tag}_bench_w32_2gpu.json 2> gpurun_out/${tag}_bench_w32_2gpu.err; cat gpurun_out/${tag}_bench_w32_2gpu.json | head -c 1500; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/${tag}_ref_w32_2gpu.json 2> gpurun_out/${tag}_ref_w32_2gpu.err; cat gpurun_out/${tag}_ref_w32_2gpu.json | head -c 700; echo
