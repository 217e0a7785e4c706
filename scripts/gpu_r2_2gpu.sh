#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --quick > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "rc=$?"; cat gpurun_out/bench_2gpu.json | cut -c1-1500; tail -3 gpurun_out/bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 --cpu-sample 16 > gpurun_out/bench_2gpu_ref.json 2>> gpurun_out/bench_2gpu.err; echo "ref rc=$?"; cat gpurun_out/bench_2gpu_ref.json | cut -c1-600
