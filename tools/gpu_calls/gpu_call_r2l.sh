#!/bin/bash
# round 2, call L (2 GPUs): bench.py under torchrun at N=2 (LLM weak scaling, UNet replicas + CFG split over NCCL, video configs[2] 64/N), reference arm
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err
echo "bench n2 exit=$?" | tee gpurun_out/summary_r2l.txt
tail -c 6000 gpurun_out/bench_r02_n2.json; tail -5 gpurun_out/bench_r02_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_r02_ref_n2.json 2> gpurun_out/bench_r02_ref_n2.err
echo "reference arm n2 exit=$?" | tee -a gpurun_out/summary_r2l.txt
tail -c 1500 gpurun_out/bench_r02_ref_n2.json
