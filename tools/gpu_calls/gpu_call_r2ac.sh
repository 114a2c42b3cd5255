#!/bin/bash
# round 2, call AC: per-op profile of the batch-2 UNet forward (what the timed DDIM step runs)
mkdir -p gpurun_out
timeout 300 python tools/kineto_unet_ops.py b2 b2 > gpurun_out/kineto_ops_b2.log 2>&1
grep -v Warn gpurun_out/kineto_ops_b2.log | head -64
