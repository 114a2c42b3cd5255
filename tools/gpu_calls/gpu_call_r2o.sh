#!/bin/bash
# round 2, call O (4 GPUs): bench.py under torchrun at N=4 (two CFG pairs, video 64/4)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_r02_n4.json 2> gpurun_out/bench_r02_n4.err
echo "bench n4 exit=$?" | tee gpurun_out/summary_r2o.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_n4.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['n_gpus'], d['video_cfg2']['value'], d['unet']['value'], d['unet']['cfg_split'])
PY
tail -3 gpurun_out/bench_r02_n4.err
