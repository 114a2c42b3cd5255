#!/bin/bash
# round 2, call AA: decode attention kernel: graph-replay timing at the bench shape + ncu --set full of two steady-state launches
mkdir -p gpurun_out
timeout 200 python tools/profile_decode_attn.py > gpurun_out/decode_attn_r02.json 2> gpurun_out/decode_attn_r02.err
cat gpurun_out/decode_attn_r02.json; tail -2 gpurun_out/decode_attn_r02.err
VB200_DEC_SPLIT_KEYS=256 timeout 200 python tools/profile_decode_attn.py > gpurun_out/decode_attn_r02_keys256.json 2>> gpurun_out/decode_attn_r02.err
cat gpurun_out/decode_attn_r02_keys256.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_decode -s 4 -c 2 -f -o gpurun_out/ncu_attn_decode_r02 python tools/profile_decode_attn.py > gpurun_out/ncu_attn_decode_r02.log 2>&1
echo "ncu exit=$?"
