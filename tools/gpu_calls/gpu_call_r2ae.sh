#!/bin/bash
# round 2, call AE: last check of the token-count assertions (tests/test_vitron_gpu.py + smoke)
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_vitron_gpu.py -q -x --timeout 100 -p no:cacheprovider > gpurun_out/t_vitron_last.log 2>&1
echo "vitron tests exit=$?"; tail -n 2 gpurun_out/t_vitron_last.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
