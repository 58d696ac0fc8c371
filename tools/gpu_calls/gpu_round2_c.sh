#!/bin/bash
# round-2 GPU call C: where does the batched decision path spend its time?  timelines + ncu of its two kernels
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
timeout 300 python tools/timeline_probe.py 4096 $O/timeline_v2.json > $O/timeline_v2.txt 2>&1; head -24 $O/timeline_v2.txt
TRL_DECIDE_V1=1 timeout 300 python tools/timeline_probe.py 4096 $O/timeline_v1.json > $O/timeline_v1.txt 2>&1; head -24 $O/timeline_v1.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:trl_decide_fc_kernel -s 2000 -c 2 -f -o $O/fc python tools/profile_target.py > $O/ncu_fc.log 2>&1; echo "ncu fc rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:trl_decide_conv_kernel -s 2000 -c 2 -f -o $O/conv python tools/profile_target.py > $O/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest gpu rc=$? $(tail -1 $O/pytest_gpu.txt)"
ls -la $O
