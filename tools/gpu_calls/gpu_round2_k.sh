#!/bin/bash
# round-2 GPU call K: ncu evidence for the shipped build -- launch list of the bench command, full reports of the step / conv / FC kernels
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2k
O=gpurun_out/r2k
# launch list of the same command the driver runs (graph launches expand to kernel launches under ncu)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --presim 0.5 --cpu-seconds 0 --config4 0 > $O/launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:trl_step_kernel -s 2600 -c 2 -f -o $O/step python tools/profile_target.py > $O/ncu_step.log 2>&1; echo "ncu step rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:trl_decide_conv_kernel -s 2000 -c 2 -f -o $O/conv python tools/profile_target.py > $O/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:trl_decide_fc_kernel -s 2000 -c 2 -f -o $O/fc python tools/profile_target.py > $O/ncu_fc.log 2>&1; echo "ncu fc rc=$?"
ls -la $O
