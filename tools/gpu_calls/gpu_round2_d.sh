#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2d
O=gpurun_out/r2d
timeout 300 python tools/fc_phase_probe.py > $O/fc_phases.txt 2>&1; cat $O/fc_phases.txt | tail -6
timeout 300 python tools/timeline_probe.py 4096 $O/timeline_v2.json > $O/timeline_v2.txt 2>&1; head -30 $O/timeline_v2.txt
TRL_DECIDE_V1=1 timeout 300 python tools/timeline_probe.py 4096 $O/timeline_v1.json > $O/timeline_v1.txt 2>&1; head -30 $O/timeline_v1.txt
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_net_torch.py -m gpu -x -q > $O/pytest.txt 2>&1; tail -1 $O/pytest.txt
timeout 300 python bench.py --steps 30 --warmup 5 --cpu-seconds 0.5 --config4 0 > $O/bench_v2.json 2> $O/bench_v2.err; python -c "
import json; d=json.loads(open('$O/bench_v2.json').read().strip().splitlines()[-1]); print('v2 dmma-conv', d['value']/1e6, d['ms_per_step'], d['roofline']['step_kernel_share_of_update'])"
