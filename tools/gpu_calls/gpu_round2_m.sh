#!/bin/bash
# round 2, call M: tcgen05 split-precision probe, env-group A/B of the step launches, group parity test
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2m
O=gpurun_out/r2m
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 600 python tools/tc_policy_probe.py --decisions 1000000 > $O/tc_policy.json 2> $O/tc_policy.err; echo "tc probe rc=$?"; tail -3 $O/tc_policy.err; cut -c1-1500 $O/tc_policy.json
timeout 300 python -m pytest tests/test_gpu_scenarios.py -m gpu -x -q -k "groups or overlap" > $O/pytest_groups.txt 2>&1; echo "groups test: $(tail -1 $O/pytest_groups.txt)"
for G in 1 2 3 4 6 8 1 2; do
  export TRL_GROUPS=$G
  timeout 300 python bench.py --steps 30 --warmup 5 --cpu-seconds 0 --config4 0 > $O/bench_groups$G.json 2> $O/bench_groups$G.err
  python -c "
import json; d=json.loads(open('$O/bench_groups$G.json').read().strip().splitlines()[-1]); print('groups $G', round(d['value']/1e6,2), 'M', round(d['ms_per_step'],3), 'ms/update  span', round(d['config']['span_ms_incl_flush']/d['steps'],3), ' step launch', round(d['roofline']['launch_ms']*1e3,1), 'e2e', round(d['e2e']['value']/1e6,2), d['clocks']['sm_mhz'])"
done
unset TRL_GROUPS
