#!/bin/bash
# round 2, call N: tcgen05 split-precision probe; register-budget variants x env groups
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2n
O=gpurun_out/r2n
timeout 700 python tools/tc_policy_probe.py --decisions 1000000 > $O/tc_policy.json 2> $O/tc_policy.err; echo "tc probe rc=$?"; tail -3 $O/tc_policy.err; cut -c1-3000 $O/tc_policy.json
for cfg in "product 2 6" "regs96 1 6" "regs96 2 6" "regs80 1 6" "regs80 2 6" "product 2 4" "product 2 7"; do
  set -- $cfg
  if [ $1 = product ]; then unset TRL_VARIANT; else export TRL_VARIANT=$1; fi
  export TRL_GROUPS=$2 TRL_LAG=$3
  timeout 300 python bench.py --steps 30 --warmup 5 --cpu-seconds 0 --config4 0 > $O/bench_$1_g$2_lag$3.json 2> $O/bench_$1_g$2_lag$3.err
  python -c "
import json; d=json.loads(open('$O/bench_$1_g$2_lag$3.json').read().strip().splitlines()[-1]); print('$1 groups $2 lag $3:', round(d['value']/1e6,2), 'M', round(d['ms_per_step'],3), 'ms/update  step launch', round(d['roofline']['launch_ms']*1e3,1), 'e2e', round(d['e2e']['value']/1e6,2), d['clocks']['sm_mhz'])"
done
