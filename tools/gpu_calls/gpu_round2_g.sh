#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2g
O=gpurun_out/r2g
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest gpu rc=$? $(tail -1 $O/pytest_gpu.txt)"
timeout 300 python tools/timeline_probe.py 4096 $O/timeline_v2_lag4.json > $O/timeline_v2_lag4.txt 2>&1; head -34 $O/timeline_v2_lag4.txt
for cfg in "v2 1" "v2 2" "v2 3" "v2 4" "v2 5" "v2 6" "v1 1" "v1 4"; do
  set -- $cfg
  if [ $1 = v1 ]; then export TRL_DECIDE_V1=1; else unset TRL_DECIDE_V1; fi
  export TRL_LAG=$2
  timeout 300 python bench.py --steps 30 --warmup 5 --cpu-seconds 0.5 --config4 0 > $O/bench_$1_lag$2.json 2> $O/bench_$1_lag$2.err
  python -c "
import json; d=json.loads(open('$O/bench_$1_lag$2.json').read().strip().splitlines()[-1]); print('$1 lag$2', round(d['value']/1e6,2), 'M', round(d['ms_per_step'],3), 'ms/update, e2e', round(d['e2e']['value']/1e6,2))"
done
