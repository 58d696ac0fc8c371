#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2h
O=gpurun_out/r2h
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest gpu rc=$? $(tail -1 $O/pytest_gpu.txt)"
TRL_PDL=1 TRL_LAG=6 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scenarios.py tests/test_gpu_ref_golden.py -m gpu -x -q > $O/pytest_pdl.txt 2>&1; echo "pytest PDL rc=$? $(tail -1 $O/pytest_pdl.txt)"
for cfg in "4 0" "6 0" "7 0" "4 1" "6 1" "7 1"; do
  set -- $cfg
  export TRL_LAG=$1 TRL_PDL=$2
  timeout 300 python bench.py --steps 30 --warmup 5 --cpu-seconds 0.5 --config4 0 > $O/bench_lag$1_pdl$2.json 2> $O/bench_lag$1_pdl$2.err
  python -c "
import json; d=json.loads(open('$O/bench_lag$1_pdl$2.json').read().strip().splitlines()[-1]); print('lag$1 pdl$2', round(d['value']/1e6,2), 'M', round(d['ms_per_step'],3), 'ms/update, e2e', round(d['e2e']['value']/1e6,2))"
done
TRL_LAG=6 TRL_PDL=1 timeout 300 python tools/timeline_probe.py 4096 $O/timeline_lag6_pdl.json > $O/timeline_lag6_pdl.txt 2>&1; head -3 $O/timeline_lag6_pdl.txt
