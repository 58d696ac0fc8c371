#!/bin/bash
# round 2, call S: ground view in registers (product) -- parity, bench x2, fresh ncu capture of the step kernel for the next round of source-level analysis
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2s
O=gpurun_out/r2s
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scenarios.py tests/test_gpu_ref_golden.py -m gpu -x -q > $O/pytest.txt 2>&1; echo "parity tests rc=$? $(tail -1 $O/pytest.txt)"
for v in product product; do
  timeout 300 python bench.py --steps 30 --warmup 5 --cpu-seconds 0 --config4 0 > $O/bench_$v.json 2> $O/bench_$v.err
  python -c "
import json; d=json.loads(open('$O/bench_$v.json').read().strip().splitlines()[-1]); print('$v:', round(d['value']/1e6,2), 'M', round(d['ms_per_step'],3), 'ms/update  step launch', round(d['roofline']['launch_ms']*1e3,1), 'e2e', round(d['e2e']['value']/1e6,2))"
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:trl_step_kernel -s 2600 -c 2 -f -o $O/step python tools/profile_target.py > $O/ncu_step.log 2>&1; echo "ncu step rc=$?"
