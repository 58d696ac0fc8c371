#!/bin/bash
# round 2, call R: lane-indexed model tables from a global-memory mirror, with 4 and 8 warps per CTA
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2r
O=gpurun_out/r2r
for v in product table_mirror cta8 cta8_mirror product table_mirror cta8 cta8_mirror; do
  if [ $v = product ]; then unset TRL_VARIANT; else export TRL_VARIANT=$v; fi
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/parity_$v.txt 2>&1; echo "$v parity: $(tail -1 $O/parity_$v.txt)"
  timeout 300 python bench.py --steps 30 --warmup 5 --cpu-seconds 0 --config4 0 > $O/bench_$v.json 2> $O/bench_$v.err
  python -c "
import json; d=json.loads(open('$O/bench_$v.json').read().strip().splitlines()[-1]); print('$v:', round(d['value']/1e6,2), 'M', round(d['ms_per_step'],3), 'ms/update  step launch', round(d['roofline']['launch_ms']*1e3,1), 'e2e', round(d['e2e']['value']/1e6,2))"
done
