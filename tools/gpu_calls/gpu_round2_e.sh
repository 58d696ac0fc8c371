#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
(cd tools/microbench && timeout 120 ./tma_stream) > $O/tma_stream.txt 2>&1; cat $O/tma_stream.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest gpu rc=$? $(tail -1 $O/pytest_gpu.txt)"
timeout 300 python tools/timeline_probe.py 4096 $O/timeline_v2.json > $O/timeline_v2.txt 2>&1; head -34 $O/timeline_v2.txt
TRL_DECIDE_V1=1 timeout 300 python tools/timeline_probe.py 4096 $O/timeline_v1.json > $O/timeline_v1.txt 2>&1; head -3 $O/timeline_v1.txt
for v in v2 v1; do
  if [ $v = v1 ]; then export TRL_DECIDE_V1=1; else unset TRL_DECIDE_V1; fi
  timeout 300 python bench.py --steps 30 --warmup 5 --cpu-seconds 0.5 --config4 0 > $O/bench_$v.json 2> $O/bench_$v.err
  python -c "
import json; d=json.loads(open('$O/bench_$v.json').read().strip().splitlines()[-1]); print('$v lag2', round(d['value']/1e6,2), 'M', round(d['ms_per_step'],3), 'ms/update, step launch', round(d['roofline']['launch_ms']*1e3,1), 'us, e2e', round(d['e2e']['value']/1e6,2))"
done
unset TRL_DECIDE_V1
TRL_SERIAL_SCHEDULE=1 timeout 300 python bench.py --steps 30 --warmup 5 --cpu-seconds 0.5 --config4 0 > $O/bench_serial.json 2> $O/bench_serial.err
python -c "
import json; d=json.loads(open('$O/bench_serial.json').read().strip().splitlines()[-1]); print('serial', round(d['value']/1e6,2), 'M', round(d['ms_per_step'],3))"
