#!/bin/bash
# round 2, call AA: tuple intake with a wrapped replay ring (the steady state of a long run), 1 GPU; trainer / comm tests
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2aa
O=gpurun_out/r2aa
timeout 600 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_comm.py -m gpu -x -q > $O/pytest.txt 2>&1; echo "tests rc=$? $(tail -1 $O/pytest.txt)"
for rs in 200000 20000; do
  timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 0.5 --replay-size $rs > $O/bench_replay$rs.json 2> $O/bench_replay$rs.err
  python -c "
import json; d=json.loads(open('$O/bench_replay$rs.json').read().strip().splitlines()[-1]); c=d['config4']; print('replay $rs: value', round(d['value']/1e6,2), 'config4', round(c['value']/1e6,2), 'M', round(c['ms_per_step'],3), 'ms  replay', c['replay_tuples'], 'iters', c['trainer_iter'])"
done
