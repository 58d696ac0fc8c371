#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2j
O=gpurun_out/r2j
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest gpu rc=$? $(tail -1 $O/pytest_gpu.txt)"
timeout 300 python tools/train_probe.py 4096 40 4 > $O/train_probe_v2.json 2> $O/train_probe_v2.err; python -c "
import json; d=json.loads(open('$O/train_probe_v2.json').read().strip().splitlines()[-1]); print('batched fwd: trainer ms/iter', round(d['trainer_ms_per_iter'],3), 'loop', round(d['train_loop_env_steps_per_s']/1e6,2), 'M')"
TRL_TRAIN_FWD_V1=1 timeout 300 python tools/train_probe.py 4096 40 4 > $O/train_probe_v1.json 2> $O/train_probe_v1.err; python -c "
import json; d=json.loads(open('$O/train_probe_v1.json').read().strip().splitlines()[-1]); print('per-layer fwd: trainer ms/iter', round(d['trainer_ms_per_iter'],3), 'loop', round(d['train_loop_env_steps_per_s']/1e6,2), 'M')"
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --train-sync 1 --cpu-seconds 0.5 > $O/bench_sync.json 2> $O/bench_sync.err
python - <<'P'
import json
for f in ("bench", "bench_sync"):
    d = json.loads(open(f"gpurun_out/r2j/{f}.json").read().strip().splitlines()[-1])
    c = d["config4"]
    print(f, "value", round(d["value"] / 1e6, 2), "M  e2e", round(d["e2e"]["value"] / 1e6, 2), " config4", round(c["value"] / 1e6, 2), "M", round(c["ms_per_step"], 3), "ms  rollout-only", round(c["rollout_only_ms_per_step"], 3),
          "gather_ms", round(c["gather_ms"], 4), "dropped", c["tuples_dropped"], "iters", c["trainer_iter"], "cpu", d.get("cpu_baseline", {}).get("value"))
P
