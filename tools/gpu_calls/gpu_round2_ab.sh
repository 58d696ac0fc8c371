#!/bin/bash
# round 2, call AB: shipped build after the trainer's merged target pass -- full GPU tier, smoke, trainer probe, bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2ab
O=gpurun_out/r2ab
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest gpu rc=$? $(tail -1 $O/pytest_gpu.txt)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$? $(tail -2 $O/smoke.txt | tr '\n' ' ')"
timeout 300 python tools/train_probe.py 4096 40 4 > $O/train_probe.json 2> $O/train_probe.err; python -c "
import json; d=json.loads(open('$O/train_probe.json').read().strip().splitlines()[-1]); print('trainer ms/iter', round(d['trainer_ms_per_iter'],3), 'launches', d['trainer_kernel_launches_per_iter'], 'loop', round(d['train_loop_env_steps_per_s']/1e6,2), 'M')"
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2ab/bench.json").read().strip().splitlines()[-1])
c = d["config4"]
print("value", round(d["value"] / 1e6, 2), "M  e2e", round(d["e2e"]["value"] / 1e6, 2), " config4", round(c["value"] / 1e6, 2), "M", round(c["ms_per_step"], 3), "ms  rollout-only", round(c["rollout_only_ms_per_step"], 3),
      "cpu", d.get("cpu_baseline", {}).get("value"), "fp64", d["roofline"]["fp64_pipe"]["frac"], "launch_ms", d["roofline"]["launch_ms"], "clocks", d["clocks"])
P
