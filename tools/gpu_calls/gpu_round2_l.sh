#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2l
O=gpurun_out/r2l
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest gpu rc=$? $(tail -1 $O/pytest_gpu.txt)"
timeout 300 python tools/fc_phase_probe.py > $O/fc_phases.txt 2>&1; tail -4 $O/fc_phases.txt
timeout 300 python tools/train_probe.py 4096 40 4 > $O/train_probe.json 2> $O/train_probe.err; python -c "
import json; d=json.loads(open('$O/train_probe.json').read().strip().splitlines()[-1]); print('trainer ms/iter', round(d['trainer_ms_per_iter'],3), 'loop', round(d['train_loop_env_steps_per_s']/1e6,2), 'M')"
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference arm rc=$?"; cut -c1-400 $O/bench_reference.json
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2l/bench.json").read().strip().splitlines()[-1])
c = d["config4"]
print("value", round(d["value"] / 1e6, 2), "M  e2e", round(d["e2e"]["value"] / 1e6, 2), " config4", round(c["value"] / 1e6, 2), "M", round(c["ms_per_step"], 3), "ms  rollout-only", round(c["rollout_only_ms_per_step"], 3),
      "cpu", d.get("cpu_baseline", {}).get("value"), "fp64", d["roofline"]["fp64_pipe"]["frac"], "clocks", d["clocks"])
P
timeout 300 python bench.py --scene raptor_narrow_gaps --envs 8192 --steps 20 --warmup 5 --cpu-seconds 0.5 > $O/bench_raptor_8192.json 2> $O/bench_raptor.err; python -c "
import json; d=json.loads(open('$O/bench_raptor_8192.json').read().strip().splitlines()[-1]); print('raptor 8192', round(d['value']/1e6,2), 'M e2e', round(d['e2e']['value']/1e6,2))"
timeout 300 python bench.py --scene goat_cliffs --envs 2048 --steps 20 --warmup 5 --cpu-seconds 0.5 > $O/bench_goat_2048.json 2> $O/bench_goat.err; python -c "
import json; d=json.loads(open('$O/bench_goat_2048.json').read().strip().splitlines()[-1]); print('goat 2048', round(d['value']/1e6,2), 'M e2e', round(d['e2e']['value']/1e6,2))"
