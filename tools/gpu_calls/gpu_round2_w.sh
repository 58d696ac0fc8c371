#!/bin/bash
# round 2, call W: the shipped build -- full GPU test tier, bench (+ reference arm), side scenes, trainer probe, ncu evidence
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2w
O=gpurun_out/r2w
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest gpu rc=$? $(tail -1 $O/pytest_gpu.txt)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$? $(tail -2 $O/smoke.txt | tr '\n' ' ')"
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference arm rc=$?"; cut -c1-300 $O/bench_reference.json
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2w/bench.json").read().strip().splitlines()[-1])
c = d["config4"]
print("value", round(d["value"] / 1e6, 2), "M  e2e", round(d["e2e"]["value"] / 1e6, 2), " config4", round(c["value"] / 1e6, 2), "M", round(c["ms_per_step"], 3), "ms  rollout-only", round(c["rollout_only_ms_per_step"], 3),
      "cpu", d.get("cpu_baseline", {}).get("value"), "roofline", d["roofline"]["frac"], "fp64", d["roofline"]["fp64_pipe"]["frac"], "launch_ms", d["roofline"]["launch_ms"], "clocks", d["clocks"])
P
timeout 300 python tools/train_probe.py 4096 40 4 > $O/train_probe.json 2> $O/train_probe.err; python -c "
import json; d=json.loads(open('$O/train_probe.json').read().strip().splitlines()[-1]); print('trainer ms/iter', round(d['trainer_ms_per_iter'],3), 'launches', d['trainer_kernel_launches_per_iter'], 'loop', round(d['train_loop_env_steps_per_s']/1e6,2), 'M')"
timeout 300 python bench.py --scene raptor_narrow_gaps --envs 8192 --steps 20 --warmup 5 --cpu-seconds 0.5 > $O/bench_raptor_8192.json 2> $O/bench_raptor.err; python -c "
import json; d=json.loads(open('$O/bench_raptor_8192.json').read().strip().splitlines()[-1]); print('raptor 8192', round(d['value']/1e6,2), 'M e2e', round(d['e2e']['value']/1e6,2))"
timeout 300 python bench.py --scene goat_cliffs --envs 2048 --steps 20 --warmup 5 --cpu-seconds 0.5 > $O/bench_goat_2048.json 2> $O/bench_goat.err; python -c "
import json; d=json.loads(open('$O/bench_goat_2048.json').read().strip().splitlines()[-1]); print('goat 2048', round(d['value']/1e6,2), 'M e2e', round(d['e2e']['value']/1e6,2))"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --presim 0.5 --cpu-seconds 0 --config4 0 > $O/launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:trl_step_kernel -s 2600 -c 2 -f -o $O/step python tools/profile_target.py > $O/ncu_step.log 2>&1; echo "ncu step rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:trl_decide_fc_kernel -s 2000 -c 2 -f -o $O/fc python tools/profile_target.py > $O/ncu_fc.log 2>&1; echo "ncu fc rc=$?"
