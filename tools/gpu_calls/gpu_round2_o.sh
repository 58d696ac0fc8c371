#!/bin/bash
# round 2, call O: full GPU test tier, trainer A/B (fused backward), bench with config 4, link_smem A/B, ncu evidence of the shipped build
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2o
O=gpurun_out/r2o
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest gpu rc=$? $(tail -1 $O/pytest_gpu.txt)"
timeout 300 python tools/train_probe.py 4096 40 4 > $O/train_probe_fused.json 2> $O/train_probe.err
TRL_TRAIN_BWD_V1=1 timeout 300 python tools/train_probe.py 4096 40 4 > $O/train_probe_unfused.json 2>> $O/train_probe.err
python - <<'P'
import json
for f in ("fused", "unfused"):
    d = json.loads(open(f"gpurun_out/r2o/train_probe_{f}.json").read().strip().splitlines()[-1])
    print(f, "trainer ms/iter", round(d["trainer_ms_per_iter"], 3), "launches/iter", d["trainer_kernel_launches_per_iter"], "loop", round(d["train_loop_env_steps_per_s"] / 1e6, 2), "M")
P
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2o/bench.json").read().strip().splitlines()[-1])
c = d["config4"]
print("value", round(d["value"] / 1e6, 2), "M  e2e", round(d["e2e"]["value"] / 1e6, 2), " config4", round(c["value"] / 1e6, 2), "M", round(c["ms_per_step"], 3), "ms  rollout-only", round(c["rollout_only_ms_per_step"], 3),
      "cpu", d.get("cpu_baseline", {}).get("value"), "fp64", d["roofline"]["fp64_pipe"]["frac"], "clocks", d["clocks"])
P
for v in product link_smem product link_smem; do
  if [ $v = product ]; then unset TRL_VARIANT; else export TRL_VARIANT=$v; fi
  timeout 300 python bench.py --steps 30 --warmup 5 --cpu-seconds 0 --config4 0 > $O/bench_ab_$v.json 2> $O/bench_ab_$v.err
  python -c "
import json; d=json.loads(open('$O/bench_ab_$v.json').read().strip().splitlines()[-1]); print('$v:', round(d['value']/1e6,2), 'M', round(d['ms_per_step'],3), 'ms/update  step launch', round(d['roofline']['launch_ms']*1e3,1), 'e2e', round(d['e2e']['value']/1e6,2))"
done
unset TRL_VARIANT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1800 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --presim 0.5 --cpu-seconds 0 --config4 0 > $O/launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:trl_step_kernel -s 2600 -c 2 -f -o $O/step python tools/profile_target.py > $O/ncu_step.log 2>&1; echo "ncu step rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:trl_tc_fc_kernel -c 4 -f -o $O/tc python tools/tc_profile_target.py > $O/ncu_tc.log 2>&1; echo "ncu tc rc=$?"; tail -2 $O/ncu_tc.log
ls -la $O/*.ncu-rep
