#!/bin/bash
# round 2, call P: trainer with the merged candidate pass, shared-memory staging variants of the step kernel (A/B, two rounds)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2p
O=gpurun_out/r2p
timeout 600 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_comm.py -m gpu -x -q > $O/pytest_trainer.txt 2>&1; echo "trainer tests rc=$? $(tail -1 $O/pytest_trainer.txt)"
timeout 300 python tools/train_probe.py 4096 40 4 > $O/train_probe.json 2> $O/train_probe.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2p/train_probe.json").read().strip().splitlines()[-1])
print("trainer ms/iter", round(d["trainer_ms_per_iter"], 3), "launches/iter", d["trainer_kernel_launches_per_iter"], "loop", round(d["train_loop_env_steps_per_s"] / 1e6, 2), "M")
P
for v in product link_smem field_smem link_field_smem product link_smem field_smem link_field_smem; do
  if [ $v = product ]; then unset TRL_VARIANT; else export TRL_VARIANT=$v; fi
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/parity_$v.txt 2>&1; echo "$v parity: $(tail -1 $O/parity_$v.txt)"
  timeout 300 python bench.py --steps 30 --warmup 5 --cpu-seconds 0 --config4 0 > $O/bench_$v.json 2> $O/bench_$v.err
  python -c "
import json; d=json.loads(open('$O/bench_$v.json').read().strip().splitlines()[-1]); print('$v:', round(d['value']/1e6,2), 'M', round(d['ms_per_step'],3), 'ms/update  step launch', round(d['roofline']['launch_ms']*1e3,1), 'e2e', round(d['e2e']['value']/1e6,2))"
done
