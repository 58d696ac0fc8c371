#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2i
O=gpurun_out/r2i
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest gpu rc=$? $(tail -1 $O/pytest_gpu.txt)"
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_async.json 2> $O/bench_async.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --train-sync 1 --cpu-seconds 0.5 > $O/bench_sync.json 2> $O/bench_sync.err
python - <<'P'
import json
for f in ("async", "sync"):
    d = json.loads(open(f"gpurun_out/r2i/bench_{f}.json").read().strip().splitlines()[-1])
    c = d["config4"]
    print(f, "value", round(d["value"] / 1e6, 2), "M  e2e", round(d["e2e"]["value"] / 1e6, 2), " config4", round(c["value"] / 1e6, 2), "M", round(c["ms_per_step"], 3), "ms  rollout-only", round(c["rollout_only_ms_per_step"], 3),
          "gather_ms", round(c["gather_ms"], 4), "dropped", c["tuples_dropped"], "iters", c["trainer_iter"], "cpu", d.get("cpu_baseline", {}).get("value"))
P
for cfg in "product 7" "ldlt_smem 6" "smem_xchg 6" "accum_ldlt 6" "xchg_no_contact 6"; do
  set -- $cfg
  if [ $1 = product ]; then unset TRL_VARIANT; else export TRL_VARIANT=$1; fi
  export TRL_LAG=$2
  timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ref_golden.py -m gpu -x -q > $O/parity_$1.txt 2>&1; echo "$1 parity: $(tail -1 $O/parity_$1.txt)"
  timeout 300 python bench.py --steps 30 --warmup 5 --cpu-seconds 0.5 --config4 0 > $O/bench_$1_lag$2.json 2> $O/bench_$1_lag$2.err
  python -c "
import json; d=json.loads(open('$O/bench_$1_lag$2.json').read().strip().splitlines()[-1]); print('$1 lag$2', round(d['value']/1e6,2), 'M', round(d['ms_per_step'],3), 'ms/update, step launch', round(d['roofline']['launch_ms']*1e3,1), 'e2e', round(d['e2e']['value']/1e6,2))"
done
