#!/bin/bash
# quick GPU round trip: parity tests + one short bench line (key fields only)
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 30 --warmup 5 --cpu-seconds 1 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
tail -2 gpurun_out/bench_quick.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
print("value %.4g  ms/step %.3f  e2e %.4g  step-launch %.1f us  share %.3f" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["launch_ms"] * 1e3, d["roofline"]["step_kernel_share_of_update"]))
P
