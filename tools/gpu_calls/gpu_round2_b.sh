#!/bin/bash
# round-2 GPU call B: the batched decision path (TMA + DMMA) on hardware -- smoke under a timeout first, then tests, timelines, A/B
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2b
O=gpurun_out/r2b
(cd tools/microbench && ./fp64_pipes) > $O/fp64_pipes.txt 2>&1; tail -6 $O/fp64_pipes.txt
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_v2.txt 2>&1; echo "smoke v2 rc=$? $(tail -2 $O/smoke_v2.txt | head -1)"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/memcheck_v2.txt 2>&1; echo "memcheck v2 rc=$? $(grep -c 'Invalid\|Misaligned' $O/memcheck_v2.txt) errors"
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest gpu rc=$? $(tail -1 $O/pytest_gpu.txt)"
timeout 300 python tools/timeline_probe.py 4096 $O/timeline_v2.json > $O/timeline_v2.txt 2>&1; head -16 $O/timeline_v2.txt
TRL_DECIDE_V1=1 timeout 300 python tools/timeline_probe.py 4096 $O/timeline_v1.json > $O/timeline_v1.txt 2>&1; head -16 $O/timeline_v1.txt
timeout 300 python bench.py --steps 30 --warmup 5 --cpu-seconds 0.5 --config4 0 > $O/bench_v2.json 2> $O/bench_v2.err; echo "bench v2 rc=$?"
TRL_DECIDE_V1=1 timeout 300 python bench.py --steps 30 --warmup 5 --cpu-seconds 0.5 --config4 0 > $O/bench_v1.json 2> $O/bench_v1.err; echo "bench v1 rc=$?"
python - <<'P'
import json
for v in ("v2", "v1"):
    try:
        d = json.loads(open(f"gpurun_out/r2b/bench_{v}.json").read().strip().splitlines()[-1])
        print(v, f"{d['value']/1e6:.2f} M env-steps/s  {d['ms_per_step']:.3f} ms/update  step launch {d['roofline']['launch_ms']*1e3:.1f} us  share {d['roofline']['step_kernel_share_of_update']:.3f}  e2e {d['e2e']['value']/1e6:.2f} M")
    except Exception as e:
        print(v, "no line", e)
P
