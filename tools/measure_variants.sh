#!/bin/bash
# Round-2 first GPU call: A/B the experimental env-step kernel builds (deepterrainrl_b200/scenario.py: VARIANTS) against the
# product build on one B200.  For each build: the GPU parity tests of the step / decision path, then one bench line.
#
#   gpurun --timeout 1500 -- 'bash tools/measure_variants.sh'            # all variants
#   gpurun --timeout 600  -- 'bash tools/measure_variants.sh smem_xchg'  # selected ones
#
# The variant libraries are prebuilt here by `TRL_VARIANT=<name> python -c "import deepterrainrl_b200 as t; t.build_library()"`;
# lib/variants/ is listed in .gpurunignore (8 x 12 MB) -- take that line out for the measurement call, or let this script
# compile the missing ones on the box (nvcc is in the image; ~2 min of box time for all of them in parallel).
# Output: gpurun_out/variants/<name>.json (bench line), <name>.parity.txt, summary.txt
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/variants
names=("$@")
if [ ${#names[@]} -eq 0 ]; then
    names=(product $(python -c "from deepterrainrl_b200 import scenario; print(' '.join(scenario.VARIANTS))"))
fi
for v in "${names[@]}"; do
    [ "$v" = product ] && continue
    ( TRL_VARIANT=$v python -c "import deepterrainrl_b200 as t; t.build_library()" > gpurun_out/variants/$v.build.txt 2>&1 ) &
done
wait
# the experiment builds have never run on a GPU: one tiny update under compute-sanitizer first (out-of-range / misaligned shared
# or global accesses would otherwise surface as a dead context in the middle of the A/B)
for v in "${names[@]}"; do
    case "$v" in smem_xchg|decide_tile4|smem_xchg_decide_tile4|all)
        TRL_VARIANT=$v timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" \
            > gpurun_out/variants/$v.memcheck.txt 2>&1
        echo "$v memcheck: exit $? $(grep -c 'Invalid\|Misaligned' gpurun_out/variants/$v.memcheck.txt) errors" ;;
    esac
done
for v in "${names[@]}"; do
    if [ "$v" = product ]; then export -n TRL_VARIANT; unset TRL_VARIANT; else export TRL_VARIANT=$v; fi
    timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ref_golden.py -m gpu -x -q > gpurun_out/variants/$v.parity.txt 2>&1
    echo "$v parity: $(tail -1 gpurun_out/variants/$v.parity.txt)"
    timeout 300 python bench.py --steps 30 --warmup 5 --cpu-seconds 0.5 --config4 0 > gpurun_out/variants/$v.json 2> gpurun_out/variants/$v.err
done
unset TRL_VARIANT
python - "${names[@]}" <<'P' | tee gpurun_out/variants/summary.txt
import json, sys
base = None
for v in sys.argv[1:]:
    try:
        d = json.loads(open(f"gpurun_out/variants/{v}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(f"{v:16s} no bench line ({e})"); continue
    if v == "product": base = d["value"]
    rel = f"{d['value'] / base:6.3f}x" if base else "      "
    print(f"{v:16s} {d['value'] / 1e6:7.2f} M env-steps/s {rel}  {d['ms_per_step']:.3f} ms/update  step launch {d['roofline']['launch_ms'] * 1e3:6.1f} us"
          f"  e2e {d['e2e']['value'] / 1e6:6.2f} M  clocks {d['clocks']['sm_mhz']:.0f} MHz {d['clocks']['reasons']}")
P
