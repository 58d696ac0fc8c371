#!/bin/bash
# round-2 GPU call A: full GPU test tier on the product build, the variant A/B, then the product bench line incl. config 4
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2a/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a/pytest_gpu.txt 2>&1; echo "pytest gpu rc=$? $(tail -1 gpurun_out/r2a/pytest_gpu.txt)"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a/bench_product.json 2> gpurun_out/r2a/bench_product.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2a/bench_product.json
bash tools/measure_variants.sh
