#!/bin/bash
# round-2 GPU call F (--gpus 2): the native exchange over NCCL between two ranks
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
N=${1:-2}
nvidia-smi -L > $O/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/train_multi.py 4096 60 4 > $O/train_multi_repl.json 2> $O/train_multi_repl.err; echo "train_multi replicated rc=$?"; tail -1 $O/train_multi_repl.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/train_multi.py 4096 60 4 --root-trainer > $O/train_multi_root.json 2> $O/train_multi_root.err; echo "train_multi root rc=$?"; tail -1 $O/train_multi_root.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err; echo "bench rc=$?"
python - $O/bench_${N}gpu.json <<'P'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", d["value"] / 1e6, "M  n_gpus", d["n_gpus"])
print("config4", json.dumps(d.get("config4"))[:900])
P
tail -3 $O/*.err | head -40
