"""Data-parallel training over N GPUs (BASELINE configs[3]): every rank rolls out its own env shard, the finished tuples of all
ranks are exchanged with ONE NCCL all-gather per outer update, and every rank feeds the same gathered block to its own
on-device trainer.  The trainers are deterministic, so the replicas stay bit-identical without a weight broadcast (checked
at the end with an all-reduce of the parameter vector).
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/train_multi.py [envs] [updates] [iters]"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import deepterrainrl_b200 as trl  # noqa: E402
from deepterrainrl_b200 import parallel  # noqa: E402
from deepterrainrl_b200.train import TrainSchedule  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
updates = int(sys.argv[2]) if len(sys.argv) > 2 else 60
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4
pack = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "assets", "dog_slopes_mixed.trlpack")
sc = trl.ScenarioExpMACE(pack, n, device=local, terrain_seeds=parallel.shard_seeds(rank, n), rng_seed=100 + rank)
tr = trl.MACETrainer(sc, replay_mem_size=200000, num_init_samples=4000 * world, freeze_target_iters=50, seed=9)
sched = TrainSchedule(init_exp_rate=0.9, exp_rate=0.2, init_exp_temp=20, exp_temp=0.025, init_exp_base_rate=0.9, exp_base_rate=0.002,
                      trainer_num_anneal_iters=2000, exp_base_anneal_iters=2000)
rows, flags, env, count = sc.DeviceTupleBlock()


def one_update(k):
    s = sched(k * iters)
    sc.EnableExplore(True, s["exp_rate"], s["exp_temp"], s["exp_base_rate"])
    sc.Update(1.0 / 30.0)
    sc.Sync()
    g = parallel.gather_tuple_blocks_fixed(rows, flags, env, count, env_offset=rank * n, block_rows=1024)
    sc.ResetTupleBuffer()
    r, f, _ = parallel.unpack_tuple_blocks(g)
    tr.AddTuplesDevice(r, f)
    tr.Train(iters)


for k in range(60):
    one_update(k)
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(60, 60 + updates):
    one_update(k)
sc.Sync(); dist.barrier(); torch.cuda.synchronize()
dt = time.perf_counter() - t0
theta = torch.from_numpy(tr.get("theta")).cuda()
hi, lo = theta.clone(), theta.clone()
dist.all_reduce(hi, op=dist.ReduceOp.MAX); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
c = tr.counters()
if rank == 0:
    print(json.dumps({"n_gpus": world, "envs_per_gpu": n, "updates": updates, "iters_per_update": iters,
                      "train_loop_env_steps_per_s": world * n * 20 * updates / dt, "ms_per_update": 1e3 * dt / updates,
                      "replica_max_abs_diff": float((hi - lo).abs().max().item()), "trainer_iter": c["iter"], "actor_iter": c["actor_iter"],
                      "replay_tuples": c["num"], "critic_loss": c["critic_loss"]}))
dist.destroy_process_group()
