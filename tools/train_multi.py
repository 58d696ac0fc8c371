"""Data-parallel training over N GPUs (BASELINE configs[3]) through the C ABI only: every rank rolls out its own env shard, the
finished tuples of all ranks are exchanged with ONE all-gather per outer update (trl_gather_tuples: device pack kernel + ncclAllGather on
a side stream, no host sync), every rank feeds the same gathered blocks to its own on-device trainer.  The trainers are deterministic,
so the replicas stay bit-identical without a weight broadcast (checked at the end: trl_trainer_replica_spread); `--root-trainer`
measures the alternative: rank 0 trains alone and broadcasts the net after every update (trl_trainer_broadcast = SyncNet).
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/train_multi.py [envs] [updates] [iters] [--root-trainer]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import deepterrainrl_b200 as trl  # noqa: E402
from deepterrainrl_b200 import parallel  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pos = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(pos[0]) if len(pos) > 0 else 4096
updates = int(pos[1]) if len(pos) > 1 else 60
iters = int(pos[2]) if len(pos) > 2 else 4
root_trainer = "--root-trainer" in sys.argv
pack = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "assets", "dog_slopes_mixed.trlpack")
sc = trl.ScenarioExpMACE(pack, n, device=local, terrain_seeds=parallel.shard_seeds(rank, n), rng_seed=100 + rank)
tr = trl.MACETrainer(sc, replay_mem_size=200000, num_init_samples=4000 * world, freeze_target_iters=50, seed=9)
comm = parallel.Comm(sc, rank, world, backend="nccl")
L = sc.L
sp = np.array([0.9, 0.2, 20.0, 0.025, 0.9, 0.002, 2000.0, 2000.0, 0.0])
L.trl_train_run_timed.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
state, ms = C.c_int64(0), C.c_double(0)


def run(k):
    if not root_trainer:
        assert L.trl_train_run_timed(tr.h, sp.ctypes.data_as(C.c_void_p), k, iters, 1024, C.c_double(1.0 / 30.0), 0, C.byref(state),
                                     C.byref(ms)) == 0, L.trl_last_error().decode()
        return ms.value
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sc.Sync(); torch.cuda.synchronize(); e0.record()
    for _ in range(k):
        sc.Update(1.0 / 30.0)
        comm.GatherTuples(1024)
        if rank == 0:
            comm.AddGathered(tr); tr.Train(iters)
        comm.BroadcastTrainer(tr, 0)
    sc.Sync(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)


run(60)
dist.barrier()
t_ms = run(updates)
t = torch.tensor([t_ms], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
spread = comm.ReplicaSpread(tr)
c = tr.counters()
if rank == 0:
    print(json.dumps({"n_gpus": world, "envs_per_gpu": n, "updates": updates, "iters_per_update": iters, "mode": "root trainer + broadcast" if root_trainer else "replicated trainers",
                      "train_loop_env_steps_per_s": world * n * 20 * updates / (float(t.item()) * 1e-3), "ms_per_update": float(t.item()) / updates,
                      "gather_ms_last": comm.LastGatherMs(), "replica_spread": spread, "trainer_iter": c["iter"], "actor_iter": c["actor_iter"],
                      "replay_tuples": c["num"], "critic_loss": c["critic_loss"], "tuples_dropped": comm.TuplesDropped()}))
tr.close(); comm.close(); sc.close()
dist.destroy_process_group()
