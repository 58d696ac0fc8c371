"""torchrun entry: explore-mode rollout on every rank + ONE NCCL all-gather of the device-resident tuple block per
outer update.  Checks that every rank ends up with the union of all ranks' tuples.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/nccl_gather_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepterrainrl_b200 as trl  # noqa: E402
from deepterrainrl_b200.parallel import (gather_tuple_blocks, gather_tuple_blocks_fixed, reduce_eval_stats,  # noqa: E402
                                         shard_seeds, unpack_tuple_blocks)


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 512
    sc = trl.ScenarioExpMACE(os.path.join(ROOT, "assets", "dog_slopes_mixed.trlpack"), n, device=local,
                             terrain_seeds=shard_seeds(rank, n), rng_seed=1234 + rank)
    sc.EnableExplore(1, 0.2, 0.025, 0.002)
    total = 0
    local_total = 0
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    gather_ms = 0.0
    for it in range(60):
        sc.Update(1.0 / 30.0)
        sc.Sync()
        rows, flags, env, count = sc.DeviceTupleBlock()
        local_n = int(count.item())
        ev0.record()
        g_rows, g_flags, g_env = gather_tuple_blocks(rows, flags, env, count, env_offset=rank * n)
        ev1.record(); torch.cuda.synchronize()
        gather_ms += ev0.elapsed_time(ev1)
        # sync-free fixed-block variant: must deliver the same rows
        ev2 = torch.cuda.Event(enable_timing=True); ev3 = torch.cuda.Event(enable_timing=True)
        ev2.record()
        gathered = gather_tuple_blocks_fixed(rows, flags, env, count, env_offset=rank * n, block_rows=512)
        ev3.record(); torch.cuda.synchronize()
        fixed_ms = locals().get("fixed_ms", 0.0) + ev2.elapsed_time(ev3)
        f_rows, f_flags, f_env = unpack_tuple_blocks(gathered)
        assert torch.equal(f_rows, g_rows) and torch.equal(f_env, g_env) and torch.equal(f_flags, g_flags)
        sc.ResetTupleBuffer()
        # every rank must hold the same union; check with a checksum all-reduce
        cs = torch.stack([g_rows.double().sum(), g_flags.double().sum(), g_env.double().sum(),
                          torch.tensor(float(g_rows.shape[0]), device="cuda")])
        mx = cs.clone(); mn = cs.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        assert torch.equal(mx, mn), (rank, it, mx, mn)
        nl = torch.tensor([local_n], device="cuda"); dist.all_reduce(nl)
        assert int(nl.item()) == g_rows.shape[0], (int(nl.item()), g_rows.shape)
        if local_n:
            mine = g_env // n == rank
            assert int(mine.sum().item()) == local_n
            assert torch.equal(g_rows[mine], rows[:local_n].float())
        total += g_rows.shape[0]; local_total += local_n
    st = reduce_eval_stats(sc._stats())
    if rank == 0:
        print(f"nccl gather ok: world={world} tuples gathered={total} (rank0 local {local_total}) env_steps={st['steps']} "
              f"avg all-gather {gather_ms / 60:.3f} ms per update (3-tensor variant with host count sync), "
              f"{fixed_ms / 60:.3f} ms (single fixed-shape block, sync-free)")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
