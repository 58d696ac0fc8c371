"""world_size-2 gloo test of the multi-GPU plumbing (env sharding + tuple-block all-gather + stats reduction)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, pack, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pyoracle import Oracle
    from deepterrainrl_b200.parallel import (gather_tuple_blocks, gather_tuple_blocks_fixed, reduce_eval_stats, shard_seeds,
                                             unpack_tuple_blocks)
    n = 4
    seeds = shard_seeds(rank, n)
    o = Oracle(pack, n, 1, terrain_seeds=seeds, rng_seed=99 + rank)   # the oracle stands in for one rank's engine
    o.set_explore(1, 0.3, 0.025, 0.02)
    for _ in range(60):
        o.update(1.0 / 30.0, 2)
    rows, flags, ids = o.tuples()
    cap = 256
    R = torch.zeros((cap, rows.shape[1]), dtype=torch.float64); R[:len(rows)] = torch.from_numpy(rows)
    F = torch.zeros(cap, dtype=torch.int32); F[:len(rows)] = torch.from_numpy(flags.astype(np.int32))
    E = torch.zeros(cap, dtype=torch.int32); E[:len(rows)] = torch.from_numpy(ids)
    cnt = torch.tensor([len(rows)], dtype=torch.int32)
    g_rows, g_flags, g_env = gather_tuple_blocks(R, F, E, cnt, env_offset=rank * n)
    f_rows, f_flags, f_env = unpack_tuple_blocks(gather_tuple_blocks_fixed(R, F, E, cnt, env_offset=rank * n, block_rows=128))
    assert torch.equal(f_rows, g_rows) and torch.equal(f_flags, g_flags) and torch.equal(f_env, g_env)
    st = reduce_eval_stats(o.eval_stats())
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), rows=g_rows.numpy(), flags=g_flags.numpy(), env=g_env.numpy(),
             local=rows.astype(np.float32), local_n=len(rows), steps=st["steps"], seeds=seeds)
    dist.destroy_process_group()


def test_gloo_world2_tuple_gather(assets, tmp_path):
    pack = os.path.join(assets, "dog_slopes_mixed.trlpack")
    port = 29500 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(2, port, pack, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "r0.npz"); r1 = np.load(tmp_path / "r1.npz")
    # every rank ends up with the same gathered set = concat of both ranks' local tuples, env ids globalised
    np.testing.assert_array_equal(r0["rows"], r1["rows"])
    np.testing.assert_array_equal(r0["env"], r1["env"])
    n0, n1 = int(r0["local_n"]), int(r1["local_n"])
    assert n0 > 0 and n1 > 0 and r0["rows"].shape[0] == n0 + n1
    np.testing.assert_array_equal(r0["rows"][:n0], r0["local"])
    np.testing.assert_array_equal(r0["rows"][n0:], r1["local"])
    assert r0["env"][:n0].max() < 4 and r0["env"][n0:].min() >= 4
    assert int(r0["steps"]) == 2 * 4 * 60 * 20
    assert set(r0["seeds"]).isdisjoint(set(r1["seeds"]))
