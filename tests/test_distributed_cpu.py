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


def _emu_worker(rank, world, port, pack, out_dir):
    """tools/train_multi.py at world size 2 without GPUs: every rank's engine AND trainer are the kernel sources on the SIMT emulator
    (tests/simt/), the exchange is the product's fixed-block all-gather over gloo."""
    import ctypes as C
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests", "simt"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from loader import open_simt
    from deepterrainrl_b200 import parallel, scenario
    import deepterrainrl_b200 as trl
    L = open_simt()
    scenario._LIB = L                                   # this process only ever sees the emulator build
    n = 6
    sc = trl.ScenarioExpMACE(pack, n, terrain_seeds=parallel.shard_seeds(rank, n), rng_seed=100 + rank)
    tr = trl.MACETrainer(sc, replay_mem_size=160, num_init_samples=32, freeze_target_iters=3, seed=9)
    sc.EnableExplore(True, 0.1, 0.025, 0.0)
    # the device tuple block as torch views (BatchedScenario.DeviceTupleBlock does the same through the CUDA array interface)
    ptrs = [C.c_void_p() for _ in range(4)]
    cap, width = C.c_int(0), C.c_int(0)
    assert L.trl_device_tuple_block(sc.h, *[C.byref(p) for p in ptrs], C.byref(cap), C.byref(width)) == 0

    def view(p, shape, ctype):
        return torch.from_numpy(np.ctypeslib.as_array(C.cast(p, C.POINTER(ctype)), shape))
    rows, flags = view(ptrs[0], (cap.value, width.value), C.c_double), view(ptrs[1], (cap.value,), C.c_int32)
    env, count = view(ptrs[2], (cap.value,), C.c_int32), view(ptrs[3], (1,), C.c_int32)
    total = 0
    for k in range(74):
        sc.Update(1.0 / 30.0)
        sc.Sync()
        g = parallel.gather_tuple_blocks_fixed(rows, flags, env, count, env_offset=rank * n, block_rows=64)
        sc.ResetTupleBuffer()
        r, f, e = parallel.unpack_tuple_blocks(g)
        if len(r):                                      # MACETrainer.AddTuplesDevice without its CUDA stream handling
            r64, f32 = r.to(torch.float64).contiguous(), f.to(torch.int32).contiguous()
            assert L.trl_trainer_add_device(tr.h, C.c_void_p(r64.data_ptr()), C.c_void_p(f32.data_ptr()), len(r64)) == 0
            sc.Sync()
            total += len(r64)
        if k >= 70:
            tr.Train(1)                                 # the first call initialises the input offset / scale
    c = tr.counters()
    np.savez(os.path.join(out_dir, f"e{rank}.npz"), theta=tr.get("theta"), total=total, iters=c["iter"], num=c["num"], critic=c["critic"], stage=c["stage"],
             local_cycles=sc._stats()["cycles"])
    tr.close(); sc.close()
    dist.destroy_process_group()


@__import__("pytest").mark.skipif(__import__("platform").machine() != "x86_64", reason="the emulator's fiber switch is written for x86-64")
def test_gloo_world2_emulated_engines_replicated_trainers(assets, tmp_path):
    """BASELINE configs[3] in miniature on the CPU tier: two ranks, each rolling out its own shard with the env-step / decision
    kernel sources, ONE fixed-block all-gather of the tuple blocks per outer update, the gathered block fed to each rank's own
    on-device trainer -- the replicas must stay bit-identical without a weight broadcast."""
    pack = os.path.join(assets, "dog_slopes_mixed.trlpack")
    port = 29500 + ((os.getpid() + 517) % 1000)
    mp.spawn(_emu_worker, args=(2, port, pack, str(tmp_path)), nprocs=2, join=True)
    e0 = np.load(tmp_path / "e0.npz"); e1 = np.load(tmp_path / "e1.npz")
    assert int(e0["total"]) == int(e1["total"]) >= 32 and int(e0["num"]) == int(e1["num"]) == int(e0["total"])
    assert int(e0["iters"]) == int(e1["iters"]) >= 2, (int(e0["total"]), int(e0["critic"]), int(e0["stage"]))
    assert int(e0["local_cycles"]) > 0 and int(e1["local_cycles"]) > 0
    np.testing.assert_array_equal(e0["theta"], e1["theta"])          # bit-identical replicas
