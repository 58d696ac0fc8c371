"""world_size-2 gloo tests of the native multi-GPU exchange (csrc/trl_comm.cu): the C ABI's pack kernel + all-gather + trainer
hand-over + broadcast, driven exactly as on GPUs, with every rank's engine and trainer running the kernel sources on the SIMT
emulator and the collectives supplied through trl_comm_init_external (torch.distributed / gloo on host memory)."""
import os
import platform
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(platform.machine() != "x86_64", reason="the emulator's fiber switch is written for x86-64")


def _setup(rank, world, port):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests", "simt"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from loader import open_simt
    from deepterrainrl_b200 import scenario
    scenario._LIB = open_simt()                          # this process only ever sees the emulator build
    import deepterrainrl_b200 as trl
    from deepterrainrl_b200 import parallel
    return trl, parallel


def _gather_worker(rank, world, port, pack, out_dir):
    trl, parallel = _setup(rank, world, port)
    n = 6
    sc = trl.ScenarioExpMACE(pack, n, terrain_seeds=parallel.shard_seeds(rank, n), rng_seed=100 + rank)
    comm = parallel.Comm(sc, rank, world, backend="external")
    sc.EnableExplore(True, 0.3, 0.025, 0.02)
    got_rows, got_flags, got_env, local_rows, counts_log = [], [], [], [], []
    for k in range(48):
        sc.Update(1.0 / 30.0)
        if k % 12 == 11:
            loc, lf, le = sc.GetTuples()                 # what this rank is about to contribute, ordered as in its block
            local_rows.append(loc)
            # a block of 4 rows: more tuples than one gather carries -> the remainder must stay queued, not be lost
            for _ in range(64):
                comm.GatherTuples(block_rows=4)
                counts, rows, flags, env = comm.Fetch()
                counts_log.append(counts.copy())
                got_rows.append(rows); got_flags.append(flags); got_env.append(env)
                if counts.sum() == 0:
                    break
            assert sc.GetNumTuples() == 0
    st = comm.EvalStats()
    np.savez(os.path.join(out_dir, f"g{rank}.npz"), rows=np.concatenate(got_rows), flags=np.concatenate(got_flags), env=np.concatenate(got_env),
             local=np.concatenate(local_rows), counts=np.stack(counts_log), steps=st["steps"], cycles=st["cycles"],
             local_cycles=sc._stats()["cycles"], dropped=comm.TuplesDropped())
    comm.close(); sc.close()
    dist.destroy_process_group()


def test_gloo_world2_native_tuple_gather(assets, tmp_path):
    pack = os.path.join(assets, "dog_slopes_mixed.trlpack")
    port = 29500 + (os.getpid() % 1000)
    mp.spawn(_gather_worker, args=(2, port, pack, str(tmp_path)), nprocs=2, join=True)
    g0 = np.load(tmp_path / "g0.npz"); g1 = np.load(tmp_path / "g1.npz")
    # every rank sees the same gathered stream
    for key in ("rows", "flags", "env", "counts"):
        np.testing.assert_array_equal(g0[key], g1[key])
    assert g0["counts"].max() <= 4 and g0["counts"][:, 0].sum() == len(g0["local"]) and g0["counts"][:, 1].sum() == len(g1["local"])
    assert len(g0["local"]) > 8 and len(g1["local"]) > 8          # more than two blocks' worth per rank: the queue was exercised
    # ... which is every rank's local tuples, nothing lost, nothing duplicated, f32 as cMACETrainer stores them, env ids globalised
    by_rank = [g0["rows"][(g0["env"] >= 6 * r) & (g0["env"] < 6 * (r + 1))] for r in range(2)]
    for r, loc in enumerate((g0["local"], g1["local"])):
        assert by_rank[r].shape == loc.shape
        np.testing.assert_array_equal(by_rank[r], loc.astype(np.float32))   # order inside a rank is preserved by the queue
    assert int(g0["steps"]) == int(g1["steps"]) == 2 * 6 * 48 * 20
    assert int(g0["cycles"]) == int(g0["local_cycles"]) + int(g1["local_cycles"])
    assert int(g0["dropped"]) == 0


def _train_worker(rank, world, port, pack, out_dir):
    trl, parallel = _setup(rank, world, port)
    n = 6
    sc = trl.ScenarioExpMACE(pack, n, terrain_seeds=parallel.shard_seeds(rank, n), rng_seed=100 + rank)
    tr = trl.MACETrainer(sc, replay_mem_size=160, num_init_samples=32, freeze_target_iters=3, seed=9)
    comm = parallel.Comm(sc, rank, world, backend="external")
    if rank == 1:
        tr.set_theta(tr.get("theta") * 0.5)             # a replica that starts out of sync ...
    assert comm.ReplicaSpread(tr) > 0
    comm.BroadcastTrainer(tr, root=0)                   # ... is brought back by cNeuralNetLearner::SyncNet across ranks
    assert comm.ReplicaSpread(tr) == 0.0
    # a critic batch needs 32 tuples: the same synthetic tuples on both ranks, then the gathered rollout tuples on top
    rng = np.random.default_rng(4)
    rows = rng.normal(size=(40, tr.W)) * 0.1
    rows[:, 1 + tr.S] = rng.integers(0, 3, 40)
    tr.AddTuples(rows, np.zeros(40, np.uint32))
    sc.EnableExplore(True, 0.1, 0.025, 0.0)
    for k in range(34):
        sc.Update(1.0 / 30.0)
        comm.GatherTuples(block_rows=64)
        comm.AddGathered(tr)
        if k >= 32:
            tr.Train(1)                                 # the first call initialises the input offset / scale
    c = tr.counters()
    spread = comm.ReplicaSpread(tr)
    np.savez(os.path.join(out_dir, f"e{rank}.npz"), theta=tr.get("theta"), iters=c["iter"], num=c["num"], total=c["total"], critic=c["critic"],
             stage=c["stage"], local_cycles=sc._stats()["cycles"], spread=spread)
    tr.close(); comm.close(); sc.close()
    dist.destroy_process_group()


def test_gloo_world2_emulated_engines_replicated_trainers(assets, tmp_path):
    """BASELINE configs[3] in miniature on the CPU tier: two ranks, each rolling out its own shard with the env-step / decision
    kernel sources, ONE all-gather of the packed tuple blocks per outer update through the C ABI, the gathered blocks fed to each
    rank's own on-device trainer -- the replicas must stay bit-identical without a weight broadcast."""
    pack = os.path.join(assets, "dog_slopes_mixed.trlpack")
    port = 29500 + ((os.getpid() + 517) % 1000)
    mp.spawn(_train_worker, args=(2, port, pack, str(tmp_path)), nprocs=2, join=True)
    e0 = np.load(tmp_path / "e0.npz"); e1 = np.load(tmp_path / "e1.npz")
    assert int(e0["total"]) == int(e1["total"]) > 40 and int(e0["num"]) == int(e1["num"]) == int(e0["total"])      # 40 seeded + gathered
    assert int(e0["iters"]) == int(e1["iters"]) >= 2, (int(e0["total"]), int(e0["critic"]), int(e0["stage"]))
    assert int(e0["local_cycles"]) > 0 and int(e1["local_cycles"]) > 0
    np.testing.assert_array_equal(e0["theta"], e1["theta"])          # bit-identical replicas
    assert float(e0["spread"]) == 0.0 and float(e1["spread"]) == 0.0
