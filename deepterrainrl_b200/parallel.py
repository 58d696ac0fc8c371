"""Multi-GPU plumbing: environments shard across ranks (one process per GPU, torch.distributed), no collective in
the rollout data path.  The only exchange is the hand-off of finished ExpTuples to the trainer side, done as one
all-gather of fixed-shape tuple blocks per outer update -- this replaces the reference's in-process
`learner->Train(tuples)` under the trainer mutex (scenarios/ScenarioTrain.cpp:388-395, learning/NeuralNetLearner.cpp:33-46).

Works on any backend: NCCL with the device-resident block from ScenarioExpMACE.DeviceTupleBlock() (zero-copy), or
gloo with CPU tensors (what the CPU tests exercise at world_size 2).
"""
import numpy as np


def shard_seeds(rank, envs_per_rank, base=1):
    """Terrain seed of env i on rank r: base + r * envs_per_rank + i."""
    return np.arange(base + rank * envs_per_rank, base + (rank + 1) * envs_per_rank, dtype=np.uint64)


def gather_tuple_blocks(rows, flags, env_ids, count, env_offset=0, pad_to=64, group=None):
    """All-gather the first `count` tuple rows of every rank.

    rows: [cap, W] float tensor (f64 on device or CPU), flags/env_ids: [cap] int32, count: [1] int32, all on the
    backend's device.  Returns (rows_f32 [total, W], flags [total], env_ids [total] globalised with env_offset) on
    every rank -- float32 rows are what cMACETrainer::AddTuples stores (learning/MACETrainer.cpp:515-539).
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count.clone(), group=group)
    counts = [int(c.item()) for c in counts]
    cap = rows.shape[0]
    m = min(max(counts), cap)
    if m == 0:
        W = rows.shape[1]
        return (torch.zeros((0, W), dtype=torch.float32, device=rows.device),
                torch.zeros((0,), dtype=torch.int32, device=rows.device),
                torch.zeros((0,), dtype=torch.int32, device=rows.device))
    m = min(cap, ((m + pad_to - 1) // pad_to) * pad_to)
    blk_rows = rows[:m].to(torch.float32).contiguous()
    blk_flags = flags[:m].to(torch.int32).contiguous()
    blk_env = (env_ids[:m].to(torch.int32) + int(env_offset)).contiguous()
    out_rows = torch.empty((world,) + tuple(blk_rows.shape), dtype=blk_rows.dtype, device=rows.device)
    out_flags = torch.empty((world, m), dtype=torch.int32, device=rows.device)
    out_env = torch.empty((world, m), dtype=torch.int32, device=rows.device)
    dist.all_gather_into_tensor(out_rows.view(-1), blk_rows.view(-1), group=group) if rows.is_cuda else \
        dist.all_gather(list(out_rows.unbind(0)), blk_rows, group=group)
    dist.all_gather_into_tensor(out_flags.view(-1), blk_flags, group=group) if rows.is_cuda else \
        dist.all_gather(list(out_flags.unbind(0)), blk_flags, group=group)
    dist.all_gather_into_tensor(out_env.view(-1), blk_env, group=group) if rows.is_cuda else \
        dist.all_gather(list(out_env.unbind(0)), blk_env, group=group)
    keep = [torch.arange(min(c, m), device=rows.device) + r * m for r, c in enumerate(counts)]
    keep = torch.cat(keep)
    return (out_rows.view(world * m, -1)[keep], out_flags.view(-1)[keep], out_env.view(-1)[keep])


def gather_tuple_blocks_fixed(rows, flags, env_ids, count, env_offset=0, block_rows=1024, group=None):
    """Sync-free variant: ONE all-gather of a fixed-shape f32 block per rank, no host round-trip.

    Block layout per rank: [block_rows + 1, W + 2] float32; row 0 carries the rank's tuple count in column 0, rows
    1.. are [flags, global env id, reward, s, a, s'].  Returns the gathered tensor [world, block_rows + 1, W + 2] on the
    backend device; `unpack_tuple_blocks` trims it on the consumer side.  Rows beyond block_rows stay queued on the
    producer (the caller only resets its tuple buffer when count <= block_rows).
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    W = rows.shape[1]
    m = min(block_rows, rows.shape[0])
    blk = torch.zeros((block_rows + 1, W + 2), dtype=torch.float32, device=rows.device)
    blk[0, 0] = torch.clamp(count[0], max=m).to(torch.float32)
    blk[1:m + 1, 0] = flags[:m].to(torch.float32)
    blk[1:m + 1, 1] = (env_ids[:m] + int(env_offset)).to(torch.float32)
    blk[1:m + 1, 2:] = rows[:m].to(torch.float32)
    out = torch.empty((world,) + tuple(blk.shape), dtype=torch.float32, device=rows.device)
    if rows.is_cuda:
        dist.all_gather_into_tensor(out.view(-1), blk.view(-1), group=group)
    else:
        dist.all_gather(list(out.unbind(0)), blk, group=group)
    return out


def unpack_tuple_blocks(gathered):
    """(rows_f32 [total, W], flags int32, env int32) from the output of gather_tuple_blocks_fixed (syncs)."""
    import torch
    rows, flags, env = [], [], []
    for r in range(gathered.shape[0]):
        c = int(gathered[r, 0, 0].item())
        rows.append(gathered[r, 1:c + 1, 2:]); flags.append(gathered[r, 1:c + 1, 0].to(torch.int32))
        env.append(gathered[r, 1:c + 1, 1].to(torch.int32))
    return torch.cat(rows), torch.cat(flags), torch.cat(env)


def reduce_eval_stats(stats, group=None):
    """Sum (cycles, episodes, steps) and episode-weighted avg_dist over ranks (cOptScenarioPoliEval::OutputResults
    merges per-thread results under a mutex: optimizer/scenarios/OptScenarioPoliEval.cpp:213-237)."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([stats["cycles"], stats["episodes"], stats["steps"], stats["avg_dist"] * stats["episodes"]],
                     dtype=torch.float64, device=dev)
    dist.all_reduce(t, group=group)
    ep = t[1].item()
    return dict(cycles=int(t[0].item()), episodes=int(ep), steps=int(t[2].item()), avg_dist=(t[3].item() / ep if ep else 0.0))
