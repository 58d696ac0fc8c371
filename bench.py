#!/usr/bin/env python
"""Benchmark: env-steps/sec of the ScenarioPoliEval step loop (dog / slopes_mixed / MACE policy, 4096 envs per GPU).

One "step" = one outer cScenario::Update(1/30 s) over the whole batch = 20 env-steps x 4096 envs (one env-step =
one iteration of the loop at scenarios/ScenarioSimChar.cpp:162-173 = 1/600 s simulated).  Data is synthetic in the
contract's sense (procedurally generated terrain from per-env seeds); the character, controller and policy weights
are the reference's shipped dog / dog_mace3_slopes_mixed assets baked into assets/dog_slopes_mixed.trlpack.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--envs E]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  `value` = whole-job env-steps/s with state resident in HBM, timed with CUDA events on
the engine's own stream; `e2e` = the same through the public C-ABI calls a user of cOptScenarioPoliEval makes per
step (Update + statistics read-back to host buffers); `roofline` = algorithmic HBM bytes of the step kernel per
launch / its measured launch duration vs the measured copy peak; `cpu_baseline` = the CPU oracle timed on this host.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PACK = os.path.join(ROOT, "assets", "dog_slopes_mixed.trlpack")
WORKLOAD = "dog/slopes_mixed MACE poli_eval (BASELINE configs[1])"
WORKLOADS = {"dog_slopes_mixed": WORKLOAD, "raptor_narrow_gaps": "raptor/narrow_gaps MACE poli_eval (BASELINE configs[2])",
             "goat_cliffs": "goat/cliffs MACE poli_eval (BASELINE configs[4] per-GPU share)", "dog_flat": "dog/flat fixed gait (BASELINE configs[0])"}
ENV_STEPS_PER_UPDATE = 20
DT = 1.0 / 30.0
# per env-step the persistent state must be read and written once (q, qd, held torque, 64-scalar
# controller block = 133 scalars each way) plus <= 42 terrain vertices read: (133 * 2) * 8 B + 42 * 4 B for f64 state
ALGO_BYTES_PER_ENV_STEP = 133 * 2 * 8 + 42 * 4
# ncu (profiles/ncu_step_kernel_r02.csv, dog / slopes_mixed, one 4096-env step launch of the shipped build): DFMA + DADD + DMUL
# thread-instructions per cycle (465.8 + 357.8 + 115.7) x 295,256 elapsed cycles = 277 M per launch (the arithmetic has not changed
# since the first round-2 capture: 278 M)
FP64_INST_PER_LAUNCH_ENV = 277.3e6 / 4096


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """SM clock / throttle reasons sampled through NVML (nvidia-ml-py) every 5 ms while the timed region runs
    (same fields as the nvidia-smi line in B200_PROFILING.md; NVML avoids nvidia-smi's start-up and pipe buffering,
    which matter because the timed region is only ~150 ms long)."""

    REASONS = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.thread = None
        self.max_mhz = None
        self.err = None

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop_flag:
                self.samples.append((time.perf_counter(), float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))))
                r = int(get_reasons(h))
                for name, bit in self.REASONS.items():
                    if r & bit:
                        self.reasons.add((time.perf_counter(), name))
                time.sleep(0.005)
            pynvml.nvmlShutdown()
        except Exception as e:   # noqa: BLE001
            self.err = repr(e)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self, t0=None, t1=None):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=2)
        sel = [v for (t, v) in self.samples if (t0 is None or t >= t0) and (t1 is None or t <= t1)]
        reasons = sorted({n for (t, n) in self.reasons if (t0 is None or t >= t0) and (t1 is None or t <= t1)})
        sel.sort()
        out = {"sm_mhz": sel[len(sel) // 2] if sel else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
               "samples": len(sel), "source": "nvml"}
        if self.err:
            out["error"] = self.err
        return out


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of one trl_step_kernel launch, from the committed `ncu --set full`
    summary of the same workload and build (profiles/ncu_step_kernel_r02.csv, regenerated whenever the kernel changes;
    tools/ncu_summary.py writes it); None if the summary is absent."""
    path = os.path.join(ROOT, "profiles", "ncu_step_kernel_r02.csv")
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", "ncu_step_kernel_r01_final.csv")
    unit_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total, seen = 0.0, 0
    try:
        for line in open(path):
            f = line.strip().split(",")
            if len(f) == 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                total += float(f[2]) * unit_scale.get(f[1], 1.0)
                seen += 1
    except OSError:
        return None
    return total if seen == 2 else None


def granted_cores():
    """CPUs this process can actually use: the affinity mask, capped by the cgroup CPU quota (cpu.max / cfs_quota).
    os.cpu_count() reports the host's CPUs, which a container may not get."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    info = {"affinity": n, "host": os.cpu_count() or 1, "cgroup_quota": None}
    quota = None
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:               # cgroup v2: "<quota|max> <period>"
            q, per = f.read().split()[:2]
            if q != "max":
                quota = float(q) / float(per)
    except (OSError, ValueError):
        try:
            with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as f, open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as g:
                q, per = float(f.read()), float(g.read())
                if q > 0:
                    quota = q / per
        except (OSError, ValueError):
            pass
    if quota is not None:
        info["cgroup_quota"] = quota
        n = max(1, min(n, int(quota + 0.5)))
    info["used"] = n
    return n, info


def ncu_metric(name):
    """one metric of the committed `ncu --set full` summary of the step kernel (profiles/ncu_step_kernel_r02.csv), or None"""
    try:
        for line in open(os.path.join(ROOT, "profiles", "ncu_step_kernel_r02.csv")):
            f = line.strip().split(",")
            if len(f) == 3 and f[0] == name:
                return float(f[2])
    except (OSError, ValueError):
        pass
    return None


def cpu_reference(num_envs, seconds_target, threads, min_updates=1):
    """Times the CPU oracle (restated reference controller + this project's reduced-coordinate physics, f64; the
    -O3 -march=x86-64-v3 build when the host has AVX2 + FMA) with thread-per-env-slice like cOptScenarioPoliEval::Run
    (optimizer/scenarios/OptScenarioPoliEval.cpp:72-110), one pinned thread per granted core."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    o = pyoracle.Oracle(PACK, num_envs, 0, fast=True)
    o.update(DT, threads)   # warm-up
    t0 = time.perf_counter()
    updates = 0
    while updates < min_updates or time.perf_counter() - t0 < seconds_target:
        o.update(DT, threads)
        updates += 1
    dt = time.perf_counter() - t0
    steps = updates * ENV_STEPS_PER_UPDATE * num_envs
    build = "x86-64-v3 (AVX2+FMA)" if pyoracle.host_has_avx2_fma() else "portable x86-64"
    return steps / dt, dt, updates, build


def run_reference_arm(args, rank):
    """The CPU arm on the SAME configuration as the GPU arm (args.envs environments, one outer update per step), on every core
    this process is granted.  A step is one outer update of all args.envs envs (81,920 env-steps at 4096); at least
    --cpu-seconds (default 12 s) of timed CPU work, so a fast host runs more than K steps and the rate is a mean over all of them."""
    if rank != 0:
        return
    cores, core_info = granted_cores()
    envs = args.envs
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    o = pyoracle.Oracle(PACK, envs, 0, fast=True)
    for _ in range(max(1, min(args.warmup, 2))):       # the CPU has no clocks to ramp: two warm-up updates fill caches / page in
        o.update(DT, cores)
    t0 = time.perf_counter()
    done = 0
    while done < args.steps or time.perf_counter() - t0 < args.cpu_seconds:
        o.update(DT, cores)
        done += 1
        if time.perf_counter() - t0 > 150.0:            # hard stop: the whole arm stays within a few minutes on any host
            break
    el = time.perf_counter() - t0
    per_step_s = el / done
    value = envs * ENV_STEPS_PER_UPDATE / per_step_s
    build = "x86-64-v3 (AVX2+FMA)" if pyoracle.host_has_avx2_fma() else "portable x86-64"
    line = {
        "impl": "reference", "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_per_gpu": envs, "env_steps_per_step": ENV_STEPS_PER_UPDATE * envs, "sim_substeps": 5,
                   "steps_timed": done, "seconds_timed": el},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "per_thread": value / cores, "core_info": core_info, "oracle_build": build,
                         "sample": f"{envs} envs x {done} outer updates of 20 env-steps ({el:.1f} s), thread-per-env-slice on {cores} pinned threads; "
                                   "restated CPU oracle (reduced-coordinate physics), not Bullet"},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_config4(args, trl, rank, local_rank, world, n, dist, torch):
    """BASELINE configs[3]: dog / slopes_mixed, args.envs envs per GPU, exploration on (args/opt_args_train_mace.txt rates), tuples of
    all ranks all-gathered once per outer update into every rank's on-device trainer, args.train_iters trainer iterations per
    update.  The whole loop is the C ABI's trl_train_run_timed (no Python between updates), device-timed on the engine stream, max
    over ranks.  Returns the `config4` block of the JSON line."""
    import ctypes as C
    import numpy as np
    from deepterrainrl_b200 import parallel
    seeds = parallel.shard_seeds(rank, n)
    sc = trl.ScenarioExpMACE(PACK, n, device=local_rank, terrain_seeds=seeds, rng_seed=100 + rank)
    tr = trl.MACETrainer(sc, replay_mem_size=args.replay_size, num_init_samples=4000 * world, freeze_target_iters=50, seed=9)
    L = sc.L
    if world > 1:
        comm = parallel.Comm(sc, rank, world, backend="nccl")          # the library's own NCCL communicator; id shipped over torch.distributed
    else:
        uid = (C.c_ubyte * 128)()
        if L.trl_comm_unique_id(uid) != 0:
            raise RuntimeError(L.trl_last_error().decode())
        comm = parallel.Comm(sc, 0, 1, backend="nccl", unique_id=bytes(uid))
    if not args.train_sync:
        if L.trl_trainer_set_async(tr.h, 1) != 0:
            raise RuntimeError(L.trl_last_error().decode())
    sp = np.array([0.9, 0.2, 20.0, 0.025, 0.9, 0.002, 2000.0, 2000.0, 0.0])    # cScenarioTrain annealing, shortened to 2000 iterations
    L.trl_train_run_timed.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    state = C.c_int64(0)
    ms = C.c_double(0)

    def run(k, flush=1):
        if L.trl_train_run_timed(tr.h, sp.ctypes.data_as(C.c_void_p), int(k), int(args.train_iters), int(args.block_rows), C.c_double(DT),
                                 int(flush), C.byref(state), C.byref(ms)) != 0:
            raise RuntimeError(L.trl_last_error().decode())
        return ms.value

    run(int(round(args.presim / DT)), flush=0)        # pre-roll: desynchronise the gait cycles, fill the replay memory past the init stage
    run(max(args.warmup, 3))
    l0, t0 = sc.KernelLaunches(), tr.KernelLaunches()
    if dist is not None:
        dist.barrier()
    sc.Sync()
    t_ms = run(args.steps)
    launches = (sc.KernelLaunches() - l0) + (tr.KernelLaunches() - t0)
    # the same K updates without exchange and trainer = what the rollout alone costs in exploration mode on this rank
    sc.Sync()
    roll_ms = sc.BenchUpdates(args.steps, DT, flush_l2=True)
    # exposed device time of one exchange (pack end -> all-gather end on the comm stream), mean over 8 updates
    gm = []
    if not args.train_sync:
        L.trl_trainer_set_async(tr.h, 0)
    for _ in range(8):
        sc.Update(DT); comm.GatherTuples(args.block_rows); comm.AddGathered(tr)
        gm.append(comm.LastGatherMs())
    counts = comm.Fetch(cap=1)[0]
    spread = comm.ReplicaSpread(tr)
    c = tr.counters()
    dropped = comm.TuplesDropped()
    if dist is not None:
        t = torch.tensor([t_ms, roll_ms, float(np.mean(gm))], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ms, roll_ms, g_ms = [float(x) for x in t.tolist()]
    else:
        g_ms = float(np.mean(gm))
    out = {"workload": "dog/slopes_mixed exploration + tuple all-gather + MACE trainer (BASELINE configs[3])",
           "value": args.steps * ENV_STEPS_PER_UPDATE * n * world / (t_ms * 1e-3), "unit": "env-steps/s", "ms_per_step": t_ms / args.steps,
           "rollout_only_ms_per_step": roll_ms / args.steps, "gather_ms": g_ms,
           "gather": "device pack kernel + one ncclAllGather of a fixed block per rank on a side stream (trl_gather_tuples), no host sync",
           "block_rows": args.block_rows, "block_bytes_per_rank": 16 + 8 * args.block_rows + 4 * args.block_rows * sc.tuple_width,
           "tuples_last_update_per_rank": [int(x) for x in counts], "tuples_dropped": int(dropped),
           "train_iters_per_update": args.train_iters, "trainer": "replicated on every rank (deterministic; no weight broadcast needed)",
           "trainer_mode": "synchronous (between the updates)" if args.train_sync else "asynchronous (own stream, overlaps the next update; policy snapshot refreshed between updates)",
           "trainer_iter": c["iter"], "actor_iter": c["actor_iter"], "replay_tuples": c["num"], "replay_capacity": args.replay_size, "replica_spread": spread,
           "gpu_launches": launches, "envs_per_gpu": n, "l2": "flushed between updates (256 MiB memset on the engine stream)",
           "timing": "cudaEvent on the engine stream around K x {update, pack + all-gather, hand-over, trainer iterations}, max over ranks"}
    tr.close(); comm.close(); sc.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--envs", type=int, default=4096, help="environments per GPU")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--scene", default="dog_slopes_mixed", choices=sorted(WORKLOADS),
                    help="asset pack; the default is the configuration the metric is quoted on, the others are side measurements")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--config4", type=int, default=1,
                    help="also measure BASELINE configs[3] in the same run: exploration on, tuple all-gather every outer update, "
                         "trainer iterations (0 = skip)")
    ap.add_argument("--train-iters", type=int, default=4, help="trainer iterations per outer update in the config-4 block")
    ap.add_argument("--block-rows", type=int, default=1024, help="tuple rows per rank and all-gather in the config-4 block")
    ap.add_argument("--replay-size", type=int, default=200000,
                    help="replay memory of the config-4 trainer (tuples); a small value puts the run into the wrapped-ring steady state of a long training run")
    ap.add_argument("--train-sync", type=int, default=0,
                    help="config-4 block: 1 = the trainer runs between the updates (synchronous); default 0 = the reference's asynchronous "
                         "trainer semantics: hand-over + training overlap the next update on their own stream")
    ap.add_argument("--presim", type=float, default=4.0,
                    help="seconds of simulated time run (untimed) before warm-up so gait cycles / episodes of the envs "
                         "are desynchronised like in a long evaluation (SURVEY §8d: warm-up 2 s sim)")
    args = ap.parse_args()
    global PACK, WORKLOAD
    PACK = os.path.join(ROOT, "assets", args.scene + ".trlpack")
    WORKLOAD = WORKLOADS[args.scene]

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import deepterrainrl_b200 as trl

    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = args.envs
    seeds = np.arange(1 + rank * n, 1 + (rank + 1) * n, dtype=np.uint64)   # SURVEY §8(d) config 4 seeding
    sc = trl.ScenarioPoliEval(PACK, n, device=local_rank, terrain_seeds=seeds, rng_seed=1234 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        sc.Sync()

    # ---- untimed pre-roll to a statistically steady state (all envs start from the same pose, so their gait cycles,
    # policy decisions and falls are synchronised at first: per-update cost is atypically low then), then W warm-up steps
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                # nvmlInit on a cold box takes a while: started before the pre-roll, checked before the timed region
    presim_updates = int(round(args.presim / DT))
    if presim_updates > 0:
        sc.BenchUpdates(presim_updates, DT, flush_l2=False)
    sc.BenchUpdates(max(args.warmup, 3), DT, flush_l2=True)
    launches0 = sc.KernelLaunches()

    # ---- timed: exactly K steps, device-timed on the engine's stream, max over ranks
    if rank == 0:
        t_wait = time.perf_counter()
        while not sampler.samples and sampler.err is None and time.perf_counter() - t_wait < 10.0:
            time.sleep(0.01)
    barrier()
    t_beg = time.perf_counter()
    ms = sc.BenchUpdates(args.steps, DT, flush_l2=True)
    span_ms = sc.BenchLastSpan()
    t_end = time.perf_counter()
    barrier()
    clocks = sampler.stop(t_beg, t_end) if rank == 0 else None
    launches = sc.KernelLaunches() - launches0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    total_env_steps = args.steps * ENV_STEPS_PER_UPDATE * n * world
    value = total_env_steps / (ms * 1e-3)

    # ---- e2e: the per-step calls of cOptScenarioPoliEval::EvalHelper through the C ABI with HOST buffers: every step
    # Update(1/30) and a read-back of the batch counters + every env's pose and velocity into host arrays.  The
    # read-back of step k is pipelined behind the launch of step k+1 (trl_snapshot / trl_snapshot_wait, pinned memory).
    pose = np.zeros((sc.num_dof, n)); vel = np.zeros((sc.num_dof, n))
    for _ in range(3):                      # untimed: first call allocates the pinned staging buffer and the copy stream
        sc.Update(DT); sc.Snapshot(); sc.SnapshotWait(pose, vel)
    barrier()
    t0 = time.perf_counter()
    sc.Update(DT); sc.Snapshot()
    for _ in range(args.steps - 1):
        sc.Update(DT)
        st = sc.SnapshotWait(pose, vel)
        sc.Snapshot()
    st = sc.SnapshotWait(pose, vel)
    e2e_s = time.perf_counter() - t0
    assert np.all(np.isfinite(pose))
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = total_env_steps / e2e_s
    d2h = 2 * sc.num_dof * n * 8 + 32

    # ---- roofline of the dominant kernel (step kernel), measured live with per-launch events
    step_ms, step_l, dec_ms, dec_l = sc.UpdateTimed(DT)
    peaks, peak_kind = measured_peaks()
    launch_s = step_ms * 1e-3 / step_l
    # the step_l = 21 launches of one update advance every env by 20 env-steps (S_0 runs the physics half only, S_end the
    # controller half only): one launch carries 20/21 of an env-step's algorithmic bytes per env
    env_steps_per_launch = ENV_STEPS_PER_UPDATE / step_l
    achieved = ALGO_BYTES_PER_ENV_STEP * n * env_steps_per_launch / launch_s / 1e9
    stats = sc._stats()
    if world > 1:
        # cOptScenarioPoliEval::OutputResults' merge of the per-scene results, over the ranks: one all-reduce through the C ABI
        from deepterrainrl_b200 import parallel
        ev_comm = parallel.Comm(sc, rank, world, backend="nccl")
        stats = ev_comm.EvalStats()
        ev_comm.close()

    # ---- BASELINE configs[3] in the same run: exploration on, ONE all-gather of the tuple blocks per outer update through the
    # C ABI (trl_gather_tuples), every rank's trainer fed with all ranks' tuples, trainer iterations between the updates
    cfg4 = None
    if args.config4 and args.scene == "dog_slopes_mixed":
        cfg4 = run_config4(args, trl, rank, local_rank, world, n, dist if world > 1 else None, torch)

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cores, core_info = granted_cores()
    cpu_envs = n if args.cpu_seconds >= 8.0 else 4 * cores      # the stated configuration unless a short smoke run asked for less
    cpu_val, cpu_dt, cpu_updates, cpu_build = cpu_reference(cpu_envs, args.cpu_seconds, cores) if world == 1 else (None, 0, 0, "")

    line = {
        "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_per_gpu": n,
                   "env_steps_per_step": ENV_STEPS_PER_UPDATE * n, "sim_substeps": 5, "parallelism": f"env-shard x{world}",
                   "l2": "flushed before every step (256 MiB memset on the engine stream, between the per-step event pairs: the flush itself is not timed)",
                   "timing": "one cudaEvent pair per graph-launched update on the engine stream, K pairs summed; span_ms_incl_flush = first start to last end with the flushes in it",
                   "span_ms_incl_flush": span_ms, "env_groups": int(os.environ.get("TRL_GROUPS", "0")) or "library default",
                   "presim_s": args.presim, "build": os.environ.get("TRL_VARIANT") or "product"},
        "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": d2h,
                "note": "poli_eval is closed-loop: no per-step host inputs exist; each step reads the counters and all poses/velocities back to host arrays (pinned staging, read-back of step k overlapped with step k+1)"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"], "traffic": ncu_traffic_bytes(), "peak_kind": peak_kind,
                     "kernel": "trl_step_kernel", "launch_ms": launch_s * 1e3,
                     "algorithmic_bytes_per_env_step": ALGO_BYTES_PER_ENV_STEP, "env_steps_per_launch": env_steps_per_launch,
                     "step_kernel_share_of_update": step_ms / (step_ms + dec_ms),
                     "note": "path is FP64-latency bound, not HBM bound (SURVEY §8d): see DESIGN.md"},
        "sim": {"episodes": stats["episodes"], "cycles": stats["cycles"], "avg_dist_m": stats["avg_dist"],
                "merged_over_ranks": world > 1},
    }
    if cpu_val is not None:
        line["cpu_baseline"] = {"value": cpu_val, "unit": "env-steps/s", "cores": cores, "kind": "port", "per_thread": cpu_val / cores,
                                "core_info": core_info, "oracle_build": cpu_build,
                                "sample": f"{cpu_envs} envs x {cpu_updates} outer updates ({cpu_dt:.1f} s) on {cores} pinned threads; "
                                          "restated CPU oracle (reduced-coordinate physics), not Bullet"}
    if cfg4 is not None:
        line["config4"] = cfg4
    if args.scene == "dog_slopes_mixed":
        # the bound that actually binds: thread-level FP64 instructions of the step kernel against the FP64 pipe
        # (64 lanes per SM and clock).  FP64_INST_PER_LAUNCH_ENV = ncu's FP64 thread-instructions of one 4096-env launch / 4096
        # (profiles/ncu_step_kernel_r0x.csv, dog / slopes_mixed only -- other scenes carry no figure); an explanatory number beside
        # the contract's HBM roofline.
        sm_mhz = float((clocks or {}).get("sm_mhz") or 1965.0)
        fp64_peak = 148 * 64 * sm_mhz * 1e6
        fp64_ach = FP64_INST_PER_LAUNCH_ENV * n / launch_s
        line["roofline"]["fp64_pipe"] = {"achieved_ginst_s": fp64_ach / 1e9, "peak_ginst_s": fp64_peak / 1e9, "frac": fp64_ach / fp64_peak,
                                         "thread_inst_per_launch_env": FP64_INST_PER_LAUNCH_ENV, "peak": "148 SMs x 64 FP64 lanes x measured SM clock",
                                         "what": "DFMA + DADD + DMUL thread-instructions (ncu) of the step kernel per launch / live launch time",
                                         "ncu_pipe_fp64_cycles_active_pct": ncu_metric("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                                         "ncu_issue_active_pct": ncu_metric("smsp__issue_active.avg.pct_of_peak_sustained_active")}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
