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
    summary of the same workload (profiles/ncu_step_kernel_r01_final.csv); None if the summary is absent."""
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


def cpu_reference(num_envs, seconds_target, threads):
    """Times the CPU oracle (restated reference controller + this project's reduced-coordinate physics, f64) with
    thread-per-env-slice like cOptScenarioPoliEval::Run (optimizer/scenarios/OptScenarioPoliEval.cpp:72-110)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from pyoracle import Oracle
    o = Oracle(PACK, num_envs, 0)
    o.update(DT, threads)   # warm-up
    t0 = time.perf_counter()
    updates = 0
    while time.perf_counter() - t0 < seconds_target:
        o.update(DT, threads)
        updates += 1
    dt = time.perf_counter() - t0
    steps = updates * ENV_STEPS_PER_UPDATE * num_envs
    return steps / dt, dt, updates


def run_reference_arm(args, rank):
    cores = os.cpu_count() or 1
    if rank != 0:
        return
    envs = 4 * cores
    per_step_s = 0.0
    # warm-up + K steps, each step a bounded sample: `envs` environments advanced by one outer update
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from pyoracle import Oracle
    o = Oracle(PACK, envs, 0)
    for _ in range(args.warmup):
        o.update(DT, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.update(DT, cores)
    el = time.perf_counter() - t0
    per_step_s = el / args.steps
    value = envs * ENV_STEPS_PER_UPDATE / per_step_s
    line = {
        "impl": "reference", "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_per_gpu": args.envs},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{envs} envs x {args.steps} outer updates of 20 env-steps, thread-per-env-slice on {cores} threads; "
                                   "restated CPU oracle (reduced-coordinate physics), not Bullet"},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


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
    presim_updates = int(round(args.presim / DT))
    if presim_updates > 0:
        sc.BenchUpdates(presim_updates, DT, flush_l2=False)
    sc.BenchUpdates(max(args.warmup, 3), DT, flush_l2=True)
    launches0 = sc.KernelLaunches()

    # ---- timed: exactly K steps, device-timed on the engine's stream, max over ranks
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.05)
    barrier()
    t_beg = time.perf_counter()
    ms = sc.BenchUpdates(args.steps, DT, flush_l2=True)
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
    achieved = ALGO_BYTES_PER_ENV_STEP * n / launch_s / 1e9
    stats = sc._stats()

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cores = os.cpu_count() or 1
    cpu_envs = 4 * cores
    cpu_val, cpu_dt, cpu_updates = cpu_reference(cpu_envs, args.cpu_seconds, cores) if world == 1 else (None, 0, 0)

    line = {
        "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_per_gpu": n,
                   "env_steps_per_step": ENV_STEPS_PER_UPDATE * n, "sim_substeps": 5, "parallelism": f"env-shard x{world}",
                   "l2": "flushed between steps (256 MiB memset on the engine stream)",
                   "timing": "cudaEvent on the engine stream around K graph-launched updates",
                   "presim_s": args.presim, "build": os.environ.get("TRL_VARIANT") or "product"},
        "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": d2h,
                "note": "poli_eval is closed-loop: no per-step host inputs exist; each step reads the counters and all poses/velocities back to host arrays (pinned staging, read-back of step k overlapped with step k+1)"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"], "traffic": ncu_traffic_bytes(), "peak_kind": peak_kind,
                     "kernel": "trl_step_kernel", "launch_ms": launch_s * 1e3,
                     "algorithmic_bytes_per_env_step": ALGO_BYTES_PER_ENV_STEP,
                     "step_kernel_share_of_update": step_ms / (step_ms + dec_ms),
                     "note": "path is FP64-latency bound, not HBM bound (SURVEY §8d): see DESIGN.md"},
        "sim": {"episodes": stats["episodes"], "cycles": stats["cycles"], "avg_dist_m": stats["avg_dist"]},
    }
    if cpu_val is not None:
        line["cpu_baseline"] = {"value": cpu_val, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                "sample": f"{cpu_envs} envs x {cpu_updates} outer updates ({cpu_dt:.1f} s) on {cores} threads; "
                                          "restated CPU oracle (reduced-coordinate physics), not Bullet"}
    try:
        # the bound that actually binds: thread-level FP64 instructions of the step kernel against the FP64 pipe
        # (64 lanes per SM and clock).  138 k per env-step = ncu's 566 M FP64 thread-instructions of one 4096-env launch
        # (profiles/ncu_step_kernel_r01_final.csv, dog / slopes_mixed); an explanatory figure beside the contract's HBM roofline.
        sm_mhz = float((clocks or {}).get("sm_mhz") or 1965.0)
        fp64_peak = 148 * 64 * sm_mhz * 1e6
        fp64_ach = 138.0e3 * n / launch_s
        line["roofline"]["fp64_pipe"] = {"achieved_ginst_s": fp64_ach / 1e9, "peak_ginst_s": fp64_peak / 1e9, "frac": fp64_ach / fp64_peak,
                                         "thread_inst_per_env_step": 138.0e3, "peak": "148 SMs x 64 FP64 lanes x measured SM clock"}
    except Exception:
        pass
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
