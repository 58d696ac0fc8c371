"""Throughput of the on-device training loop (rollout + tuple hand-over + K trainer iterations per outer update).
Usage: python tools/train_probe.py [envs] [updates] [iters_per_update]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import deepterrainrl_b200 as trl  # noqa: E402
from deepterrainrl_b200.train import ScenarioTrainMACE, TrainSchedule  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
updates = int(sys.argv[2]) if len(sys.argv) > 2 else 60
k = int(sys.argv[3]) if len(sys.argv) > 3 else 8
pack = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "assets", "dog_slopes_mixed.trlpack")
sched = TrainSchedule(init_exp_rate=0.9, exp_rate=0.2, init_exp_temp=20, exp_temp=0.025, init_exp_base_rate=0.9, exp_base_rate=0.002,
                      trainer_num_anneal_iters=50000, exp_base_anneal_iters=50000)
st = ScenarioTrainMACE(pack, n, schedule=sched, iters_per_update=k,
                       trainer_params=dict(replay_mem_size=500000, num_init_samples=20000, freeze_target_iters=500))
st.Run(90)                                    # fill the replay memory past the init stage, desynchronise the gaits
st.exp.Sync()
c0 = st.trainer.counters()
t0 = time.perf_counter()
st.Run(updates)
st.exp.Sync()
dt = time.perf_counter() - t0
c1 = st.trainer.counters()
# rollout alone, same exploration settings
t1 = time.perf_counter()
for _ in range(updates):
    st.exp.Update(1.0 / 30.0)
st.exp.Sync()
dr = time.perf_counter() - t1
st.exp.ResetTupleBuffer()
# trainer alone
lt0 = st.trainer.KernelLaunches()
t2 = time.perf_counter()
st.trainer.Train(updates * k)
st.exp.Sync()
lt1 = st.trainer.KernelLaunches()
dtr = time.perf_counter() - t2
print(json.dumps({"envs": n, "updates": updates, "iters_per_update": k,
                  "train_loop_env_steps_per_s": n * 20 * updates / dt, "rollout_only_env_steps_per_s": n * 20 * updates / dr,
                  "ms_per_update_with_training": 1e3 * dt / updates, "ms_per_update_rollout": 1e3 * dr / updates,
                  "trainer_ms_per_iter": 1e3 * dtr / (updates * k), "trainer_iters": c1["iter"] - c0["iter"],
                  "actor_iters": c1["actor_iter"] - c0["actor_iter"], "tuples_per_update": (c1["total"] - c0["total"]) / updates,
                  "critic_loss": c1["critic_loss"], "actor_loss": c1["actor_loss"],
                  "trainer_kernel_launches_per_iter": (lt1 - lt0) / (updates * k)}))
