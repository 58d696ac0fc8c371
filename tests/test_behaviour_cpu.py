"""Behavioural acceptance of the physics stand-in: Bullet is absent, so the contact / constraint model cannot be pinned
numerically against the reference; what can be tracked is whether the reference's SHIPPED policies, trained against Bullet, still
work on it.  64 terrain seeds per scene on the CPU oracle (the CUDA path equals the oracle to ~1e-11, tests/test_gpu_parity.py);
the thresholds sit just under the figures measured when the table in DESIGN §3 was written, so a contact-model regression fails here.
tools/behaviour_stats.py prints the full statistic."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def _stats(scene, seconds):
    import behaviour_stats
    r = behaviour_stats.run(scene, seeds=64, seconds=seconds, threads=min(8, os.cpu_count() or 1), fast=False)
    r["dist"] = np.array(r["dist"])
    return r


def test_dog_slopes_mixed_policy_runs():
    r = _stats("dog_slopes_mixed", 20.0)                     # measured: 90.6 % without a fall, mean 79.6 m, 85.9 % beyond 80 m
    assert r["no_fall_frac"] >= 0.85, r["no_fall_frac"]
    assert r["mean_first_episode_dist"] >= 75.0, r["mean_first_episode_dist"]
    assert (r["dist"] > 80.0).mean() >= 0.80
    assert 3.8 <= r["dist_no_fall_mean"] / 20.0 <= 4.4       # the controller's target speed is 4 m/s


def test_raptor_narrow_gaps_policy_runs():
    r = _stats("raptor_narrow_gaps", 10.0)                   # measured: 53.1 % without a fall, mean 32.8 m, surviving runs at 4.1 m/s
    assert r["no_fall_frac"] >= 0.45, r["no_fall_frac"]
    assert r["mean_first_episode_dist"] >= 28.0, r["mean_first_episode_dist"]
    assert 3.8 <= r["dist_no_fall_mean"] / 10.0 <= 4.6


def test_goat_cliffs_known_gap():
    """KNOWN GAP, tracked: on this contact model the shipped goat policy flips backwards at the first cliff (every episode ends
    after ~3 m); on flat ground the same controller runs indefinitely.  The test pins the current state so that a change of the
    model shows up either way."""
    r = _stats("goat_cliffs", 5.0)
    assert r["no_fall_frac"] <= 0.25
    assert 2.0 <= r["mean_first_episode_dist"] <= 8.0
