"""CPU tests of the MACE trainer oracle (oracle/trainer.h): the backward pass against finite differences, the SGD step against
a numpy restatement of Caffe's update, replay-buffer bookkeeping, and that a few solver steps reduce the loss."""
import os

import numpy as np
import pytest

from pyoracle import OracleTrainer


def _problem(tr, B, seed):
    rng = np.random.default_rng(seed)
    S, no = tr.n_in, tr.n_out
    in_off, in_scale = tr.get("in_off"), tr.get("in_scale")
    # inputs around the data the shipped scale file was fitted to
    X = rng.normal(size=(B, S)) / np.where(in_scale == 0, 1.0, in_scale) - in_off
    Y = tr.eval_batch(X)
    Y = Y + 0.3 * rng.normal(size=Y.shape) / tr.get("out_scale")
    return X, Y


def test_backward_matches_finite_differences(assets):
    tr = OracleTrainer(os.path.join(assets, "dog_slopes_mixed.trlpack"))
    assert tr.num_params == 570474 and tr.n_in == 283 and tr.n_out == 90
    X, Y = _problem(tr, 4, 0)
    loss, g = tr.loss_grad(X, Y)
    theta = tr.get("theta")
    rng = np.random.default_rng(1)
    # probe the largest-gradient parameter of every blob family plus random ones
    idx = list(np.argsort(-np.abs(g))[:12]) + list(rng.integers(0, theta.size, 12))
    for i in idx:
        h = 1e-6 * max(1.0, abs(theta[i]))
        tp = theta.copy(); tp[i] += h
        tm = theta.copy(); tm[i] -= h
        tr.set_theta(tp); lp, _ = tr.loss_grad(X, Y, False)
        tr.set_theta(tm); lm, _ = tr.loss_grad(X, Y, False)
        fd = (lp - lm) / (2 * h)
        assert abs(fd - g[i]) <= 1e-6 * max(1.0, abs(g[i])) + 2e-7, (i, fd, g[i])
    tr.set_theta(theta)


def _blob_table(n_char=83, frag=29, n_frags=3):
    sizes = [16 * 8, 16, 32 * 16 * 4, 32, 32 * 32 * 4, 32, 64 * 32 * 187, 64, 256 * (64 + n_char), 256]
    for h in range(4):
        nout = n_frags if h == 0 else frag
        sizes += [128 * 256, 128, nout * 128, nout]
    return sizes


def test_sgd_step_is_caffe_update(assets):
    """One critic step from a hand-made replay memory: theta' = theta - (lr*(g + wd*theta) + mom*hist) with the per-blob
    lr_mult / decay_mult of dog_mace3_train.prototxt; history carried into the second step."""
    pack = os.path.join(assets, "dog_slopes_mixed.trlpack")
    tr = OracleTrainer(pack, num_init_samples=32, init_input_offset_scale=0, replay_cap=64, seed=3)
    S, A, W = tr.n_in, 30, tr.W
    rng = np.random.default_rng(5)
    in_off, in_scale = tr.get("in_off"), tr.get("in_scale")
    rows = np.zeros((32, W))
    rows[:, 0] = rng.uniform(0, 1, 32)
    for k in (1, 1 + S + A):
        rows[:, k:k + S] = rng.normal(size=(32, S)) / np.where(in_scale == 0, 1.0, in_scale) - in_off
    rows[:, 1 + S] = rng.integers(0, 3, 32)
    rows[:, 2 + S:1 + S + A] = rng.normal(size=(32, A - 1))
    flags = np.zeros(32, np.uint32); flags[:4] = 1          # a few failures
    tr.add_tuples(rows, flags)
    c = tr.counters()
    assert c["num"] == 32 and c["critic"] == 32 and c["actor"] == 0
    theta0 = tr.get("theta")
    Ytar_all = None
    tr.train()
    theta1, hist1 = tr.get("theta"), tr.get("history")
    assert tr.counters()["iter"] == 1 and tr.counters()["stage"] == 1
    np.testing.assert_allclose(theta0 - theta1, hist1, rtol=0, atol=1e-15)
    # independent restatement of BuildProblemY + the Caffe update on the batch the trainer drew
    ids = tr.lists("last_critic")
    assert ids.size == 32
    r, fl = tr.rows(ids)
    r = r.astype(np.float64)
    tr.set_theta(theta0)                                       # evaluate the problem at the pre-step weights
    X = r[:, 1:1 + S]
    Yt = tr.eval_batch(r[:, 1 + S + A:], target=True)          # target net == initial net here
    q = r[:, 0] * (1 - 0.9) + np.where(fl & 1, 0.0, 0.9 * Yt[:, :3].max(1))
    Y = tr.eval_batch(X)
    Y[np.arange(32), r[:, 1 + S].astype(int)] = q
    loss, g = tr.loss_grad(X, Y)
    sizes = _blob_table()
    off = np.concatenate([[0], np.cumsum(sizes)])
    assert off[-1] == theta0.size
    expect = np.zeros_like(theta0)
    for b in range(26):
        sl = slice(off[b], off[b + 1])
        lr = 1e-3 * (2.0 if b & 1 else 1.0)
        decay = 5e-4 * ((1.0 if b < 6 else 0.0) if b & 1 else 1.0)
        expect[sl] = lr * (g[sl] + decay * theta0[sl])
    np.testing.assert_allclose(hist1, expect, rtol=1e-9, atol=1e-16)
    assert abs(loss - tr.losses()[0]) <= 1e-12 * max(1.0, loss)
    tr.set_theta(theta1)
    tr.train()
    theta2, hist2 = tr.get("theta"), tr.get("history")
    np.testing.assert_allclose(theta1 - theta2, hist2, rtol=0, atol=1e-15)
    assert np.linalg.norm(hist2 - 0.9 * hist1) < np.linalg.norm(hist2)       # momentum term present


def test_replay_buffers_follow_flags(assets):
    tr = OracleTrainer(os.path.join(assets, "dog_slopes_mixed.trlpack"), replay_cap=8, num_init_samples=1000)
    W = tr.W
    rows = np.ones((12, W)) * 0.25
    flags = np.array([0, 4, 0, 6, 0, 4, 0, 0, 4, 0, 2, 0], np.uint32)       # bit 2 = actor exploration
    rows[5, 3] = np.nan                                                      # rejected by CheckTuple
    tr.add_tuples(rows, flags)
    c = tr.counters()
    assert c["total"] == 11 and c["num"] == 8 and c["head"] == 11 % 8
    # slots after the wrap: accepted tuples 0..10 (input 5 skipped) land in slot k % 8
    accepted = [i for i in range(12) if i != 5]
    slot_flag = {}
    for k, i in enumerate(accepted):
        slot_flag[k % 8] = int(flags[i])
    actor = sorted(s for s, f in slot_flag.items() if f & 4)
    critic = sorted(s for s, f in slot_flag.items() if not (f & 4))
    assert sorted(tr.lists("actor").tolist()) == actor
    assert sorted(tr.lists("critic").tolist()) == critic
    assert c["stage"] == 0
    tr.train()
    assert tr.counters()["iter"] == 0                                        # still collecting initial samples


def test_training_reduces_loss_and_updates_target(assets):
    pack = os.path.join(assets, "dog_slopes_mixed.trlpack")
    tr = OracleTrainer(pack, num_init_samples=64, init_input_offset_scale=1, replay_cap=256, freeze_target_iters=2, seed=11)
    S, A, W = tr.n_in, 30, tr.W
    rng = np.random.default_rng(2)
    in_off, in_scale = tr.get("in_off"), tr.get("in_scale")
    rows = np.zeros((96, W))
    rows[:, 0] = rng.uniform(0, 1, 96)
    for k in (1, 1 + S + A):
        rows[:, k:k + S] = rng.normal(size=(96, S)) / np.where(in_scale == 0, 1.0, in_scale) - in_off
    rows[:, 1 + S] = rng.integers(0, 3, 96)
    rows[:, 2 + S:1 + S + A] = rng.normal(size=(96, A - 1)) * 0.1
    flags = np.where(rng.uniform(size=96) < 0.4, 4, 0).astype(np.uint32)
    tr.add_tuples(rows, flags)
    t0 = tr.get("target")
    np.testing.assert_array_equal(t0, tr.get("theta"))
    # offsets refitted from the replay memory at the stage switch
    tr.train()
    mean = rows[:, 1:1 + S].astype(np.float32).astype(np.float64).mean(0)
    np.testing.assert_allclose(tr.get("in_off"), -mean, rtol=1e-12, atol=1e-12)
    for _ in range(4):
        tr.train()
    c = tr.counters()
    assert c["iter"] == 5 and c["stage"] == 1
    assert not np.array_equal(tr.get("theta"), t0)
    assert not np.array_equal(tr.get("target"), t0)                          # copied at iter 2 / 4
    assert np.isfinite(tr.losses()).all()
