"""CPU: size-independent properties of the oracle restatement (hypothesis): the ITM sampling rule, the bit-reproducible
exponential behind the hard-negative sampler, the hard sampler itself, clip_loss symmetries and the evaluation scores."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import restatement as R

SET = dict(max_examples=40, deadline=None)


@settings(**SET)
@given(st.integers(1, 300), st.integers(0, 2 ** 31 - 1))
def test_uniform_rule_invariants(B, seed):
    rs = np.random.RandomState(seed)
    u_coin, u_pick = rs.rand(B).astype(np.float32), rs.rand(B).astype(np.float32)
    lbl, src = R.itm_sample_uniform(u_coin, u_pick)
    i = np.arange(B)
    assert set(np.unique(lbl)) <= {0, 1} and src.min() >= 0 and src.max() < B
    if B == 1:
        assert lbl.tolist() == [1] and src.tolist() == [0]        # mm_late.py:410-411: a single sample is always a match
        return
    assert np.array_equal(lbl == 0, u_coin < 0.5)                  # swap iff the coin says so
    assert np.all(src[lbl == 1] == i[lbl == 1])                    # kept rows point at themselves
    assert np.all(src[lbl == 0] != i[lbl == 0])                    # swapped rows never pick themselves (:400)


@settings(**SET)
@given(st.integers(2, 200), st.integers(0, 2 ** 31 - 1))
def test_stream_decisions_round_trip_through_uniforms(B, seed):
    """numpy-stream decisions -> uniforms -> the uniform rule reproduces them: how the CLI replays the reference's stream on
    the device kernel."""
    swap, src = R.itm_decisions_from_stream(B, np.random.RandomState(seed))
    u_coin, u_pick = R.uniforms_from_decisions(swap, src)
    lbl, src2 = R.itm_sample_uniform(u_coin, u_pick)
    assert np.array_equal(lbl, (~swap).astype(lbl.dtype)) and np.array_equal(src2, src)


@settings(**SET)
@given(st.lists(st.floats(-80.0, 0.0, allow_nan=False, width=32), min_size=1, max_size=64))
def test_det_exp_is_accurate_and_monotone(xs):
    x = np.sort(np.asarray(xs, dtype=np.float32))
    y = R.det_exp_f32(x)
    ref = np.exp(x.astype(np.float64))
    assert np.all(np.abs(y - ref) <= 4e-7 * ref + 1e-38)
    assert np.all(np.diff(y.astype(np.float64)) >= -1e-7 * ref[1:])


@settings(max_examples=15, deadline=None)
@given(st.integers(2, 48), st.integers(0, 2 ** 31 - 1))
def test_hard_sampler_invariants(B, seed):
    rs = np.random.RandomState(seed)
    S = (rs.randn(B, B) * 3).astype(np.float32)
    u_coin, u_pick = rs.rand(B).astype(np.float32), rs.rand(B).astype(np.float32)
    lbl, src = R.itm_sample_hard(S, u_coin, u_pick)
    lbl2, src2 = R.itm_sample_hard(S.copy(), u_coin.copy(), u_pick.copy())
    i = np.arange(B)
    assert np.array_equal(lbl, lbl2) and np.array_equal(src, src2)             # deterministic
    assert np.array_equal(lbl == 0, u_coin < 0.5)
    assert np.all(src[lbl == 1] == i[lbl == 1]) and np.all(src[lbl == 0] != i[lbl == 0])
    # a row whose similarity to one other row dominates by > 40 nats picks that row whatever the uniform says
    S2 = S.copy()
    tgt = (i + 1) % B
    S2[i, tgt] += 60.0
    lbl3, src3 = R.itm_sample_hard(S2, u_coin, u_pick)
    assert np.all(src3[lbl3 == 0] == tgt[lbl3 == 0])


@settings(**SET)
@given(st.integers(1, 40), st.integers(0, 2 ** 31 - 1))
def test_clip_loss_symmetries(B, seed):
    g = torch.Generator().manual_seed(seed)
    S = torch.randn(B, B, generator=g, dtype=torch.float64) * 3
    base = R.clip_loss(S)
    assert torch.allclose(R.clip_loss(S.t()), base)                                # text <-> image
    perm = torch.randperm(B, generator=g)
    assert torch.allclose(R.clip_loss(S[perm][:, perm]), base)                     # relabelling the batch
    assert torch.allclose(R.clip_loss(S + 7.5), base)                              # softmax shift invariance
    assert float(base) >= -1e-12


@settings(**SET)
@given(st.integers(2, 9), st.integers(1, 400), st.integers(0, 2 ** 31 - 1))
def test_metric_invariants(C, n, seed):
    rs = np.random.RandomState(seed)
    t, p = rs.randint(0, C, n), rs.randint(0, C, n)
    conf = R.confusion_matrix(p, t, C)
    assert conf.sum() == n and np.array_equal(conf.sum(axis=1), np.bincount(t, minlength=C))
    m = R.metrics_from_confusion(conf)
    assert all(0.0 <= v <= 1.0 + 1e-12 for v in m.values())
    assert abs(m["recall_weighted"] - np.mean(p == t)) < 1e-12                     # weighted recall == accuracy
    perm = rs.permutation(C)                                                        # renaming the classes changes nothing
    m2 = R.metrics_from_confusion(R.confusion_matrix(perm[p], perm[t], C))
    assert all(abs(m[k] - m2[k]) < 1e-12 for k in m)
