"""Sample-tensor consumers: per-draw histogramdd and HPD across draws.
CPU: the oracle's hpd_vectorized against outputs of the reference's own statutils.hpd_vectorized
(tests/golden/ref_stats_hpd.npz, tools/make_reference_goldens.py).  GPU: libnazb against numpy.histogramdd (the call the
reference itself makes, bflow_jax_maf.py:436-441) — bit-exact counts — and against the reference HPD outputs."""
import os

import numpy as np
import pytest
import torch

from oracle import stats_oracle as so

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_stats_hpd.npz")


def test_hpd_oracle_matches_reference_outputs():
    g = np.load(GOLD)
    for i in range(int(g["n"])):
        out = so.hpd_vectorized(g[f"v_{i}"], float(g[f"alpha_{i}"]))
        assert np.array_equal(out, g[f"hpd_{i}"]), i
    with pytest.raises(ValueError):
        so.hpd_vectorized(np.zeros((3, 2, 2)), 0.0)          # statutils.py:33-34


TN_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_truncnorm.npz")


def test_truncnorm_oracle_matches_reference_outputs():
    """oracle (fp64) vs the reference's own TruncatedNormalTransform executed in fp32 (tools/make_reference_goldens.py)."""
    g = np.load(TN_GOLD)
    for i in range(int(g["n"])):
        y, log_q = so.truncnorm_sample(g[f"x_{i}"], g[f"loc_{i}"], g[f"scale_{i}"], g[f"low_{i}"], g[f"high_{i}"])
        assert np.allclose(y, g[f"y_{i}"], rtol=2e-5, atol=2e-6)
        assert np.allclose(log_q, g[f"log_q_{i}"], rtol=2e-5)
        lo, hi = np.broadcast_to(g[f"low_{i}"], y.shape[1:]), np.broadcast_to(g[f"high_{i}"], y.shape[1:])
        assert (y >= lo - 1e-6).all() and (y <= hi + 1e-6).all()


@pytest.mark.gpu
def test_truncnorm_matches_reference_outputs_and_feeds_importance():
    from naz_b200.stats import truncnorm_sample
    g = np.load(TN_GOLD)
    for i in range(int(g["n"])):
        y, log_q = truncnorm_sample(torch.from_numpy(g[f"x_{i}"]).cuda(), g[f"loc_{i}"], g[f"scale_{i}"], g[f"low_{i}"], g[f"high_{i}"])
        assert np.allclose(y.cpu().numpy(), g[f"y_{i}"], rtol=2e-5, atol=2e-6)          # vs the reference's fp32 outputs
        assert np.allclose(log_q.cpu().numpy(), g[f"log_q_{i}"], rtol=2e-5)
        yo, lqo = so.truncnorm_sample(g[f"x_{i}"], g[f"loc_{i}"], g[f"scale_{i}"], g[f"low_{i}"], g[f"high_{i}"])
        assert np.allclose(y.cpu().numpy(), yo, rtol=2e-5, atol=2e-6) and np.allclose(log_q.cpu().numpy(), lqo, rtol=2e-5)
    # big shape: 256 draws x 51 k parameters (cfg 4's flow), log_q sums in double
    S, P = 256, 51_000
    x = torch.rand((S, P), device="cuda") * 0.96 + 0.02
    y, log_q = truncnorm_sample(x, 0.1, 0.1, -1.0, 1.0)
    yo, lqo = so.truncnorm_sample(x.cpu().numpy(), 0.1, 0.1, -1.0, 1.0)
    assert np.allclose(y.cpu().numpy(), yo, rtol=1e-4, atol=1e-5) and np.allclose(log_q.cpu().numpy(), lqo, rtol=1e-5)


def _edges(rng, D, nb):
    # quantile-like, unequal bin widths
    return [np.concatenate([[-3.0], np.sort(rng.uniform(-2.5, 2.5, size=nb[d] - 1)), [3.0]]) for d in range(D)]


@pytest.mark.gpu
@pytest.mark.parametrize("S,N,D,nb", [(3, 5000, 2, (7, 5)), (5, 20000, 4, (4, 4, 4, 4)), (2, 999, 1, (16,)),
                                      (2, 30000, 3, (32, 32, 16)), (4, 1, 2, (3, 3))])
def test_histogramdd_matches_numpy_bit_exact(S, N, D, nb):
    from naz_b200.stats import histogramdd_draws
    rng = np.random.default_rng(S * 1000 + N)
    edges = _edges(rng, D, nb)
    x = (rng.normal(size=(S, N, D)) * 1.6).astype(np.float32)
    # edge cases: samples exactly on interior edges, on the outer edges, outside, NaN
    k = min(N, 8)
    for d in range(D):
        x[0, :k, d] = np.resize(edges[d], k).astype(np.float32)
    if N > 20:
        x[-1, 10, 0] = np.nan
        x[-1, 11, :] = 3.0
        x[-1, 12, :] = -3.0
        x[-1, 13, 0] = np.float32(3.0000002)
    counts, dens = histogramdd_draws(torch.from_numpy(x).cuda(), edges)
    c_ref, d_ref = so.histogramdd_draws(x, edges)
    assert counts.shape == (S, *nb)
    assert np.array_equal(counts.cpu().numpy(), c_ref)                      # integer work: bit-exact
    ok = np.isfinite(d_ref)
    assert np.allclose(dens.cpu().numpy()[ok], d_ref[ok], rtol=2e-6, atol=0)


@pytest.mark.gpu
def test_hpd_matches_reference_outputs_and_oracle():
    from naz_b200.stats import hpd_draws
    g = np.load(GOLD)
    for i in range(int(g["n"])):
        out = hpd_draws(torch.from_numpy(g[f"v_{i}"]).cuda(), float(g[f"alpha_{i}"]))
        assert np.array_equal(out.cpu().numpy(), g[f"hpd_{i}"]), i          # order statistics: exact
    rng = np.random.default_rng(3)
    v = rng.gamma(2.0, 1.0, size=(4400, 16, 16)).astype(np.float32)        # calibrate_4p: 4400 draws
    out = hpd_draws(torch.from_numpy(v).cuda(), 0.1)
    assert np.array_equal(out.cpu().numpy(), so.hpd_vectorized(v, 0.1))
    with pytest.raises(ValueError):
        hpd_draws(torch.zeros((3, 2, 2)).cuda(), 0.0)


@pytest.mark.gpu
def test_sample_then_histogram_then_hpd_pipeline():
    """sample_draws -> per-draw histogram -> HPD band, all on device (the calibrate() data flow)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import engine_for, make_case
    from naz_b200.stats import histogramdd_draws, hpd_draws
    S, N = 24, 4000
    spec, draws, _, rng = make_case("maf", 2, 2, [32, 32], 3, S, seed=8, scale=0.1)
    eng = engine_for(spec, draws)
    z = torch.from_numpy(rng.normal(size=(N, 2)).astype(np.float32))
    x = eng.forward(z, torch.tensor([0.3, 0.6]))                            # [S, N, 2] on device
    xs = x.cpu().numpy()
    edges = [np.quantile(xs[0][:, d], np.linspace(0, 1, 6)) for d in range(2)]
    counts, dens = histogramdd_draws(x, edges)
    c_ref, d_ref = so.histogramdd_draws(xs, edges)
    assert np.array_equal(counts.cpu().numpy(), c_ref)
    band = hpd_draws(dens, 0.32)
    assert np.array_equal(band.cpu().numpy(), so.hpd_vectorized(dens.cpu().numpy(), 0.32))


@pytest.mark.gpu
@pytest.mark.parametrize("S", [9000, 20000, 32768])
def test_hpd_many_draws(S):
    """More than 8192 draws: fewer columns per CTA so the sort still fits shared memory (2 up to 16384 draws, 1 up to 32768),
    exact against the restatement of statutils.hpd_vectorized."""
    from naz_b200.stats import hpd_draws
    rng = np.random.default_rng(S)
    v = rng.gamma(2.0, 1.0, size=(S, 7)).astype(np.float32)
    got = hpd_draws(torch.from_numpy(v).cuda(), 0.1).cpu().numpy()
    want = so.hpd_vectorized(v, 0.1)
    assert np.array_equal(got, want.astype(np.float32))
