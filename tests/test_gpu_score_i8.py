"""The tcgen05 scoring path (int8 digit planes, exact integer accumulation; csrc/score.cu, `score_i8_kernel`)
against the CPU oracle and against the FP64 DMMA path it replaces from n_pad = 512.
Reference: train/log_gaus_likelihood.py:128-218, bayes_opt/helpers.py:51-68,151-171."""
import math

import numpy as np
import pytest
import torch

from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def fitted_engine(eng, N, d, s2, seed, kvec_zero_frac=0.0):
    rs = np.random.RandomState(seed)
    theta = rs.randn(N, d)
    y = np.sin(theta.sum(1)) + 0.1 * rs.randn(N)
    e = eng.GPEngine(d)
    e.params.set("log_sigma", 0.5 * math.log(s2))
    e.params.set("log_bw", 0.3 * rs.randn(d))
    e.params.set("log_scale", 0.4)
    e.params.set("mean", 0.2)
    e.set_observations(theta, y)
    e.build_gram()
    e.factorize(need_inverse=True, need_kinv=False)
    e.likelihood()
    e.read_loss()
    kv = 0.3 + 0.7 * rs.rand(N)
    if kvec_zero_frac:
        kv[rs.rand(N) < kvec_zero_frac] = 0.0
    e.kvec.zero_()
    e.kvec[:N] = torch.from_numpy(kv).to(e.dev)
    return e, theta, y, kv, rs


def both_paths(e, cand, **kw):
    out = {}
    for on in (False, True):
        e.use_score_i8, e.score_i8_min_npad = on, 128
        r = e.score(cand, **kw)
        torch.cuda.synchronize()
        out[on] = {k: (v.clone().cpu().numpy() if v is not None else None) for k, v in r.items()}
    return out[False], out[True]


@pytest.mark.parametrize("N,d,M,s2", [(50, 1, 77, 1e-2), (128, 3, 128, 1e-2), (333, 5, 777, 1e-5),
                                      (640, 2, 1500, 1e-5), (1000, 16, 2049, 1e-2), (1537, 8, 19000, 1e-5)])
@pytest.mark.parametrize("acq", ["ei", "ucb"])
def test_i8_matches_oracle_and_dmma(cuda_engine_mod, N, d, M, s2, acq):
    eng = cuda_engine_mod
    e, theta, y, kv, rs = fitted_engine(eng, N, d, s2, seed=N + d)
    cand = rs.randn(M, d)
    kw = dict(acq=acq, y_max=float(y.max()), xi=0.01, kappa=2.5)
    dm, i8 = both_paths(e, cand, **kw)
    p = {k: e.params.get(k) for k in e.params.offsets}
    mean, var = O.predict(O.NP, p, "dist", O.TrainBatch(h_params=theta, obs=y[:, None]), cand,
                          k_dist=np.ones((N, N)), k_dist_pred=kv[:, None])
    std = O.std_from_var(var)
    a = O.acq_ei(mean, std, float(y.max()), 0.01) if acq == "ei" else O.acq_ucb(mean, std, 2.5)
    scale = math.exp(float(p["log_scale"][0]))
    # the tolerance north_star states: 1e-8 relative (variance relative to the prior scale)
    assert rel(i8["mean"], mean[:, 0]) < 1e-8
    assert float(np.max(np.abs(i8["std"] ** 2 - np.maximum(var[:, 0], 1e-32)))) < 1e-8 * scale
    assert int(i8["best_idx"][0]) == int(np.argmax(a))
    # ... and the two device paths agree far below that
    assert rel(i8["mean"], dm["mean"]) < 1e-10
    assert float(np.max(np.abs(i8["std"] ** 2 - dm["std"] ** 2))) < 1e-10 * scale
    assert int(i8["best_idx"][0]) == int(dm["best_idx"][0])


def test_i8_pad_rows_zero_kvec_and_scaler(cuda_engine_mod):
    """Rows with kvec = 0 (other tasks under kernel 'params'-style masks), candidates through the frozen
    StandardScaler, index offset of a candidate shard."""
    eng = cuda_engine_mod
    e, theta, y, kv, rs = fitted_engine(eng, 700, 4, 1e-3, seed=5, kvec_zero_frac=0.3)
    M = 4000
    raw = rs.randn(M, 4) * 3.0 + 1.0
    shift = torch.from_numpy(np.full(4, 1.0)).to(e.dev)
    scl = torch.from_numpy(np.full(4, 3.0)).to(e.dev)
    kw = dict(acq="ei", y_max=float(y.max()), xi=0.0, scaler=(shift, scl), index_offset=12345)
    dm, i8 = both_paths(e, torch.from_numpy(raw).to(e.dev), **kw)
    assert rel(i8["mean"], dm["mean"]) < 1e-10
    assert int(i8["best_idx"][0]) == int(dm["best_idx"][0]) >= 12345
    p = {k: e.params.get(k) for k in e.params.offsets}
    mean, _ = O.predict(O.NP, p, "dist", O.TrainBatch(h_params=theta, obs=y[:, None]), (raw - 1.0) / 3.0,
                        k_dist=np.ones((700, 700)), k_dist_pred=kv[:, None])
    assert rel(i8["mean"], mean[:, 0]) < 1e-8


def test_i8_digit_planes_reconstruct_w(cuda_engine_mod):
    """The digit planes are an exact 55-bit fixed-point image of W: sum_s d_s 2^(48-8s) 2^(e_i-54) equals W to
    2^-55 of the row maximum; digits stay in [-128, 127], the leading one in [-65, 65]."""
    eng = cuda_engine_mod
    e, theta, y, kv, rs = fitted_engine(eng, 300, 3, 1e-4, seed=9)
    e.use_score_i8, e.score_i8_min_npad = True, 128
    e.score(rs.randn(10, 3), acq="ei", y_max=0.0, xi=0.0)
    torch.cuda.synchronize()
    W8, rowscale = e._w8
    n = e.n_pad
    assert W8.shape[0] == 7
    W = np.tril(e.W.cpu().numpy()[:n, :n])
    dg = W8.cpu().numpy().astype(np.int64)
    assert dg.min() >= -128 and dg.max() <= 127 and np.abs(dg[0]).max() <= 65
    fixed = sum(dg[s] * (1 << (48 - 8 * s)) for s in range(7))
    rs_ = rowscale.cpu().numpy()
    back = fixed.astype(np.float64) * (rs_[:, None] * 2.0 ** -54)
    rowmax = np.abs(W).max(axis=1)
    assert np.all(rs_ >= rowmax) and np.all(rs_ <= 2.0 * np.maximum(rowmax, 1e-300) + (rowmax == 0))
    assert np.max(np.abs(back - W) / rs_[:, None]) <= 2.0 ** -55


def test_i8_refreshes_after_refit(cuda_engine_mod):
    """W changes with every factorisation: the digit planes must follow (engine.W_version)."""
    eng = cuda_engine_mod
    e, theta, y, kv, rs = fitted_engine(eng, 600, 2, 1e-2, seed=3)
    cand = rs.randn(1000, 2)
    kw = dict(acq="ei", y_max=float(y.max()), xi=0.01)
    both_paths(e, cand, **kw)
    e.params.set("log_sigma", 0.5 * math.log(0.3))
    e.build_gram()
    e.factorize(need_inverse=True, need_kinv=False)
    e.likelihood()
    dm, i8 = both_paths(e, cand, **kw)
    assert rel(i8["mean"], dm["mean"]) < 1e-10
    assert float(np.max(np.abs(i8["std"] - dm["std"]))) < 1e-9
