"""The fit's two inverse products on the tcgen05 digit-plane kernel (csrc/score.cu): the top levels of W = L^-1
(`distbo_trtri_i8_f64`) and K^-1 = W^T W (`distbo_lauum_i8_f64`), against the FP64 DMMA route and dense algebra,
and through one full loss-and-gradient evaluation against torch autograd.
Reference: tf.cholesky / tf.matrix_triangular_solve + autodiff, train/log_gaus_likelihood.py:119-125, train/train.py:31-40."""
import math

import numpy as np
import pytest
import torch

from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu


def make(eng, N, d, s2, seed, trtri, lauum, min_rows=128):
    rs = np.random.RandomState(seed)
    theta = rs.randn(N, d)
    y = np.sin(theta.sum(1)) + 0.1 * rs.randn(N)
    e = eng.GPEngine(d)
    e.use_trtri_i8, e.trtri_i8_min_rows = trtri, min_rows
    e.use_lauum_i8, e.lauum_i8_min_npad = lauum, 128
    e.params.set("log_sigma", 0.5 * math.log(s2))
    e.params.set("log_bw", 0.2 * rs.randn(d))
    e.set_observations(theta, y)
    return e, theta, y


# N = 700 -> 6 diagonal blocks (uneven halving 3 + 3 -> 1 + 2), 1100 -> 9, 2500 -> 20, 4096 -> 32
@pytest.mark.parametrize("N,s2,min_rows", [(300, 1e-2, 128), (700, 1e-5, 128), (1100, 1e-2, 256), (2500, 1e-5, 512),
                                           (4096, 1e-2, 1024)])
def test_inverse_products_match_dmma_and_dense_algebra(cuda_engine_mod, N, s2, min_rows):
    eng = cuda_engine_mod
    out = {}
    for on in (False, True):
        e, theta, y = make(eng, N, 3, s2, seed=N, trtri=on, lauum=on, min_rows=min_rows)
        e.build_gram()
        e.factorize(need_inverse=True, need_kinv=True)
        e.likelihood()
        nll = e.read_loss()
        assert (e._trtri_i8 is not None) == on and (e._lauum_i8 is not None) == on
        out[on] = (torch.tril(e.W[:N, :N]).clone(), torch.tril(e.Tm[:N, :N]).clone(), nll,
                   torch.tril(e.K[:N, :N]).clone())
        e.close()
    W0, K0, nll0, L0 = out[False]
    W1, K1, nll1, L1 = out[True]
    assert torch.equal(L0, L1)                                   # POTRF itself is untouched
    eye = torch.eye(N, dtype=torch.float64, device=W1.device)
    assert float((L1 @ W1 - eye).abs().max()) < 1e-11            # W really is the inverse of L
    assert float((W1 - W0).abs().max() / W0.abs().max()) < 1e-11
    kref = torch.tril(W1.T @ W1)
    assert float((K1 - kref).abs().max() / kref.abs().max()) < 1e-12
    assert float((K1 - K0).abs().max() / K0.abs().max()) < 1e-10
    assert abs(nll1 - nll0) <= 1e-10 * abs(nll0)


def test_loss_and_gradient_through_the_int8_inverse(cuda_engine_mod):
    """One evaluation of the NLL and its gradient with every inverse product on the digit-plane kernel, against
    torch autograd through the Cholesky (the reference's route), 1e-8."""
    eng = cuda_engine_mod
    N, d = 900, 4
    e, theta, y = make(eng, N, d, 1e-3, seed=11, trtri=True, lauum=True, min_rows=128)
    e.build_gram()
    e.factorize(need_inverse=True, need_kinv=True)
    e.likelihood()
    nll = e.read_loss()
    e.gram_gradient()
    p = {k: e.params.get(k) for k in e.params.offsets}
    B = O.TH()
    tp = {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=True) for k, v in p.items()}
    k = torch.exp(tp["log_scale"]) * O.matern32_kernel(B, torch.from_numpy(theta), torch.from_numpy(theta),
                                                      torch.exp(tp["log_bw"]))
    ks = k + (torch.exp(2.0 * tp["log_sigma"]) + O.JITTER) * torch.eye(N, dtype=torch.float64)
    Lref = torch.linalg.cholesky(ks)
    loss = -O.multivariate_normal_logpdf(B, torch.from_numpy(y[:, None]),
                                         tp["mean"] * torch.ones(N, 1, dtype=torch.float64), Lref).reshape(())
    loss.backward()
    assert abs(nll - float(loss.detach())) <= 1e-9 * abs(float(loss.detach()))
    for name in ["log_bw", "log_scale", "mean", "log_sigma"]:
        got = np.asarray(e.params.get_grad(name), dtype=np.float64).reshape(-1)
        ref = tp[name].grad.numpy().reshape(-1)
        assert float(np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-300)) < 1e-8, name
    e.close()
