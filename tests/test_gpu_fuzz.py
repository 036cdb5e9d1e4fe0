"""Randomised shape sweep on the GPU: ragged task sizes, N not a multiple of the 128 tile, 1 <= d <= 8, embedding
widths 1..40, candidate counts around the tile / chunk edges.  Checker: the CPU oracle (N is small enough for
seconds) for K, L, NLL, every gradient, the posterior and the arg-max."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import gp_oracle as O  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


CASES = [(137, 1, [137], 0, 129), (255, 8, [1, 200, 54], 1, 127), (384, 3, [128, 128, 128], 40, 128),
         (513, 5, [13] * 39 + [6], 7, 1), (640, 2, [300, 1, 1, 338], 23, 18944 + 1), (1100, 7, [550, 550], 12, 257),
         (1260, 2, [30] * 42, 100, 300),      # Parkinson scale: 42 patients x 30 evaluations, joint embedding P = 100
         (5, 2, [2, 3], 3, 3)]


@pytest.mark.parametrize("N,d,sizes,P,M", CASES)
def test_random_shape(cuda_engine_mod, N, d, sizes, P, M):
    eng = cuda_engine_mod
    rs = np.random.RandomState(N + d)
    theta = rs.randn(N, d)
    y = rs.randn(N)
    T = len(sizes)
    e = eng.GPEngine(d, n_embed=P)
    var = 10 ** rs.uniform(-3, -1)
    e.params.set("log_sigma", 0.5 * math.log(var))
    e.params.set("log_scale", [rs.randn() * 0.3])
    e.params.set("mean", [rs.randn() * 0.3])
    e.params.set("log_bw", rs.randn(d) * 0.3)
    E = None
    if P:
        e.params.set("log_bw_embed", rs.randn(P) * 0.3)
        E = rs.randn(T, P) * 0.4
        e.E = torch.from_numpy(E).to(e.dev)
    e.set_observations(theta, y, sizes if P else None)
    KE = e.task_gram(e.E, T) if P else None
    e.build_gram(KE)
    e.factorize(need_inverse=True, need_kinv=True)
    e.likelihood()
    nll = e.read_loss()
    e.gram_gradient()
    dE = e.task_gram_backward(e.E, T).cpu().numpy() if P else None

    p = {k: e.params.get(k) for k in e.params.offsets}
    B = O.TH()
    tp = {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=True) for k, v in p.items()}
    k = O.matern32_kernel(B, torch.from_numpy(theta), torch.from_numpy(theta), torch.exp(tp["log_bw"]))
    if P:
        Et = torch.tensor(E, requires_grad=True)
        emb = B.repeat_rows(Et, sizes)
        k = k * O.matern32_kernel(B, emb, emb, torch.exp(tp["log_bw_embed"]))
    k = torch.exp(tp["log_scale"]) * k
    ks = k + (torch.exp(2.0 * tp["log_sigma"]) + O.JITTER) * torch.eye(N, dtype=torch.float64)
    Lref = torch.linalg.cholesky(ks)
    loss = -O.multivariate_normal_logpdf(B, torch.from_numpy(y[:, None]), tp["mean"] * torch.ones(N, 1, dtype=torch.float64),
                                         Lref).reshape(())
    loss.backward()
    assert abs(nll - float(loss.detach())) <= 1e-9 * abs(float(loss.detach()))
    assert rel(np.tril(e.K.cpu().numpy())[:N, :N], Lref.detach().numpy()) < 1e-9
    for name in ["log_bw", "log_scale", "mean", "log_sigma"] + (["log_bw_embed"] if P else []):
        got = e.params.get_grad(name)
        assert rel(got, tp[name].grad.numpy().reshape(got.shape)) < 1e-7, name
    if P:
        assert rel(dE[:, :P], Et.grad.numpy()) < 1e-7

    kv = rs.rand(N) * 0.5 + 0.5 if P else np.ones(N)
    e.kvec.zero_()
    e.kvec[:N] = torch.from_numpy(kv).to(e.dev)
    cand = rs.randn(M, d) * 1.3
    y_max = float(y.max())
    kind = ["ei", "ucb", "lcb"][N % 3]
    out = e.score(cand, acq=kind, y_max=y_max, xi=0.01, kappa=2.58)
    b = O.TrainBatch(h_params=theta, obs=y[:, None])
    kd = np.ones((N, N))
    if P:
        emb = np.repeat(E, sizes, axis=0)
        kd = O.matern32_kernel(O.NP, emb, emb, np.exp(p["log_bw_embed"]))
    mean, v = O.predict(O.NP, p, "dist", b, cand, k_dist=kd, k_dist_pred=kv[:, None])
    std = O.std_from_var(v)
    acq = O.acquisition(kind, mean, std, y_max=y_max, xi=0.01, kappa=2.58)
    assert rel(out["mean"].cpu().numpy(), mean[:, 0]) < 1e-8
    assert np.max(np.abs(out["acq"].cpu().numpy() - acq)) < 1e-8 * max(1.0, np.max(np.abs(acq)))
    assert int(out["best_idx"].item()) == int(np.argmax(acq))
