"""Full-size checks (BASELINE config 5: N = 8192, 64 tasks, d_theta = 8), where the CPU oracle would take
minutes: size-independent properties of the factorisation (L L^T = K, W L = I, K^-1 K = I on random vectors),
the likelihood against an independent device path (torch / cuSOLVER fp64, used here only as a checker), the
scoring pass against dense torch algebra on a candidate sample, and shard-count invariance of the arg-max."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def big(cuda_engine_mod):
    eng = cuda_engine_mod
    rs = np.random.RandomState(0)
    N, d, T, P = 8192, 8, 64, 23
    theta = rs.uniform(-5, 5, size=(N, d))
    theta = (theta - theta.mean(0)) / theta.std(0)
    y = np.sin(theta @ rs.randn(d) / math.sqrt(d)) + 0.1 * rs.randn(N)
    e = eng.GPEngine(d, n_embed=P)
    e.params.set("log_sigma", 0.5 * math.log(1e-2))
    e.params.set("log_scale", [0.2])
    e.params.set("mean", [0.1])
    e.params.set("log_bw", rs.randn(d) * 0.2)
    e.params.set("log_bw_embed", rs.randn(P) * 0.2)
    e.set_observations(theta, y, [N // T] * T)
    e.E = torch.from_numpy(rs.randn(T, P) * 0.3).to(e.dev)
    KE = e.task_gram(e.E, T)
    e.build_gram(KE)
    K = torch.tril(e.K.clone())
    K = K + torch.tril(K, -1).T                     # full symmetric K_sigma (n_pad == N here)
    e.factorize(need_inverse=True, need_kinv=True)
    e.likelihood()
    nll = e.read_loss()
    return dict(e=e, K=K, N=N, d=d, T=T, theta=theta, y=y, nll=nll, rs=rs)


def relerr(a, b):
    return float((a - b).norm() / b.norm())


def test_factorisation_properties(big):
    e, K, N = big["e"], big["K"], big["N"]
    g = torch.Generator(device="cpu").manual_seed(1)
    v = torch.randn(N, 4, generator=g, dtype=torch.float64).to(e.dev)
    L = torch.tril(e.K)
    W = torch.tril(e.W)
    assert relerr(L @ (L.T @ v), K @ v) < 1e-12
    assert relerr(W @ (L @ v), v) < 1e-10
    Kinv = torch.tril(e.Tm)
    Kinv = Kinv + torch.tril(Kinv, -1).T
    assert relerr(Kinv @ (K @ v), v) < 1e-9
    r = torch.from_numpy(big["y"]).to(e.dev) - 0.1
    assert relerr(K @ e.alpha, r) < 1e-9          # alpha = K^-1 (y - m)
    assert relerr(L @ e.V, r) < 1e-11             # V = L^-1 (y - m)


def test_likelihood_against_cusolver(big):
    e, K, N = big["e"], big["K"], big["N"]
    Lref = torch.linalg.cholesky(K)
    r = (torch.from_numpy(big["y"]).to(e.dev) - 0.1)[:, None]
    a = torch.linalg.solve_triangular(Lref, r, upper=False)
    ref = 0.5 * float((a * a).sum()) + 0.5 * N * math.log(2 * math.pi) + float(torch.log(torch.diagonal(Lref)).sum())
    assert abs(big["nll"] - ref) <= 1e-10 * abs(ref)
    assert relerr(torch.tril(e.K), Lref) < 1e-11


def test_scoring_against_dense_algebra_and_sharding(big):
    e, N, d, T = big["e"], big["N"], big["d"], big["T"]
    rs = np.random.RandomState(5)
    M = 3 * 128 + 57
    cand = rs.randn(M, d) * 1.2
    kv = rs.rand(N) * 0.5 + 0.5
    e.kvec.zero_()
    e.kvec[:N] = torch.from_numpy(kv).to(e.dev)
    y_max = float(big["y"].max())
    out = e.score(cand, acq="ei", y_max=y_max, xi=0.01)
    # dense checker on the device (torch fp64): k_* = scale matern(theta, cand) kvec; A = L^-1 k_*
    bw = torch.exp(e.params.view("log_bw"))
    th = torch.from_numpy(big["theta"]).to(e.dev) / bw
    c = torch.from_numpy(cand).to(e.dev) / bw
    d2 = (-2.0 * th @ c.T + (th * th).sum(1, keepdim=True) + (c * c).sum(1)[None, :]).clamp_min(1e-32)
    v = math.sqrt(3.0) * d2.sqrt()
    scale = float(torch.exp(e.params.view("log_scale")))
    ks = scale * (1 + v) * torch.exp(-v) * torch.from_numpy(kv).to(e.dev)[:, None]
    A = torch.linalg.solve_triangular(torch.tril(e.K), ks, upper=False)
    mean = (A.T @ e.V[:N]) + 0.1
    var = scale - (A * A).sum(0)
    assert float((out["mean"] - mean).abs().max()) < 1e-9 * max(1.0, float(mean.abs().max()))
    assert float((out["std"] ** 2 - var.clamp_min(1e-32)).abs().max()) < 1e-9 * scale
    std = var.clamp_min(1e-32).sqrt()
    z = (mean - y_max - 0.01) / std
    ei = (mean - y_max - 0.01) * 0.5 * torch.erfc(-z / math.sqrt(2.0)) + std * torch.exp(-0.5 * z * z) / math.sqrt(2 * math.pi)
    best_i, best_v = int(out["best_idx"].item()), float(out["best_val"].item())   # engine-owned scalars
    assert best_i == int(torch.argmax(ei).item())
    # shard invariance: scoring [lo, hi) ranges separately and merging gives the same (value, index)
    from distbo_b200.sharding import merge_best, shard_range
    vals, idxs = [], []
    for r in range(3):
        lo, hi = shard_range(M, r, 3)
        o = e.score(cand[lo:hi], acq="ei", y_max=y_max, xi=0.01, want_arrays=False, index_offset=lo)
        vals.append(float(o["best_val"].item()))
        idxs.append(int(o["best_idx"].item()))
    v_best, i_best = merge_best(vals, idxs)
    assert i_best == best_i and v_best == best_v


def test_gradient_directional_derivative(big):
    """Closed-form gradient at full size: a central difference of the NLL along a random direction of
    (log_bw, log_scale, mean, log_sigma, log_bw_embed, E) matches g . direction."""
    e, T = big["e"], big["T"]
    rs = np.random.RandomState(9)
    names = ["log_bw", "log_scale", "mean", "log_sigma", "log_bw_embed"]
    base = {k: e.params.get(k).copy() for k in names}
    E0 = e.E.clone()

    def nll_at(step, dirs, dE):
        for k in names:
            e.params.set(k, base[k] + step * dirs[k])
        e.E = E0 + step * dE
        e.build_gram(e.task_gram(e.E, T))
        e.factorize(need_inverse=True, need_kinv=True)
        e.likelihood()
        return e.read_loss()

    dirs = {k: rs.randn(*np.shape(base[k])) if np.ndim(base[k]) else np.asarray(rs.randn()) for k in names}
    dE = torch.from_numpy(rs.randn(*E0.shape) * 0.1).to(e.dev)
    nll_at(0.0, dirs, dE)
    e.gram_gradient()
    gE = e.task_gram_backward(e.E, T)
    g_dot = sum(float(np.sum(e.params.get_grad(k) * dirs[k])) for k in names) + float((gE * dE).sum())
    h = 1e-5
    fd = (nll_at(h, dirs, dE) - nll_at(-h, dirs, dE)) / (2 * h)
    nll_at(0.0, dirs, dE)
    assert abs(fd - g_dot) <= 2e-6 * max(1.0, abs(g_dot)), (fd, g_dot)


def test_fit_kernels_bitwise_reproducible():
    """The multi-stream POTRF schedule (panel / next / window / far streams), TRTRI and LAUUM use fixed tile
    products in fixed orders: L, L^-1, K^-1 and the likelihood of the same matrix are bitwise identical from run to
    run, eagerly and as a CUDA-graph replay (a race between the streams would show up here; tools/stress_potrf.py
    is the long version)."""
    from distbo_b200 import engine as eng
    N = 7300                                   # 58 column blocks: the near / far schedule
    rs = np.random.RandomState(1)
    e = eng.GPEngine(8)
    e.params.set("log_sigma", 0.5 * np.log(1e-2))
    e.set_observations(rs.randn(N, 8), rs.randn(N))

    def run():
        e.build_gram(None)
        e.factorize(need_inverse=True, need_kinv=True)
        e.likelihood()

    run()
    torch.cuda.synchronize()
    assert int(e.info.item()) == 0
    ref = [e.K.clone(), e.W.clone(), e.Tm.clone(), e.nll.clone()]
    for _ in range(4):
        run()
        torch.cuda.synchronize()
        assert all(torch.equal(a, b) for a, b in zip((e.K, e.W, e.Tm, e.nll), ref))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    for _ in range(4):
        g.replay()
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip((e.K, e.W, e.Tm, e.nll), ref))
