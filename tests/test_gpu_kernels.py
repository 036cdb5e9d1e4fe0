"""GPU parity tests: every kernel of libdistbo_b200.so against the CPU oracle on the same seeded
inputs (called through the C ABI via distbo_b200.engine).  Tolerances are the north_star's:
fp64 GP quantities 1e-8 relative, fp32 embeddings 1e-5 relative."""
import math

import numpy as np
import pytest
import scipy.linalg

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import gp_oracle as O  # noqa: E402


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def make_problem(N, d, sizes=None, P=0, var=1e-2, seed=0):
    rs = np.random.RandomState(seed)
    theta = rs.randn(N, d)
    y = rs.randn(N)
    E = rs.randn(len(sizes), P) * 0.3 if P else None
    return rs, theta, y, E


def setup_engine(eng, theta, y, sizes=None, E=None, var=1e-2, log_bw=None, log_scale=0.3, mean=0.2,
                 log_bw_embed=None):
    N, d = theta.shape
    P = 0 if E is None else E.shape[1]
    e = eng.GPEngine(d, n_embed=P)
    e.params.set("log_sigma", 0.5 * math.log(var))
    e.params.set("log_scale", [log_scale])
    e.params.set("mean", [mean])
    if log_bw is not None:
        e.params.set("log_bw", log_bw)
    if P and log_bw_embed is not None:
        e.params.set("log_bw_embed", log_bw_embed)
    e.set_observations(theta, y, sizes)
    KE = None
    if P:
        e.E = torch.from_numpy(E).to(e.dev)
        KE = e.task_gram(e.E, len(sizes))
    e.build_gram(KE)
    return e


def oracle_params(e):
    p = {k: e.params.get(k) for k in e.params.offsets}
    return p


def oracle_K(p, theta, sizes=None, E=None):
    k = O.matern32_kernel(O.NP, theta, theta, np.exp(p["log_bw"]))
    if E is not None:
        emb = np.repeat(E, sizes, axis=0)
        k = k * O.matern32_kernel(O.NP, emb, emb, np.exp(p["log_bw_embed"]))
    k = np.exp(p["log_scale"]) * k
    ks = k + (np.exp(2.0 * p["log_sigma"]) + O.JITTER) * np.eye(len(theta))
    return k, ks


@pytest.mark.parametrize("N,d,sizes,P", [(200, 3, None, 0), (300, 2, [100, 50, 150], 5), (465, 1, [31] * 15, 10)])
def test_gram(cuda_engine_mod, N, d, sizes, P):
    rs, theta, y, E = make_problem(N, d, sizes, P)
    lbw = rs.randn(d) * 0.3
    lbe = rs.randn(P) * 0.3 if P else None
    e = setup_engine(cuda_engine_mod, theta, y, sizes, E, log_bw=lbw, log_bw_embed=lbe)
    p = oracle_params(e)
    _, ks = oracle_K(p, theta, sizes, E)
    K = e.K.cpu().numpy()
    got = np.tril(K[:N, :N])
    assert rel(got, np.tril(ks)) < 1e-13
    # pad block is the identity
    if e.n_pad > N:
        assert np.array_equal(np.tril(K[N:, N:]), np.eye(e.n_pad - N))
        assert not K[N:, :N].any()
    if P:
        emb_k = O.matern32_kernel(O.NP, E, E, np.exp(p["log_bw_embed"]))
        assert rel(e.KE.cpu().numpy(), emb_k) < 1e-13


@pytest.mark.parametrize("N", [100, 384, 700, 1300])
def test_cholesky_inverse(cuda_engine_mod, N):
    rs, theta, y, _ = make_problem(N, 4, seed=N)
    e = setup_engine(cuda_engine_mod, theta, y)
    p = oracle_params(e)
    _, ks = oracle_K(p, theta)
    e.factorize(need_inverse=True, need_kinv=True)
    e.likelihood()
    nll = e.read_loss()
    Lref = scipy.linalg.cholesky(ks, lower=True)
    L = np.tril(e.K.cpu().numpy())[:N, :N]
    assert rel(L, Lref) < 1e-10
    W = np.tril(e.W.cpu().numpy())[:N, :N]
    Wref = scipy.linalg.solve_triangular(Lref, np.eye(N), lower=True)
    assert rel(W, Wref) < 1e-9
    Kinv = np.tril(e.Tm.cpu().numpy())[:N, :N]
    assert rel(Kinv, np.tril(np.linalg.inv(ks))) < 1e-8
    b = O.TrainBatch(h_params=theta, obs=y[:, None])
    ref = float(O.neg_log_marg_likhd(O.NP, p, "params", b))
    assert abs(nll - ref) <= 1e-10 * abs(ref)
    V = e.V.cpu().numpy()[:N]
    Vref = scipy.linalg.solve_triangular(Lref, y - p["mean"][0], lower=True)
    assert rel(V, Vref) < 1e-9
    alpha = e.alpha.cpu().numpy()[:N]
    assert rel(alpha, scipy.linalg.solve_triangular(Lref.T, Vref, lower=False)) < 1e-8
    if e.n_pad > N:
        assert not e.V.cpu().numpy()[N:].any() and not e.alpha.cpu().numpy()[N:].any()


@pytest.mark.parametrize("far,group", [("3", "2"), ("6", "4"), ("0", "4")])
def test_potrf_schedules_agree(cuda_engine_mod, monkeypatch, far, group):
    """The near / far trailing-update schedule (DISTBO_POTRF_FAR / _FARG; default at N >= 7168) and the uniform one
    give the same factor, inverse and likelihood at a size the multi-launch path handles (n_pad = 3328: 26 column
    blocks, above the single-launch kernel's 16)."""
    eng = cuda_engine_mod
    monkeypatch.setenv("DISTBO_POTRF_FAR", far)
    monkeypatch.setenv("DISTBO_POTRF_FARG", group)
    N = 3300
    # (the schedule is chosen when the engine's plan is built, i.e. in setup_engine below)
    rs, theta, y, _ = make_problem(N, 4, seed=3)
    e = setup_engine(eng, theta, y)
    Kfull = torch.tril(e.K[:N, :N].clone())
    Kfull = Kfull + torch.tril(Kfull, -1).T
    e.factorize(need_inverse=True, need_kinv=True)
    e.likelihood()
    nll = e.read_loss()
    Lref = torch.linalg.cholesky(Kfull)
    assert rel(torch.tril(e.K[:N, :N]).cpu().numpy(), Lref.cpu().numpy()) < 1e-11
    Wref = torch.linalg.solve_triangular(Lref, torch.eye(N, dtype=torch.float64, device=e.dev), upper=False)
    assert rel(torch.tril(e.W[:N, :N]).cpu().numpy(), Wref.cpu().numpy()) < 1e-9
    yc = torch.from_numpy(y - float(e.params.get("mean")[0])).to(e.dev)
    a = torch.linalg.solve_triangular(Lref, yc[:, None], upper=False)
    ref = float(0.5 * (a * a).sum() + 0.5 * N * np.log(2 * np.pi) + torch.log(torch.diagonal(Lref)).sum())
    assert abs(nll - ref) <= 1e-10 * abs(ref)


def test_cholesky_not_pd_raises(cuda_engine_mod):
    rs, theta, y, _ = make_problem(150, 2)
    e = setup_engine(cuda_engine_mod, theta, y)
    e.K[10, 10] = -1.0
    e.factorize(need_inverse=False)
    e.likelihood()
    with pytest.raises(RuntimeError, match="not successful"):
        e.read_loss()


@pytest.mark.parametrize("N,d,sizes,P", [(200, 3, None, 0), (300, 2, [100, 50, 150], 5), (450, 1, [30] * 15, 10)])
def test_gram_gradient(cuda_engine_mod, N, d, sizes, P):
    rs, theta, y, E = make_problem(N, d, sizes, P, seed=3)
    lbw = rs.randn(d) * 0.3
    lbe = rs.randn(P) * 0.3 if P else None
    e = setup_engine(cuda_engine_mod, theta, y, sizes, E, log_bw=lbw, log_bw_embed=lbe)
    e.factorize(need_inverse=True, need_kinv=True)
    e.likelihood()
    e.gram_gradient()
    p = oracle_params(e)
    if P:
        dE = e.task_gram_backward(e.E, len(sizes)).cpu().numpy()
    # autograd reference (torch CPU fp64), embeddings as a leaf
    tp = {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=True) for k, v in p.items()}
    B = O.TH()
    th = torch.from_numpy(theta)
    k = O.matern32_kernel(B, th, th, torch.exp(tp["log_bw"]))
    if P:
        Et = torch.tensor(E, requires_grad=True)
        emb = B.repeat_rows(Et, sizes)
        k = k * O.matern32_kernel(B, emb, emb, torch.exp(tp["log_bw_embed"]))
    k = torch.exp(tp["log_scale"]) * k
    ks = k + (torch.exp(2.0 * tp["log_sigma"]) + O.JITTER) * torch.eye(N, dtype=torch.float64)
    L = torch.linalg.cholesky(ks)
    mean = tp["mean"] * torch.ones(N, 1, dtype=torch.float64)
    loss = -O.multivariate_normal_logpdf(B, torch.from_numpy(y[:, None]), mean, L).reshape(())
    loss.backward()
    for name in ["log_bw", "log_scale", "mean", "log_sigma"] + (["log_bw_embed"] if P else []):
        got = e.params.get_grad(name)
        want = tp[name].grad.numpy().reshape(got.shape)
        assert rel(got, want) < 1e-8, name
    if P:
        assert rel(dE[:, :P], Et.grad.numpy()) < 1e-8


@pytest.mark.parametrize("N,d,M,acq", [(200, 3, 1000, "ei"), (300, 8, 5000, "ucb"), (520, 2, 300, "lcb"),
                                       (200, 1, 20000, "ei")])
def test_score(cuda_engine_mod, N, d, M, acq):
    rs, theta, y, _ = make_problem(N, d, seed=7 + N)
    e = setup_engine(cuda_engine_mod, theta, y)
    e.factorize(need_inverse=True)
    e.likelihood()
    e.read_loss()
    kv = rs.rand(N) * 0.5 + 0.5
    e.kvec.zero_()
    e.kvec[:N] = torch.from_numpy(kv).to(e.dev)
    cand = rs.randn(M, d) * 1.5
    y_max, xi, kappa = float(np.max(y)), 0.01, 2.58
    out = e.score(cand, acq=acq, y_max=y_max, xi=xi, kappa=kappa)
    p = oracle_params(e)
    b = O.TrainBatch(h_params=theta, obs=y[:, None])
    # kernel 'dist' form with every column of k_dist_pred equal to kvec and k_dist = 1
    mean, var = O.predict(O.NP, p, "dist", b, cand, k_dist=np.ones((N, N)), k_dist_pred=kv[:, None])
    std = O.std_from_var(var)
    a = O.acquisition(acq, mean, std, y_max=y_max, xi=xi, kappa=kappa)
    scale = float(np.exp(p["log_scale"][0]))
    assert rel(out["mean"].cpu().numpy(), mean[:, 0]) < 1e-9
    assert np.max(np.abs(out["std"].cpu().numpy() ** 2 - std[:, 0] ** 2)) < 1e-9 * scale
    assert np.max(np.abs(out["acq"].cpu().numpy() - a)) < 1e-9 * max(1.0, np.max(np.abs(a)))
    assert int(out["best_idx"].item()) == int(np.argmax(a))
    assert float(out["best_val"].item()) == pytest.approx(float(np.max(a)), rel=1e-9, abs=1e-12)


def _nn_case(eng, mode, rs, sizes, seed):
    """'nn' operator cases (train/distbo_net.py:18-20): one two- or three-layer net over [x | y]; classification
    appends the class ratios (:43-48).  Returns (net, oracle params, X, Y, XY, model_type)."""
    S = sum(sizes)
    if mode == "nn3class":
        dx, dy, model_type = 16, 2, "classification"
        X = (rs.rand(S, dx) < 0.3).astype(np.float64)
        Y = np.eye(2)[(rs.rand(S) < 0.3).astype(int)]
    else:
        dx, dy, model_type = 7, 1, "regression"
        X = rs.randn(S, dx)
        Y = rs.rand(S, dy)
    wd = [20, 10] if mode == "nn2" else [20, 12, 10]
    net = eng.NetSpec(dx + dy, dy, wd[0], wd[-1], 0, 0, 4 if model_type == "classification" else 0,
                      pm=wd[1] if len(wd) == 3 else 0, xname="xy")
    p = O.build_params(2, "dist", in_dim_x=dx, in_dim_y=dy, weight_dim_xy=wd, model_type=model_type,
                       embed_op="nn", seed=seed)
    p["bias_xy_1"] = rs.randn(wd[0]) * 0.1
    if len(wd) == 3:
        p["bias_xy_2"] = rs.randn(wd[1]) * 0.1
    return net, p, X, Y, np.hstack([X, Y]), model_type


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-12), ("f32", 1e-5)])
@pytest.mark.parametrize("mode", ["nn2", "nn3", "nn3class"])
def test_embed_forward_nn(cuda_engine_mod, dtype, tol, mode):
    eng = cuda_engine_mod
    rs = np.random.RandomState(21)
    sizes = [300, 17, 1000, 129, 1]
    net, p, X, Y, XY, model_type = _nn_case(eng, mode, rs, sizes, 5)
    e = eng.GPEngine(2, n_embed=net.pooled_width, net=net)
    for k, v in p.items():
        if e.params.has(k) and k.startswith(("weights", "bias")):
            e.params.set(k, v)
    tdt = torch.float64 if dtype == "f64" else torch.float32
    off = torch.from_numpy(np.r_[0, np.cumsum(sizes)].astype(np.int32)).to(e.dev)
    Eg = e.embed(torch.from_numpy(XY).to(e.dev).to(tdt), torch.from_numpy(Y).to(e.dev).to(tdt), off,
                 dtype=tdt).double().cpu().numpy()
    want = O.distbo_net(O.NP, X, Y, sizes, p, embed_op="nn", model_type=model_type)
    assert want.shape[1] == net.pooled_width
    assert rel(Eg[:, :want.shape[1]], want) < tol


@pytest.mark.parametrize("mode", ["nn2", "nn3", "nn3class"])
def test_embed_backward_nn(cuda_engine_mod, mode):
    eng = cuda_engine_mod
    rs = np.random.RandomState(22)
    sizes = [200, 130, 77, 1]
    net, p, X, Y, XY, model_type = _nn_case(eng, mode, rs, sizes, 6)
    e = eng.GPEngine(2, n_embed=net.pooled_width, net=net)
    for k, v in p.items():
        if e.params.has(k) and k.startswith(("weights", "bias")):
            e.params.set(k, v)
    off = torch.from_numpy(np.r_[0, np.cumsum(sizes)].astype(np.int32)).to(e.dev)
    dE = rs.randn(len(sizes), net.pooled_width)
    e.embed_backward(torch.from_numpy(XY).to(e.dev), torch.from_numpy(Y).to(e.dev), off, torch.from_numpy(dE).to(e.dev))
    tp = {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=True) for k, v in p.items()}
    out = O.distbo_net(O.TH(), torch.from_numpy(X), torch.from_numpy(Y), sizes, tp, embed_op="nn", model_type=model_type)
    (out * torch.from_numpy(dE)).sum().backward()
    n_checked = 0
    for name in e.params.offsets:
        if name.startswith(("weights", "bias")):
            assert rel(e.params.get_grad(name), tp[name].grad.numpy()) < 1e-10, name
            n_checked += 1
    assert n_checked == (3 if mode == "nn2" else 5)


@pytest.mark.parametrize("dtype,tol", [("f64", 1e-12), ("f32", 1e-5)])
@pytest.mark.parametrize("mode", ["none", "joint", "class", "concat"])
def test_embed_forward(cuda_engine_mod, dtype, tol, mode):
    eng = cuda_engine_mod
    rs = np.random.RandomState(11)
    sizes = [300, 17, 1000, 129, 1]
    S = sum(sizes)
    if mode == "class":
        dx, dy, y_mode, model_type, embed_op = 16, 2, 2, "classification", "joint"
        X = (rs.rand(S, dx) < 0.3).astype(np.float64)
        lab = (rs.rand(S) < 0.3).astype(int)
        Y = np.eye(2)[lab]
    else:
        dx, dy, model_type = 7, 1, "regression"
        y_mode, embed_op = {"none": (0, "none"), "joint": (1, "joint"), "concat": (3, "concat")}[mode]
        X = rs.randn(S, dx)
        Y = rs.rand(S, dy)
    net = eng.NetSpec(dx, dy, 20, 10, 20 if y_mode in (1, 3) else 0, 10 if y_mode in (1, 3) else 0, y_mode)
    e = eng.GPEngine(2, n_embed=net.pooled_width, net=net)
    p = O.build_params(2, "dist", in_dim_x=dx, in_dim_y=dy, weight_dim_x=[20, 10], weight_dim_y=[20, 10],
                       model_type=model_type, embed_op=embed_op, seed=5)
    p["bias_x_1"] = rs.randn(20) * 0.1
    if "bias_y_1" in p:
        p["bias_y_1"] = rs.randn(20) * 0.1
    if "log_reg" in p:
        p["log_reg"] = np.asarray(-0.7)
    for k, v in p.items():
        if e.params.has(k) and k.startswith(("weights", "bias", "log_reg")):
            e.params.set(k, v)
    tdt = torch.float64 if dtype == "f64" else torch.float32
    Xd = torch.from_numpy(X).to(e.dev).to(tdt)
    Yd = torch.from_numpy(Y).to(e.dev).to(tdt)
    off = torch.from_numpy(np.r_[0, np.cumsum(sizes)].astype(np.int32)).to(e.dev)
    Eg = e.embed(Xd, Yd, off, dtype=tdt).double().cpu().numpy()
    want = O.distbo_net(O.NP, X, Y, sizes, p, embed_op=embed_op, model_type=model_type)
    assert Eg.shape[1] >= want.shape[1]
    if mode == "concat" and dtype == "f64":
        # the reference's Woodbury form (1/lambda)(M - M (G + lambda I)^-1 G) cancels ~cond(G + lambda I)
        # digits; the kernel evaluates the same operator as M (G + lambda I)^-1, so agreement is bounded by
        # the oracle's own rounding (measured 1.1e-12 here), still far inside north_star's 1e-8
        tol = 1e-10
    assert rel(Eg[:, :want.shape[1]], want) < tol


@pytest.mark.parametrize("case", [
    # (dx, dy, hx, p, classification, binary X, sizes, rows sliced off the front of the buffer)
    (166, 2, 20, 10, True, True, [1037, 128, 4434, 1, 255, 129], 0),     # protein fingerprints (data.py:27-38)
    (166, 2, 20, 10, True, False, [640, 3, 500], 1),                       # real-valued rows: the lo(x) MMA runs
    (17, 0, 20, 10, False, False, [97, 1, 130, 260], 3),                   # 'none' pooling, 68-byte rows
    (7, 3, 5, 3, True, False, [300, 17, 1000, 129, 1], 1),                 # 28-byte rows, 3 classes, one column tile
    (40, 4, 32, 16, True, False, [129, 383, 2], 0),                        # widest net / widest one-hot handled
    (1, 0, 20, 10, False, False, [500] * 4 + [3], 1),                      # toy: d_x = 1 (toy_datasets.py:44-46)
])
def test_embed_forward_tensor_core(cuda_engine_mod, case, monkeypatch):
    """fp32 prediction-time embedding on the tensor-core kernel (bulk-copy ring + 3xTF32): against the fp64 oracle
    (north_star: 1e-5 relative) and against the generic fp32 kernel on the same inputs; ragged datasets and row
    blocks whose ends are not 16-byte aligned (head / tail copies)."""
    eng = cuda_engine_mod
    dx, dy, hx, pw, classification, binary, sizes, skip = case
    rs = np.random.RandomState(dx * 31 + hx)
    S = sum(sizes)
    X = (rs.rand(S + skip, dx) < 0.3).astype(np.float64) if binary else rs.randn(S + skip, dx)
    if classification:
        Y = np.eye(dy)[rs.randint(0, dy, size=S + skip)]
        y_mode, model_type, embed_op = 2, "classification", "joint"
    else:
        dy = 1
        Y = rs.rand(S + skip, 1)
        y_mode, model_type, embed_op = 0, "regression", "none"
    net = eng.NetSpec(dx, dy, hx, pw, 0, 0, y_mode)
    e = eng.GPEngine(2, n_embed=net.pooled_width, net=net)
    p = O.build_params(2, "dist", in_dim_x=dx, in_dim_y=dy, weight_dim_x=[hx, pw], weight_dim_y=[hx, pw],
                       model_type=model_type, embed_op=embed_op, seed=7)
    p["bias_x_1"] = rs.randn(hx) * 0.1
    for k, v in p.items():
        if e.params.has(k) and k.startswith(("weights", "bias")):
            e.params.set(k, v)
    # fp32 inputs; the oracle sees the same fp32-rounded samples and weights in fp64
    Xd = torch.from_numpy(X).to(e.dev).float()[skip:]
    Yd = torch.from_numpy(Y).to(e.dev).float()[skip:]
    assert Xd.is_contiguous() and Yd.is_contiguous()
    off = torch.from_numpy(np.r_[0, np.cumsum(sizes)].astype(np.int32)).to(e.dev)
    p32 = {k: np.asarray(v, dtype=np.float32).astype(np.float64) for k, v in p.items()}
    want = O.distbo_net(O.NP, Xd.double().cpu().numpy(), Yd.double().cpu().numpy(), sizes, p32, embed_op=embed_op,
                        model_type=model_type)
    monkeypatch.setenv("DISTBO_EMBED_TC", "1")
    assert eng._lib.load().distbo_embed_f32_on_tensor_cores(dx, dy, hx, pw, 0, y_mode) == 1
    got = e.embed(Xd, Yd, off, dtype=torch.float32).double().cpu().numpy()[:, :want.shape[1]]
    monkeypatch.setenv("DISTBO_EMBED_TC", "0")
    assert eng._lib.load().distbo_embed_f32_on_tensor_cores(dx, dy, hx, pw, 0, y_mode) == 0
    generic = e.embed(Xd, Yd, off, dtype=torch.float32).double().cpu().numpy()[:, :want.shape[1]]
    assert rel(got, want) < 1e-5
    assert rel(generic, want) < 1e-5
    assert rel(got, generic) < 1e-5
    # deterministic: fixed-order reductions, no atomics
    monkeypatch.setenv("DISTBO_EMBED_TC", "1")
    again = e.embed(Xd, Yd, off, dtype=torch.float32).double().cpu().numpy()[:, :want.shape[1]]
    assert np.array_equal(got, again)


@pytest.mark.parametrize("mode", ["joint", "class", "concat"])
def test_embed_three_layer_x_net_with_y_side(cuda_engine_mod, mode):
    """x net of three layers (nn_net's weights_3 branch, distbo_net.py:11-13) under the poolings that have a y side:
    forward (fp64, fp32) and the closed-form backward against torch autograd through the oracle."""
    eng = cuda_engine_mod
    rs = np.random.RandomState(31)
    sizes = [150, 64, 77, 1, 200]
    S = sum(sizes)
    if mode == "class":
        dx, dy, y_mode, model_type, embed_op = 16, 2, 2, "classification", "joint"
        X = (rs.rand(S, dx) < 0.3).astype(np.float64)
        Y = np.eye(2)[(rs.rand(S) < 0.3).astype(int)]
    else:
        dx, dy, model_type = 9, 1, "regression"
        y_mode, embed_op = {"joint": (1, "joint"), "concat": (3, "concat")}[mode]
        X = rs.randn(S, dx)
        Y = rs.rand(S, dy)
    net = eng.NetSpec(dx, dy, 20, 10, 20 if y_mode in (1, 3) else 0, 10 if y_mode in (1, 3) else 0, y_mode, pm=12)
    e = eng.GPEngine(2, n_embed=net.pooled_width, net=net)
    p = O.build_params(2, "dist", in_dim_x=dx, in_dim_y=dy, weight_dim_x=[20, 12, 10], weight_dim_y=[20, 10],
                       model_type=model_type, embed_op=embed_op, seed=8)
    p["bias_x_1"] = rs.randn(20) * 0.1
    p["bias_x_2"] = rs.randn(12) * 0.1
    if "log_reg" in p:
        p["log_reg"] = np.asarray(0.3)
    for k, v in p.items():
        if e.params.has(k) and k.startswith(("weights", "bias", "log_reg")):
            e.params.set(k, v)
    off = torch.from_numpy(np.r_[0, np.cumsum(sizes)].astype(np.int32)).to(e.dev)
    want = O.distbo_net(O.NP, X, Y, sizes, p, embed_op=embed_op, model_type=model_type)
    Xd, Yd = torch.from_numpy(X).to(e.dev), torch.from_numpy(Y).to(e.dev)
    got32 = e.embed(Xd.float(), Yd.float(), off, dtype=torch.float32).double().cpu().numpy()[:, :want.shape[1]]
    assert rel(got32, want) < 1e-5
    got = e.embed(Xd, Yd, off).cpu().numpy()[:, :want.shape[1]]       # fp64 last: its per-dataset sums feed the backward
    assert rel(got, want) < (1e-10 if mode == "concat" else 1e-12)
    dE = rs.randn(len(sizes), net.pooled_width)
    e.embed_backward(Xd, Yd, off, torch.from_numpy(dE).to(e.dev))
    tp = {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=True) for k, v in p.items()}
    out = O.distbo_net(O.TH(), torch.from_numpy(X), torch.from_numpy(Y), sizes, tp, embed_op=embed_op, model_type=model_type)
    (out * torch.from_numpy(dE)).sum().backward()
    n = 0
    for name in e.params.offsets:
        if name.startswith(("weights", "bias", "log_reg")):
            assert rel(e.params.get_grad(name), tp[name].grad.numpy()) < 1e-10, name
            n += 1
    assert n == {"joint": 8, "class": 5, "concat": 9}[mode]


@pytest.mark.parametrize("T,dtype", [(2100, "f64"), (2100, "f32"), (1, "f32"), (1, "f64")])
def test_embed_many_tiny_datasets_and_single_dataset(cuda_engine_mod, T, dtype):
    """Chunk table beyond its shared-memory scan (T > 2048: serial fill path) and the degenerate single short dataset
    (one partial chunk); fp32 goes through the tensor-core kernel."""
    eng = cuda_engine_mod
    rs = np.random.RandomState(T)
    sizes = list(rs.randint(1, 4, size=T)) if T > 1 else [5]
    S = int(sum(sizes))
    dx = 5
    X, Y = rs.randn(S, dx), rs.rand(S, 1)
    net = eng.NetSpec(dx, 1, 20, 10, 0, 0, 0)
    e = eng.GPEngine(2, n_embed=net.pooled_width, net=net)
    p = O.build_params(2, "dist", in_dim_x=dx, in_dim_y=1, weight_dim_x=[20, 10], weight_dim_y=[20, 10],
                       model_type="regression", embed_op="none", seed=3)
    for k, v in p.items():
        if e.params.has(k) and k.startswith(("weights", "bias")):
            e.params.set(k, v)
    tdt = torch.float64 if dtype == "f64" else torch.float32
    off = torch.from_numpy(np.r_[0, np.cumsum(sizes)].astype(np.int32)).to(e.dev)
    got = e.embed(torch.from_numpy(X).to(e.dev).to(tdt), torch.from_numpy(Y).to(e.dev).to(tdt), off, dtype=tdt)
    want = O.distbo_net(O.NP, X, Y, sizes, p, embed_op="none", model_type="regression")
    assert rel(got.double().cpu().numpy()[:, :10], want) < (1e-12 if dtype == "f64" else 1e-5)


@pytest.mark.parametrize("mode", ["none", "joint", "class", "concat"])
def test_embed_backward(cuda_engine_mod, mode):
    eng = cuda_engine_mod
    rs = np.random.RandomState(12)
    sizes = [200, 130, 77, 1]
    S = sum(sizes)
    if mode == "class":
        dx, dy, y_mode, model_type, embed_op = 16, 2, 2, "classification", "joint"
        X = (rs.rand(S, dx) < 0.3).astype(np.float64)
        Y = np.eye(2)[(rs.rand(S) < 0.3).astype(int)]
    else:
        dx, dy, model_type = 40, 1, "regression"
        y_mode, embed_op = {"none": (0, "none"), "joint": (1, "joint"), "concat": (3, "concat")}[mode]
        X = rs.randn(S, dx)
        Y = rs.rand(S, dy)
    net = eng.NetSpec(dx, dy, 20, 10, 20 if y_mode in (1, 3) else 0, 10 if y_mode in (1, 3) else 0, y_mode)
    e = eng.GPEngine(2, n_embed=net.pooled_width, net=net)
    p = O.build_params(2, "dist", in_dim_x=dx, in_dim_y=dy, weight_dim_x=[20, 10], weight_dim_y=[20, 10],
                       model_type=model_type, embed_op=embed_op, seed=6)
    p["bias_x_1"] = rs.randn(20) * 0.1
    if "log_reg" in p:
        p["log_reg"] = np.asarray(0.4)
    for k, v in p.items():
        if e.params.has(k) and k.startswith(("weights", "bias", "log_reg")):
            e.params.set(k, v)
    Xd = torch.from_numpy(X).to(e.dev)
    Yd = torch.from_numpy(Y).to(e.dev)
    off = torch.from_numpy(np.r_[0, np.cumsum(sizes)].astype(np.int32)).to(e.dev)
    Pw = net.pooled_width
    dE = rs.randn(len(sizes), Pw)
    if y_mode == 3:
        e.embed(Xd, Yd, off)                      # the forward pass stores the per-dataset sums backward reads
    e.embed_backward(Xd, Yd, off, torch.from_numpy(dE).to(e.dev))
    tp = {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=True) for k, v in p.items()}
    out = O.distbo_net(O.TH(), torch.from_numpy(X), torch.from_numpy(Y), sizes, tp, embed_op=embed_op,
                       model_type=model_type)
    (out * torch.from_numpy(dE)).sum().backward()
    for name in e.params.offsets:
        if name.startswith(("weights", "bias", "log_reg")):
            assert rel(e.params.get_grad(name), tp[name].grad.numpy()) < 1e-10, name


@pytest.mark.parametrize("T", [1, 5, 42, 65])
def test_multi_task_kernel(cuda_engine_mod, T):
    """kernel 'multi' (train/kernels.py:32-40): task covariance and its gradient vs the oracle / torch autograd."""
    eng = cuda_engine_mod
    rs = np.random.RandomState(T)
    e = eng.GPEngine(2, n_multi=T)
    x = rs.normal(size=T * (T + 1) // 2)
    e.params.set("log_triangular_task", x)
    KE = e.task_tri().cpu().numpy()
    want = O.task_kernel_matrix(O.NP, {"log_triangular_task": x}, T)
    assert rel(KE, want) < 1e-13
    H = rs.randn(T, T)
    e.H = torch.from_numpy(H).to(e.dev)
    e.task_tri_backward()
    xt = torch.tensor(x, requires_grad=True)
    (O.task_kernel_matrix(O.TH(), {"log_triangular_task": xt}, T) * torch.from_numpy(H)).sum().backward()
    if T > 1:
        assert rel(e.params.get_grad("log_triangular_task"), xt.grad.numpy()) < 1e-10
    else:
        assert abs(e.params.get_grad("log_triangular_task")[0]) < 1e-12


def test_adam_matches_tf_form(cuda_engine_mod):
    eng = cuda_engine_mod
    e = eng.GPEngine(3)
    rs = np.random.RandomState(0)
    x0 = rs.randn(e.params.n)
    e.params.flat.copy_(torch.from_numpy(x0))
    adam = O.TFAdam(0.005)
    ref = {"x": x0.copy()}
    for step in range(5):
        g = rs.randn(e.params.n) * 10
        e.params.grad.copy_(torch.from_numpy(g))
        e.adam_step(0.005, clip_norm=5.0 if step % 2 else 0.0)
        ref = adam.step(ref, {"x": g}, clip_norm=5.0 if step % 2 else None)
        assert rel(e.params.flat.cpu().numpy(), ref["x"]) < 1e-14
