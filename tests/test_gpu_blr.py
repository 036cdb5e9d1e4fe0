"""GPU parity of the BLR surrogate (algorithm='BLR', reference train/log_gaus_likelihood_feature.py) against the
oracle restatement: loss, full gradient (vs torch autograd = TF autodiff), posterior, acquisition arg-max and the
gp_regressor-level training trajectory."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import blr_oracle as BO  # noqa: E402
from oracle import gp_oracle as O  # noqa: E402
from oracle.gp_regressor_oracle import OracleGPRegressor  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("kernel,weight_dim,N", [("params", [50, 50, 50], 333), ("manual", [50, 50, 50], 64),
                                                 ("multi", [24, 17], 200), ("manual", [64, 64, 64], 1000),
                                                 ("params", [7, 5], 1)])
def test_blr_fit_and_score(cuda_engine_mod, kernel, weight_dim, N):
    from distbo_b200 import blr_engine
    rs = np.random.RandomState(len(weight_dim) * 1000 + N)
    d, T = 5, 6
    sizes = np.full(T, N // T)
    sizes[: N - sizes.sum()] += 1
    sizes = [int(s) for s in sizes if s > 0]
    T = len(sizes)
    theta, y = rs.randn(N, d), rs.randn(N, 1)
    Pe = {"params": 0, "manual": 9, "multi": T}[kernel]
    p, rff_in = BO.build_params_blr(d, kernel, weight_dim, in_dim_meta=9, n_tasks=T, likhd_var_init=0.02, seed=5)
    for k in p:
        if k.startswith("bias"):
            p[k] = rs.randn(*p[k].shape) * 0.2
    p["log_alpha"] = np.asarray(-0.4)
    meta = rs.rand(T + 1, 9)
    batch = O.TrainBatch(h_params=theta, obs=y, sizes=sizes, embed=meta[:T] if kernel == "manual" else None)
    want_loss, want_g = BO.blr_loss_and_grads(p, kernel, batch, rff_in, 0, n_tasks=T)

    e = blr_engine.BLREngine(d, weight_dim, n_embed=Pe)
    for k, v in p.items():
        e.params.set(k, v)
    e.set_observations(theta, y, sizes if kernel != "params" else None)
    E = None
    if kernel == "manual":
        E = torch.from_numpy(meta[:T]).to(e.dev)
    elif kernel == "multi":
        E = torch.eye(T, dtype=torch.float64, device=e.dev)
    e.loss_and_grad(E_const=E)
    assert e.read_loss() == pytest.approx(want_loss, rel=1e-11)
    for name in p:
        if name == "log_bw":
            assert not e.params.get_grad(name).any()       # unused without the random Fourier map
            continue
        assert rel(e.params.get_grad(name), want_g[name]) < 1e-9, name

    # posterior + acquisition over candidates (ragged count: not a multiple of the 64-row tile)
    M = 1000 + 37
    cand = rs.uniform(-2, 2, size=(M, d))
    if kernel == "manual":
        emb_t = meta[T]
        fea_pred = np.concatenate([np.repeat(emb_t[None], M, 0), cand], 1)
    elif kernel == "multi":
        emb_t = np.eye(T)[T - 1]
        fea_pred = np.concatenate([np.repeat(emb_t[None], M, 0), cand], 1)
    else:
        emb_t, fea_pred = None, cand
    fea = BO.train_feature(O.NP, p, kernel, batch, n_tasks=T)
    phi, phi_c = BO.nn_features(O.NP, p, fea, rff_in, 0), BO.nn_features(O.NP, p, fea_pred, rff_in, 0)
    mean, var = BO.blr_predict(O.NP, p, phi, y, phi_c)
    std = O.std_from_var(var)
    e.prepare_posterior(E, None if emb_t is None else torch.from_numpy(emb_t).to(e.dev))
    for kind, kw in (("ei", dict(y_max=float(y.max()), xi=0.01)), ("ucb", dict(kappa=2.58)), ("lcb", dict(kappa=1.0))):
        out = e.score(cand, acq=kind, want_arrays=True, index_offset=11, **kw)
        a = O.acquisition(kind, mean, std, **kw).ravel()
        assert rel(out["mean"].cpu().numpy(), mean.ravel()) < 1e-9
        assert rel(out["std"].cpu().numpy(), std.ravel()) < 1e-9
        assert np.allclose(out["acq"].cpu().numpy(), a, rtol=1e-8, atol=1e-12)
        assert int(out["best_idx"].item()) == 11 + int(np.argmax(a))
        assert float(out["best_val"].item()) == pytest.approx(float(a.max()), rel=1e-8, abs=1e-12)


def make_tasks(rs, T, per_task, d, dx, dy, n_rows, classification):
    sizes = [per_task] * T
    theta = rs.uniform(-3, 3, size=(T * per_task, d))
    y = np.sin(theta).sum(1) + 0.1 * rs.randn(T * per_task)
    X, Y = [], []
    for t in range(T + 1):
        n = n_rows + 7 * t
        if classification:
            X.append((rs.rand(n, dx) < 0.3).astype(np.float64))
            Y.append(np.eye(dy)[(rs.rand(n) < 0.4).astype(int)])
        else:
            X.append(rs.randn(n, dx) + 0.3 * t)
            Y.append(rs.rand(n, dy))
    return theta, y[:, None], sizes, X, Y


@pytest.mark.parametrize("case", ["params", "manual", "multi", "dist_joint", "dist_class", "dist_concat",
                                  "dist_batched"])
def test_blr_regressor_parity(cuda_engine_mod, case):
    from distbo_b200.train import gp_regressor
    rs = np.random.RandomState(43)
    classification = case == "dist_class"
    d, T, per_task = 3, 5, 24
    dx, dy = (12, 2) if classification else (6, 1)
    theta, y, sizes, X, Y = make_tasks(rs, T, per_task, d, dx, dy, 150, classification)
    kernel = case.split("_")[0]
    common = dict(model_type="classification" if classification else "regression", seed=7, float64=True)
    fit = dict(learning_rate=0.005, max_epochs=6, likhd_var_init=1e-2, first_early_stop_epoch=3, algorithm="BLR",
               weight_dim=[50, 50, 50], num_of_rff=0, gradient_clip=(case == "dist_joint"))
    if kernel == "dist":
        common.update(target_X=X[T], target_Y=Y[T], target_embed_ind=np.arange(100))
        fit.update(data_X=X[:T], data_Y=Y[:T], source_embed_ind=[np.arange(120)] * T, sizes=sizes,
                   embed_op="concat" if case == "dist_concat" else "joint", weight_dim_x=[20, 10],
                   weight_dim_y=[20, 10], batch_size=64 if case == "dist_batched" else 500, stratify=False,
                   concat_data_size=True, max_size=400)
    elif kernel == "manual":
        emb = rs.rand(T + 1, 9)
        fit.update(data_embed=emb[:T], target_embed=emb[T:], sizes=sizes)
    elif kernel == "multi":
        fit.update(sizes=sizes)
    g, o = gp_regressor(kernel, **common), OracleGPRegressor(kernel, **common)
    lg = g.maximise_likhd(theta, y, **fit)
    lo = o.maximise_likhd(theta, y, **fit)
    assert len(g.loss_history) == len(o.loss_history) == 6
    assert rel(g.loss_history, o.loss_history) < 1e-8
    assert abs(lg - lo) <= 1e-8 * abs(lo)
    for name, want in o.p.items():
        assert g.engine.params.has(name), name
        assert rel(g.engine.params.get(name), np.asarray(want).reshape(g.engine.params.get(name).shape)) < 1e-8, name
    g.initialise_predict()
    cand = rs.uniform(-3, 3, size=(700, d))
    mg, sg = g.predict_y(cand, compute_embed=True)
    mo, so = o.predict_y(cand, compute_embed=True)
    assert mg.shape == mo.shape == (700, 1) and sg.shape == so.shape
    assert rel(mg, mo) < 1e-7
    assert rel(sg, so) < 1e-7
    assert g.similarity() == ['none']
    y_max = float(y.max())
    ei = O.acq_ei(mo, so, y_max, 0.01)
    v, i = g.acquire(cand, "ei", y_max=y_max, xi=0.01)
    assert i == int(np.argmax(ei))
    assert v == pytest.approx(float(ei.max()), rel=1e-6, abs=1e-12)
    # second fit: scaler frozen, parameters and Adam slots carried over
    theta2 = np.vstack([theta, rs.uniform(-3, 3, size=(1, d))])
    y2 = np.vstack([y, [[0.3]]])
    if kernel == "dist":
        fit.update(data_X=X, data_Y=Y, source_embed_ind=[np.arange(120)] * (T + 1), sizes=sizes + [1])
    elif kernel == "manual":
        fit.update(data_embed=emb, sizes=sizes + [1])
    elif kernel == "multi":
        fit.update(sizes=sizes[:-1] + [sizes[-1] + 1])
    g.maximise_likhd(theta2, y2, **fit)
    o.maximise_likhd(theta2, y2, **fit)
    assert rel(g.loss_history, o.loss_history) < 1e-8
    mg, _ = g.predict_y(cand)
    mo, _ = o.predict_y(cand)
    assert rel(mg, mo) < 1e-7
    g.close()


def test_blr_rejects_fourier_map(cuda_engine_mod):
    from distbo_b200.train import gp_regressor
    g = gp_regressor("params", seed=1, float64=True)
    with pytest.raises(NotImplementedError):
        g.maximise_likhd(np.zeros((4, 2)), np.zeros((4, 1)), algorithm="BLR", weight_dim=[8, 8], num_of_rff=100)
