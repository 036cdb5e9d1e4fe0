"""GPU parity of the surrogate facade: distbo_b200.train.gp_regressor (CUDA, through the C ABI) against
the oracle's restatement of train/gp_regressor.py + train/train.py on the same seeded inputs - Adam
trajectory of the marginal likelihood, posterior mean / std, similarity and the acquisition arg-max.
Tolerance: 1e-8 relative (north_star, fp64)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import gp_oracle as O  # noqa: E402
from oracle.gp_regressor_oracle import OracleGPRegressor  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def make_tasks(rs, T, per_task, d, dx, dy, rows, classification):
    sizes = [per_task] * T
    theta = rs.uniform(-3, 3, size=(T * per_task, d))
    y = np.sin(theta.sum(1)) + np.repeat(rs.randn(T) * 0.2, per_task) + 0.05 * rs.randn(T * per_task)
    X, Y = [], []
    for t in range(T + 1):
        n = rows + 7 * t
        if classification:
            X.append((rs.rand(n, dx) < 0.3).astype(np.float64))
            Y.append(np.eye(dy)[(rs.rand(n) < 0.4).astype(int)])
        else:
            X.append(rs.randn(n, dx) + 0.3 * t)
            Y.append(rs.rand(n, dy))
    return theta, y[:, None], sizes, X, Y


def pair(kernel, cuda_engine_mod, **kw):
    from distbo_b200.train import gp_regressor
    return gp_regressor(kernel, **kw), OracleGPRegressor(kernel, **kw)


@pytest.mark.parametrize("case", ["params", "manual", "multi", "dist_none", "dist_joint", "dist_class",
                                  "dist_batched", "dist_concat", "dist_nn", "dist_nn3", "dist_nn3class", "dist_x3"])
def test_fit_predict_parity(cuda_engine_mod, case):
    rs = np.random.RandomState(42)
    classification = case in ("dist_class", "dist_nn3class")
    d, T, per_task = 3, 5, 24
    dx, dy = (12, 2) if classification else (6, 1)
    theta, y, sizes, X, Y = make_tasks(rs, T, per_task, d, dx, dy, 150, classification)
    kernel = case.split("_")[0]
    model_type = "classification" if classification else "regression"
    common = dict(model_type=model_type, seed=7, float64=True)
    fit = dict(learning_rate=0.005, max_epochs=6, likhd_var_init=1e-2, first_early_stop_epoch=3,
               gradient_clip=(case == "dist_joint"))
    if kernel == "dist":
        ind = np.arange(100)
        common.update(target_X=X[T], target_Y=Y[T], target_embed_ind=ind)
        fit.update(data_X=X[:T], data_Y=Y[:T], source_embed_ind=[np.arange(120)] * T, sizes=sizes,
                   embed_op={"dist_none": "none", "dist_concat": "concat", "dist_nn": "nn", "dist_nn3": "nn",
                             "dist_nn3class": "nn", "dist_x3": "none"}.get(case, "joint"),
                   # 'nn' (train/distbo_net.py:18-20): one net over [x | y], two or three layers; dist_x3: a three-layer
                   # x net under the 'none' pooling (nn_net's weights_3 branch, :11-13)
                   weight_dim_x=[20, 12, 10] if case == "dist_x3" else [20, 10], weight_dim_y=[20, 10],
                   weight_dim_xy=[20, 10] if case == "dist_nn" else [20, 12, 10],
                   batch_size=64 if case == "dist_batched" else 500, stratify=False,
                   concat_data_size=(case != "dist_none"), max_size=400)
    elif kernel == "manual":
        emb = rs.rand(T + 1, 9)
        fit.update(data_embed=emb[:T], target_embed=emb[T:], sizes=sizes)
    elif kernel == "multi":
        fit.update(sizes=sizes)                    # the last task is the target (kernels.py:47-49)
    g, o = pair(kernel, cuda_engine_mod, **common)
    lg = g.maximise_likhd(theta, y, **fit)
    lo = o.maximise_likhd(theta, y, **fit)
    assert len(g.loss_history) == len(o.loss_history) == 6
    assert rel(g.loss_history, o.loss_history) < 1e-8
    assert abs(lg - lo) <= 1e-8 * abs(lo)
    for name, want in o.p.items():
        if g.engine.params.has(name):
            assert rel(g.engine.params.get(name), np.asarray(want).reshape(g.engine.params.get(name).shape)) < 1e-8, name

    g.initialise_predict()
    cand = rs.uniform(-3, 3, size=(700, d))
    mg, sg = g.predict_y(cand, compute_embed=True)
    mo, so = o.predict_y(cand, compute_embed=True)
    assert mg.shape == mo.shape == (700, 1) and sg.shape == so.shape
    assert rel(mg, mo) < 1e-8
    scale = float(np.exp(o.p["log_scale"][0]))
    assert np.max(np.abs(sg ** 2 - so ** 2)) < 1e-8 * scale
    if kernel != "params":
        assert rel(g.similarity()[0], o.similarity()[0]) < 1e-8
    y_max = float(y.max())
    ei = O.acq_ei(mo, so, y_max, 0.01)
    v, i = g.acquire(cand, "ei", y_max=y_max, xi=0.01)
    assert i == int(np.argmax(ei))
    assert v == pytest.approx(float(ei.max()), rel=1e-7, abs=1e-12)

    # second fit: scaler frozen, parameters and Adam slots carried over (gp_regressor.py:58-64,141-143)
    theta2 = np.vstack([theta, rs.uniform(-3, 3, size=(1, d))])
    y2 = np.vstack([y, [[0.3]]])
    if kernel == "dist":
        fit.update(data_X=X, data_Y=Y, source_embed_ind=[np.arange(120)] * (T + 1), sizes=sizes + [1])
    elif kernel == "manual":
        fit.update(data_embed=emb, sizes=sizes + [1])
    elif kernel == "multi":
        fit.update(sizes=sizes[:-1] + [sizes[-1] + 1])
    lg = g.maximise_likhd(theta2, y2, **fit)
    lo = o.maximise_likhd(theta2, y2, **fit)
    assert rel(g.loss_history, o.loss_history) < 1e-8
    mg, sg = g.predict_y(cand)
    mo, so = o.predict_y(cand)
    assert rel(mg, mo) < 1e-8
    g.close()


@pytest.mark.parametrize("case", ["dist_joint", "dist_class"])
def test_fp32_embedding_at_prediction(cuda_engine_mod, case):
    """float64=False: the datasets' rows are embedded in fp32 at prediction time (north_star: embeddings within
    1e-5 relative); the GP itself stays fp64, so the posterior moves by ~1e-5 at most relative to the oracle."""
    rs = np.random.RandomState(17)
    classification = case == "dist_class"
    d, T, per = 2, 4, 20
    dx, dy = (16, 2) if classification else (5, 1)
    theta, y, sizes, X, Y = make_tasks(rs, T, per, d, dx, dy, 300, classification)
    model_type = "classification" if classification else "regression"
    fit = dict(learning_rate=0.005, max_epochs=4, likhd_var_init=1e-2, first_early_stop_epoch=2, data_X=X[:T],
               data_Y=Y[:T], source_embed_ind=[np.arange(len(x)) for x in X[:T]], sizes=sizes, embed_op="joint",
               weight_dim_x=[20, 10], weight_dim_y=[20, 10], batch_size=1000, stratify=False, concat_data_size=True,
               max_size=400)
    common = dict(model_type=model_type, seed=3, target_X=X[T], target_Y=Y[T], target_embed_ind=np.arange(len(X[T])))
    from distbo_b200.train import gp_regressor
    g = gp_regressor("dist", float64=False, **common)
    o = OracleGPRegressor("dist", float64=True, **common)
    g.maximise_likhd(theta, y, **fit)
    o.maximise_likhd(theta, y, **fit)
    assert rel(g.loss_history, o.loss_history) < 1e-8          # training is fp64 in both
    cand = rs.uniform(-3, 3, size=(500, d))
    mg, sg = g.predict_y(cand)
    mo, so = o.predict_y(cand)
    E32 = g._E_all.cpu().numpy()
    assert g._prediction_sets()[0].dtype == torch.float32
    assert rel(mg, mo) < 2e-4 and np.max(np.abs(sg - so)) < 2e-4
    # the embeddings themselves: fp32 kernel vs fp64 oracle
    p = o.p
    want = O.dist_features(O.NP, p, np.vstack(X), np.vstack(Y), [1] * (T + 1), [len(x) for x in X],
                           data_sizes=[len(x) for x in X], max_size=400, embed_op="joint", model_type=model_type,
                           return_pooled=True)
    assert rel(E32[:, :want.shape[1]], want) < 1e-5
