"""GPU parity at the FULL shapes of the reference's shipped configurations (task counts,
observations per task, hyper-parameter and dataset dimensions, embedding width, batch sizes as the experiment
scripts set them), on synthetic data of those shapes: a few optimiser steps and one acquisition pass of the CUDA
surrogate against the CPU oracle.  The toy configuration (config 1) runs end to end in test_gpu_trajectory.py and
config 5 in test_gpu_fullsize.py / bench.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import gp_oracle as O  # noqa: E402
from oracle.gp_regressor_oracle import OracleGPRegressor  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def protein_like(rs):
    """experiments/protein_config.py: 6 source proteins + target, 10 random + 10 BO evaluations each, forest with
    4 hyper-parameters, 166 binary features, one-hot binary label, 0.6 n_i training rows (n_i = 1037..4434), batch
    1000, concat_data_size -> P = 10*2 + 2 + 1 = 23."""
    T, per, d, dx, dy = 6, 20, 4, 166, 2
    n_rows = [622, 2660, 1391, 1024, 1876, 933, 1203]           # 0.6 * sizes of the seven protein files
    X = [(rs.rand(n, dx) < 0.25).astype(np.float64) for n in n_rows]
    Y = [np.eye(dy)[(rs.rand(n) < 0.3).astype(int)] for n in n_rows]
    kw = dict(embed_op="joint", weight_dim_x=[20, 10], weight_dim_y=[20, 10], batch_size=1000, stratify=True,
              concat_data_size=True, max_size=max(n_rows))
    return "dist", "classification", T, per, d, X, Y, None, kw


def parkinson_like(rs):
    """experiments/train_test.py defaults on the Parkinson data: 41 source patients + target, 30 evaluations each,
    ridge with 2 hyper-parameters, 17 features, scalar label in [0, 1], 60..100 rows per patient, joint -> P = 100."""
    T, per, d, dx, dy = 41, 30, 2, 17, 1
    n_rows = [int(n) for n in rs.randint(60, 101, size=T + 1)]
    X = [rs.randn(n, dx) + 0.1 * t for t, n in enumerate(n_rows)]
    Y = [rs.rand(n, dy) for n in n_rows]
    kw = dict(embed_op="joint", weight_dim_x=[20, 10], weight_dim_y=[20, 10], batch_size=250, stratify=True,
              concat_data_size=False)
    return "dist", "regression", T, per, d, X, Y, None, kw


def counter_manual_like(rs):
    """experiments/counter_manual_config.py with manualBO: 4 sources + target, 75 + 75 evaluations each, 7
    hyper-parameters, <= 26 MinMax-scaled meta-features per dataset."""
    T, per, d = 4, 150, 7
    meta = rs.rand(T + 1, 26)
    return "manual", "regression", T, per, d, None, None, meta, {}


@pytest.mark.parametrize("make", [protein_like, parkinson_like, counter_manual_like])
def test_config_shape_parity(cuda_engine_mod, make):
    from distbo_b200.train import gp_regressor
    rs = np.random.RandomState(2024)
    kernel, model_type, T, per, d, X, Y, meta, kw = make(rs)
    sizes = [per] * T
    theta = rs.uniform(0, 1, size=(T * per, d)) * rs.uniform(1, 30, size=d)
    y = np.sin(3 * theta / theta.max(0)).sum(1, keepdims=True) / d + 0.05 * rs.randn(T * per, 1)
    common = dict(model_type=model_type, seed=11, float64=True)
    fit = dict(learning_rate=0.005, max_epochs=4, likhd_var_init=1e-5, first_early_stop_epoch=2, sizes=sizes, **kw)
    if kernel == "dist":
        common.update(target_X=X[T], target_Y=Y[T], target_embed_ind=np.arange(len(X[T])))
        fit.update(data_X=X[:T], data_Y=Y[:T], source_embed_ind=[np.arange(len(x)) for x in X[:T]])
    else:
        fit.update(data_embed=meta[:T], target_embed=meta[T:])
    g, o = gp_regressor(kernel, **common), OracleGPRegressor(kernel, **common)
    g.maximise_likhd(theta, y, **fit)
    o.maximise_likhd(theta, y, **fit)
    assert rel(g.loss_history, o.loss_history) < 1e-8
    for name, want in o.p.items():
        if g.engine.params.has(name):
            assert rel(g.engine.params.get(name), np.asarray(want).reshape(g.engine.params.get(name).shape)) < 1e-8, name
    g.initialise_predict()
    M = 100000                                       # a third of the configs' n_warmup keeps the oracle to seconds
    cand = rs.uniform(0, 1, size=(M, d)) * theta.max(0)
    mo, so = o.predict_y(cand, compute_embed=True)
    y_max = float(y[-per:].max())
    ei = O.acq_ei(mo, so, y_max, 0.01)
    v, i = g.acquire(cand, "ei", y_max=y_max, xi=0.01)
    assert i == int(np.argmax(ei))
    assert v == pytest.approx(float(ei.max()), rel=1e-6, abs=1e-12)
    mg, sg = g.predict_y(cand[:5000])
    assert rel(mg, mo[:5000]) < 1e-7
    scale = float(np.exp(o.p["log_scale"][0]))
    assert np.max(np.abs(sg ** 2 - so[:5000] ** 2)) < 1e-7 * scale
    assert rel(g.similarity()[0], o.similarity()[0]) < 1e-8
    g.close()
