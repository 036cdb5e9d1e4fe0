"""Parity against numbers computed by the REFERENCE'S OWN SOURCES.

tests/golden/ref_*.npz were written by tools/make_reference_golden.py: hcllaw/distBO's distBO/train/gp_regressor.py
(with train.py, log_gaus_likelihood.py, kernels.py, distbo_net.py, feature_map.py, base.py, gp_utils.py) imported
unmodified from /root/reference and executed over tools/tf1_shim.py, a lazy TensorFlow-1.x op shim evaluated with
torch CPU fp64 (TensorFlow 1.7 itself cannot be installed).  Each file holds the inputs, the initial parameter
values, the loss of every Adam epoch, the fitted parameters, posterior mean / std on 60 candidates and the
similarities, for one kernel / embedding-operator combination.

  * CPU (`-m "not gpu"`): the oracle restatement reproduces them (1e-9) - this is what pins the oracle.
  * GPU (`-m gpu`): distbo_b200.train.gp_regressor (CUDA through the C ABI) reproduces them (1e-8, north_star).
"""
import glob
import json
import os

import numpy as np
import pytest

from oracle.gp_regressor_oracle import OracleGPRegressor

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLD, "ref_*.npz")))


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def load_case(name):
    z = np.load(os.path.join(GOLD, "ref_%s.npz" % name), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    T = meta["T"]
    common = dict(model_type=meta["model_type"], seed=meta["seed"], float64=True)
    fit = dict(meta["fit"])
    sizes = [int(v) for v in z["sizes"]]
    if meta["kernel"] == "dist":
        X = [z["X%d" % t] for t in range(T + 1)]
        Y = [z["Y%d" % t] for t in range(T + 1)]
        common.update(target_X=X[T], target_Y=Y[T], target_embed_ind=z["target_embed_ind"])
        fit.update(data_X=X[:T], data_Y=Y[:T], source_embed_ind=list(z["source_embed_ind"]), sizes=sizes)
    elif meta["kernel"] == "manual":
        fit.update(data_embed=z["embed"][:T], target_embed=z["embed"][T:], sizes=sizes)
    elif meta["kernel"] == "multi":
        fit.update(sizes=sizes)
    init = {k[len("init__"):]: z[k] for k in z.files if k.startswith("init__")}
    final = {k[len("final__"):]: z[k] for k in z.files if k.startswith("final__")}
    return z, meta, common, fit, init, final


def check(reg, z, meta, fit, final, get_param, tol):
    loss = reg.maximise_likhd(z["theta"], z["y"], **fit)
    want = z["loss_log"]
    assert len(reg.loss_history) == len(want)
    assert rel(reg.loss_history, want) < tol
    assert abs(loss - float(z["final_loss"])) <= tol * abs(float(z["final_loss"]))
    n = 0
    for name, w in final.items():
        got = get_param(name)
        if got is not None:
            assert rel(np.asarray(got).reshape(-1), np.asarray(w).reshape(-1)) < tol, name
            n += 1
    assert n == len(final)
    reg.initialise_predict()
    mean, std = reg.predict_y(z["cand"], compute_embed=True)
    assert mean.shape == z["pred_mean"].shape and std.shape == z["pred_std"].shape
    assert rel(mean, z["pred_mean"]) < tol
    scale = float(np.exp(np.asarray(final["log_scale"]).reshape(-1)[0]))
    assert np.max(np.abs(std ** 2 - z["pred_std"] ** 2)) < tol * scale
    if "similarity" in z.files:
        assert rel(reg.similarity()[0], z["similarity"]) < tol


def test_fixture_set_is_complete():
    assert {"params", "manual", "multi", "dist_none", "dist_joint_clip", "dist_joint_batched", "dist_joint_stratified",
            "dist_class", "dist_concat", "dist_nn3", "dist_nn2class"} <= set(CASES)


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference(name):
    z, meta, common, fit, init, final = load_case(name)
    o = OracleGPRegressor(meta["kernel"], **common)
    o.initial_weights = init
    check(o, z, meta, fit, final, lambda k: o.p.get(k), 1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_reproduces_reference(cuda_engine_mod, name):
    from distbo_b200.train import gp_regressor
    z, meta, common, fit, init, final = load_case(name)
    g = gp_regressor(meta["kernel"], **common)
    g.initial_weights = init
    check(g, z, meta, fit, final, lambda k: g.engine.params.get(k) if g.engine.params.has(k) else None, 1e-8)
    g.close()
