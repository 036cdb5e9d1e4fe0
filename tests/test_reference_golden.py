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
    # the random Fourier block runs in float32 inside the reference (feature_map.py:20-25): its posterior agrees to
    # fp32 rounding of cos / matmul between libraries, not to 1e-9
    pred_tol = 1e-5 if fit.get("num_of_rff") else tol
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
    assert rel(mean, z["pred_mean"]) < pred_tol
    tol = pred_tol
    # variance relative to the prior scale: exp(log_scale) for the GP, alpha |phi|^2 ~ the largest variance for BLR
    scale = float(np.exp(np.asarray(final["log_scale"]).reshape(-1)[0])) if "log_scale" in final else \
        float(np.max(z["pred_std"] ** 2))
    assert np.max(np.abs(std ** 2 - z["pred_std"] ** 2)) < tol * scale
    if "similarity" in z.files:
        assert rel(reg.similarity()[0], z["similarity"]) < tol


def test_fixture_set_is_complete():
    assert {"params", "manual", "multi", "dist_none", "dist_joint_clip", "dist_joint_batched", "dist_joint_stratified",
            "dist_class", "dist_concat", "dist_nn3", "dist_nn2class", "blr_params", "blr_manual", "blr_multi",
            "blr_dist_joint", "blr_dist_class"} <= set(CASES)


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
    if fit.get("num_of_rff"):        # not on the sm_100a path: refused loudly, no CPU fallback
        with pytest.raises(NotImplementedError):
            g.maximise_likhd(z["theta"], z["y"], **fit)
        return
    check(g, z, meta, fit, final, lambda k: g.engine.params.get(k) if g.engine.params.has(k) else None, 1e-8)
    g.close()


# ---------------------------------------------------------------------------------------------------------------
# BO trajectories written by the reference's own driver (bayes_opt_wrap -> BayesianOptimization.maximize -> acq_max,
# UtilityFunction, TargetSpace, set_model('toy')) over the same shim: tests/golden/refbo_*.npz.
BO_CASES = sorted(os.path.basename(p)[6:-4] for p in glob.glob(os.path.join(GOLD, "refbo_*.npz")))


def bo_optim_config(**over):
    cfg = dict(weight_dim=None, weight_dim_x=[20, 10], weight_dim_y=[20, 10], weight_dim_xy=None, n_warmup=20000,
               n_opt_iterations=0, n_epochs=25, lkh_var_init=1e-5, lr=0.005, float64=True, n_cpus=1, stratify=False,
               embed_op='none', concat_data_size=False, algorithm='GP', gradient_clip=False, num_of_rff=0,
               batch_size=500, first_early_stop_epoch=10, max_size=None)
    cfg.update(over)
    return cfg


def run_bo(name, make_regressor):
    from distbo_b200.data.toy_datasets import one_dim_split_tasks
    from distbo_b200.model_embed import bayes_opt_wrap
    z = np.load(os.path.join(GOLD, "refbo_%s.npz" % name), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    init = {k[len("init__"):]: z[k] for k in z.files if k.startswith("init__")}
    target, sources, _ = one_dim_split_tasks(0, 3, 2, n_train=120, n_test=120)
    if meta["method"] == "manual":
        for dset, e in zip(sources + [target], z["embed"]):
            dset.embed = e
    init_list = [z["init_list%d" % i].copy() for i in range(len(sources))]

    def factory(kernel, **kw):
        reg = make_regressor(kernel, **kw)
        reg.initial_weights = init
        return reg
    out = bayes_opt_wrap(meta["method"], target, acq='ei', model='toy', rs=np.random.RandomState(7),
                         optim_config=bo_optim_config(embed_op=meta["embed_op"], **meta["extra"]), xi=0.01, kappa=2.58,
                         bo_it=5, rand_it=meta["rand_it"], init_list=init_list, include_init=False, source_data=sources,
                         gp_class=factory, verbose=0, acq_one_shot=meta["acq_one_shot"])
    traj, sim = (out if meta["method"] in ("dist", "manual", "multi") else (out, None))
    want = z["trajectory"]
    assert traj.shape == want.shape
    if meta["extra"].get("n_opt_iterations"):
        # L-BFGS-B polish (helpers.py:70-86): finite-difference steps through the posterior; same optimum to 1e-6
        assert np.allclose(traj, want, rtol=1e-6, atol=1e-8)
    else:
        # suggestions are rows of the same RandomState candidate stream: identical picks are bit-equal
        assert np.array_equal(traj[:, 0], want[:, 0]), (traj[:, 0], want[:, 0])
        assert np.array_equal(traj[:, 1], want[:, 1])
    if sim is not None and "similarity" in z.files:
        for got, ref in zip(sim, z["similarity"]):
            ref = ref[~np.isnan(ref)]
            # several refits at sigma^2 = 1e-5 amplify rounding-level differences (measured 4e-7 after five fits)
            assert np.allclose(np.asarray(got, dtype=float).reshape(-1), ref, rtol=1e-5, atol=0)


def test_bo_fixture_set_is_complete():
    assert {"dist_none", "dist_joint", "params", "multi", "manual", "params_polish"} <= set(BO_CASES)


@pytest.mark.parametrize("name", BO_CASES)
def test_oracle_bo_trajectory_matches_reference(name):
    run_bo(name, OracleGPRegressor)


@pytest.mark.gpu
@pytest.mark.parametrize("name", BO_CASES)
def test_cuda_bo_trajectory_matches_reference(cuda_engine_mod, name):
    from distbo_b200.train import gp_regressor
    run_bo(name, gp_regressor)


# ---------------------------------------------------------------------------------------------------------------
# fp32 embeddings computed by the reference's distbo_net in tf.float32 on the reference's own data (subsets):
# tests/golden/refembed_f32.npz.  north_star: embeddings within 1e-5 relative (fp32).
EMBED_CASES = {"protein_class": ("joint", "classification", 2), "parkinson_none": ("none", "regression", 0),
               "parkinson_joint": ("joint", "regression", 1)}


def load_embed(tag):
    z = np.load(os.path.join(GOLD, "refembed_f32.npz"))
    w = {k[len(tag) + 2:]: z[k] for k in z.files if k.startswith(tag + "__weights") or k.startswith(tag + "__bias")}
    return z[tag + "__X"], z[tag + "__Y"], [int(v) for v in z[tag + "__sizes"]], w, z[tag + "__E"]


@pytest.mark.parametrize("tag", sorted(EMBED_CASES))
def test_oracle_embedding_matches_reference_fp32(tag):
    from oracle import gp_oracle as O
    X, Y, sizes, w, E = load_embed(tag)
    embed_op, model_type, _ = EMBED_CASES[tag]
    p = {k: v.astype(np.float64) for k, v in w.items()}
    got = O.distbo_net(O.NP, X.astype(np.float64), Y.astype(np.float64), sizes, p, embed_op=embed_op, model_type=model_type)
    assert got.shape == E.shape
    assert rel(got, E) < 1e-5          # fp64 evaluation of the same fp32 inputs vs the reference's fp32 arithmetic


@pytest.mark.gpu
@pytest.mark.parametrize("tag", sorted(EMBED_CASES))
def test_cuda_fp32_embedding_matches_reference(cuda_engine_mod, tag):
    import torch
    eng = cuda_engine_mod
    X, Y, sizes, w, E = load_embed(tag)
    embed_op, model_type, y_mode = EMBED_CASES[tag]
    dx, dy = X.shape[1], Y.shape[1]
    net = eng.NetSpec(dx, dy, 20, 10, 20 if y_mode == 1 else 0, 10 if y_mode == 1 else 0, y_mode)
    e = eng.GPEngine(2, n_embed=net.pooled_width, net=net)
    for k, v in w.items():
        e.params.set(k, v.astype(np.float64))
    off = torch.from_numpy(np.r_[0, np.cumsum(sizes)].astype(np.int32)).to(e.dev)
    on_tc = eng._lib.load().distbo_embed_f32_on_tensor_cores(dx, dy, 20, 10, 0, y_mode)
    assert on_tc == (0 if y_mode == 1 else 1)      # fingerprints / 'none' pooling: the bulk-copy + TF32 tensor-core kernel
    got = e.embed(torch.from_numpy(X).to(e.dev), torch.from_numpy(Y).to(e.dev), off, dtype=torch.float32)
    got = got.double().cpu().numpy()[:, :E.shape[1]]
    assert rel(got, E) < 1e-5


# ---------------------------------------------------------------------------------------------------------------
# acquisition values from the reference's UtilityFunction (bayes_opt/helpers.py:114-177): tests/golden/refacq.npz
def test_acquisition_functions_match_reference():
    from oracle import gp_oracle as O
    from distbo_b200.bayes_opt.helpers import UtilityFunction
    z = np.load(os.path.join(GOLD, "refacq.npz"))
    mean, std = z["mean"], z["std"]
    y_max, xi, kappa = float(z["y_max"]), float(z["xi"]), float(z["kappa"])
    assert np.allclose(np.reshape(O.acq_ei(mean, std, y_max, xi), -1), z["ei"], rtol=1e-12, atol=1e-300)
    assert np.allclose(np.reshape(O.acq_ucb(mean, std, kappa), -1), z["ucb"], rtol=1e-14)
    assert np.allclose(np.reshape(O.acq_lcb(mean, std, kappa), -1), z["lcb"], rtol=1e-14)

    class StubGP(object):                       # duck-typed surrogate without the fused device path
        def predict_y(self, x_, **kw):
            return mean[:len(x_)], std[:len(x_)]
    x = np.zeros((len(mean), 2))
    for kind in ("ei", "ucb", "lcb", "poi"):
        got = UtilityFunction(kind=kind, kappa=kappa, xi=xi).utility(x, StubGP(), y_max)
        assert np.allclose(np.reshape(got, -1), z[kind], rtol=1e-12, atol=1e-300), kind
        v, i = UtilityFunction(kind=kind, kappa=kappa, xi=xi).argmax(x, StubGP(), y_max)
        assert i == int(np.argmax(z[kind])) and v == pytest.approx(float(np.max(z[kind])), rel=1e-12)


def test_early_stopping_epoch_matches_reference():
    z = np.load(os.path.join(GOLD, "ref_params_early_stop.npz"))
    assert 100 < len(z["loss_log"]) < 140      # the reference's run stopped on its countdown, not on max_epochs


# ---------------------------------------------------------------------------------------------------------------
# manual meta-features from the reference's model_embed/meta_fea.py meta(): tests/golden/refmeta.npz
@pytest.mark.parametrize("tag", ["protein", "parkinson"])
def test_manual_meta_features_match_reference(tag):
    import warnings
    from sklearn.utils import check_random_state
    from distbo_b200.data.load_datasets import parkinson_tasks, protein_tasks
    from distbo_b200.model_embed.meta_fea import meta
    z = np.load(os.path.join(GOLD, "refmeta.npz"))
    if tag == "protein":
        zz = np.load(os.path.join(GOLD, "protein_subset.npz"))
        bits = np.unpackbits(zz["bits"], axis=2)[:, :, :166].astype(np.float64)
        tables = [np.hstack([np.zeros((bits.shape[1], 1)), bits[i], zz["labels"][i][:, None].astype(np.float64)])
                  for i in range(bits.shape[0])]
        t, s, _ = protein_tasks(tables, [str(n) for n in zz["names"]], target=2, random_state=check_random_state(3),
                                test_size=0.4)
        model = "forest"
    else:
        zz = np.load(os.path.join(GOLD, "parkinson_subset.npz"))
        t, s, _ = parkinson_tasks([zz["patient_%d" % i] for i in range(10)], target=4,
                                  random_state=check_random_state(1), test_size=0.4)
        model = "ridge"
    dsets = [t] + s
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # constant fingerprint bits give NaN skewness, as in the reference
        scaler = meta(dsets, model, max_size=max(len(d.train_x) for d in dsets), random_state=5)
    got = np.asarray([d.embed for d in dsets])
    assert got.shape == z[tag].shape
    assert np.allclose(got, z[tag], rtol=1e-10, atol=1e-12, equal_nan=True)
    assert np.allclose(scaler.data_min_, z[tag + "__data_min"], rtol=1e-10, atol=1e-12, equal_nan=True)
    assert np.allclose(scaler.data_max_, z[tag + "__data_max"], rtol=1e-10, atol=1e-12, equal_nan=True)


# ---------------------------------------------------------------------------------------------------------------
# data loaders and the train/test container vs the reference's own (tests/golden/refdata.npz holds digests)
def _digest(d):
    return {"shape": np.asarray(d.train_x.shape + d.test_x.shape), "colsum": d.train_x.sum(0), "ysum": d.train_y.sum(0),
            "test_colsum": d.test_x.sum(0), "row0": d.train_x[0], "embed_ind": np.asarray(list(d.embed_ind))[:64],
            "n_embed": np.asarray(len(list(d.embed_ind)))}


def test_train_test_container_matches_reference():
    from sklearn.utils import check_random_state
    from distbo_b200.data.data import data
    z = np.load(os.path.join(GOLD, "refdata.npz"))
    x, yr, yc = z["container__x"], z["container__yr"], z["container__yc"]
    d1 = data(x[:200], x[200:], yr[:200], yr[200:], embed_size=50, random_state=check_random_state(9))
    d2 = data(x[:200], x[200:], yc[:200, None], yc[200:, None], embed_size=50, random_state=check_random_state(9),
              prob_type='classification')
    assert np.array_equal(np.asarray(list(d1.embed_ind)), z["container__reg_embed_ind"])
    assert np.array_equal(np.asarray(list(d2.embed_ind)), z["container__cls_embed_ind"])
    assert np.array_equal(d2.train_y, z["container__cls_train_y"])


@pytest.mark.skipif(not os.path.isdir("/root/reference/protein"), reason="the reference's data folders exist only in the build container")
@pytest.mark.parametrize("tag", ["protein", "parkinson"])
def test_loaders_match_reference_on_its_data_folders(tag):
    from sklearn.utils import check_random_state
    from distbo_b200.data.load_datasets import load_parkinson, load_protein
    z = np.load(os.path.join(GOLD, "refdata.npz"))
    if tag == "protein":
        t, s, names = load_protein(2, "/root/reference/protein", random_state=check_random_state(3), test_size=0.4)
    else:
        t, s, names = load_parkinson(4, "/root/reference/parkinson", random_state=check_random_state(1), test_size=0.4)
    assert [str(n) for n in names] == [str(n) for n in z[tag + "__names"]]
    for i, d in enumerate([t] + s):
        for k, v in _digest(d).items():
            want = z["%s__%d__%s" % (tag, i, k)]
            assert np.allclose(v, want, rtol=1e-13, atol=1e-13), (tag, i, k)


# ---------------------------------------------------------------------------------------------------------------
# toy task family (config 1) from the reference's generate_data / one_d_split: tests/golden/reftoy.npz
@pytest.mark.parametrize("tag", ["a", "b"])
def test_toy_tasks_match_reference(tag):
    from distbo_b200.data.toy_datasets import one_dim_split_tasks
    z = np.load(os.path.join(GOLD, "reftoy.npz"))
    seed, fake, true, n = (int(v) for v in z[tag + "__cfg"])
    target, sources, supp = one_dim_split_tasks(seed, fake, true, n_train=n, n_test=n)
    assert np.array_equal(np.asarray(supp, dtype=np.float64), z[tag + "__supp"])
    for i, d in enumerate([target] + sources):
        assert np.array_equal(d.train_x, z["%s__%d__train_x" % (tag, i)])
        assert float(d.test_x.sum()) == float(z["%s__%d__test_sum" % (tag, i)])


# ---------------------------------------------------------------------------------------------------------------
# BO trajectory at the Parkinson configuration through the reference's driver and ridge objective:
# tests/golden/refpark_bo_ridge.npz (joint embedding of 17-feature patient datasets, P = 100)
def run_bo_parkinson(make_regressor):
    from sklearn.utils import check_random_state
    from distbo_b200.data.load_datasets import parkinson_tasks
    from distbo_b200.model_embed import bayes_opt_wrap
    z = np.load(os.path.join(GOLD, "refpark_bo_ridge.npz"))
    zp = np.load(os.path.join(GOLD, "parkinson_subset.npz"))
    target, sources, _ = parkinson_tasks([zp["patient_%d" % i] for i in range(10)], target=4,
                                         random_state=check_random_state(1), test_size=0.4)
    sources = sources[:5]
    init = {k[len("init__"):]: z[k] for k in z.files if k.startswith("init__")}
    init_list = [z["init_list%d" % i].copy() for i in range(len(sources))]

    def factory(kernel, **kw):
        reg = make_regressor(kernel, **kw)
        reg.initial_weights = init
        return reg
    cfg = bo_optim_config(embed_op='joint', n_warmup=5000, n_epochs=20, stratify=False, batch_size=1000)
    traj, _ = bayes_opt_wrap('dist', target, acq='ei', model='ridge', rs=np.random.RandomState(11), optim_config=cfg,
                             xi=0.01, kappa=2.58, bo_it=4, rand_it=0, init_list=init_list, include_init=False,
                             source_data=sources, gp_class=factory, verbose=0, acq_one_shot='ei')
    want = z["trajectory"]
    assert traj.shape == want.shape == (4, 3)
    assert np.array_equal(traj[:, :2], want[:, :2]), (traj, want)          # identical suggestions
    assert np.allclose(traj[:, 2], want[:, 2], rtol=1e-12)                 # the ridge objective's R^2 at them


def test_init_list_of_parkinson_run_is_the_reference_ridge_objective():
    """The stored previous evaluations were produced by the reference's own ridge objective: ours returns the same."""
    from sklearn.utils import check_random_state
    from distbo_b200.data.load_datasets import parkinson_tasks
    from distbo_b200.model_embed import set_model
    z = np.load(os.path.join(GOLD, "refpark_bo_ridge.npz"))
    zp = np.load(os.path.join(GOLD, "parkinson_subset.npz"))
    _, sources, _ = parkinson_tasks([zp["patient_%d" % i] for i in range(10)], target=4,
                                    random_state=check_random_state(1), test_size=0.4)
    for i, s in enumerate(sources[:5]):
        f, bounds, _, _ = set_model('ridge', s, None)
        rows = z["init_list%d" % i]
        for la, lg, val in rows[:3]:
            assert f(log_alpha=la, log_gamma=lg) == pytest.approx(val, rel=1e-10)


def test_oracle_parkinson_bo_trajectory_matches_reference():
    run_bo_parkinson(OracleGPRegressor)


@pytest.mark.gpu
def test_cuda_parkinson_bo_trajectory_matches_reference(cuda_engine_mod):
    from distbo_b200.train import gp_regressor
    run_bo_parkinson(gp_regressor)


# ---------------------------------------------------------------------------------------------------------------
# The headline configuration at full size (N = 8192, 64 tasks, d_theta = 8, P = 23) through the reference's
# gp_regressor: tests/golden/refconfig5.npz (outputs only; the inputs are bench.make_workload's seeded streams)
# Two files: refconfig5.npz (sigma^2 = 1e-2, 256 posterior candidates) and refconfig5_s1e-5.npz (the reference's
# DEFAULT noise sigma^2 = 1e-5 + 1e-6 jitter, experiments/train_test.py:79, 4096 posterior candidates, with
# cond(K_sigma) of the matrix the reference factorises recorded in the file).
CONFIG5 = {"s1e-2": ("refconfig5.npz", 1e-2, 256), "s1e-5": ("refconfig5_s1e-5.npz", 1e-5, 4096)}


def run_config5(make_regressor, tol, variant="s1e-2"):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    import bench
    fname, lkh_var, n_cand = CONFIG5[variant]
    z = np.load(os.path.join(GOLD, fname))
    if "lkh_var_init" in z.files:
        assert float(z["lkh_var_init"]) == lkh_var
        assert 1.0 < float(z["cond_K_sigma"]) < 1e12
    work = bench.make_workload(bench.N_OBS, n_cand, 1000)
    T = bench.T_SRC
    X = [work["data_X"][t][:1000 - 3 * t] for t in range(T + 1)]
    Y = [work["data_Y"][t][:1000 - 3 * t] for t in range(T + 1)]
    reg = make_regressor("dist", model_type="classification", seed=0, float64=True, target_X=X[T], target_Y=Y[T],
                         target_embed_ind=np.arange(len(X[T])))
    reg.initial_weights = {k[len("init__"):]: z[k] for k in z.files if k.startswith("init__")}
    loss = reg.maximise_likhd(
        work["theta"], work["y"][:, None], data_X=X[:T], data_Y=Y[:T],
        source_embed_ind=[np.arange(len(x)) for x in X[:T]], embed_op="joint", algorithm="GP", weight_dim_x=[20, 10],
        weight_dim_y=[20, 10], batch_size=1000, sizes=work["sizes"], concat_data_size=True, learning_rate=0.005,
        max_epochs=2, max_size=1000, likhd_var_init=lkh_var, stratify=False, gradient_clip=False,
        first_early_stop_epoch=10 ** 9)
    assert rel(reg.loss_history, z["loss_log"]) < tol
    assert abs(loss - float(z["final_loss"])) <= tol * abs(float(z["final_loss"]))
    reg.initialise_predict()
    mean, std = reg.predict_y(work["cand"], compute_embed=True)
    assert rel(mean, z["pred_mean"]) < tol
    scale = float(np.exp(z["final__log_scale"].reshape(-1)[0]))
    assert np.max(np.abs(std ** 2 - z["pred_std"] ** 2)) < tol * scale
    assert rel(reg.similarity()[0], z["similarity"]) < tol
    # acquisition sweep: the reference's acq_max picked a row of the bench's candidate stream
    from distbo_b200.bayes_opt.helpers import UtilityFunction
    cand = np.random.RandomState(2).uniform(-5.0, 5.0, size=(4096, 8))
    want_i = int(np.argmax(z["acq_ei"]))
    assert np.array_equal(cand[want_i], z["acq_x_max"])
    util = UtilityFunction(kind='ei', kappa=2.58, xi=0.01)
    ei = np.reshape(util.utility(cand, reg, work["y_max"]), -1)
    assert int(np.argmax(ei)) == want_i
    assert np.max(np.abs(ei - z["acq_ei"])) < max(tol, 1e-8) * max(1.0, float(np.max(np.abs(z["acq_ei"]))))
    if hasattr(reg, "acquire"):          # the fused device pass (posterior + EI + arg-max in one sweep)
        v, i = util.argmax(cand, reg, work["y_max"])
        assert i == want_i
        assert v == pytest.approx(float(z["acq_ei"][want_i]), rel=1e-7)
    return reg


@pytest.mark.parametrize("variant", sorted(CONFIG5))
def test_oracle_reproduces_reference_at_config5_full_size(variant):
    o = run_config5(OracleGPRegressor, 1e-9, variant)
    z = np.load(os.path.join(GOLD, CONFIG5[variant][0]))
    for k in z.files:
        if k.startswith("final__"):      # parameters after two Adam steps: the full-size gradient is the reference's
            assert rel(np.asarray(o.p[k[7:]]).reshape(-1), z[k].reshape(-1)) < 1e-9, k


@pytest.mark.gpu
@pytest.mark.parametrize("variant", sorted(CONFIG5))
def test_cuda_reproduces_reference_at_config5_full_size(cuda_engine_mod, variant):
    from distbo_b200.train import gp_regressor
    g = run_config5(gp_regressor, 1e-8, variant)
    z = np.load(os.path.join(GOLD, CONFIG5[variant][0]))
    for k in z.files:
        if k.startswith("final__") and g.engine.params.has(k[7:]):
            assert rel(g.engine.params.get(k[7:]).reshape(-1), z[k].reshape(-1)) < 1e-8, k
    g.close()


# ---------------------------------------------------------------------------------------------------------------
# Config 4 (counter_manual): the reference's manualBO pipeline end to end - generate_data('counter_manual'),
# meta_fea.meta(), the 7-d dim_ridge objective, manual-kernel BO loop: tests/golden/refcounter_manual_bo.npz
def counter_manual_setup():
    import warnings
    from distbo_b200.data.toy_datasets import counter_manual_tasks
    from distbo_b200.model_embed.meta_fea import meta
    z = np.load(os.path.join(GOLD, "refcounter_manual_bo.npz"))
    target, sources, _ = counter_manual_tasks(3, noise=0.5, n_train=300, n_test=300)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        meta([target] + sources, 'dim_ridge', max_size=None, random_state=5)
    return z, target, sources


def test_counter_manual_inputs_match_reference():
    """Datasets, meta-features and objective of config 4 are the reference's."""
    from distbo_b200.model_embed import set_model
    z, target, sources = counter_manual_setup()
    assert np.array_equal(target.train_x, z["target_train_x"]) and np.array_equal(target.train_y, z["target_train_y"])
    assert np.array_equal(sources[0].train_x, z["source0_train_x"])
    assert np.allclose(np.asarray([d.embed for d in [target] + sources]), z["embed"], rtol=1e-10, atol=1e-12)
    names = ['log_alpha'] + ['log_bw%d' % i for i in range(1, 7)]
    for i, s in enumerate(sources):
        f, bounds, _, model_type = set_model('dim_ridge', s, None)
        assert model_type == 'regression' and sorted(bounds) == sorted(names)
        for row in z["init_list%d" % i][:3]:
            assert f(**dict(zip(names, row[:-1]))) == pytest.approx(row[-1], rel=1e-9, abs=1e-12)


def run_counter_manual_bo(make_regressor):
    from distbo_b200.model_embed import bayes_opt_wrap
    z, target, sources = counter_manual_setup()
    init = {k[len("init__"):]: z[k] for k in z.files if k.startswith("init__")}
    init_list = [z["init_list%d" % i].copy() for i in range(len(sources))]

    def factory(kernel, **kw):
        reg = make_regressor(kernel, **kw)
        reg.initial_weights = init
        return reg
    cfg = bo_optim_config(n_warmup=5000, n_epochs=20, lkh_var_init=1e-5)
    traj, _ = bayes_opt_wrap('manual', target, acq='ei', model='dim_ridge', rs=np.random.RandomState(13),
                             optim_config=cfg, xi=0.01, kappa=2.58, bo_it=4, rand_it=0, init_list=init_list,
                             include_init=False, source_data=sources, gp_class=factory, verbose=0, acq_one_shot='ei')
    want = z["trajectory"]
    assert traj.shape == want.shape == (4, 8)
    assert np.array_equal(traj[:, :7], want[:, :7]), (traj, want)          # identical 7-d suggestions
    assert np.allclose(traj[:, 7], want[:, 7], rtol=1e-10)


def test_oracle_counter_manual_bo_matches_reference():
    run_counter_manual_bo(OracleGPRegressor)


@pytest.mark.gpu
def test_cuda_counter_manual_bo_matches_reference(cuda_engine_mod):
    from distbo_b200.train import gp_regressor
    run_counter_manual_bo(gp_regressor)


# ---------------------------------------------------------------------------------------------------------------
# BO trajectory at the protein configuration (config 2) through the reference's driver with its deterministic
# jaccard_svm objective: tests/golden/refprotein_bo_jaccard.npz (6 + 1 ChEMBL targets, 166 bits, 2 classes, P = 23)
def run_bo_protein(make_regressor):
    from sklearn.utils import check_random_state
    from distbo_b200.data.load_datasets import protein_tasks
    from distbo_b200.model_embed import bayes_opt_wrap, set_model
    z = np.load(os.path.join(GOLD, "refprotein_bo_jaccard.npz"))
    zz = np.load(os.path.join(GOLD, "protein_subset.npz"))
    bits = np.unpackbits(zz["bits"], axis=2)[:, :, :166].astype(np.float64)
    tables = [np.hstack([np.zeros((bits.shape[1], 1)), bits[i], zz["labels"][i][:, None].astype(np.float64)])
              for i in range(bits.shape[0])]
    target, sources, _ = protein_tasks(tables, [str(n) for n in zz["names"]], target=2,
                                       random_state=check_random_state(3), test_size=0.4)
    init_list = [z["init_list%d" % i].copy() for i in range(len(sources))]
    f, _, _, _ = set_model('jaccard_svm', sources[0], None)          # the stored evaluations are this objective's
    assert f(log_C=init_list[0][0, 0]) == pytest.approx(init_list[0][0, 1], rel=1e-12)
    init = {k[len("init__"):]: z[k] for k in z.files if k.startswith("init__")}

    def factory(kernel, **kw):
        reg = make_regressor(kernel, **kw)
        reg.initial_weights = init
        return reg
    cfg = bo_optim_config(embed_op='joint', concat_data_size=True, max_size=int(z["max_size"]), n_warmup=5000,
                          n_epochs=20, stratify=True, batch_size=1000)
    traj, _ = bayes_opt_wrap('dist', target, acq='ei', model='jaccard_svm', rs=np.random.RandomState(11),
                             optim_config=cfg, xi=0.01, kappa=2.58, bo_it=4, rand_it=0, init_list=init_list,
                             include_init=False, source_data=sources, gp_class=factory, verbose=0, acq_one_shot='lcb')
    want = z["trajectory"]
    assert traj.shape == want.shape == (4, 2)
    assert np.array_equal(traj, want), (traj, want)


def test_oracle_protein_bo_trajectory_matches_reference():
    run_bo_protein(OracleGPRegressor)


@pytest.mark.gpu
def test_cuda_protein_bo_trajectory_matches_reference(cuda_engine_mod):
    from distbo_b200.train import gp_regressor
    run_bo_protein(gp_regressor)


# ---------------------------------------------------------------------------------------------------------------
# observation store vs the reference's TargetSpace (bayes_opt/target_space.py): tests/golden/reftargetspace.npz
def test_target_space_matches_reference():
    from distbo_b200.bayes_opt.target_space import TargetSpace
    z = np.load(os.path.join(GOLD, "reftargetspace.npz"))
    calls = []

    def f(a, b, c):
        calls.append((a, b, c))
        return a * b - c
    sp = TargetSpace(f, {'b': (0, 1), 'a': (-2.0, 3.0), 'c': (10, 20)}, random_state=4)
    assert [str(k) for k in sp.keys] == [str(k) for k in z["keys"]]      # dict order, not sorted (target_space.py:47)
    pts = sp.random_points(6)
    assert np.array_equal(np.asarray(pts), z["points"])
    ys = [sp.observe_point(x) for x in pts]
    assert np.array_equal(np.asarray(ys), z["ys"])
    sp.observe_point(pts[2])
    sp.add_observation(np.array([0.5, 1.0, 15.0]), 7.0)
    sp.set_bounds({'a': (-1.0, 1.0)})
    assert np.array_equal(np.asarray(sp.random_points(3)), z["more"])
    assert len(calls) == int(z["n_calls"])
    assert np.array_equal(sp.X, z["X"]) and np.array_equal(sp.Y, z["Y"]) and np.array_equal(sp.bounds, z["bounds"])
    mp = sp.max_point()
    assert mp['max_val'] == float(z["max_val"])
    assert np.array_equal(np.asarray([mp['max_params'][k] for k in sp.keys]), z["max_params"])
