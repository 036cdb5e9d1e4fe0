#!/usr/bin/env python
"""Golden vectors produced by the REFERENCE'S OWN SOURCES (run in the build container, where /root/reference exists).

hcllaw/distBO needs Python 2 + TensorFlow 1.7 (README.md:13); neither can be installed here.  Its GP path is
nevertheless plain Python over ~55 TensorFlow ops, so this script imports distBO/train/gp_regressor.py UNMODIFIED
from /root/reference with `tensorflow` bound to tools/tf1_shim.py (a lazy graph evaluated op for op with torch CPU
fp64) and drives the reference's public surrogate interface:

    gp_regressor(...).maximise_likhd(...)   -> train/train.py train_network + gp_utils.loop_batches + Adam
    .initialise_predict(); .predict_y(...)  -> log_gaus_likelihood.build_predict(_dist), embed_network, eval_network
    .similarity()

The recorded numbers (loss per epoch, fitted parameters, posterior mean / std, similarities) are what the
reference's code computes; only the TF runtime under it is substituted.  Initial network weights come from the
shim's stand-in for glorot_normal (TF's Philox stream is not reproducible) and are stored, so that the oracle and
the CUDA path start from the same values.

    python tools/make_reference_golden.py [/root/reference]      -> tests/golden/ref_*.npz
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = next((a for a in sys.argv[1:] if not a.startswith("--")), "/root/reference")
sys.path.insert(0, HERE)
import tf1_shim  # noqa: E402

tf = tf1_shim.install()
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "distBO", "train"))     # train.py:8 `from gp_utils import ...` (py2 implicit relative)
np.float = float                                                # py2-era alias (bayes_opt/target_space.py:50)
from distBO.train.gp_regressor import gp_regressor  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


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


CASES = {
    # name: (kernel, classification, embed_op, extra fit kwargs)
    "params": ("params", False, None, {}),
    # early stopping (train/train.py:53-65): with a tiny learning rate no epoch improves the loss by 1e-4, so the
    # countdown set at first_early_stop_epoch runs out after 100 more epochs, well before max_epochs
    "params_early_stop": ("params", False, None, dict(learning_rate=1e-9, max_epochs=140, first_early_stop_epoch=5)),
    "manual": ("manual", False, None, {}),
    "multi": ("multi", False, None, {}),
    "dist_none": ("dist", False, "none", dict(concat_data_size=False)),
    "dist_joint_clip": ("dist", False, "joint", dict(concat_data_size=True, gradient_clip=True)),
    "dist_joint_batched": ("dist", False, "joint", dict(concat_data_size=True, batch_size=64, stratify=False)),
    "dist_joint_stratified": ("dist", False, "joint", dict(concat_data_size=False, batch_size=64, stratify=True)),
    "dist_class": ("dist", True, "joint", dict(concat_data_size=True)),
    # stratified class mini-batches (gp_utils.py:38-84): the reference shuffles with the global `random`, seeded
    # below with 0 = the product's / oracle's shuffle_seed
    "dist_class_stratified": ("dist", True, "joint", dict(concat_data_size=True, batch_size=48, stratify=True,
                                                          max_epochs=12)),
    "dist_concat": ("dist", False, "concat", dict(concat_data_size=True)),
    "dist_nn3": ("dist", False, "nn", dict(concat_data_size=True, weight_dim_xy=[20, 12, 10])),
    "dist_nn2class": ("dist", True, "nn", dict(concat_data_size=False, weight_dim_xy=[20, 10])),
    # BLR surrogate (train/log_gaus_likelihood_feature.py), the shipped setting: weight_dim (50,50,50), num_of_rff 0
    "blr_params": ("params", False, None, dict(algorithm="BLR", weight_dim=[50, 50, 50], num_of_rff=0)),
    "blr_manual": ("manual", False, None, dict(algorithm="BLR", weight_dim=[50, 50, 50], num_of_rff=0)),
    "blr_multi": ("multi", False, None, dict(algorithm="BLR", weight_dim=[50, 50, 50], num_of_rff=0)),
    "blr_dist_joint": ("dist", False, "joint", dict(algorithm="BLR", weight_dim=[50, 50, 50], num_of_rff=0,
                                                    concat_data_size=True)),
    "blr_dist_class": ("dist", True, "joint", dict(algorithm="BLR", weight_dim=[30, 20], num_of_rff=0,
                                                   concat_data_size=True)),
    # random Fourier Matern features in front of the BLR (feature_map.py:19-26, RandomFourierMaternFeatureMapper.py):
    # the reference evaluates this block in float32, so agreement is limited by fp32 cos / matmul rounding.  Oracle
    # only: the CUDA product raises NotImplementedError for num_of_rff > 0
    "blr_rff_params": ("params", False, None, dict(algorithm="BLR", weight_dim=[30, 20], num_of_rff=40)),
    "blr_dist_nn3": ("dist", False, "nn", dict(algorithm="BLR", weight_dim=[30, 20], num_of_rff=0,
                                               concat_data_size=True, weight_dim_xy=[20, 12, 10])),
    # the reference's own data (subsets committed as tests/golden/protein_subset.npz / parkinson_subset.npz by
    # tools/make_real_fixtures.py): MACCS fingerprints of ChEMBL targets, 166 bits, 2 classes (config 2) and
    # per-patient Parkinson telemonitoring rows, 17 features (config 3)
    "real_protein": ("dist", True, "joint", dict(concat_data_size=True, batch_size=100, stratify=False)),
    "real_parkinson": ("dist", False, "joint", dict(concat_data_size=False)),
}


def real_tasks(name, rs, T, per_task, d):
    gold = os.path.join(os.path.dirname(HERE), "tests", "golden")
    if name == "real_protein":
        z = np.load(os.path.join(gold, "protein_subset.npz"))
        bits = np.unpackbits(z["bits"], axis=2)[:, :, :166].astype(np.float64)
        # unequal dataset sizes, so that the size-ratio column (concat_data_size) carries a gradient
        X = [bits[i][:360 - 23 * i] for i in range(T + 1)]
        Y = [np.eye(2)[z["labels"][i].astype(int)][:360 - 23 * i] for i in range(T + 1)]
    else:
        from sklearn import preprocessing
        z = np.load(os.path.join(gold, "parkinson_subset.npz"))
        X, Y = [], []
        for i in range(T + 1):
            pat = z["patient_%d" % i]
            X.append(preprocessing.scale(pat[:, :-2][:, 2:]))
            yy = pat[:, -1:]
            Y.append((yy - yy.min()) / (yy.max() - yy.min()))
    sizes = [per_task] * T
    theta = rs.uniform(-3, 3, size=(T * per_task, d))
    y = np.sin(theta.sum(1)) + np.repeat(rs.randn(T) * 0.2, per_task) + 0.05 * rs.randn(T * per_task)
    return theta, y[:, None], sizes, X, Y


def run_case(name):
    kernel, classification, embed_op, extra = CASES[name]
    import random
    random.seed(0)
    tf.reset_default_graph()
    tf1_shim._EAGER_SEED[0] = 0
    rs = np.random.RandomState(abs(hash(name)) % 1000 if False else sum(map(ord, name)))
    d, T, per_task = 3, 4, 20
    dx, dy = (12, 2) if classification else (6, 1)
    if name.startswith("real_"):
        theta, y, sizes, X, Y = real_tasks(name, rs, T, per_task, d)
    else:
        theta, y, sizes, X, Y = make_tasks(rs, T, per_task, d, dx, dy, 110, classification)
    model_type = "classification" if classification else "regression"
    common = dict(model_type=model_type, seed=7, float64=True)
    fit = dict(learning_rate=0.005, max_epochs=8, likhd_var_init=1e-2, first_early_stop_epoch=3)
    arrays = {"theta": theta, "y": y, "sizes": np.asarray(sizes)}
    if kernel == "dist":
        tind = np.arange(90)
        sind = [np.arange(100)] * T
        common.update(target_X=X[T], target_Y=Y[T], target_embed_ind=tind)
        fit.update(data_X=X[:T], data_Y=Y[:T], source_embed_ind=sind, sizes=sizes, embed_op=embed_op,
                   weight_dim_x=[20, 10], weight_dim_y=[20, 10], batch_size=500, stratify=False, max_size=400)
        fit.update(extra)
        for t in range(T + 1):
            arrays["X%d" % t], arrays["Y%d" % t] = X[t], Y[t]
        arrays["target_embed_ind"] = tind
        arrays["source_embed_ind"] = np.asarray(sind)
    elif kernel == "manual":
        emb = rs.rand(T + 1, 9)
        fit.update(data_embed=emb[:T], target_embed=emb[T:], sizes=sizes)
        fit.update(extra)
        arrays["embed"] = emb
    elif kernel == "multi":
        fit.update(sizes=sizes)
        fit.update(extra)
    else:
        fit.update(extra)
    g = gp_regressor(kernel, **common)
    loss = g.maximise_likhd(theta, y, **fit)
    init = {k: tf1_shim.INIT_VALUES[id(v)] for k, v in g.net.params.items()}
    final = {k: v.value.detach().numpy().copy() for k, v in g.net.params.items()}
    g.initialise_predict()
    cand = rs.uniform(-3, 3, size=(60, d))
    mean, std = g.predict_y(cand, compute_embed=True)
    out = dict(arrays)
    out.update(cand=cand, loss_log=np.asarray(g.sess.loss_log), final_loss=np.asarray(loss), pred_mean=mean, pred_std=std)
    if kernel != "params" and fit.get("algorithm", "GP") == "GP":
        out["similarity"] = np.asarray(g.similarity()[0])
    for k, v in init.items():
        out["init__" + k] = v
    for k, v in final.items():
        out["final__" + k] = v
    meta = {"kernel": kernel, "model_type": model_type, "seed": 7, "T": T,
            "fit": {k: v for k, v in fit.items() if k not in ("data_X", "data_Y", "source_embed_ind", "data_embed",
                                                               "target_embed", "sizes")}}
    out["meta"] = np.asarray(json.dumps(meta))
    np.savez_compressed(os.path.join(OUT, "ref_%s.npz" % name), **out)
    print("%-24s epochs %d  loss %.10f -> %.10f  mean[0] %.8f" % (name, len(g.sess.loss_log), g.sess.loss_log[0], loss,
                                                                 mean[0, 0]))
    g.close()




# ---------------------------------------------------------------------------------------------------------------
# BO trajectories: the reference's own BO driver (model_embed/bayes_optimise.py bayes_opt_wrap ->
# bayes_opt/bayesian_optimization.py BayesianOptimization.maximize -> helpers.py acq_max / UtilityFunction,
# target_space.py TargetSpace, model_embed/models.py set_model + n_toy) over the same shim.  The datasets come from
# this repo's py3 toy loader (the reference's imports sklearn.cluster.k_means_, removed from sklearn); they are
# inputs and are stored in the file.
BO_CASES = {
    # name: (method, embed_op, rand_it, acq_one_shot, extra optim_config)
    "bo_dist_none": ("dist", "none", 0, "lcb", {}),
    "bo_dist_joint": ("dist", "joint", 0, "lcb", {}),
    "bo_params": ("params", "none", 0, "lcb", {}),
    "bo_multi": ("multi", "none", 3, "lcb", {}),
    "bo_manual": ("manual", "none", 0, "ei", {}),
    "bo_params_polish": ("params", "none", 0, "ucb", dict(n_opt_iterations=3, n_warmup=2000)),
    # BLR surrogate: acq_max's warm start from the best previous theta of every source task (helpers.py:54-65)
    "bo_blr_dist_joint": ("dist", "joint", 0, "lcb", dict(algorithm="BLR", weight_dim=[50, 50, 50], lkh_var_init=1e-2,
                                                         n_epochs=10, first_early_stop_epoch=4, n_warmup=2000)),
    "bo_blr_multi": ("multi", "none", 3, "lcb", dict(algorithm="BLR", weight_dim=[50, 50, 50], lkh_var_init=1e-2,
                                                    n_epochs=10, first_early_stop_epoch=4, n_warmup=2000)),
}


def bo_optim_config(**over):
    cfg = dict(weight_dim=None, weight_dim_x=[20, 10], weight_dim_y=[20, 10], weight_dim_xy=None, n_warmup=20000,
               n_opt_iterations=0, n_epochs=25, lkh_var_init=1e-5, lr=0.005, float64=True, n_cpus=1, stratify=False,
               embed_op='none', concat_data_size=False, algorithm='GP', gradient_clip=False, num_of_rff=0,
               batch_size=500, first_early_stop_epoch=10, max_size=None)
    cfg.update(over)
    return cfg


def run_bo_case(name):
    import distBO.train.gp_regressor as ref_gp
    from distBO.model_embed.bayes_optimise import bayes_opt_wrap as ref_wrap
    from distBO.model_embed.models import set_model as ref_set_model
    sys.path.insert(0, os.path.dirname(HERE))
    from distbo_b200.data.toy_datasets import one_dim_split_tasks

    method, embed_op, rand_it, one_shot, extra = BO_CASES[name]
    tf.reset_default_graph()
    tf1_shim._EAGER_SEED[0] = 0
    target, sources, _ = one_dim_split_tasks(0, 3, 2, n_train=120, n_test=120)
    rs = np.random.RandomState(1)
    init_list = []
    for s in sources:                                   # 12 evaluated configurations per source task
        f, _, _, _ = ref_set_model('toy', s, None)
        th = rs.uniform(-8, 8, size=12)
        init_list.append(np.column_stack([th, [f(theta=t) for t in th]]))
    if method == "manual":                              # manual meta-features: 4 numbers per dataset (inputs)
        ers = np.random.RandomState(5)
        for dset in sources + [target]:
            dset.embed = ers.rand(4)
    # record the initial values of the surrogate's variables (close() drops the graph)
    inits = []
    orig_close = ref_gp.gp_regressor.close

    def recording_close(self):
        inits.append({k: tf1_shim.INIT_VALUES[id(v)] for k, v in self.net.params.items()})
        return orig_close(self)
    ref_gp.gp_regressor.close = recording_close

    # numpy >= 1.24 refuses the ragged np.array(list of similarity vectors) at bayes_optimise.py:60 that older numpy
    # turned into an object array: give that module a numpy proxy with the old behaviour (nothing numeric changes)
    import distBO.model_embed.bayes_optimise as ref_bo_mod

    class _NumpyCompat(object):
        def __getattr__(self, k):
            return getattr(np, k)

        @staticmethod
        def array(obj, *a, **kw):
            try:
                return np.array(obj, *a, **kw)
            except ValueError:
                out = np.empty(len(obj), dtype=object)
                for i, v in enumerate(obj):
                    out[i] = v
                return out
    ref_bo_mod.np = _NumpyCompat()
    # scipy >= 1.11 rejects the 2-D x0 that helpers.py:75-77 passes to minimize (older scipy ravelled it)
    import distBO.bayes_opt.helpers as ref_helpers
    from scipy.optimize import minimize as _minimize
    ref_helpers.minimize = lambda fun, x0, *a, **kw: _minimize(fun, np.asarray(x0).ravel(), *a, **kw)
    try:
        out = ref_wrap(method, target, acq='ei', model='toy', rs=np.random.RandomState(7),
                       optim_config=bo_optim_config(embed_op=embed_op, **extra), xi=0.01, kappa=2.58, bo_it=5,
                       rand_it=rand_it, init_list=[a.copy() for a in init_list], include_init=False,
                       source_data=sources, acq_one_shot=one_shot)
    finally:
        ref_gp.gp_regressor.close = orig_close
        ref_bo_mod.np = np
        ref_helpers.minimize = _minimize
    traj, sim = (out if method in ("dist", "manual", "multi") else (out, None))
    arrays = {"trajectory": np.asarray(traj, dtype=np.float64)}
    if sim is not None and extra.get("algorithm", "GP") == "GP":   # (BLR logs the string 'none', gp_regressor.py:221-222)
        # ragged (one more task after the first suggestion): NaN padded
        rows = [np.asarray(s_, dtype=np.float64).reshape(-1) for s_ in sim]
        arrays["similarity"] = np.asarray([np.r_[r, np.full(max(map(len, rows)) - len(r), np.nan)] for r in rows])
    for k, v in inits[0].items():
        arrays["init__" + k] = v
    for i, a in enumerate(init_list):
        arrays["init_list%d" % i] = a
    if method == "manual":
        arrays["embed"] = np.asarray([d.embed for d in sources + [target]])
    arrays["meta"] = np.asarray(json.dumps({"method": method, "embed_op": embed_op, "rand_it": rand_it,
                                           "acq_one_shot": one_shot, "extra": extra}))
    np.savez_compressed(os.path.join(OUT, "refbo_%s.npz" % name[3:]), **arrays)
    print("%-24s suggestions %s" % (name, np.array2string(np.asarray(traj)[:, 0], precision=6)))




# ---------------------------------------------------------------------------------------------------------------
# fp32 embeddings: the reference's distbo_net / nn_net / sparse_joint_op (train/distbo_net.py, base.py) evaluated in
# tf.float32, the arithmetic type of its CLI default (train_test.py:87) and of north_star's embedding tolerance
# (1e-5 relative).  Inputs: the committed subsets of the reference's own data.
def run_embed_f32():
    from distBO.train.distbo_net import distbo_net
    from distBO.train.base import sum_matrix as ref_sum_matrix
    gold = os.path.join(os.path.dirname(HERE), "tests", "golden")
    rs = np.random.RandomState(11)
    out = {}

    def weights(din, h, p, name):
        return {"weights_%s_1" % name: (rs.randn(din, h) * np.sqrt(2.0 / (din + h))).astype(np.float32),
                "bias_%s_1" % name: (rs.randn(h) * 0.1).astype(np.float32),
                "weights_%s_2" % name: (rs.randn(h, p) * np.sqrt(2.0 / (h + p))).astype(np.float32)}

    def embed(tag, X, Y, sizes, w, embed_op, model_type):
        tf.reset_default_graph()
        params = {k: tf.constant(v, dtype=tf.float32) for k, v in w.items()}
        sm = ref_sum_matrix(None, np.asarray(sizes).reshape(-1, 1))
        sm_t = type(sm)(indices=tf.constant(sm.indices, dtype=tf.int64), values=tf.constant(sm.values, dtype=tf.float32),
                        dense_shape=tf.constant(sm.dense_shape, dtype=tf.int64))
        E = distbo_net(tf.constant(X, dtype=tf.float32), tf.constant(Y, dtype=tf.float32),
                       tf.constant(np.asarray(sizes, dtype=np.float32).reshape(-1, 1), dtype=tf.float32), params,
                       embed_op=embed_op, model_type=model_type, sum_matrix=sm_t, dtype=tf.float32)
        e = tf.Session().run(E)
        assert e.dtype == np.float32
        out[tag + "__E"] = e
        out[tag + "__X"] = np.ascontiguousarray(X, dtype=np.float32)
        out[tag + "__Y"] = np.ascontiguousarray(Y, dtype=np.float32)
        out[tag + "__sizes"] = np.asarray(sizes)
        for k, v in w.items():
            out[tag + "__" + k] = v
        print("%-28s E %s  |E|max %.6f" % (tag, e.shape, np.abs(e).max()))

    z = np.load(os.path.join(gold, "protein_subset.npz"))
    bits = np.unpackbits(z["bits"], axis=2)[:, :, :166].astype(np.float32)
    sizes = [360 - 37 * i for i in range(7)]            # ragged: 360, 323, ..., 138 rows
    X = np.vstack([bits[i][:n] for i, n in enumerate(sizes)])
    Y = np.vstack([np.eye(2, dtype=np.float32)[z["labels"][i].astype(int)][:n] for i, n in enumerate(sizes)])
    embed("protein_class", X, Y, sizes, weights(166, 20, 10, "x"), "joint", "classification")

    from sklearn import preprocessing
    zp = np.load(os.path.join(gold, "parkinson_subset.npz"))
    Xs, Ys = [], []
    for i in range(10):
        pat = zp["patient_%d" % i]
        Xs.append(preprocessing.scale(pat[:, :-2][:, 2:]).astype(np.float32))
        yy = pat[:, -1:]
        Ys.append(((yy - yy.min()) / (yy.max() - yy.min())).astype(np.float32))
    sizes = [len(x) for x in Xs]
    w = weights(17, 20, 10, "x")
    embed("parkinson_none", np.vstack(Xs), np.vstack(Ys), sizes, w, "none", "regression")
    w2 = dict(w)
    w2.update(weights(1, 20, 10, "y"))
    embed("parkinson_joint", np.vstack(Xs), np.vstack(Ys), sizes, w2, "joint", "regression")
    np.savez_compressed(os.path.join(OUT, "refembed_f32.npz"), **out)




# ---------------------------------------------------------------------------------------------------------------
# acquisition functions: the reference's UtilityFunction (bayes_opt/helpers.py:114-177) on fixed (mean, std) pairs
def run_acquisition():
    from distBO.bayes_opt.helpers import UtilityFunction
    rs = np.random.RandomState(3)
    M = 400
    mean = rs.randn(M, 1) * 0.7
    std = np.abs(rs.randn(M, 1)) * 0.5 + 1e-3
    std[:5] = 1e-16                                        # the clamped-variance regime (gp_regressor.py:197)
    x = rs.rand(M, 2)

    class StubGP(object):
        def predict_y(self, x_, **kw):
            return mean[:len(x_)], std[:len(x_)]
    out = {"mean": mean, "std": std, "y_max": np.asarray(0.4), "xi": np.asarray(0.01), "kappa": np.asarray(2.58)}
    for kind in ("ei", "ucb", "lcb", "poi"):
        u = UtilityFunction(kind=kind, kappa=2.58, xi=0.01)
        out[kind] = np.asarray(u.utility(x, StubGP(), 0.4), dtype=np.float64).reshape(-1)
        print("%-8s max %.8f argmax %d" % (kind, out[kind].max(), int(out[kind].argmax())))
    np.savez_compressed(os.path.join(OUT, "refacq.npz"), **out)




# ---------------------------------------------------------------------------------------------------------------
# manual meta-features: the reference's model_embed/meta_fea.py meta() on datasets built from the committed subsets
# of its own data (sklearn.datasets.load_boston, imported but unused by that module, no longer exists: stubbed).
def meta_datasets():
    sys.path.insert(0, os.path.dirname(HERE))
    from sklearn.utils import check_random_state
    from distbo_b200.data.load_datasets import parkinson_tasks, protein_tasks
    gold = os.path.join(os.path.dirname(HERE), "tests", "golden")
    z = np.load(os.path.join(gold, "protein_subset.npz"))
    bits = np.unpackbits(z["bits"], axis=2)[:, :, :166].astype(np.float64)
    tables = [np.hstack([np.zeros((bits.shape[1], 1)), bits[i], z["labels"][i][:, None].astype(np.float64)])
              for i in range(bits.shape[0])]
    pt, ps, _ = protein_tasks(tables, [str(n) for n in z["names"]], target=2, random_state=check_random_state(3),
                              test_size=0.4)
    zp = np.load(os.path.join(gold, "parkinson_subset.npz"))
    kt, ks, _ = parkinson_tasks([zp["patient_%d" % i] for i in range(10)], target=4,
                                random_state=check_random_state(1), test_size=0.4)
    return {"protein": ([pt] + ps, "forest"), "parkinson": ([kt] + ks, "ridge")}


def run_meta_features():
    import sklearn.datasets
    sklearn.datasets.__dict__.setdefault("load_boston", lambda *a, **k: None)   # removed in scikit-learn 1.2
    from distBO.model_embed.meta_fea import meta as ref_meta
    import contextlib
    import io
    out = {}
    for tag, (dsets, model) in meta_datasets().items():
        max_size = max(len(d.train_x) for d in dsets)
        with contextlib.redirect_stdout(io.StringIO()):
            scaler = ref_meta(dsets, model, max_size=max_size, random_state=5)
        out[tag] = np.asarray([d.embed for d in dsets])
        out[tag + "__data_min"] = scaler.data_min_
        out[tag + "__data_max"] = scaler.data_max_
        print("%-12s meta-features %s" % (tag, out[tag].shape))
    np.savez_compressed(os.path.join(OUT, "refmeta.npz"), **out)




# ---------------------------------------------------------------------------------------------------------------
# data loaders and the train/test container: the reference's data/load_datasets.py load_protein / load_parkinson and
# data/data.py data (py2 `range(...).remove` at load_datasets.py:42-43,82-83 -> the module gets a list-returning range)
def _digest(d):
    return {"shape": np.asarray(d.train_x.shape + d.test_x.shape), "colsum": d.train_x.sum(0), "ysum": d.train_y.sum(0),
            "test_colsum": d.test_x.sum(0), "row0": d.train_x[0], "embed_ind": np.asarray(list(d.embed_ind))[:64],
            "n_embed": np.asarray(len(list(d.embed_ind)))}


def run_loaders():
    import builtins
    import contextlib
    import io
    from sklearn.utils import check_random_state
    import distBO.data.load_datasets as ref_ld
    from distBO.data.data import data as ref_data
    ref_ld.range = lambda *a: list(builtins.range(*a))
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        pt, ps, pn = ref_ld.load_protein(2, random_state=check_random_state(3), test_size=0.4)
        kt, ks, kn = ref_ld.load_parkinson(4, random_state=check_random_state(1), test_size=0.4)
    for tag, dsets, names in (("protein", [pt] + ps, pn), ("parkinson", [kt] + ks, kn)):
        out[tag + "__names"] = np.asarray([str(n) for n in names])
        for i, d in enumerate(dsets):
            for k, v in _digest(d).items():
                out["%s__%d__%s" % (tag, i, k)] = v
        print("%-10s %d datasets, target %s" % (tag, len(dsets), names[0]))
    # embed_ind subsampling of the container (data.py:40-62): regression and class-proportional classification
    rs = np.random.RandomState(0)
    x, yr = rs.randn(300, 4), rs.rand(300)
    yc = (rs.rand(300) < 0.3).astype(float)
    d1 = ref_data(x[:200], x[200:], yr[:200], yr[200:], embed_size=50, random_state=check_random_state(9))
    d2 = ref_data(x[:200], x[200:], yc[:200, None], yc[200:, None], embed_size=50, random_state=check_random_state(9),
                  prob_type='classification')
    out["container__x"], out["container__yr"], out["container__yc"] = x, yr, yc
    out["container__reg_embed_ind"] = np.asarray(list(d1.embed_ind))
    out["container__cls_embed_ind"] = np.asarray(list(d2.embed_ind))
    out["container__cls_train_y"] = d2.train_y
    np.savez_compressed(os.path.join(OUT, "refdata.npz"), **out)




# ---------------------------------------------------------------------------------------------------------------
# toy task family of config 1: the reference's data/generate_data.py generate_data (dataset 'one_dim_split') and
# toy_datasets.one_d_split (its import of sklearn.cluster.k_means_._init_centroids, unused here, is stubbed)
def run_toy_tasks():
    import types as _types
    import sklearn.datasets
    sklearn.datasets.__dict__.setdefault("load_boston", lambda *a, **k: None)
    km = _types.ModuleType("sklearn.cluster.k_means_")
    km._init_centroids = None
    sys.modules.setdefault("sklearn.cluster.k_means_", km)
    from distBO.data.generate_data import generate_data
    out = {}
    for tag, (seed, fake, true, n) in {"a": (0, 3, 2, 120), "b": (7, 12, 3, 500)}.items():
        args = _types.SimpleNamespace(dataset='one_dim_split', data_seed=seed, dim=None, preprocess='standardise',
                                      max_embed_size=10000, n_train=n, n_test=n, fake_source_num=fake,
                                      true_source_num=true, mu_sd=1.0)
        res = generate_data(args)
        target, sources, supp = res[0], res[1], res[2]
        out[tag + "__cfg"] = np.asarray([seed, fake, true, n])
        out[tag + "__supp"] = np.asarray(supp, dtype=np.float64)
        for i, d in enumerate([target] + list(sources)):
            out["%s__%d__train_x" % (tag, i)] = d.train_x
            out["%s__%d__test_sum" % (tag, i)] = np.asarray(d.test_x.sum())
        print("toy %s: %d sources, mu_target %.6f" % (tag, len(sources), supp[0]))
    np.savez_compressed(os.path.join(OUT, "reftoy.npz"), **out)




# ---------------------------------------------------------------------------------------------------------------
# BO trajectory at the Parkinson configuration (config 3): the reference's driver and its own ridge objective
# (model_embed/models.py ridge: 200 random Fourier features under np.random.seed(23)) on the committed subset of its
# patient files, joint embedding (17 features x 1 label -> P = 100).
def run_bo_parkinson():
    import distBO.train.gp_regressor as ref_gp
    import distBO.model_embed.bayes_optimise as ref_bo_mod
    from distBO.model_embed.models import set_model as ref_set_model
    from sklearn.utils import check_random_state
    sys.path.insert(0, os.path.dirname(HERE))
    from distbo_b200.data.load_datasets import parkinson_tasks
    tf.reset_default_graph()
    tf1_shim._EAGER_SEED[0] = 0
    zp = np.load(os.path.join(os.path.dirname(HERE), "tests", "golden", "parkinson_subset.npz"))
    target, sources, _ = parkinson_tasks([zp["patient_%d" % i] for i in range(10)], target=4,
                                         random_state=check_random_state(1), test_size=0.4)
    sources = sources[:5]
    rs = np.random.RandomState(6)
    init_list = []
    for s_ in sources:                                   # 6 random-search evaluations per source patient
        f, bounds, _, _ = ref_set_model('ridge', s_, None)
        names = sorted(bounds)
        rows = []
        for _ in range(6):
            th = {k: rs.uniform(*bounds[k]) for k in names}
            rows.append([th[k] for k in names] + [f(**th)])
        init_list.append(np.array(rows))
    inits = []
    orig_close = ref_gp.gp_regressor.close

    def recording_close(self):
        inits.append({k: tf1_shim.INIT_VALUES[id(v)] for k, v in self.net.params.items()})
        return orig_close(self)
    ref_gp.gp_regressor.close = recording_close

    class _NumpyCompat(object):
        def __getattr__(self, k):
            return getattr(np, k)

        @staticmethod
        def array(obj, *a, **kw):
            try:
                return np.array(obj, *a, **kw)
            except ValueError:
                out = np.empty(len(obj), dtype=object)
                for i, v in enumerate(obj):
                    out[i] = v
                return out
    ref_bo_mod.np = _NumpyCompat()
    cfg = bo_optim_config(embed_op='joint', n_warmup=5000, n_epochs=20, stratify=False, batch_size=1000)
    try:
        traj, sim = ref_bo_mod.bayes_opt_wrap('dist', target, acq='ei', model='ridge', rs=np.random.RandomState(11),
                                              optim_config=cfg, xi=0.01, kappa=2.58, bo_it=4, rand_it=0,
                                              init_list=[a.copy() for a in init_list], include_init=False,
                                              source_data=sources, acq_one_shot='ei')
    finally:
        ref_gp.gp_regressor.close = orig_close
        ref_bo_mod.np = np
    arrays = {"trajectory": np.asarray(traj, dtype=np.float64)}
    for k, v in inits[0].items():
        arrays["init__" + k] = v
    for i, a in enumerate(init_list):
        arrays["init_list%d" % i] = a
    np.savez_compressed(os.path.join(OUT, "refpark_bo_ridge.npz"), **arrays)
    print("bo_parkinson suggestions\n%s" % np.array2string(np.asarray(traj), precision=6))




# ---------------------------------------------------------------------------------------------------------------
# The headline configuration at FULL size (BASELINE config 5: N = 8192 observations = 64 tasks x 128, d_theta = 8,
# protein-like datasets of 166 binary features / 2 classes, P = 23): two Adam epochs and the posterior on 256
# candidates through the reference's gp_regressor.  Inputs come from bench.make_workload (seeded), so the file holds
# outputs only.
def config5_inputs(lkh_var=1e-2, n_cand=256):
    sys.path.insert(0, os.path.dirname(HERE))
    import bench
    work = bench.make_workload(bench.N_OBS, n_cand, 1000)
    T = bench.T_SRC
    # unequal dataset sizes (1000, 997, ...), so that the size-ratio column carries a gradient
    X = [work["data_X"][t][:1000 - 3 * t] for t in range(T + 1)]
    Y = [work["data_Y"][t][:1000 - 3 * t] for t in range(T + 1)]
    common = dict(model_type="classification", seed=0, float64=True, target_X=X[T], target_Y=Y[T],
                  target_embed_ind=np.arange(len(X[T])))
    fit = dict(data_X=X[:T], data_Y=Y[:T], source_embed_ind=[np.arange(len(x)) for x in X[:T]], embed_op="joint",
               algorithm="GP", weight_dim_x=[20, 10], weight_dim_y=[20, 10], batch_size=1000, sizes=work["sizes"],
               concat_data_size=True, learning_rate=0.005, max_epochs=2, max_size=1000, likhd_var_init=lkh_var,
               stratify=False, gradient_clip=False, first_early_stop_epoch=10 ** 9)
    return work, common, fit


def run_config5(lkh_var=1e-2, n_cand=256, fname="refconfig5.npz"):
    """lkh_var 1e-2 = the well-conditioned variant of SURVEY 8d; 1e-5 = the reference's default noise
    (experiments/train_test.py:79), written by run_config5_default_noise with 4096 posterior candidates and
    cond(K_sigma) of the fitted surrogate."""
    import time
    tf.reset_default_graph()
    tf1_shim._EAGER_SEED[0] = 0
    work, common, fit = config5_inputs(lkh_var, n_cand)
    t0 = time.time()
    g = gp_regressor("dist", **common)
    loss = g.maximise_likhd(work["theta"], work["y"][:, None], **fit)
    init = {k: tf1_shim.INIT_VALUES[id(v)] for k, v in g.net.params.items()}
    final = {k: v.value.detach().numpy().copy() for k, v in g.net.params.items()}
    g.initialise_predict()
    mean, std = g.predict_y(work["cand"], compute_embed=True)
    out = {"loss_log": np.asarray(g.sess.loss_log), "final_loss": np.asarray(loss), "pred_mean": mean, "pred_std": std,
           "similarity": np.asarray(g.similarity()[0])}
    # the acquisition sweep of the BO loop (HOT LOOP 2): the reference's acq_max over the first 4096 rows of the
    # bench's candidate stream (RandomState(2), U(-5, 5)^8; the full 2^20 would need 69 GB per [N, M] temporary)
    from distBO.bayes_opt.helpers import UtilityFunction, acq_max
    util = UtilityFunction(kind='ei', kappa=2.58, xi=0.01)
    bounds = np.array([[-5.0, 5.0]] * 8)
    x_max = acq_max(ac=util.utility, gp=g, y_max=work["y_max"], bounds=bounds, random_state=np.random.RandomState(2),
                    n_warmup=4096, n_iter=0)
    cand = np.random.RandomState(2).uniform(-5.0, 5.0, size=(4096, 8))
    out["acq_x_max"] = np.asarray(x_max)
    out["acq_ei"] = np.asarray(util.utility(cand, g, work["y_max"]), dtype=np.float64).reshape(-1)
    for k, v in init.items():
        out["init__" + k] = v
    for k, v in final.items():
        out["final__" + k] = v
    # conditioning of the matrix the reference factorises (log_gaus_likelihood.py:122-123) at the fitted parameters:
    # K_sigma evaluated by the reference's own graph on the training feed
    # (the feed of the last predict run = full sample sets + the training observations)
    try:
        ksig = g.sess.run(g.net.k_sigma_sq, feed_dict=g.sess.last_feed)
        ev = np.linalg.eigvalsh(np.asarray(ksig))
        out["cond_K_sigma"] = np.asarray(ev[-1] / ev[0])
        out["eig_min_max"] = np.asarray([ev[0], ev[-1]])
    except Exception as e:                      # pragma: no cover - diagnostic only
        print("cond(K) not recorded:", repr(e))
    out["lkh_var_init"] = np.asarray(lkh_var)
    np.savez_compressed(os.path.join(OUT, fname), **out)
    print("config5 N=%d sigma^2=%g: loss %s, mean[0] %.10f, cond %s, %.0f s" % (
        len(work["theta"]), lkh_var, g.sess.loss_log, mean[0, 0], out.get("cond_K_sigma"), time.time() - t0))
    g.close()


def run_config5_default_noise():
    run_config5(lkh_var=1e-5, n_cand=4096, fname="refconfig5_s1e-5.npz")




# ---------------------------------------------------------------------------------------------------------------
# Config 4 (counter_manual): the reference's whole manualBO pipeline - generate_data('counter_manual') datasets,
# meta_fea.meta() features, the 7-d dim_ridge objective and the manual-kernel BO loop of bayes_opt_wrap.
def run_bo_counter_manual():
    import contextlib
    import io
    import types as _types
    import sklearn.datasets
    sklearn.datasets.__dict__.setdefault("load_boston", lambda *a, **k: None)
    km = _types.ModuleType("sklearn.cluster.k_means_")
    km._init_centroids = None
    sys.modules.setdefault("sklearn.cluster.k_means_", km)
    import distBO.train.gp_regressor as ref_gp
    import distBO.model_embed.bayes_optimise as ref_bo_mod
    from distBO.data.generate_data import generate_data
    from distBO.model_embed.meta_fea import meta as ref_meta
    from distBO.model_embed.models import set_model as ref_set_model
    tf.reset_default_graph()
    tf1_shim._EAGER_SEED[0] = 0
    args = _types.SimpleNamespace(dataset='counter_manual', data_seed=3, dim=6, preprocess='standardise',
                                  max_embed_size=10000, n_train=300, n_test=300, noise=0.5)
    with contextlib.redirect_stdout(io.StringIO()):
        res = generate_data(args)
        target, sources = res[0], list(res[1])
        ref_meta([target] + sources, 'dim_ridge', max_size=None, random_state=5)
    rs = np.random.RandomState(8)
    init_list = []
    for s_ in sources:                                   # 10 random-search evaluations per source task
        f, bounds, _, _ = ref_set_model('dim_ridge', s_, None)
        names = ['log_alpha'] + ['log_bw%d' % i for i in range(1, 7)]
        rows = []
        for _ in range(10):                              # drawn from the well-behaved corner of the search box
            th = {k: rs.uniform(np.log(1e-3), np.log(0.1)) if k == 'log_alpha' else rs.uniform(0.0, np.log(16.0))
                  for k in names}
            with contextlib.redirect_stdout(io.StringIO()):
                rows.append([th[k] for k in names] + [f(**th)])
        init_list.append(np.array(rows))
    inits = []
    orig_close = ref_gp.gp_regressor.close

    def recording_close(self):
        inits.append({k: tf1_shim.INIT_VALUES[id(v)] for k, v in self.net.params.items()})
        return orig_close(self)
    ref_gp.gp_regressor.close = recording_close

    class _NumpyCompat(object):          # ragged np.array(...) at bayes_optimise.py:60, see run_bo_case
        def __getattr__(self, k):
            return getattr(np, k)

        @staticmethod
        def array(obj, *a, **kw):
            try:
                return np.array(obj, *a, **kw)
            except ValueError:
                out = np.empty(len(obj), dtype=object)
                for i, v in enumerate(obj):
                    out[i] = v
                return out
    ref_bo_mod.np = _NumpyCompat()
    cfg = bo_optim_config(n_warmup=5000, n_epochs=20, lkh_var_init=1e-5)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            traj, sim = ref_bo_mod.bayes_opt_wrap('manual', target, acq='ei', model='dim_ridge',
                                                  rs=np.random.RandomState(13), optim_config=cfg, xi=0.01, kappa=2.58,
                                                  bo_it=4, rand_it=0, init_list=[a.copy() for a in init_list],
                                                  include_init=False, source_data=sources, acq_one_shot='ei')
    finally:
        ref_gp.gp_regressor.close = orig_close
        ref_bo_mod.np = np
    arrays = {"trajectory": np.asarray(traj, dtype=np.float64),
              "embed": np.asarray([d.embed for d in [target] + sources]),
              "target_train_x": target.train_x, "target_train_y": target.train_y,
              "source0_train_x": sources[0].train_x}
    for k, v in inits[0].items():
        arrays["init__" + k] = v
    for i, a in enumerate(init_list):
        arrays["init_list%d" % i] = a
    np.savez_compressed(os.path.join(OUT, "refcounter_manual_bo.npz"), **arrays)
    print("counter_manual: %d meta-features, suggestions\n%s" % (arrays["embed"].shape[1],
                                                                np.array2string(np.asarray(traj), precision=4)))




# ---------------------------------------------------------------------------------------------------------------
# BO trajectory at the protein configuration (config 2): 6 source ChEMBL targets + 1 target (committed subset of the
# reference's files), 166 binary features, 2 classes, P = 10*2 + 2 + 1 = 23, through the reference's driver with its
# deterministic jaccard_svm objective (the default random forest is unseeded, models.py:225-237).
def run_bo_protein():
    import contextlib
    import io
    import distBO.train.gp_regressor as ref_gp
    import distBO.model_embed.bayes_optimise as ref_bo_mod
    from distBO.model_embed.models import set_model as ref_set_model
    from sklearn.utils import check_random_state
    sys.path.insert(0, os.path.dirname(HERE))
    from distbo_b200.data.load_datasets import protein_tasks
    tf.reset_default_graph()
    tf1_shim._EAGER_SEED[0] = 0
    z = np.load(os.path.join(os.path.dirname(HERE), "tests", "golden", "protein_subset.npz"))
    bits = np.unpackbits(z["bits"], axis=2)[:, :, :166].astype(np.float64)
    tables = [np.hstack([np.zeros((bits.shape[1], 1)), bits[i], z["labels"][i][:, None].astype(np.float64)])
              for i in range(bits.shape[0])]
    target, sources, _ = protein_tasks(tables, [str(n) for n in z["names"]], target=2,
                                       random_state=check_random_state(3), test_size=0.4)
    rs = np.random.RandomState(5)
    init_list = []
    for s_ in sources:                                   # 8 random-search evaluations per source target
        f, bounds, _, _ = ref_set_model('jaccard_svm', s_, None)
        rows = []
        for _ in range(8):
            c = rs.uniform(*bounds['log_C'])
            with contextlib.redirect_stdout(io.StringIO()):
                rows.append([c, f(log_C=c)])
        init_list.append(np.array(rows))
    inits = []
    orig_close = ref_gp.gp_regressor.close

    def recording_close(self):
        inits.append({k: tf1_shim.INIT_VALUES[id(v)] for k, v in self.net.params.items()})
        return orig_close(self)
    ref_gp.gp_regressor.close = recording_close

    class _NumpyCompat(object):          # ragged np.array(...) at bayes_optimise.py:60, see run_bo_case
        def __getattr__(self, k):
            return getattr(np, k)

        @staticmethod
        def array(obj, *a, **kw):
            try:
                return np.array(obj, *a, **kw)
            except ValueError:
                out = np.empty(len(obj), dtype=object)
                for i, v in enumerate(obj):
                    out[i] = v
                return out
    ref_bo_mod.np = _NumpyCompat()
    max_size = max([len(target.train_x)] + [len(s_.train_x) for s_ in sources])
    cfg = bo_optim_config(embed_op='joint', concat_data_size=True, max_size=max_size, n_warmup=5000, n_epochs=20,
                          stratify=True, batch_size=1000)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            traj, sim = ref_bo_mod.bayes_opt_wrap('dist', target, acq='ei', model='jaccard_svm',
                                                  rs=np.random.RandomState(11), optim_config=cfg, xi=0.01, kappa=2.58,
                                                  bo_it=4, rand_it=0, init_list=[a.copy() for a in init_list],
                                                  include_init=False, source_data=sources, acq_one_shot='lcb')
    finally:
        ref_gp.gp_regressor.close = orig_close
        ref_bo_mod.np = np
    arrays = {"trajectory": np.asarray(traj, dtype=np.float64), "max_size": np.asarray(max_size)}
    for k, v in inits[0].items():
        arrays["init__" + k] = v
    for i, a in enumerate(init_list):
        arrays["init_list%d" % i] = a
    np.savez_compressed(os.path.join(OUT, "refprotein_bo_jaccard.npz"), **arrays)
    print("bo_protein suggestions\n%s" % np.array2string(np.asarray(traj), precision=6))


# ---------------------------------------------------------------------------------------------------------------
# observation store: the reference's bayes_opt/target_space.py TargetSpace (np.float alias restored above)
def run_target_space():
    from distBO.bayes_opt.target_space import TargetSpace
    calls = []

    def f(a, b, c):
        calls.append((a, b, c))
        return a * b - c
    sp = TargetSpace(f, {'b': (0, 1), 'a': (-2.0, 3.0), 'c': (10, 20)}, random_state=4)
    pts = sp.random_points(6)
    ys = [sp.observe_point(x) for x in pts]
    sp.observe_point(pts[2])                               # cached: no second evaluation
    sp.add_observation(np.array([0.5, 1.0, 15.0]), 7.0)
    sp.set_bounds({'a': (-1.0, 1.0)})
    more = sp.random_points(3)
    mp = sp.max_point()
    out = {"keys": np.asarray(sp.keys), "bounds": sp.bounds, "points": np.asarray(pts), "ys": np.asarray(ys),
           "n_calls": np.asarray(len(calls)), "X": sp.X, "Y": sp.Y, "more": np.asarray(more),
           "max_val": np.asarray(mp['max_val']), "max_params": np.asarray([mp['max_params'][k] for k in sp.keys])}
    np.savez_compressed(os.path.join(OUT, "reftargetspace.npz"), **out)
    print("target space: keys %s, %d evaluations, %d rows" % (list(sp.keys), len(calls), len(sp.X)))


def main():
    os.makedirs(OUT, exist_ok=True)
    only = [a[len("--only="):] for a in sys.argv[1:] if a.startswith("--only=")]
    if only:                                    # e.g. --only=run_config5_default_noise
        for fn in only:
            globals()[fn]()
        return
    if "--bo-only" not in sys.argv:
        for name in CASES:
            run_case(name)
    for name in BO_CASES:
        run_bo_case(name)
    run_embed_f32()
    run_acquisition()
    run_meta_features()
    run_loaders()
    run_toy_tasks()
    run_bo_parkinson()
    run_config5()
    run_config5_default_noise()
    run_bo_counter_manual()
    run_bo_protein()
    run_target_space()


if __name__ == "__main__":
    main()
