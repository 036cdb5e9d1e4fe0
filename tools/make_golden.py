#!/usr/bin/env python
"""Generates tests/golden/*.npz from the CPU oracle with fixed seeds (the reference itself cannot run
here and ships no recorded outputs - SURVEY 8c).  Small cases only; inputs AND outputs are stored so
the GPU tests need nothing but the file.   python tools/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gp_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def stage_case(name, seed, T, per, d, dx, dy, rows, classification, embed_op, concat):
    rs = np.random.RandomState(seed)
    N = T * per
    sizes = [per] * T
    theta = rs.randn(N, d)
    y = rs.randn(N, 1)
    row_counts = [rows + 13 * t for t in range(T + 1)]
    if classification:
        X = [(rs.rand(n, dx) < 0.3).astype(np.float64) for n in row_counts]
        Y = [np.eye(dy)[(rs.rand(n) < 0.4).astype(int)] for n in row_counts]
        model_type = "classification"
    else:
        X = [rs.randn(n, dx) + 0.2 * t for t, n in enumerate(row_counts)]
        Y = [rs.rand(n, dy) for n in row_counts]
        model_type = "regression"
    p = O.build_params(d, "dist", in_dim_x=dx, in_dim_y=dy, weight_dim_x=[20, 10], weight_dim_y=[20, 10],
                       model_type=model_type, embed_op=embed_op, concat_data_size=concat, likhd_var_init=1e-2, seed=seed)
    p["bias_x_1"] = rs.randn(20) * 0.1
    if "bias_y_1" in p:
        p["bias_y_1"] = rs.randn(20) * 0.1
    p["log_bw"] = rs.randn(d) * 0.3
    p["log_bw_embed"] = rs.randn(len(p["log_bw_embed"])) * 0.3
    p["log_scale"] = np.array([0.2])
    p["mean"] = np.array([-0.1])
    max_size = max(row_counts) if concat else None
    batch = O.TrainBatch(h_params=theta, obs=y, sizes=sizes, batch_X=np.vstack(X[:T]), batch_y=np.vstack(Y[:T]),
                         embed_sizes=row_counts[:T], data_sizes=row_counts[:T], max_size=max_size)
    E = O.dist_features(O.NP, p, batch.batch_X, batch.batch_y, sizes, row_counts[:T], data_sizes=row_counts[:T],
                        max_size=max_size, embed_op=embed_op, model_type=model_type, return_pooled=True)
    loss, L, k, ks = O.neg_log_marg_likhd(O.NP, p, "dist", batch, embed_op=embed_op, model_type=model_type,
                                          return_all=True)
    _, grads = O.loss_and_grads(p, "dist", batch, embed_op=embed_op, model_type=model_type)
    k_dist, k_dist_pred = O.embed_grams(O.NP, p, batch, X[T], Y[T], [row_counts[T]], [row_counts[T]], 1,
                                        embed_op=embed_op, model_type=model_type)
    cand = rs.randn(400, d) * 1.5
    mean, var = O.predict(O.NP, p, "dist", batch, cand, k_dist=k_dist, k_dist_pred=k_dist_pred)
    std = O.std_from_var(var)
    y_max = float(y.max())
    out = dict(theta=theta, y=y, sizes=np.array(sizes), row_counts=np.array(row_counts),
               X=np.vstack(X), Y=np.vstack(Y), cand=cand, y_max=y_max, xi=0.01, kappa=2.58,
               classification=classification, embed_op=embed_op, concat=concat,
               E=E, K_sigma=ks, L=L, nll=float(loss), kvec=k_dist_pred[:, 0], mean=mean, var=var,
               ei=O.acq_ei(mean, std, y_max, 0.01), ucb=O.acq_ucb(mean, std, 2.58), lcb=O.acq_lcb(mean, std, 2.58))
    for kname, v in p.items():
        out["p_" + kname] = np.asarray(v)
    for kname, v in grads.items():
        out["g_" + kname] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "N=%d" % N, "nll=%.6f" % float(loss), "argmax(ei)=%d" % int(np.argmax(out["ei"])))


def multi_case(name, seed, T, per, d):
    """kernel 'multi' (train/kernels.py:32-55): loss, gradients, posterior, similarity."""
    rs = np.random.RandomState(seed)
    N, sizes = T * per, [per] * T
    theta, y = rs.randn(N, d), rs.randn(N, 1)
    p = O.build_params(d, "multi", n_tasks=T, likhd_var_init=1e-2, seed=seed)
    p["log_bw"] = rs.randn(d) * 0.3
    p["log_scale"] = np.array([0.1])
    p["mean"] = np.array([0.05])
    batch = O.TrainBatch(h_params=theta, obs=y, sizes=sizes)
    loss, grads = O.loss_and_grads(p, "multi", batch)
    cand = rs.randn(300, d) * 1.5
    mean, var = O.predict(O.NP, p, "multi", batch, cand)
    out = dict(theta=theta, y=y, sizes=np.array(sizes), cand=cand, nll=loss, mean=mean, var=var,
               KE=O.task_kernel_matrix(O.NP, p, T), sim=O.similarity(O.NP, p, "multi", batch))
    out.update({"p_" + k: np.asarray(v) for k, v in p.items()})
    out.update({"g_" + k: np.asarray(v) for k, v in grads.items()})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "N=%d" % N, "nll=%.6f" % loss)


def blr_case(name, seed, T, per, d, P, weight_dim):
    """algorithm='BLR' on meta-features (train/log_gaus_likelihood_feature.py): loss, gradients, posterior."""
    from oracle import blr_oracle as BO
    rs = np.random.RandomState(seed)
    N, sizes = T * per, [per] * T
    theta, y = rs.randn(N, d), rs.randn(N, 1)
    meta = rs.rand(T + 1, P)
    p, rff_in = BO.build_params_blr(d, "manual", weight_dim, in_dim_meta=P, likhd_var_init=2e-2, seed=seed)
    for k in p:
        if k.startswith("bias"):
            p[k] = rs.randn(*p[k].shape) * 0.2
    p["log_alpha"] = np.asarray(0.25)
    batch = O.TrainBatch(h_params=theta, obs=y, sizes=sizes, embed=meta[:T])
    loss, grads = BO.blr_loss_and_grads(p, "manual", batch, rff_in, 0)
    cand = rs.randn(300, d) * 1.5
    phi = BO.nn_features(O.NP, p, BO.train_feature(O.NP, p, "manual", batch), rff_in, 0)
    phi_c = BO.nn_features(O.NP, p, np.concatenate([np.repeat(meta[T:], 300, 0), cand], 1), rff_in, 0)
    mean, var = BO.blr_predict(O.NP, p, phi, y, phi_c)
    out = dict(theta=theta, y=y, sizes=np.array(sizes), meta=meta, cand=cand, loss=loss, mean=mean, var=var,
               weight_dim=np.array(weight_dim))
    out.update({"p_" + k: np.asarray(v) for k, v in p.items()})
    out.update({"g_" + k: np.asarray(v) for k, v in grads.items()})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "N=%d" % N, "loss=%.6f" % loss)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if "--new-only" in sys.argv:      # the three original fixtures stay byte-identical
        stage_case("stages_reg_concat", 104, T=4, per=30, d=2, dx=5, dy=1, rows=70, classification=False,
                   embed_op="concat", concat=True)
        multi_case("multi_small", 105, T=5, per=24, d=3)
        blr_case("blr_manual", 106, T=4, per=33, d=3, P=7, weight_dim=[50, 50, 50])
        sys.exit(0)
    stage_case("stages_class_concat", 101, T=4, per=40, d=3, dx=16, dy=2, rows=120, classification=True,
               embed_op="joint", concat=True)
    stage_case("stages_reg_joint", 102, T=3, per=50, d=2, dx=7, dy=1, rows=90, classification=False,
               embed_op="joint", concat=False)
    stage_case("stages_reg_none", 103, T=5, per=31, d=1, dx=1, dy=1, rows=200, classification=False,
               embed_op="none", concat=False)
    stage_case("stages_reg_concat", 104, T=4, per=30, d=2, dx=5, dy=1, rows=70, classification=False,
               embed_op="concat", concat=True)
    multi_case("multi_small", 105, T=5, per=24, d=3)
    blr_case("blr_manual", 106, T=4, per=33, d=3, P=7, weight_dim=[50, 50, 50])
