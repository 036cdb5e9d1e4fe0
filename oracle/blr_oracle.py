"""CPU restatement of the reference's Bayesian-linear-regression surrogate (algorithm='BLR').

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs as the checker.
The product path (distbo_b200/) never imports this module.

Follows, in order (file:line relative to the reference repository):
  train/log_gaus_likelihood_feature.py:13-143   parameters, feature assembly, marginal likelihood
  train/feature_map.py:12-46                    nn_features (deep feature map, optional random Fourier map),
                                                manual_features, task_features
  train/RandomFourierMaternFeatureMapper.py:12-66   Student-t frequencies, uniform phases (numpy, seed 23)
  train/log_gaus_likelihood_feature.py:145-225  posterior mean / variance
  train/gp_regressor.py:70-74,153-156,179-196   how gp_regressor drives it

Parity: the reference ships no tests and TF 1.7 cannot run here; tests/test_reference_golden.py pins this file
against numbers the reference's own log_gaus_likelihood_feature.py produces over a TF op shim (ref_blr_*.npz:
num_of_rff = 0 as shipped, and one case with the random Fourier map).  Further pins in tests/test_oracle.py:
the loss equals -2 log N(y | 0, sigma^2 I + alpha Phi Phi^T) - N log 2 pi (scipy), the posterior equals the
textbook weight-space posterior, and the random Fourier map converges to sklearn's Matern(nu=1.5) (the
author's own check, RandomFourierMaternFeatureMapper.py:69-82).
"""
import math

import numpy as np
from sklearn.utils import check_random_state

from . import gp_oracle as O


def build_params_blr(in_dim_params, kernel, weight_dim, num_of_rff=0, in_dim_x=None, in_dim_y=None,
                     weight_dim_x=None, weight_dim_y=None, model_type="regression", embed_op="joint",
                     concat_data_size=False, in_dim_meta=None, n_tasks=None, likhd_var_init=0.01, seed=23,
                     dtype=np.float64, weight_dim_xy=None):
    """log_gaus_likelihood_feature.py:30-79.  Returns (params, rff_input_dim)."""
    params = {}
    if kernel == "dist":
        if embed_op not in ("none", "joint", "concat", "nn"):
            raise NotImplementedError(embed_op)
        if embed_op == "nn":                                 # :49-53: one net over [x | y]
            O.nn_weight_setter(params, in_dim_x + in_dim_y, weight_dim_xy, "xy", seed=seed, dtype=dtype)
        else:
            O.nn_weight_setter(params, in_dim_x, weight_dim_x, "x", seed=seed, dtype=dtype)
            if embed_op in ("joint", "concat") and model_type == "regression":
                O.nn_weight_setter(params, in_dim_y, weight_dim_y, "y", seed=seed, dtype=dtype)
                if embed_op == "concat":
                    params["log_reg"] = np.zeros((), dtype=dtype)
        embed_dim = O.num_bw_embed(embed_op, model_type, in_dim_y, weight_dim_x, weight_dim_y, weight_dim_xy,
                                   concat_data_size)
    elif kernel == "manual":
        embed_dim = in_dim_meta
    elif kernel == "multi":
        embed_dim = n_tasks
    else:
        embed_dim = 0
    total_dim = in_dim_params + embed_dim
    weight_dim = list(weight_dim or [])
    O.nn_weight_setter(params, total_dim, weight_dim, "", seed=seed, dtype=dtype)
    rff_input_dim = weight_dim[-1] if len(weight_dim) != 0 else total_dim
    params["log_sigma"] = np.asarray(0.5 * math.log(likhd_var_init), dtype=dtype)   # gp_regressor.py:83
    params["log_bw"] = np.zeros(rff_input_dim, dtype=dtype)
    params["log_alpha"] = np.asarray(0.0, dtype=dtype)
    return params, rff_input_dim


def rff_constants(in_dim, n_rff, nu=1.5, seed=23):
    """RandomFourierMaternFeatureMapper.sample/map (:12-60): frequencies ~ multivariate Student-t with 2 nu
    degrees of freedom, phases ~ U(0, 2 pi), both drawn from RandomState(seed) in this order.  A fresh mapper
    (same seed) is built on every nn_features call, so training and prediction share the constants."""
    rs = check_random_state(seed)
    x = rs.chisquare(2.0 * nu, n_rff) / (2.0 * nu)
    z = rs.multivariate_normal(np.zeros(in_dim), np.eye(in_dim), (n_rff,))
    samples = np.transpose(z / np.sqrt(x)[:, None])                # [in_dim, n_rff]
    from scipy.stats import uniform
    bias = 2.0 * math.pi * uniform.rvs(size=n_rff, random_state=rs)
    return samples, bias


def nn_features(B, params, feature, rff_input_dim, num_of_rff):
    """train/feature_map.py:12-27."""
    if "weights__1" in params and "weights__2" in params:
        feature = B.tanh(B.matmul(feature, params["weights__1"]) + params["bias__1"])
        feature = B.matmul(feature, params["weights__2"])
        if "weights__3" in params:
            feature = B.matmul(B.tanh(feature + params["bias__2"]), params["weights__3"])
    if num_of_rff > 0:
        # the reference casts to float32 around the random Fourier map (:21,26) and back afterwards
        samples, bias = rff_constants(rff_input_dim, num_of_rff)
        feature = feature / B.exp(params["log_bw"])
        if B is O.NP:
            f32 = feature.astype(np.float32)
            phi = np.cos(np.matmul(f32, samples.astype(np.float32)) + bias.astype(np.float32))
            feature = (np.float32(math.sqrt(2.0 / num_of_rff)) * phi).astype(np.float64)
        else:
            t = B.t
            f32 = feature.to(t.float32)
            phi = t.cos(f32 @ t.as_tensor(samples.astype(np.float32)) + t.as_tensor(bias.astype(np.float32)))
            feature = (math.sqrt(2.0 / num_of_rff) * phi).to(t.float64)
    return feature


def task_one_hot(B, n_tasks, sizes, like, pred=False):
    """train/feature_map.py:37-46: rows of the identity repeated per task; candidates take the LAST task."""
    t = np.full(int(np.sum(sizes)), n_tasks - 1) if pred else O._task_rows(sizes)
    eye = np.eye(n_tasks)[t]
    return eye if B is O.NP else B.t.as_tensor(eye)


def train_feature(B, params, kernel, batch, embed_op="joint", model_type=None, n_tasks=None):
    """[embedding part | h_params]  (log_gaus_likelihood_feature.py:86-105)."""
    if kernel == "dist":
        fea = O.dist_features(B, params, batch.batch_X, batch.batch_y, batch.sizes, batch.embed_sizes,
                              data_sizes=batch.data_sizes, max_size=batch.max_size, embed_op=embed_op,
                              model_type=model_type)
    elif kernel == "manual":
        fea = B.repeat_rows(batch.embed, np.asarray(batch.sizes).reshape(-1))
    elif kernel == "multi":
        fea = task_one_hot(B, n_tasks, np.asarray(batch.sizes).reshape(-1), batch.h_params)
    else:
        return batch.h_params
    return B.cat([fea, batch.h_params], 1)


def blr_factor(B, params, nn_fea, obs):
    """L_inv and e_term of log_gaus_likelihood_feature.py:121-130."""
    sigma_sq_inv = 1.0 / B.exp(2.0 * params["log_sigma"])
    alpha = B.exp(params["log_alpha"])
    phi_t_phi = B.matmul(B.transpose(nn_fea), nn_fea)
    inside = B.eye(nn_fea.shape[1], nn_fea) + alpha * sigma_sq_inv * phi_t_phi
    L = B.cholesky(inside)
    L_inv = B.inv(L)
    e_term = B.matmul(L_inv, B.matmul(B.transpose(nn_fea), obs))
    return L, L_inv, e_term


def blr_loss(B, params, kernel, batch, rff_input_dim, num_of_rff, embed_op="joint", model_type=None,
             n_tasks=None):
    """net.loss of log_gaus_likelihood_feature.py:107-143."""
    feature = train_feature(B, params, kernel, batch, embed_op=embed_op, model_type=model_type, n_tasks=n_tasks)
    nn_fea = nn_features(B, params, feature, rff_input_dim, num_of_rff)
    sigma_sq = B.exp(2.0 * params["log_sigma"])
    alpha = B.exp(params["log_alpha"])
    y = batch.obs
    L, L_inv, e_term = blr_factor(B, params, nn_fea, y)
    term_1 = B.sum(B.square(y))
    term_2 = alpha / sigma_sq * B.sum(B.square(e_term))
    term_3 = B.log(sigma_sq) * float(y.shape[0])
    term_4 = 2.0 * B.sum(B.log(B.diag(L)))
    return ((1.0 / sigma_sq) * (term_1 - term_2) + term_3 + term_4).reshape(())


def blr_predict(B, params, nn_fea, obs, nn_fea_pred):
    """pred_mean / pred_var of log_gaus_likelihood_feature.py:184-190,216-224 (variance not clamped here)."""
    sigma_sq = B.exp(2.0 * params["log_sigma"])
    alpha = B.exp(params["log_alpha"])
    _, L_inv, e_term = blr_factor(B, params, nn_fea, obs)
    L_inv_fea_pred_t = B.matmul(L_inv, B.transpose(nn_fea_pred))
    mean = (alpha / sigma_sq) * B.transpose(B.matmul(B.transpose(e_term), L_inv_fea_pred_t))
    var = alpha * B.sum(B.square(L_inv_fea_pred_t), axis=0).reshape(-1, 1)
    return mean, var


def blr_loss_and_grads(params, kernel, batch, rff_input_dim, num_of_rff, embed_op="joint", model_type=None,
                       n_tasks=None):
    """Loss and gradient through torch autograd (TF autodiff, train/train.py:31-40).  Parameters the loss does
    not depend on (log_bw when num_of_rff == 0) get a zero gradient: TF returns None for them and Adam
    skips the variable, which leaves it at its initial value either way."""
    import torch
    B = O.TH()
    tp = {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=True) for k, v in params.items()}
    as_t = lambda a: None if a is None else torch.as_tensor(np.asarray(a, dtype=np.float64))
    tb = O.TrainBatch(h_params=as_t(batch.h_params), obs=as_t(batch.obs), sizes=batch.sizes,
                      batch_X=as_t(batch.batch_X), batch_y=as_t(batch.batch_y), embed_sizes=batch.embed_sizes,
                      data_sizes=batch.data_sizes, max_size=batch.max_size, embed=as_t(batch.embed))
    loss = blr_loss(B, tp, kernel, tb, rff_input_dim, num_of_rff, embed_op=embed_op, model_type=model_type,
                    n_tasks=n_tasks)
    loss.backward()
    grads = {k: (np.zeros_like(np.asarray(params[k], dtype=np.float64)) if v.grad is None
                 else v.grad.numpy().copy()) for k, v in tp.items()}
    return float(loss.detach()), grads
