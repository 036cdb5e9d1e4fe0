"""CPU restatement of the distBO joint-GP hot path (TEST INFRASTRUCTURE ONLY).

PARITY PIN (see oracle/__init__.py): the reference ships no golden vectors and TensorFlow 1.7 cannot
run here; this file restates its TensorFlow-1.7 graph op-for-op and is checked against numbers the
reference's own sources produce over a TF op shim (tests/golden/ref_*.npz, tests/test_reference_golden.py).  All
``file:line`` citations are relative to the reference repository (hcllaw/distBO).

The same forward code runs on two array backends:
  * ``NP``  - numpy + scipy LAPACK (the oracle proper; stands in for TF's Eigen LLT,
              triangular solve and matmul);
  * ``TH``  - torch CPU float64 with autograd, which reproduces TensorFlow's
              reverse-mode gradient of the same graph (train/train.py:33-40).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.linalg
import scipy.stats

SQRT3 = math.sqrt(3.0)
JITTER = 0.000001          # train/log_gaus_likelihood.py:122,210
DIST_CLAMP = 1e-32         # train/kernels.py:25
VAR_CLAMP = 1e-32          # train/gp_regressor.py:197


# --------------------------------------------------------------------------- backends
class _NP:
    name = "numpy"

    @staticmethod
    def asarray(x, dtype=np.float64):
        return np.asarray(x, dtype=dtype)

    matmul = staticmethod(np.matmul)
    exp = staticmethod(np.exp)
    log = staticmethod(np.log)
    sqrt = staticmethod(np.sqrt)
    tanh = staticmethod(np.tanh)
    square = staticmethod(np.square)

    @staticmethod
    def clamp_min(x, c):
        return np.maximum(x, c)

    @staticmethod
    def sum(x, axis=None, keepdims=False):
        return np.sum(x, axis=axis, keepdims=keepdims)

    @staticmethod
    def cat(xs, axis):
        return np.concatenate(xs, axis=axis)

    @staticmethod
    def stack(xs, axis=0):
        return np.stack(xs, axis=axis)

    @staticmethod
    def eye(n, like):
        return np.eye(n, dtype=like.dtype)

    @staticmethod
    def cholesky(a):
        return scipy.linalg.cholesky(a, lower=True)

    @staticmethod
    def solve_lower(l, b):
        return scipy.linalg.solve_triangular(l, b, lower=True)

    @staticmethod
    def inv(a):
        return np.linalg.inv(a)

    @staticmethod
    def diag(a):
        return np.diagonal(a)

    @staticmethod
    def repeat_rows(emb, sizes):
        return np.repeat(emb, np.asarray(sizes, dtype=np.int64), axis=0)

    @staticmethod
    def transpose(a):
        return a.T

    @staticmethod
    def full_like_col(n, value, like):
        return np.full((n, 1), value, dtype=like.dtype)


class _TH:
    name = "torch"

    def __init__(self):
        import torch
        self.t = torch

    def asarray(self, x, dtype=None):
        t = self.t
        if isinstance(x, t.Tensor):
            return x
        return t.as_tensor(np.asarray(x, dtype=np.float64))

    def matmul(self, a, b):
        return a @ b

    def exp(self, x):
        return self.t.exp(x)

    def log(self, x):
        return self.t.log(x)

    def sqrt(self, x):
        return self.t.sqrt(x)

    def tanh(self, x):
        return self.t.tanh(x)

    def square(self, x):
        return x * x

    def clamp_min(self, x, c):
        # tf.maximum routes the gradient to x where x >= c, else to the constant
        return self.t.clamp(x, min=c)

    def sum(self, x, axis=None, keepdims=False):
        if axis is None:
            return x.sum()
        return x.sum(dim=axis, keepdim=keepdims)

    def cat(self, xs, axis):
        return self.t.cat(list(xs), dim=axis)

    def stack(self, xs, axis=0):
        return self.t.stack(list(xs), dim=axis)

    def eye(self, n, like):
        return self.t.eye(n, dtype=like.dtype)

    def cholesky(self, a):
        return self.t.linalg.cholesky(a)

    def solve_lower(self, l, b):
        return self.t.linalg.solve_triangular(l, b, upper=False)

    def inv(self, a):
        return self.t.linalg.inv(a)

    def diag(self, a):
        return self.t.diagonal(a)

    def repeat_rows(self, emb, sizes):
        idx = self.t.as_tensor(np.repeat(np.arange(len(sizes)), np.asarray(sizes, dtype=np.int64)))
        return emb[idx]

    def transpose(self, a):
        return a.T

    def full_like_col(self, n, value, like):
        return self.t.full((n, 1), float(value), dtype=like.dtype)


NP = _NP()
_TH_SINGLETON = None


def TH():
    global _TH_SINGLETON
    if _TH_SINGLETON is None:
        _TH_SINGLETON = _TH()
    return _TH_SINGLETON


# --------------------------------------------------------------------------- weights
def glorot_truncated_normal(rs, shape):
    """Stand-in for tf.keras.initializers.glorot_normal (train/gp_utils.py:13-26).

    TF 1.7 draws a Philox truncated normal (|z| <= 2 sigma) with
    sigma = sqrt(2 / (fan_in + fan_out)) in float32; Philox is not reproducible outside
    TF, so both this oracle and the product use numpy's MT19937 with
    resampling.  float32 values, cast by the caller.
    """
    fan_in, fan_out = shape
    sigma = math.sqrt(2.0 / (fan_in + fan_out))
    z = rs.standard_normal(size=shape)
    bad = np.abs(z) > 2.0
    while bad.any():
        z[bad] = rs.standard_normal(size=int(bad.sum()))
        bad = np.abs(z) > 2.0
    return (z * sigma).astype(np.float32)


def nn_weight_setter(params, in_dim, weight_dim, name, seed=23, dtype=np.float64):
    """train/gp_utils.py:13-26 (layer shapes, zero biases, no bias on the last layer)."""
    if len(weight_dim) >= 2:
        rs = np.random.RandomState(seed)
        params["weights_%s_1" % name] = glorot_truncated_normal(rs, (in_dim, weight_dim[0])).astype(dtype)
        params["bias_%s_1" % name] = np.zeros(weight_dim[0], dtype=dtype)
        params["weights_%s_2" % name] = glorot_truncated_normal(rs, (weight_dim[0], weight_dim[1])).astype(dtype)
        if len(weight_dim) == 3:
            params["bias_%s_2" % name] = np.zeros(weight_dim[1], dtype=dtype)
            params["weights_%s_3" % name] = glorot_truncated_normal(rs, (weight_dim[1], weight_dim[2])).astype(dtype)
    return params


def num_bw_embed(embed_op, model_type, in_dim_y, weight_dim_x, weight_dim_y, weight_dim_xy,
                 concat_data_size):
    """train/log_gaus_likelihood.py:33-59 (width P of the dataset embedding)."""
    if embed_op in ("none", "joint", "concat"):
        if embed_op in ("joint", "concat") and model_type == "regression":
            if embed_op == "joint":
                num_bw = weight_dim_x[-1] * weight_dim_y[-1]
            else:
                num_bw = weight_dim_x[-1] * (weight_dim_y[-1] + 1)
        else:
            num_bw = weight_dim_x[-1] * in_dim_y
    elif embed_op == "nn":
        num_bw = weight_dim_xy[-1]
    else:
        raise ValueError(embed_op)
    if concat_data_size:
        num_bw += 1
    if model_type == "classification":
        num_bw += in_dim_y
    return num_bw


def build_params(in_dim_params, kernel, in_dim_x=None, in_dim_y=None, weight_dim_x=None,
                 weight_dim_y=None, weight_dim_xy=None, model_type="regression",
                 embed_op="joint", concat_data_size=False, in_dim_meta=None,
                 likhd_var_init=0.01, bw_params_init=1.0, seed=23, dtype=np.float64, n_tasks=None):
    """Trainable parameters and their initial values.

    train/log_gaus_likelihood.py:33-84; log_sigma_init = 0.5*log(likhd_var_init)
    (train/gp_regressor.py:83).
    """
    params = {}
    if kernel == "dist":
        if embed_op in ("none", "joint", "concat"):
            nn_weight_setter(params, in_dim_x, weight_dim_x, "x", seed=seed, dtype=dtype)
            if embed_op in ("joint", "concat") and model_type == "regression":
                nn_weight_setter(params, in_dim_y, weight_dim_y, "y", seed=seed, dtype=dtype)
                if embed_op == "concat":
                    params["log_reg"] = np.zeros((), dtype=dtype)
        elif embed_op == "nn":
            nn_weight_setter(params, in_dim_x + in_dim_y, weight_dim_xy, "xy", seed=seed, dtype=dtype)
        nbw = num_bw_embed(embed_op, model_type, in_dim_y, weight_dim_x, weight_dim_y,
                           weight_dim_xy, concat_data_size)
        params["log_bw_embed"] = np.zeros(nbw, dtype=dtype)
    elif kernel == "manual":
        params["log_bw_embed"] = np.zeros(in_dim_meta, dtype=dtype)
    elif kernel == "multi":
        # log_gaus_likelihood.py:68-72: np.random.seed(seed); np.random.normal(size=T(T+1)/2).  RandomState(seed)
        # draws the same legacy stream without reseeding the caller's global generator.
        params["log_triangular_task"] = np.random.RandomState(seed).normal(
            size=n_tasks * (n_tasks + 1) // 2).astype(dtype)
    elif kernel != "params":
        raise NotImplementedError("kernel %r" % kernel)
    params["log_bw"] = np.log(np.full(in_dim_params, bw_params_init, dtype=dtype))
    params["log_scale"] = np.zeros(1, dtype=dtype)
    params["mean"] = np.zeros(1, dtype=dtype)
    params["log_sigma"] = np.asarray(0.5 * math.log(likhd_var_init), dtype=dtype)
    return params


# --------------------------------------------------------------------------- embedding
def nn_net(B, params, x, name):
    """train/distbo_net.py:7-14."""
    hidden = B.tanh(B.matmul(x, params["weights_%s_1" % name]) + params["bias_%s_1" % name])
    out = B.matmul(hidden, params["weights_%s_2" % name])
    if ("weights_%s_3" % name) in params:
        out = B.tanh(out + params["bias_%s_2" % name])
        out = B.matmul(out, params["weights_%s_3" % name])
    return out


def _segment_mean(B, feat, embed_sizes):
    """base.py:9-11,24-43: sparse [T,S] matrix with 1/size entries times feat."""
    bounds = np.r_[0, np.cumsum(np.asarray(embed_sizes, dtype=np.int64))]
    rows = []
    for s, e in zip(bounds[:-1], bounds[1:]):
        # 1/size is applied per element before the sum, as the sparse matmul does
        rows.append(B.sum(feat[s:e] * (1.0 / float(e - s)), axis=0))
    return B.stack(rows, axis=0)


def distbo_net(B, x, y, embed_sizes, params, embed_op="joint", model_type=None):
    """train/distbo_net.py:16-90 -> pooled embeddings [T, P']."""
    embed_sizes = np.asarray(embed_sizes, dtype=np.int64).reshape(-1)
    if embed_op == "nn":
        feature_xy = nn_net(B, params, B.cat([x, y], 1), "xy")
    else:
        feature_xy = feature_x = nn_net(B, params, x, "x")
        if embed_op in ("joint", "concat") and model_type == "regression":
            feature_y = nn_net(B, params, y, "y")
        elif model_type == "classification":
            feature_y = y
    if embed_op == "joint" or (model_type == "classification" and embed_op != "nn"):
        # einsum('bil,blj->bij') then reshape -> index i*q + j   (distbo_net.py:33-42)
        p, q = feature_x.shape[1], feature_y.shape[1]
        feature_xy = (feature_x[:, :, None] * feature_y[:, None, :]).reshape(feature_x.shape[0], p * q)
    if embed_op != "concat" or model_type == "classification":
        pool = _segment_mean(B, feature_xy, embed_sizes)
        if model_type == "classification":
            pool = B.cat([pool, _segment_mean(B, y, embed_sizes)], 1)   # class ratios, :45-48
        return pool
    # conditional-embedding operator branch, distbo_net.py:49-89
    bounds = np.r_[0, np.cumsum(embed_sizes)]
    reg = B.exp(params["log_reg"])
    p = feature_x.shape[1]
    marg, cond = [], []
    for s, e in zip(bounds[:-1], bounds[1:]):
        fx, fy = feature_x[s:e], feature_y[s:e]
        yx = B.matmul(B.transpose(fy), fx)
        xx = B.matmul(B.transpose(fx), fx)
        inv_term = B.inv(xx + reg * B.eye(p, fx))
        c = (1.0 / reg) * (yx - B.matmul(yx, B.matmul(inv_term, xx)))
        cond.append(c.reshape(-1))
        marg.append(B.sum(fx, axis=0) / float(e - s))           # tf.reduce_mean
    return B.cat([B.stack(marg, 0), B.stack(cond, 0)], 1)


def dist_features(B, params, x, y, sizes, embed_sizes, data_sizes=None, max_size=None,
                  embed_op="joint", model_type=None, return_pooled=False):
    """train/feature_map.py:48-65: pooled [T,P'] -> repeated [N,P] (+ size-ratio column)."""
    pooled = distbo_net(B, x, y, embed_sizes, params, embed_op=embed_op, model_type=model_type)
    if max_size is not None:
        ratio = np.asarray(data_sizes, dtype=np.float64).reshape(-1, 1) / float(max_size)
        pooled = B.cat([pooled, B.asarray(ratio) if B is not NP else ratio.astype(pooled.dtype)], 1)
    sizes = np.asarray(sizes, dtype=np.int64).reshape(-1)
    if return_pooled:
        return pooled
    return B.repeat_rows(pooled, sizes)


# --------------------------------------------------------------------------- kernels
def matern32_kernel(B, X, Y, stddev=1.0):
    """train/kernels.py:19-28 (expanded-distance form, clamp 1e-32)."""
    X = X / stddev
    Y = Y / stddev
    xs = B.sum(B.square(X), axis=1, keepdims=True)
    ys = B.transpose(B.sum(B.square(Y), axis=1, keepdims=True))
    xy = B.matmul(X, B.transpose(Y))
    dist = B.sqrt(B.clamp_min(-2 * xy + xs + ys, DIST_CLAMP))
    value = SQRT3 * dist
    return (1.0 + value) * B.exp(-value)


def fill_triangular_index(n):
    """Index map of tf.contrib.distributions.fill_triangular (TF 1.7, lower): the n x n row-major matrix is
    concat(x[n:], reverse(x)) and then band-limited to the lower triangle.  Returns (idx [n,n], mask [n,n])
    with L = x[idx] * mask.  TF's docstring example: fill_triangular([1..6]) = [[4,0,0],[6,5,0],[3,2,1]]."""
    m = n * (n + 1) // 2
    src = np.concatenate([np.arange(n, m), np.arange(m)[::-1]])
    return src.reshape(n, n), np.tril(np.ones((n, n)))


def task_kernel_matrix(B, params, n_tasks):
    """train/kernels.py:32-40 with log_gaus_likelihood.py:75: L = fill_triangular(exp(log_triangular_task)),
    L L^T rescaled to unit diagonal (rows first, then columns).  [T,T]."""
    idx, mask = fill_triangular_index(n_tasks)
    x = B.exp(params["log_triangular_task"])
    if B is NP:
        L = x[idx] * mask
    else:
        L = x[B.t.as_tensor(idx)] * B.t.as_tensor(mask)
    k = B.matmul(L, B.transpose(L))
    ds = B.sqrt(B.diag(k))
    k = k / ds.reshape(-1, 1)
    k = k / ds.reshape(1, -1)
    return k


def _task_rows(sizes):
    sizes = np.asarray(sizes, dtype=np.int64).reshape(-1)
    return np.repeat(np.arange(len(sizes)), sizes)


def task_kernel(B, params, sizes):
    """train/kernels.py:41-55 (sizes_2 None): the [T,T] matrix expanded to [N,N] by the task of each row."""
    t = _task_rows(sizes)
    k = task_kernel_matrix(B, params, len(np.asarray(sizes).reshape(-1)))
    if B is not NP:
        t = B.t.as_tensor(t)
    return k[t][:, t]


def task_kernel_pred(B, params, sizes, n_pred):
    """train/kernels.py:47-49 (sizes_2 given): every candidate belongs to the LAST task -> [N, M]."""
    t = _task_rows(sizes)
    k = task_kernel_matrix(B, params, len(np.asarray(sizes).reshape(-1)))
    if B is not NP:
        t = B.t.as_tensor(t)
    col = k[t][:, -1].reshape(-1, 1)
    return np.repeat(col, n_pred, axis=1) if B is NP else col.expand(-1, n_pred)


@dataclass
class TrainBatch:
    """What one sess.run of the likelihood graph is fed (train/base.py:91-127)."""
    h_params: object                 # [N, d_theta] scaled hyper-parameters
    obs: object                      # [N, 1]
    sizes: object = None             # [T] observations per task
    batch_X: object = None           # [S, d_x] stacked embedding samples
    batch_y: object = None           # [S, d_y]
    embed_sizes: object = None       # [T] rows per dataset in batch_X
    data_sizes: object = None        # [T] len(train_x) per dataset
    max_size: object = None
    embed: object = None             # [T, P] manual meta-features
    extra: dict = field(default_factory=dict)


def kernel_train(B, params, kernel, batch, embed_op="joint", model_type=None):
    """K without noise: scale * k_params * k_dist * k_man (log_gaus_likelihood.py:86-119)."""
    scale = B.exp(params["log_scale"])
    k = matern32_kernel(B, batch.h_params, batch.h_params, stddev=B.exp(params["log_bw"]))
    if kernel == "dist":
        emb = dist_features(B, params, batch.batch_X, batch.batch_y, batch.sizes, batch.embed_sizes,
                            data_sizes=batch.data_sizes, max_size=batch.max_size,
                            embed_op=embed_op, model_type=model_type)
        k = k * matern32_kernel(B, emb, emb, stddev=B.exp(params["log_bw_embed"]))
    elif kernel == "manual":
        emb = B.repeat_rows(batch.embed, np.asarray(batch.sizes).reshape(-1))     # feature_map.py:29-35
        k = k * matern32_kernel(B, emb, emb, stddev=B.exp(params["log_bw_embed"]))
    elif kernel == "multi":
        k = k * task_kernel(B, params, batch.sizes)
    return scale * k


def multivariate_normal_logpdf(B, x, mu, L):
    """train/gp_utils.py:118-147."""
    d = x - mu
    alpha = B.solve_lower(L, d)
    n = d.shape[0]
    p = -0.5 * B.sum(B.square(alpha), axis=0)
    p = p - 0.5 * n * math.log(2 * math.pi)
    p = p - B.sum(B.log(B.diag(L)))
    return p


def neg_log_marg_likhd(B, params, kernel, batch, embed_op="joint", model_type=None,
                       return_all=False):
    """loss of train/log_gaus_likelihood.py:119-125."""
    k = kernel_train(B, params, kernel, batch, embed_op=embed_op, model_type=model_type)
    n = k.shape[0]
    sigma_sq = B.exp(2.0 * params["log_sigma"])
    k_sigma = k + (sigma_sq + JITTER) * B.eye(n, k)
    L = B.cholesky(k_sigma)
    mean = params["mean"] * B.full_like_col(n, 1.0, k)
    loss = -multivariate_normal_logpdf(B, batch.obs, mean, L)
    loss = loss.reshape(())
    if return_all:
        return loss, L, k, k_sigma
    return loss


# --------------------------------------------------------------------------- posterior
def predict_factor(B, params, kernel, batch, k_dist=None):
    """Candidate-independent half of the posterior graph: K, L = chol(K + (sigma^2 + 1e-6) I) and
    V = L^-1 (y - m)  (log_gaus_likelihood.py:178-184 / :199-214).  The reference re-evaluates this
    on every predict call; it is split out so the CPU baseline can time it separately."""
    scale = B.exp(params["log_scale"])
    stddev = B.exp(params["log_bw"])
    sigma_sq = B.exp(2.0 * params["log_sigma"])
    n = batch.h_params.shape[0]
    k = matern32_kernel(B, batch.h_params, batch.h_params, stddev=stddev)
    if kernel == "dist":
        k = scale * k * k_dist
    elif kernel == "manual":
        emb = B.repeat_rows(batch.embed, np.asarray(batch.sizes).reshape(-1))
        k = scale * k * matern32_kernel(B, emb, emb, stddev=B.exp(params["log_bw_embed"]))
    elif kernel == "multi":
        k = scale * k * task_kernel(B, params, batch.sizes)
    else:
        k = scale * k
    L = B.cholesky(k + (sigma_sq + JITTER) * B.eye(n, k))
    V = B.solve_lower(L, batch.obs - params["mean"])
    return L, V


def predict_apply(B, params, kernel, batch, L, V, h_params_pred, k_dist_pred=None, embed_pred=None):
    """Per-candidate half: K_*, A = L^-1 K_*, mean = A^T V + m, var = scale - colsum(A^2)
    (log_gaus_likelihood.py:178-187 / :203-217)."""
    scale = B.exp(params["log_scale"])
    stddev = B.exp(params["log_bw"])
    k_pred = matern32_kernel(B, batch.h_params, h_params_pred, stddev=stddev)
    if kernel == "dist":
        k_pred = scale * k_pred * k_dist_pred
    elif kernel == "manual":
        emb = B.repeat_rows(batch.embed, np.asarray(batch.sizes).reshape(-1))
        emb_p = B.repeat_rows(embed_pred, [h_params_pred.shape[0]])
        k_pred = scale * k_pred * matern32_kernel(B, emb, emb_p, stddev=B.exp(params["log_bw_embed"]))
    elif kernel == "multi":
        k_pred = scale * k_pred * task_kernel_pred(B, params, batch.sizes, h_params_pred.shape[0])
    else:
        k_pred = scale * k_pred
    A = B.solve_lower(L, k_pred)
    mean = B.matmul(B.transpose(A), V) + params["mean"].reshape(1, 1)
    var = scale - B.sum(B.square(A), axis=0)
    return mean, var.reshape(-1, 1)


def predict(B, params, kernel, batch, h_params_pred, k_dist=None, k_dist_pred=None,
            embed_pred=None):
    """Posterior mean/variance.

    kernel == 'dist': build_predict_dist (log_gaus_likelihood.py:190-218) with the
    pre-computed embedding Grams fed in.  Otherwise build_predict (:128-188).
    Returns (mean[M,1], var[M,1]); var is NOT clamped here (gp_regressor.py:197 does it).
    """
    L, V = predict_factor(B, params, kernel, batch, k_dist=k_dist)
    return predict_apply(B, params, kernel, batch, L, V, h_params_pred, k_dist_pred=k_dist_pred,
                         embed_pred=embed_pred)


def embed_grams(B, params, batch, X_pred, y_pred, embed_sizes_pred, data_sizes_pred, n_pred,
                embed_op="joint", model_type=None):
    """k_dist [N,N] and k_dist_pred [N,M] as embed_network returns them
    (train/train.py:78-84; graph: kernels.py:69-88 via log_gaus_likelihood.py:146-161).
    Every column of k_dist_pred is identical (sizes_pred = [M], gp_regressor.py:171).
    """
    e1 = dist_features(B, params, batch.batch_X, batch.batch_y, batch.sizes, batch.embed_sizes,
                       data_sizes=batch.data_sizes, max_size=batch.max_size,
                       embed_op=embed_op, model_type=model_type)
    e2 = dist_features(B, params, X_pred, y_pred, [n_pred], embed_sizes_pred,
                       data_sizes=data_sizes_pred, max_size=batch.max_size,
                       embed_op=embed_op, model_type=model_type)
    bw = B.exp(params["log_bw_embed"])
    return matern32_kernel(B, e1, e1, stddev=bw), matern32_kernel(B, e1, e2, stddev=bw)


def similarity(B, params, kernel, batch, X_pred=None, y_pred=None, embed_sizes_pred=None,
               data_sizes_pred=None, embed_pred=None, embed_op="joint", model_type=None):
    """train/train.py:86-95 with sizes all 1 (base.py:137-139): [T,1] kernel between every
    source embedding and the target embedding."""
    T = len(np.asarray(batch.sizes).reshape(-1))
    ones = [1] * T
    if kernel == "multi":                       # net.k_task_pred with sizes all 1: column of the last task
        return task_kernel_pred(B, params, ones, 1)
    bw = B.exp(params["log_bw_embed"])
    if kernel == "dist":
        e1 = dist_features(B, params, batch.batch_X, batch.batch_y, ones, batch.embed_sizes,
                           data_sizes=batch.data_sizes, max_size=batch.max_size,
                           embed_op=embed_op, model_type=model_type)
        e2 = dist_features(B, params, X_pred, y_pred, [1], embed_sizes_pred,
                           data_sizes=data_sizes_pred, max_size=batch.max_size,
                           embed_op=embed_op, model_type=model_type)
    elif kernel == "manual":
        e1 = B.repeat_rows(batch.embed, ones)
        e2 = B.repeat_rows(embed_pred, [1])
    else:
        raise ValueError(kernel)
    return matern32_kernel(B, e1, e2, stddev=bw)


# --------------------------------------------------------------------------- acquisition
def std_from_var(var):
    """train/gp_regressor.py:197."""
    return np.sqrt(np.maximum(var, VAR_CLAMP))


def acq_ei(mean, std, y_max, xi):
    """bayes_opt/helpers.py:165-171."""
    mean = np.squeeze(mean)
    std = np.squeeze(std)
    z = (mean - y_max - xi) / std
    return (mean - y_max - xi) * scipy.stats.norm.cdf(z) + std * scipy.stats.norm.pdf(z)


def acq_ucb(mean, std, kappa):
    """bayes_opt/helpers.py:151-156."""
    return np.squeeze(mean) + kappa * np.squeeze(std)


def acq_lcb(mean, std, kappa):
    """bayes_opt/helpers.py:158-163."""
    return np.squeeze(mean) - kappa * np.squeeze(std)


def acquisition(kind, mean, std, y_max=0.0, xi=0.0, kappa=1.0):
    if kind == "ei":
        return acq_ei(mean, std, y_max, xi)
    if kind == "ucb":
        return acq_ucb(mean, std, kappa)
    if kind == "lcb":
        return acq_lcb(mean, std, kappa)
    raise NotImplementedError(kind)


# --------------------------------------------------------------------------- gradients
def loss_and_grads(params, kernel, batch, embed_op="joint", model_type=None):
    """Loss and d loss / d params through torch autograd (mirrors TF autodiff,
    train/train.py:31-40).  Returns (float, {name: ndarray})."""
    import torch
    B = TH()
    tp = {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=True) for k, v in params.items()}
    tb = TrainBatch(
        h_params=torch.as_tensor(np.asarray(batch.h_params, dtype=np.float64)),
        obs=torch.as_tensor(np.asarray(batch.obs, dtype=np.float64)),
        sizes=batch.sizes,
        batch_X=None if batch.batch_X is None else torch.as_tensor(np.asarray(batch.batch_X, dtype=np.float64)),
        batch_y=None if batch.batch_y is None else torch.as_tensor(np.asarray(batch.batch_y, dtype=np.float64)),
        embed_sizes=batch.embed_sizes, data_sizes=batch.data_sizes, max_size=batch.max_size,
        embed=None if batch.embed is None else torch.as_tensor(np.asarray(batch.embed, dtype=np.float64)),
    )
    loss = neg_log_marg_likhd(B, tp, kernel, tb, embed_op=embed_op, model_type=model_type)
    loss.backward()
    grads = {k: (np.zeros_like(np.asarray(params[k], dtype=np.float64)) if v.grad is None
                 else v.grad.numpy().copy()) for k, v in tp.items()}
    return float(loss.detach()), grads


class TFAdam:
    """tf.train.AdamOptimizer defaults (beta1=.9, beta2=.999, eps=1e-8), 'epsilon-hat'
    update lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps).
    State persists across maximise_likhd calls (train/gp_regressor.py:141-143)."""

    def __init__(self, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.t = 0
        self.m, self.v = {}, {}

    def step(self, params, grads, clip_norm=None):
        if clip_norm is not None:           # tf.clip_by_global_norm, train/train.py:33-35
            gn = math.sqrt(sum(float(np.sum(np.square(g))) for g in grads.values()))
            if gn > clip_norm:
                grads = {k: g * (clip_norm / gn) for k, g in grads.items()}
        self.t += 1
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        for k, g in grads.items():
            if k not in self.m:
                self.m[k] = np.zeros_like(g)
                self.v[k] = np.zeros_like(g)
            self.m[k] = self.b1 * self.m[k] + (1 - self.b1) * g
            self.v[k] = self.b2 * self.v[k] + (1 - self.b2) * g * g
            params[k] = params[k] - lr_t * self.m[k] / (np.sqrt(self.v[k]) + self.eps)
        return params
