"""Device-side engine of the BLR surrogate (algorithm='BLR').

Bayesian linear regression on deep features of [dataset embedding | theta]
(reference train/log_gaus_likelihood_feature.py:13-225, train/feature_map.py:12-46).  Everything is linear
in the number of observations N: no N x N matrix exists; the only factorisation is D x D (D <= 64).
State in HBM: theta [N,d], y [N], task_id / task_off, the flat parameter / gradient / Adam vectors, the
row activations inside one workspace (6 x [N,64] doubles), Linv [64,64], e [64].

Shares the embedding kernels (K1) and the parameter store with GPEngine; the BLR kernels are
distbo_blr_fit_f64 / distbo_blr_score_f64 of libdistbo_b200.so.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .engine import ACQ, D_MAX, GPEngine, ParamStore, _stream, net_param_shapes, on_device

BLR_W = 64          # widest layer the kernels take (blr.cu)


class BLREngine(GPEngine):
    def __init__(self, d_theta, weight_dim, n_embed=0, net=None, device=None):
        if not torch.cuda.is_available():
            raise _lib.DistboError("distbo_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        weight_dim = [int(w) for w in (weight_dim or [])]
        if len(weight_dim) not in (2, 3) or max(weight_dim) > BLR_W:
            raise NotImplementedError("BLR feature map: weight_dim of 2 or 3 layers, widths <= %d (got %r)"
                                      % (BLR_W, weight_dim))
        if not 1 <= int(d_theta) <= D_MAX:
            raise ValueError("d_theta = %d: 1 <= d_theta <= %d" % (int(d_theta), D_MAX))
        self.lib = _lib.load()
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.dev.index is None:
            self.dev = torch.device("cuda", torch.cuda.current_device())
        self._plan = None
        self.d = int(d_theta)
        self.P = int(n_embed)            # width of the embedding part of a feature row (0: kernel 'params')
        self.net = net
        self.weight_dim = weight_dim
        self.n_layers = len(weight_dim)
        self.h1, self.h2 = weight_dim[0], weight_dim[1]
        self.D = weight_dim[-1]
        shapes = net_param_shapes(net)
        shapes["weights__1"] = (self.P + self.d, self.h1)
        shapes["bias__1"] = (self.h1,)
        shapes["weights__2"] = (self.h1, self.h2)
        if self.n_layers == 3:
            shapes["bias__2"] = (self.h2,)
            shapes["weights__3"] = (self.h2, self.D)
        shapes["log_sigma"] = ()
        shapes["log_bw"] = (self.D,)     # only the random Fourier map reads it (feature_map.py:20): stays 0 here
        shapes["log_alpha"] = ()
        self.params = ParamStore(shapes, self.dev)
        self.N = 0
        self.T = 1
        self.info = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.nll = torch.zeros(1, dtype=torch.float64, device=self.dev)      # the BLR loss (read_loss reads it)
        self.best_val = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self.best_idx = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self.Linv = torch.zeros(BLR_W * BLR_W, dtype=torch.float64, device=self.dev)
        self.evec = torch.zeros(BLR_W, dtype=torch.float64, device=self.dev)
        nb = int(self.lib.distbo_blr_score_blocks())
        self._bbv = torch.empty(nb, dtype=torch.float64, device=self.dev)
        self._bbi = torch.empty(nb, dtype=torch.int64, device=self.dev)
        self._emb_ws = {}
        self._emb_raw = None
        self.E = None
        self._ws = None
        self.emb_target = None           # [P] embedding of the candidates' task

    @on_device
    def set_observations(self, theta, y, sizes=None):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1))
        N = theta.shape[0]
        assert theta.shape[1] == self.d and y.shape[0] == N
        self.N = N
        self.theta = torch.from_numpy(theta).to(self.dev)
        self.y = torch.from_numpy(y).to(self.dev)
        if sizes is None:
            sizes = [N]
        sizes = np.asarray(sizes, dtype=np.int64).reshape(-1)
        assert sizes.sum() == N
        self.T = len(sizes)
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        self.task_off = torch.from_numpy(off).to(self.dev)
        self.task_id = torch.from_numpy(np.repeat(np.arange(self.T), sizes).astype(np.int32)).to(self.dev)
        need = int(self.lib.distbo_blr_workspace(N, self.T, self.d, self.h1, self.h2, self.D))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.float64, device=self.dev)
        self.dE = self._f64(self.T, max(self.P, 1))

    def _net_args(self):
        p = self.params
        w = lambda n: _lib.ptr(p.view(n)) if p.has(n) else None
        return (w("weights__1"), w("bias__1"), w("weights__2"), w("bias__2"), w("weights__3"), self.n_layers,
                self.h1, self.h2, self.D)

    @on_device
    def fit(self, E, want_grad=True):
        """Loss (into self.nll), Linv, e and - with want_grad - the gradient of every BLR parameter and
        dE = d loss / d E  (one evaluation of net.loss and tf.gradients, log_gaus_likelihood_feature.py:107-143)."""
        p = self.params
        g = lambda n: _lib.ptr(p.gview(n)) if (want_grad and p.has(n)) else None
        rc = self.lib.distbo_blr_fit_f64(
            _lib.ptr(self.theta), _lib.ptr(self.y), self.N, self.d,
            _lib.ptr(E) if self.P else None, E.stride(0) if self.P else 0, self.P,
            _lib.ptr(self.task_id), _lib.ptr(self.task_off), self.T, *self._net_args(),
            _lib.ptr(p.view("log_sigma")), _lib.ptr(p.view("log_alpha")), _lib.ptr(self.Linv), _lib.ptr(self.evec),
            _lib.ptr(self.nll), _lib.ptr(self.info),
            g("weights__1"), g("bias__1"), g("weights__2"), g("bias__2"), g("weights__3"), g("log_sigma"),
            g("log_alpha"), _lib.ptr(self.dE) if (want_grad and self.P) else None, _lib.ptr(self._ws),
            1 if want_grad else 0, _stream())
        _lib.check(rc, "distbo_blr_fit_f64")

    @on_device
    def loss_and_grad(self, X=None, Y=None, seg_off=None, E_const=None, size_ratio=None):
        """One evaluation of the BLR loss and its full gradient, queued on the current stream.
        kernel 'dist': X, Y, seg_off = this step's embedding mini-batch (size_ratio fills the last column);
        'manual' / 'multi': E_const [T,P] (meta-features / identity);  'params': all None."""
        E = None
        if self.net is not None:
            if self.E is None or self.E.shape[0] != self.T or self.E.shape[1] != self.P:
                self.E = self._f64(self.T, self.P)
            E = self.E
            if size_ratio is not None:
                E[:, self.P - 1].copy_(size_ratio)
            self.embed(X, Y, seg_off, out=E)
        elif E_const is not None:
            E = E_const
        self.fit(E, want_grad=True)
        if self.net is not None:
            self.embed_backward(X, Y, seg_off, self.dE)

    @on_device
    def prepare_posterior(self, E, emb_target):
        """Linv and e at the current parameters for features built from E [T,P]; emb_target [P] is the embedding
        part of every candidate's feature row (build_predict_feature / build_predict_dist_feature, :145-225)."""
        self.fit(E, want_grad=False)
        self.emb_target = emb_target

    @on_device
    def score(self, theta_c, acq="ei", y_max=0.0, xi=0.0, kappa=1.0, want_arrays=True, index_offset=0,
              kvec=None, scaler=None):
        if not torch.is_tensor(theta_c):
            theta_c = torch.from_numpy(np.ascontiguousarray(theta_c, dtype=np.float64)).to(self.dev)
        M = int(theta_c.shape[0])
        mean = std = acqv = None
        if want_arrays:
            mean = torch.empty(M, dtype=torch.float64, device=self.dev)
            std = torch.empty(M, dtype=torch.float64, device=self.dev)
            acqv = torch.empty(M, dtype=torch.float64, device=self.dev)
        p = self.params
        rc = self.lib.distbo_blr_score_f64(
            _lib.ptr(theta_c), _lib.ptr(scaler[0]) if scaler is not None else None,
            _lib.ptr(scaler[1]) if scaler is not None else None, M, self.d,
            _lib.ptr(self.emb_target) if self.P else None, self.P, *self._net_args(),
            _lib.ptr(self.Linv), _lib.ptr(self.evec), _lib.ptr(p.view("log_sigma")), _lib.ptr(p.view("log_alpha")),
            ACQ[acq], float(y_max), float(xi), float(kappa), _lib.ptr(mean), _lib.ptr(std), _lib.ptr(acqv),
            _lib.ptr(self._bbv), _lib.ptr(self._bbi), _lib.ptr(self.best_val), _lib.ptr(self.best_idx),
            int(index_offset), _stream())
        _lib.check(rc, "distbo_blr_score_f64")
        return {"best_val": self.best_val, "best_idx": self.best_idx, "mean": mean, "std": std, "acq": acqv}
