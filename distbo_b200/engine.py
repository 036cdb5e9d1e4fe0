"""Device-side engine of the joint GP over (theta, dataset embedding).

Thin host layer over the C-ABI library: owns the HBM-resident state of one surrogate (padded
N x N matrices, vectors, hyper-parameters, Adam slots) as torch CUDA tensors and sequences the
kernels on torch's current stream.  PyTorch is used for device memory and streams only - every
arithmetic step on the hot path is a kernel of libdistbo_b200.so.

Layout in HBM (row-major, fp64 unless noted; n_pad = roundup(N, 128)):
  K/L   [n_pad, n_pad]  Gram -> Cholesky factor in place (lower 128-tiles valid; pad = identity)
  W     [n_pad, n_pad]  L^-1 (lower)
  T     [n_pad, n_pad]  TRTRI temp, reused as Kinv = W^T W (lower tiles)
  theta_s [n_pad, d], sqnorm [n_pad], V, alpha [n_pad], task_id int32 [N], task_off int32 [T+1]
  E     [T(+1), P]      dataset embeddings (last row = target at predict time), KE [T,T]
"""
from __future__ import annotations

import functools
import os
import math

import numpy as np
import torch

from . import _lib

JITTER = 0.000001  # reference train/log_gaus_likelihood.py:122,210
# slots of the device-side training control block (include/distbo_b200.h DISTBO_CTRL_*)
CTRL_CUR_MIN, CTRL_COUNTDOWN, CTRL_STOP, CTRL_EPOCHS, CTRL_LAST_LOSS, CTRL_FIRST_EPOCH, CTRL_INFO = range(7)
CTRL_THRESHOLD, CTRL_PATIENCE, CTRL_SLOTS = 7, 8, 16
ACQ = {"ei": 0, "ucb": 1, "lcb": 2}
D_MAX = 16         # gram.cu DMAX / score.cu SCORE_DMAX
I8_DIGITS = 7      # digit_planes.cuh I8_NS: int8 planes per fp64 operand of the tcgen05 contractions
I8_MAX_NPAD = 16384  # ... whose s32 accumulators are exact up to this size (larger problems stay on FP64 DMMA)


def round_up(n, m):
    return (n + m - 1) // m * m


def _stream():
    return _lib.C.c_void_p(torch.cuda.current_stream().cuda_stream)


def on_device(fn):
    """Runs an engine method with the engine's device current: kernel launches, stream lookups and the
    library's per-device state all follow the CURRENT device, not the tensors' device."""
    @functools.wraps(fn)
    def wrapped(self, *a, **kw):
        if torch.cuda.current_device() == self.dev.index:
            return fn(self, *a, **kw)
        with torch.cuda.device(self.dev):
            return fn(self, *a, **kw)
    return wrapped


class FitPlan:
    """The factorisation plan of ONE engine (tile lists, internal streams, events, inter-CTA flags and the relay
    buffer).  A plan carries mutable device state and is single-stream, so engines never share one; it lives on
    the device that was current when it was created and is destroyed with its engine."""

    def __init__(self, n_pad):
        self.lib = _lib.load()
        self.n_pad = n_pad
        self.device = torch.cuda.current_device()
        self.handle = _lib.C.c_void_p()
        _lib.check(self.lib.distbo_plan_create(n_pad, _lib.C.byref(self.handle)), "distbo_plan_create")

    def destroy(self):
        h, self.handle = self.handle, None
        if h is not None and self.lib is not None:
            try:
                with torch.cuda.device(self.device):
                    self.lib.distbo_plan_destroy(h)
            except Exception:       # interpreter shutdown: the driver reclaims everything
                pass

    def __del__(self):
        self.destroy()


class NetSpec:
    """Shapes of the embedding nets (reference train/gp_utils.py:13-26).

    The x net is dx -> hx -> p, or dx -> hx -> pm -> p when `pm` > 0 (the three-layer form nn_net takes when
    'weights_3' exists, train/distbo_net.py:11-13).  `xname` is the reference's parameter-name infix: 'x', or 'xy'
    for the 'nn' operator, whose single net reads the concatenated [x | y] rows (dx then counts both)."""

    H_MAX, P_MAX = 32, 16       # hidden / output widths the embedding kernels hold in registers (embed.cu)

    def __init__(self, dx, dy, hx, p, hy=0, q=0, y_mode=0, pm=0, xname="x"):
        self.dx, self.dy, self.hx, self.p, self.hy, self.q, self.y_mode = dx, dy, hx, p, hy, q, y_mode
        self.pm, self.xname = int(pm), xname
        if y_mode not in (0, 1, 2, 3, 4):
            raise ValueError("y_mode %r: 0 (x only), 1 (joint), 2 (classification), 3 (concat), 4 (x only + class "
                             "ratios)" % (y_mode,))
        if not (1 <= hx <= self.H_MAX and 1 <= p <= self.P_MAX and 0 <= self.pm <= self.P_MAX):
            raise ValueError("embedding net %s: hidden width <= %d, middle and output widths <= %d (got %d, %d, %d); "
                             "the shipped configs use (20, 10)" % (xname, self.H_MAX, self.P_MAX, hx, self.pm, p))
        if y_mode in (1, 3) and not (1 <= hy <= self.H_MAX and 1 <= q <= self.P_MAX):
            raise ValueError("embedding net y: hidden width <= %d and output width <= %d (got %d, %d)"
                             % (self.H_MAX, self.P_MAX, hy, q))

    @property
    def pooled_width(self):
        if self.y_mode == 0:
            return self.p
        if self.y_mode == 1:
            return self.p * self.q
        if self.y_mode == 3:                      # 'concat': mean(phi_x) + conditional embedding operator
            return self.p + self.q * self.p
        if self.y_mode == 4:                      # 'nn' on classification data: mean(phi_xy) + class ratios
            return self.p + self.dy
        return self.p * self.dy + self.dy

    @property
    def has_ynet(self):
        return self.y_mode in (1, 3)

    @property
    def n_sums(self):
        return self.p + self.q * self.p + self.p * self.p if self.y_mode == 3 else self.pooled_width


class ParamStore:
    """Flat fp64 parameter vector on the device with named views, plus gradient and Adam slots."""

    def __init__(self, shapes, device):
        self.shapes = dict(shapes)
        self.offsets = {}
        off = 0
        for k, shp in shapes.items():
            n = int(np.prod(shp)) if len(shp) else 1
            self.offsets[k] = (off, n)
            off += n
        self.n = off
        self.flat = torch.zeros(off, dtype=torch.float64, device=device)
        self.grad = torch.zeros(off, dtype=torch.float64, device=device)
        self.m = torch.zeros(off, dtype=torch.float64, device=device)
        self.v = torch.zeros(off, dtype=torch.float64, device=device)
        self.step = 0

    def view(self, name, buf=None):
        off, n = self.offsets[name]
        buf = self.flat if buf is None else buf
        return buf[off:off + n].view(*self.shapes[name]) if len(self.shapes[name]) else buf[off:off + 1]

    def gview(self, name):
        return self.view(name, self.grad)

    def set(self, name, value):
        v = torch.as_tensor(np.asarray(value, dtype=np.float64)).reshape(-1)
        off, n = self.offsets[name]
        assert v.numel() == n, (name, v.numel(), n)
        self.flat[off:off + n].copy_(v)

    def get(self, name):
        off, n = self.offsets[name]
        return self.flat[off:off + n].cpu().numpy().reshape(self.shapes[name])

    def get_grad(self, name):
        off, n = self.offsets[name]
        return self.grad[off:off + n].cpu().numpy().reshape(self.shapes[name])

    def has(self, name):
        return name in self.offsets


def net_param_shapes(net):
    """Trainable tensors of the embedding nets (train/gp_utils.py:13-26; log_gaus_likelihood.py:33-48)."""
    shapes = {}
    if net is not None:
        x = net.xname
        shapes["weights_%s_1" % x] = (net.dx, net.hx)
        shapes["bias_%s_1" % x] = (net.hx,)
        shapes["weights_%s_2" % x] = (net.hx, net.pm if net.pm else net.p)
        if net.pm:
            shapes["bias_%s_2" % x] = (net.pm,)
            shapes["weights_%s_3" % x] = (net.pm, net.p)
        if net.has_ynet:
            shapes["weights_y_1"] = (net.dy, net.hy)
            shapes["bias_y_1"] = (net.hy,)
            shapes["weights_y_2"] = (net.hy, net.q)
        if net.y_mode == 3:
            shapes["log_reg"] = ()                # log of the ridge lambda of the 'concat' operator, init 0
    return shapes


class GPEngine:
    def __init__(self, d_theta, n_embed=0, net=None, device=None, n_multi=0):
        if not torch.cuda.is_available():
            raise _lib.DistboError("distbo_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if not 1 <= int(d_theta) <= D_MAX:
            raise ValueError("d_theta = %d: the Gram / scoring kernels hold a hyper-parameter row in registers, "
                             "1 <= d_theta <= %d (the reference's search spaces have 1..7 dimensions)"
                             % (int(d_theta), D_MAX))
        self.lib = _lib.load()
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.dev.index is None:
            self.dev = torch.device("cuda", torch.cuda.current_device())
        self._plan = None
        self.d = int(d_theta)
        self.P = int(n_embed)            # embedding width incl. size-ratio column (0: kernel 'params')
        self.net = net
        shapes = net_param_shapes(net)
        if self.P > 0:
            shapes["log_bw_embed"] = (self.P,)
        self.n_multi = int(n_multi)      # kernel 'multi': number of tasks of the free-form task covariance
        if self.n_multi > 0:
            shapes["log_triangular_task"] = (self.n_multi * (self.n_multi + 1) // 2,)
            self._tri_ws = torch.empty(3 * self.n_multi * self.n_multi, dtype=torch.float64, device=self.dev)
        shapes["log_bw"] = (self.d,)
        shapes["log_scale"] = (1,)
        shapes["mean"] = (1,)
        shapes["log_sigma"] = ()
        self.params = ParamStore(shapes, self.dev)
        self.N = 0
        self.n_pad = 0
        self.info = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.nll = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self.best_val = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self.best_idx = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self._score_ws = None
        self._score_i8_ws = None
        self._w8 = None                  # int8 digit planes of W + row exponents (tcgen05 scoring path)
        self.W_version = 0               # bumped whenever self.W changes
        self._w8_version = -1
        # scoring path: tcgen05 int8 digit planes from n_pad = 512 (measured cross-over: 300 000 candidates take
        # 0.99 / 1.87 ms at N = 120, 6.5 / 5.1 ms at N = 600, 308 / 134 ms per 151 552 at N = 8192, DMMA / tcgen05);
        # DISTBO_SCORE_I8=0 keeps the FP64 DMMA kernel everywhere (A/B measurements)
        self.score_i8_min_npad = int(os.environ.get("DISTBO_SCORE_I8_MIN", "512"))
        self.use_score_i8 = os.environ.get("DISTBO_SCORE_I8", "1") == "1"
        # K^-1 = W^T W (LAUUM) on the same tcgen05 digit-plane kernel from n_pad = 2048; DISTBO_LAUUM_I8=0 = DMMA
        self.use_lauum_i8 = os.environ.get("DISTBO_LAUUM_I8", "1") == "1"
        self.lauum_i8_min_npad = int(os.environ.get("DISTBO_LAUUM_I8_MIN", "2048"))
        self._lauum_i8 = None
        # the top levels of W = L^-1 (blocks of at least DISTBO_TRTRI_I8_MIN rows) on the same kernel
        self.use_trtri_i8 = os.environ.get("DISTBO_TRTRI_I8", "1") == "1"
        self.trtri_i8_min_rows = int(os.environ.get("DISTBO_TRTRI_I8_MIN", "1024"))
        self._trtri_i8 = None
        self._emb_ws = {}
        self._emb_raw = None
        self.E = None
        self.KE = None
        self.T = 1

    def close(self):
        """Frees the factorisation plan (streams, events, flag buffers); the tensors go with the object."""
        if self._plan is not None:
            self._plan.destroy()
            self._plan = None
            self.plan = None

    # ------------------------------------------------------------------ buffers
    def _f64(self, *shape):
        return torch.zeros(*shape, dtype=torch.float64, device=self.dev)

    @on_device
    def set_observations(self, theta, y, sizes=None):
        """theta [N,d] (already StandardScaler-transformed), y [N]; sizes = observations per task."""
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1))
        N = theta.shape[0]
        assert theta.shape[1] == self.d and y.shape[0] == N
        n_pad = round_up(N, 128)
        if n_pad != self.n_pad:
            self.n_pad = n_pad
            self.K = self._f64(n_pad, n_pad)
            self.W = self._f64(n_pad, n_pad)
            self.Tm = self._f64(n_pad, n_pad)
            self.theta_s = self._f64(n_pad, self.d)
            self.sqnorm = self._f64(n_pad)
            self.V = self._f64(n_pad)
            self.alpha = self._f64(n_pad)
            self.kvec = self._f64(n_pad)
            self.nll_ws = self._f64(int(self.lib.distbo_nll_workspace(n_pad)))
            if self._plan is not None:
                self._plan.destroy()
            self._plan = FitPlan(n_pad)
            self.plan = self._plan.handle
            self._trtri_i8 = None
            if self.use_trtri_i8 and 2 * self.trtri_i8_min_rows <= n_pad <= I8_MAX_NPAD:
                mr = self.trtri_i8_min_rows
                first = int(self.lib.distbo_trtri_i8_first_level(n_pad, mr))
                if 1 <= first < int(self.lib.distbo_trtri_levels(self.plan)):
                    sched = torch.zeros(int(self.lib.distbo_trtri_i8_sched_words(n_pad, mr)), dtype=torch.int32,
                                        device=self.dev)
                    _lib.check(self.lib.distbo_trtri_i8_plan(n_pad, mr, _lib.ptr(sched)), "distbo_trtri_i8_plan")
                    self._trtri_i8 = (first, sched, self._f64(int(self.lib.distbo_trtri_i8_workspace(n_pad))))
            self._lauum_i8 = None
            if self.use_lauum_i8 and self.lauum_i8_min_npad <= n_pad <= I8_MAX_NPAD:
                sched = torch.zeros(int(self.lib.distbo_lauum_i8_sched_words(n_pad)), dtype=torch.int32,
                                    device=self.dev)
                _lib.check(self.lib.distbo_lauum_i8_plan(n_pad, _lib.ptr(sched)), "distbo_lauum_i8_plan")
                self._lauum_i8 = (sched, self._f64(int(self.lib.distbo_lauum_i8_workspace(n_pad))))
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
        self.gws = self._f64(n_pad * (self.T + self.d + 3) + self.T * self.T)
        self.H = self._f64(self.T, self.T)
        self.gscal = self._f64(3)
        self.W_valid = False

    # ------------------------------------------------------------------ embedding
    def _embed_workspace(self, S, Tn, elem_bytes, backward):
        net = self.net
        need = int(self.lib.distbo_embed_workspace(S, Tn, net.dx, net.dy, net.hx, net.p, net.pm, net.hy, net.q,
                                                   net.y_mode, elem_bytes, 1 if backward else 0))
        key = "bwd" if backward else "fwd"
        ws = self._emb_ws.get(key)
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, dtype=torch.uint8, device=self.dev)
            self._emb_ws[key] = ws
        return ws

    @on_device
    def embed(self, X, Y, seg_off, out=None, n_rows=None, dtype=torch.float64):
        """K1 forward: X [S,dx], Y [S,dy] device tensors, seg_off int32 [T+1] -> E rows [T, P]."""
        net = self.net
        Tn = seg_off.numel() - 1
        S = int(X.shape[0])
        X, Y = X.contiguous(), Y.contiguous()     # the kernels read row-major [S, d] (no-op for the regressor's buffers)
        if out is None:
            out = torch.zeros(Tn, max(self.P, net.pooled_width), dtype=dtype, device=self.dev)
        p = self.params
        if dtype == torch.float64:
            fn = self.lib.distbo_embed_fwd_f64
            cast = lambda t: t
        else:
            fn = self.lib.distbo_embed_fwd_f32
            cast = lambda t: t.float().contiguous()
        w = lambda name: cast(p.view(name)) if p.has(name) else None
        x = net.xname
        wx1, bx1, wx2 = w("weights_%s_1" % x), w("bias_%s_1" % x), w("weights_%s_2" % x)
        bx2, wx3 = w("bias_%s_2" % x), w("weights_%s_3" % x)
        wy1, by1, wy2 = w("weights_y_1"), w("bias_y_1"), w("weights_y_2")
        ws = self._embed_workspace(S, Tn, 8 if dtype == torch.float64 else 4, False)
        log_reg = raw = None
        if net.y_mode == 3:
            log_reg = cast(p.view("log_reg"))
            raw = self._emb_ws.get(("raw", Tn))
            if raw is None:
                raw = self._emb_ws[("raw", Tn)] = torch.empty(Tn, net.n_sums, dtype=torch.float64, device=self.dev)
            self._emb_raw = raw                   # kept for embed_backward on the same batch
        rc = fn(_lib.ptr(X), _lib.ptr(Y), _lib.ptr(seg_off), Tn, S, net.dx, net.dy,
                _lib.ptr(wx1), _lib.ptr(bx1), _lib.ptr(wx2), net.hx, net.p, net.pm, _lib.ptr(bx2), _lib.ptr(wx3),
                _lib.ptr(wy1), _lib.ptr(by1), _lib.ptr(wy2), net.hy, net.q,
                net.y_mode, _lib.ptr(log_reg), _lib.ptr(raw), _lib.ptr(out), out.stride(0), _lib.ptr(ws), ws.numel(),
                _stream())
        _lib.check(rc, "distbo_embed_fwd")
        return out

    @on_device
    def embed_backward(self, X, Y, seg_off, dE):
        """K1 backward into the gradient slots of the net weights."""
        net, p = self.net, self.params
        Tn = seg_off.numel() - 1
        S = int(X.shape[0])
        X, Y = X.contiguous(), Y.contiguous()
        ws = self._embed_workspace(S, Tn, 8, True)
        g = lambda name: _lib.ptr(p.gview(name)) if p.has(name) else None
        w = lambda name: _lib.ptr(p.view(name)) if p.has(name) else None
        x = net.xname
        rc = self.lib.distbo_embed_bwd_f64(
            _lib.ptr(X), _lib.ptr(Y), _lib.ptr(seg_off), Tn, S, net.dx, net.dy,
            w("weights_%s_1" % x), w("bias_%s_1" % x), w("weights_%s_2" % x), net.hx, net.p, net.pm,
            w("bias_%s_2" % x), w("weights_%s_3" % x),
            w("weights_y_1"), w("bias_y_1"), w("weights_y_2"), net.hy, net.q, net.y_mode,
            w("log_reg"), _lib.ptr(self._emb_raw) if net.y_mode == 3 else None,
            _lib.ptr(dE), dE.stride(0),
            g("weights_%s_1" % x), g("bias_%s_1" % x), g("weights_%s_2" % x), g("bias_%s_2" % x), g("weights_%s_3" % x),
            g("weights_y_1"), g("bias_y_1"), g("weights_y_2"), g("log_reg"),
            _lib.ptr(ws), ws.numel(), _stream())
        _lib.check(rc, "distbo_embed_bwd_f64")

    # ------------------------------------------------------------------ Gram and fit
    @on_device
    def task_gram(self, E, Tn):
        """KE [Tn,Tn] over the first Tn rows of E (width P)."""
        KE = self._f64(Tn, Tn)
        rc = self.lib.distbo_task_gram_f64(_lib.ptr(E), E.stride(0), Tn, self.P,
                                           _lib.ptr(self.params.view("log_bw_embed")), _lib.ptr(KE), _stream())
        _lib.check(rc, "distbo_task_gram_f64")
        return KE

    @on_device
    def task_tri(self):
        """Multi-task covariance KE [T,T] from log_triangular_task (train/kernels.py:32-40)."""
        T = self.n_multi
        KE = self._f64(T, T)
        rc = self.lib.distbo_task_tri_f64(_lib.ptr(self.params.view("log_triangular_task")), T, _lib.ptr(KE),
                                          _lib.ptr(self._tri_ws), _stream())
        _lib.check(rc, "distbo_task_tri_f64")
        return KE

    @on_device
    def task_tri_backward(self):
        rc = self.lib.distbo_task_tri_bwd_f64(_lib.ptr(self.params.view("log_triangular_task")), self.n_multi,
                                              _lib.ptr(self.H), _lib.ptr(self.params.gview("log_triangular_task")),
                                              _lib.ptr(self._tri_ws), _stream())
        _lib.check(rc, "distbo_task_tri_bwd_f64")

    @on_device
    def build_gram(self, KE=None):
        p = self.params
        rc = self.lib.distbo_scale_theta_f64(_lib.ptr(self.theta), self.N, self.d, _lib.ptr(p.view("log_bw")),
                                             _lib.ptr(self.theta_s), _lib.ptr(self.sqnorm), self.n_pad, _stream())
        _lib.check(rc, "distbo_scale_theta_f64")
        rc = self.lib.distbo_gram_f64(_lib.ptr(self.theta_s), _lib.ptr(self.sqnorm), self.N, self.n_pad, self.d,
                                      _lib.ptr(self.task_id), _lib.ptr(KE), self.T if KE is not None else 1,
                                      _lib.ptr(p.view("log_scale")), _lib.ptr(p.view("log_sigma")), JITTER,
                                      _lib.ptr(self.K), _stream())
        _lib.check(rc, "distbo_gram_f64")
        self.KE = KE

    @on_device
    def factorize(self, need_inverse=True, need_kinv=False):
        """K (in self.K) -> L in place; W = L^-1; optionally Kinv = W^T W (into self.Tm)."""
        st = _stream()
        _lib.check(self.lib.distbo_potrf_f64(self.plan, _lib.ptr(self.K), _lib.ptr(self.W), _lib.ptr(self.info), st),
                   "distbo_potrf_f64")
        if (need_inverse or need_kinv) and self._trtri_i8 is not None:
            first, sched, ws = self._trtri_i8
            _lib.check(self.lib.distbo_trtri_range_f64(self.plan, _lib.ptr(self.K), _lib.ptr(self.W), _lib.ptr(self.Tm),
                                                       0, first, st), "distbo_trtri_range_f64")
            _lib.check(self.lib.distbo_trtri_i8_f64(_lib.ptr(self.K), _lib.ptr(self.W), _lib.ptr(self.Tm), self.n_pad,
                                                    self.trtri_i8_min_rows, _lib.ptr(sched), _lib.ptr(ws), st),
                       "distbo_trtri_i8_f64")
        elif need_inverse or need_kinv:
            _lib.check(self.lib.distbo_trtri_f64(self.plan, _lib.ptr(self.K), _lib.ptr(self.W), _lib.ptr(self.Tm), st),
                       "distbo_trtri_f64")
        if need_inverse or need_kinv:
            self.W_valid = True
            self.W_version += 1
        if need_kinv and self._lauum_i8 is not None:
            _lib.check(self.lib.distbo_lauum_i8_f64(_lib.ptr(self.W), self.n_pad, _lib.ptr(self._lauum_i8[0]),
                                                    _lib.ptr(self._lauum_i8[1]), _lib.ptr(self.Tm), st),
                       "distbo_lauum_i8_f64")
        elif need_kinv:
            _lib.check(self.lib.distbo_lauum_f64(self.plan, _lib.ptr(self.W), _lib.ptr(self.Tm), st),
                       "distbo_lauum_f64")

    @on_device
    def likelihood(self):
        rc = self.lib.distbo_nll_f64(_lib.ptr(self.K), _lib.ptr(self.W), self.N, self.n_pad, _lib.ptr(self.y),
                                     _lib.ptr(self.params.view("mean")), _lib.ptr(self.V), _lib.ptr(self.alpha),
                                     _lib.ptr(self.nll), _lib.ptr(self.nll_ws), _stream())
        _lib.check(rc, "distbo_nll_f64")

    @on_device
    def gram_gradient(self):
        """Fills the gradient slots of log_bw, log_scale, mean, log_sigma and self.H = dNLL/dKE."""
        p = self.params
        multi = self.KE is not None
        rc = self.lib.distbo_gram_grad_f64(
            _lib.ptr(self.Tm), _lib.ptr(self.alpha), _lib.ptr(self.theta_s), _lib.ptr(self.sqnorm), self.N,
            self.n_pad, self.d, _lib.ptr(self.task_id) if multi else None,
            _lib.ptr(self.task_off) if multi else None, _lib.ptr(self.KE) if multi else None,
            self.T if multi else 1, _lib.ptr(p.view("log_scale")), _lib.ptr(p.view("log_sigma")),
            _lib.ptr(p.gview("log_bw")), _lib.ptr(self.gscal), _lib.ptr(self.H), _lib.ptr(self.gws), _stream())
        _lib.check(rc, "distbo_gram_grad_f64")
        p.gview("log_scale").copy_(self.gscal[0:1])
        p.gview("mean").copy_(self.gscal[1:2])
        p.gview("log_sigma").copy_(self.gscal[2:3])

    @on_device
    def task_gram_backward(self, E, Tn):
        dE = self._f64(Tn, E.shape[1])
        ws = torch.empty(Tn * self.P, dtype=torch.float64, device=self.dev)
        rc = self.lib.distbo_task_gram_bwd_f64(_lib.ptr(E), E.stride(0), Tn, self.P,
                                               _lib.ptr(self.params.view("log_bw_embed")), _lib.ptr(self.H),
                                               _lib.ptr(self.params.gview("log_bw_embed")), _lib.ptr(dE),
                                               _lib.ptr(ws), _stream())
        _lib.check(rc, "distbo_task_gram_bwd_f64")
        return dE

    @staticmethod
    def adam_lr_t(lr, step, beta1=0.9, beta2=0.999):
        """TF AdamOptimizer's per-step rate lr * sqrt(1 - beta2^t) / (1 - beta1^t)."""
        return lr * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)

    @on_device
    def adam_step(self, lr, clip_norm=0.0, lr_t_dev=None, ctrl=None, loss_hist=None):
        """TF-form Adam update of the flat parameter vector.  With lr_t_dev (device scalar) the launch
        takes its rate from device memory - the form a CUDA graph replays; the caller advances
        params.step and refreshes the scalar.  With ctrl (TrainControl.dev) the same launch also records
        the epoch's loss and evaluates the early-stopping rule on the device (include/distbo_b200.h)."""
        p = self.params
        if lr_t_dev is None:
            p.step += 1
        rc = self.lib.distbo_adam_f64(_lib.ptr(p.flat), _lib.ptr(p.grad), _lib.ptr(p.m), _lib.ptr(p.v), p.n,
                                      float(lr), 0.9, 0.999, 1e-8, max(p.step, 1), float(clip_norm),
                                      _lib.ptr(lr_t_dev), _lib.ptr(ctrl), _lib.ptr(self.nll) if ctrl is not None else None,
                                      _lib.ptr(self.info) if ctrl is not None else None, _lib.ptr(loss_hist),
                                      int(loss_hist.numel()) if loss_hist is not None else 0, _stream())
        _lib.check(rc, "distbo_adam_f64")

    def check_info(self, info):
        """Raises for a non-zero factorisation status, like TF's InvalidArgumentError."""
        info = int(info)
        if info < 0:
            raise _lib.DistboError("Cholesky decomposition was not successful: an inter-CTA wait timed out "
                                   "(info %d; the GPU is shared with other work and a producer CTA was not "
                                   "resident) - rerun, or set DISTBO_POTRF_SMALL=0" % info)
        if info != 0:
            raise _lib.DistboError("Cholesky decomposition was not successful: non-positive pivot at row %d "
                                   "(matrix not positive definite)" % info)

    @on_device
    def read_loss(self):
        """One device->host sync: (nll, info).  Raises on a failed Cholesky like TF's
        InvalidArgumentError."""
        out = torch.stack([self.nll[0], self.info[0].double()]).cpu().numpy()
        self.check_info(out[1])
        return float(out[0])

    # ------------------------------------------------------------------ scoring
    @on_device
    def score(self, theta_c, acq="ei", y_max=0.0, xi=0.0, kappa=1.0, want_arrays=True, index_offset=0,
              kvec=None, scaler=None):
        """K4 on candidates theta_c [M,d] (device or host array).  `scaler` = (mean_, scale_) device
        tensors [d] of the frozen StandardScaler, applied inside the kernel; None = theta_c is already
        transformed.  Returns dict(best_val, best_idx, mean, std, acq) - arrays only when want_arrays."""
        assert self.W_valid, "factorize(need_inverse=True) first"
        if not torch.is_tensor(theta_c):
            theta_c = torch.from_numpy(np.ascontiguousarray(theta_c, dtype=np.float64)).to(self.dev)
        M = int(theta_c.shape[0])
        if self.use_score_i8 and self.score_i8_min_npad <= self.n_pad <= I8_MAX_NPAD:
            return self._score_i8(theta_c, M, acq, y_max, xi, kappa, want_arrays, index_offset, kvec, scaler)
        chunk = self.lib.distbo_score_chunk()
        need = int(self.lib.distbo_score_workspace(self.n_pad))
        if self._score_ws is None or self._score_ws[0].numel() < need:
            self._score_ws = (torch.empty(need, dtype=torch.float64, device=self.dev),
                              torch.empty(chunk // 128, dtype=torch.float64, device=self.dev),
                              torch.empty(chunk // 128, dtype=torch.int64, device=self.dev))
        kst, bbv, bbi = self._score_ws
        if kvec is None:
            kvec = self.kvec
        mean = std = acqv = None
        if want_arrays:
            mean = torch.empty(M, dtype=torch.float64, device=self.dev)
            std = torch.empty(M, dtype=torch.float64, device=self.dev)
            acqv = torch.empty(M, dtype=torch.float64, device=self.dev)
        p = self.params
        rc = self.lib.distbo_score_f64(
            _lib.ptr(self.W), _lib.ptr(self.V), self.N, self.n_pad, _lib.ptr(self.theta_s), _lib.ptr(self.sqnorm),
            self.d, _lib.ptr(p.view("log_bw")), _lib.ptr(kvec), _lib.ptr(p.view("log_scale")),
            _lib.ptr(p.view("mean")), _lib.ptr(theta_c),
            _lib.ptr(scaler[0]) if scaler is not None else None,
            _lib.ptr(scaler[1]) if scaler is not None else None, M, ACQ[acq], float(y_max), float(xi), float(kappa),
            _lib.ptr(kst), _lib.ptr(mean), _lib.ptr(std), _lib.ptr(acqv), _lib.ptr(bbv), _lib.ptr(bbi),
            _lib.ptr(self.best_val), _lib.ptr(self.best_idx), int(index_offset), _stream())
        _lib.check(rc, "distbo_score_f64")
        return {"best_val": self.best_val, "best_idx": self.best_idx, "mean": mean, "std": std, "acq": acqv}

    def _score_i8(self, theta_c, M, acq, y_max, xi, kappa, want_arrays, index_offset, kvec, scaler):
        """The same operator on the tcgen05 tensor cores (int8 digit planes, exact integer accumulation in
        tensor memory; include/distbo_b200.h).  W's digit planes are refreshed when W changed."""
        n_pad = self.n_pad
        if self._w8 is None or self._w8[0].shape[1] != n_pad:
            self._w8 = (torch.empty(I8_DIGITS, n_pad, n_pad, dtype=torch.int8, device=self.dev), self._f64(n_pad))
            self._w8_version = -1
        if self._w8_version != self.W_version:
            _lib.check(self.lib.distbo_score_i8_prepare(_lib.ptr(self.W), n_pad, _lib.ptr(self._w8[0]),
                                                        _lib.ptr(self._w8[1]), _stream()), "distbo_score_i8_prepare")
            self._w8_version = self.W_version
        chunk = self.lib.distbo_score_chunk()
        need = int(self.lib.distbo_score_i8_workspace(n_pad))
        if self._score_i8_ws is None or self._score_i8_ws[0].numel() < need:
            self._score_i8_ws = (torch.empty(need, dtype=torch.float64, device=self.dev),
                                 torch.empty(chunk // 128, dtype=torch.float64, device=self.dev),
                                 torch.empty(chunk // 128, dtype=torch.int64, device=self.dev))
        ws, bbv, bbi = self._score_i8_ws
        if kvec is None:
            kvec = self.kvec
        mean = std = acqv = None
        if want_arrays:
            mean = torch.empty(M, dtype=torch.float64, device=self.dev)
            std = torch.empty(M, dtype=torch.float64, device=self.dev)
            acqv = torch.empty(M, dtype=torch.float64, device=self.dev)
        p = self.params
        rc = self.lib.distbo_score_i8_f64(
            _lib.ptr(self._w8[0]), _lib.ptr(self._w8[1]), _lib.ptr(self.V), self.N, n_pad, _lib.ptr(self.theta_s),
            _lib.ptr(self.sqnorm), self.d, _lib.ptr(p.view("log_bw")), _lib.ptr(kvec), _lib.ptr(p.view("log_scale")),
            _lib.ptr(p.view("mean")), _lib.ptr(theta_c),
            _lib.ptr(scaler[0]) if scaler is not None else None,
            _lib.ptr(scaler[1]) if scaler is not None else None, M, ACQ[acq], float(y_max), float(xi), float(kappa),
            _lib.ptr(ws), _lib.ptr(mean), _lib.ptr(std), _lib.ptr(acqv), _lib.ptr(bbv), _lib.ptr(bbi),
            _lib.ptr(self.best_val), _lib.ptr(self.best_idx), int(index_offset), _stream())
        _lib.check(rc, "distbo_score_i8_f64")
        return {"best_val": self.best_val, "best_idx": self.best_idx, "mean": mean, "std": std, "acq": acqv}

    # ------------------------------------------------------------------ one evaluation of loss + gradient
    @on_device
    def loss_and_grad(self, X=None, Y=None, seg_off=None, E_const=None, size_ratio=None):
        """One evaluation of the negative log marginal likelihood and its full gradient (what one
        sess.run([optimize_step, loss]) does in the reference, train/train.py:43-52), as a chain of
        asynchronous launches on the current stream.  No host synchronisation.

        kernel 'dist'  : X [S,dx], Y [S,dy], seg_off [T+1] = this step's embedding mini-batch;
                         size_ratio [T] fills the last embedding column (concat_data_size).
        kernel 'manual': E_const [T,P] meta-features.      kernel 'params': all None.
        """
        T = self.T
        E = None
        if self.net is not None:
            if self.E is None or self.E.shape[0] != T or self.E.shape[1] != self.P:
                self.E = self._f64(T, self.P)
            E = self.E
            if size_ratio is not None:
                E[:, self.P - 1].copy_(size_ratio)
            self.embed(X, Y, seg_off, out=E)
        elif E_const is not None:
            E = E_const
        if self.n_multi:
            assert T == self.n_multi, "kernel 'multi': the number of tasks is fixed at the first fit"
            KE = self.task_tri()
        else:
            KE = self.task_gram(E, T) if E is not None else None
        self.build_gram(KE)
        self.factorize(need_inverse=True, need_kinv=True)
        self.likelihood()
        self.gram_gradient()
        if self.n_multi:
            self.task_tri_backward()
        if E is not None:
            dE = self.task_gram_backward(E, T)
            if self.net is not None:
                self.embed_backward(X, Y, seg_off, dE)
