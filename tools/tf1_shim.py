"""A minimal lazy-graph stand-in for the TensorFlow 1.x API surface that hcllaw/distBO's GP path uses, evaluated
with torch CPU tensors (fp64 / fp32 as the graph asks).

PURPOSE: golden-vector generation only (tools/make_reference_golden.py).  TensorFlow 1.7 cannot be installed in
the build container, so the reference's OWN Python sources (distBO/train/*.py, imported from /root/reference,
unmodified) are executed over this shim: every `tf.*` call builds a node, `Session.run(fetches, feed_dict)`
evaluates the nodes with the one-to-one torch operation, `tf.gradients` / `AdamOptimizer` use torch autograd and
TF's documented Adam update.  What is restated here is the semantics of single TF ops (matmul, cholesky,
triangular solve, tile, fill_triangular, ...), not the reference's algorithm: the composition - kernels, pooling,
likelihood, posterior, training loop, early stopping, batching - is the reference's code running as written.
Nothing in the product, the tests' GPU path or bench.py imports this file.
"""
from __future__ import annotations

import math
import sys
import types

import numpy as np
import torch


# ------------------------------------------------------------------ dtypes
class DType(object):
    def __init__(self, name, tdt):
        self.name, self.torch = name, tdt

    def __repr__(self):
        return "tf." + self.name


float32 = DType("float32", torch.float32)
float64 = DType("float64", torch.float64)
int32 = DType("int32", torch.int32)
int64 = DType("int64", torch.int64)
bool_ = DType("bool", torch.bool)
_BY_TORCH = {d.torch: d for d in (float32, float64, int32, int64, bool_)}


def _tdt(dtype):
    if dtype is None:
        return None
    if isinstance(dtype, DType):
        return dtype.torch
    if isinstance(dtype, torch.dtype):
        return dtype
    return {np.float32: torch.float32, np.float64: torch.float64, np.int32: torch.int32, np.int64: torch.int64}[dtype]


class _Shape(object):
    def __init__(self, ndims):
        self.ndims = ndims


# ------------------------------------------------------------------ graph nodes
class Tensor(object):
    """Lazy node: fn(env, *evaluated_inputs) -> torch tensor (or tuple for multi-output nodes)."""

    def __init__(self, fn, inputs=(), rank=None, dtype=None, name=None):
        self.fn, self.inputs, self.rank, self._dtype, self.name = fn, tuple(inputs), rank, dtype, name

    @property
    def shape(self):
        return _Shape(self.rank)

    def get_shape(self):
        return _Shape(self.rank)

    @property
    def dtype(self):
        return self._dtype

    # arithmetic
    def __add__(self, o): return _binary(torch.add, self, o)
    def __radd__(self, o): return _binary(torch.add, o, self)
    def __sub__(self, o): return _binary(torch.sub, self, o)
    def __rsub__(self, o): return _binary(torch.sub, o, self)
    def __mul__(self, o): return _binary(torch.mul, self, o)
    def __rmul__(self, o): return _binary(torch.mul, o, self)
    def __truediv__(self, o): return _binary(torch.true_divide, self, o)
    def __rtruediv__(self, o): return _binary(torch.true_divide, o, self)
    def __floordiv__(self, o): return _binary(torch.floor_divide, self, o)
    def __neg__(self): return Tensor(lambda env, a: -a, [self], self.rank, self._dtype)
    def __pow__(self, o): return _binary(torch.pow, self, o)
    def __lt__(self, o): return _binary(torch.lt, self, o)
    def __le__(self, o): return _binary(torch.le, self, o)
    def __gt__(self, o): return _binary(torch.gt, self, o)
    def __ge__(self, o): return _binary(torch.ge, self, o)
    __hash__ = object.__hash__

    def __getitem__(self, idx):
        def fn(env, a, ix):
            return a[_index(ix)]
        rank = None
        if self.rank is not None and not isinstance(idx, tuple):
            rank = self.rank if isinstance(idx, slice) else self.rank - 1
        return Tensor(fn, [self, idx], rank, self._dtype)


def _index(ix):
    if isinstance(ix, tuple):
        return tuple(_index(i) for i in ix)
    if isinstance(ix, slice):
        f = lambda v: None if v is None else int(v)
        return slice(f(ix.start), f(ix.stop), f(ix.step))
    if torch.is_tensor(ix) and ix.numel() == 1:
        return int(ix)
    return ix


def _rank_of(x):
    if isinstance(x, Tensor):
        return x.rank
    return np.ndim(x)


def _dtype_of(*xs):
    for x in xs:
        if isinstance(x, Tensor) and x._dtype is not None:
            return x._dtype
    return None


def _binary(op, a, b):
    ra, rb = _rank_of(a), _rank_of(b)
    rank = max(ra, rb) if ra is not None and rb is not None else None
    dt = _dtype_of(a, b)

    def fn(env, x, y):
        x, y = _promote(x, y)
        return op(x, y)
    out_dt = bool_ if op in (torch.lt, torch.le, torch.gt, torch.ge) else dt
    return Tensor(fn, [a, b], rank, out_dt)


def _promote(x, y):
    """Python / numpy operands take the dtype of the tensor operand (TF's constant conversion)."""
    tx, ty = torch.is_tensor(x), torch.is_tensor(y)
    if tx and not ty:
        y = torch.as_tensor(np.asarray(y), dtype=x.dtype) if x.dtype.is_floating_point or float(y) == int(y) \
            else torch.as_tensor(np.asarray(y, dtype=np.float64))
    elif ty and not tx:
        x = torch.as_tensor(np.asarray(x), dtype=y.dtype) if y.dtype.is_floating_point or float(x) == int(x) \
            else torch.as_tensor(np.asarray(x, dtype=np.float64))
    elif not tx and not ty:
        x, y = torch.as_tensor(x), torch.as_tensor(y)
    if x.dtype != y.dtype and x.dtype.is_floating_point != y.dtype.is_floating_point:
        # int tensor with float tensor (shape arithmetic): follow the float side
        if x.dtype.is_floating_point:
            y = y.to(x.dtype)
        else:
            x = x.to(y.dtype)
    return x, y


def _ev(x, env):
    """Evaluate a node / nested structure of nodes; plain values pass through."""
    if isinstance(x, Tensor):
        key = id(x)
        if key not in env:
            if x.fn is None:
                raise KeyError("placeholder / loop symbol %r has no value in this run" % (x.name,))
            env[key] = x.fn(env, *[_ev(i, env) for i in x.inputs])
        return env[key]
    if isinstance(x, (list, tuple)):
        return type(x)(_ev(i, env) for i in x)
    if isinstance(x, slice):
        return slice(_ev(x.start, env), _ev(x.stop, env), _ev(x.step, env))
    return x


def _as_t(x, dtype=None):
    if torch.is_tensor(x):
        return x if dtype is None else x.to(dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype)


def _const_node(value):
    """Wrap an evaluated torch value as a node (used to re-enter Python callbacks in while_loop / map_fn)."""
    rank = value.dim() if torch.is_tensor(value) else np.ndim(value)
    dt = _BY_TORCH.get(value.dtype) if torch.is_tensor(value) else None
    return Tensor(lambda env: value, [], rank, dt)


# ------------------------------------------------------------------ variables, placeholders, session
_VARIABLES = []
_EAGER_SEED = [0]
INIT_VALUES = {}      # id(variable) -> numpy array it was initialised with (read by tools/make_reference_golden.py)


class Placeholder(Tensor):
    def __init__(self, dtype, shape=None, name=None):
        Tensor.__init__(self, None, [], None if shape is None else len(shape), dtype if isinstance(dtype, DType) else
                        _BY_TORCH.get(_tdt(dtype)), name)


def placeholder(dtype, shape=None, name=None):
    return Placeholder(dtype, shape, name)


class Variable(Tensor):
    def __init__(self, initial_value, dtype=None, name=None, trainable=True):
        init = initial_value if isinstance(initial_value, Tensor) else constant(initial_value, dtype=dtype)
        dt = dtype if isinstance(dtype, DType) else (init._dtype if init._dtype is not None else None)
        Tensor.__init__(self, self._read, [], init.rank, dt, name)
        self.init, self.value, self.trainable = init, None, trainable
        self.adam_m = self.adam_v = None
        _VARIABLES.append(self)

    def _read(self, env):
        if self.value is None:
            raise RuntimeError("variable %r read before tf.global_variables_initializer()" % (self.name,))
        return self.value

    def load(self, array):
        self.value = torch.tensor(np.asarray(array), dtype=self.value.dtype, requires_grad=True)


class _Op(Tensor):
    """Node evaluated for its side effect (initializer, optimizer step); Session.run returns None for it."""


def global_variables_initializer():
    def fn(env):
        for v in _VARIABLES:
            if v.value is None:
                val = _ev(v.init, {})
                tdt = _tdt(v._dtype) or val.dtype
                v.value = val.detach().clone().to(tdt).requires_grad_(val.dtype.is_floating_point)
                INIT_VALUES[id(v)] = v.value.detach().numpy().copy()
        return None
    return _Op(fn, [])


def trainable_variables():
    return [v for v in _VARIABLES if v.trainable]


def reset_default_graph():
    del _VARIABLES[:]


class GraphKeys(object):
    UPDATE_OPS = "update_ops"


def get_collection(key):
    return []


class _NullContext(object):
    def __enter__(self): return self
    def __exit__(self, *a): return False


def control_dependencies(ops):
    return _NullContext()


class ConfigProto(object):
    def __init__(self, **kw):
        self.kw = kw


class Session(object):
    def __init__(self, config=None, **kw):
        self.closed = False
        self.loss_log = []           # the second fetch of every [optimize_step, loss] run (train/train.py:49-50)

    def run(self, fetches, feed_dict=None):
        self.last_feed = feed_dict
        env = {}
        for k, v in (feed_dict or {}).items():
            if v is None:
                continue
            env[id(k)] = _as_t(v, _tdt(k._dtype))
        single = not isinstance(fetches, (list, tuple))
        flist = [fetches] if single else list(fetches)
        vals = [None] * len(flist)
        for i, f in enumerate(flist):          # values first: a fetched loss is the pre-update loss
            if not isinstance(f, _Op):
                vals[i] = _ev(f, env)
        for i, f in enumerate(flist):
            if isinstance(f, _Op):
                _ev(f, env)
        out = [None if v is None else v.detach().cpu().numpy() for v in vals]
        if len(flist) == 2 and isinstance(flist[0], _Op) and out[1] is not None:
            self.loss_log.append(float(out[1]))
        return out[0] if single else out

    def close(self):
        self.closed = True

    def __enter__(self): return self
    def __exit__(self, *a): self.close()


# ------------------------------------------------------------------ op helpers
def _unary(op, same_rank=True):
    def build(x, name=None):
        return Tensor(lambda env, a: op(_as_t(a)), [x], _rank_of(x) if same_rank else None, _dtype_of(x))
    return build


exp = _unary(torch.exp)
log = _unary(torch.log)
sqrt = _unary(torch.sqrt)
square = _unary(torch.square)
tanh = _unary(torch.tanh)
cos = _unary(torch.cos)
sin = _unary(torch.sin)
ones_like = _unary(torch.ones_like)


def constant(value, dtype=None, shape=None, name=None):
    if isinstance(value, Tensor):
        return cast(value, dtype)
    arr = np.asarray(value)
    tdt = _tdt(dtype)
    if tdt is None:
        tdt = torch.float32 if arr.dtype.kind == "f" else (torch.int32 if arr.dtype.kind in "iu" else None)
    t = torch.as_tensor(arr, dtype=tdt)
    return Tensor(lambda env: t, [], t.dim(), _BY_TORCH.get(t.dtype))


def cast(x, dtype, name=None):
    tdt = _tdt(dtype)
    dt = dtype if isinstance(dtype, DType) else _BY_TORCH.get(tdt)
    return Tensor(lambda env, a: _as_t(a) if tdt is None else _as_t(a).to(tdt), [x], _rank_of(x), dt)


def shape(x, name=None):
    return Tensor(lambda env, a: torch.tensor(list(a.shape), dtype=torch.int32), [x], 1, int32)


def expand_dims(x, axis, name=None):
    r = _rank_of(x)
    return Tensor(lambda env, a: torch.unsqueeze(_as_t(a), int(axis)), [x], None if r is None else r + 1, _dtype_of(x))


def squeeze(x, axis=None, name=None):
    return Tensor(lambda env, a: torch.squeeze(_as_t(a)) if axis is None else torch.squeeze(_as_t(a), axis), [x], None,
                  _dtype_of(x))


def _ints(seq):
    return [int(v) for v in (seq.tolist() if torch.is_tensor(seq) else seq)]


def reshape(x, shp, name=None):
    def fn(env, a, s):
        if torch.is_tensor(s):
            s = s.tolist()
        return torch.reshape(a, _ints(list(s)))
    rank = len(shp) if isinstance(shp, (list, tuple)) else None
    return Tensor(fn, [x, shp], rank, _dtype_of(x))


def transpose(x, perm=None, name=None):
    def fn(env, a):
        return a.t() if perm is None and a.dim() == 2 else a.permute(*(perm if perm is not None else reversed(range(a.dim()))))
    return Tensor(fn, [x], _rank_of(x), _dtype_of(x))


def concat(values, axis, name=None):
    def fn(env, vs):
        return torch.cat([_as_t(v) for v in vs], int(axis))
    rank = _rank_of(values[0])
    return Tensor(fn, [list(values)], rank, _dtype_of(*values))


def tile(x, multiples, name=None):
    def fn(env, a, m):
        m = _ints(m if isinstance(m, (list, tuple)) else (m.reshape(-1) if torch.is_tensor(m) else [m]))
        return a.repeat(*m)
    return Tensor(fn, [x, multiples], _rank_of(x), _dtype_of(x))


def _reduce(op):
    def build(x, axis=None, keepdims=False, name=None, keep_dims=None):
        kd = keepdims if keep_dims is None else keep_dims
        r = _rank_of(x)

        def fn(env, a):
            return op(a) if axis is None else op(a, dim=axis, keepdim=kd)
        rank = None if r is None else (0 if axis is None else (r if kd else r - 1))
        return Tensor(fn, [x], rank, _dtype_of(x))
    return build


reduce_sum = _reduce(torch.sum)
reduce_mean = _reduce(torch.mean)


def divide(a, b, name=None): return _binary(torch.true_divide, a, b)
def multiply(a, b, name=None): return _binary(torch.mul, a, b)
def maximum(a, b, name=None): return _binary(torch.maximum, a, b)
def less(a, b, name=None): return _binary(torch.lt, a, b)
def add(a, b, name=None): return _binary(torch.add, a, b)
def subtract(a, b, name=None): return _binary(torch.sub, a, b)


def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    def fn(env, x, y):
        x = x.transpose(-1, -2) if transpose_a else x
        y = y.transpose(-1, -2) if transpose_b else y
        return torch.matmul(x, y)
    return Tensor(fn, [a, b], 2, _dtype_of(a, b))


def einsum(eq, *ops):
    return Tensor(lambda env, *xs: torch.einsum(eq, *xs), list(ops), None, _dtype_of(*ops))


def eye(n, dtype=float32, name=None):
    return Tensor(lambda env, k: torch.eye(int(k), dtype=_tdt(dtype)), [n], 2, dtype if isinstance(dtype, DType) else None)


def ones(shp, dtype=float32, name=None):
    def fn(env, s):
        s = _ints(s) if isinstance(s, (list, tuple)) or (torch.is_tensor(s) and s.dim() > 0) else [int(s)]
        return torch.ones(*s, dtype=_tdt(dtype))
    rank = len(shp) if isinstance(shp, (list, tuple)) else 1
    return Tensor(fn, [shp], rank, dtype if isinstance(dtype, DType) else None)


def range_(start, limit=None, delta=1, dtype=None, name=None):
    def fn(env, a, b):
        lo, hi = (0, int(a)) if b is None else (int(a), int(b))
        return torch.arange(lo, hi, delta, dtype=torch.int32)
    return Tensor(fn, [start, limit], 1, int32)


def cumsum(x, axis=0, name=None):
    return Tensor(lambda env, a: torch.cumsum(a, axis).to(a.dtype), [x], _rank_of(x), _dtype_of(x))


def cholesky(x, name=None):
    return Tensor(lambda env, a: torch.linalg.cholesky(a), [x], 2, _dtype_of(x))


def matrix_triangular_solve(matrix, rhs, lower=True, adjoint=False, name=None):
    def fn(env, m, r):
        m = m.transpose(-1, -2).conj() if adjoint else m
        return torch.linalg.solve_triangular(m, r, upper=(not lower) != bool(adjoint))
    return Tensor(fn, [matrix, rhs], 2, _dtype_of(matrix, rhs))


def matrix_diag_part(x, name=None):
    return Tensor(lambda env, a: torch.diagonal(a, dim1=-2, dim2=-1), [x], 1, _dtype_of(x))


diag_part = matrix_diag_part


def matrix_diag(x, name=None):
    return Tensor(lambda env, a: torch.diag_embed(a), [x], 2, _dtype_of(x))


def matrix_inverse(x, name=None):
    return Tensor(lambda env, a: torch.linalg.inv(a), [x], 2, _dtype_of(x))


def self_adjoint_eig(x, name=None):
    node = Tensor(lambda env, a: tuple(torch.linalg.eigh(a)), [x])
    return node[0], node[1]


def norm(x, ord=2, name=None):
    return Tensor(lambda env, a: torch.linalg.vector_norm(a.reshape(-1), ord=ord), [x], 0, _dtype_of(x))


def where(cond, x=None, y=None, name=None):
    return Tensor(lambda env, c, a, b: torch.where(c, *_promote(a, b)), [cond, x, y], _rank_of(x), _dtype_of(x, y))


def Print(x, data, message=None, summarize=None, name=None):
    return x


def check_numerics(x, message, name=None):
    return x


def identity(x, name=None):
    return x


class _Symbol(Tensor):
    """Loop variable / map element: bound in the evaluation environment of one iteration."""

    def __init__(self, like=None, rank=None, dtype=None):
        Tensor.__init__(self, None, [], _rank_of(like) if like is not None else rank,
                        _dtype_of(like) if like is not None else dtype)


def while_loop(cond, body, loop_vars, **kw):
    """cond / body are traced ONCE, here, with symbolic loop variables - as TensorFlow does at graph construction (the
    reference rebinds names its lambdas close over right after the call, e.g. feature_map.py:31-34).  Evaluation
    iterates: each pass binds the symbols in a copy of the outer environment."""
    loop_vars = tuple(loop_vars)
    syms = [_Symbol(like=lv) for lv in loop_vars]
    cond_node = cond(*syms)
    body_nodes = tuple(body(*syms))

    def fn(env, vals):
        vals = [_as_t(v) for v in vals]
        while True:
            it_env = dict(env)
            for sym, v in zip(syms, vals):
                it_env[id(sym)] = v
            if not bool(_ev(cond_node, it_env)):
                break
            vals = [_as_t(v) for v in _ev(body_nodes, it_env)]
        return tuple(vals)
    node = Tensor(fn, [list(loop_vars)])
    outs = []
    for i, lv in enumerate(loop_vars):
        outs.append(Tensor(lambda env, t, i=i: t[i], [node], _rank_of(lv), _dtype_of(lv)))
    return tuple(outs)


def map_fn(fn, elems, dtype=None, **kw):
    sym = _Symbol(rank=None if _rank_of(elems) is None else _rank_of(elems) - 1, dtype=_dtype_of(elems))
    out_node = fn(sym)

    def run(env, es):
        outs = []
        for e in es:
            it_env = dict(env)
            it_env[id(sym)] = e
            outs.append(_ev(out_node, it_env))
        return torch.stack(outs, 0)
    return Tensor(run, [elems], None, dtype if isinstance(dtype, DType) else None)


# ------------------------------------------------------------------ sparse (base.py:9-11)
class SparseTensor(object):
    def __init__(self, indices, values, dense_shape):
        self.indices, self.values, self.dense_shape = indices, values, dense_shape


SparseMatrix = SparseTensor


def sparse_tensor_dense_matmul(sp, dense, name=None):
    def fn(env, idx, val, shp, d):
        idx = _as_t(idx).long()
        rows, cols = _ints(shp)
        m = torch.zeros(rows, cols, dtype=d.dtype)
        m = m.index_put((idx[:, 0], idx[:, 1]), _as_t(val).to(d.dtype), accumulate=True)
        return m @ d
    return Tensor(fn, [sp.indices, sp.values, sp.dense_shape, dense], 2, _dtype_of(dense))


# ------------------------------------------------------------------ gradients, clipping, Adam
def _all_grads(loss, variables):
    """One node holding d loss / d v for every variable (None where the loss does not reach v)."""
    def fn(env, l):
        leaves = [v.value for v in variables]
        gs = torch.autograd.grad(l, leaves, retain_graph=True, allow_unused=True)
        return tuple(None if g is None else g.detach() for g in gs)
    return Tensor(fn, [loss])


def gradients(ys, xs, **kw):
    xs = list(xs) if isinstance(xs, (list, tuple)) else [xs]
    node = _all_grads(ys, xs)
    return [Tensor(lambda env, t, i=i: t[i], [node], x.rank, x._dtype) for i, x in enumerate(xs)]


def global_norm(t_list):
    def fn(env, ts):
        return torch.sqrt(sum(torch.sum(t * t) for t in ts if t is not None))
    return Tensor(fn, [list(t_list)], 0)


def clip_by_global_norm(t_list, clip_norm, use_norm=None, name=None):
    gn = global_norm(t_list)

    def scale(env, t, g):
        if t is None:
            return None
        c = torch.as_tensor(float(clip_norm), dtype=g.dtype)
        return t * (c / torch.maximum(g, c))
    return [Tensor(scale, [t, gn], _rank_of(t), _dtype_of(t)) for t in t_list], gn


class AdamOptimizer(object):
    """tf.train.AdamOptimizer: lr_t = lr sqrt(1 - b2^t) / (1 - b1^t); m, v moments; var -= lr_t m / (sqrt(v) + eps)."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, **kw):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon
        self.t = 0

    def compute_gradients(self, loss, var_list=None):
        vs = list(var_list) if var_list is not None else trainable_variables()
        return list(zip(gradients(loss, vs), vs))

    def apply_gradients(self, grads_and_vars, global_step=None, name=None):
        gv = list(grads_and_vars)

        def fn(env, gs):
            self.t += 1
            lr_t = self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
            with torch.no_grad():
                for g, (_, v) in zip(gs, gv):
                    if g is None:
                        continue
                    if v.adam_m is None:
                        v.adam_m, v.adam_v = torch.zeros_like(v.value), torch.zeros_like(v.value)
                    v.adam_m.mul_(self.b1).add_(g, alpha=1.0 - self.b1)
                    v.adam_v.mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
                    v.value.sub_(lr_t * v.adam_m / (torch.sqrt(v.adam_v) + self.eps))
            return None
        return _Op(fn, [[g for g, _ in gv]])

    def minimize(self, loss, var_list=None, **kw):
        return self.apply_gradients(self.compute_gradients(loss, var_list))


# ------------------------------------------------------------------ initializers (values are recorded by the caller)
class _GlorotNormal(object):
    """Stand-in for tf.keras.initializers.glorot_normal(seed): a truncated normal with stddev sqrt(2/(fan_in+fan_out)).
    TF's Philox stream cannot be reproduced; the golden files store the values this produced, and the parity
    tests inject them."""

    def __init__(self, seed=None):
        self.seed = 0 if seed is None else int(seed)

    def __call__(self, shape, dtype=None, partition_info=None):
        _EAGER_SEED[0] += 1
        rs = np.random.RandomState((self.seed * 1000 + _EAGER_SEED[0]) % (2 ** 32))
        std = math.sqrt(2.0 / (shape[0] + shape[1]))
        vals = rs.normal(size=tuple(shape))
        bad = np.abs(vals) > 2.0
        while bad.any():
            vals[bad] = rs.normal(size=int(bad.sum()))
            bad = np.abs(vals) > 2.0
        t = torch.as_tensor((vals * std).astype(np.float32))
        return Tensor(lambda env: t, [], 2, float32)


def fill_triangular(x, upper=False, name=None):
    """tf.contrib.distributions.fill_triangular (lower): n = (sqrt(8 m + 1) - 1) / 2;
    reshape(concat([x[n:], reverse(x)]), [n, n]) restricted to the lower triangle."""
    def fn(env, a):
        m = a.shape[-1]
        n = int(round(math.sqrt(0.25 + 2.0 * m) - 0.5))
        assert n * (n + 1) // 2 == m
        xc = torch.cat([a[n:], torch.flip(a, dims=[-1])], -1).reshape(n, n)
        return torch.tril(xc)
    return Tensor(fn, [x], 2, _dtype_of(x))


# ------------------------------------------------------------------ module assembly
def install():
    """Registers this shim as `tensorflow` (and the sub-modules the reference imports) in sys.modules."""
    tf = types.ModuleType("tensorflow")
    me = sys.modules[__name__]
    for k in dir(me):
        if not k.startswith("_") or k in ("_VARIABLES",):
            setattr(tf, k, getattr(me, k))
    tf.range = range_
    tf.bool = bool_
    tf.nn = types.SimpleNamespace(tanh=tanh)
    tf.linalg = types.SimpleNamespace(inv=matrix_inverse)
    tf.train = types.SimpleNamespace(AdamOptimizer=AdamOptimizer)
    tf.keras = types.SimpleNamespace(initializers=types.SimpleNamespace(glorot_normal=_GlorotNormal))
    contrib = types.ModuleType("tensorflow.contrib")
    km = types.ModuleType("tensorflow.contrib.kernel_methods")
    km.RandomFourierFeatureMapper = object
    dist = types.ModuleType("tensorflow.contrib.distributions")
    dist.fill_triangular = fill_triangular
    contrib.kernel_methods, contrib.distributions = km, dist
    tf.contrib = contrib
    tf.__version__ = "1.7.0-shim"
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow.contrib"] = contrib
    sys.modules["tensorflow.contrib.kernel_methods"] = km
    sys.modules["tensorflow.contrib.distributions"] = dist
    return tf
