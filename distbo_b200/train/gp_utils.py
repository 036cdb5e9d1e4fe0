"""Host-side helpers of the surrogate (reference distBO/train/gp_utils.py, distBO/utils.py).

Weight initialisation and the per-epoch embedding mini-batch sampler.  Host logic only: the
arithmetic on the sampled rows happens in the CUDA library.
"""
from __future__ import annotations

import math

import numpy as np


def check_dims(array):
    """Column-vector view of a 1-d array (reference distBO/utils.py:109-112)."""
    array = np.asarray(array)
    return array[:, None] if array.ndim == 1 else array


def truncated_glorot(rs, fan_in, fan_out):
    """Glorot-normal weights, |z| <= 2 sigma, float32 values (reference train/gp_utils.py:13-26 uses
    tf.keras.initializers.glorot_normal, a Philox truncated normal that cannot be reproduced
    outside TensorFlow - SURVEY F9; numpy's MT19937 with rejection is used here and callers that
    need a fixed trajectory inject weights through gp_regressor.initial_weights)."""
    sigma = math.sqrt(2.0 / (fan_in + fan_out))
    z = rs.standard_normal(size=(fan_in, fan_out))
    out_of_range = np.abs(z) > 2.0
    while out_of_range.any():
        z[out_of_range] = rs.standard_normal(size=int(out_of_range.sum()))
        out_of_range = np.abs(z) > 2.0
    return (z * sigma).astype(np.float32)


def init_net_weights(in_dim, weight_dim, name, seed):
    """{weights_<name>_1, bias_<name>_1, weights_<name>_2[, bias_<name>_2, weights_<name>_3]}: zero biases, no
    bias on the last layer (reference train/gp_utils.py:13-26)."""
    if weight_dim is None or len(weight_dim) not in (2, 3):
        raise NotImplementedError("feature maps of 2 or 3 layers (weight_dim of length 2 or 3, as every shipped "
                                  "config uses)")
    rs = np.random.RandomState(23 if seed is None else seed)
    h, p = int(weight_dim[0]), int(weight_dim[1])
    out = {
        "weights_%s_1" % name: truncated_glorot(rs, in_dim, h).astype(np.float64),
        "bias_%s_1" % name: np.zeros(h),
        "weights_%s_2" % name: truncated_glorot(rs, h, p).astype(np.float64),
    }
    if len(weight_dim) == 3:
        out["bias_%s_2" % name] = np.zeros(p)
        out["weights_%s_3" % name] = truncated_glorot(rs, p, int(weight_dim[2])).astype(np.float64)
    return out


def label_bins(labels, lo=0.0, hi=1.01, intervals=4):
    """Bin regression labels for stratified sampling (reference distBO/utils.py:13-16)."""
    return np.digitize(np.squeeze(labels), np.linspace(lo, hi, intervals))


class PyShuffle:
    """`random.Random(seed).shuffle` on int64 numpy arrays at C speed: the Mersenne-Twister state of the Python
    generator is handed to `distbo_py_shuffle_i64` (csrc/hostutil.cu), which advances it exactly as CPython's
    shuffle / _randbelow / getrandbits do.  tests/test_bo_host.py checks it against `random.shuffle` itself."""

    def __init__(self, seed):
        import random
        from .. import _lib
        state = random.Random(seed).getstate()[1]
        self._mt = np.ascontiguousarray(np.asarray(state[:624], dtype=np.uint32))
        self._idx = np.asarray([state[624]], dtype=np.int32)
        self._fn = _lib.load().distbo_py_shuffle_i64
        self._C = _lib.C
        self.version = 0            # number of shuffles drawn so far

    def state(self):
        """(version, Mersenne-Twister words, position): restore() rewinds the stream to this point."""
        return self.version, self._mt.copy(), self._idx.copy()

    def restore(self, state):
        self.version = state[0]
        self._mt[:] = state[1]
        self._idx[:] = state[2]

    def shuffle(self, members):
        self.version += 1
        out = np.ascontiguousarray(members, dtype=np.int64).copy()
        vp = self._C.c_void_p
        rc = self._fn(vp(self._mt.ctypes.data), vp(self._idx.ctypes.data), vp(out.ctypes.data), int(out.size))
        if rc:
            raise RuntimeError("distbo_py_shuffle_i64 failed (%d)" % rc)
        return out


class EmbedBatchSampler:
    """Row indices of each epoch's embedding mini-batch, per dataset (reference train/gp_utils.py:28-104).

    Yields, per epoch, a list of index arrays (one per dataset; None = "all rows, in order").
    Draw order from `rs` matches the reference: one permutation (or stratified split) per dataset
    that is larger than batch_size, datasets in order.  The reference's stratified-classification
    branch shuffles Python lists with the unseeded global `random` (gp_utils.py:53,78); here the same
    `random.shuffle` algorithm runs on a private `random.Random(shuffle_seed)`, i.e. the stream the reference
    draws after `random.seed(shuffle_seed)` (tests/golden/ref_dist_class_stratified.npz).  Shuffles
    happen at set-up and when a class runs out of unseen rows; between them the members are numpy arrays, so
    an epoch costs a few slices per dataset.
    """

    def __init__(self, data_Y, batch_size, rs, model_type, stratify, shuffle_seed=0, pyrand=None):
        self.Y = data_Y
        self.n = [len(y) for y in data_Y]
        self.batch_size = int(batch_size)
        self.rs = rs
        self.model_type = model_type
        self.stratify = bool(stratify)
        self.constant = all(n <= self.batch_size for n in self.n)
        self._cls = None
        if not self.constant and self.stratify and model_type == "classification":
            # the reference never re-seeds the global `random`, so its stream runs on from fit to fit: a regressor
            # that refits hands in its own persistent PyShuffle
            self._pyrand = pyrand if pyrand is not None else PyShuffle(shuffle_seed)
            n_classes = data_Y[0].shape[1]
            self._cls = []
            for y in data_Y:
                labels = (y[:, 0] > 0.5).astype(int) if n_classes == 2 else np.argmax(y, axis=1)
                per_class = []
                for c in range(n_classes):
                    members = self._shuffled(np.where(labels == c)[0].astype(np.int64))
                    take = int((float(len(members)) / len(y)) * (self.batch_size + 4))
                    per_class.append({"members": members, "take": take, "pos": 0})
                self._cls.append(per_class)

    def _shuffled(self, members):
        return self._pyrand.shuffle(members)    # Python's list shuffle, as the reference calls it

    @property
    def uses_shuffle_stream(self):
        return self._cls is not None

    def sizes(self):
        """Rows per dataset in every epoch's mini-batch.  Class-stratified batches hold sum_c int(frac_c *
        (batch_size + 4)) rows cut at batch_size (gp_utils.py:57,72-83), which falls below batch_size once there
        are six or more classes; the count is the same in every epoch of a fit."""
        out = []
        for i, n in enumerate(self.n):
            if n <= self.batch_size:
                out.append(n)
            elif self._cls is not None:
                out.append(min(self.batch_size, sum(c["take"] for c in self._cls[i])))
            else:
                out.append(self.batch_size)
        return out

    def next_indices(self):
        out = []
        for i, n in enumerate(self.n):
            if n <= self.batch_size:
                out.append(None)
            elif self.stratify and self.model_type == "classification":
                picked = []
                for c in self._cls[i]:
                    end = c["pos"] + c["take"]
                    if end > len(c["members"]):
                        c["members"] = self._shuffled(c["members"])
                        c["pos"], end = 0, c["take"]
                    picked.append(c["members"][c["pos"]:end])
                    c["pos"] = end
                out.append(np.concatenate(picked)[:self.batch_size])
            elif self.stratify:
                from sklearn.model_selection import train_test_split
                _, idx = train_test_split(np.arange(n), stratify=label_bins(self.Y[i]),
                                          test_size=self.batch_size, random_state=self.rs)
                out.append(np.asarray(idx, dtype=np.int64))
            else:
                out.append(self.rs.permutation(n)[:self.batch_size].astype(np.int64))
        return out
