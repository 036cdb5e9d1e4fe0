"""Surrogate facade: same interface as the reference's ``gp_regressor`` (train/gp_regressor.py:19-223)
over the sm_100a engine.

    gp = gp_regressor(kernel, target_X=..., target_Y=..., model_type=..., seed=..., float64=True)
    loss = gp.maximise_likhd(params, y, **max_lkd_dict)      # Adam on the GP marginal likelihood
    gp.initialise_predict()
    mean, std = gp.predict_y(params_pred)                    # numpy [M,1], [M,1]
    sim = gp.similarity()
    gp.close()

plus the fused form the BO loop uses when it only needs the maximiser (UtilityFunction.argmax):

    value, index = gp.acquire(params_pred, "ei", y_max=..., xi=...)

State kept across calls, as in the reference: the StandardScaler (frozen after the first fit,
:58-64), the trainable parameters and Adam slots (:141-143).  Every arithmetic step runs in
libdistbo_b200.so on the GPU; there is no CPU fallback (GPEngine raises without a device).
GP arithmetic is fp64 for both values of `float64` (the flag only selects the storage type the
datasets' rows are embedded from at prediction time: fp32 when False).
"""
from __future__ import annotations

import collections
import functools
import math
import os

import numpy as np
import torch
from sklearn.preprocessing import StandardScaler
from sklearn.utils import check_random_state

from .. import blr_engine as _blr
from .. import engine as _eng
from ..sharding import ShardGroup
from .gp_utils import EmbedBatchSampler, PyShuffle, check_dims, init_net_weights

EARLY_STOP_THRESHOLD = 0.0001   # train/train.py:54
EARLY_STOP_PATIENCE = 100       # train/train.py:55
LOSS_HISTORY_CAP = 1 << 22      # device ring of per-epoch losses (the shipped configs run <= 10 000 epochs per fit)


def _on_engine_device(fn):
    """Public entry points run with the surrogate's device current (kernel launches, streams and graph capture
    follow the CURRENT device): a process may hold regressors on several GPUs."""
    @functools.wraps(fn)
    def wrapped(self, *a, **kw):
        dev = self.engine.dev if self.engine is not None else None
        if dev is None and self.device is not None:
            dev = torch.device(self.device)
        if dev is None or dev.index is None or torch.cuda.current_device() == dev.index:
            return fn(self, *a, **kw)
        with torch.cuda.device(dev):
            return fn(self, *a, **kw)
    return wrapped


class TrainControl:
    """Device-side state of one fit's training loop (include/distbo_b200.h DISTBO_CTRL_*): the early-stopping rule
    of train/train.py:53-65 is evaluated by the Adam launch itself, the losses go to a device ring, and the host
    only looks at an asynchronous copy of this block every few epochs."""

    def __init__(self, dev, max_epochs, first_early_stop_epoch):
        init = np.zeros(_eng.CTRL_SLOTS)
        init[_eng.CTRL_CUR_MIN] = init[_eng.CTRL_COUNTDOWN] = np.inf
        init[_eng.CTRL_FIRST_EPOCH] = float(first_early_stop_epoch)
        init[_eng.CTRL_THRESHOLD] = EARLY_STOP_THRESHOLD
        init[_eng.CTRL_PATIENCE] = EARLY_STOP_PATIENCE
        self.dev = torch.from_numpy(init).to(dev)
        self.loss_hist = torch.zeros(max(1, min(int(max_epochs), LOSS_HISTORY_CAP)), dtype=torch.float64, device=dev)
        self._free, self._pending = [], collections.deque()

    def post_check(self):
        """Queues an asynchronous copy of the control block into pinned memory."""
        slot = self._free.pop() if self._free else (torch.empty(_eng.CTRL_SLOTS, dtype=torch.float64).pin_memory(),
                                                    torch.cuda.Event())
        slot[0].copy_(self.dev, non_blocking=True)
        slot[1].record()
        self._pending.append(slot)

    def stopped(self):
        """True once a completed check shows STOP != 0.  Never blocks."""
        seen = False
        while self._pending and self._pending[0][1].query():
            slot = self._pending.popleft()
            seen = seen or slot[0][_eng.CTRL_STOP].item() != 0.0
            self._free.append(slot)
        return seen

    def final(self):
        """Blocking read of the block after the last epoch has been queued."""
        return self.dev.cpu().numpy()


class gp_regressor(object):
    def __init__(self, kernel, n_cpus=1, target_X=None, target_Y=None, model_type=None, seed=None,
                 target_embed_ind=None, float64=False):
        if kernel not in ("dist", "manual", "params", "multi"):
            raise ValueError("kernel %r: expected 'dist', 'manual', 'params' or 'multi'" % (kernel,))
        self.model_type = model_type
        self.kernel = kernel
        self.n_cpus = n_cpus
        self.float64 = bool(float64)
        self.seed = seed
        self.initialise = True
        self.params_scaler = None
        self.initial_weights = None      # optional {name: array} injected before the first fit
        self.shuffle_seed = 0
        self.engine = None
        self.shards = ShardGroup()
        self.fit_rank = 0                # with torch.distributed: the rank that runs the Adam loop
        self.device = None               # optional "cuda:k" the engine is built on (default: the current device)
        self.check_every = int(os.environ.get("DISTBO_CHECK_EVERY", "8"))   # epochs between stop-flag checks
        self._pyshuffle = None           # the stratified sampler's shuffle stream: ONE per regressor, never re-seeded
        self.handoff_ms = None
        self.loss_history = []
        self._fit_generation = 0
        self._predict_generation = -1
        self._pred_sets = None
        self._target_dev = None
        if target_X is not None:
            self.data_sizes_target = check_dims([len(target_X)])
            self.target_embed_ind = target_embed_ind
            self.target_X, self.target_Y = target_X, target_Y

    # ------------------------------------------------------------------ helpers
    def _dev(self, a, dtype=torch.float64):
        return torch.from_numpy(np.ascontiguousarray(a)).to(self.engine.dev).to(dtype)

    def _stack(self, X_list, Y_list, dtype=torch.float64):
        sizes = [len(x) for x in X_list]
        off = np.r_[0, np.cumsum(sizes)].astype(np.int32)
        X = self._dev(np.vstack([np.asarray(x, dtype=np.float64) for x in X_list]), dtype)
        Y = self._dev(np.vstack([check_dims(np.asarray(y, dtype=np.float64)) for y in Y_list]), dtype)
        return self._join_xy(X, Y), Y, torch.from_numpy(off).to(self.engine.dev)

    def _join_xy(self, X, Y):
        """'nn' operator: the net reads the rows [x | y] (tf.concat([x, y], axis=1), train/distbo_net.py:19)."""
        if getattr(self, "embed_op", None) == "nn" and self.kernel == "dist":
            return torch.cat([X, Y.to(X.dtype)], 1).contiguous()
        return X

    def _build_engine(self, weight_dim_x, weight_dim_y, embed_op, concat_data_size, bw_params_init,
                      likhd_var_init, target_embed, n_tasks=0, weight_dim=None, weight_dim_xy=None):
        d = self.in_dim_params
        net = None
        P = 0
        weights = {}
        if self.kernel == "dist":
            if embed_op not in ("none", "joint", "concat", "nn"):
                raise ValueError("embed_op %r: expected 'none', 'joint', 'concat' or 'nn'" % (embed_op,))
        if self.kernel == "dist" and embed_op == "nn":
            # one net over the concatenated [x | y] rows, two or three layers (train/distbo_net.py:18-20,
            # log_gaus_likelihood.py:51-55); pooling = plain mean (+ class ratios on classification data, :43-48)
            wd = [int(v) for v in weight_dim_xy]
            if len(wd) not in (2, 3):
                raise ValueError("weight_dim_xy must have 2 or 3 entries (train/gp_utils.py:13-26), got %r" % (wd,))
            din = self.in_dim_x + self.in_dim_y
            weights.update(init_net_weights(din, wd, "xy", self.seed))
            net = _eng.NetSpec(din, self.in_dim_y, wd[0], wd[-1], 0, 0,
                               y_mode=4 if self.model_type == "classification" else 0,
                               pm=wd[1] if len(wd) == 3 else 0, xname="xy")
            P = net.pooled_width + (1 if concat_data_size else 0)
        elif self.kernel == "dist":
            if len(weight_dim_x) not in (2, 3) or (weight_dim_y is not None and len(weight_dim_y) not in (0, 2)):
                raise NotImplementedError("the sm_100a embedding kernel implements x nets of two or three layers and "
                                          "y nets of two (the reference's CLI builds two-layer x / y nets only, "
                                          "experiments/train_test.py:195-202)")
            weights.update(init_net_weights(self.in_dim_x, weight_dim_x, "x", self.seed))
            hx, p = int(weight_dim_x[0]), int(weight_dim_x[-1])
            pm = int(weight_dim_x[1]) if len(weight_dim_x) == 3 else 0
            if self.model_type == "classification":
                net = _eng.NetSpec(self.in_dim_x, self.in_dim_y, hx, p, 0, 0, y_mode=2, pm=pm)
            elif embed_op in ("joint", "concat"):
                weights.update(init_net_weights(self.in_dim_y, weight_dim_y, "y", self.seed))
                net = _eng.NetSpec(self.in_dim_x, self.in_dim_y, hx, p, int(weight_dim_y[0]),
                                   int(weight_dim_y[1]), y_mode=1 if embed_op == "joint" else 3, pm=pm)
                if embed_op == "concat":
                    weights["log_reg"] = 0.0       # log(1.0), log_gaus_likelihood.py:45-48
            else:
                net = _eng.NetSpec(self.in_dim_x, self.in_dim_y, hx, p, 0, 0, y_mode=0, pm=pm)
            # width of the embedding: train/log_gaus_likelihood.py:33-59
            P = net.pooled_width + (1 if concat_data_size else 0)
        elif self.kernel == "manual":
            P = len(target_embed[0])
        if self.algorithm == "BLR":
            # log_gaus_likelihood_feature.py:30-79: feature row = [embedding part | theta]
            if self.kernel == "multi":
                P = int(n_tasks)
            weights.update(init_net_weights(P + d, weight_dim, "", self.seed))
            e = _blr.BLREngine(d, weight_dim, n_embed=P, net=net, device=self.device)
            for k, v in weights.items():
                e.params.set(k, v)
            if self.initial_weights:
                for k, v in self.initial_weights.items():
                    e.params.set(k, v)
            e.params.set("log_sigma", 0.5 * math.log(likhd_var_init))
            self.engine = e
            return
        n_multi = 0
        if self.kernel == "multi":
            # log_gaus_likelihood.py:68-72 (np.random.seed(seed); np.random.normal): the same legacy stream,
            # drawn from a private RandomState so the caller's global generator is left alone
            n_multi = int(n_tasks)
            seed = self.seed if self.seed is not None else 23
            weights["log_triangular_task"] = np.random.RandomState(seed).normal(size=n_multi * (n_multi + 1) // 2)
        e = _eng.GPEngine(d, n_embed=P, net=net, n_multi=n_multi, device=self.device)
        for k, v in weights.items():
            e.params.set(k, v)
        if self.initial_weights:
            for k, v in self.initial_weights.items():
                e.params.set(k, v)
        e.params.set("log_bw", np.log(np.full(d, float(bw_params_init))))
        e.params.set("log_sigma", 0.5 * math.log(likhd_var_init))   # gp_regressor.py:83
        self.engine = e

    # ------------------------------------------------------------------ fit
    @_on_engine_device
    def maximise_likhd(self, params, y, data_X=None, data_Y=None, data_embed=None,
                       source_embed_ind=None, embed_op='joint', algorithm='GP',
                       weight_dim=None, weight_dim_x=None, num_of_rff=100,
                       weight_dim_xy=None, weight_dim_y=None, batch_size=250,
                       sizes=None, target_embed=None, concat_data_size=False,
                       learning_rate=0.001, max_epochs=200, max_size=None,
                       bw_params_init=1.0, likhd_var_init=0.01, stratify=True,
                       gradient_clip=False, first_early_stop_epoch=30):
        assert len(params) == len(y), 'Configurations and y must be same length'
        if algorithm not in ('GP', 'BLR'):
            raise ValueError("algorithm %r: expected 'GP' or 'BLR'" % (algorithm,))
        if algorithm == 'BLR' and num_of_rff:
            raise NotImplementedError("BLR with a random Fourier map (num_of_rff > 0, feature_map.py:19-26) is not on "
                                      "the sm_100a path; every shipped config uses num_of_rff = 0")
        self.target_embed = target_embed
        self.algorithm = algorithm
        self.embed_op = embed_op
        self.concat_data_size = bool(concat_data_size) and self.kernel == "dist"
        if self.kernel == 'dist':
            assert len(params) == np.sum(sizes), 'Total sizes must match with number of configurations'
            assert len(data_X) == len(sizes), 'Number of datasets must equal number of sizes'
            assert len(data_X) == len(data_Y), 'Number of datasets X must be equal to number of labels'
            for m, n in zip(data_X, data_Y):
                assert len(m) == len(n), 'Number of datapoints must equal number of labels'
            self.in_dim_x = np.shape(data_X[0])[1]
            self.in_dim_y = np.shape(check_dims(data_Y[0]))[1]
        self.in_dim_params = np.shape(params)[1]
        params = np.asarray(params, dtype=np.float64)
        if self.params_scaler is None:                      # fitted once, then frozen
            self.params_scaler = StandardScaler()
            self.params_scaler.fit(params)
        theta = self.params_scaler.transform(params)

        if self.initialise:
            self._build_engine(weight_dim_x, weight_dim_y, embed_op, concat_data_size, bw_params_init,
                               likhd_var_init, target_embed, n_tasks=len(sizes) if sizes is not None else 0,
                               weight_dim=weight_dim, weight_dim_xy=weight_dim_xy)
        e = self.engine
        task_sizes = None if sizes is None or self.kernel == "params" else np.asarray(sizes).reshape(-1)
        e.set_observations(theta, np.asarray(y, dtype=np.float64).reshape(-1), task_sizes)

        self._E_const = None
        self._ratio = None
        sampler = None
        if self.kernel == 'dist':
            data_Y = [check_dims(np.asarray(yy)) for yy in data_Y]
            self._data_sizes = np.asarray([len(x) for x in data_X], dtype=np.float64)
            self._max_size = max_size
            if self.concat_data_size:
                self._ratio = self._dev(self._data_sizes / float(max_size))
            # rows the embeddings are computed from at prediction time (gp_regressor.py:95-102)
            if source_embed_ind is None:
                source_embed_ind = [np.arange(len(x)) for x in data_X]
            self._source_embed_ind = source_embed_ind
            self._pred_sets = None
            if self._pyshuffle is None and stratify and self.model_type == "classification" and \
                    any(len(yy) > int(batch_size) for yy in data_Y):
                self._pyshuffle = PyShuffle(self.shuffle_seed)
            sampler = EmbedBatchSampler(data_Y, batch_size, check_random_state(self.seed), self.model_type,
                                        stratify, shuffle_seed=self.shuffle_seed, pyrand=self._pyshuffle)
            self._full = self._stack(data_X, data_Y)          # all rows of every dataset, HBM resident
            full_off = self._full[2].cpu().numpy().astype(np.int64)
            bs = np.r_[0, np.cumsum(sampler.sizes())].astype(np.int32)
            batch_off = torch.from_numpy(bs).to(e.dev)
        elif self.kernel == 'manual':
            self._E_const = self._dev(np.asarray(data_embed, dtype=np.float64))
        elif self.kernel == 'multi' and algorithm == 'BLR':
            self._E_const = torch.eye(e.T, dtype=torch.float64, device=e.dev)   # task_features, feature_map.py:37-46

        self._sampler, self._full_off = sampler, (full_off if sampler is not None else None)
        self._batch_off = batch_off if sampler is not None else None
        if self.initialise or not hasattr(self, "_lr"):
            # the optimiser is built once, at the first fit: later calls keep ITS learning rate and clipping
            # (train/train.py:29-41 runs only when `initialise`, gp_regressor.py:130-143)
            self._lr, self._clip = learning_rate, (5.0 if gradient_clip else 0.0)
        self._scaler_dev = (self._dev(self.params_scaler.mean_), self._dev(self.params_scaler.scale_))
        self._ctrl = None
        # ---- training loop: train/train.py:11-71
        if first_early_stop_epoch is None:
            first_early_stop_epoch = max_epochs // 5
        max_epochs = int(max_epochs)
        if self.shards.world > 1 and self.shards.rank != self.fit_rank:
            # the fit is single-GPU (SURVEY 8e "replicas only"): this rank waits for the fitting rank's result
            # and receives the posterior state in _prepare_predict
            loss = self._receive_fit_result()
        else:
            loss, failure = None, None
            try:
                loss = self._train(max_epochs, first_early_stop_epoch)
            except _eng._lib.DistboError as exc:
                failure = exc
            if self.shards.world > 1:
                self._send_fit_result(loss, failure)
            if failure is not None:
                raise failure
        self._fit_generation += 1
        self.initialise = False
        self.pre_compute_embed = True
        return loss

    def _train(self, max_epochs, first_early_stop_epoch):
        """The Adam loop with NO host synchronisation per epoch: every epoch is one CUDA-graph replay whose last
        launch (Adam) also records the loss and applies the early-stopping rule on the device; the host queues
        epochs, looks at an asynchronously copied stop flag every `check_every` epochs and, once it is set, stops
        queueing - the epochs already queued behind the stop are no-ops on the device."""
        e = self.engine
        self._setup_epoch_state()
        ctrl = self._ctrl = TrainControl(e.dev, max_epochs, first_early_stop_epoch)
        step0 = e.params.step
        self._shuffle_marks = []
        for epoch in range(max_epochs):
            self.run_epoch()
            if (epoch + 1) % self.check_every == 0:
                ctrl.post_check()
                if ctrl.stopped():
                    break
        final = ctrl.final()
        done = int(final[_eng.CTRL_EPOCHS])
        e.params.step = step0 + done                    # Adam's step count = updates actually applied
        if done:
            self.loss_history.extend(ctrl.loss_hist[:min(done, ctrl.loss_hist.numel())].cpu().numpy().tolist())
        self._rewind_shuffle_stream(done)
        self.last_epoch = done - 1
        if int(final[_eng.CTRL_STOP]) == 2:
            e.check_info(final[_eng.CTRL_INFO])         # raises, as TF's InvalidArgumentError would
        return float(final[_eng.CTRL_LAST_LOSS]) if done else None

    def _rewind_shuffle_stream(self, epochs_done):
        """Mini-batches are drawn ahead of the device; the reference draws one per epoch it runs (its generator
        is advanced only after the stop test, train/train.py:43-65).  The shuffle stream persists across fits, so
        it is put back to where it stood after the draw of the last epoch that ran."""
        if self._pyshuffle is None or not self._shuffle_marks:
            return
        keep = [st for ep, st in self._shuffle_marks if ep < max(epochs_done, 1)]
        self._pyshuffle.restore(keep[-1] if keep else self._shuffle_marks[0][1])

    def _send_fit_result(self, loss, failure):
        rec = torch.tensor([0.0 if failure is None else 1.0, float("nan") if loss is None else loss,
                            float(getattr(self, "last_epoch", -1))], dtype=torch.float64, device=self.engine.dev)
        self.shards.broadcast_fit([rec], src=self.fit_rank)

    def _receive_fit_result(self):
        rec = torch.zeros(3, dtype=torch.float64, device=self.engine.dev)
        self.shards.broadcast_fit([rec], src=self.fit_rank)
        rec = rec.cpu().numpy()
        if rec[0] != 0.0:
            raise _eng._lib.DistboError("the fit failed on rank %d (Cholesky decomposition was not successful)"
                                        % self.fit_rank)
        self.last_epoch = int(rec[2])
        return None if np.isnan(rec[1]) else float(rec[1])

    # One optimisation step = one CUDA graph.  The step's launch sequence (embedding forward, task Gram,
    # Gram, POTRF with its look-ahead stream, TRTRI, LAUUM, likelihood, gradient contractions, embedding
    # backward, Adam) is identical from epoch to epoch, so it is captured once per fit and replayed; the
    # only per-epoch inputs - the mini-batch row indices and Adam's lr_t - reach the graph through
    # static device buffers refreshed from a small ring of pinned host slots.
    RING = 4

    def _setup_epoch_state(self):
        e, sampler = self.engine, self._sampler
        self._graph = None
        self._graph_warm = 0
        self._epochs_staged = 0
        if not hasattr(self, "_shuffle_marks"):
            self._shuffle_marks = None
        self._lrt_dev = torch.zeros(1, dtype=torch.float64, device=e.dev)
        self._ring = []
        self._ring_pos = 0
        n_idx = 0
        if sampler is not None and not sampler.constant:
            n_idx = int(sum(sampler.sizes()))
            self._gather_dev = torch.zeros(n_idx, dtype=torch.int64, device=e.dev)
            self._bX = torch.empty(n_idx, self._full[0].shape[1], dtype=torch.float64, device=e.dev)
            self._bY = torch.empty(n_idx, self._full[1].shape[1], dtype=torch.float64, device=e.dev)
        for _ in range(self.RING):
            self._ring.append({"idx": torch.empty(max(n_idx, 1), dtype=torch.int64).pin_memory(),
                               "lrt": torch.empty(1, dtype=torch.float64).pin_memory(),
                               "ev": torch.cuda.Event()})
        self._use_graph = os.environ.get("DISTBO_GRAPH", "1") != "0"

    def _stage_inputs(self):
        """Host side of one epoch: draw the mini-batch rows, compute lr_t, queue both copies."""
        e, sampler = self.engine, self._sampler
        slot = self._ring[self._ring_pos]
        self._ring_pos = (self._ring_pos + 1) % self.RING
        slot["ev"].synchronize()                 # the copies that last used this slot have completed
        e.params.step += 1
        slot["lrt"][0] = e.adam_lr_t(self._lr, e.params.step)
        self._lrt_dev.copy_(slot["lrt"], non_blocking=True)
        if sampler is not None and not sampler.constant:
            fo = self._full_off
            rows = [np.arange(fo[i], fo[i + 1]) if ix is None else fo[i] + ix
                    for i, ix in enumerate(sampler.next_indices())]
            slot["idx"].copy_(torch.from_numpy(np.concatenate(rows)))
            self._gather_dev.copy_(slot["idx"], non_blocking=True)
            if sampler.uses_shuffle_stream and self._shuffle_marks is not None:
                ps = self._pyshuffle
                if not self._shuffle_marks or self._shuffle_marks[-1][1][0] != ps.version:
                    self._shuffle_marks.append((self._epochs_staged, ps.state()))
        self._epochs_staged += 1
        slot["ev"].record()

    def _epoch_body(self):
        e, sampler = self.engine, self._sampler
        if sampler is None:
            e.loss_and_grad(E_const=self._E_const)
        elif sampler.constant:
            e.loss_and_grad(self._full[0], self._full[1], self._full[2], size_ratio=self._ratio)
        else:
            L = _eng._lib
            for src, dst in ((self._full[0], self._bX), (self._full[1], self._bY)):
                assert src.is_contiguous() and dst.is_contiguous() and src.shape[1] == dst.shape[1]
                L.check(e.lib.distbo_gather_rows(L.ptr(src), src.shape[1] * src.element_size(),
                                                 L.ptr(self._gather_dev), int(self._gather_dev.numel()),
                                                 L.ptr(dst), _eng._stream()), "distbo_gather_rows")
            e.loss_and_grad(self._bX, self._bY, self._batch_off, size_ratio=self._ratio)
        ctrl = self._ctrl
        e.adam_step(self._lr, clip_norm=self._clip, lr_t_dev=self._lrt_dev,
                    ctrl=ctrl.dev if ctrl is not None else None, loss_hist=ctrl.loss_hist if ctrl is not None else None)

    @_on_engine_device
    def run_epoch(self):
        """One optimisation step (loss, full gradient, Adam update) queued on the current stream with
        no host synchronisation: what one sess.run([optimize_step, loss]) is in train/train.py:43-52."""
        self._stage_inputs()
        if self._graph is not None:
            self._graph.replay()
        elif self._use_graph and self._graph_warm >= 1:
            g = torch.cuda.CUDAGraph()
            try:
                # an explicit capture stream ON THE ENGINE'S DEVICE: torch.cuda.graph's default capture stream is
                # created once per process, on whichever device was current then - a capture for another device
                # through it records nothing (and its replays would silently do nothing)
                if getattr(self, "_capture_stream", None) is None or self._capture_stream.device != self.engine.dev:
                    self._capture_stream = torch.cuda.Stream(device=self.engine.dev)
                with torch.cuda.graph(g, stream=self._capture_stream):
                    self._epoch_body()
                self._graph = g
                g.replay()
            except Exception as exc:        # same kernels either way: run the step eagerly
                import warnings
                warnings.warn("distbo_b200: CUDA graph capture of the training step failed (%s); "
                              "launching eagerly" % (exc,))
                self._use_graph = False
                torch.cuda.synchronize()
                self._epoch_body()
        else:
            self._epoch_body()              # first epoch of a fit: eager (allocates persistent buffers)
            self._graph_warm += 1
        self._fit_generation += 1           # (also bumped by maximise_likhd: covers callers that drive run_epoch)

    # ------------------------------------------------------------------ predict
    def initialise_predict(self):
        """The reference builds its prediction graphs here (gp_regressor.py:148-157); the engine has
        nothing to build, the factorisation for prediction is prepared lazily in predict_y."""
        return None

    def _prediction_sets(self):
        """Device-resident rows the embeddings are computed from at prediction time: the `embed_ind`
        subsample of every source dataset followed by the target's (gp_regressor.py:95-102,163-172).
        Built once per fit call and cached: the source rows are gathered ON THE DEVICE from the already
        resident training rows (no second upload); the target set is uploaded once."""
        if self._pred_sets is not None:
            return self._pred_sets
        e = self.engine
        dt = torch.float64 if self.float64 else torch.float32
        full_off = self._full_off
        sizes, rows, identity = [], [], True
        for i, ind in enumerate(self._source_embed_ind):
            ind = np.asarray(list(ind), dtype=np.int64)
            n_i = int(full_off[i + 1] - full_off[i])
            identity = identity and len(ind) == n_i and np.array_equal(ind, np.arange(n_i))
            rows.append(full_off[i] + ind)
            sizes.append(len(ind))
        if identity:
            sX, sY = self._full[0], self._full[1]
        else:
            gather = torch.from_numpy(np.concatenate(rows)).to(e.dev)
            sX, sY = self._full[0].index_select(0, gather), self._full[1].index_select(0, gather)
        if self._target_dev is None:
            tind = list(self.target_embed_ind)
            tX = np.asarray(self.target_X, dtype=np.float64)[tind]
            tY = check_dims(np.asarray(self.target_Y, dtype=np.float64))[tind]
            tYd = self._dev(tY)
            self._target_dev = (self._join_xy(self._dev(tX), tYd), tYd)
        X = torch.cat([sX, self._target_dev[0]], 0).to(dt)
        Y = torch.cat([sY, self._target_dev[1]], 0).to(dt)
        sizes.append(int(self._target_dev[0].shape[0]))
        off = torch.from_numpy(np.r_[0, np.cumsum(sizes)].astype(np.int32)).to(e.dev)
        self._pred_sets = (X, Y, off)
        return self._pred_sets

    def _prepare_predict(self):
        """Posterior state at the CURRENT parameters.  Single process: computed here.  With torch.distributed
        (candidates sharded over the ranks) the fit is single-GPU: the fitting rank computes the state
        and BROADCASTS it - W = L^-1, V, the scaled observations, kvec, the parameter vector, the task Gram - and
        the other ranks only receive (north_star: "L and alpha are broadcast").  `handoff_ms` keeps the device
        times of the last hand-off: {"refresh_ms", "bcast_ms", "bytes"}."""
        if self._predict_generation == self._fit_generation:
            return
        sh = self.shards
        timed = self.engine.dev.type == "cuda"
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if timed else None
        if timed:
            ev[0].record()
        failure = None
        if sh.world == 1 or sh.rank == self.fit_rank:
            try:
                self._compute_posterior_state()
            except _eng._lib.DistboError as exc:
                if sh.world == 1:
                    raise
                failure = exc
        if timed:
            ev[1].record()
        nbytes = 0
        if sh.world > 1:
            status = torch.tensor([0.0 if failure is None else 1.0], dtype=torch.float64, device=self.engine.dev)
            sh.broadcast_fit([status], src=self.fit_rank)
            if status.item() != 0.0:
                raise failure if failure is not None else _eng._lib.DistboError(
                    "the posterior state could not be prepared on rank %d (Cholesky decomposition was not "
                    "successful)" % self.fit_rank)
            state = self._posterior_state_tensors()
            sh.broadcast_fit(state, src=self.fit_rank)
            nbytes = sum(t.numel() * t.element_size() for t in state)
        if timed:
            ev[2].record()
            ev[2].synchronize()
            self.handoff_ms = {"refresh_ms": ev[0].elapsed_time(ev[1]), "bcast_ms": ev[1].elapsed_time(ev[2]),
                               "bytes": nbytes}
        self._predict_generation = self._fit_generation

    @_on_engine_device
    def refresh_posterior(self):
        """Recomputes (and, with torch.distributed, re-broadcasts) the posterior state at the current parameters.
        Collective: with world > 1 every rank must call it.  maximise_likhd makes the next predict_y / acquire do
        this implicitly; callers that step the optimiser themselves through run_epoch() call it explicitly."""
        self._fit_generation += 1
        self._prepare_predict()
        return self.handoff_ms

    def _posterior_state_tensors(self):
        """The tensors scoring and similarity() read, in broadcast order; allocated here on receiving ranks."""
        e = self.engine
        T = e.T
        if self.algorithm == "BLR":
            if e.P and e.emb_target is None:
                e.emb_target = torch.zeros(e.P, dtype=torch.float64, device=e.dev)
            return [e.Linv, e.evec, e.params.flat] + ([e.emb_target] if e.P else [])
        out = [e.W, e.V, e.theta_s, e.sqnorm, e.kvec, e.params.flat]
        if self.kernel != "params":
            shape = (T, T) if self.kernel == "multi" else (T + 1, T + 1)
            if getattr(self, "_KE_all", None) is None or tuple(self._KE_all.shape) != shape:
                self._KE_all = torch.zeros(*shape, dtype=torch.float64, device=e.dev)
            out.append(self._KE_all)
        e.W_valid = True
        e.W_version = getattr(e, "W_version", 0) + 1   # the broadcast overwrites W: the digit planes of the tcgen05 path are stale
        return out

    def _compute_posterior_state(self):
        """Embeddings of the full sample sets, task Gram over sources + target, K, L, W = L^-1, V, and kvec
        (what build_predict / build_predict_dist recompute on every call, train/log_gaus_likelihood.py:128-218)."""
        e = self.engine
        T = e.T
        KE = None
        if self.kernel == "dist":
            X, Y, off = self._prediction_sets()
            E_all = torch.zeros(T + 1, e.P, dtype=X.dtype, device=e.dev)
            e.embed(X, Y, off, out=E_all, dtype=X.dtype)
            E_all = E_all.double()
            if self.concat_data_size:
                E_all[:T, e.P - 1] = self._ratio
                E_all[T, e.P - 1] = float(self.data_sizes_target[0, 0]) / float(self._max_size)
            self._E_all = E_all
        elif self.kernel == "manual":
            self._E_all = torch.cat([self._E_const, self._dev(np.asarray(self.target_embed, dtype=np.float64))], 0)
        if self.algorithm == "BLR":
            # build_predict_feature / build_predict_dist_feature (log_gaus_likelihood_feature.py:145-225): the
            # D x D factor from the training rows; every candidate's feature row carries the target's embedding part
            E = emb_t = None
            if self.kernel in ("dist", "manual"):
                E, emb_t = self._E_all[:T].contiguous(), self._E_all[T].contiguous()
            elif self.kernel == "multi":
                E, emb_t = self._E_const, self._E_const[T - 1].contiguous()
            e.prepare_posterior(E, emb_t)
            e.read_loss()
            return
        if self.kernel == "multi":
            # candidates belong to the last task (kernels.py:47-49): kvec = column T-1 of the task covariance
            KE = e.task_tri()
            self._KE_all = KE
            e.kvec.zero_()
            e.kvec[:e.N] = KE[:, T - 1][e.task_id.long()]
        elif self.kernel in ("dist", "manual"):
            KE_all = e.task_gram(self._E_all, T + 1)
            self._KE_all = KE_all
            KE = KE_all[:T, :T].contiguous()
            e.kvec.zero_()
            e.kvec[:e.N] = KE_all[:T, T][e.task_id.long()]
        else:
            e.kvec.zero_()
            e.kvec[:e.N] = 1.0
        e.build_gram(KE)
        e.factorize(need_inverse=True, need_kinv=False)
        e.likelihood()
        e.read_loss()          # raises on a failed Cholesky, as TF would

    @_on_engine_device
    def predict_y(self, params_pred, refresh_net=False, compute_embed=False):
        self._prepare_predict()
        self.pre_compute_embed = bool(compute_embed)
        # the StandardScaler transform (gp_regressor.py:161) is applied inside the kernel
        out = self.engine.score(np.asarray(params_pred, dtype=np.float64), acq="ucb", kappa=0.0,
                                want_arrays=True, scaler=self._scaler_dev)
        mean = out["mean"].cpu().numpy()[:, None]
        std = out["std"].cpu().numpy()[:, None]
        return mean, std

    @_on_engine_device
    def acquire(self, params_pred, kind, y_max=0.0, xi=0.0, kappa=1.0, want_values=False):
        """Fused posterior + acquisition + argmax over the candidate rows `params_pred` [M,d] (host
        numpy array or host torch tensor, raw scale).  With torch.distributed initialised, every rank passes the SAME array and
        scores its own contiguous shard; the per-GPU (best value, lowest index) are all-gathered.
        Returns (best value, best row index[, acquisition values of the local shard])."""
        self._prepare_predict()
        e = self.engine
        M = len(params_pred)
        lo, hi = self.shards.range(M)
        if hi > lo:
            if torch.is_tensor(params_pred):     # host tensor (pinned for an asynchronous copy)
                cand = params_pred[lo:hi].to(e.dev, non_blocking=True)
            else:
                cand = torch.from_numpy(np.ascontiguousarray(params_pred[lo:hi], dtype=np.float64)).to(e.dev)
            out = e.score(cand, acq=kind, y_max=y_max, xi=xi, kappa=kappa, want_arrays=want_values,
                          index_offset=lo, scaler=self._scaler_dev)
            val, idx = out["best_val"], out["best_idx"]
        else:
            out = {"acq": None}
            val = torch.full((1,), float("-inf"), dtype=torch.float64, device=e.dev)
            idx = torch.full((1,), np.iinfo(np.int64).max, dtype=torch.int64, device=e.dev)
        best_v, best_i = self.shards.argmax(val, idx)
        if want_values:
            return best_v, best_i, (None if out["acq"] is None else out["acq"].cpu().numpy())
        return best_v, best_i

    @_on_engine_device
    def similarity(self, sample_sim=False):
        """[array [T,1]]: embedding-kernel value between every source task and the target
        (gp_regressor.py:205-223; train/train.py:86-95)."""
        if self.algorithm == "BLR":
            return ['none']                      # gp_regressor.py:221-222
        if self.kernel == "params":
            raise ValueError("similarity is defined for the 'dist', 'manual' and 'multi' kernels")
        self._prepare_predict()
        T = self.engine.T
        if self.kernel == "multi":               # sim_network -> net.k_task_pred with sizes all 1 (train.py:91-92)
            return [self._KE_all[:, T - 1:T].cpu().numpy()]
        return [self._KE_all[:T, T:T + 1].cpu().numpy()]

    def close(self):
        self.initialise = True
        if self.engine is not None:
            self.engine.close()         # frees the factorisation plan (streams, events, flag buffers)
        self.engine = None
        self._graph = None
