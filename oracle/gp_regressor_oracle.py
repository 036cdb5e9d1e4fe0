"""Oracle surrogate facade: CPU restatement of train/gp_regressor.py + train/train.py
(TEST INFRASTRUCTURE ONLY; pinned by tests/test_reference_golden.py against the reference's own sources run over a
TF op shim, see oracle/__init__.py).

Same duck-typed interface as the reference ``gp_regressor`` so that the BO loop can be
driven with either this class or the CUDA product and the trajectories compared.
"""
from __future__ import annotations


import numpy as np
from sklearn.model_selection import train_test_split
from sklearn.preprocessing import StandardScaler
from sklearn.utils import check_random_state

from . import blr_oracle as BO
from . import gp_oracle as O


def check_dims(array):
    """distBO/utils.py:109-112."""
    array = np.asarray(array)
    if np.ndim(array) == 1:
        array = np.expand_dims(array, 1)
    return array


def binarize_labels(labels, min_bin=0, max_bin=1.01, intervals=4):
    """distBO/utils.py:13-16."""
    bins = np.linspace(min_bin, max_bin, intervals)
    return np.digitize(np.squeeze(labels), bins)


def sample(data_X, data_Y, embed_ind=None, source_list=True):
    """distBO/utils.py:86-97."""
    if not source_list:
        return data_X[list(embed_ind)], data_Y[list(embed_ind)]
    sx, sy = [], []
    for i in range(len(data_X)):
        sx.append(data_X[i][list(embed_ind[i])])
        sy.append(data_Y[i][list(embed_ind[i])])
    return sx, sy


def loop_batches(train, batch_size, max_epochs, rs, model_type, stratify=False, shuffle_seed=0, pyrand=None):
    """train/gp_utils.py:28-104.  The reference's stratified-classification branch shuffles Python lists with
    the unseeded global ``random.shuffle`` (:53,78); a private ``random.Random(shuffle_seed)`` reproduces the stream
    the reference draws after ``random.seed(shuffle_seed)``.  The global stream is never re-seeded by
    the reference, so it runs on from fit to fit: callers that refit pass their own persistent `pyrand`."""
    if "X" in train and "y" in train:
        X_list, Y_list = train["X"], train["y"]
        n_datasets = len(X_list)
        import random
        if pyrand is None:
            pyrand = random.Random(shuffle_seed)
        if model_type == "classification":
            n_classes = Y_list[0].shape[1]
        if model_type == "classification" and stratify:
            cls_list, cls_ratio, cls_index = [], [], []
            for i in range(n_datasets):
                dataY = Y_list[i]
                # LabelBinarizer.inverse_transform (:46-50): binary -> threshold column 0,
                # multi-class -> argmax
                if n_classes == 2:
                    class_y = (dataY[:, 0] > 0.5).astype(int)
                else:
                    class_y = np.argmax(dataY, axis=1)
                c_l, c_r, c_i = [], [], []
                for l in range(n_classes):
                    index = np.where(class_y == l)[0].tolist()
                    pyrand.shuffle(index)
                    c_l.append(index)
                    c_r.append(int((float(len(index)) / len(dataY)) * (batch_size + 4)))
                    c_i.append(0)
                cls_list.append(c_l)
                cls_index.append(c_i)
                cls_ratio.append(c_r)
        while max_epochs > 0:
            xs, ys = [], []
            for i in range(n_datasets):
                dataX, dataY = X_list[i], Y_list[i]
                n = len(dataX)
                if n <= batch_size:
                    sX, sY = dataX, dataY
                elif stratify:
                    if model_type == "classification":
                        class_index = []
                        for k in range(n_classes):
                            end = cls_index[i][k] + cls_ratio[i][k]
                            if end > len(cls_list[i][k]):
                                pyrand.shuffle(cls_list[i][k])
                                cls_index[i][k] = 0
                                end = cls_ratio[i][k]
                            class_index = class_index + cls_list[i][k][cls_index[i][k]:end]
                            cls_index[i][k] = end
                        class_index = class_index[:batch_size]
                        sX, sY = dataX[class_index, :], dataY[class_index, :]
                    else:
                        _, sX, _, sY = train_test_split(dataX, dataY, stratify=binarize_labels(dataY),
                                                        test_size=batch_size, random_state=rs)
                else:
                    perm = rs.permutation(n)[:batch_size]
                    sX, sY = dataX[perm, :], dataY[perm, :]
                xs.append(sX)
                ys.append(sY)
            embed_sizes = check_dims([len(s) for s in xs])
            max_epochs -= 1
            yield np.vstack(xs), np.vstack(ys), embed_sizes
    else:
        while max_epochs > 0:
            max_epochs -= 1
            yield None, None, None


class OracleGPRegressor(object):
    """train/gp_regressor.py:19-223 on the CPU oracle (GP algorithm; kernels dist/manual/params)."""

    def __init__(self, kernel, n_cpus=1, target_X=None, target_Y=None, model_type=None,
                 seed=None, target_embed_ind=None, float64=False):
        self.model_type = model_type
        self.kernel = kernel
        self.seed = seed
        self.initialise = True
        self.params_scaler = None
        self.loss_history = []
        if target_X is not None:
            self.data_sizes_target = check_dims([len(target_X)])
            self.target_embed_ind = target_embed_ind
            self.target_X, self.target_Y = target_X, target_Y

    # -- train -----------------------------------------------------------------------
    def maximise_likhd(self, params, y, data_X=None, data_Y=None, data_embed=None,
                       source_embed_ind=None, embed_op="joint", algorithm="GP",
                       weight_dim=None, weight_dim_x=None, num_of_rff=100,
                       weight_dim_xy=None, weight_dim_y=None, batch_size=250,
                       sizes=None, target_embed=None, concat_data_size=False,
                       learning_rate=0.001, max_epochs=200, max_size=None,
                       bw_params_init=1.0, likhd_var_init=0.01, stratify=True,
                       gradient_clip=False, first_early_stop_epoch=30):
        assert len(params) == len(y)
        assert algorithm in ("GP", "BLR")
        self.num_of_rff = num_of_rff
        self.target_embed = target_embed
        self.algorithm = algorithm
        self.embed_op = embed_op
        if self.kernel == "dist":
            data_sizes = check_dims([len(d) for d in data_X])
            assert len(params) == np.sum(sizes)
            assert len(data_X) == len(sizes) == len(data_Y)
            self.in_dim_x = np.shape(data_X[0])[1]
            self.in_dim_y = np.shape(data_Y[0])[1]
        if sizes is not None:
            sizes = check_dims(sizes)
        self.in_dim_params = np.shape(params)[1]
        if self.params_scaler is None:                       # gp_regressor.py:58-64
            self.params_scaler = StandardScaler()
            self.params_scaler.fit(params)
        params = self.params_scaler.transform(params)

        train = {"h_params": params, "obs": np.asarray(y, dtype=np.float64)}
        if self.initialise:
            kw = dict(likhd_var_init=likhd_var_init, bw_params_init=bw_params_init,
                      seed=self.seed if self.seed is not None else 23)
            if algorithm == "BLR":                           # gp_regressor.py:70-74
                self.n_tasks = len(sizes) if sizes is not None else None
                self.p, self.rff_input_dim = BO.build_params_blr(
                    self.in_dim_params, self.kernel, weight_dim, num_of_rff=num_of_rff,
                    in_dim_x=getattr(self, "in_dim_x", None), in_dim_y=getattr(self, "in_dim_y", None),
                    weight_dim_x=weight_dim_x, weight_dim_y=weight_dim_y, weight_dim_xy=weight_dim_xy,
                    model_type=self.model_type, embed_op=embed_op, concat_data_size=concat_data_size,
                    in_dim_meta=len(target_embed[0]) if self.kernel == "manual" else None,
                    n_tasks=self.n_tasks, likhd_var_init=likhd_var_init, seed=kw["seed"])
                if self.kernel == "dist":
                    self.samp_target_X, self.samp_target_Y = sample(
                        self.target_X, self.target_Y, embed_ind=self.target_embed_ind, source_list=False)
            elif self.kernel == "dist":
                self.p = O.build_params(self.in_dim_params, "dist", in_dim_x=self.in_dim_x,
                                        in_dim_y=self.in_dim_y, weight_dim_x=weight_dim_x,
                                        weight_dim_y=weight_dim_y, weight_dim_xy=weight_dim_xy,
                                        model_type=self.model_type, embed_op=embed_op,
                                        concat_data_size=concat_data_size, **kw)
                self.samp_target_X, self.samp_target_Y = sample(
                    self.target_X, self.target_Y, embed_ind=self.target_embed_ind, source_list=False)
            elif self.kernel == "manual":
                self.p = O.build_params(self.in_dim_params, "manual",
                                        in_dim_meta=len(target_embed[0]), **kw)
            elif self.kernel == "multi":                     # gp_regressor.py:124-129
                self.p = O.build_params(self.in_dim_params, "multi", n_tasks=len(sizes), **kw)
            else:
                self.p = O.build_params(self.in_dim_params, "params", **kw)
            # optional injected initial values {name: array} (TF's glorot stream is not reproducible, so
            # parity runs start every implementation from the same recorded weights)
            for k, v in (getattr(self, "initial_weights", None) or {}).items():
                if k in self.p:
                    self.p[k] = np.asarray(v, dtype=np.float64).reshape(np.shape(self.p[k]))
            self.adam = O.TFAdam(learning_rate)
            # the optimiser (learning rate, clip_by_global_norm) is part of the graph built at the first fit only
            # (train/train.py:29-41): later fits keep it, whatever they pass
            self._gradient_clip = bool(gradient_clip)
        if self.kernel == "dist":
            self.source_X, self.source_Y = sample(data_X, data_Y, embed_ind=source_embed_ind,
                                                  source_list=True)
            self.samp_source_X = np.vstack(self.source_X)
            self.samp_source_Y = np.vstack(self.source_Y)
            train["X"], train["y"] = data_X, data_Y
            train["data_sizes"] = data_sizes
            train["sizes"] = sizes
            train["max_size"] = max_size if concat_data_size else None
        elif self.kernel == "manual":
            train["embed"] = np.asarray(data_embed, dtype=np.float64)
            train["sizes"] = sizes
        elif self.kernel == "multi":
            train["sizes"] = sizes
        self.train = train
        loss = self._train_network(max_epochs, first_early_stop_epoch, batch_size, stratify,
                                   self._gradient_clip)
        self.initialise = False
        self.pre_compute_embed = True
        return loss

    def _batch(self, batch_X, batch_y, embed_sizes):
        t = self.train
        return O.TrainBatch(h_params=t["h_params"], obs=t["obs"], sizes=t.get("sizes"),
                            batch_X=batch_X, batch_y=batch_y, embed_sizes=embed_sizes,
                            data_sizes=t.get("data_sizes"), max_size=t.get("max_size"),
                            embed=t.get("embed"))

    def _train_network(self, max_epochs, first_early_stop_epoch, batch_size, stratify, gradient_clip):
        """train/train.py:11-71."""
        if first_early_stop_epoch is None:
            first_early_stop_epoch = max_epochs // 5
        cur_min, countdown = np.inf, np.inf
        rs = check_random_state(self.seed)
        loss = None
        epoch = -1
        if getattr(self, "_pyrand", None) is None:
            import random
            self._pyrand = random.Random(getattr(self, "shuffle_seed", 0))   # one stream for all fits of this regressor
        for epoch, (bX, by, es) in enumerate(loop_batches(self.train, batch_size, max_epochs, rs,
                                                          self.model_type, stratify=stratify, pyrand=self._pyrand)):
            batch = self._batch(bX, by, es)
            if self.algorithm == "BLR":
                loss, grads = BO.blr_loss_and_grads(self.p, self.kernel, batch, self.rff_input_dim, self.num_of_rff,
                                                    embed_op=self.embed_op, model_type=self.model_type,
                                                    n_tasks=self.n_tasks)
            else:
                loss, grads = O.loss_and_grads(self.p, self.kernel, batch, embed_op=self.embed_op,
                                               model_type=self.model_type)
            self.p = self.adam.step(self.p, grads, clip_norm=5.0 if gradient_clip else None)
            self.loss_history.append(loss)
            if epoch >= first_early_stop_epoch:
                if (cur_min - loss) > 0.0001:
                    countdown = 100
                    cur_min = loss
                else:
                    countdown -= 1
            if epoch >= first_early_stop_epoch and countdown <= 0:
                break
        self.last_epoch = epoch
        return loss

    # -- predict ---------------------------------------------------------------------
    def initialise_predict(self):
        pass

    def predict_y(self, params_pred, refresh_net=False, compute_embed=False):
        """train/gp_regressor.py:159-198."""
        params_pred = self.params_scaler.transform(params_pred)
        B = O.NP
        if self.algorithm == "BLR":
            return self._predict_blr(params_pred, compute_embed)
        if self.kernel == "dist":
            batch = self._batch(self.samp_source_X, self.samp_source_Y,
                                check_dims([len(s) for s in self.source_X]))
            if compute_embed or self.pre_compute_embed:
                # one target embedding, repeated: keep [N,1] and broadcast (identical columns)
                self.dist, self.dist_pred = O.embed_grams(
                    B, self.p, batch, self.samp_target_X, self.samp_target_Y,
                    check_dims([len(self.target_embed_ind)]), self.data_sizes_target, 1,
                    embed_op=self.embed_op, model_type=self.model_type)
                self.pre_compute_embed = bool(compute_embed)
            mean, var = O.predict(B, self.p, "dist", batch, params_pred, k_dist=self.dist,
                                  k_dist_pred=self.dist_pred)
        elif self.kernel == "manual":
            batch = self._batch(None, None, None)
            mean, var = O.predict(B, self.p, "manual", batch, params_pred,
                                  embed_pred=np.asarray(self.target_embed, dtype=np.float64))
        else:
            batch = self._batch(None, None, None)
            mean, var = O.predict(B, self.p, self.kernel, batch, params_pred)
        return mean, O.std_from_var(var)

    def _predict_blr(self, params_pred, compute_embed):
        """gp_regressor.py:163-198 with algorithm == 'BLR' (build_predict_feature / build_predict_dist_feature)."""
        B = O.NP
        M = params_pred.shape[0]
        if self.kernel == "dist":
            batch = self._batch(self.samp_source_X, self.samp_source_Y,
                                check_dims([len(s) for s in self.source_X]))
            if compute_embed or self.pre_compute_embed:     # embed_network -> (dist_fea, dist_fea_pred)
                self.dist = O.dist_features(B, self.p, batch.batch_X, batch.batch_y, batch.sizes,
                                            batch.embed_sizes, data_sizes=batch.data_sizes,
                                            max_size=batch.max_size, embed_op=self.embed_op,
                                            model_type=self.model_type)
                self.dist_pred = O.dist_features(B, self.p, self.samp_target_X, self.samp_target_Y, [1],
                                                 check_dims([len(self.target_embed_ind)]),
                                                 data_sizes=self.data_sizes_target, max_size=batch.max_size,
                                                 embed_op=self.embed_op, model_type=self.model_type)
                self.pre_compute_embed = bool(compute_embed)
            fea = np.concatenate([self.dist, batch.h_params], 1)
            fea_pred = np.concatenate([np.repeat(self.dist_pred, M, axis=0), params_pred], 1)
        else:
            batch = self._batch(None, None, None)
            fea = BO.train_feature(B, self.p, self.kernel, batch, n_tasks=self.n_tasks)
            if self.kernel == "manual":
                emb = np.repeat(np.asarray(self.target_embed, dtype=np.float64), M, axis=0)
                fea_pred = np.concatenate([emb, params_pred], 1)
            elif self.kernel == "multi":
                fea_pred = np.concatenate([BO.task_one_hot(B, self.n_tasks, [M], params_pred, pred=True),
                                           params_pred], 1)
            else:
                fea_pred = params_pred
        nn_fea = BO.nn_features(B, self.p, fea, self.rff_input_dim, self.num_of_rff)
        nn_fea_pred = BO.nn_features(B, self.p, fea_pred, self.rff_input_dim, self.num_of_rff)
        mean, var = BO.blr_predict(B, self.p, nn_fea, batch.obs, nn_fea_pred)
        return mean, O.std_from_var(var)

    def similarity(self, sample_sim=False):
        """train/gp_regressor.py:205-223."""
        if self.algorithm == "BLR":
            return ['none']
        B = O.NP
        if self.kernel == "dist":
            batch = self._batch(self.samp_source_X, self.samp_source_Y,
                                check_dims([len(s) for s in self.source_X]))
            sim = O.similarity(B, self.p, "dist", batch, X_pred=self.samp_target_X,
                               y_pred=self.samp_target_Y,
                               embed_sizes_pred=check_dims([len(self.target_embed_ind)]),
                               data_sizes_pred=self.data_sizes_target,
                               embed_op=self.embed_op, model_type=self.model_type)
        elif self.kernel == "manual":
            batch = self._batch(None, None, None)
            sim = O.similarity(B, self.p, "manual", batch,
                               embed_pred=np.asarray(self.target_embed, dtype=np.float64))
        elif self.kernel == "multi":
            sim = O.similarity(B, self.p, "multi", self._batch(None, None, None))
        else:
            raise ValueError(self.kernel)
        return [sim]

    def close(self):
        self.initialise = True
