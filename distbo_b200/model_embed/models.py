"""Objectives being tuned (reference distBO/model_embed/models.py).  They are CPU black boxes outside
the GP hot path; the closed-form toy objective and the two sklearn objectives of the real-data
configs (Parkinson: kernel ridge on random Fourier features; protein: random forest) are kept so the
BO loop can be driven end to end."""
from __future__ import annotations

from functools import partial

import numpy as np


def n_toy(data, theta):
    """exp(-(theta - mean(x))^2 / 2): optimum at the sample mean of the dataset (models.py:239-242)."""
    pooled = np.vstack((data.train_x, data.test_x))
    return np.exp(-0.5 * (theta - np.squeeze(np.mean(pooled))) ** 2)


def ridge(data, log_alpha, log_gamma):
    """R^2 (floored at -1) of ridge regression on 200 random Fourier features drawn under
    np.random.seed(23) (models.py:211-223)."""
    from sklearn.kernel_approximation import RBFSampler
    from sklearn.linear_model import Ridge
    np.random.seed(23)
    rff = RBFSampler(gamma=np.exp(log_gamma), n_components=200)
    train = rff.fit_transform(data.train_x)
    test = rff.transform(data.test_x)
    model = Ridge(alpha=np.exp(log_alpha))
    model.fit(train, data.train_y)
    return max(model.score(test, data.test_y), -1.0)


def dim_ridge(data, log_alpha, log_bw1, log_bw2, log_bw3, log_bw4, log_bw5, log_bw6):
    """R^2 (floored at -1) of ridge regression on 200 random Fourier features (gamma 0.5, random_state 0) of the six
    features divided by per-dimension bandwidths: the 7-d objective of config 4 (models.py:177-195)."""
    from sklearn.kernel_approximation import RBFSampler
    from sklearn.linear_model import Ridge
    bw = np.exp(np.array([log_bw1, log_bw2, log_bw3, log_bw4, log_bw5, log_bw6]))
    rff = RBFSampler(gamma=0.5, n_components=200, random_state=0)
    train = rff.fit_transform(np.divide(data.train_x, bw))
    test = rff.transform(np.divide(data.test_x, bw))
    model = Ridge(alpha=np.exp(log_alpha))
    model.fit(train, data.train_y)
    return max(model.score(test, data.test_y), -1.0)


def jaccard_svm(data, log_C):
    """Test accuracy of an SVM on the precomputed Jaccard kernel of the binary fingerprints: the deterministic
    1-d objective the reference offers for the protein tasks (models.py:122-139)."""
    from scipy.spatial.distance import cdist
    from sklearn.svm import SVC
    train_y, test_y = data.lb.inverse_transform(data.train_y), data.lb.inverse_transform(data.test_y)
    k_train = 1 - cdist(data.train_x, data.train_x, 'jaccard')
    k_test = 1 - cdist(data.test_x, data.train_x, 'jaccard')
    clf = SVC(C=np.exp(log_C), kernel='precomputed')
    clf.fit(k_train, train_y)
    return clf.score(k_test, test_y)


def random_forest(data, depth, min_samples_leaf, min_samples_split, n_trees):
    """Test AUC of an (unseeded, as in the reference) random forest (models.py:225-237)."""
    from sklearn.ensemble import RandomForestClassifier
    from sklearn.metrics import roc_auc_score
    forest = RandomForestClassifier(n_estimators=int(n_trees), max_depth=int(depth),
                                    min_samples_split=min_samples_split, min_samples_leaf=min_samples_leaf)
    forest.fit(data.train_x, data.train_y[:, 0])
    return roc_auc_score(data.test_y[:, 0], np.array(forest.predict_proba(data.test_x)[:, 1]))


_BOUNDS = {
    'toy': (n_toy, 'regression', {'theta': (-8, 8)}, ['theta']),
    'ridge': (ridge, 'regression', {'log_alpha': (np.log(10.0 ** -10), np.log(0.1)),
                                    'log_gamma': (np.log(2.0 ** -7), np.log(2.0 ** 5))},
              ['log_alpha', 'log_gamma']),
    'dim_ridge': (dim_ridge, 'regression',
                  # insertion order as in the reference's dict literal (models.py:84-88): it is the column order of
                  # theta inside the GP and of the candidate stream (target_space.py:47 keeps dict order)
                  dict([('log_bw%d' % i, (np.log(2.0 ** -7), np.log(2.0 ** 6))) for i in range(1, 7)] +
                       [('log_alpha', (np.log(10.0 ** -8), np.log(0.1)))]),
                  ['log_alpha'] + ['log_bw%d' % i for i in range(1, 7)]),
    'jaccard_svm': (jaccard_svm, 'classification', {'log_C': (np.log(2 ** -7), np.log(2 ** 10))}, ['log_C']),
    'forest': (random_forest, 'classification', {'n_trees': (1, 200), 'depth': (1, 32),
                                                 'min_samples_leaf': (0.01, 0.5),
                                                 'min_samples_split': (0.01, 1.0)},
               ['depth', 'min_samples_leaf', 'min_samples_split', 'n_trees']),
}


def register_model(name, fn, model_type, bounds, init_cols=None):
    """Adds a user objective `fn(data, **params)` under `name`: bounds = {param: (lo, hi)}; init_cols = order
    of the parameter columns in previous-evaluation arrays (default: sorted names, as bayes_opt_wrap returns)."""
    _BOUNDS[name] = (fn, model_type, dict(bounds), list(init_cols) if init_cols else sorted(bounds))


def set_model(model, data, init):
    """-> (objective(**params), bounds dict, previous evaluations dict or None, model_type).
    `init` rows are [params sorted by name..., value] (models.py:18-104)."""
    if model not in _BOUNDS:
        raise NotImplementedError("objective %r is a CPU black box outside the GP hot path" % (model,))
    fn, model_type, bounds, init_cols = _BOUNDS[model]
    previous = None
    if init is not None:
        previous = {'target': init[:, -1]}
        for c, name in enumerate(init_cols):
            previous[name] = init[:, c]
    return partial(fn, data), dict(bounds), previous, model_type
