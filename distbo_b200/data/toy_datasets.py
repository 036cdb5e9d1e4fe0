"""The 1-d toy task family of BASELINE config 1 (reference distBO/data/toy_datasets.py:39-52,
distBO/data/generate_data.py:31-59, distBO/utils.py:99-107)."""
from __future__ import annotations

import numpy as np
from sklearn.utils import check_random_state

from .data import data


def data_split(p_data, s_label, n_train, n_test, rs, embed_size=10000, name=None):
    order = rs.permutation(n_train + n_test)
    tr, te = order[:n_train], order[n_train:]
    return data(p_data[tr], p_data[te], s_label[tr], s_label[te], embed_size=embed_size, name=name)


def one_d_split(n_train, n_test, sd=1.0, dim=None, preprocess='standardise', embed_size=10000,
                true_dist=True, seed=23):
    """x ~ N(mu, 1) with mu ~ N(0, sd) ("true" family) or N(4, sd) ("fake" family); y is unused."""
    rs = check_random_state(seed)
    mu = rs.normal(loc=0.0 if true_dist else 4.0, scale=sd, size=1)[0]
    X = rs.normal(loc=mu, scale=1.0, size=(n_train + n_test, 1))
    Y = np.array([rs.normal(loc=float(x), scale=1.0, size=1) for x in X.ravel()])
    return data_split(X, Y, n_train, n_test, rs, embed_size=embed_size), mu


def one_dim_split_tasks(data_seed, fake_source_num, true_source_num, mu_sd=1.0, n_train=500, n_test=500,
                        max_embed_size=10000):
    """(target, [sources], [mu_target, mu_sources...]) exactly as generate_data does for
    `one_dim_split`: one seed per task from RandomState(data_seed).randint(2**32)."""
    rs = check_random_state(data_seed)
    n_src = fake_source_num + true_source_num
    seeds = rs.randint(2 ** 32, size=n_src + 1)
    cfg = dict(sd=mu_sd, n_train=n_train, n_test=n_test, embed_size=max_embed_size)
    target, mu_t = one_d_split(seed=seeds[0], true_dist=True, **cfg)
    sources, supp = [], [mu_t]
    for i in range(n_src):
        s, mu = one_d_split(seed=seeds[1 + i], true_dist=(i >= fake_source_num), **cfg)
        sources.append(s)
        supp.append(mu)
    return target, sources, supp
