"""The 1-d toy task family of BASELINE config 1 (reference distBO/data/toy_datasets.py:39-52,
distBO/data/generate_data.py:31-59, distBO/utils.py:99-107) and the 'counter' family of config 4
(toy_datasets.py:22-37, model_embed/meta_fea.py:170-177, generate_data.py:94-112)."""
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


def counter_dataset(sig_dim, noise=0.5, data_size=1000, dimension=5, random_seed=23):
    """x ~ N(0, 2^2)^dimension with the sign of column sig_dim-1 tied to x0 x1; y = log(1 + (x0 x1 x_sig)^3) + noise
    (model_embed/meta_fea.py:170-177; draws from the GLOBAL numpy stream reseeded with random_seed, as the reference)."""
    assert 2 < sig_dim <= dimension
    np.random.seed(random_seed)
    x = np.random.normal(loc=0.0, scale=2.0, size=(data_size, dimension))
    x[:, sig_dim - 1] = np.sign(np.multiply(x[:, 0], x[:, 1])) * np.absolute(x[:, sig_dim - 1])
    y = np.log(1 + (x[:, 0] * x[:, 1] * x[:, sig_dim - 1]) ** 3) + np.random.normal(size=data_size, scale=noise)
    return x, y


def counter_dataset_make(n_train, n_test, dim, sig_dim=3, seed=23, noise=0.5, name=None, embed_size=10000,
                         preprocess='standardise'):
    """Standardised features, labels rescaled to [0, 1], seeded train/test split (toy_datasets.py:22-37)."""
    from sklearn import preprocessing
    x, y = counter_dataset(sig_dim, noise=noise, data_size=n_train + n_test, dimension=dim, random_seed=seed)
    if preprocess == 'standardise':
        x = preprocessing.scale(x)
        y = (y - np.min(y)) / (np.max(y) - np.min(y))
    return data_split(x, y, n_train, n_test, check_random_state(seed), embed_size=embed_size, name=name)


def counter_manual_tasks(data_seed, noise=0.5, n_train=1000, n_test=5000, max_embed_size=10000):
    """(target, [4 sources], names): the target's signal dimension is 4, the sources' 3..6; six features
    (generate_data.py:94-112)."""
    rs = check_random_state(data_seed)
    seeds = rs.randint(2 ** 32, size=4 + 1)
    cfg = dict(dim=6, preprocess='standardise', embed_size=max_embed_size, n_train=n_train, n_test=n_test)
    names = ['counter_noise_{}_dim_{}'.format(noise, 4)]
    target = counter_dataset_make(sig_dim=4, seed=seeds[0], noise=noise, name=names[0], **cfg)
    sources = []
    for i in range(1, 6 - 1):
        names.append('counter_noise_{}_dim_{}'.format(noise, i + 2))
        sources.append(counter_dataset_make(sig_dim=i + 2, seed=seeds[i], noise=noise, name=names[-1], **cfg))
    return target, sources, names
