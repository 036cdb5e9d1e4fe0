"""Hand-crafted dataset meta-features for the `manual` kernel (reference distBO/model_embed/meta_fea.py:21-170).

Host-side statistics computed once per dataset (they are an INPUT of the GP path: `gp_regressor('manual')`
receives them as `data_embed` / `target_embed`); restated for Python 3 and current scikit-learn.  `meta()` fills
`dataset.embed` of every dataset in the list with the MinMax-scaled vector and returns the fitted scaler, exactly
as the reference does.
"""
from __future__ import annotations

import numpy as np
from scipy.stats import entropy, kurtosis, skew
from sklearn.decomposition import PCA
from sklearn.discriminant_analysis import LinearDiscriminantAnalysis
from sklearn.linear_model import LinearRegression
from sklearn.model_selection import train_test_split
from sklearn.naive_bayes import GaussianNB
from sklearn.neighbors import KNeighborsClassifier, KNeighborsRegressor
from sklearn.preprocessing import MinMaxScaler
from sklearn.tree import DecisionTreeClassifier, DecisionTreeRegressor

CLASSIFICATION_MODELS = ('forest', 'svm', 'logistic', 'jaccard_svm', 'logistic_gamma', 'dim_logistic')
REGRESSION_MODELS = ('dim_ridge', 'ridge', 'toy', 'ridge_gamma', 'ridge_alpha')


def _four(v):
    return [np.min(v), np.max(v), np.mean(v), np.std(v)]


def skewness_meta(dataset):
    """min / max / mean / std of the per-feature skewness (meta_fea.py:83-87)."""
    return _four(skew(dataset, axis=0))


def kurtosis_meta(dataset):
    """Same summary of the per-feature excess kurtosis (Fisher; meta_fea.py:90-93)."""
    return _four(kurtosis(dataset, axis=0))


def corr_cov_meta(dataset, include_corr_meta=True):
    """Summaries of the strict upper triangles of the correlation and covariance matrices (meta_fea.py:69-81)."""
    iu = np.triu_indices(dataset.shape[1], k=1)
    corr = _four(np.corrcoef(dataset.T)[iu]) if include_corr_meta else []
    return corr + _four(np.cov(dataset.T)[iu])


def pca_meta(dataset, random_state=None):
    """Skewness and kurtosis of the first principal component and the fraction of components that explain 95 %
    of the variance (meta_fea.py:95-109)."""
    pca = PCA(random_state=random_state)
    first_pc = pca.fit_transform(dataset)[:, 0]
    cumsum = np.cumsum(pca.explained_variance_ratio_)
    intrinsic_dim = int(np.argmax(cumsum > 0.95)) + 1
    return [skew(first_pc), kurtosis(first_pc), float(intrinsic_dim) / dataset.shape[1]]


def landmark_meta(dataset_x, dataset_y, lb=None, model_type='regression', random_state=None):
    """Hold-out scores of cheap learners (meta_fea.py:111-152): 1-NN, linear model / LDA, decision tree and, for
    classification, a fourth score."""
    X_train, X_test, y_train, y_test = train_test_split(dataset_x, dataset_y, test_size=0.25, random_state=random_state)
    if model_type == 'regression':
        neigh = KNeighborsRegressor(n_neighbors=1).fit(X_train, y_train)
        linear = LinearRegression().fit(X_train, y_train)
        tree = DecisionTreeRegressor(random_state=random_state).fit(X_train, y_train)
        return [neigh.score(X_test, y_test), linear.score(X_test, y_test), tree.score(X_test, y_test)]
    if model_type == 'classification':
        y_train, y_test = lb.inverse_transform(y_train), lb.inverse_transform(y_test)
        neigh = KNeighborsClassifier(n_neighbors=1).fit(X_train, y_train)
        lda = LinearDiscriminantAnalysis().fit(X_train, y_train)
        tree = DecisionTreeClassifier(random_state=random_state).fit(X_train, y_train)
        GaussianNB().fit(X_train, y_train)
        # the reference scores the TREE a second time where the naive-Bayes score is meant (meta_fea.py:147-148:
        # `nb_score = tree.score(...)`); kept, because the embedding must be the reference's
        return [neigh.score(X_test, y_test), lda.score(X_test, y_test), tree.score(X_test, y_test),
                tree.score(X_test, y_test)]
    raise ValueError('Must be regression or classification')


def label_meta(dataset_y, lb=None, model_type='regression'):
    """Label statistics (meta_fea.py:154-168): mean / std / skew / kurtosis, or class ratios and their entropy."""
    if model_type == 'regression':
        return [np.mean(dataset_y), np.std(dataset_y), skew(dataset_y)[0], kurtosis(dataset_y)[0]]
    if model_type == 'classification':
        labels = lb.inverse_transform(dataset_y).astype(np.int64)
        count = np.bincount(labels)
        ratio = count[count != 0] / len(labels)
        return ratio.tolist() + [entropy(ratio)]
    raise ValueError('Must be regression or classification')


def meta(data_list, model, max_size=None, scaler=None, random_state=None, include_corr_meta=True):
    """Meta-feature vector of every dataset (rows `embed_ind` of its training split), constant columns removed,
    MinMax-scaled to [0, 1] with `scaler` (fitted here when None); stored in `dataset.embed` (meta_fea.py:21-67)."""
    if model in CLASSIFICATION_MODELS:
        model_type = 'classification'
    elif model in REGRESSION_MODELS:
        model_type = 'regression'
    else:
        raise ValueError('Model not recognised')
    rows = []
    for source in data_list:
        ind = list(source.embed_ind)
        x, y = source.train_x[ind], source.train_y[ind]
        if model != 'toy':
            vec = (skewness_meta(x) + kurtosis_meta(x) + pca_meta(x, random_state=random_state) +
                   landmark_meta(x, y, model_type=model_type, lb=source.lb, random_state=random_state) +
                   corr_cov_meta(x, include_corr_meta=include_corr_meta) +
                   label_meta(y, model_type=model_type, lb=source.lb))
            if max_size is not None:
                vec = vec + [len(source.train_x) / max_size]
        else:
            vec = [np.squeeze(np.mean(x, axis=0)), np.squeeze(np.std(x, axis=0)), np.squeeze(skew(x, axis=0)),
                   np.squeeze(kurtosis(x, axis=0))]
        rows.append(vec)
    table = np.array(rows, dtype=np.float64)
    table = table[:, ~np.all(table[1:] == table[:-1], axis=0)]      # drop columns that are the same everywhere
    if scaler is None:
        scaler = MinMaxScaler()
        scaler.fit(table)
    table = scaler.transform(table)
    for i, source in enumerate(data_list):
        source.embed = table[i, :]
    return scaler
