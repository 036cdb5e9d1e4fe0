"""Real-data task families of BASELINE configs 2 and 3 (reference distBO/data/load_datasets.py:14-84),
restated for Python 3.  Each loader returns (target dataset, [source datasets], [names]).

The raw arrays are passed in (or read from a directory of the reference's files) so that the same code serves
the reference's `protein/` and `parkinson/` folders and the small fixtures under tests/golden/.
"""
from __future__ import annotations

import os

import numpy as np
from sklearn import preprocessing
from sklearn.model_selection import train_test_split

from .data import data


def standardise(labels, low=0.0, high=1.0):
    """Min-max rescaling to [low, high] (distBO/utils.py:114-118)."""
    lo, hi = np.min(labels), np.max(labels)
    return low + (labels - lo) / (hi - lo) * (high - low)


def binarize_labels(labels, min_bin=0, max_bin=1.01, intervals=4):
    return np.digitize(np.squeeze(labels), np.linspace(min_bin, max_bin, intervals))


def _split_target(datasets, names, target):
    assert target < len(datasets), 'Target index is more than total group.'
    others = [i for i in range(len(datasets)) if i != target]
    return datasets[target], [datasets[i] for i in others], [names[i] for i in [target] + others]


def protein_tasks(tables, names, target, random_state=None, test_size=0.2):
    """tables: list of [n, 1 + 166 + 1] arrays (column 0 unused, last column = binary label), one per
    ChEMBL target, in sorted file-name order (load_datasets.py:14-45)."""
    out = []
    for table, name in zip(tables, names):
        x, y = table[:, 1:-1], np.expand_dims(table[:, -1], 1)
        tr_x, te_x, tr_y, te_y = train_test_split(x, y, stratify=y, test_size=test_size, random_state=random_state)
        out.append(data(tr_x, te_x, tr_y, te_y, prob_type='classification', random_state=random_state, name=name))
    return _split_target(out, list(names), target)


def parkinson_tasks(patients, target, random_state=None, preprocess='standardise', label='total', test_size=0.2):
    """patients: list of [n, 21] arrays (age, sex, 17 features, motor, total) (load_datasets.py:47-84)."""
    out, names = [], []
    for ind, patient in enumerate(patients):
        x = patient[:, :-2][:, 2:]                      # age and sex are constant per patient
        if preprocess == 'standardise':
            x = preprocessing.scale(x)
        y = np.expand_dims(patient[:, -2] if label == 'motor' else patient[:, -1], 1)
        y = standardise(y)
        tr_x, te_x, tr_y, te_y = train_test_split(x, y, test_size=test_size, stratify=binarize_labels(y),
                                                  random_state=random_state)
        name = 'pat_{}_{}'.format(label, ind)
        out.append(data(tr_x, te_x, tr_y, te_y, name=name))
        names.append(name)
    return _split_target(out, names, target)


def load_protein(target, directory, random_state=None, preprocess='standardise', test_size=0.2):
    files = sorted(os.listdir(directory))
    tables = [np.genfromtxt(os.path.join(directory, f), delimiter=',') for f in files]
    return protein_tasks(tables, files, target, random_state=random_state, test_size=test_size)


def load_parkinson(target, directory, random_state=None, preprocess='standardise', label='total', test_size=0.2):
    patients = [np.load(os.path.join(directory, 'patient_{}.npy'.format(i))) for i in range(42)]
    return parkinson_tasks(patients, target, random_state=random_state, preprocess=preprocess, label=label,
                           test_size=test_size)
