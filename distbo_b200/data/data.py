"""Train/test container handed to the BO loop (reference distBO/data/data.py:14-74)."""
from __future__ import annotations

import numpy as np
from sklearn.preprocessing import LabelBinarizer
from sklearn.utils import check_random_state


def check_dims(array):
    array = np.asarray(array)
    return array[:, None] if array.ndim == 1 else array


class data(object):
    """train_x/test_x, train_y/test_y (one-hot with >= 2 columns for classification) and
    `embed_ind`: the rows the dataset embedding is computed from at prediction time (all rows, or a
    seeded subsample of at most `embed_size` rows, class-proportional for classification)."""

    def __init__(self, train_x, test_x, train_y, test_y, name=None, path=None, index=None,
                 embed_size=None, random_state=None, prob_type='regression'):
        self.train_x, self.test_x = train_x, test_x
        self.prob_type = prob_type
        if prob_type == 'regression':
            self.train_y, self.test_y = check_dims(train_y), check_dims(test_y)
            self.lb = None
        elif prob_type == 'classification':
            n_classes = len(np.unique(train_y))
            self.lb = LabelBinarizer()
            if n_classes == 2:      # force two columns for a binary problem
                self.lb.fit(np.squeeze(train_y).tolist() + [2])
            else:
                self.lb.fit(train_y)
            self.train_y, self.test_y = self.lb.transform(train_y), self.lb.transform(test_y)
            if n_classes == 2:
                self.train_y, self.test_y = self.train_y[:, :-1], self.test_y[:, :-1]
        else:
            raise ValueError('Must be regression or classification')
        n = len(train_x)
        if embed_size is None or n <= embed_size:
            self.embed_ind = range(0, n)
        else:
            rs = check_random_state(23) if random_state is None else random_state
            rows = np.arange(n)
            if prob_type == 'regression':
                self.embed_ind = rs.permutation(rows)[:embed_size].tolist()
            else:
                picked = []
                for c in range(n_classes):
                    members = rows[np.squeeze(train_y) == c]
                    share = int(len(members) / len(train_y) * (embed_size + 10))
                    picked += rs.permutation(members)[:share].tolist()
                self.embed_ind = picked[:embed_size]
        self.embed = None
        if name is not None:
            self.name = name
        if path is not None:
            self.path = path
        if index is not None:
            self.index = index

    def __len__(self):
        return len(self.train_x) + len(self.test_x)

    def train_len(self):
        return len(self.train_x)

    def dim(self):
        return self.train_x.shape[1]
