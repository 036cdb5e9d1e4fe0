"""Observation store of the BO loop (reference distBO/bayes_opt/target_space.py:12-305)."""
from __future__ import annotations

import numpy as np

from .helpers import ensure_rng, unique_rows


def _hashable(x):
    return tuple(map(float, x))


class TargetSpace(object):
    """Growable (X, Y) record of evaluated points plus the objective and its box bounds.

    Same public surface as the reference: keys, bounds, dim, X, Y, observe_point, add_observation,
    random_points (one scalar draw per coordinate, so that the seed stream matches,
    target_space.py:254-258), max_point, set_bounds.  Repeated points are stored again, as the
    reference does (its duplicate check is commented out, :131-137,171-172)."""

    def __init__(self, target_func, pbounds, random_state=None):
        self.random_state = ensure_rng(random_state)
        self.target_func = target_func
        self.keys = list(pbounds.keys())
        self.bounds = np.array(list(pbounds.values()), dtype=float)
        self.dim = len(self.keys)
        self._Xarr = None
        self._Yarr = None
        self._length = 0
        self._cache = {}

    # -- views --------------------------------------------------------------------------------
    @property
    def X(self):
        return None if self._Xarr is None else self._Xarr[:self._length]

    @property
    def Y(self):
        return None if self._Yarr is None else self._Yarr[:self._length]

    def __contains__(self, x):
        return _hashable(x) in self._cache

    def __len__(self):
        return self._length

    @property
    def _n_alloc_rows(self):
        return 0 if self._Xarr is None else self._Xarr.shape[0]

    # -- conversions --------------------------------------------------------------------------
    def _dict_to_points(self, points_dict):
        lengths = {len(list(points_dict[k])) for k in self.keys}
        if len(lengths) != 1:
            raise ValueError('The same number of initialization points must be entered for every parameter.')
        return [list(row) for row in zip(*(points_dict[k] for k in self.keys))]

    # -- observations -------------------------------------------------------------------------
    def observe_point(self, x):
        x = np.asarray(x).ravel()
        assert x.size == self.dim, 'x must have the same dimensions'
        y = self.target_func(**dict(zip(self.keys, x)))
        self.add_observation(x, y)
        return y

    def add_observation(self, x, y):
        if self._length >= self._n_alloc_rows:
            self._allocate((self._length + 1) * 2)
        x = np.asarray(x).ravel()
        self._cache[_hashable(x)] = y
        self._Xarr[self._length] = x
        self._Yarr[self._length] = y
        self._length += 1

    def _allocate(self, num):
        if num <= self._n_alloc_rows:
            raise ValueError('num must be larger than current array length')
        X = np.empty((num, self.bounds.shape[0]))
        Y = np.empty(num)
        if self._Xarr is not None:
            X[:self._length] = self._Xarr[:self._length]
            Y[:self._length] = self._Yarr[:self._length]
        self._Xarr, self._Yarr = X, Y

    def random_points(self, num):
        data = np.empty((num, self.dim))
        for i in range(num):
            for col, (lower, upper) in enumerate(self.bounds):
                data[i, col] = self.random_state.uniform(lower, upper, size=1)[0]
        return data

    def max_point(self):
        return {'max_val': self.Y.max(), 'max_params': dict(zip(self.keys, self.X[self.Y.argmax()]))}

    def set_bounds(self, new_bounds):
        for row, key in enumerate(self.keys):
            if key in new_bounds:
                self.bounds[row] = new_bounds[key]

    def _assert_internal_invariants(self, fast=True):
        if self._Xarr is None:
            assert self._Yarr is None
        else:
            assert len(self.X) == len(self.Y) == self._length
            assert len(self._Xarr) == len(self._Yarr)
            if not fast:
                assert np.all(unique_rows(self.X))
