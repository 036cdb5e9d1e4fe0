"""Acquisition layer of the BO loop (reference distBO/bayes_opt/helpers.py).

`UtilityFunction` and `acq_max` keep the reference's names, arguments and decisions
(helpers.py:9-111,114-177).  When the surrogate is the sm_100a `gp_regressor` the candidate sweep
(posterior mean/variance, EI/UCB/LCB, arg-max; HOT LOOP 2 of SURVEY section 3) is ONE fused device
pass through `gp.acquire`; for any other duck-typed surrogate exposing only `predict_y` the
reference formulae are evaluated from its (mean, std).
"""
from __future__ import annotations

from datetime import datetime

import numpy as np
from scipy.optimize import minimize
from scipy.stats import norm

FUSED_KINDS = ("ei", "ucb", "lcb")


class UtilityFunction(object):
    """Acquisition functions (helpers.py:114-177)."""

    def __init__(self, kind, kappa, xi):
        self.kappa = kappa
        self.xi = xi
        if kind not in ['ucb', 'lcb', 'ei', 'poi', 'max']:
            raise NotImplementedError("The utility function {} has not been implemented, please choose one of "
                                      "ucb, ei, lcb, max or poi.".format(kind))
        self.kind = kind

    # -- fused device path ---------------------------------------------------------------------
    def _fused(self, gp):
        return self.kind in FUSED_KINDS and hasattr(gp, "acquire")

    def argmax(self, x, gp, y_max, compute_embed=False):
        """(max acquisition value, row index of its first occurrence) over the rows of x."""
        if self._fused(gp):
            return gp.acquire(x, self.kind, y_max=y_max, xi=self.xi, kappa=self.kappa)
        ys = self.utility(x, gp, y_max, compute_embed=compute_embed)
        i = int(np.argmax(ys))
        return float(np.reshape(ys, -1)[i]), i

    # -- reference interface -------------------------------------------------------------------
    def utility(self, x, gp, y_max, compute_embed=False):
        if self._fused(gp) and getattr(gp.shards, "world", 1) == 1:
            _, _, ys = gp.acquire(x, self.kind, y_max=y_max, xi=self.xi, kappa=self.kappa, want_values=True)
            return ys if len(ys) > 1 else np.squeeze(ys)
        if self.kind == 'ucb':
            return self._ucb(x, gp, self.kappa, compute_embed=compute_embed)
        if self.kind == 'lcb':
            return self._lcb(x, gp, self.kappa, compute_embed=compute_embed)
        if self.kind == 'ei':
            return self._ei(x, gp, y_max, self.xi, compute_embed=compute_embed)
        if self.kind == 'poi':
            return self._poi(x, gp, y_max, self.xi)
        return self._mean_max(x, gp)

    @staticmethod
    def _moments(x, gp, **kw):
        mean, std = gp.predict_y(x, **kw)
        return np.squeeze(mean), np.squeeze(std)

    @staticmethod
    def _mean_max(x, gp):
        return gp.predict_y(x)[0]

    @staticmethod
    def _ucb(x, gp, kappa, compute_embed=False):
        mean, std = UtilityFunction._moments(x, gp, compute_embed=compute_embed)
        return mean + kappa * std

    @staticmethod
    def _lcb(x, gp, kappa, compute_embed=False):
        mean, std = UtilityFunction._moments(x, gp, compute_embed=compute_embed)
        return mean - kappa * std

    @staticmethod
    def _ei(x, gp, y_max, xi, compute_embed=False):
        mean, std = UtilityFunction._moments(x, gp, compute_embed=compute_embed)
        improvement = mean - y_max - xi
        z = improvement / std
        return improvement * norm.cdf(z) + std * norm.pdf(z)

    @staticmethod
    def _poi(x, gp, y_max, xi):
        mean, std = gp.predict_y(x)
        return norm.cdf((mean - y_max - xi) / std)


def _as_utility(ac):
    """The UtilityFunction behind a bound `utility` method, if that is what `ac` is."""
    owner = getattr(ac, "__self__", None)
    return owner if isinstance(owner, UtilityFunction) else None


class LockstepEvaluator(object):
    """One device call for the pending evaluations of SEVERAL optimiser threads.

    Each L-BFGS-B run of the polish lives in its own thread and calls `evaluate(X)` (rows = points it needs now);
    the call blocks until every still-running thread is waiting, then the last one to arrive evaluates the
    concatenation of all requests with ONE call of `f_batch` and hands every thread its slice.  An evaluation
    depends only on its own row (the scoring kernel treats candidates independently), so each run sees exactly
    the numbers it would see alone - the batching changes the number of device calls, not a single bit."""

    def __init__(self, f_batch, n_threads):
        import threading
        self.f_batch = f_batch
        self.cv = threading.Condition()
        self.active = n_threads
        self.requests = {}          # thread ident -> X
        self.results = {}
        self.calls = 0              # device calls made
        self.points = 0             # rows evaluated
        self.error = None

    def _flush(self):
        """Caller holds the lock and every active thread has a request pending."""
        order = sorted(self.requests)
        X = np.concatenate([self.requests[k] for k in order], axis=0)
        try:
            vals = np.reshape(np.asarray(self.f_batch(X), dtype=np.float64), -1)
        except BaseException as exc:      # hand the failure to every waiting thread
            self.error = exc
            vals = np.full(len(X), np.nan)
        self.calls += 1
        self.points += len(X)
        pos = 0
        for k in order:
            n = len(self.requests[k])
            self.results[k] = vals[pos:pos + n]
            pos += n
        self.requests.clear()
        self.cv.notify_all()

    def evaluate(self, X):
        import threading
        me = threading.get_ident()
        with self.cv:
            self.requests[me] = np.atleast_2d(np.asarray(X, dtype=np.float64))
            if len(self.requests) == self.active:
                self._flush()
            while me not in self.results:
                self.cv.wait()
            out = self.results.pop(me)
            if self.error is not None:
                raise self.error
            return out

    def retire(self):
        with self.cv:
            self.active -= 1
            if self.active > 0 and len(self.requests) == self.active:
                self._flush()


class _SpeculativeObjective(object):
    """Negated acquisition for one L-BFGS-B run, evaluated through a LockstepEvaluator.

    scipy estimates the gradient with its forward-difference stencil (eps given, steps flipped at the upper
    bound): d more evaluations after every f(x).  `fun` asks for x TOGETHER with the d stencil points it expects
    scipy to want next and keeps them; `map` (handed to scipy as its `workers` map, so the stencil is scipy's own,
    point for point) serves them from that cache and evaluates only what is missing.  A wrong guess costs one more
    call, never a different number."""

    def __init__(self, evaluator, bounds, eps):
        self.ev, self.lo, self.hi, self.eps = evaluator, bounds[:, 0], bounds[:, 1], eps
        self.cache = {}

    def _stencil(self, x):
        pts = []
        for i in range(x.size):
            p = x.copy()
            h = self.eps
            if x[i] + h > self.hi[i] and x[i] - h >= self.lo[i]:
                h = -h
            p[i] = x[i] + h
            pts.append(p)
        return pts

    def fun(self, x):
        x = np.asarray(x, dtype=np.float64).ravel()
        hit = self.cache.get(x.tobytes())
        if hit is not None:
            return hit
        pts = [x] + self._stencil(x)
        vals = self.ev.evaluate(np.stack(pts))
        self.cache = {p.tobytes(): -float(v) for p, v in zip(pts, vals)}
        return self.cache[x.tobytes()]

    def map(self, wrapped_fun, xs):
        xs = [np.asarray(x, dtype=np.float64).ravel() for x in xs]
        missing = [x for x in xs if x.tobytes() not in self.cache]
        if missing:
            vals = self.ev.evaluate(np.stack(missing))
            for p, v in zip(missing, vals):
                self.cache[p.tobytes()] = -float(v)
        return [self.cache[x.tobytes()] for x in xs]


POLISH_STATS = {"device_calls": 0, "points": 0, "runs": 0}    # cumulative, for profiles/ (tools/probe_polish.py)


def _polish_runs_batched(ac, gp, y_max, x_seeds, bounds, eps, maxfun):
    """The L-BFGS-B runs of `_polish`, all seeds in lock-step on the fused device path: one `gp.acquire`-style call
    serves the pending evaluations of every run (value + scipy's finite-difference stencil)."""
    import threading
    seeds = [np.ravel(x) for x in x_seeds]
    ev = LockstepEvaluator(lambda X: ac(X, gp=gp, y_max=y_max), len(seeds))
    results = [None] * len(seeds)

    def run(k):
        try:
            obj = _SpeculativeObjective(ev, np.asarray(bounds, dtype=np.float64), eps)
            results[k] = minimize(obj.fun, seeds[k], bounds=bounds, method="L-BFGS-B",
                                  options={'eps': eps, 'ftol': 1e-04, 'maxfun': maxfun, 'maxls': 10,
                                           'workers': obj.map})
        except BaseException as exc:
            results[k] = exc
        finally:
            ev.retire()
    threads = [threading.Thread(target=run, args=(k,), daemon=True) for k in range(len(seeds))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for r in results:
        if isinstance(r, BaseException):
            raise r
    POLISH_STATS["device_calls"] += ev.calls
    POLISH_STATS["points"] += ev.points
    POLISH_STATS["runs"] += len(seeds)
    return results


def _polish_runs_serial(ac, gp, y_max, x_seeds, bounds, eps, maxfun):
    """One L-BFGS-B run after the other, one single-point evaluation per function value (helpers.py:70-86)."""
    return [minimize(lambda x: -float(np.squeeze(ac(x.reshape(1, -1), gp=gp, y_max=y_max))), np.ravel(x_try),
                     bounds=bounds, method="L-BFGS-B",
                     options={'eps': eps, 'ftol': 1e-04, 'maxfun': maxfun, 'maxls': 10}) for x_try in x_seeds]


def _polish(ac, gp, y_max, x_seeds, bounds, evaluatedX, x_max, max_acq, eps, maxfun):
    """L-BFGS-B from each seed with finite-difference gradients (helpers.py:70-86 / :94-107).  A result
    replaces the incumbent only when `evaluatedX` is given and shares no coordinate value with it -
    the reference's `res.x not in evaluatedX` is numpy's element-wise membership test.  On the fused device
    path the runs advance in lock-step and share their device calls (LockstepEvaluator); the runs themselves,
    their stencils and the order in which their results are compared are the reference's."""
    util = _as_utility(ac)
    batched = util is not None and util._fused(gp) and getattr(gp.shards, "world", 1) == 1 and len(x_seeds) > 0
    runs = (_polish_runs_batched if batched else _polish_runs_serial)(ac, gp, y_max, x_seeds, bounds, eps, maxfun)
    for res in runs:
        if not res.success:
            continue
        value = -float(np.squeeze(res.fun))
        if max_acq is None or value > max_acq:
            if evaluatedX is not None and not np.any(np.asarray(evaluatedX) == res.x):
                x_max, max_acq = res.x, value
    return x_max, max_acq


def _sweep(ac, util, gp, y_max, x_tries, n_iter):
    """One pass over the random candidates -> (x_max, max_acq, seeds for the polish)."""
    if n_iter == 0 and util is not None:
        max_acq, best = util.argmax(x_tries, gp, y_max, compute_embed=True)
        # the reference's x_seeds = x_tries[argsort(ys)[-0:]] is unused when n_iter == 0
        return x_tries[best], max_acq, x_tries[:0]
    ys = ac(x_tries, gp=gp, y_max=y_max, compute_embed=True)
    return x_tries[ys.argmax()], ys.max(), x_tries[np.argsort(ys)[-n_iter:]]


def acq_max(ac, gp, y_max, bounds, random_state, algorithm=None, kernel=None,
            size_evals=None, evaluatedX=None, prev_X=None, prev_Y=None, kappa=1.96,
            n_warmup=100000, n_iter=250):
    """Maximiser of the acquisition function: `n_warmup` uniform candidates, arg-max, optional
    L-BFGS-B polish from the best `n_iter` of them, all-zero-EI -> UCB fallback (helpers.py:9-111)."""
    x_tries = random_state.uniform(bounds[:, 0], bounds[:, 1], size=(n_warmup, bounds.shape[0]))
    if algorithm == 'BLR' and kernel in ('dist', 'manual', 'multi'):
        # BLR warm start (helpers.py:54-65): the candidates are the best theta of every source task but the last
        # (the random draw above still advances the caller's stream, as in the reference); no polish
        if kernel == 'multi':
            size_evals = size_evals[:-1]
        cumsum_size = [0] + np.cumsum(size_evals).tolist()
        init_loc = []
        for i in range(len(size_evals) - 1):
            init_Y_task = prev_Y[cumsum_size[i]:cumsum_size[i + 1]]
            init_loc.append(cumsum_size[i] + init_Y_task.argsort()[-1])
        x_tries = prev_X[init_loc]
        n_iter = 0
    util = _as_utility(ac)
    x_max, max_acq, x_seeds = _sweep(ac, util, gp, y_max, x_tries, n_iter)
    if n_iter != 0:
        x_max, max_acq = _polish(ac, gp, y_max, x_seeds, bounds, evaluatedX, x_max, max_acq, 0.05, 2000)
    if max_acq == 0.0:   # every EI value is exactly zero: fall back to UCB (helpers.py:87-107)
        print('All acquisition value 0.0, using UCB with {}'.format(kappa))
        util = UtilityFunction(kind='ucb', kappa=kappa, xi=0.01)
        ac = util.utility
        x_max, max_acq, x_seeds = _sweep(ac, util, gp, y_max, x_tries, n_iter)
        if n_iter != 0:
            x_max, max_acq = _polish(ac, gp, y_max, x_seeds, bounds, evaluatedX, x_max, max_acq, 0.1, 1500)
    gp.predict_y(x_max.reshape(1, -1))
    return np.clip(x_max, bounds[:, 0], bounds[:, 1])


def unique_rows(a):
    """Mask of the first occurrence of every distinct row (helpers.py:180-200)."""
    if a.size == 0:
        return np.empty((0,))
    order = np.lexsort(a.T)
    back = np.argsort(order)
    s = a[order]
    keep = np.ones(len(s), 'bool')
    keep[1:] = (np.diff(s, axis=0) != 0).any(axis=1)
    return keep[back]


def ensure_rng(random_state=None):
    """None -> fresh RandomState, int -> seeded RandomState, RandomState -> itself (helpers.py:203-216)."""
    if random_state is None:
        return np.random.RandomState()
    if isinstance(random_state, (int, np.integer)):
        return np.random.RandomState(int(random_state))
    assert isinstance(random_state, np.random.RandomState)
    return random_state


class PrintLog(object):
    """Plain-text progress table of the BO loop (helpers.py:228-315, colours dropped)."""

    def __init__(self, params):
        self.ymax = None
        self.xmax = None
        self.params = params
        self.ite = 1
        self.start_time = self.last_round = datetime.now()
        self.sizes = [max(len(ps), 7) for ps in params]
        self.sorti = sorted(range(len(params)), key=params.__getitem__)

    def reset_timer(self):
        self.start_time = self.last_round = datetime.now()

    def print_header(self, initialization=True):
        print("Initialization" if initialization else "Bayesian Optimization")
        print("-" * (29 + sum(s + 5 for s in self.sizes)))
        cells = ["{:>5}".format("Step"), "{:>6}".format("Time"), "{:>10}".format("Value")]
        cells += ["{0:>{1}}".format(self.params[i], self.sizes[i] + 2) for i in self.sorti]
        print(" | ".join(cells) + " | ")

    def print_step(self, x, y, warning=False):
        m, s = divmod((datetime.now() - self.last_round).total_seconds(), 60)
        if self.ymax is None or self.ymax < y:
            self.ymax, self.xmax = y, x
        cells = ["{:>5d}".format(self.ite), "{:>02d}m{:>02d}s".format(int(m), int(s)), "{: >10.5f}".format(y)]
        cells += ["{0: >{1}.{2}f}".format(x[i], self.sizes[i] + 2, min(self.sizes[i] - 3, 4)) for i in self.sorti]
        print(" | ".join(cells) + " | " + ("Warning: Test point chose at random due to repeated sample."
                                          if warning else ""))
        self.last_round = datetime.now()
        self.ite += 1

    def print_summary(self):
        pass
