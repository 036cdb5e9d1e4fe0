"""CPU tests of the host glue that surrounds the hot path: TargetSpace, acq_max decisions and the
BayesianOptimization loop, driven with the oracle's duck-typed regressor (no GPU needed).  The same
loop is run against the CUDA regressor in tests/test_gpu_trajectory.py."""
import numpy as np
import pytest

from distbo_b200.bayes_opt import BayesianOptimization, UtilityFunction, acq_max
from distbo_b200.bayes_opt.helpers import ensure_rng, unique_rows
from distbo_b200.bayes_opt.target_space import TargetSpace
from distbo_b200.data.toy_datasets import one_dim_split_tasks
from distbo_b200.model_embed import bayes_opt_wrap, set_model
from oracle.gp_regressor_oracle import OracleGPRegressor


def toy_optim_config(**over):
    cfg = dict(weight_dim=None, weight_dim_x=[20, 10], weight_dim_y=[20, 10], weight_dim_xy=None, n_warmup=300,
               n_opt_iterations=0, n_epochs=4, lkh_var_init=1e-2, lr=0.005, float64=True, n_cpus=1, stratify=False,
               embed_op='none', concat_data_size=False, algorithm='GP', gradient_clip=False, num_of_rff=0,
               batch_size=500, first_early_stop_epoch=2, max_size=None)
    cfg.update(over)
    return cfg


def test_random_points_draw_scalar_by_scalar():
    """The fork draws one scalar per coordinate, row-major (target_space.py:254-258); its docstring
    sample (:244-250) is the upstream column-wise order and does not match its own code."""
    space = TargetSpace(lambda p2, p1: p1 + p2, {'p2': (1, 100), 'p1': (0, 1)}, random_state=0)
    u = np.random.RandomState(0).uniform(size=6)
    want = np.array([[1 + 99 * u[0], u[1]], [1 + 99 * u[2], u[3]], [1 + 99 * u[4], u[5]]])
    assert np.allclose(space.random_points(3), want, rtol=1e-15)
    assert want[0, 0] == pytest.approx(55.33253689)      # first value of the reference docstring


def test_target_space_growth_and_duplicates():
    space = TargetSpace(lambda a, b: a - b, {'a': (0, 1), 'b': (0, 1)}, random_state=1)
    assert space.X is None and len(space) == 0
    for i in range(9):
        space.observe_point([0.5, 0.25])            # duplicates are stored again, as in the reference
    assert len(space) == 9 and space.X.shape == (9, 2) and np.all(space.Y == 0.25)
    assert [0.5, 0.25] in space
    assert space.max_point()['max_val'] == 0.25
    space._assert_internal_invariants()
    assert unique_rows(space.X).sum() == 1
    with pytest.raises(ValueError):
        space._dict_to_points({'a': [1, 2], 'b': [1]})


def test_ensure_rng():
    assert isinstance(ensure_rng(None), np.random.RandomState)
    assert ensure_rng(3).uniform() == np.random.RandomState(3).uniform()
    rs = np.random.RandomState(4)
    assert ensure_rng(rs) is rs


class _FakeGP:
    """predict_y only: mean peaks at x = 0.3, constant std."""

    def predict_y(self, x, **kw):
        x = np.asarray(x)
        return -np.sum((x - 0.3) ** 2, axis=1, keepdims=True), np.full((len(x), 1), 0.1)


def test_acq_max_argmax_and_fallback():
    gp = _FakeGP()
    bounds = np.array([[0.0, 1.0], [0.0, 1.0]])
    util = UtilityFunction('ucb', kappa=1.0, xi=0.0)
    x = acq_max(util.utility, gp, 0.0, bounds, np.random.RandomState(0), n_warmup=2000, n_iter=0)
    tries = np.random.RandomState(0).uniform(bounds[:, 0], bounds[:, 1], size=(2000, 2))
    assert np.array_equal(x, tries[np.argmax(util.utility(tries, gp, 0.0))])
    # EI identically zero (y_max far above the mean) -> UCB fallback picks the same point
    ei = UtilityFunction('ei', kappa=1.0, xi=0.01)
    assert np.max(ei.utility(tries, gp, 1e6)) == 0.0
    x2 = acq_max(ei.utility, gp, 1e6, bounds, np.random.RandomState(0), kappa=1.0, n_warmup=2000, n_iter=0)
    assert np.array_equal(x2, x)
    # polish: accepted only with evaluatedX (helpers.py:84); result stays inside the bounds
    x3 = acq_max(util.utility, gp, 0.0, bounds, np.random.RandomState(0), n_warmup=50, n_iter=3,
                 evaluatedX=np.array([[9.0, 9.0]]))
    assert np.all(x3 >= 0) and np.all(x3 <= 1) and np.sum((x3 - 0.3) ** 2) < 1e-3
    x4 = acq_max(util.utility, gp, 0.0, bounds, np.random.RandomState(0), n_warmup=50, n_iter=3)
    t50 = np.random.RandomState(0).uniform(bounds[:, 0], bounds[:, 1], size=(50, 2))
    assert np.array_equal(x4, t50[np.argmax(util.utility(t50, gp, 0.0))])
    with pytest.raises(NotImplementedError):
        UtilityFunction('es', 1.0, 0.0)


@pytest.mark.parametrize("method", ["params", "dist", "multi"])
def test_bo_loop_with_oracle_regressor(method):
    target, sources, supp = one_dim_split_tasks(0, 1, 1, n_train=40, n_test=40)
    rs = np.random.RandomState(5)
    init_list = []
    for s in sources:          # source evaluations: random search on each source objective
        f, bounds, _, _ = set_model('toy', s, None)
        th = rs.uniform(-8, 8, size=6)
        init_list.append(np.column_stack([th, [f(theta=t) for t in th]]))
    cfg = toy_optim_config()
    kw = dict(acq='ei', model='toy', rs=np.random.RandomState(11), optim_config=cfg, xi=0.01, kappa=2.58,
              bo_it=3, rand_it=2 if method == 'multi' else 0, init_list=init_list, include_init=False, source_data=sources,
              gp_class=OracleGPRegressor, verbose=0, acq_one_shot='lcb')
    out = bayes_opt_wrap(method, target, **kw)
    if method == 'dist':
        out, sim = out
        # one similarity row per acquisition: T_src = 2 entries, then 3 once the target task has joined
        assert [np.size(s) for s in sim] == [2, 3, 3]
        assert all(np.all(np.asarray(s) > 0) and np.all(np.asarray(s) <= 1 + 1e-12) for s in sim)
    if method == 'multi':
        out, sim = out
        # the target is the last of the T = 3 tasks from the start (bayesian_optimization.py:282-283)
        assert [np.size(s) for s in sim] == [3, 3, 3]
        assert all(np.asarray(s, dtype=float)[-1] == pytest.approx(1.0) for s in sim)
        out = out[-3:]                              # the 2 random target initialisations come first
    assert out.shape == (3, 2)
    f, _, _, _ = set_model('toy', target, None)
    assert np.allclose(out[:, 1], [f(theta=t) for t in out[:, 0]])
    assert np.all(np.abs(out[:, 0]) <= 8)


def test_bo_bookkeeping_matches_reference_rules():
    """size_evals grows by one task then by one observation per acquisition; the scaler is fitted
    once; first_early_stop_epoch drops to 600 from the second in-loop refit on."""
    target, sources, _ = one_dim_split_tasks(1, 1, 1, n_train=30, n_test=30)
    calls = []

    class Spy(OracleGPRegressor):
        def maximise_likhd(self, params, y, **kw):
            calls.append((len(params), list(np.ravel(kw['sizes'])), kw['first_early_stop_epoch'], len(kw['data_X'])))
            return super().maximise_likhd(params, y, **kw)

    rs = np.random.RandomState(2)
    init_list = [np.column_stack([rs.uniform(-8, 8, 5), rs.rand(5)]) for _ in sources]
    cfg = toy_optim_config(first_early_stop_epoch=1)
    bayes_opt_wrap('dist', target, acq='ei', model='toy', rs=3, optim_config=cfg, xi=0.01, kappa=2.58, bo_it=4,
                   rand_it=0, init_list=init_list, source_data=sources, gp_class=Spy, verbose=0)
    assert [c[0] for c in calls] == [10, 11, 12, 13]
    assert [c[1] for c in calls] == [[5, 5], [5, 5, 1], [5, 5, 2], [5, 5, 3]]
    assert [c[2] for c in calls] == [1, 1, 600, 600]
    assert [c[3] for c in calls] == [2, 3, 3, 3]


def test_blr_warm_start_candidates():
    """algorithm='BLR' with a transfer kernel (helpers.py:54-65): the one-shot acquisition is evaluated on the best
    theta of every source task except the last one, the uniform draw still advances the stream, no polish."""
    seen = {}

    class GP:
        def predict_y(self, x, compute_embed=False):
            seen.setdefault('x', []).append(np.array(x))
            return np.asarray(x)[:, :1] * 1.0, np.full((len(x), 1), 0.5)

    util = UtilityFunction('ucb', 1.0, 0.0)
    bounds = np.array([[0.0, 1.0], [0.0, 1.0]])
    prev_X = np.arange(24, dtype=float).reshape(12, 2) / 24.0
    prev_Y = np.array([0.1, 0.9, 0.3, 0.2, 0.5, 0.4, 0.8, 0.7, 0.0, 0.6, 0.95, 0.05])
    rs = np.random.RandomState(0)
    x = acq_max(util.utility, GP(), 0.0, bounds, rs, algorithm='BLR', kernel='dist', size_evals=[4, 5, 3],
                prev_X=prev_X, prev_Y=prev_Y, n_warmup=50, n_iter=3)
    np.testing.assert_array_equal(seen['x'][0], prev_X[[1, 4 + 2]])         # argmax of tasks 0 and 1 only
    assert np.array_equal(x, prev_X[6])                                    # larger first coordinate -> larger UCB
    assert rs.uniform() == np.random.RandomState(0).uniform(size=101)[-1]  # 50 x 2 draws were consumed
    seen.clear()
    acq_max(util.utility, GP(), 0.0, bounds, np.random.RandomState(0), algorithm='BLR', kernel='multi',
            size_evals=[4, 5, 3], prev_X=prev_X, prev_Y=prev_Y, n_warmup=50, n_iter=3)
    np.testing.assert_array_equal(seen['x'][0], prev_X[[1]])                # target dropped, then last source skipped


def test_bo_loop_blr_with_oracle_regressor():
    target, sources, supp = one_dim_split_tasks(0, 2, 1, n_train=40, n_test=40)
    rs = np.random.RandomState(5)
    init_list = []
    for s in sources:
        f, bounds, _, _ = set_model('toy', s, None)
        th = rs.uniform(-8, 8, size=6)
        init_list.append(np.column_stack([th, [f(theta=t) for t in th]]))
    cfg = toy_optim_config(algorithm='BLR', weight_dim=[12, 12, 12], embed_op='joint')
    out, sim = bayes_opt_wrap('dist', target, acq='ei', model='toy', rs=np.random.RandomState(11), optim_config=cfg,
                              xi=0.01, kappa=2.58, bo_it=3, rand_it=0, init_list=init_list, include_init=False,
                              source_data=sources, gp_class=OracleGPRegressor, verbose=0, acq_one_shot='lcb')
    assert out.shape == (3, 2) and all(s == 'none' for s in np.ravel(sim))
    # first suggestion = best theta of one of the first two source tasks (BLR warm start)
    firsts = [init[np.argmax(init[:, 1]), 0] for init in init_list[:2]]
    assert out[0, 0] in firsts


def test_c_shuffle_is_pythons_random_shuffle():
    """distbo_py_shuffle_i64 (host helper of the stratified sampler) = random.Random(seed).shuffle, call after call."""
    import random
    from distbo_b200.train.gp_utils import PyShuffle
    for seed in (0, 7, 2 ** 31 + 5):
        ref, mine = random.Random(seed), PyShuffle(seed)
        for n in (1, 2, 3, 17, 624, 625, 1000, 4434, 33, 0):
            lst = list(range(100, 100 + n))
            ref.shuffle(lst)
            got = mine.shuffle(np.arange(100, 100 + n))
            assert got.tolist() == lst, (seed, n)


def _reference_style_batches(Y_list, batch_size, epochs, pyrand):
    """loop_batches' class-stratified branch (reference train/gp_utils.py:38-84) on index lists, driven by `pyrand`."""
    from oracle.gp_regressor_oracle import loop_batches
    X_list = [np.arange(len(y), dtype=np.float64)[:, None] for y in Y_list]
    gen = loop_batches({"X": X_list, "y": Y_list}, batch_size, epochs, np.random.RandomState(0), "classification",
                       stratify=True, pyrand=pyrand)
    return [(bx[:, 0].astype(np.int64), es.reshape(-1).astype(int)) for bx, _, es in gen]


def test_class_stratified_sampler_many_classes_and_persistent_stream():
    """Six or more classes: the stratified batch has fewer than batch_size rows (gp_utils.py:57,72-83) and sizes()
    must say so; and the shuffle stream is ONE stream over all fits of a regressor (the reference never re-seeds
    the global `random`), rewound by restore() when mini-batches were drawn ahead of an early stop."""
    import random
    from distbo_b200.train.gp_utils import EmbedBatchSampler, PyShuffle
    rs = np.random.RandomState(3)
    n_classes, batch = 12, 40
    Y_list = [np.eye(n_classes)[rs.randint(0, n_classes, size=n)] for n in (150, 30, 97)]
    ref_rand = random.Random(5)
    stream = PyShuffle(5)
    for fit in range(3):                                   # three fits, one stream
        want = _reference_style_batches(Y_list, batch, 9, ref_rand)
        s = EmbedBatchSampler(Y_list, batch, np.random.RandomState(0), "classification", True, pyrand=stream)
        sizes = s.sizes()
        assert sizes[1] == 30 and sizes[0] < batch and sizes[2] < batch      # int() truncation per class: < 40 rows
        marks = []
        for ep in range(9 + 4):                            # the host runs four epochs ahead of the "stop"
            idx = s.next_indices()
            marks.append(stream.state())
            if ep < 9:
                rows = np.concatenate([np.arange(len(Y_list[i])) if ix is None else ix for i, ix in enumerate(idx)])
                off = np.r_[0, np.cumsum(want[ep][1])]
                got_sizes = [len(Y_list[i]) if ix is None else len(ix) for i, ix in enumerate(idx)]
                assert got_sizes == sizes == list(want[ep][1])
                # dataset-local row indices, dataset by dataset
                assert np.array_equal(rows, want[ep][0]), (fit, ep)
        stream.restore(marks[8])                           # back to the state after the ninth draw


def test_lockstep_polish_is_the_serial_polish():
    """The batched polish (all L-BFGS-B runs in lock-step, value + scipy's own forward-difference stencil per device
    call; reference helpers.py:70-86) returns, run for run, the serial result bit for bit - x, value, nfev, nit -
    including seeds on / next to a bound (flipped stencil steps), with ~25x fewer acquisition calls."""
    from distbo_b200.bayes_opt import helpers as H
    rs = np.random.RandomState(0)
    d = 4
    bounds = np.array([[-2.0, 3.0]] * d)
    C = rs.randn(5, d)
    calls = [0]

    def acq(X, gp=None, y_max=None):           # evaluated row by row: independent of the batch by construction
        calls[0] += 1
        X = np.atleast_2d(X)
        out = np.array([sum(np.exp(-np.sum((x - c) ** 2)) for c in C) + 0.1 * np.sin(3 * x).sum() for x in X])
        return out if len(out) > 1 else np.squeeze(out)
    seeds = rs.uniform(-2, 3, size=(10, d))
    seeds[3, 1], seeds[5, 0], seeds[7, 2] = 3.0, 2.98, -2.0
    serial = H._polish_runs_serial(acq, None, 0.0, seeds, bounds, 0.05, 2000)
    n_serial, calls[0] = calls[0], 0
    batched = H._polish_runs_batched(acq, None, 0.0, seeds, bounds, 0.05, 2000)
    for a, b in zip(serial, batched):
        assert np.array_equal(a.x, b.x) and a.fun == b.fun and a.nfev == b.nfev and a.nit == b.nit
        assert a.success == b.success
    assert calls[0] * 10 < n_serial
    # a tiny maxfun budget ends the runs on the same evaluation count as well
    a = H._polish_runs_serial(acq, None, 0.0, seeds[:3], bounds, 0.05, 12)
    b = H._polish_runs_batched(acq, None, 0.0, seeds[:3], bounds, 0.05, 12)
    assert [(r.nfev, r.success, r.fun) for r in a] == [(r.nfev, r.success, r.fun) for r in b]


def test_lockstep_evaluator_propagates_errors():
    from distbo_b200.bayes_opt import helpers as H

    def bad(X, gp=None, y_max=None):
        raise RuntimeError("device call failed")
    with pytest.raises(RuntimeError, match="device call failed"):
        H._polish_runs_batched(bad, None, 0.0, np.zeros((3, 2)), np.array([[-1.0, 1.0]] * 2), 0.05, 100)
