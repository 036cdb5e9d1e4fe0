"""GPU parity on the real-data-shaped configurations (BASELINE configs 2-4) from small fixtures of the
reference's own data (tests/golden/protein_subset.npz, parkinson_subset.npz; tools/make_real_fixtures.py):
the same BO loop once on the CUDA surrogate and once on the CPU oracle must suggest IDENTICAL points."""
import os

import numpy as np
import pytest
from sklearn.utils import check_random_state

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from distbo_b200.data.load_datasets import parkinson_tasks, protein_tasks  # noqa: E402
from distbo_b200.model_embed import bayes_opt_wrap, set_model  # noqa: E402
from distbo_b200.model_embed.models import register_model  # noqa: E402
from oracle.gp_regressor_oracle import OracleGPRegressor  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def config(**over):
    cfg = dict(weight_dim=None, weight_dim_x=[20, 10], weight_dim_y=[20, 10], weight_dim_xy=None, n_warmup=20000,
               n_opt_iterations=0, n_epochs=20, lkh_var_init=1e-5, lr=0.005, float64=True, n_cpus=1, stratify=True,
               embed_op='joint', concat_data_size=True, algorithm='GP', gradient_clip=False, num_of_rff=0,
               batch_size=1000, first_early_stop_epoch=10, max_size=None)
    cfg.update(over)
    return cfg


def seeded_forest(data, depth, min_samples_leaf, min_samples_split, n_trees):
    from sklearn.ensemble import RandomForestClassifier
    from sklearn.metrics import roc_auc_score
    f = RandomForestClassifier(n_estimators=max(1, int(n_trees) // 8), max_depth=int(depth),
                               min_samples_split=min_samples_split, min_samples_leaf=min_samples_leaf, random_state=0)
    f.fit(data.train_x, data.train_y[:, 0])
    return roc_auc_score(data.test_y[:, 0], f.predict_proba(data.test_x)[:, 1])


register_model('forest_seeded', seeded_forest, 'classification',
               {'n_trees': (1, 200), 'depth': (1, 32), 'min_samples_leaf': (0.01, 0.5),
                'min_samples_split': (0.01, 1.0)},
               ['depth', 'min_samples_leaf', 'min_samples_split', 'n_trees'])


def random_search(sources, model, n, seed):
    rs = np.random.RandomState(seed)
    out = []
    for s in sources:
        f, bounds, _, _ = set_model(model, s, None)
        names = sorted(bounds)
        rows = []
        for _ in range(n):
            th = {k: rs.uniform(*bounds[k]) for k in names}
            rows.append([th[k] for k in names] + [f(**th)])
        out.append(np.array(rows))
    return out


def both(method, target, sources, model, init_list, cfg, bo_it, acq_one_shot='lcb'):
    runs = []
    for gp_class in (None, OracleGPRegressor):
        runs.append(bayes_opt_wrap(method, target, acq='ei', model=model, rs=np.random.RandomState(11),
                                   optim_config=dict(cfg), xi=0.01, kappa=2.58, bo_it=bo_it, rand_it=0,
                                   init_list=[a.copy() for a in init_list], include_init=False, source_data=sources,
                                   gp_class=gp_class, verbose=0, acq_one_shot=acq_one_shot))
    return runs


def assert_same_trajectory(a, b):
    if isinstance(a, tuple):
        (a, sa), (b, sb) = a, b
        for x, y in zip(sa, sb):
            assert np.allclose(np.asarray(x, dtype=float), np.asarray(y, dtype=float), rtol=1e-7, atol=0)
    assert a.shape == b.shape
    assert np.array_equal(a[:, :-1], b[:, :-1]), (a[:, :-1], b[:, :-1])     # identical suggested points
    assert np.array_equal(a[:, -1], b[:, -1])


def protein_fixture():
    z = np.load(os.path.join(GOLD, "protein_subset.npz"))
    bits = np.unpackbits(z["bits"], axis=2)[:, :, :166].astype(np.float64)
    tables = [np.hstack([np.zeros((bits.shape[1], 1)), bits[i], z["labels"][i][:, None].astype(np.float64)])
              for i in range(bits.shape[0])]
    return tables, [str(n) for n in z["names"]]


def test_protein_shaped_trajectory(cuda_engine_mod):
    """Config 2: 6 source ChEMBL targets + 1 target, 166 binary features, 2 classes, P = 23, 4-d theta;
    stratified class mini-batches (batch 100 of 216 rows)."""
    tables, names = protein_fixture()
    target, sources, _ = protein_tasks(tables, names, target=2, random_state=check_random_state(3), test_size=0.4)
    init_list = random_search(sources, 'forest_seeded', 6, seed=5)
    max_size = max([target.train_len()] + [s.train_len() for s in sources])
    cfg = config(batch_size=100, max_size=max_size)
    a, b = both('dist', target, sources, 'forest_seeded', init_list, cfg, bo_it=4)
    assert a[0].shape == (4, 5)
    assert_same_trajectory(a, b)


def test_parkinson_shaped_trajectory(cuda_engine_mod):
    """Config 3: per-patient regression datasets (17 features), joint embedding P = 100, ridge objective."""
    z = np.load(os.path.join(GOLD, "parkinson_subset.npz"))
    patients = [z["patient_%d" % i] for i in range(10)]
    target, sources, _ = parkinson_tasks(patients, target=4, random_state=check_random_state(1), test_size=0.4)
    init_list = random_search(sources, 'ridge', 6, seed=6)
    cfg = config(concat_data_size=False, stratify=False, batch_size=1000)
    a, b = both('dist', target, sources, 'ridge', init_list, cfg, bo_it=4, acq_one_shot='ei')
    assert a[0].shape == (4, 3)
    assert_same_trajectory(a, b)


def test_manual_meta_feature_trajectory(cuda_engine_mod):
    """Config 4 shape: 7-d theta, manual kernel over 26 MinMax-scaled meta-features per dataset."""
    def bowl(data, **th):
        v = np.array([th[k] for k in sorted(th)])
        return float(np.exp(-0.5 * np.sum((v - data.centre) ** 2)))
    register_model('bowl7', bowl, 'regression', {'t%d' % i: (-2.0, 2.0) for i in range(7)})
    rs = np.random.RandomState(9)

    class D(object):
        pass
    tasks = []
    for t in range(5):
        d = D()
        d.centre = rs.uniform(-1, 1, size=7)
        d.embed = np.clip(0.5 + 0.3 * d.centre.sum() * np.linspace(-1, 1, 26) + 0.05 * rs.randn(26), 0, 1)
        d.train_x = d.train_y = np.zeros((3, 1))
        d.embed_ind = range(3)
        tasks.append(d)
    target, sources = tasks[0], tasks[1:]
    init_list = random_search(sources, 'bowl7', 20, seed=7)
    cfg = config(n_epochs=25, concat_data_size=False)
    a, b = both('manual', target, sources, 'bowl7', init_list, cfg, bo_it=4)
    assert a[0].shape == (4, 8)
    assert_same_trajectory(a, b)
