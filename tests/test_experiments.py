"""The experiment driver (distbo_b200/experiments.py; reference experiments/train_test.py:28-338) on CPU: flag
surface, results.npz keys, data generation and objectives against the reference's own full-horizon run
(tests/golden/reffull_toy.npz, written by the reference's main()), and a short end-to-end run with the CPU oracle's
regressor standing in for the CUDA surrogate."""
import json
import os

import numpy as np
import pytest

from distbo_b200 import experiments as X
from oracle.gp_regressor_oracle import OracleGPRegressor

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_flags_and_defaults_are_the_references():
    a = X.parse_args(["patient_vary", "out"])
    # experiments/train_test.py:28-91 defaults (config 3 runs on them)
    assert (a.source_rand_iterations, a.source_iterations, a.target_iterations) == (30, 0, 30)
    assert (a.n_warmup, a.n_opt_iterations, a.n_epochs, a.first_early_stop_epoch) == (300000, 10, 5000, 1000)
    assert (a.lkh_var_init, a.learning_rate, a.batch_size, a.float64) == (1e-5, 0.005, 1000, False)
    assert (a.acq_one_shot, a.acq_target, a.bo_target_type, a.bo_source_type) == ('lcb', 'ei', 'distBO', 'RS')
    assert (a.target_group, a.park_labels, a.test_size) == (36, 'total', 0.4)
    oc = X.optim_config_from(a)
    assert set(oc) == {'weight_dim', 'weight_dim_x', 'weight_dim_y', 'weight_dim_xy', 'n_warmup', 'n_opt_iterations',
                       'n_epochs', 'lkh_var_init', 'lr', 'float64', 'n_cpus', 'stratify', 'embed_op',
                       'concat_data_size', 'algorithm', 'gradient_clip', 'num_of_rff', 'batch_size',
                       'first_early_stop_epoch', 'max_size'}
    assert oc['weight_dim'] == [50, 50, 50] and oc['weight_dim_xy'] == [20, 20] and oc['stratify'] is True


def test_reference_command_line_parses_and_data_matches_reference_run():
    """The exact command line of experiments/one_dim_toy_config.py (stored in the golden) parses; the toy tasks it
    generates and the toy objective reproduce the objective values the reference's run recorded, bit for bit."""
    z = np.load(os.path.join(GOLD, "reffull_toy.npz"))
    args = X.parse_args(json.loads(str(z["argv"])) + ["out"])
    assert (args.n_epochs, args.first_early_stop_epoch, args.n_warmup, args.target_iterations) == (10000, 5000, 300000, 15)
    target, sources, supp = X.generate_data(args)
    assert args.model == 'toy' and len(sources) == 15 and len(target.train_x) == 500
    from distbo_b200.model_embed.models import set_model
    f, bounds, _, _ = set_model('toy', target, None)
    assert np.array_equal([f(theta=t) for t in z["target_params"][:, 0]], z["target_evals"])
    for s, sp, se in zip(sources, z["source_params"], z["source_evals"]):
        fs = set_model('toy', s, None)[0]
        assert np.array_equal([fs(theta=t) for t in sp[:, 0]], se)
    # results.npz of the reference holds exactly these keys (train_test.py:170-338)
    assert set(str(k) for k in z["results_keys"]) == set(X.RESULT_KEYS)


def test_driver_runs_end_to_end_and_writes_results_npz(tmp_path):
    argv = ["one_dim_split", "--n-train", "60", "--n-test", "60", "--fake_source_num", "2", "--true_source_num", "1",
            "--data-seed", "3", "--opt-seed", "4", "--float64", "--bo-source-type", "BO",
            "--source-rand-iterations", "3", "--source-iterations", "2", "--target-iterations", "3",
            "--embed-op", "none", "--batch-size", "60", "--dont-stratify", "--n-epochs", "12", "--n-warmup", "2000",
            "--n-opt-iterations", "0", "--first_early_stop_epoch", "5", "--lkh_var_init", "0.01",
            str(tmp_path / "run")]
    args = X.parse_args(argv)
    d = X.run(args, surrogate=OracleGPRegressor, verbose=0)
    assert set(d) == set(X.RESULT_KEYS)
    assert len(d['source_params']) == 3 and all(p.shape == (5, 1) for p in d['source_params'])
    assert d['target_params'].shape == (3, 1) and d['target_evals'].shape == (3,)
    assert len(d['target_sim']) == 3
    path = X.save_results(d, args.out_dir)
    z = np.load(path, allow_pickle=True)
    assert set(z.files) == set(X.RESULT_KEYS)
    assert np.array_equal(z['target_params'], d['target_params'])
    # same seeds -> same run
    d2 = X.run(X.parse_args(argv), surrogate=OracleGPRegressor, verbose=0)
    assert np.array_equal(d2['target_params'], d['target_params'])
    # replaying recorded source searches skips them and leaves the target search unchanged
    src = [np.hstack([p, e[:, None]]) for p, e in zip(d['source_params'], d['source_evals'])]
    d3 = X.run(X.parse_args(argv), surrogate=OracleGPRegressor, source_evals=src, verbose=0)
    assert np.array_equal(d3['target_params'], d['target_params'])


@pytest.mark.parametrize("kind", ["manualBO", "multiBO", "initBO", "allBO", "noneBO", "RS"])
def test_driver_baseline_methods(kind, tmp_path):
    """The baselines the reference's CLI offers beside distBO (train_test.py:283-325)."""
    argv = ["counter_manual", "--n-train", "80", "--n-test", "80", "--data-seed", "1", "--opt-seed", "2", "--float64",
            "--source-rand-iterations", "4", "--target-iterations", "2", "--target-rand-iterations", "2",
            "--bo-target-type", kind, "--n-epochs", "6", "--n-warmup", "500", "--n-opt-iterations", "0",
            "--first_early_stop_epoch", "3", "--lkh_var_init", "0.01", "--warm-start-tasks", "2",
            "--n-warm-per-task", "1", str(tmp_path / kind)]
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        d = X.run(X.parse_args(argv), surrogate=OracleGPRegressor, verbose=0)
    total = {"initBO": 2 * 1 + 2, "noneBO": 4, "multiBO": 4, "RS": 2}.get(kind, 2)
    assert d['target_params'].shape == (total, 7) and np.all(np.isfinite(d['target_evals']))
    assert ('target_sim' in d) == (kind in ("manualBO", "multiBO"))
