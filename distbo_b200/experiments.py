"""Experiment driver: one transfer-BO run end to end, written to `results.npz` with the reference's keys
(reference experiments/train_test.py:167-338; flags :28-165).

    python -m distbo_b200.experiments one_dim_split --data-seed 0 --opt-seed 0 --float64 ... out_dir
    python -m distbo_b200.experiments protein_vary --data-dir /path/to/protein ... out_dir

Same sub-commands, flags, defaults and control flow as the reference's CLI for the four data sets of the BASELINE
configurations (`one_dim_split`, `counter_manual`, `patient_vary`, `protein_vary`); the surrogate behind it is the
sm_100a `gp_regressor`.  Host glue only: data generation, the source searches, the target search, `np.savez`.
`run(args)` is the importable form; `surrogate=` swaps the regressor class (the parity tools pass the CPU oracle's)
and `source_evals=` replays recorded source searches instead of running them (the reference's --use-prev-results).
"""
from __future__ import annotations

import argparse
import multiprocessing as mp
import os
import time
from copy import deepcopy

import numpy as np
from sklearn.utils import check_random_state

from .data.load_datasets import load_parkinson, load_protein, parkinson_tasks, protein_tasks
from .data.toy_datasets import counter_manual_tasks, one_dim_split_tasks
from .model_embed.bayes_optimise import bayes_opt_wrap
from .model_embed.meta_fea import meta

_COMMON = ["--float64", "--kappa", "2.58", "--xi", "0.01", "--learning-rate", "0.005", "--n-warmup", "300000",
           "--num-of-rff", "0", "--warm-start-method", "manual", "--warm-start-tasks", "3", "--n-warm-per-task", "3",
           "--weight_dim_1", "0", "--weight_dim_2", "0", "--weight_dim_3", "0",
           "--weight_dim_1_xy", "0", "--weight_dim_2_xy", "0", "--weight_dim_3_xy", "0"]

# The command lines the reference ships for the BASELINE configurations (seed 0, GP surrogate): what
# experiments/one_dim_toy_config.py, protein_config.py and counter_manual_config.py print for their first grid
# entry, and the CLI defaults for Parkinson (no config script; train_test.py:28-91) with --float64.
SHIPPED_COMMAND_LINES = {
    "config1": ["one_dim_split", "--n-train", "500", "--n-test", "500", "--data-seed", "0", "--opt-seed", "0",
                "--bo-source-type", "BO", "--source-rand-iterations", "10", "--source-iterations", "20",
                "--algorithm", "GP", "--bo-target-type", "distBO", "--target-iterations", "15",
                "--target-rand-iterations", "0", "--embed-op", "none", "--max-embed-size", "500",
                "--batch-size", "500", "--dont-stratify", "--n-epochs", "10000", "--n-opt-iterations", "0",
                "--first_early_stop_epoch", "5000", "--weight_dim_1_x", "20", "--weight_dim_2_x", "10",
                "--weight_dim_1_y", "0", "--weight_dim_2_y", "0"] + _COMMON,
    "config2": ["protein_vary", "--target-group", "0", "--test-size", "0.4", "--data-seed", "0", "--opt-seed", "0",
                "--bo-source-type", "BO", "--source-rand-iterations", "10", "--source-iterations", "10",
                "--concat-data-size", "--algorithm", "GP", "--bo-target-type", "distBO",
                "--target-iterations", "20", "--target-rand-iterations", "0", "--embed-op", "joint",
                "--max-embed-size", "5000", "--batch-size", "1000", "--n-epochs", "10000",
                "--n-opt-iterations", "10", "--first_early_stop_epoch", "5000", "--weight_dim_1_x", "20",
                "--weight_dim_2_x", "10", "--weight_dim_1_y", "20", "--weight_dim_2_y", "10"] + _COMMON,
    "config3": ["patient_vary", "--data-seed", "0", "--opt-seed", "0", "--float64"],
    "config4": ["counter_manual", "--n-train", "5000", "--n-test", "5000", "--data-seed", "0", "--opt-seed", "0",
                "--bo-source-type", "BO", "--source-rand-iterations", "75", "--source-iterations", "75",
                "--algorithm", "GP", "--bo-target-type", "manualBO", "--target-iterations", "100",
                "--target-rand-iterations", "0", "--embed-op", "none", "--max-embed-size", "5000",
                "--batch-size", "0", "--n-epochs", "10000", "--n-opt-iterations", "10",
                "--first_early_stop_epoch", "5000", "--weight_dim_1_x", "0", "--weight_dim_2_x", "0",
                "--weight_dim_1_y", "0", "--weight_dim_2_y", "0"] + _COMMON,
}

RESULT_KEYS = ("args", "supp", "source_params", "source_evals", "source_time", "target_sim", "target_params",
               "target_evals", "target_time")         # train_test.py:170-338


def _add_args(sub):
    """train_test.py:28-91, flag for flag."""
    a = sub.add_argument
    a('--use-prev-results', action='store_true', default=False)
    a('--bo-source-type', choices=['BO', 'RS'], default='RS')
    a('--bo-target-type', choices=['distBO', 'allBO', 'manualBO', 'noneBO', 'RS', 'multiBO', 'initBO'],
      default='distBO')
    a('--acq-one-shot', choices=['lcb', 'ei', 'ucb', 'es'], default='lcb')
    a('--acq-source', choices=['ei', 'ucb', 'es'], default='ei')
    a('--acq-target', choices=['ei', 'ucb', 'es'], default='ei')
    a('--source-rand-iterations', type=int, default=30)
    a('--source-iterations', type=int, default=0)
    a('--target-rand-iterations', type=int, default=0)
    a('--target-iterations', type=int, default=30)
    a('--warm-start-tasks', type=int, default=3)
    a('--n-warm-per-task', type=int, default=3)
    a('--warm-start-method', choices=['dist', 'manual'], default='manual')
    a('--data-seed', type=int, default=None)
    a('--opt-seed', type=int, default=None)
    a('--embed-op', choices=['none', 'joint', 'concat', 'nn'], default='joint')
    a('--concat-data-size', action='store_true', default=False)
    a('--max-embed-size', type=int, default=20000)
    a('--xi', type=float, default=0.01)
    a('--kappa', type=float, default=2.58)
    a('--algorithm', choices=['GP', 'BLR'], default='GP')
    for i in (1, 2, 3):
        a('--weight_dim_%d' % i, type=int, default=50)
    a('--num-of-rff', type=int, default=0)
    a('--weight_dim_1_x', type=int, default=20)
    a('--weight_dim_2_x', type=int, default=10)
    a('--weight_dim_1_y', type=int, default=20)
    a('--weight_dim_2_y', type=int, default=10)
    a('--weight_dim_1_xy', type=int, default=20)
    a('--weight_dim_2_xy', type=int, default=20)
    a('--weight_dim_3_xy', type=int, default=0)
    a('--gradient-clip', action='store_true', default=False)
    a('--lkh_var_init', type=float, default=0.00001)
    a('--dont-stratify', action='store_true', default=False)
    a('--n-warmup', type=int, default=300000)
    a('--n-opt-iterations', type=int, default=10)
    a('--n-epochs', type=int, default=5000)
    a('--learning-rate', type=float, default=0.005)
    a('--first_early_stop_epoch', type=int, default=1000)
    a('--batch-size', type=int, default=1000)
    a('--float64', action='store_true', default=False)
    a('out_dir')
    a('--n-cpus', type=int, default=min(1, mp.cpu_count()))
    a('--data-dir', default=None, help="folder holding the reference's protein/ or parkinson/ files")
    a('--quiet', action='store_true', default=False)


def make_parser():
    """train_test.py:93-150 (the `bw_dim_vary` toy family is not part of any BASELINE configuration)."""
    parser = argparse.ArgumentParser(prog="distbo_b200.experiments")
    subs = parser.add_subparsers(help="The dataset to run on")

    def add(name):
        sub = subs.add_parser(name)
        sub.set_defaults(dataset=name)
        _add_args(sub)
        return sub

    def sim_args(sub):
        sub.add_argument('--preprocess', choices=['None', 'standardise'], default='standardise')
        sub.add_argument('--n-train', type=int, default=1000)
        sub.add_argument('--n-test', type=int, default=1000)
        sub.add_argument('--dim', '-d', type=int, default=10)

    def split_args(sub):
        sub.add_argument('--preprocess', choices=['None', 'standardise'], default='standardise')
        sub.add_argument('--test-size', type=float, default=.4)

    s = add('one_dim_split')
    s.add_argument('--fake_source_num', type=int, default=12)
    s.add_argument('--true_source_num', type=int, default=3)
    s.add_argument('--mu_sd', type=float, default=0.5)
    sim_args(s)
    s = add('counter_manual')
    s.add_argument('--noise', type=float, default=0.5)
    sim_args(s)
    s = add('patient_vary')
    s.add_argument('--target-group', type=int, default=36)
    s.add_argument('--park-labels', choices=['motor', 'total', 'both'], default='total')
    split_args(s)
    s = add('protein_vary')
    s.add_argument('--target-group', type=int, default=0)
    split_args(s)
    return parser


def parse_args(argv=None):
    args = make_parser().parse_args(argv)
    if args.data_seed is None:
        args.data_seed = int(np.random.randint(2 ** 32))
    if args.opt_seed is None:
        args.opt_seed = int(np.random.randint(2 ** 23))
    return args


def generate_data(args):
    """distBO/data/generate_data.py:11-118 for the BASELINE data sets; sets args.model as the reference does."""
    rs = check_random_state(args.data_seed)
    if args.dataset == 'patient_vary':
        args.model = 'ridge'
        if args.park_labels == 'both':
            raise NotImplementedError("--park-labels both (84 tasks) is not part of any BASELINE configuration")
        src = _data_dir(args, 'parkinson')
        if src.endswith('.npz'):      # the 42 patient tables packed into one file (tests/golden/parkinson_full.npz)
            z = np.load(src)
            return parkinson_tasks([z["patient_%d" % i] for i in range(42)], args.target_group, random_state=rs,
                                   preprocess=args.preprocess, label=args.park_labels, test_size=args.test_size)
        return load_parkinson(args.target_group, src, random_state=rs,
                              preprocess=args.preprocess, label=args.park_labels, test_size=args.test_size)
    if args.dataset == 'protein_vary':
        args.model = getattr(args, "protein_model", 'forest')     # generate_data.py:22 (or 'jaccard_svm')
        src = _data_dir(args, 'protein')
        if src.endswith('.npz'):      # the 7 ChEMBL tables, fingerprints bit-packed (tests/golden/protein_full.npz)
            z = np.load(src)
            tables = []
            for i in range(len(z["names"])):
                bits = np.unpackbits(z["bits_%d" % i], axis=1)[:, :166].astype(np.float64)
                tables.append(np.hstack([np.zeros((len(bits), 1)), bits, z["labels_%d" % i].astype(np.float64)[:, None]]))
            return protein_tasks(tables, [str(n) for n in z["names"]], args.target_group, random_state=rs,
                                 test_size=args.test_size)
        return load_protein(args.target_group, src, random_state=rs,
                            preprocess=args.preprocess, test_size=args.test_size)
    if args.dataset == 'one_dim_split':
        args.model = 'toy'
        return one_dim_split_tasks(rs, args.fake_source_num, args.true_source_num, mu_sd=args.mu_sd,
                                   n_train=args.n_train, n_test=args.n_test, max_embed_size=args.max_embed_size)
    if args.dataset == 'counter_manual':
        args.dim, args.model = 6, 'dim_ridge'
        return counter_manual_tasks(rs, noise=args.noise, n_train=args.n_train, n_test=args.n_test,
                                    max_embed_size=args.max_embed_size)
    raise ValueError("unknown dataset {}".format(args.dataset))


def _data_dir(args, name):
    d = args.data_dir or os.environ.get("DISTBO_DATA_DIR")
    if d is None:
        raise ValueError("--data-dir (or DISTBO_DATA_DIR) must point at the folder that holds the reference's "
                         "%s/ files" % name)
    if os.path.isfile(d):
        return d
    return d if os.path.basename(os.path.normpath(d)) == name or not os.path.isdir(os.path.join(d, name)) \
        else os.path.join(d, name)


def optim_config_from(args):
    """train_test.py:189-218."""
    if args.weight_dim_1 != 0 and args.weight_dim_2 != 0:
        weight_dim = [args.weight_dim_1, args.weight_dim_2] + ([args.weight_dim_3] if args.weight_dim_3 != 0 else [])
    else:
        weight_dim = []
    weight_dim_xy = [args.weight_dim_1_xy, args.weight_dim_2_xy] + \
        ([args.weight_dim_3_xy] if args.weight_dim_3_xy != 0 else [])
    return {'weight_dim': weight_dim, 'weight_dim_x': [args.weight_dim_1_x, args.weight_dim_2_x],
            'weight_dim_y': [args.weight_dim_1_y, args.weight_dim_2_y], 'weight_dim_xy': weight_dim_xy,
            'n_warmup': args.n_warmup, 'n_opt_iterations': args.n_opt_iterations, 'n_epochs': args.n_epochs,
            'lkh_var_init': args.lkh_var_init, 'lr': args.learning_rate, 'float64': args.float64,
            'n_cpus': args.n_cpus, 'stratify': not args.dont_stratify, 'embed_op': args.embed_op,
            'concat_data_size': args.concat_data_size, 'algorithm': args.algorithm,
            'gradient_clip': args.gradient_clip, 'num_of_rff': args.num_of_rff, 'batch_size': args.batch_size,
            'first_early_stop_epoch': args.first_early_stop_epoch, 'max_size': None}


def run(args, surrogate=None, source_evals=None, data=None, on_target_start=None, verbose=1):
    """One experiment (train_test.py:167-338) -> the dict the reference saves.  `surrogate`: regressor class or
    factory (default: the sm_100a gp_regressor); `source_evals`: list of [n_i, d+1] arrays replayed instead of the
    source searches; `data`: (target, sources, supp) instead of generate_data(args); `on_target_start()`: hook
    called right before the target search (the parity replays restore recorded random states there)."""
    d = {'args': args}
    target_data, source_datas, supp = data if data is not None else generate_data(args)
    d['supp'] = supp
    args.include_corr_meta = args.dataset != 'protein_vary'
    opt_rs = check_random_state(args.opt_seed)
    if args.bo_target_type not in ['noneBO', 'RS', 'multiBO']:
        args.target_rand_iterations = 0
    bo_config = {'model': args.model, 'rs': opt_rs, 'xi': args.xi, 'kappa': args.kappa}
    optim_config = optim_config_from(args)
    source_optim_config = deepcopy(optim_config)
    extra = {'gp_class': surrogate, 'verbose': verbose}
    max_size = None

    if args.bo_target_type in ['distBO', 'allBO', 'multiBO', 'manualBO', 'initBO']:
        if args.concat_data_size:
            optim_config['max_size'] = max_size = np.max([target_data.train_len()] +
                                                         [s.train_len() for s in source_datas])
        if args.bo_source_type == 'RS':
            args.source_iterations = 0
        source_bo_config = {'model': args.model, 'rs': opt_rs, 'xi': 0.01, 'kappa': args.kappa,
                            'acq': args.acq_source, 'bo_it': args.source_iterations,
                            'rand_it': args.source_rand_iterations}
        source_optim_config['algorithm'] = 'GP'
        source_optim_config['first_early_stop_epoch'] = 1000
        source_bo_config['optim_config'] = source_optim_config
        start = time.time()
        if source_evals is None:
            source_evals = [bayes_opt_wrap('params', s_data, include_init=True, **source_bo_config, **extra)
                            for s_data in source_datas]
        d['source_params'] = [sp[:, :-1] for sp in source_evals]
        d['source_evals'] = [sp[:, -1] for sp in source_evals]
        d['source_time'] = time.time() - start

    start = time.time()
    bo_config['rs'] = check_random_state(args.opt_seed + 1)
    tcfg = dict(bo_config, acq=args.acq_target, bo_it=args.target_iterations, rand_it=args.target_rand_iterations,
                optim_config=optim_config, **extra)
    kind = args.bo_target_type
    if on_target_start is not None:
        on_target_start()
    if kind == 'distBO':
        summary, d['target_sim'] = bayes_opt_wrap('dist', target_data, init_list=source_evals, include_init=False,
                                                  source_data=source_datas, acq_one_shot=args.acq_one_shot, **tcfg)
    elif kind == 'manualBO':
        meta(source_datas + [target_data], args.model, max_size=max_size, random_state=opt_rs,
             include_corr_meta=args.include_corr_meta)
        summary, d['target_sim'] = bayes_opt_wrap('manual', target_data, init_list=source_evals, include_init=False,
                                                  source_data=source_datas, acq_one_shot=args.acq_one_shot, **tcfg)
    elif kind == 'initBO':
        args.target_rand_iterations = 0
        assert args.warm_start_tasks <= len(source_datas), \
            'Warm start tasks should not be greater than the number of source data.'
        if args.warm_start_method == 'manual':
            meta(source_datas + [target_data], args.model, max_size=max_size, random_state=opt_rs,
                 include_corr_meta=args.include_corr_meta)
        summary = bayes_opt_wrap('params', target_data, init_list=source_evals, include_init=False,
                                 warm_start_tasks=args.warm_start_tasks, n_warm_per_task=args.n_warm_per_task,
                                 warm_start_method=args.warm_start_method, source_data=source_datas, **tcfg)
    elif kind == 'allBO':
        summary = bayes_opt_wrap('params', target_data, init_list=source_evals, source_data=source_datas,
                                 include_init=False, acq_one_shot=args.acq_one_shot, **tcfg)
    elif kind == 'multiBO':
        summary, d['target_sim'] = bayes_opt_wrap('multi', target_data, init_list=source_evals,
                                                  source_data=source_datas, include_init=False, **tcfg)
    elif kind == 'noneBO':
        summary = bayes_opt_wrap('params', target_data, include_init=True, **tcfg)
    else:   # 'RS'
        args.target_iterations = tcfg['bo_it'] = 0
        summary = bayes_opt_wrap('params', target_data, include_init=True, **tcfg)
    if kind == 'initBO':
        total_it = args.warm_start_tasks * args.n_warm_per_task + args.target_iterations
    else:
        total_it = args.target_rand_iterations + args.target_iterations
    d['target_params'] = summary[-total_it:, :-1]
    d['target_evals'] = summary[-total_it:, -1]
    d['target_time'] = time.time() - start
    return d


def save_results(d, out_dir):
    """np.savez(out_dir/results.npz, **d) as train_test.py:338 (ragged lists stored as object arrays)."""
    os.makedirs(out_dir, exist_ok=True)
    out = {}
    for k, v in d.items():
        if k in ('source_params', 'source_evals'):
            arr = np.empty(len(v), dtype=object)
            for i, a in enumerate(v):
                arr[i] = a
            out[k] = arr
        else:
            out[k] = v
    path = os.path.join(out_dir, 'results.npz')
    np.savez(path, **out)
    return path


def main(argv=None):
    args = parse_args(argv)
    if os.path.exists(args.out_dir) and set(os.listdir(args.out_dir)) - {'output'}:
        raise SystemExit("Output directory {} exists. Change the name or delete it.".format(args.out_dir))
    d = run(args, verbose=0 if args.quiet else 1)
    path = save_results(d, args.out_dir)
    print("results written to", path)
    return 0


if __name__ == '__main__':
    raise SystemExit(main())
