"""One hyper-parameter search (reference distBO/model_embed/bayes_optimise.py:11-63)."""
from __future__ import annotations

from copy import deepcopy

import numpy as np

from ..bayes_opt import BayesianOptimization
from .models import set_model


def similarity_l2(source_data, target):
    """Negative L2 distance between manual meta-feature vectors (distBO/utils.py:21-24)."""
    return [-np.linalg.norm(target.embed - s.embed, ord=2) for s in source_data]


def bayes_opt_wrap(method, target_data, acq, model, rs, optim_config, xi=0, kappa=1.0, bo_it=100,
                   warm_start_tasks=0, n_warm_per_task=1, warm_start_method='manual', rand_it=10,
                   init_list=None, max_size=None, acq_one_shot='ei', include_init=False, source_data=None,
                   gp_class=None, verbose=1):
    """Runs BayesianOptimization.maximize with kernel `method` on `target_data`.

    Returns [params sorted by name..., value] rows (and the similarity log for 'dist' / 'manual' / 'multi').
    `init_list` = per-source arrays of previous evaluations (last column = value)."""
    size_evals, stacked = None, None
    if init_list is not None:
        size_evals = [len(init) for init in init_list]
        stacked = np.vstack(init_list)
    target, bounds, previous, model_type = set_model(model, deepcopy(target_data), stacked)
    if source_data is not None:
        bo = BayesianOptimization(target, bounds, source_data=source_data, target_data=target_data,
                                  size_evals=deepcopy(size_evals), include_init=include_init,
                                  model_type=model_type, random_state=rs, verbose=verbose)
    else:
        bo = BayesianOptimization(target, bounds, include_init=include_init, model_type=model_type,
                                  random_state=rs, verbose=verbose)
    if warm_start_tasks != 0 and source_data is not None:
        assert n_warm_per_task > 0, 'n_warm_per_task must be greater than 0'
        similarity = np.array(similarity_l2(source_data, target_data))
        nearest = similarity.argsort()[-warm_start_tasks:][::-1]
        starts = [0] + np.cumsum(size_evals).tolist()
        rows = []
        for s in nearest:
            assert n_warm_per_task < len(init_list[s]), 'n_warm_per_task exceeds source evaluations'
            rows += (starts[s] + init_list[s][:, -1].argsort()[-n_warm_per_task:][::-1]).tolist()
        del previous['target']
        previous = {k: v[rows] for k, v in previous.items()}
        bo = BayesianOptimization(target, bounds, model_type=model_type, random_state=rs, verbose=verbose)
        bo.explore(previous)
        init_list = None
    if gp_class is not None:
        bo.gp_class = gp_class
    if init_list is not None:
        bo.initialize(previous)
    bo.maximize(method, optim_config, init_points=rand_it, n_iter=bo_it, acq=acq, xi=xi, kappa=kappa,
                acq_one_shot=acq_one_shot)
    names = sorted(bo.space.keys)
    params = np.array([[p[k] for k in names] for p in bo.res['all']['params']], dtype=float).reshape(-1, len(names))
    out = np.hstack((params, np.expand_dims(bo.res['all']['values'], 1)))
    if method in ('dist', 'manual', 'multi'):
        sims = bo.res['all']['similarity']
        # one row per acquisition; the first has one entry per source task, the later ones also carry the
        # target task (appended to the GP after the one-shot acquisition) -> ragged, as in the reference
        ragged = len({np.size(s) for s in sims}) > 1
        return out, (np.array(sims, dtype=object) if ragged else np.array(sims))
    return out
