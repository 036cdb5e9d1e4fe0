"""BO driver (reference distBO/bayes_opt/bayesian_optimization.py:15-387).

Host glue around the surrogate: replays the source-task evaluations, fits the joint GP over
(theta, dataset embedding), maximises the acquisition, evaluates the objective, refits.  The
surrogate class is `gp_class` (the sm_100a gp_regressor); the parity tests swap in the CPU oracle's
duck-typed regressor to compare trajectories.
"""
from __future__ import annotations

from copy import deepcopy

import numpy as np

from .helpers import PrintLog, UtilityFunction, acq_max, ensure_rng
from .target_space import TargetSpace


def _default_gp_class():
    from ..train.gp_regressor import gp_regressor
    return gp_regressor


def _row_shares_value(x, rows):
    """numpy's `x in rows`: True when any element of rows equals the element of x in its column
    (what the reference's repeated-point test evaluates, bayesian_optimization.py:333)."""
    rows = np.asarray(rows)
    return bool(rows.size) and bool(np.any(rows == x))


class BayesianOptimization(object):
    gp_class = None     # None -> distbo_b200.train.gp_regressor

    def __init__(self, f, pbounds, target_data=None, source_data=None, size_evals=None,
                 model_type=None, include_init=False, random_state=None, verbose=1):
        have_embed = source_data is not None and source_data[0].embed is not None
        self.source_embed = [s.embed for s in source_data] if have_embed else [None]
        if target_data is not None and target_data.embed is not None:
            self.target_embed = np.expand_dims(target_data.embed, 0)
        else:
            self.target_embed = None
        if source_data is not None:
            self.source_X = [s.train_x for s in source_data]
            self.source_Y = [s.train_y for s in source_data]
            self.source_embed_ind = [s.embed_ind for s in source_data]
        else:
            self.source_X, self.source_Y, self.source_embed_ind = [None], [None], [None]
        if target_data is not None:
            self.target_X, self.target_Y = target_data.train_x, target_data.train_y
            self.target_embed_ind = target_data.embed_ind
        else:
            self.target_X = self.target_Y = self.target_embed_ind = None
        self.size_evals = size_evals if size_evals is not None else [0]
        self.model_type = model_type
        self.pbounds = pbounds
        self.para_dim = len(pbounds)
        self.include_init = include_init
        self.random_state = ensure_rng(random_state)
        # the reference hands the RAW random_state on: an int seed gives the space its own stream
        self.space = TargetSpace(f, pbounds, random_state)
        self.initialized = False
        self.init_points, self.x_init, self.y_init = [], [], []
        self.current_x, self.current_y = [], []
        self.i = 0
        self.util = None
        self.plog = PrintLog(self.space.keys)
        self.res = {'max': {'max_val': None, 'max_params': None},
                    'all': {'values': [], 'params': [], 'similarity': []}}
        self.verbose = verbose

    # ------------------------------------------------------------------ initial design
    def init(self, init_points):
        self.init_points.extend(self.space.random_points(init_points))
        if self.x_init:     # evaluations replayed from the source tasks
            for x, y in zip(np.vstack(self.x_init), np.hstack(self.y_init)):
                self.space.add_observation(x, y)
                if self.include_init:
                    self.res['all']['values'].append(y)
                    self.res['all']['params'].append(dict(zip(self.space.keys, x)))
                if self.verbose:
                    self.plog.print_step(x, y)
        for x in self.init_points:
            y = self._observe_point(x)
            self.current_y.append(y)
            self.res['all']['values'].append(y)
            self.res['all']['params'].append(dict(zip(self.space.keys, x)))
        self.initialized = True

    def _observe_point(self, x):
        y = self.space.observe_point(x)
        if self.verbose:
            self.plog.print_step(x, y)
        return y

    def explore(self, points_dict, eager=False):
        points = self.space._dict_to_points(points_dict)
        if eager:
            self.plog.reset_timer()
            if self.verbose:
                self.plog.print_header(initialization=True)
            for x in points:
                self._observe_point(x)
        else:
            self.init_points = points

    def initialize(self, points_dict):
        self.y_init.extend(points_dict['target'])
        for i in range(len(points_dict['target'])):
            self.x_init.append([points_dict[key][i] for key in self.space.keys])

    def initialize_df(self, points_df):
        for i in points_df.index:
            self.y_init.append(points_df.loc[i, 'target'])
            self.x_init.append([points_df.loc[i, key] for key in self.space.keys])

    def set_bounds(self, new_bounds):
        self.pbounds.update(new_bounds)
        self.space.set_bounds(new_bounds)

    # ------------------------------------------------------------------ the loop
    def maximize(self, kernel, optim_config, warm_start=False, init_points=5, n_iter=25, acq='ucb',
                 kappa=1.0, xi=0.0, acq_one_shot='ei'):
        self.plog.reset_timer()
        if not self.initialized:
            if self.verbose:
                self.plog.print_header()
            self.init(init_points)
        y_max = np.max(self.current_y) if len(self.current_y) != 0 else 0.0

        oc = optim_config
        gp_dict = {'target_X': self.target_X, 'target_Y': self.target_Y, 'n_cpus': oc['n_cpus'],
                   'model_type': self.model_type, 'target_embed_ind': self.target_embed_ind,
                   'float64': oc['float64']}
        fit = {'data_X': self.source_X, 'data_Y': self.source_Y, 'source_embed_ind': self.source_embed_ind,
               'data_embed': self.source_embed, 'target_embed': self.target_embed,
               'weight_dim': oc['weight_dim'], 'weight_dim_x': oc['weight_dim_x'],
               'weight_dim_y': oc['weight_dim_y'], 'weight_dim_xy': oc['weight_dim_xy'],
               'sizes': self.size_evals, 'stratify': oc['stratify'], 'learning_rate': oc['lr'],
               'batch_size': oc['batch_size'], 'num_of_rff': oc['num_of_rff'], 'algorithm': oc['algorithm'],
               'embed_op': oc['embed_op'], 'max_epochs': oc['n_epochs'],
               'concat_data_size': oc['concat_data_size'], 'max_size': oc['max_size'],
               'likhd_var_init': oc['lkh_var_init'], 'gradient_clip': oc['gradient_clip'],
               'first_early_stop_epoch': oc['first_early_stop_epoch']}
        self.util_one_shot = UtilityFunction(kind=acq_one_shot, kappa=kappa, xi=xi)
        if kernel == 'multi':
            self.size_evals.append(init_points)     # the target task already holds its initial points (:282-283)

        if n_iter != 0:
            seed = self.random_state.randint(2 ** 32)
            gp_class = self.gp_class or _default_gp_class()
            self.gp = gp_class(kernel, seed=seed, **gp_dict)
            self.gp.maximise_likhd(self.space.X, np.expand_dims(self.space.Y, 1), **fit)
            init_X, init_Y = deepcopy(self.space.X), deepcopy(self.space.Y)
            self.gp.initialise_predict()
            # one-shot acquisition: no evaluatedX, so a polished point is never accepted (helpers.py:84)
            x_max = acq_max(ac=self.util_one_shot.utility, algorithm=oc['algorithm'], kernel=kernel, gp=self.gp,
                            prev_X=init_X, prev_Y=np.squeeze(init_Y), size_evals=self.size_evals, y_max=y_max,
                            bounds=self.space.bounds, random_state=self.random_state,
                            n_warmup=oc['n_warmup'], n_iter=oc['n_opt_iterations'])
        if self.verbose:
            self.plog.print_header(initialization=False)

        # the target data set joins the GP as a new task with one (upcoming) observation
        if kernel == 'manual':
            self.source_embed.append(np.squeeze(self.target_embed))
            self.size_evals.append(1)
        elif kernel == 'dist':
            self.source_X.append(self.target_X)
            self.source_Y.append(self.target_Y)
            self.size_evals.append(1)
            fit['source_embed_ind'] = self.source_embed_ind + [self.target_embed_ind]
        elif kernel == 'multi':
            self.size_evals[-1] = self.size_evals[-1] + 1
        if kernel in ('dist', 'manual', 'multi'):
            self.res['all']['similarity'].append(np.squeeze(self.gp.similarity()))

        for i in range(n_iter):
            target_rows = self.space.X[int(np.sum(self.size_evals[:-1])):]
            pwarning = False
            while _row_shares_value(x_max, target_rows):
                x_max = self.space.random_points(1)[0]
                pwarning = True
            y = self.space.observe_point(x_max)
            self.current_y.append(y)
            if self.verbose:
                self.plog.print_step(x_max, y, pwarning)
            self.res['all']['values'].append(y)
            self.res['all']['params'].append(dict(zip(self.space.keys, x_max)))
            if i == n_iter - 1:
                break
            self.gp.maximise_likhd(self.space.X, np.expand_dims(self.space.Y, 1), **fit)
            fit['first_early_stop_epoch'] = 600      # from the second in-loop refit on (:354-356)
            y_max = np.max(self.current_y)
            self.util = UtilityFunction(kind=acq, kappa=kappa, xi=xi)
            self.gp.initialise_predict()
            x_max = acq_max(ac=self.util.utility, gp=self.gp, y_max=y_max, bounds=self.space.bounds, kappa=kappa,
                            evaluatedX=self.space.X[int(np.sum(self.size_evals[:-1])):],
                            random_state=self.random_state, n_warmup=oc['n_warmup'],
                            n_iter=oc['n_opt_iterations'])
            if kernel in ('dist', 'manual', 'multi'):
                self.res['all']['similarity'].append(np.squeeze(self.gp.similarity()))
                self.size_evals[-1] = self.size_evals[-1] + 1
        if hasattr(self, 'gp'):
            self.gp.close()
        if self.verbose:
            self.plog.print_summary()

    def points_to_csv(self, file_name):
        points = np.hstack((self.space.X, np.expand_dims(self.space.Y, axis=1)))
        np.savetxt(file_name, points, header=','.join(self.space.keys + ['target']), delimiter=',', comments='')
