#!/usr/bin/env python
"""Full-horizon BO trajectories written by the REFERENCE'S OWN DRIVER at its shipped settings.

`experiments/train_test.py main()` of /root/reference is executed unmodified (runpy, its own argparse command line
as printed by experiments/one_dim_toy_config.py / protein_config.py, or the CLI defaults for Parkinson) with
`tensorflow` bound to tools/tf1_shim.py, exactly as tools/make_reference_golden.py does for the short cases:
10 000 Adam epochs with the reference's early stopping, 300 000 acquisition candidates, the source searches and
the whole target search.  The file it writes (`results.npz`, train_test.py:326-338) is the golden; next to it this
script records what a replay needs and what a divergence report needs:

  init__<name>          initial values of the target surrogate's variables (TF's glorot stream is not reproducible)
  fit_epochs, fit_loss  epochs run and last loss of every maximise_likhd call of the target search
  fit_params__<j>__<name>  fitted parameters after target fit j;  fit_loss_log__<j>: its loss at every epoch
  sweep_top2            the two largest acquisition values of every candidate sweep and their row indices
                        (the margin that decides the suggestion: arbiter input)

    python tools/make_full_horizon_golden.py toy        -> tests/golden/reffull_toy.npz       (config 1)
    python tools/make_full_horizon_golden.py protein    -> tests/golden/reffull_protein.npz   (config 2)
    python tools/make_full_horizon_golden.py parkinson  -> tests/golden/reffull_parkinson.npz (config 3)

Long: tens of minutes to hours per configuration on the build container's CPU.
"""
import json
import os
import runpy
import shutil
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("DISTBO_REFERENCE", "/root/reference")
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import tf1_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

from distbo_b200.experiments import SHIPPED_COMMAND_LINES  # noqa: E402

# the command lines experiments/*_config.py print for seed 0, method distBO, algorithm GP (config 3: CLI defaults);
# protein runs the reference's default objective, an unseeded random forest (models.py:225-237) drawing from numpy's
# global stream, which main() seeds once - its state at the start of the target search is recorded for the replays
CONFIGS = {"toy": SHIPPED_COMMAND_LINES["config1"], "protein": SHIPPED_COMMAND_LINES["config2"],
           "parkinson": SHIPPED_COMMAND_LINES["config3"]}


def install_py3_shims():
    """Environment shims only (nothing numeric): see tools/make_reference_golden.py."""
    import sklearn.datasets
    np.float = float                                            # bayes_opt/target_space.py:50
    sklearn.datasets.__dict__.setdefault("load_boston", lambda *a, **k: None)   # model_embed/meta_fea.py:8
    km = types.ModuleType("sklearn.cluster.k_means_")           # data/toy_datasets.py:11
    km._init_centroids = None
    sys.modules.setdefault("sklearn.cluster.k_means_", km)
    sys.modules.setdefault("cPickle", __import__("pickle"))     # distBO/utils.py:70


class _NumpyCompat(object):
    """numpy >= 1.24 refuses the ragged np.array(list of similarity vectors) at bayes_optimise.py:60."""

    def __getattr__(self, k):
        return getattr(np, k)

    @staticmethod
    def array(obj, *a, **kw):
        try:
            return np.array(obj, *a, **kw)
        except ValueError:
            out = np.empty(len(obj), dtype=object)
            for i, v in enumerate(obj):
                out[i] = v
            return out


def main():
    name = sys.argv[1]
    argv = CONFIGS[name]
    out_dir = os.path.join(ROOT, "gpurun_out", "fullhorizon", name)
    shutil.rmtree(out_dir, ignore_errors=True)
    import random
    random.seed(0)                  # gp_utils.py:53,78 shuffle with the unseeded global `random`
    np.random.seed(0)               # the protein objective (an unseeded random forest, models.py:225-237) draws from
    #                                 numpy's global stream; its state at the start of the target search is recorded
    import torch
    torch.set_num_threads(int(os.environ.get("DISTBO_GOLD_THREADS", "4")))
    tf = tf1_shim.install()
    tf.reset_default_graph()
    tf1_shim._EAGER_SEED[0] = 0
    install_py3_shims()
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "distBO", "train"))    # train.py:8 py2 implicit relative import
    sys.path.insert(0, os.path.join(REF, "experiments"))

    import distBO.train.gp_regressor as ref_gp
    import distBO.model_embed.bayes_optimise as ref_bo_mod
    import distBO.bayes_opt.helpers as ref_helpers
    import distBO.data.load_datasets as ref_load
    from scipy.optimize import minimize as _minimize
    ref_bo_mod.np = _NumpyCompat()
    # scipy >= 1.11 rejects the 2-D x0 that helpers.py:75-77 passes to minimize (older scipy ravelled it)
    ref_helpers.minimize = lambda fun, x0, *a, **kw: _minimize(fun, np.asarray(x0).ravel(), *a, **kw)
    # py2 `range(...).remove(...)` in the real-data loaders (load_datasets.py:42-43,82-83): give that module a list-
    # returning range
    ref_load.range = lambda *a: list(range(*a))

    rec = {"inits": [], "fit_epochs": [], "fit_loss": [], "fit_params": [], "fit_kernel": [], "fit_seconds": [],
           "sweeps": [], "fit_loss_log": []}
    orig_fit = ref_gp.gp_regressor.maximise_likhd
    orig_close = ref_gp.gp_regressor.close
    t_start = time.time()

    def recording_fit(self, *a, **kw):
        t0 = time.time()
        n0 = len(self.sess.loss_log) if hasattr(self, "sess") else 0
        loss = orig_fit(self, *a, **kw)
        n1 = len(self.sess.loss_log)
        rec["fit_epochs"].append(n1 - n0)
        rec["fit_loss"].append(float(loss))
        rec["fit_kernel"].append(self.kernel)
        rec["fit_seconds"].append(time.time() - t0)
        rec["fit_params"].append({k: v.value.detach().numpy().copy() for k, v in self.net.params.items()})
        rec["fit_loss_log"].append(np.asarray(self.sess.loss_log[n0:n1]) if self.kernel != "params" else None)
        print("[full-horizon %s] fit %d kernel=%s N=%d epochs=%d loss=%.8f %.1fs (total %.0fs)" % (
            name, len(rec["fit_loss"]), self.kernel, len(a[0]), n1 - n0, loss, time.time() - t0,
            time.time() - t_start), file=sys.stderr, flush=True)
        return loss

    def recording_close(self):
        rec["inits"].append((self.kernel, {k: tf1_shim.INIT_VALUES[id(v)] for k, v in self.net.params.items()}))
        return orig_close(self)

    orig_utility = ref_helpers.UtilityFunction.utility

    def recording_utility(self, x, gp, y_max, compute_embed=False):
        ys = orig_utility(self, x, gp, y_max, compute_embed=compute_embed)
        flat = np.reshape(np.asarray(ys, dtype=np.float64), -1)
        if flat.size > 1 and gp.kernel != "params":
            top = np.argsort(flat, kind="stable")[-2:][::-1]
            rec["sweeps"].append([float(flat[top[0]]), float(flat[top[1]]), int(top[0]), int(top[1]), float(y_max),
                                  float(flat.size)])
        return ys

    orig_wrap = ref_bo_mod.bayes_opt_wrap

    def recording_wrap(method, *a, **kw):
        if method != "params":
            st = np.random.get_state()
            rec["np_state"] = (np.asarray(st[1], dtype=np.uint32), int(st[2]))
            rec["py_random_state"] = random.getstate()
        return orig_wrap(method, *a, **kw)
    ref_bo_mod.bayes_opt_wrap = recording_wrap
    ref_gp.gp_regressor.maximise_likhd = recording_fit
    ref_gp.gp_regressor.close = recording_close
    ref_helpers.UtilityFunction.utility = recording_utility

    old_argv, old_cwd = sys.argv, os.getcwd()
    sys.argv = ["train_test.py"] + argv + [out_dir]
    os.chdir(REF)                   # the real-data loaders read protein/ and parkinson/ relative to the repository
    try:
        runpy.run_path(os.path.join(REF, "experiments", "train_test.py"), run_name="__main__")
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)
        ref_gp.gp_regressor.maximise_likhd = orig_fit
        ref_gp.gp_regressor.close = orig_close
        ref_helpers.UtilityFunction.utility = orig_utility

    z = np.load(os.path.join(out_dir, "results.npz"), allow_pickle=True, encoding="latin1")
    arrays = {}
    for k in ("source_params", "source_evals", "target_params", "target_evals"):
        v = z[k]
        if v.dtype == object:
            for i, a in enumerate(v):
                arrays["%s__%d" % (k, i)] = np.asarray(a, dtype=np.float64)
        else:
            arrays[k] = np.asarray(v, dtype=np.float64)
    if "target_sim" in z.files:
        rows = [np.asarray(s_, dtype=np.float64).reshape(-1) for s_ in z["target_sim"]]
        w = max(map(len, rows))
        arrays["target_sim"] = np.asarray([np.r_[r, np.full(w - len(r), np.nan)] for r in rows])
    arrays["source_time"] = np.asarray(float(z["source_time"]))
    arrays["target_time"] = np.asarray(float(z["target_time"]))
    arrays["results_keys"] = np.asarray(sorted(z.files))
    tgt = [i for i, k in enumerate(rec["fit_kernel"]) if k != "params"]
    arrays["fit_epochs"] = np.asarray([rec["fit_epochs"][i] for i in tgt])
    arrays["fit_loss"] = np.asarray([rec["fit_loss"][i] for i in tgt])
    arrays["fit_seconds"] = np.asarray([rec["fit_seconds"][i] for i in tgt])
    arrays["source_fit_epochs"] = np.asarray([e for e, k in zip(rec["fit_epochs"], rec["fit_kernel"]) if k == "params"])
    for j, i in enumerate(tgt):
        for k, v in rec["fit_params"][i].items():
            arrays["fit_params__%d__%s" % (j, k)] = v
        arrays["fit_loss_log__%d" % j] = rec["fit_loss_log"][i]      # every epoch's loss (divergence reports)
    arrays["sweep_top2"] = np.asarray(rec["sweeps"], dtype=np.float64).reshape(-1, 6)
    for kernel, init in rec["inits"]:
        if kernel != "params":
            for k, v in init.items():
                arrays["init__" + k] = v
    if "np_state" in rec:
        arrays["np_random_keys"], arrays["np_random_pos"] = rec["np_state"][0], np.asarray(rec["np_state"][1])
        arrays["py_random_state"] = np.asarray(json.dumps(rec["py_random_state"]))
    arrays["argv"] = np.asarray(json.dumps(argv))
    np.savez_compressed(os.path.join(OUT, "reffull_%s.npz" % name), **arrays)
    print("[full-horizon %s] done in %.0f s: target suggestions\n%s" % (
        name, time.time() - t_start, np.array2string(np.asarray(z["target_params"], dtype=np.float64), precision=6)))


if __name__ == "__main__":
    main()
