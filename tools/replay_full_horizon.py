#!/usr/bin/env python
"""Replays a full-horizon reference run (tests/golden/reffull_<name>.npz, written by the reference's own
experiments/train_test.py through tools/make_full_horizon_golden.py) with the CUDA product or the CPU oracle, and
reports where - if anywhere - the trajectories part.

    python tools/replay_full_horizon.py toy --impl cuda      -> profiles/r02_full_horizon_toy_cuda.json
    python tools/replay_full_horizon.py toy --impl oracle    (CPU; checker-only code, never the product path)

The target search runs through the product's own driver (distbo_b200.experiments.run -> bayes_opt_wrap ->
BayesianOptimization.maximize -> acq_max) at the reference's shipped settings, from the reference's recorded source
evaluations and initial network weights.  Per BO iteration the report holds: suggestion equal (bit for bit when no
polish is configured, 1e-6 otherwise), epochs of the fit, last loss, drift of the fitted parameters against the
reference's.  At the first differing suggestion the acquisition values of both candidates under this
implementation's surrogate are recorded next to the reference's top-2 values of that sweep (the margin that decided).
"""
import argparse
import json
import os
import random
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def real_data(name, args):
    """The reference's data files, committed as fixtures (bit-packed fingerprints / patient tables)."""
    from distbo_b200 import experiments as X
    args.data_dir = os.path.join(GOLD, "%s_full.npz" % name)
    return X.generate_data(args)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("name", choices=["toy", "protein", "parkinson"])
    ap.add_argument("--impl", choices=["cuda", "oracle"], default="cuda")
    ap.add_argument("--out", default=None)
    ap.add_argument("--max-iterations", type=int, default=None, help="replay only the first k target iterations")
    ap.add_argument("--teacher-forced", action="store_true",
                    help="skip training: load the REFERENCE'S fitted parameters of every fit into the surrogate, so that "
                         "each of the acquisitions (posterior + EI over 300 000 candidates + arg-max [+ polish]) is "
                         "compared on the same surrogate, free of the training run's sensitivity")
    opt = ap.parse_args()

    from distbo_b200 import experiments as X
    from distbo_b200.bayes_opt import helpers as H
    import distbo_b200.bayes_opt.bayesian_optimization as BOmod

    z = np.load(os.path.join(GOLD, "reffull_%s.npz" % opt.name), allow_pickle=False)
    argv = json.loads(str(z["argv"]))
    args = X.parse_args(argv + ["/tmp/replay_out"])
    if opt.max_iterations is not None:
        args.target_iterations = min(args.target_iterations, opt.max_iterations)
    polish = args.n_opt_iterations != 0
    init = {k[len("init__"):]: z[k] for k in z.files if k.startswith("init__")}
    if "source_params" in z.files:           # equal-sized source searches: one [T, n, d] array
        source_evals = [np.hstack([sp, se[:, None]]) for sp, se in zip(z["source_params"], z["source_evals"])]
    else:
        n_src = len([k for k in z.files if k.startswith("source_params__")])
        source_evals = [np.hstack([z["source_params__%d" % i], z["source_evals__%d" % i][:, None]])
                        for i in range(n_src)]
    want_x, want_y = z["target_params"], z["target_evals"]

    if opt.impl == "cuda":
        from distbo_b200.train import gp_regressor as base
    else:
        from oracle.gp_regressor_oracle import OracleGPRegressor as base   # checker code: tools/ only

    fits, first_fit_losses = [], []

    def factory(kernel, **kw):
        reg = base(kernel, **kw)
        reg.initial_weights = init
        orig = reg.maximise_likhd

        def recording(*a, **k):
            t0 = time.time()
            if opt.teacher_forced:
                k = dict(k, max_epochs=0)
            loss = orig(*a, **k)
            if opt.teacher_forced:
                j = len(fits)
                for name in list(reg.engine.params.offsets if opt.impl == "cuda" else reg.p):
                    key = "fit_params__%d__%s" % (j, name)
                    if key in z.files:
                        if opt.impl == "cuda":
                            reg.engine.params.set(name, z[key])
                        else:
                            reg.p[name] = np.asarray(z[key], dtype=np.float64).reshape(np.shape(reg.p[name]))
                if opt.impl == "cuda":
                    reg._fit_generation += 1
                loss = float(z["fit_loss"][j])
                reg.last_epoch = int(z["fit_epochs"][j]) - 1
            elif not fits:
                first_fit_losses.extend(reg.loss_history)
            if opt.impl == "cuda":
                p = {n: reg.engine.params.get(n) for n in reg.engine.params.offsets}
            else:
                p = {n: np.asarray(v) for n, v in reg.p.items()}
            fits.append({"epochs": int(reg.last_epoch) + 1, "loss": float(loss), "seconds": time.time() - t0,
                         "params": p})
            print("[replay %s/%s] fit %d: N=%d epochs=%d loss=%.8f %.1fs" % (
                opt.name, opt.impl, len(fits), len(a[0]), fits[-1]["epochs"], loss, fits[-1]["seconds"]),
                file=sys.stderr, flush=True)
            return loss
        reg.maximise_likhd = recording
        return reg

    # first differing suggestion: acquisition values of both candidates under THIS implementation's surrogate
    sweep = {"i": 0, "first_diff": None}
    orig_acq_max = H.acq_max

    def watching_acq_max(ac, gp, y_max, bounds, random_state, **kw):
        x = orig_acq_max(ac, gp, y_max, bounds, random_state, **kw)
        i = sweep["i"]
        sweep["i"] += 1
        if sweep["first_diff"] is None and i < len(want_x):
            same = np.allclose(x, want_x[i], rtol=1e-6, atol=1e-8) if polish else np.array_equal(x, want_x[i])
            if not same:
                both = np.vstack([np.ravel(x), np.ravel(want_x[i])])
                vals = np.reshape(ac(both, gp=gp, y_max=y_max), -1)
                sweep["first_diff"] = {"iteration": i, "own_choice": np.ravel(x).tolist(),
                                       "reference_choice": np.ravel(want_x[i]).tolist(),
                                       "acq_own_choice": float(vals[0]), "acq_reference_choice": float(vals[1]),
                                       "y_max": float(y_max)}
        return x
    BOmod.acq_max = watching_acq_max

    def restore_streams():
        if "np_random_keys" in z.files:
            np.random.set_state(("MT19937", z["np_random_keys"], int(z["np_random_pos"]), 0, 0.0))
            st = json.loads(str(z["py_random_state"]))
            random.setstate((st[0], tuple(st[1]), st[2]))

    data = None if opt.name == "toy" else real_data(opt.name, args)
    t0 = time.time()
    try:
        d = X.run(args, surrogate=factory, source_evals=source_evals, data=data, on_target_start=restore_streams,
                  verbose=0)
    finally:
        BOmod.acq_max = orig_acq_max
    wall = time.time() - t0

    got_x, got_y = np.asarray(d["target_params"], dtype=np.float64), np.asarray(d["target_evals"], dtype=np.float64)
    n = min(len(got_x), len(want_x))
    eq = [bool(np.allclose(got_x[i], want_x[i], rtol=1e-6, atol=1e-8) if polish else np.array_equal(got_x[i], want_x[i]))
          for i in range(n)]
    first_bad = next((i for i, e in enumerate(eq) if not e), None)
    rows = []
    for j, f in enumerate(fits):
        row = {"fit": j, "epochs": f["epochs"], "loss": f["loss"], "seconds": round(f["seconds"], 2)}
        if j < len(z["fit_epochs"]):
            row["reference_epochs"] = int(z["fit_epochs"][j])
            row["reference_loss"] = float(z["fit_loss"][j])
            drift = {}
            for name, v in f["params"].items():
                key = "fit_params__%d__%s" % (j, name)
                if key in z.files:
                    ref = np.asarray(z[key], dtype=np.float64).reshape(-1)
                    drift[name] = float(np.max(np.abs(np.asarray(v).reshape(-1) - ref)) / max(np.max(np.abs(ref)), 1e-300))
            row["max_param_drift_rel"] = max(drift.values()) if drift else None
            row["worst_param"] = max(drift, key=drift.get) if drift else None
        rows.append(row)
    sim_drift = None
    if "target_sim" in z.files and d.get("target_sim") is not None:
        worst = 0.0
        for got, ref in zip(d["target_sim"], z["target_sim"]):
            ref = ref[~np.isnan(ref)]
            g = np.asarray(got, dtype=np.float64).reshape(-1)
            if len(g) == len(ref):
                worst = max(worst, float(np.max(np.abs(g - ref) / np.maximum(np.abs(ref), 1e-300))))
        sim_drift = worst
    report = {
        "config": opt.name, "impl": opt.impl, "teacher_forced": bool(opt.teacher_forced), "argv": argv,
        "iterations": n, "polish": polish,
        "suggestions_equal": eq, "all_equal": first_bad is None, "first_differing_iteration": first_bad,
        "objective_values_equal": bool(np.array_equal(got_y[:n], want_y[:n])),
        "max_abs_suggestion_diff": float(np.max(np.abs(got_x[:n] - want_x[:n]))) if n else None,
        "fits": rows, "similarity_max_rel_drift": sim_drift, "first_difference": sweep["first_diff"],
        "reference_sweep_top2": z["sweep_top2"].tolist() if "sweep_top2" in z.files else None,
        "reference_target_seconds": float(z["target_time"]), "replay_target_seconds": float(d["target_time"]),
        "wall_seconds": wall,
        "polish_stats": dict(H.POLISH_STATS),
    }
    tag = "%s_%s%s" % (opt.name, opt.impl, "_teacher_forced" if opt.teacher_forced else "")
    out = opt.out or os.path.join(ROOT, "profiles", "r02_full_horizon_%s.json" % tag)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as fh:
        json.dump(report, fh, indent=1)
    if first_fit_losses:                  # every epoch's loss of the first fit: input of the divergence report
        np.save(out[:-5] + "_first_fit_losses.npy", np.asarray(first_fit_losses))
    print("[replay %s/%s] %d iterations, all suggestions equal: %s (first difference: %s); target search %.1f s "
          "(reference over the shim: %.1f s) -> %s" % (opt.name, opt.impl, n, first_bad is None, first_bad,
                                                       d["target_time"], float(z["target_time"]), out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
