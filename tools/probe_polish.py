"""Cost of the acquisition step at the shipped-config sizes: the candidate sweep (n_warmup = 300000) and the
L-BFGS-B polish (n_opt_iterations = 10) that follows it, through distbo_b200.bayes_opt.helpers.acq_max."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from distbo_b200.bayes_opt.helpers import UtilityFunction, acq_max
from distbo_b200.train import gp_regressor

out = {}
for name, N, d in (("protein_like", 140, 4), ("parkinson_like", 1260, 2), ("counter_like", 700, 7)):
    rs = np.random.RandomState(0)
    theta = rs.uniform(-5, 5, size=(N, d))
    y = np.sin(theta).sum(1, keepdims=True) + 0.05 * rs.randn(N, 1)
    g = gp_regressor("params", seed=1, float64=True)
    g.maximise_likhd(theta, y, learning_rate=0.005, max_epochs=30, likhd_var_init=1e-2)
    util = UtilityFunction("ei", kappa=2.58, xi=0.01)
    bounds = np.array([[-5.0, 5.0]] * d)
    res = {}
    for n_iter in (0, 10):
        calls = {"n": 0, "rows": 0}
        orig = g.predict_y

        def counted(x, **kw):
            calls["n"] += 1
            calls["rows"] += len(x)
            return orig(x, **kw)

        g.predict_y = counted
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x = acq_max(util.utility, g, float(y.max()), bounds, np.random.RandomState(3), n_warmup=300000, n_iter=n_iter,
                    evaluatedX=theta[-5:])
        torch.cuda.synchronize()
        res["n_iter_%d" % n_iter] = {"seconds": time.perf_counter() - t0, "predict_calls": calls["n"],
                                     "rows": calls["rows"]}
        g.predict_y = orig
    t0 = time.perf_counter()
    for _ in range(200):
        g.predict_y(theta[:1])
    res["single_point_predict_us"] = (time.perf_counter() - t0) / 200 * 1e6
    out[name] = res
print(json.dumps(out, indent=1))
