import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_round2 import toy_problem, manual_fit_kwargs
from distbo_b200.train import gp_regressor
theta, y, sizes, embed = toy_problem(2, T=5, per_task=60, d=3)
for dev in ("cuda:0", "cuda:0", "cuda:1", "cuda:1"):
    for graph in ("1", "0"):
        os.environ["DISTBO_GRAPH"] = graph
        g = gp_regressor("manual", seed=3, float64=True)
        g.device = dev
        loss = g.maximise_likhd(theta, y, **manual_fit_kwargs(sizes, embed, max_epochs=10))
        print(dev, "graph", graph, "loss", loss, "hist", np.round(g.loss_history, 4).tolist(), "step", g.engine.params.step, flush=True)
        g.close()
