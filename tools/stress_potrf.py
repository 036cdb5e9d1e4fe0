"""Race / determinism stress of the multi-stream POTRF schedules: the factor, its inverse and K^-1 of the same matrix
must be BITWISE identical over many eager runs and CUDA-graph replays (fixed tile products, fixed orders)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from distbo_b200 import engine as eng

out = {}
for N in (8192, 7300, 3000):
    rs = np.random.RandomState(1)
    theta, y = rs.randn(N, 8), rs.randn(N)
    e = eng.GPEngine(8)
    e.params.set("log_sigma", 0.5 * np.log(1e-2))
    e.set_observations(theta, y)

    def run():
        e.build_gram(None)
        e.factorize(need_inverse=True, need_kinv=True)
        e.likelihood()

    run()
    torch.cuda.synchronize()
    ref = [e.K.clone(), e.W.clone(), e.Tm.clone(), e.nll.clone()]
    bad = 0
    reps = 60 if N > 4000 else 150
    for _ in range(reps):
        run()
        torch.cuda.synchronize()
        bad += int(not (torch.equal(e.K, ref[0]) and torch.equal(e.W, ref[1]) and torch.equal(e.Tm, ref[2])
                        and torch.equal(e.nll, ref[3])))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    bad_g = 0
    for _ in range(reps):
        g.replay()
        torch.cuda.synchronize()
        bad_g += int(not (torch.equal(e.K, ref[0]) and torch.equal(e.W, ref[1]) and torch.equal(e.Tm, ref[2])))
    # back-to-back replays without host synchronisation in between (what the training loop does)
    for _ in range(reps):
        g.replay()
    torch.cuda.synchronize()
    bad_b = int(not (torch.equal(e.K, ref[0]) and torch.equal(e.W, ref[1]) and torch.equal(e.Tm, ref[2])))
    out[str(N)] = {"eager_mismatches": bad, "graph_mismatches": bad_g, "back_to_back_mismatch": bad_b, "reps": reps,
                   "info": int(e.info.item())}
print(json.dumps(out))
