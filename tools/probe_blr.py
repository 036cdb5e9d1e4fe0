"""BLR surrogate at the config-5 shape: one training step (loss + full gradient + Adam), posterior preparation and
the acquisition pass over 2^20 candidates, CUDA-event timed; the oracle's CPU time for the same pieces beside it."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from distbo_b200 import blr_engine

N, T, d, P, M = 8192, 64, 8, 23, 1 << 20
rs = np.random.RandomState(0)
theta, y = rs.randn(N, d), rs.randn(N, 1)
meta = rs.rand(T + 1, P)
e = blr_engine.BLREngine(d, [50, 50, 50], n_embed=P)
from distbo_b200.train.gp_utils import init_net_weights
for k, v in init_net_weights(P + d, [50, 50, 50], "", 3).items():
    e.params.set(k, v)
e.params.set("log_sigma", 0.5 * np.log(1e-2))
e.set_observations(theta, y, [N // T] * T)
E = torch.from_numpy(meta[:T]).cuda()


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def step():
    e.loss_and_grad(E_const=E)
    e.adam_step(0.005)


out = {"N": N, "T": T, "d": d, "P": P, "weight_dim": [50, 50, 50], "candidates": M}
out["fit_step_ms"] = timed(step, 50)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step()
out["fit_step_graph_ms"] = timed(g.replay, 200)
out["prepare_posterior_ms"] = timed(lambda: e.prepare_posterior(E, torch.from_numpy(meta[T]).cuda()), 50)
cand = torch.from_numpy(rs.uniform(-3, 3, size=(M, d))).cuda()
out["score_ms"] = timed(lambda: e.score(cand, acq="ei", y_max=float(y.max()), xi=0.01, want_arrays=False), 10)
out["score_candidates_per_s"] = M / out["score_ms"] * 1e3
out["score_dfma_per_candidate"] = (d + 50 + 50 + 50) * 64 + 100
out["score_tdfma_per_s"] = out["score_dfma_per_candidate"] * M / out["score_ms"] * 1e-9

# CPU oracle beside it (numpy, host threads as configured)
from oracle import blr_oracle as BO, gp_oracle as O
p = {k: e.params.get(k) for k in e.params.offsets}
batch = O.TrainBatch(h_params=theta, obs=y, sizes=[N // T] * T, embed=meta[:T])
t0 = time.perf_counter()
BO.blr_loss_and_grads(p, "manual", batch, 50, 0)
out["cpu_fit_step_ms"] = (time.perf_counter() - t0) * 1e3
fea = BO.train_feature(O.NP, p, "manual", batch)
phi = BO.nn_features(O.NP, p, fea, 50, 0)
Ms = 1 << 17
cs = cand[:Ms].cpu().numpy()
t0 = time.perf_counter()
phi_c = BO.nn_features(O.NP, p, np.concatenate([np.repeat(meta[T][None], Ms, 0), cs], 1), 50, 0)
mean, var = BO.blr_predict(O.NP, p, phi, y, phi_c)
a = O.acq_ei(mean, O.std_from_var(var), float(y.max()), 0.01)
out["cpu_score_candidates_per_s"] = Ms / (time.perf_counter() - t0)
res = e.score(cand[:Ms], acq="ei", y_max=float(y.max()), xi=0.01, want_arrays=False)
out["argmax_matches_cpu_on_sample"] = int(res["best_idx"].item()) == int(np.argmax(a))
print(json.dumps(out))
