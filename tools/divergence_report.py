#!/usr/bin/env python
"""Quantifies how the first target fit of BASELINE config 1 (shipped settings: 10 000 Adam epochs, early stopping
from epoch 5000, sigma^2 = 1e-5) separates between implementations: per-epoch losses of
  ref3   the reference's own code over tools/tf1_shim.py, 3 BLAS threads (= tests/golden/reffull_toy.npz)
  ref1   the same unmodified code, 1 BLAS thread            (tools/reference_first_fit.py)
  oracle the CPU restatement                                  (tools/replay_full_horizon.py --impl oracle)
  cuda   the sm_100a product                                  (tools/replay_full_horizon.py --impl cuda, on a B200)
-> a markdown table on stdout (committed as profiles/r02_full_horizon_divergence.md)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
S = os.path.join(ROOT, "gpurun_out", "selfdiv")
runs = {
    "ref3": np.load(os.path.join(S, "ref_threads3.npz"))["loss_log"],
    "ref1": np.load(os.path.join(S, "ref_threads1.npz"))["loss_log"],
    "oracle": np.load(os.path.join(S, "replay_toy_oracle_2it_first_fit_losses.npy")),
    "cuda": np.load(os.path.join(ROOT, sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r2b/full_horizon_toy_cuda_first_fit_losses.npy")),
}
base = runs["ref3"]
print("| run | epochs until the early stop | last loss | first epoch with rel. loss difference vs ref3 > 1e-13 | > 1e-11 | > 1e-9 | > 1e-7 | > 1e-5 | max rel. difference over the common epochs |")
print("|---|---|---|---|---|---|---|---|---|")
for name, l in runs.items():
    n = min(len(l), len(base))
    rel = np.abs(l[:n] - base[:n]) / np.abs(base[:n])
    first = []
    for thr in (1e-13, 1e-11, 1e-9, 1e-7, 1e-5):
        idx = np.nonzero(rel > thr)[0]
        first.append(str(int(idx[0])) if len(idx) else "never")
    print("| %s | %d | %.8f | %s | %.2e |" % (name, len(l), l[-1], " | ".join(first), rel.max()))
# growth between thresholds for the reference against itself
rel = np.abs(runs["ref1"][:min(len(runs["ref1"]), len(base))] - base[:min(len(runs["ref1"]), len(base))]) / np.abs(base[:min(len(runs["ref1"]), len(base))])
for a in (50, 100, 200, 500, 1000, 2000, 4000, 5000):
    if a < len(rel):
        print("ref1 vs ref3: max rel. loss difference up to epoch %5d: %.3e" % (a, rel[:a + 1].max()))
