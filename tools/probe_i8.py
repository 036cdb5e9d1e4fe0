"""A/B of the two scoring paths on one fitted GP: DMMA (fp64 tensor cores) against tcgen05 int8 digit planes.
Usage: python tools/probe_i8.py N d M [sigma2] [reps]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from distbo_b200 import engine as eng  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    M = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
    s2 = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-2
    reps = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    torch.cuda.set_device(0)
    rs = np.random.RandomState(0)
    theta = rs.randn(N, d)
    y = np.sin(theta.sum(1)) + 0.1 * rs.randn(N)
    e = eng.GPEngine(d)
    e.params.set("log_sigma", 0.5 * np.log(s2))
    e.set_observations(theta, y)
    e.build_gram()
    e.factorize(need_inverse=True, need_kinv=False)
    e.likelihood()
    e.read_loss()
    e.kvec[:N] = torch.from_numpy(0.3 + 0.7 * rs.rand(N)).to(e.dev)
    cand = torch.from_numpy(rs.randn(M, d)).cuda()
    out = {}
    for mode in (0, 1):
        e.use_score_i8 = bool(mode)
        r = e.score(cand, acq="ei", y_max=float(y.max()), xi=0.01)
        torch.cuda.synchronize()
        out[mode] = {k: v.clone() for k, v in r.items() if v is not None}
    scale = float(np.exp(e.params.view("log_scale").cpu().numpy()[0]))
    dm = (out[0]["mean"] - out[1]["mean"]).abs().max().item()
    var0 = out[0]["std"] ** 2
    var1 = out[1]["std"] ** 2
    dv = (var0 - var1).abs().max().item() / scale
    res = {"N": N, "d": d, "M": M, "sigma2": s2, "max_abs_dmean": dm, "max_dvar_rel_scale": dv,
           "argmax": [int(out[0]["best_idx"].item()), int(out[1]["best_idx"].item())],
           "best": [float(out[0]["best_val"].item()), float(out[1]["best_val"].item())],
           "mean_range": [float(out[0]["mean"].min()), float(out[0]["mean"].max())]}
    if reps:
        for mode in (0, 1):
            e.use_score_i8 = bool(mode)
            e.score(cand, acq="ei", y_max=float(y.max()), xi=0.01, want_arrays=False)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                e.score(cand, acq="ei", y_max=float(y.max()), xi=0.01, want_arrays=False)
            torch.cuda.synchronize()
            res["ms_mode%d" % mode] = (time.perf_counter() - t0) / reps * 1e3
    print(json.dumps(res))


if __name__ == "__main__":
    main()
