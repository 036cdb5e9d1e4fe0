"""W = L^-1: top levels on the tcgen05 digit-plane kernel against the all-DMMA route on one factorisation.
Usage: python tools/probe_trtri_i8.py N [sigma2] [reps] [min_rows]"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from distbo_b200 import engine as eng  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    s2 = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-2
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    mr = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
    d = 4
    torch.cuda.set_device(0)
    rs = np.random.RandomState(0)
    theta = rs.randn(N, d)
    y = np.sin(theta.sum(1)) + 0.1 * rs.randn(N)
    res = {"N": N, "sigma2": s2, "min_rows": mr}
    Ws = {}
    for mode in (0, 1):
        e = eng.GPEngine(d)
        e.use_trtri_i8, e.trtri_i8_min_rows = bool(mode), mr
        e.use_lauum_i8 = False
        e.params.set("log_sigma", 0.5 * np.log(s2))
        e.set_observations(theta, y)
        e.build_gram()
        e.factorize(need_inverse=True, need_kinv=False)
        e.likelihood()
        e.read_loss()
        res["i8_levels_mode%d" % mode] = None if e._trtri_i8 is None else int(e._trtri_i8[0])
        Ws[mode] = torch.tril(e.W[:N, :N]).clone()
        if mode == 1:
            Lm = torch.tril(e.K[:N, :N])
            I = Lm @ Ws[1]
            res["resid_LW_minus_I"] = float((I - torch.eye(N, device=I.device, dtype=I.dtype)).abs().max())
            I0 = Lm @ Ws[0]
            res["resid_dmma"] = float((I0 - torch.eye(N, device=I.device, dtype=I.dtype)).abs().max())
        if reps:
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for it in range(reps + 1):
                if it == 1:
                    e0.record()
                e.build_gram()
                e.factorize(need_inverse=True, need_kinv=False)
            e1.record()
            torch.cuda.synchronize()
            res["ms_gram_potrf_trtri_mode%d" % mode] = e0.elapsed_time(e1) / reps
        e.close()
    res["i8_vs_dmma_rel"] = float((Ws[1] - Ws[0]).abs().max() / Ws[0].abs().max())
    print(json.dumps(res))


if __name__ == "__main__":
    main()
