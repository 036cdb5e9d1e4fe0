"""Kinv = W^T W: the tcgen05 digit-plane kernel against the DMMA tile GEMM (and torch) on one factorisation.
Usage: python tools/probe_lauum_i8.py N [sigma2] [reps]"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from distbo_b200 import engine as eng  # noqa: E402
from distbo_b200 import _lib  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    s2 = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-2
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    d = 4
    torch.cuda.set_device(0)
    rs = np.random.RandomState(0)
    theta = rs.randn(N, d)
    y = np.sin(theta.sum(1)) + 0.1 * rs.randn(N)
    res = {"N": N, "sigma2": s2}
    kin = {}
    for mode in (0, 1):
        e = eng.GPEngine(d)
        e.use_lauum_i8, e.lauum_i8_min_npad = bool(mode), 128
        e.params.set("log_sigma", 0.5 * np.log(s2))
        e.set_observations(theta, y)
        e.build_gram()
        e.factorize(need_inverse=True, need_kinv=True)
        e.likelihood()
        e.read_loss()
        kin[mode] = torch.tril(e.Tm[:N, :N]).clone()
        if mode == 1:
            W = torch.tril(e.W[:N, :N])
            ref = torch.tril(W.T @ W)
            res["i8_vs_torch_rel"] = float((kin[1] - ref).abs().max() / ref.abs().max())
        if reps:
            st = torch.cuda.current_stream().cuda_stream
            lib = e.lib
            def run():
                if mode:
                    lib.distbo_lauum_i8_f64(_lib.ptr(e.W), e.n_pad, _lib.ptr(e._lauum_i8[0]), _lib.ptr(e._lauum_i8[1]),
                                            _lib.ptr(e.Tm), st)
                else:
                    lib.distbo_lauum_f64(e.plan, _lib.ptr(e.W), _lib.ptr(e.Tm), st)
            run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                run()
            e1.record()
            torch.cuda.synchronize()
            res["ms_mode%d" % mode] = e0.elapsed_time(e1) / reps
        e.close()
    res["i8_vs_dmma_rel"] = float((kin[1] - kin[0]).abs().max() / kin[0].abs().max())
    print(json.dumps(res))


if __name__ == "__main__":
    main()
