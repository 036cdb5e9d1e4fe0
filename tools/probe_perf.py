"""GPU probe: FP64 peak (cuBLAS DGEMM via torch.matmul) and timings of the fit / scoring stages
at the BASELINE config-5 shape.  Prints one JSON line per measurement."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from distbo_b200 import engine as eng  # noqa: E402


def ev_time(fn, iters=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 128 * 2
    d = 8
    dev = torch.device("cuda")
    a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    b = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    best, med = ev_time(lambda: torch.matmul(a, b), iters=10, warm=3)
    print(json.dumps({"what": "cublas_dgemm_8192", "ms_best": best, "ms_med": med,
                      "tflops_best": 2 * 8192**3 / best / 1e9, "tflops_med": 2 * 8192**3 / med / 1e9}), flush=True)
    t0 = time.time()
    n = 0
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 3.0:
        torch.matmul(a, b); n += 1
        if n % 20 == 0:
            torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"what": "cublas_dgemm_8192_sustained", "tflops": 2 * 8192**3 * n / e0.elapsed_time(e1) / 1e9}), flush=True)
    best, med = ev_time(lambda: torch.linalg.cholesky(a @ a.T + 8192 * torch.eye(8192, dtype=torch.float64, device=dev)), iters=3)
    spd = a @ a.T + 8192 * torch.eye(8192, dtype=torch.float64, device=dev)
    best, med = ev_time(lambda: torch.linalg.cholesky(spd), iters=5)
    print(json.dumps({"what": "cusolver_potrf_8192", "ms_best": best, "tflops": 8192**3 / 3 / best / 1e9}), flush=True)
    del a, b, spd

    rs = np.random.RandomState(0)
    T = 64
    theta = rs.uniform(-5, 5, size=(N, d))
    theta = (theta - theta.mean(0)) / theta.std(0)
    y = rs.randn(N)
    sizes = [N // T] * T
    E = rs.randn(T, 23) * 0.3
    e = eng.GPEngine(d, n_embed=23)
    e.params.set("log_sigma", 0.5 * np.log(1e-2))
    e.set_observations(theta, y, sizes)
    e.E = torch.from_numpy(E).to(e.dev)
    KE = e.task_gram(e.E, T)

    def t(name, fn, flops=None, iters=5):
        best, med = ev_time(fn, iters=iters)
        rec = {"what": name, "ms_best": best, "ms_med": med}
        if flops:
            rec["tflops_best"] = flops / best / 1e9
        print(json.dumps(rec), flush=True)

    t("gram", lambda: e.build_gram(KE))
    def potrf():
        e.build_gram(KE)
        eng._lib.check(e.lib.distbo_potrf_f64(e.plan, eng._lib.ptr(e.K), eng._lib.ptr(e.W), eng._lib.ptr(e.info), eng._stream()), "potrf")
    t("gram+potrf", potrf, N**3 / 3)
    t("trtri", lambda: eng._lib.check(e.lib.distbo_trtri_f64(e.plan, eng._lib.ptr(e.K), eng._lib.ptr(e.W), eng._lib.ptr(e.Tm), eng._stream()), "trtri"), N**3 / 3)
    t("lauum", lambda: eng._lib.check(e.lib.distbo_lauum_f64(e.plan, eng._lib.ptr(e.W), eng._lib.ptr(e.Tm), eng._stream()), "lauum"), N**3 / 3)
    e.W_valid = True
    t("nll", lambda: e.likelihood())
    t("gram_grad", lambda: e.gram_gradient())
    def fit():
        e.build_gram(KE); e.factorize(True, True); e.likelihood(); e.gram_gradient()
    t("fit_total", fit, N**3)
    print(json.dumps({"what": "nll_value", "nll": e.read_loss()}), flush=True)
    e.kvec[:N] = 1.0
    cand = torch.from_numpy(rs.randn(M, d)).to(e.dev)
    t("score_M%d" % M, lambda: e.score(cand, acq="ei", y_max=1.0, xi=0.01, want_arrays=False), float(N) * N * M, iters=3)


if __name__ == "__main__":
    main()
