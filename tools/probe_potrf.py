"""POTRF alone at N = 8192 (and 4096): device time per call (CUDA events, eager launches and graph replay) and
max |L - torch.linalg.cholesky| for the schedule selected by the environment (DISTBO_POTRF_FAR, DISTBO_POTRF_FARG,
DISTBO_POTRF_DOUBLE)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from distbo_b200 import _lib, engine as eng

out = {"env": {k: v for k, v in os.environ.items() if k.startswith("DISTBO_")}}
for N in [int(v) for v in os.environ.get("DISTBO_PROBE_SIZES", "8192,4096,3000").split(",")]:
    rs = np.random.RandomState(0)
    theta, y = rs.randn(N, 8), rs.randn(N)
    e = eng.GPEngine(8)
    e.params.set("log_sigma", 0.5 * np.log(1e-2))
    e.set_observations(theta, y)
    def potrf():
        st = _lib.C.c_void_p(torch.cuda.current_stream().cuda_stream)     # the capture stream inside a graph
        _lib.check(e.lib.distbo_potrf_f64(e.plan, _lib.ptr(e.K), _lib.ptr(e.W), _lib.ptr(e.info), st), "potrf")

    e.build_gram(None)
    Kfull = torch.tril(e.K.clone())
    Kfull = Kfull + torch.tril(Kfull, -1).T
    potrf()
    torch.cuda.synchronize()
    Lref = torch.linalg.cholesky(Kfull)
    err = float((torch.tril(e.K) - Lref).abs().max() / Lref.abs().max())
    times = []
    for rep in range(6):
        e.build_gram(None)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        potrf()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        e.build_gram(None)
        potrf()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    out[str(N)] = {"info": int(e.info.item()), "rel_err": err, "potrf_ms_min": min(times),
                   "gram_plus_potrf_graph_ms": a.elapsed_time(b) / 10}
print(json.dumps(out))
