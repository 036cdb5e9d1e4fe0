"""GPU probe: per-stage device time of one loss+gradient+Adam step at the bench shape (eager launches,
CUDA events between stages), next to the whole step eager and as a CUDA graph."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from distbo_b200 import _lib  # noqa: E402
from distbo_b200.train import gp_regressor  # noqa: E402


def main():
    work = bench.make_workload(bench.N_OBS, 1024, bench.ROWS)
    gp = gp_regressor("dist", target_X=work["data_X"][64], target_Y=work["data_Y"][64], model_type="classification",
                      seed=0, target_embed_ind=np.arange(bench.ROWS), float64=False)
    gp.initial_weights = work["weights"]
    gp.maximise_likhd(work["theta"], work["y"][:, None], **bench.fit_kwargs(work, 3, 1e-2))
    e = gp.engine
    torch.cuda.synchronize()
    T = e.T
    names, evs = [], []

    def mark(name):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        names.append(name)
        evs.append(ev)

    for rep in range(3):
        names.clear(); evs.clear()
        gp._stage_inputs()
        mark("start")
        torch.index_select(gp._full[0], 0, gp._gather_dev, out=gp._bX)
        torch.index_select(gp._full[1], 0, gp._gather_dev, out=gp._bY)
        mark("gather")
        e.E[:, e.P - 1].copy_(gp._ratio)
        e.embed(gp._bX, gp._bY, gp._batch_off, out=e.E)
        mark("embed_fwd")
        KE = e.task_gram(e.E, T)
        mark("task_gram")
        e.build_gram(KE)
        mark("gram")
        _lib.check(e.lib.distbo_potrf_f64(e.plan, _lib.ptr(e.K), _lib.ptr(e.W), _lib.ptr(e.info),
                                          _lib.C.c_void_p(torch.cuda.current_stream().cuda_stream)), "potrf")
        mark("potrf")
        _lib.check(e.lib.distbo_trtri_f64(e.plan, _lib.ptr(e.K), _lib.ptr(e.W), _lib.ptr(e.Tm),
                                          _lib.C.c_void_p(torch.cuda.current_stream().cuda_stream)), "trtri")
        mark("trtri")
        _lib.check(e.lib.distbo_lauum_f64(e.plan, _lib.ptr(e.W), _lib.ptr(e.Tm),
                                          _lib.C.c_void_p(torch.cuda.current_stream().cuda_stream)), "lauum")
        mark("lauum")
        e.W_valid = True
        e.likelihood()
        mark("nll")
        e.gram_gradient()
        mark("gram_grad")
        dE = e.task_gram_backward(e.E, T)
        mark("task_gram_bwd")
        e.embed_backward(gp._bX, gp._bY, gp._batch_off, dE)
        mark("embed_bwd")
        e.adam_step(gp._lr, clip_norm=0.0, lr_t_dev=gp._lrt_dev)
        mark("adam")
        torch.cuda.synchronize()
        if rep == 2:
            stages = {names[i]: evs[i - 1].elapsed_time(evs[i]) for i in range(1, len(names))}
            stages["total"] = evs[0].elapsed_time(evs[-1])
            print(json.dumps({"what": "stages_ms", **stages}), flush=True)

    def timed(fn, n=5):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    gp._use_graph = False
    gp._graph = None
    print(json.dumps({"what": "epoch_eager_ms", "ms": timed(gp.run_epoch)}), flush=True)
    gp._use_graph = True
    gp._graph_warm = 1
    print(json.dumps({"what": "epoch_graph_ms", "ms": timed(gp.run_epoch)}), flush=True)
    os.environ["X"] = "1"

    def potrf_only():
        e.build_gram(e.KE)
        _lib.check(e.lib.distbo_potrf_f64(e.plan, _lib.ptr(e.K), _lib.ptr(e.W), _lib.ptr(e.info),
                                          _lib.C.c_void_p(torch.cuda.current_stream().cuda_stream)), "potrf")
    print(json.dumps({"what": "gram+potrf_eager_ms", "ms": timed(potrf_only)}), flush=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        potrf_only()
    print(json.dumps({"what": "gram+potrf_graph_ms", "ms": timed(g.replay)}), flush=True)
    print(json.dumps({"what": "info", "info": int(e.info.item())}), flush=True)
    # K1 at prediction time: fp32 embedding of the 65 full sample sets (218 MB of rows)
    X, Y, off = gp._prediction_sets()
    out = torch.zeros(e.T + 1, e.P, dtype=X.dtype, device=e.dev)
    ms = timed(lambda: e.embed(X, Y, off, out=out, dtype=X.dtype), n=10)
    nbytes = X.numel() * X.element_size() + Y.numel() * Y.element_size()
    print(json.dumps({"what": "embed_fwd_predict_sets", "dtype": str(X.dtype), "rows": int(X.shape[0]), "ms": ms,
                      "GB_per_s": nbytes / ms / 1e6}), flush=True)
    Xd, Yd = X.double(), Y.double()
    outd = torch.zeros(e.T + 1, e.P, dtype=torch.float64, device=e.dev)
    ms = timed(lambda: e.embed(Xd, Yd, off, out=outd, dtype=torch.float64), n=10)
    print(json.dumps({"what": "embed_fwd_predict_sets", "dtype": "float64", "rows": int(X.shape[0]), "ms": ms,
                      "GB_per_s": 2 * nbytes / ms / 1e6}), flush=True)


if __name__ == "__main__":
    main()
