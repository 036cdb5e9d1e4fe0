"""K1 probe: fp32 prediction-time embedding of the config-5 sample sets (65 x 5000 x 166 protein-like rows, 218 MB)
on the tensor-core kernel and on the generic kernel; CUDA-event time per call -> GB/s of algorithmic bytes."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distbo_b200 import engine as eng  # noqa: E402


def run(dx, dy, y_mode, T, rows, binary, reps=20):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    S = T * rows
    if binary:
        X = (torch.rand(S, dx, device=dev, generator=g) < 0.3).float()
    else:
        X = torch.randn(S, dx, device=dev, generator=g)
    if y_mode == 2:
        lab = (torch.rand(S, device=dev, generator=g) < 0.3).long()
        Y = torch.nn.functional.one_hot(lab, dy).float().contiguous()
    else:
        Y = torch.rand(S, max(dy, 1), device=dev, generator=g)
    net = eng.NetSpec(dx, Y.shape[1], 20, 10, 0, 0, y_mode)
    e = eng.GPEngine(2, n_embed=net.pooled_width, net=net)
    rs = np.random.RandomState(1)
    e.params.set("weights_x_1", rs.randn(dx, 20) * np.sqrt(2.0 / (dx + 20)))
    e.params.set("weights_x_2", rs.randn(20, 10) * np.sqrt(2.0 / 30))
    off = torch.arange(0, S + 1, rows, dtype=torch.int32, device=dev)
    out = {}
    res = {}
    for knob in ("1", "0"):
        os.environ["DISTBO_EMBED_TC"] = knob
        for _ in range(3):
            E = e.embed(X, Y, off, dtype=torch.float32)
        torch.cuda.synchronize()
        # one call captured as a CUDA graph, so that the host-side cost of the Python wrapper (~0.1 ms) is not
        # what the events see
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            E = e.embed(X, Y, off, dtype=torch.float32)
        graph.replay()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(reps):
            graph.replay()
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / reps
        nbytes = S * (dx + (Y.shape[1] if y_mode == 2 else 0)) * 4
        res[knob] = E.double().cpu().numpy()
        out["tensor_core" if knob == "1" else "generic"] = {"ms": ms, "GB_per_s": nbytes / ms / 1e6, "bytes": nbytes}
    out["max_rel_diff"] = float(np.max(np.abs(res["1"] - res["0"])) / np.max(np.abs(res["0"])))
    return out


if __name__ == "__main__":
    r = {"protein_like_65x5000x166": run(166, 2, 2, 65, 5000, True),
         "real_valued_65x5000x166": run(166, 2, 2, 65, 5000, False),
         "parkinson_like_none_65x5000x17": run(17, 1, 0, 65, 5000, False)}
    print(json.dumps(r))
