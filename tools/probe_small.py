"""GPU probe: per-epoch time of the training loop at the toy configuration's size (N = 465, 15 tasks x 500 rows,
1-d theta; experiments/one_dim_toy_config.py), where launch latency rather than arithmetic dominates."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distbo_b200.train import gp_regressor

rs = np.random.RandomState(0)
T, per, rows = 15, 31, 500
theta = rs.uniform(-8, 8, size=(T * per, 1)); y = np.exp(-0.5 * (theta - 1.0) ** 2)
X = [rs.randn(rows, 1) + t * 0.1 for t in range(T + 1)]; Y = [rs.randn(rows, 1) for _ in range(T + 1)]
gp = gp_regressor("dist", target_X=X[T], target_Y=Y[T], model_type="regression", seed=1, target_embed_ind=range(rows), float64=True)
kw = dict(data_X=X[:T], data_Y=Y[:T], source_embed_ind=[range(rows)] * T, embed_op="none", weight_dim_x=[20, 10],
          weight_dim_y=[20, 10], batch_size=500, sizes=[per] * T, learning_rate=0.005, likhd_var_init=1e-5,
          stratify=False, first_early_stop_epoch=10 ** 9)
gp.maximise_likhd(theta, y, max_epochs=20, **kw)
torch.cuda.synchronize()
for mode in ("graph", "eager"):
    os.environ["DISTBO_GRAPH"] = "1" if mode == "graph" else "0"
    t0 = time.perf_counter()
    gp.maximise_likhd(theta, y, max_epochs=300, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"what": "toy_epoch_with_loss_readback", "mode": mode, "us_per_epoch": dt / 300 * 1e6}), flush=True)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
os.environ["DISTBO_GRAPH"] = "1"
gp.maximise_likhd(theta, y, max_epochs=5, **kw)
a.record()
for _ in range(200):
    gp.run_epoch()
b.record(); torch.cuda.synchronize()
print(json.dumps({"what": "toy_epoch_device_only_graph", "us_per_epoch": a.elapsed_time(b) / 200 * 1e3}), flush=True)
cand = rs.uniform(-8, 8, size=(300000, 1))
t0 = time.perf_counter(); v, i = gp.acquire(cand, "ei", y_max=float(y.max()), xi=0.01); dt = time.perf_counter() - t0
t0 = time.perf_counter(); v, i = gp.acquire(cand, "ei", y_max=float(y.max()), xi=0.01); dt2 = time.perf_counter() - t0
print(json.dumps({"what": "toy_acquire_300k", "first_ms": dt * 1e3, "second_ms": dt2 * 1e3}), flush=True)
