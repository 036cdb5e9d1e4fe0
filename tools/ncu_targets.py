"""Small driver for ncu captures: the bench-shaped problem (N=8192, 64 tasks, protein-like datasets), ONE eager
training step, ONE fit->score hand-off (fp32 embedding of the 65 full sample sets) and ONE scoring chunk."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DISTBO_GRAPH"] = "0"
import bench  # noqa: E402
from distbo_b200.train import gp_regressor  # noqa: E402

work = bench.make_workload(bench.N_OBS, 18944, bench.ROWS)
gp = gp_regressor("dist", target_X=work["data_X"][64], target_Y=work["data_Y"][64], model_type="classification",
                  seed=0, target_embed_ind=np.arange(bench.ROWS), float64=False)
gp.initial_weights = work["weights"]
loss = gp.maximise_likhd(work["theta"], work["y"][:, None], **bench.fit_kwargs(work, 2, 1e-2))
v, i = gp.acquire(work["cand"], "ei", y_max=work["y_max"], xi=0.01)
torch.cuda.synchronize()
print("loss %.6f best %d %.6g" % (loss, i, v))
