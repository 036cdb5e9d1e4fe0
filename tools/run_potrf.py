import sys, numpy as np, torch
sys.path.insert(0, ".")
from distbo_b200 import engine as eng
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rs = np.random.RandomState(0)
theta = rs.randn(N, 8); y = rs.randn(N)
e = eng.GPEngine(8)
e.params.set("log_sigma", 0.5 * np.log(1e-2))
e.set_observations(theta, y)
for _ in range(reps):
    e.build_gram(None)
    e.factorize(need_inverse=False)
torch.cuda.synchronize()
print("info", int(e.info.item()))
