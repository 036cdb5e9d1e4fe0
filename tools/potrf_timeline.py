"""GPU debug tool: device-side timeline (globaltimer) of one distbo_potrf_f64 at N=8192.
Prints, per column block kb: panel start / potf2 start / inverse published / panel end, trailing-update
start / end, all in microseconds since the first panel started."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distbo_b200 import _lib, engine as eng

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rs = np.random.RandomState(0)
theta = rs.randn(N, 8); y = rs.randn(N)
e = eng.GPEngine(8)
e.params.set("log_sigma", 0.5 * np.log(1e-2))
e.set_observations(theta, y)
nb = e.n_pad // 128
st = _lib.C.c_void_p(torch.cuda.current_stream().cuda_stream)
def potrf():
    e.build_gram(None)
    _lib.check(e.lib.distbo_potrf_f64(e.plan, _lib.ptr(e.K), _lib.ptr(e.W), _lib.ptr(e.info), st), "potrf")
for _ in range(3):
    potrf()
torch.cuda.synchronize()
buf = torch.zeros(3 * nb, 4, dtype=torch.int64, device=e.dev)
buf[:, 0] = torch.iinfo(torch.int64).max     # atomicMin target (as unsigned: large)
e.lib.distbo_debug_timeline(_lib.ptr(buf))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e.build_gram(None); torch.cuda.synchronize()
a.record()
_lib.check(e.lib.distbo_potrf_f64(e.plan, _lib.ptr(e.K), _lib.ptr(e.W), _lib.ptr(e.info), st), "potrf")
b.record(); torch.cuda.synchronize()
e.lib.distbo_debug_timeline(None)
t = buf.cpu().numpy().astype(np.float64)
t0 = t[0, 0]
print("potrf %.3f ms (instrumented), info %d" % (a.elapsed_time(b), int(e.info.item())))
print("%3s | %9s %9s %9s %9s | %9s %9s | %7s %7s %7s | %9s %9s" % ("kb", "p.start", "potf2", "publish", "p.end", "r.start", "r.end", "upd", "potf2", "solve", "far.start", "far.end"))
for kb in range(nb):
    p = (t[2 * kb] - t0) / 1e3
    r = (t[2 * kb + 1] - t0) / 1e3
    has_r = t[2 * kb + 1, 1] > 0
    f = (t[2 * nb + kb] - t0) / 1e3
    has_f = t[2 * nb + kb, 1] > 0
    print("%3d | %9.1f %9.1f %9.1f %9.1f | %9s %9s | %7.1f %7.1f %7.1f | %9s %9s" % (
        kb, p[0], p[2], p[3], p[1], ("%9.1f" % r[0]) if has_r else "-", ("%9.1f" % r[1]) if has_r else "-",
        p[2] - p[0], p[3] - p[2], p[1] - p[3], ("%9.1f" % f[0]) if has_f else "-", ("%9.1f" % f[1]) if has_f else "-"))
