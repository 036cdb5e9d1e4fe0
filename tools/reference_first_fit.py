#!/usr/bin/env python
"""First target fit of BASELINE config 1 at the shipped settings, by the REFERENCE'S OWN code over tools/tf1_shim.py,
with a chosen number of host threads: every epoch's loss -> gpurun_out/selfdiv/ref_threads<k>.npz.

Purpose (profiles/r02_full_horizon_divergence.md): the training run is thousands of Adam steps on an ill-conditioned
likelihood (sigma^2 = 1e-5); rounding-level differences grow along it.  Running the unmodified reference twice with
nothing changed but the BLAS thread count (i.e. the summation order inside matmul / cholesky) shows how far the
reference tracks ITSELF, which bounds what any other implementation can be asked to reproduce.

    DISTBO_GOLD_THREADS=1 python tools/reference_first_fit.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import make_full_horizon_golden as G  # noqa: E402
import tf1_shim  # noqa: E402


def main():
    threads = int(os.environ.get("DISTBO_GOLD_THREADS", "4"))
    import torch
    torch.set_num_threads(threads)
    tf = tf1_shim.install()
    tf.reset_default_graph()
    tf1_shim._EAGER_SEED[0] = 0
    G.install_py3_shims()
    sys.path.insert(0, G.REF)
    sys.path.insert(0, os.path.join(G.REF, "distBO", "train"))
    import distBO.train.gp_regressor as ref_gp
    import distBO.model_embed.bayes_optimise as ref_bo_mod
    from sklearn.utils import check_random_state
    from distbo_b200 import experiments as X
    import json
    z = np.load(os.path.join(ROOT, "tests", "golden", "reffull_toy.npz"))
    args = X.parse_args(json.loads(str(z["argv"])) + ["/tmp/unused"])
    target, sources, _ = X.generate_data(args)          # bit-equal to the reference's generate_data (reftoy.npz)
    source_evals = [np.hstack([sp, se[:, None]]) for sp, se in zip(z["source_params"], z["source_evals"])]
    ref_bo_mod.np = G._NumpyCompat()
    logs = {}
    orig_fit, orig_close = ref_gp.gp_regressor.maximise_likhd, ref_gp.gp_regressor.close

    def recording_fit(self, *a, **kw):
        loss = orig_fit(self, *a, **kw)
        logs["loss_log"] = np.asarray(self.sess.loss_log)
        logs["params"] = {k: v.value.detach().numpy().copy() for k, v in self.net.params.items()}
        logs["init"] = {k: tf1_shim.INIT_VALUES[id(v)] for k, v in self.net.params.items()}
        return loss
    ref_gp.gp_regressor.maximise_likhd = recording_fit
    oc = X.optim_config_from(args)
    oc["n_warmup"] = 1000                                # the sweep after the fit is not what is measured here
    try:
        ref_bo_mod.bayes_opt_wrap('dist', target, acq='ei', model='toy', rs=check_random_state(args.opt_seed + 1),
                                  optim_config=oc, xi=0.01, kappa=2.58, bo_it=1, rand_it=0,
                                  init_list=source_evals, include_init=False, source_data=sources, acq_one_shot='lcb')
    finally:
        ref_gp.gp_regressor.maximise_likhd, ref_gp.gp_regressor.close = orig_fit, orig_close
    for k in z.files:                                    # same initial values as the golden run
        if k.startswith("init__"):
            assert np.array_equal(z[k], logs["init"][k[6:]]), k
    out = os.path.join(ROOT, "gpurun_out", "selfdiv")
    os.makedirs(out, exist_ok=True)
    np.savez_compressed(os.path.join(out, "ref_threads%d.npz" % threads), loss_log=logs["loss_log"],
                        **{"p__" + k: v for k, v in logs["params"].items()})
    print("threads %d: %d epochs, last loss %.10f (golden: %d epochs, %.10f)" % (
        threads, len(logs["loss_log"]), logs["loss_log"][-1], int(z["fit_epochs"][0]), float(z["fit_loss"][0])))


if __name__ == "__main__":
    main()
