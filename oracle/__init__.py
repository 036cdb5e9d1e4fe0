"""CPU oracle for the distBO joint-GP hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under ``distbo_b200/``
imports it: the product path is the sm_100a CUDA library and fails loudly without it.

PARITY PIN: the reference (hcllaw/distBO, Python 2 + TensorFlow 1.7) ships no tests, golden
vectors or recorded outputs for this path, and TensorFlow 1.7 cannot be installed in this image.
Its GP path is, however, plain Python over ~55 TensorFlow ops, so the reference's OWN SOURCES
(distBO/train/gp_regressor.py, train.py, log_gaus_likelihood(_feature).py, kernels.py,
distbo_net.py, feature_map.py, base.py, gp_utils.py, and the BO driver bayes_opt/*.py,
model_embed/bayes_optimise.py, models.py) are imported unmodified from /root/reference and
executed over ``tools/tf1_shim.py`` - a lazy TF-1.x op shim evaluated with torch CPU fp64 - by
``tools/make_reference_golden.py``.  The numbers they produce (loss per Adam epoch, fitted
parameters, posterior mean / std, similarities for 18 kernel / operator / surrogate
combinations, fp32 embeddings on the reference's own data, and 6 whole BO trajectories) are committed as ``tests/golden/ref_*.npz`` /
``refembed_f32.npz`` / ``refbo_*.npz``; ``tests/test_reference_golden.py`` checks that this oracle reproduces them
(1e-9; suggestions bit-equal) and that the CUDA path does (1e-8).  What the shim restates is
the semantics of single TF ops, not the reference's algorithm; the TF runtime itself remains
absent, which is the residual gap of this pin.  The restatement below follows the reference
source op-for-op (each function cites file:line) and is additionally cross-checked against
independent implementations (sklearn Matern(nu=1.5), scipy LAPACK, mpmath high precision,
torch autograd) in ``tests/test_oracle.py``.
"""
