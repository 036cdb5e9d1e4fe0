"""CPU oracle for the distBO joint-GP hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under ``distbo_b200/``
imports it: the product path is the sm_100a CUDA library and fails loudly without it.

PARITY UNPINNED: the reference (hcllaw/distBO, Python 2 + TensorFlow 1.7) ships no
tests, golden vectors or recorded outputs for this path and cannot be imported or run
in this image (no tensorflow, no python2).  The restatement below follows the reference
source op-for-op (each function cites file:line) and is cross-checked against
independent implementations (sklearn Matern(nu=1.5), scipy LAPACK, mpmath high
precision, torch autograd) in ``tests/test_oracle.py``; that is the strongest pin
available, and it is not the reference itself.
"""
