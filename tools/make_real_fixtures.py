#!/usr/bin/env python
"""Builds small fixtures from the reference's data folders (run in the build container, where
/root/reference exists): tests/golden/protein_subset.npz and tests/golden/parkinson_subset.npz.
  protein   : the 7 ChEMBL tables, first 360 rows each, 166 MACCS bits packed + label
  parkinson : the first 10 patients, all rows, 21 float64 columns
The GPU box has no /root/reference; the tests read only these files."""
import os
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

files = sorted(os.listdir(os.path.join(REF, "protein")))
bits, labels, col0, names = [], [], [], []
for f in files:
    t = np.genfromtxt(os.path.join(REF, "protein", f), delimiter=",")[:360]
    assert set(np.unique(t[:, 1:-1])) <= {0.0, 1.0}
    bits.append(np.packbits(t[:, 1:-1].astype(np.uint8), axis=1))
    labels.append(t[:, -1].astype(np.uint8))
    names.append(f)
np.savez_compressed(os.path.join(OUT, "protein_subset.npz"), bits=np.stack(bits), labels=np.stack(labels),
                    names=np.array(names))
pats = [np.load(os.path.join(REF, "parkinson", "patient_%d.npy" % i)) for i in range(10)]
np.savez_compressed(os.path.join(OUT, "parkinson_subset.npz"), **{"patient_%d" % i: p for i, p in enumerate(pats)})
print("protein", np.stack(bits).shape, [int(l.sum()) for l in labels])
print("parkinson", [p.shape for p in pats])

# Full data sets of BASELINE configs 2 and 3 (the full-horizon replays, tools/replay_full_horizon.py, and the
# end-to-end bench workloads run on the GPU box, which has no /root/reference): every row of the 7 ChEMBL tables
# (166 bits packed into 21 bytes + label) and of the 42 patient tables.
full = {"names": np.array(files)}
for i, f in enumerate(files):
    t = np.genfromtxt(os.path.join(REF, "protein", f), delimiter=",")
    assert set(np.unique(t[:, 1:-1])) <= {0.0, 1.0} and set(np.unique(t[:, -1])) <= {0.0, 1.0}
    full["bits_%d" % i] = np.packbits(t[:, 1:-1].astype(np.uint8), axis=1)
    full["labels_%d" % i] = t[:, -1].astype(np.uint8)
np.savez_compressed(os.path.join(OUT, "protein_full.npz"), **full)
pats = [np.load(os.path.join(REF, "parkinson", "patient_%d.npy" % i)) for i in range(42)]
np.savez_compressed(os.path.join(OUT, "parkinson_full.npz"), **{"patient_%d" % i: p for i, p in enumerate(pats)})
print("protein_full", [full["bits_%d" % i].shape for i in range(len(files))])
print("parkinson_full", len(pats), sum(len(p) for p in pats))
