"""Golden vectors for ksw_align2: outputs of the REFERENCE's own function (src/ksw.c compiled into oracle/_ref/libforkksw.so) on the
seeded job sets of tests/test_sw_oracle.py.  Run where /root/reference exists:  python tests/golden/make_sw_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402
from tools import synth  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_sw_oracle import CASES  # noqa: E402

O.build_oracle()
out = {}
for name, jk, pk in CASES:
    if name == "overflow_byte":
        continue
    jobs = synth.make_sw_jobs(**jk)
    r = O.fork_sw_align2_batch(jobs, O.make_params(**pk))
    out[name] = np.stack([r[f] for f in r.dtype.names], axis=1).astype(np.int32)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sw_golden.npz"), **out)
print({k: v.shape for k, v in out.items()})
