#!/usr/bin/env python
"""Randomised stress of the re-seeding kernels against the oracle (GPU box): random thresholds, seed lengths, read shapes, genomes with
repeats, narrow and wide rows.  Exits non-zero on the first mismatch."""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
from oracle import oracle_py as O
from tools import synth

pkg = ge.load_package(); pkg.build()
rng = np.random.default_rng(4321)
tmp = tempfile.mkdtemp()
bad = 0
for gi in range(3):
    g = synth.make_genome(int(rng.choice([60_000, 250_000, 1_000_000])), seed=50 + gi, repeats=bool(gi % 2 == 0))
    prefix = os.path.join(tmp, f"g{gi}")
    pkg.build_index(g, prefix, sa_intv=int(rng.choice([8, 16, 32])), n_threads=8)
    oi = O.OracleIndex(prefix + ".bwt", prefix + ".sa")
    for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
        if it % 3 == 2:
            os.environ["BWA_B200_WIDE_ROWS"] = "1"
        else:
            os.environ.pop("BWA_B200_WIDE_ROWS", None)
        idx = pkg.Index.load(prefix + ".bwt", prefix + ".sa", 0)
        L = int(rng.choice([76, 101, 150, 250]))
        reads, _, _ = synth.make_reads(g, 2500, L, seed=900 + 10 * gi + it, sub_rate=float(rng.choice([0.0, 0.01, 0.04])), n_rate=float(rng.choice([0.0, 0.004])))
        flat = reads.reshape(-1).copy(); off = (np.arange(2501) * L).astype(np.uint64)
        msl = int(rng.choice([12, 19, 25])); mo = int(rng.choice([3, 50, 500]))
        sf, sw, mmi = float(rng.choice([1.0, 1.5, 2.5])), int(rng.choice([1, 10, 100])), int(rng.choice([0, 5, 20, 200]))
        packed, woff, rl = pkg.pack_codes(flat, off)
        sd = pkg.Seeder(idx, 2500, packed.size)
        got = sd.seed_host(packed, woff, rl, params=pkg.seed_params(msl, mo, True, sf, sw, mmi))
        sm = sd.smems(2500, int(flat.size) * 8)
        sd.destroy(); idx.free()
        rs = O.reseed(sf, sw, mmi)
        wsm = oi.smem_batch(flat, off, msl, rs=rs, cap=int(flat.size) * 8)
        want = oi.seed_batch(flat, off, msl, mo, n_threads=8, rs=rs)
        ok = (sm["n_smems"] == wsm["n_smems"]).all() and all((sm[k] == wsm[k]).all() for k in ("qbeg", "qend", "k", "s")) and \
            got["total"] == want["total"] and (got["rbeg"] == want["rbeg"]).all() and (got["score"] == want["score"]).all() and \
            (got["qq"][:, 0] == want["qbeg"]).all() and (got["qq"][:, 1] == want["qend"]).all()
        print(f"genome {gi} it {it} L={L} min_seed={msl} max_occ={mo} split=({sf},{sw}) mmi={mmi} wide={int(it % 3 == 2)} smems={int(sm['n_smems'].sum())} "
              f"max/read={int(sm['n_smems'].max())} {'ok' if ok else 'MISMATCH'}", flush=True)
        bad += not ok
    oi.close()
sys.exit(1 if bad else 0)
