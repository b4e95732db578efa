"""Golden vectors for the re-seeding passes, from the REFERENCE's own mem_collect_intv
(bwa_index/bwamem.c:114-162, reached through oracle/ref_collect_shim.c inside oracle/_ref/libbwaref.so).

Run in the build container only:   python tests/golden/make_reseed_golden.py
  reseed_golden.npz   reads (ragged), the sorted interval list of every read (start, end, x0, x2) with stock
                      parameters (split_factor 1.5, split_width 10, max_mem_intv 20) and with two other settings,
                      and the located seeds (bwt_sa with the mem_chain sampling rule, bwa_index/bwamem.c:278-283).
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402
from tools import synth  # noqa: E402

SETTINGS = [(1.5, 10, 20), (1.0, 3, 0), (2.0, 50, 5)]
GENOME_LEN, GENOME_SEED, SA_INTV, MAX_OCC = 80000, 5151, 16, 20


def make_reads(g):
    r1, _, _ = synth.make_reads(g, 300, 150, seed=199)
    r2, _, _ = synth.make_reads(g, 150, 150, seed=200, n_rate=0.01, sub_rate=0.03)
    r3, _, _ = synth.make_reads(g, 60, 250, seed=201, sub_rate=0.002)
    r4, _, _ = synth.make_reads(g, 60, 150, seed=202, sub_rate=0.0, ins_rate=0.0, del_rate=0.0)
    rnd = np.random.default_rng(6).integers(0, 4, size=(20, 150), dtype=np.uint8)
    reads = [x for blk in (r1, r2, r3, r4, rnd) for x in blk]
    reads += [np.full(40, 4, np.uint8), np.zeros(30, np.uint8), g[-60:].copy(), g[:19].copy(), g[100:118].copy()]
    rng = np.random.default_rng(7)
    for i in range(0, len(reads), 7):                       # ragged lengths
        reads[i] = reads[i][: int(rng.integers(20, reads[i].size + 1))]
    return reads


def main():
    assert O.have_ref(), "oracle/_ref missing: run oracle/build_ref.sh"
    tmp = tempfile.mkdtemp()
    g = synth.make_repeat_genome(GENOME_LEN, seed=GENOME_SEED)
    fa = os.path.join(tmp, "g.fa")
    synth.genome_to_fasta(g, fa)
    prefix = os.path.join(tmp, "g")
    O.ref_build_index(fa, prefix, SA_INTV)
    R = O.ref_lib()
    h = R.ref_load((prefix + ".bwt128").encode(), (prefix + ".sa").encode())
    assert h
    reads = make_reads(g)
    flat = np.concatenate(reads)
    off = np.zeros(len(reads) + 1, np.uint64)
    off[1:] = np.cumsum([r.size for r in reads])
    out = dict(genome_len=GENOME_LEN, genome_seed=GENOME_SEED, sa_intv=SA_INTV, max_occ=MAX_OCC, reads=flat, read_off=off,
               settings=np.array(SETTINGS, np.float64))
    for si, (sf, sw, mmi) in enumerate(SETTINGS):
        n_smems, iv = O.ref_collect_batch(h, flat, off, 19, sf, sw, mmi)
        out[f"n_smems_{si}"] = n_smems
        out[f"intv_{si}"] = iv[:, [3, 4, 0, 2]].copy()          # start, end, x0, x2
        if si == 0:
            rbeg, score, n_seeds = [], [], np.zeros(len(reads), np.uint32)
            j = 0
            for r in range(len(reads)):
                for _ in range(int(n_smems[r])):
                    x0, s = int(iv[j, 0]), int(iv[j, 2])
                    step = s // MAX_OCC if s > MAX_OCC else 1
                    k = cnt = 0
                    while k < s and cnt < MAX_OCC:
                        rbeg.append(R.ref_sa(h, x0 + k)); score.append(s if cnt == 0 else 0)
                        k += step; cnt += 1
                    n_seeds[r] += cnt
                    j += 1
            out["rbeg"] = np.array(rbeg, np.uint64); out["score"] = np.array(score, np.uint32); out["n_seeds"] = n_seeds
        print(f"setting {si}: {iv.shape[0]} intervals over {len(reads)} reads")
    np.savez_compressed(os.path.join(HERE, "reseed_golden.npz"), **out)


if __name__ == "__main__":
    main()
