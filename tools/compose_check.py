"""reads -> regions -> finished regions -> alignments, all three device stages composed (bwa_b200_align_host ->
bwa_b200_finish_regions_host -> bwa_b200_reg2aln_host) on real reads, every stage compared with the oracle on the same records.
Each stage has its own GPU parity tests; this is the composition.  NOT YET RUN ON A GPU BOX when it was written (round 1 ended without
GPU minutes): run it in the first GPU visit of the next round and promote it to tests/ once green.

    python tools/compose_check.py [n_reads=2000]
"""
import importlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("bwa-mem_gpu_b200")
from oracle import chain_py as CP, oracle_py as O, region_py as RP  # noqa: E402  (checkers only)
from tools import synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    g = synth.make_genome(200_000, repeats=True)
    prefix = os.path.join(tempfile.mkdtemp(), "g")
    pkg.build_index(g, prefix, sa_intv=16, n_threads=4)
    idx = pkg.Index.load(prefix + ".bwt", prefix + ".sa", 0)
    idx.attach_ref(g)
    reads, _, _ = synth.make_reads(g, n, 150, seed=77, sub_rate=0.03, ins_rate=0.004, del_rate=0.004)
    L = reads.shape[1]
    packed, woff, rl = pkg.pack_codes(reads.reshape(-1).copy(), (np.arange(n + 1) * L).astype(np.uint64))
    ctg = CP.Contigs((g.size,))
    # stage 1: regions
    al = pkg.Aligner(idx, n, packed.size)
    res = al.align_host(packed, woff, rl, pkg.seed_params(19, 500, True), pkg.chain_params(max_occ=500, w=100), pkg.ext_params(w=100, zdrop=100, use_band=1))
    al.destroy()
    regs = res["regions"]
    read_of = np.repeat(np.arange(n), res["n_regions"])
    keep = (regs["qe"] > regs["qb"]) & (regs["re"] > regs["rb"])
    recs = pkg.alnregs_from_regions(regs[keep])
    per_read = np.bincount(read_of[keep], minlength=n)
    off = np.concatenate([[0], np.cumsum(per_read)]).astype(np.uint64)
    # stage 2: finished regions, against the oracle on the same records
    fin, n_pri = pkg.finish_regions(idx, packed, woff, rl, recs, off, pkg.region_opt(), ctg_alt=ctg.alt, first_read_id=0)
    opt = RP.default_opt()
    bad_fin = 0
    for i in range(n):
        want, wp = RP.oracle_finish(opt, ctg, g, reads[i], recs[int(off[i]):int(off[i + 1])], i)
        bad_fin += int(not (int(n_pri[i]) == wp and RP.equal(fin[i], want)))
    # stage 3: alignments of the primary regions, against the oracle's mem_reg2aln
    rows = [(i, r) for i in range(n) for r in fin[i] if r["secondary"] < 0]
    alns = np.zeros(len(rows), pkg.ALN_IN_DTYPE)
    for k, (i, r) in enumerate(rows):
        alns[k] = (i, r["qb"], r["qe"], r["rb"], r["re"], r["truesc"], r["w"])
    cg = pkg.Cigar(0)
    got, flat = cg.reg2aln_host(idx, ctg.off, packed, woff, rl, alns, pkg.ext_params(w=100), 1)
    cg.destroy()
    idx.free()
    copt, kp = CP.default_opt(w=100), O.make_params()
    bad_aln = 0
    for k, (i, r) in enumerate(rows):
        want, wc = CP.oracle_reg2aln(copt, kp, ctg, g, reads[i], int(r["qb"]), int(r["qe"]), int(r["rb"]), int(r["re"]), int(r["truesc"]), int(r["w"]))
        o = got[k]
        mine = flat[int(o["cigar_off"]):int(o["cigar_off"]) + int(o["n_cigar"])]
        ok = (int(o["pos"]), int(o["rid"]), int(o["is_rev"]), int(o["nm"])) == (int(want["pos"]), int(want["rid"]), int(want["is_rev"]), int(want["nm"])) \
            and mine.size == wc.size and bool((mine == wc).all())
        bad_aln += int(not ok)
    info = dict(reads=n, regions=int(keep.sum()), finished=int(sum(len(f) for f in fin)), primary=len(rows), finish_mismatches=bad_fin,
                reg2aln_mismatches=bad_aln, mapq60=int(sum(int((f["mapq"] == 60).sum()) for f in fin)))
    print(json.dumps(info))
    sys.exit(0 if bad_fin == 0 and bad_aln == 0 else 1)


if __name__ == "__main__":
    main()
