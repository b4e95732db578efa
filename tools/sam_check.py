"""SAM-level comparison on BASELINE config 1 (10 k synthetic 150 bp reads vs a 5 Mb genome), or a smaller case.

Runs the reference's UNMODIFIED `gase_aln` driver (oracle/build_fork_driver.sh) linked three ways on the same index and reads
 - bwa-gasal2-b200  over libbwamem_b200.so (needs a GPU),
 - bwa-gasal2-cpu   over oracle/cpu_compat.cpp (the reference's bwt_smem1 / bwt_sa / ksw_extend2 on the CPU),
 - bwa-gasal2-ref   over the reference's own GPUSeed + GASAL2 compiled for sm_100 (needs a GPU; optional),
plus stock `bwa mem -r 100 -y 0` (oracle/_ref/bwa7: re-seeding passes 2-3 disabled from the command line) and compares
the SAM records.  The driver is the reference's program; nothing here is product code.

    python tools/sam_check.py [--reads 10000] [--genome 5000000] [--out DIR] [--no-gpu] [--threads 1]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")


def prepare(workdir, genome_bases, n_reads, read_len=150, repeats=False, n_threads=4, n_rate=0.0):
    """FASTA + every index file the driver opens: <prefix>.bwt (GPU layout) .sa .bwt128 (stock layout) from the repo's builder,
    .pac .ann .amb from the reference's own `bwa fa2pac`"""
    import importlib
    pkg = importlib.import_module("bwa-mem_gpu_b200")
    g = synth.make_genome(genome_bases, repeats=repeats)
    prefix = os.path.join(workdir, "ref.fa")
    synth.genome_to_fasta(g, prefix)
    pkg.build_index(g, prefix, sa_intv=16, also_stock_layout=True, n_threads=n_threads)
    subprocess.check_call([os.path.join(REF, "bwa7"), "fa2pac", "-f", prefix, prefix], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    reads, pos, strand = synth.make_reads(g, n_reads, read_len, n_rate=n_rate)
    fa = os.path.join(workdir, "reads.fa")
    synth.reads_to_fasta(reads, pos, strand, fa)
    return prefix, fa


def stock_prefix(prefix):
    """a second prefix whose .bwt is the stock-layout file, for the unmodified CPU `bwa mem` (same .sa / .pac / .ann / .amb)"""
    sp = prefix + ".stock"
    for ext, src in ((".bwt", ".bwt128"), (".sa", ".sa"), (".pac", ".pac"), (".ann", ".ann"), (".amb", ".amb")):
        if os.path.lexists(sp + ext):
            os.unlink(sp + ext)
        os.symlink(prefix + src, sp + ext)
    return sp


def run(cmd, out_path, cwd):
    t0 = time.time()
    with open(out_path, "wb") as fo, open(out_path + ".log", "wb") as fe:
        try:
            rc = subprocess.call(cmd, stdout=fo, stderr=fe, cwd=cwd, timeout=600)
        except subprocess.TimeoutExpired:
            rc = -999
    return rc, time.time() - t0


def sam_records(path):
    recs = {}
    order = []
    with open(path) as f:
        for line in f:
            if line.startswith("@"):
                continue
            t = line.rstrip("\n").split("\t")
            key = (t[0], int(t[1]) & 0x900)
            recs.setdefault(key, []).append(t)
            order.append(key)
    return recs, order


def classify(a, b):
    """a, b: SAM field lists of the same read (primary line).  Returns a short class name."""
    if a[:11] == b[:11] and sorted(a[11:]) == sorted(b[11:]):
        return "identical"
    if a[2] != b[2] or a[3] != b[3] or (int(a[1]) & 16) != (int(b[1]) & 16):
        return "position"
    if a[5] != b[5]:
        return "cigar"
    if a[4] != b[4]:
        return "mapq"
    ta, tb = {x[:2]: x for x in a[11:]}, {x[:2]: x for x in b[11:]}
    return "tags:" + "+".join(sorted(k for k in set(ta) | set(tb) if ta.get(k) != tb.get(k)))


def compare(pa, pb):
    ra, _ = sam_records(pa)
    rb, _ = sam_records(pb)
    cls = {}
    examples = {}
    for key in ra:
        if key not in rb:
            cls["missing"] = cls.get("missing", 0) + 1
            continue
        la, lb = ra[key], rb[key]
        if len(la) != len(lb):
            c = "n_records"
        else:
            c = "identical"
            for x, y in zip(la, lb):
                c = classify(x, y)
                if c != "identical":
                    break
        cls[c] = cls.get(c, 0) + 1
        if c != "identical" and c not in examples:
            examples[c] = ("\t".join(la[0][:9] + la[0][11:]), "\t".join(lb[0][:9] + lb[0][11:]))
    for key in rb:
        if key not in ra:
            cls["extra"] = cls.get("extra", 0) + 1
    return cls, examples


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=10_000)
    ap.add_argument("--genome", type=int, default=5_000_000)
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-gpu", action="store_true")
    ap.add_argument("--threads", type=int, default=1)
    ap.add_argument("--repeats", action="store_true")
    args = ap.parse_args()
    work = args.out or tempfile.mkdtemp()
    os.makedirs(work, exist_ok=True)
    prefix, fa = prepare(work, args.genome, args.reads, repeats=args.repeats)
    sp = stock_prefix(prefix)
    t = str(args.threads)
    runs = {"cpu": [os.path.join(REF, "bwa-gasal2-cpu"), "gase_aln", "-t", t, "-l", "150", prefix, fa],
            "stock": [os.path.join(REF, "bwa7p"), "mem", "-t", t, "-r", "100", "-y", "0", sp, fa],
            "stock_default": [os.path.join(REF, "bwa7p"), "mem", "-t", t, sp, fa]}
    if not args.no_gpu:
        runs["b200"] = [os.path.join(REF, "bwa-gasal2-b200"), "gase_aln", "-t", t, "-l", "150", prefix, fa]
        if os.path.exists(os.path.join(REF, "bwa-gasal2-ref")):
            runs["refgpu"] = [os.path.join(REF, "bwa-gasal2-ref"), "gase_aln", "-t", t, "-l", "150", prefix, fa]
    info = {"reads": args.reads, "genome": args.genome, "threads": args.threads, "runs": {}}
    for name, cmd in runs.items():
        rc, dt = run(cmd, os.path.join(work, name + ".sam"), work)
        info["runs"][name] = {"rc": rc, "seconds": round(dt, 2)}
    ok = lambda n: n in info["runs"] and info["runs"][n]["rc"] == 0  # noqa: E731
    pairs = [("b200", "cpu"), ("b200", "refgpu"), ("cpu", "stock"), ("b200", "stock"), ("stock", "stock_default")]
    info["compare"] = {}
    for a, b in pairs:
        if ok(a) and ok(b):
            cls, ex = compare(os.path.join(work, a + ".sam"), os.path.join(work, b + ".sam"))
            same_bytes = open(os.path.join(work, a + ".sam"), "rb").read().split(b"\n@PG")[0] == open(os.path.join(work, b + ".sam"), "rb").read().split(b"\n@PG")[0]
            body = lambda p: [ln for ln in open(p) if not ln.startswith("@PG")]  # noqa: E731
            info["compare"][f"{a}_vs_{b}"] = {"classes": cls, "examples": ex,
                                              "byte_identical_except_PG": body(os.path.join(work, a + ".sam")) == body(os.path.join(work, b + ".sam"))}
            del same_bytes
    print(json.dumps(info, indent=1))
    need = "b200_vs_cpu"
    if need in info["compare"]:
        sys.exit(0 if info["compare"][need]["byte_identical_except_PG"] else 1)


if __name__ == "__main__":
    main()
