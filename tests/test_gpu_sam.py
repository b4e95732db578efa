"""SAM-level parity (north star: "the final SAM must be identical ..."): the reference's UNMODIFIED `gase_aln` driver
(src/fastmap.c:139-557, src/bwamem.c worker) linked over libbwamem_b200.so must write the same SAM, byte for byte, as the same
driver linked over the CPU checker oracle/cpu_compat.cpp (the reference's own bwt_smem1 / bwt_sa / ksw_extend2).  The binaries are
built by oracle/build_fork_driver.sh from the sources under /root/reference (here; they travel to the GPU box prebuilt).
The CPU test pins the checker itself against the reference's CPU program `bwa mem -r 100 -y 0` (oracle/_ref/bwa7p): same
position, strand, CIGAR, NM, MD and mapping quality on every read; AS / XS differ on a minority of reads because the fork extends
both sides of a seed from h0 = seed length and adds the two scores (src/bwamem.c:1356-1424, 2297-2302; SURVEY 8c caveat B)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import sam_check as SC  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
need = [os.path.join(REF, f) for f in ("bwa-gasal2-cpu", "bwa7p", "libbwaref.so")]
have_cpu = all(os.path.exists(p) for p in need)
have_b200 = os.path.exists(os.path.join(REF, "bwa-gasal2-b200"))


def body(path):
    return [ln for ln in open(path) if not ln.startswith("@PG")]


@pytest.mark.skipif(not have_cpu, reason="oracle/_ref driver binaries not built (no /root/reference)")
def test_cpu_checker_driver_vs_reference_bwa_mem(pkg, tmp_path):
    work = str(tmp_path)
    prefix, fa = SC.prepare(work, 600_000, 2500, repeats=True, n_rate=0.001)
    sp = SC.stock_prefix(prefix)
    rc, _ = SC.run([need[0], "gase_aln", "-t", "2", "-l", "150", prefix, fa], os.path.join(work, "cpu.sam"), work)
    assert rc == 0, open(os.path.join(work, "cpu.sam.log")).read()[-2000:]
    rc, _ = SC.run([need[1], "mem", "-t", "2", "-r", "100", "-y", "0", sp, fa], os.path.join(work, "stock.sam"), work)
    assert rc == 0, open(os.path.join(work, "stock.sam.log")).read()[-2000:]
    cls, ex = SC.compare(os.path.join(work, "cpu.sam"), os.path.join(work, "stock.sam"))
    n = sum(cls.values())
    assert n == 2500
    # every difference is in the score tags only; positions, CIGARs, NM/MD and mapping qualities agree
    assert set(cls) <= {"identical", "tags:AS", "tags:XS", "tags:AS+XS"}, (cls, ex)
    assert cls["identical"] >= 0.85 * n, cls


@pytest.mark.gpu
@pytest.mark.parametrize("threads,n_reads,genome,repeats", [(1, 3000, 1_000_000, True), (3, 10_000, 5_000_000, False)])
def test_b200_driver_sam_identical_to_cpu_checker(pkg, tmp_path, threads, n_reads, genome, repeats):
    assert have_cpu and have_b200, "oracle/_ref/bwa-gasal2-{b200,cpu} must travel to the GPU box (oracle/build_fork_driver.sh)"
    work = str(tmp_path)
    prefix, fa = SC.prepare(work, genome, n_reads, repeats=repeats, n_rate=0.001 if repeats else 0.0)
    rc, _ = SC.run([os.path.join(REF, "bwa-gasal2-b200"), "gase_aln", "-t", str(threads), "-l", "150", prefix, fa], os.path.join(work, "b200.sam"), work)
    assert rc == 0, open(os.path.join(work, "b200.sam.log")).read()[-3000:]
    rc, _ = SC.run([need[0], "gase_aln", "-t", str(threads), "-l", "150", prefix, fa], os.path.join(work, "cpu.sam"), work)
    assert rc == 0, open(os.path.join(work, "cpu.sam.log")).read()[-3000:]
    a, b = body(os.path.join(work, "b200.sam")), body(os.path.join(work, "cpu.sam"))
    assert len(a) == len(b) and len([x for x in a if not x.startswith("@")]) >= n_reads
    diff = [(x, y) for x, y in zip(a, b) if x != y]
    assert not diff, diff[:3]
