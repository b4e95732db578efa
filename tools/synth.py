"""Deterministic synthetic inputs: genomes, reads, extension jobs.

Everything is produced from numpy's PCG64 with fixed seeds, so the CPU oracle, the
reference arm and the CUDA path all see byte-identical inputs on any machine.
Base codes follow the reference: A=0 C=1 G=2 T=3 N=4 (nst_nt4_table, src/bntseq.c).
"""
from __future__ import annotations

import numpy as np

GENOME_SEED = 20261017
READS_SEED = 20261018
EXT_SEED = 777


def make_genome(n_bases: int, seed: int = GENOME_SEED, repeats: bool = False) -> np.ndarray:
    """i.i.d. uniform ACGT genome as uint8 codes.  `repeats=True` overwrites ~5 % of it with
    diverged copies of a few elements (stress profile for the max_occ path)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    g = rng.integers(0, 4, size=n_bases, dtype=np.uint8)
    if repeats and n_bases >= 4000:
        n_el = 4
        target = n_bases // 20
        placed = 0
        els = [rng.integers(0, 4, size=int(rng.integers(300, min(6000, n_bases // 8))), dtype=np.uint8) for _ in range(n_el)]
        while placed < target:
            e = els[int(rng.integers(0, n_el))].copy()
            mut = rng.random(e.size) < 0.10
            e[mut] = (e[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) & 3
            p = int(rng.integers(0, n_bases - e.size))
            g[p:p + e.size] = e
            placed += e.size
    return g


def make_repeat_genome(n_bases: int, unit_len: int = 1000, copies: int = 30, div: float = 0.02, seed: int = 7) -> np.ndarray:
    """random genome whose second half holds `copies` diverged tandem copies of one unit:
    SMEMs inside the unit have tens of occurrences (exercises the max_occ sampling rule)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    g = rng.integers(0, 4, size=n_bases, dtype=np.uint8)
    unit = rng.integers(0, 4, size=unit_len, dtype=np.uint8)
    p = n_bases - copies * unit_len
    assert p > 0
    for _ in range(copies):
        e = unit.copy()
        mut = rng.random(unit_len) < div
        e[mut] = (e[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) & 3
        g[p:p + unit_len] = e
        p += unit_len
    return g


def revcomp(codes: np.ndarray) -> np.ndarray:
    out = codes[..., ::-1].copy()
    m = out < 4
    out[m] = 3 - out[m]
    return out


def make_reads(genome: np.ndarray, n_reads: int, read_len: int = 150, seed: int = READS_SEED,
               sub_rate: float = 0.01, ins_rate: float = 0.0005, del_rate: float = 0.0005,
               n_rate: float = 0.0):
    """Returns (reads[n_reads, read_len] uint8, pos[n_reads] int64, strand[n_reads] uint8).
    Odd read index = reverse-complement strand; substitutions go to a different base;
    indels have length 1."""
    rng = np.random.Generator(np.random.PCG64(seed))
    G = genome.size
    pad = 16
    pos = rng.integers(0, G - read_len - pad, size=n_reads, dtype=np.int64)
    idx = pos[:, None] + np.arange(read_len + pad, dtype=np.int64)[None, :]
    raw = genome[idx]                                    # n x (L+pad)
    sub = rng.random((n_reads, read_len + pad)) < sub_rate
    shift = rng.integers(1, 4, size=(n_reads, read_len + pad), dtype=np.uint8)
    raw = np.where(sub, (raw + shift) & 3, raw).astype(np.uint8)
    ins = rng.random((n_reads, read_len)) < ins_rate
    dele = rng.random((n_reads, read_len)) < del_rate
    ins_base = rng.integers(0, 4, size=(n_reads, read_len), dtype=np.uint8)
    reads = raw[:, :read_len].copy()
    rows = np.nonzero(ins.any(axis=1) | dele.any(axis=1))[0]
    for r in rows:
        src = raw[r]
        out = []
        j = 0
        while len(out) < read_len and j < src.size:
            if j < read_len and ins[r, j]:
                out.append(ins_base[r, j])
                if len(out) >= read_len:
                    break
            if j < read_len and dele[r, j]:
                j += 1
                continue
            out.append(src[j])
            j += 1
        while len(out) < read_len:
            out.append(0)
        reads[r] = np.asarray(out[:read_len], dtype=np.uint8)
    if n_rate > 0:
        nm = rng.random((n_reads, read_len)) < n_rate
        reads[nm] = 4
    strand = (np.arange(n_reads) & 1).astype(np.uint8)
    odd = strand == 1
    reads[odd] = revcomp(reads[odd])
    return reads, pos, strand


def reads_to_fasta(reads: np.ndarray, pos: np.ndarray, strand: np.ndarray, path: str) -> None:
    """single-line FASTA, one read per line (boundary requirement of seed_gpu, seed_gen.cu:1708)."""
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    with open(path, "wb") as fh:
        for i in range(reads.shape[0]):
            fh.write(b">r%d_%d_%d\n" % (i, int(pos[i]), int(strand[i])))
            fh.write(lut[reads[i]].tobytes())
            fh.write(b"\n")


def genome_to_fasta(genome: np.ndarray, path: str, name: str = "chr1", width: int = 60) -> None:
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    txt = lut[genome]
    with open(path, "wb") as fh:
        fh.write(b">" + name.encode() + b"\n")
        for i in range(0, txt.size, width):
            fh.write(txt[i:i + width].tobytes())
            fh.write(b"\n")


def make_ext_jobs(n_jobs: int, qlen_choices=(100, 150, 200, 250, 300), w: int = 100, seed: int = EXT_SEED,
                  sub_rate: float = 0.05, indel_rate: float = 0.01, n_job_frac: float = 0.01,
                  h0_range=(19, 150), pad8: bool = True, qlen_range=None, tail_random: bool = True):
    """Extension-only sweep jobs (C4): target = query with `sub_rate` substitutions and
    `indel_rate` indels, then a random tail up to tlen = qlen + min(qlen, 2w).
    Returns a dict of GASAL-style host arrays: byte codes 0..4 with per-job offsets (each
    sequence padded to a multiple of 8 with code 4 when pad8), lengths and h0."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if qlen_range is not None:
        qlens = rng.integers(qlen_range[0], qlen_range[1] + 1, size=n_jobs)
    else:
        qlens = rng.choice(np.asarray(qlen_choices), size=n_jobs)
    tlens = qlens + np.minimum(qlens, 2 * w)
    h0 = rng.integers(h0_range[0], h0_range[1] + 1, size=n_jobs).astype(np.uint32)

    def padded(n):
        return (n + 7) // 8 * 8 if pad8 else n

    qoff = np.zeros(n_jobs, dtype=np.uint32)
    toff = np.zeros(n_jobs, dtype=np.uint32)
    qp = np.asarray([padded(int(x)) for x in qlens], dtype=np.int64)
    tp = np.asarray([padded(int(x)) for x in tlens], dtype=np.int64)
    qoff[1:] = np.cumsum(qp)[:-1]
    toff[1:] = np.cumsum(tp)[:-1]
    qseq = np.full(int(qp.sum()), 4, dtype=np.uint8)
    tseq = np.full(int(tp.sum()), 4, dtype=np.uint8)
    for a in range(n_jobs):
        ql, tl = int(qlens[a]), int(tlens[a])
        q = rng.integers(0, 4, size=ql, dtype=np.uint8)
        t = q.copy()
        sm = rng.random(ql) < sub_rate
        t[sm] = (t[sm] + rng.integers(1, 4, size=int(sm.sum()), dtype=np.uint8)) & 3
        ev = rng.random(ql) < indel_rate
        if ev.any():
            parts = []
            last = 0
            for p in np.nonzero(ev)[0]:
                parts.append(t[last:p])
                if rng.random() < 0.5:
                    parts.append(rng.integers(0, 4, size=int(rng.integers(1, 4)), dtype=np.uint8))  # insertion in target
                    last = p
                else:
                    last = min(ql, p + int(rng.integers(1, 4)))                                     # deletion from target
            parts.append(t[last:])
            t = np.concatenate(parts)
        if t.size < tl:
            tail = rng.integers(0, 4, size=tl - t.size, dtype=np.uint8) if tail_random else np.zeros(tl - t.size, np.uint8)
            t = np.concatenate([t, tail])
        t = t[:tl]
        if rng.random() < n_job_frac:
            q[int(rng.integers(0, ql))] = 4
            t[int(rng.integers(0, tl))] = 4
        qseq[qoff[a]:qoff[a] + ql] = q
        tseq[toff[a]:toff[a] + tl] = t
    return dict(qseq=qseq, tseq=tseq, qoff=qoff, toff=toff,
                qlen=qlens.astype(np.uint32), tlen=tlens.astype(np.uint32), h0=h0)


def make_global_jobs(n_jobs: int, qlen_range=(30, 150), seed: int = 991, sub_rate: float = 0.03, indel_rate: float = 0.01,
                     w_extra=(0, 0), n_frac: float = 0.01, w_cap: int = 100):
    """End-to-end alignment jobs for ksw_global2 / bwa_gen_cigar2: target = query with substitutions and short indels (no
    tail: both ends are fixed).  Band per job = |tlen - qlen| + 3 + U[w_extra], capped at w_cap but never below the
    feasible minimum |tlen - qlen| + 1.  Byte codes 0..4, sequences padded to a multiple of 8 with code 4."""
    rng = np.random.Generator(np.random.PCG64(seed))
    qs, ts = [], []
    for _ in range(n_jobs):
        ql = int(rng.integers(qlen_range[0], qlen_range[1] + 1))
        q = rng.integers(0, 4, size=ql, dtype=np.uint8)
        if rng.random() < n_frac:
            q[int(rng.integers(0, ql))] = 4
        t = q.copy()
        sm = rng.random(ql) < sub_rate
        t[sm] = (t[sm] + rng.integers(1, 4, size=int(sm.sum()), dtype=np.uint8)) & 3
        ev = np.nonzero(rng.random(ql) < indel_rate)[0]
        if ev.size:
            parts, last = [], 0
            for p in ev:
                parts.append(t[last:p])
                if rng.random() < 0.5:
                    parts.append(rng.integers(0, 4, size=int(rng.integers(1, 5)), dtype=np.uint8)); last = p
                else:
                    last = min(ql, p + int(rng.integers(1, 5)))
            parts.append(t[last:])
            t = np.concatenate(parts)
        if t.size == 0:
            t = q[:1].copy()
        qs.append(q); ts.append(t)
    qlen = np.array([x.size for x in qs], np.uint32); tlen = np.array([x.size for x in ts], np.uint32)
    qp = (qlen.astype(np.int64) + 7) // 8 * 8; tp = (tlen.astype(np.int64) + 7) // 8 * 8
    qoff = np.zeros(n_jobs, np.uint32); toff = np.zeros(n_jobs, np.uint32)
    qoff[1:] = np.cumsum(qp)[:-1]; toff[1:] = np.cumsum(tp)[:-1]
    qseq = np.full(int(qp.sum()), 4, np.uint8); tseq = np.full(int(tp.sum()), 4, np.uint8)
    for a in range(n_jobs):
        qseq[qoff[a]:qoff[a] + qlen[a]] = qs[a]; tseq[toff[a]:toff[a] + tlen[a]] = ts[a]
    diff = np.abs(tlen.astype(np.int64) - qlen.astype(np.int64))
    w = diff + 3 + rng.integers(w_extra[0], w_extra[1] + 1, size=n_jobs)
    w = np.maximum(np.minimum(w, w_cap), diff + 1).astype(np.uint32)
    return dict(qseq=qseq, tseq=tseq, qoff=qoff, toff=toff, qlen=qlen, tlen=tlen, w=w)


def make_sw_jobs(n_jobs: int, qlen_range=(30, 150), tlen_range=(100, 700), seed: int = 4242, sub_rate: float = 0.04, indel_rate: float = 0.01,
                 none_frac: float = 0.1, n_frac: float = 0.05, xtra=None):
    """Local-alignment jobs shaped like mate rescue (mem_matesw, src/bwamem_pair.c:119-175): a query and a longer target window that
    holds a diverged copy of it (or of a part of it, or nothing: none_frac) somewhere inside; codes 0..4.  Byte-per-base buffers with
    per-job offsets, like make_ext_jobs.  xtra: per-job flags of ksw_align2 (array or scalar); default = the mate-rescue value
    KSW_XSUBO | KSW_XSTART | (qlen < 250 ? KSW_XBYTE : 0) | 19."""
    rng = np.random.Generator(np.random.PCG64(seed))
    qs, ts, qoff, toff, qlen, tlen = [], [], [], [], [], []
    qo = to = 0
    for _ in range(n_jobs):
        ql = int(rng.integers(qlen_range[0], qlen_range[1] + 1))
        tl = int(rng.integers(max(tlen_range[0], 1), tlen_range[1] + 1))
        q = rng.integers(0, 4, size=ql, dtype=np.uint8)
        t = rng.integers(0, 4, size=tl, dtype=np.uint8)
        if rng.random() >= none_frac:
            a = int(rng.integers(0, max(1, ql // 3)))
            b = int(rng.integers(min(ql, a + 10), ql + 1))
            part = q[a:b].copy()
            mut = rng.random(part.size) < sub_rate
            part[mut] = (part[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) & 3
            out = []
            for c in part:
                u = rng.random()
                if u < indel_rate / 2:
                    continue
                out.append(c)
                if u > 1 - indel_rate / 2:
                    out.extend(rng.integers(0, 4, size=int(rng.integers(1, 4)), dtype=np.uint8))
            part = np.asarray(out, dtype=np.uint8)
            if 0 < part.size <= tl:
                p = int(rng.integers(0, tl - part.size + 1))
                t[p:p + part.size] = part
        if rng.random() < n_frac:
            q[rng.integers(0, ql, size=max(1, ql // 30))] = 4
        if rng.random() < n_frac:
            t[rng.integers(0, tl, size=max(1, tl // 40))] = 4
        qs.append(q); ts.append(t); qoff.append(qo); toff.append(to); qlen.append(ql); tlen.append(tl)
        qo += ql; to += tl
    ql_a = np.asarray(qlen, np.uint32)
    if xtra is None:
        x = (0x40000 | 0x80000 | 19) | np.where(ql_a < 250, 0x10000, 0).astype(np.uint32)
    else:
        x = np.broadcast_to(np.asarray(xtra, np.uint32), ql_a.shape).copy()
    return dict(qseq=np.concatenate(qs), tseq=np.concatenate(ts), qoff=np.asarray(qoff, np.uint32), toff=np.asarray(toff, np.uint32),
                qlen=ql_a, tlen=np.asarray(tlen, np.uint32), xtra=x.astype(np.uint32))


def make_flank_jobs(n_jobs: int, seed: int = 515, qlen_range=(1, 160), h0_range=(1, 200), w: int = 100, pad8: bool = True):
    """Extension jobs as a maximal exact match leaves them in a read with few differences: the query equals the head of the target,
    mostly with its first base changed (the base that ended the match) and zero to six further substitutions, a third of them over
    a tandem repeat (where a shifted diagonal matches as well as the main one).  Many jobs miss the closed-form shape by one detail -- a
    third difference, adjacent differences, an N on either side, a target shorter than the query, a small h0 -- so that the shortcut is
    tested on both sides of every condition.  Same dict as make_ext_jobs."""
    rng = np.random.Generator(np.random.PCG64(seed))
    qlens = rng.integers(qlen_range[0], qlen_range[1] + 1, size=n_jobs)
    h0 = rng.integers(h0_range[0], h0_range[1] + 1, size=n_jobs).astype(np.uint32)
    qs, ts = [], []
    for a in range(n_jobs):
        ql = int(qlens[a])
        tl = ql + int(rng.integers(0, min(ql, 2 * w) + 1))
        t = rng.integers(0, 4, size=tl, dtype=np.uint8)
        q = t[:ql].copy()
        if rng.random() < 0.35 and tl > 8:                       # a tandem repeat (period 1 .. 12) somewhere: shifted diagonals match there
            per = int(rng.integers(1, 13)); at = int(rng.integers(0, tl - 4)); ln = int(rng.integers(4, 120))
            unit = t[at:at + per].copy()
            for x in range(at, min(tl, at + ln)):
                t[x] = unit[(x - at) % per]
            q = t[:ql].copy()
        u = rng.random()
        if u < 0.85:
            q[0] = (q[0] + int(rng.integers(1, 4))) & 3          # the mismatch that ended the seed
        for _ in range(int(rng.choice([0, 0, 1, 1, 1, 2, 2, 3, 3, 4, 5, 6]))):   # further substitutions, anywhere (next to the first one included)
            if ql > 1:
                k = int(rng.integers(1, ql)) if rng.random() < 0.8 else min(ql - 1, int(rng.integers(1, 8)))
                q[k] = (q[k] + int(rng.integers(1, 4))) & 3
        v = rng.random()
        if v < 0.04 and ql > 1:                                   # one more difference somewhere
            k = int(rng.integers(1, ql)); q[k] = (q[k] + 1) & 3
        elif v < 0.13:                                            # N in the query / in the compared part of the target / beyond it
            q[int(rng.integers(0, ql))] = 4
        elif v < 0.18:
            t[int(rng.integers(0, ql))] = 4
        elif v < 0.21 and tl > ql:
            t[int(rng.integers(ql, tl))] = 4
        elif v < 0.26 and ql > 1:                                 # target shorter than the query
            t = t[:int(rng.integers(1, ql))]
        elif v < 0.31:
            h0[a] = int(rng.integers(1, 8))
        qs.append(q); ts.append(t)
    tlens = np.array([len(t) for t in ts], np.int64)

    def padded(n):
        return (n + 7) // 8 * 8 if pad8 else n
    qp = np.asarray([padded(int(x)) for x in qlens], dtype=np.int64)
    tp = np.asarray([padded(int(x)) for x in tlens], dtype=np.int64)
    qoff = np.zeros(n_jobs, dtype=np.uint32); toff = np.zeros(n_jobs, dtype=np.uint32)
    qoff[1:] = np.cumsum(qp)[:-1]; toff[1:] = np.cumsum(tp)[:-1]
    qseq = np.full(int(qp.sum()) + 8, 4, dtype=np.uint8); tseq = np.full(int(tp.sum()) + 8, 4, dtype=np.uint8)
    for a in range(n_jobs):
        qseq[qoff[a]:qoff[a] + len(qs[a])] = qs[a]; tseq[toff[a]:toff[a] + len(ts[a])] = ts[a]
    return dict(qseq=qseq, tseq=tseq, qoff=qoff, toff=toff, qlen=qlens.astype(np.uint32), tlen=tlens.astype(np.uint32), h0=h0)


def make_repeat_flank_jobs(n_jobs: int, seed: int, kw: dict) -> dict:
    """Adversarial jobs for the closed-form answer (closed_form_job, ext_pair_core.cuh): targets that are tandem repeats of period
    1 .. dmax_k + 2 with a few point breaks (so shifted diagonals are clean or nearly clean), two-letter sequences and random ones;
    2 .. 7 substitutions planted in the query at spacings around the dmax_k + 2 boundary of the test; a third of the jobs with a small
    h0.  kw: the extension parameters (they set dmax_k).  Same dict as make_ext_jobs."""
    rng = np.random.Generator(np.random.PCG64(seed))
    dm = closed_form_eligible(**kw) or {2: 4}
    qs, ts, h0s = [], [], []
    for _ in range(n_jobs):
        k = int(rng.integers(2, 8))
        dk = dm.get(min(k, max(dm)), 4)
        ql = int(rng.integers(max(4, k * 3), 40 + k * (dk + 8)))
        tl = ql + int(rng.integers(0, 40))
        mode = rng.random()
        if mode < 0.6:       # tandem repeat with breaks
            per = int(rng.integers(1, dk + 3))
            unit = rng.integers(0, 4, size=per, dtype=np.uint8)
            t = np.tile(unit, tl // per + 2)[:tl].copy()
            nb = int(rng.integers(0, 2 + ql // 12))
            for x in rng.integers(0, tl, size=nb):
                t[x] = (t[x] + int(rng.integers(1, 4))) & 3
        elif mode < 0.8:     # low-complexity: two-letter alphabet
            t = rng.integers(0, 2, size=tl, dtype=np.uint8) * int(rng.integers(1, 4))
        else:
            t = rng.integers(0, 4, size=tl, dtype=np.uint8)
        q = t[:ql].copy()
        # k differences: spacings near the boundary dk + 2
        pos = [int(rng.integers(0, max(1, ql // (k + 1))))]
        for _m in range(k - 1):
            gap = dk + 2 + int(rng.integers(-2, 12)) if rng.random() < 0.7 else int(rng.integers(1, 60))
            pos.append(pos[-1] + max(1, gap))
        for x in pos:
            if x < ql:
                q[x] = (q[x] + int(rng.integers(1, 4))) & 3
        qs.append(q); ts.append(t)
        h0s.append(int(rng.integers(1, 40)) if rng.random() < 0.3 else int(rng.integers(19, 200)))
    qlens = np.array([len(x) for x in qs]); tlens = np.array([len(x) for x in ts])
    qp = (qlens + 7) // 8 * 8; tp = (tlens + 7) // 8 * 8
    qoff = np.zeros(n_jobs, np.uint32); toff = np.zeros(n_jobs, np.uint32)
    qoff[1:] = np.cumsum(qp)[:-1]; toff[1:] = np.cumsum(tp)[:-1]
    qseq = np.full(int(qp.sum()) + 8, 4, np.uint8); tseq = np.full(int(tp.sum()) + 8, 4, np.uint8)
    for a in range(n_jobs):
        qseq[qoff[a]:qoff[a] + qlens[a]] = qs[a]; tseq[toff[a]:toff[a] + tlens[a]] = ts[a]
    return dict(qseq=qseq, tseq=tseq, qoff=qoff, toff=toff, qlen=qlens.astype(np.uint32), tlen=tlens.astype(np.uint32), h0=np.array(h0s, np.uint32))


def closed_form_mask(jobs: dict, a: int = 1, b: int = 4, dmax=None, zdrop: int = 100) -> np.ndarray:
    """Which jobs the extender answers in closed form (an independent statement of closed_form_job's predicate, ext_pair_core.cuh):
    target at least as long as the query, every compared base in A/C/G/T, at most kcap = max(dmax) substituted bases, h0 > k b,
    k b <= zdrop, and -- with k >= 2 -- between each two neighbouring differences every diagonal shifted by 1 .. dmax[k] either way has
    a mismatch (or leaves the matrix) on the rows strictly between the earlier difference + dmax[k] and the later one.
    dmax: {k: dmax_k} as closed_form_eligible returns it (default: the default penalties' table)."""
    if dmax is None:
        dmax = closed_form_eligible()
    n = jobs["qlen"].size
    out = np.zeros(n, bool)
    kcap = max(dmax)
    for j in range(n):
        ql, tl, h0 = int(jobs["qlen"][j]), int(jobs["tlen"][j]), int(jobs["h0"][j])
        if ql == 0 or tl < ql:
            continue
        q = jobs["qseq"][int(jobs["qoff"][j]):int(jobs["qoff"][j]) + ql]
        t = jobs["tseq"][int(jobs["toff"][j]):int(jobs["toff"][j]) + ql]
        if (q > 3).any() or (t > 3).any():
            continue
        diff = np.nonzero(q != t)[0]
        k = len(diff)
        if k > kcap or h0 <= k * b or (zdrop > 0 and k * b > zdrop):
            continue
        ok = True
        dk = dmax[k] if k >= 2 else 0
        if k >= 2 and dk > 0:
            for m in range(k - 1):
                lo, hi = int(diff[m]) + dk + 1, int(diff[m + 1])
                if lo >= hi:
                    ok = False
                    break
                rows = np.arange(lo, hi)
                for s_ in [x for d in range(1, dk + 1) for x in (-d, d)]:
                    cols = rows + s_
                    inside = cols < ql
                    if not ((~inside).any() or (q[cols[inside]] != t[rows[inside]]).any()):
                        ok = False
                        break
                if not ok:
                    break
        out[j] = ok
    return out


CF_KMAX, CF_DMAX_CAP = 6, 24


def closed_form_eligible(w=100, zdrop=100, use_band=1, a=1, b=4, o_del=6, e_del=1, o_ins=6, e_ins=1, **_):
    """closed_params_from (ext_pair_core.cuh) restated: None when the parameters rule the closed-form answer out, else {k: dmax_k} for
    the numbers of differences k = 2 .. kcap that are taken, dmax_k the longest gap that costs no more than k mismatches"""
    g = min(o_del + e_del, o_ins + e_ins)
    if a < 1 or b < 1 or g <= a + b:
        return None

    def dmax_of(k):
        return max((k * (a + b) - o_del) // e_del, (k * (a + b) - o_ins) // e_ins, 0)
    if dmax_of(2) > 16 or (use_band and w < dmax_of(2) + 2):
        return None
    out = {2: dmax_of(2)}
    for k in range(3, CF_KMAX + 1):
        d = dmax_of(k)
        if d > CF_DMAX_CAP or (use_band and w < d + 2):
            break
        out[k] = d
    return out


def subset_jobs(jobs: dict, keep: np.ndarray) -> dict:
    """the jobs where keep is set, sequences copied into fresh arrays (same dict layout)"""
    idx = np.nonzero(keep)[0]
    qs, ts, qo, to = [], [], [], []
    nq = nt = 0
    for k in idx:
        ql, tl = int(jobs["qlen"][k]), int(jobs["tlen"][k])
        qp, tp = (ql + 7) // 8 * 8, (tl + 7) // 8 * 8
        q = np.full(qp, 4, np.uint8); t = np.full(tp, 4, np.uint8)
        q[:ql] = jobs["qseq"][int(jobs["qoff"][k]):int(jobs["qoff"][k]) + ql]; t[:tl] = jobs["tseq"][int(jobs["toff"][k]):int(jobs["toff"][k]) + tl]
        qs.append(q); ts.append(t); qo.append(nq); to.append(nt); nq += qp; nt += tp
    pad = np.full(8, 4, np.uint8)
    return dict(qseq=np.concatenate(qs + [pad]), tseq=np.concatenate(ts + [pad]), qoff=np.array(qo, np.uint32), toff=np.array(to, np.uint32),
                qlen=jobs["qlen"][idx].astype(np.uint32), tlen=jobs["tlen"][idx].astype(np.uint32), h0=jobs["h0"][idx].astype(np.uint32))


def dp_cells(oracle, jobs: dict, kw: dict, cnt_all: dict):
    """(cells the extender evaluates, jobs it answers in closed form) for a batch: the oracle's cell count without the jobs of the
    closed-form shape when the parameters admit the shortcut"""
    dmax = closed_form_eligible(**kw)
    if dmax is None:
        return cnt_all["cells"], 0
    m = closed_form_mask(jobs, kw.get("a", 1), kw.get("b", 4), dmax, kw.get("zdrop", 100))
    if not m.any():
        return cnt_all["cells"], 0
    _, c = oracle.ksw_batch(subset_jobs(jobs, m), oracle.make_params(**kw), n_threads=4)
    return cnt_all["cells"] - c["cells"], int(m.sum())
