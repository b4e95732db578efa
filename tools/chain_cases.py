"""Seed sets for the chaining / extension-job tests: synthetic reads' worth of SMEM groups built to hit the
branches of mem_chain / test_and_merge / mem_chain_flt / mem_chain2aln (collinear runs, off-diagonal seeds,
repeats above max_occ, equal reference positions, many chains per read so the chain B-tree splits, seeds that
bridge contigs or the forward/reverse boundary, seeds at contig ends)."""
import numpy as np


def random_genome(rng, lens):
    return rng.integers(0, 4, size=int(np.sum(lens)), dtype=np.uint8)


def make_read_seeds(rng, l_pac, l_query, max_occ, style):
    """-> (rbeg u64[n], qq i32[n,2], score u32[n]) in the reference's full layout (a group of s rows per SMEM)"""
    rbeg, qq, score = [], [], []
    n_groups = int(rng.integers(1, 5)) if style != "many" else int(rng.integers(6, 16))
    starts = np.sort(rng.integers(0, max(1, l_query - 19), size=n_groups))
    loci = rng.integers(0, 2 * l_pac - l_query - 1, size=3)
    prev_end = 0
    if style == "full":          # one SMEM spanning the whole read: no extension at all (align_sides == 0)
        n_groups, starts = 1, np.array([0])
    for g in range(n_groups):
        qb = int(max(starts[g], 0))
        ln = int(rng.integers(19, max(20, min(l_query - qb, 80) + 1)))
        qe = min(l_query, qb + ln)
        if style == "full":
            qe = l_query
        if qe - qb < 19:
            continue
        if qe <= prev_end:           # SMEMs are not nested: ends ascend with starts
            qe = min(l_query, prev_end + 1)
            if qe - qb < 19:
                continue
        prev_end = qe
        if style == "repeat" and rng.random() < 0.4:
            s = int(rng.integers(max_occ + 1, 3 * max_occ))
        elif style == "many":
            s = int(rng.integers(1, 12))
        else:
            s = int(rng.integers(1, 4))
        rows = []
        for _ in range(s):
            u = rng.random()
            locus = int(loci[int(rng.integers(0, 3))])
            if u < 0.55:        # on the diagonal of a locus (chains grow)
                r = locus + qb + int(rng.integers(-2, 3)) * int(rng.random() < 0.3)
            elif u < 0.7:       # near a locus, off-diagonal by up to a few hundred
                r = locus + qb + int(rng.integers(-400, 400))
            elif u < 0.8:       # exactly an earlier position (equal keys in the chain tree)
                r = int(rbeg[int(rng.integers(0, len(rbeg)))]) if rbeg else locus
            elif u < 0.85:      # around the forward/reverse boundary or a contig edge
                r = l_pac + int(rng.integers(-60, 60))
            else:
                r = int(rng.integers(0, 2 * l_pac - l_query))
            r = min(max(r, 0), 2 * l_pac - (qe - qb))
            rows.append(r)
        rows.sort()              # SA rows of one interval come out in suffix order, not position order; any order is legal input
        if rng.random() < 0.5:
            rng.shuffle(rows)
        for k, r in enumerate(rows):
            rbeg.append(r); qq.append((qb, qe)); score.append(s if k == 0 else 0)
    if not rbeg:
        return np.zeros(0, np.uint64), np.zeros((0, 2), np.int32), np.zeros(0, np.uint32)
    return np.array(rbeg, dtype=np.uint64), np.array(qq, dtype=np.int32), np.array(score, dtype=np.uint32)


def to_compact(rbeg, qq, score, max_occ):
    """full layout -> the sampled layout of bwa_b200_seeds_t (max_occ > 0)"""
    rb, q2, sc = [], [], []
    i, n = 0, len(rbeg)
    while i < n:
        s = int(score[i])
        step = s // max_occ if s > max_occ else 1
        k = cnt = 0
        while k < s and cnt < max_occ:
            rb.append(rbeg[i + k]); q2.append(qq[i]); sc.append(s if cnt == 0 else 0)
            k += step; cnt += 1
        i += s
    if not rb:
        return np.zeros(0, np.uint64), np.zeros((0, 2), np.int32), np.zeros(0, np.uint32)
    return np.array(rb, dtype=np.uint64), np.array(q2, dtype=np.int32).reshape(-1, 2), np.array(sc, dtype=np.uint32)


def make_cases(seed, n_reads, contig_lens=(30000, 1500, 20000), max_occ=50):
    rng = np.random.default_rng(seed)
    fwd = random_genome(rng, contig_lens)
    l_pac = len(fwd)
    cases = []
    styles = ["plain", "repeat", "many", "plain", "repeat", "full", "many"]
    for r in range(n_reads):
        l_query = int(rng.choice([150, 150, 101, 250, 60]))
        query = rng.integers(0, 4, size=l_query, dtype=np.uint8)
        if rng.random() < 0.1:
            query[int(rng.integers(0, l_query))] = 4
        rb, qq, sc = make_read_seeds(rng, l_pac, l_query, max_occ, styles[r % len(styles)])
        cases.append((query, rb, qq, sc))
    return fwd, cases


def make_long_cases(seed, n_reads, contig_lens=(30000, 1500, 20000), max_occ=50):
    """Reads long enough for mem_flt_chained_seeds to act (>= 757 bases at the default options): cut from the text fwd + revcomp(fwd) with
    2-30 % substitutions (so that the local alignment around a seed scores on either side of min_HSP_score), seeds on the read's true
    diagonal, at random places, around the forward/reverse boundary, at contig and text ends, and a few of 200 bases and more (the
    alignment is skipped for those)."""
    rng = np.random.default_rng(seed)
    fwd = random_genome(rng, contig_lens)
    l_pac = len(fwd)
    text = np.concatenate([fwd, (3 - fwd)[::-1]])
    edges = np.concatenate([[0], np.cumsum(contig_lens)])
    edges = np.concatenate([edges, 2 * l_pac - edges])
    cases = []
    for r in range(n_reads):
        l_query = int(rng.choice([760, 800, 1000, 1500, 2500]))
        pos = int(rng.integers(0, 2 * l_pac - l_query))
        query = text[pos:pos + l_query].copy()
        rate = float(rng.choice([0.02, 0.1, 0.2, 0.3]))
        mut = rng.random(l_query) < rate
        query[mut] = (query[mut] + rng.integers(1, 4, size=int(mut.sum()))) & 3
        if rng.random() < 0.3:
            query[rng.integers(0, l_query, size=3)] = 4
        n_groups = int(rng.integers(3, 30))
        qbs = np.sort(rng.integers(0, l_query - 19, size=n_groups))
        rbeg, qq, score = [], [], []
        prev_end = 0
        for g in range(n_groups):
            qb = int(qbs[g])
            ln = int(rng.integers(19, 90)) if rng.random() < 0.93 else int(rng.integers(190, 260))
            qe = min(l_query, qb + ln)
            if qe <= prev_end:
                qe = min(l_query, prev_end + 1)
            if qe - qb < 19:
                continue
            prev_end = qe
            s = int(rng.integers(1, 5)) if rng.random() < 0.9 else int(rng.integers(max_occ + 1, 2 * max_occ))
            rows = []
            for _ in range(s):
                u = rng.random()
                if u < 0.6:
                    rr = pos + qb + int(rng.integers(-3, 4)) * int(rng.random() < 0.2)
                elif u < 0.7:
                    rr = l_pac + int(rng.integers(-120, 120))
                elif u < 0.8:
                    rr = int(rng.choice(edges)) + int(rng.integers(-120, 120))
                else:
                    rr = int(rng.integers(0, 2 * l_pac - l_query))
                rows.append(min(max(rr, 0), 2 * l_pac - (qe - qb)))
            if rng.random() < 0.5:
                rows.sort()
            for k, rr in enumerate(rows):
                rbeg.append(rr); qq.append((qb, qe)); score.append(s if k == 0 else 0)
        if not rbeg:
            rbeg, qq, score = [pos], [(0, 30)], [1]
        cases.append((query, np.array(rbeg, dtype=np.uint64), np.array(qq, dtype=np.int32).reshape(-1, 2), np.array(score, dtype=np.uint32)))
    return fwd, cases
