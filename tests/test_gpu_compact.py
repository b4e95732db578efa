"""The compact host boundary of the aligner (bwa_b200_align_host_compact: 2-bit reads in, 40-byte region records out) and the
multi-device dispatcher over it (bwa_b200_multi_*), through the C ABI.  Both must return exactly what the full-record call
(bwa_b200_align_host_view, itself compared with the oracle in test_gpu_align.py) returns on the same reads: ragged lengths, N
bases, a read with no seed, an empty batch; chunks of every size dealt to two workers of one device, and to two devices when the
box has them (the replica made by bwa_b200_index_clone_to must seed and extend identically)."""
import ctypes as C

import numpy as np
import pytest

from tools import synth

pytestmark = pytest.mark.gpu
FIELDS = ["rb", "re", "qb", "qe", "score", "truesc", "seedcov", "rid", "w", "seedlen0", "frac_rep"]


@pytest.fixture(scope="module")
def setup(pkg, tmp_path_factory):
    assert pkg.lib().bwa_b200_device_count() > 0, "no CUDA device: these tests must run on the GPU box"
    g = synth.make_genome(150_000, seed=901)
    prefix = str(tmp_path_factory.mktemp("cmp") / "g")
    pkg.build_index(g, prefix, sa_intv=16, n_threads=4)
    idx = pkg.Index.load(prefix + ".bwt", prefix + ".sa", 0)
    idx.attach_ref(g)
    yield g, idx
    idx.free()


def ragged_reads(g, n, seed, uniform=None):
    rng = np.random.default_rng(seed)
    reads, _, _ = synth.make_reads(g, n, 150, seed=seed)
    out = []
    for r in range(n):
        L = uniform if uniform else int(rng.integers(30, 151))
        x = reads[r, :L].copy()
        if r % 17 == 3:
            x[rng.integers(0, L, size=3)] = 4          # N bases
        if r % 97 == 5:
            x[:] = rng.integers(0, 4, size=L)            # a read from nowhere
        out.append(x)
    lens = np.array([len(x) for x in out], np.uint64)
    off = np.zeros(n + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    return np.concatenate(out).astype(np.uint8), off


def full_records(pkg, idx, flat, off, params):
    packed, woff, rl = pkg.pack_codes(flat, off)
    al = pkg.Aligner(idx, rl.size, packed.size)
    got = al.align_host_view(packed.ctypes.data, woff.ctypes.data, rl.ctypes.data, rl.size, *params)
    al.destroy()
    return got


def same(compact, full):
    c = compact
    assert (c["n_regions"] == full["n_regions"]).all()
    assert (c["region_off"] == full["region_off"]).all()
    for f in FIELDS:
        assert (c[f] == full["regions"][f]).all(), f


@pytest.mark.parametrize("uniform", [None, 101])
def test_compact_equals_full_records(pkg, setup, uniform):
    g, idx = setup
    flat, off = ragged_reads(g, 3000, 11, uniform)
    params = (pkg.seed_params(19, 500), pkg.chain_params(w=100), pkg.ext_params())
    full = full_records(pkg, idx, flat, off, params)
    assert full["regions"].size > 3000
    p2, rl, nl = pkg.pack2_codes(flat, off, with_lengths=uniform is None)
    assert nl.size > 100
    al = pkg.Aligner(idx, 3000, int(((off[1:] - off[:-1] + 7) // 8).sum()))
    l0 = al.launches
    got = al.align_host_compact(p2.ctypes.data, rl.ctypes.data if rl is not None else None, uniform or 0, 3000, nl.ctypes.data, nl.size, *params)
    assert al.launches > l0
    same(pkg.unpack_compact(got), full)
    # twice on the same handle (buffers reused), then an empty batch
    got2 = al.align_host_compact(p2.ctypes.data, rl.ctypes.data if rl is not None else None, uniform or 0, 3000, nl.ctypes.data, nl.size, *params)
    assert got2["regions"].tobytes() == got["regions"].tobytes()
    empty = al.align_host_compact(None, None, 100, 0, None, 0, *params)
    assert empty["regions"].size == 0
    # a bad N entry is refused
    bad = nl.copy(); bad[0] = np.uint64(5000) << np.uint64(32)
    with pytest.raises(pkg.B200Error):
        al.align_host_compact(p2.ctypes.data, rl.ctypes.data if rl is not None else None, uniform or 0, 3000, bad.ctypes.data, bad.size, *params)
    al.destroy()


def run_multi(pkg, idx, devices, flat, off, params, chunk, workers=2):
    p2, rl, nl = pkg.pack2_codes(flat, off)
    n = rl.size
    m = pkg.MultiAligner(idx, devices, workers, chunk, 150)
    got = m.align_compact(p2.ctypes.data, rl.ctypes.data, 0, n, nl.ctypes.data, nl.size, *params)
    again = m.align_compact(p2.ctypes.data, rl.ctypes.data, 0, n, nl.ctypes.data, nl.size, *params)
    chunks = m.worker_chunks()
    launches = m.launches
    m.destroy()
    assert again["regions"].tobytes() == got["regions"].tobytes()
    return got, chunks, launches


@pytest.mark.parametrize("chunk", [64, 333, 5000])
def test_multi_one_device_equals_single_call(pkg, setup, chunk):
    g, idx = setup
    flat, off = ragged_reads(g, 2500, 12)
    params = (pkg.seed_params(19, 500), pkg.chain_params(w=100), pkg.ext_params())
    full = full_records(pkg, idx, flat, off, params)
    got, chunks, launches = run_multi(pkg, idx, [0], flat, off, params, chunk)
    same(pkg.unpack_compact(got), full)
    assert sum(chunks) == 2 * ((2500 + chunk - 1) // chunk) and launches > 0
    if chunk == 64:
        assert min(chunks) > 0                       # both workers took part


def test_multi_two_devices_equals_single_call(pkg, setup):
    if pkg.lib().bwa_b200_device_count() < 2:
        pytest.skip("one device on this box")
    g, idx = setup
    flat, off = ragged_reads(g, 4000, 13)
    params = (pkg.seed_params(19, 500, True), pkg.chain_params(w=100), pkg.ext_params())
    full = full_records(pkg, idx, flat, off, params)
    got, chunks, _ = run_multi(pkg, idx, [0, 1], flat, off, params, 250)
    same(pkg.unpack_compact(got), full)
    assert min(chunks) > 0                           # every worker of both devices took part


def test_multi_refuses_unsorted_n_list(pkg, setup):
    g, idx = setup
    flat, off = ragged_reads(g, 300, 14)
    params = (pkg.seed_params(19, 500), pkg.chain_params(w=100), pkg.ext_params())
    p2, rl, nl = pkg.pack2_codes(flat, off)
    assert nl.size > 2
    bad = nl[::-1].copy()
    m = pkg.MultiAligner(idx, [0], 1, 100, 150)
    with pytest.raises(pkg.B200Error):
        m.align_compact(p2.ctypes.data, rl.ctypes.data, 0, 300, bad.ctypes.data, bad.size, *params)
    m.destroy()


@pytest.mark.parametrize("chunk,workers", [(200, 2), (5000, 1), (77, 3)])
def test_multi_two_batches_in_flight(pkg, setup, chunk, workers):
    """bwa_b200_multi_submit_compact / _wait: two different batches in flight, several rounds through the two slots; every batch's
    results equal the synchronous call's, whatever the order the workers finished the chunks in; a third submit is refused"""
    g, idx = setup
    params = (pkg.seed_params(19, 500), pkg.chain_params(w=100), pkg.ext_params())
    batches = []
    for seed, n in ((21, 1800), (22, 2300), (23, 900)):
        flat, off = ragged_reads(g, n, seed)
        p2, rl, nl = pkg.pack2_codes(flat, off)
        batches.append((p2, rl, nl, n))
    m = pkg.MultiAligner(idx, [0], workers, chunk, 150)
    want = [m.align_compact(p2.ctypes.data, rl.ctypes.data, 0, n, nl.ctypes.data, nl.size, *params) for p2, rl, nl, n in batches]

    def submit(k):
        p2, rl, nl, n = batches[k]
        return m.submit_compact(p2.ctypes.data, rl.ctypes.data, 0, n, nl.ctypes.data, nl.size, *params)
    order = [0, 1, 2, 1, 0, 2, 2, 0]
    t_prev = submit(order[0])
    for i in range(1, len(order)):
        t_next = submit(order[i])                    # two in flight
        if i == 1:
            with pytest.raises(pkg.B200Error):
                submit(2)                            # a third is refused, and refusing it leaves the two alone
        got = m.wait(t_prev)
        w = want[order[i - 1]]
        assert (got["n_regions"] == w["n_regions"]).all() and got["regions"].tobytes() == w["regions"].tobytes()
        t_prev = t_next
    got = m.wait(t_prev)
    assert got["regions"].tobytes() == want[order[-1]]["regions"].tobytes()
    with pytest.raises(pkg.B200Error):
        m.wait(t_prev)                               # nothing behind the ticket any more... until it is submitted again
    m.destroy()
