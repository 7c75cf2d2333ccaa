"""GPU parity: the CUDA path (through the C ABI) must equal the CPU oracle bit for bit on LAS
records and trace-point integers, on seeded inputs the oracle finishes in seconds."""
import numpy as np
import pytest

from dentist_b200 import synth

pytestmark = pytest.mark.gpu

ORC = dict(k=14, w=6, h=35, t=32, cdiff=20, xdrop=300, wmax=30, rounds=3, poolmul=64)


def run_both(A, B, tspace, minlen, self_block=0, a_mask=None, b_mask=None, **over):
    from dentist_b200 import dazzler
    from oracle import oracle
    gpu_only = {k: over.pop(k) for k in list(over) if k in ("join_mode",)}
    o = dict(ORC); o.update(over)
    am = bm = None
    if a_mask is not None:
        am = np.zeros(A.total, np.uint8)
        for r, iv in enumerate(a_mask):
            for b, e in iv:
                am[A.off[r] + b:A.off[r] + e] = 1
    if b_mask is not None:
        bm = np.zeros(B.total, np.uint8)
        for r, iv in enumerate(b_mask):
            for b, e in iv:
                bm[B.off[r] + b:B.off[r] + e] = 1
    la, tr, st = oracle.align(A.off, A.bases, B.off, B.bases, a_mask=am, b_mask=bm, tspace=tspace, minlen=minlen,
                              self=self_block, **o)
    ga = dazzler.Block(A.off, A.bases, mask=a_mask)
    gb = ga if B is A else dazzler.Block(B.off, B.bases, mask=b_mask)
    g = {k: v for k, v in o.items() if k not in ("cdiff",)}
    g.update(gpu_only)
    rec, toff, gtr, gst = dazzler.align_blocks(ga, gb, tspace=tspace, minlen=minlen, self_block=self_block, e=0.7, **g)
    return (la, tr, st), (rec, toff, gtr, gst)


def assert_same(orc, gpu):
    (la, tr, st), (rec, toff, gtr, gst) = orc, gpu
    assert gst["hits"] == st["nhits"]
    assert gst["seeds"] == st["nseeds"]
    assert len(rec) == len(la)
    for f in ("aread", "bread", "abpos", "aepos", "bbpos", "bepos", "diffs", "tlen", "flags"):
        assert np.array_equal(rec[f], la[f]), f
    assert np.array_equal(toff, la["toff"])
    assert np.array_equal(gtr, tr)


def small_case(seed, cov=4, rl=6000, err=0.13, glen=120000, gaps=2):
    sc = synth.make_scaffolds(2, glen, seed, n_repeats=1, repeat_copies=4)
    ref, _ = synth.contigs_from(sc, synth.make_gaps(sc, gaps, seed + 1))
    reads, _ = synth.simulate_reads(sc, cov, rl, rl // 3, err, seed + 2)
    return ref, reads


@pytest.mark.parametrize("seed,tspace", [(1, 100), (2, 126), (3, 100)])
def test_ref_vs_reads_matches_oracle(seed, tspace):
    ref, reads = small_case(seed)
    orc, gpu = run_both(ref, reads, tspace, 500)
    assert len(orc[0]) > 50
    assert_same(orc, gpu)


@pytest.mark.parametrize("join_mode", [1, 2])
def test_both_join_strategies_match_oracle(join_mode):
    """sorted-merge join (both tuple lists radix sorted) and index-lookup join give identical LAS."""
    ref, reads = small_case(11, cov=3)
    orc, gpu = run_both(ref, reads, 100, 500, join_mode=join_mode)
    assert_same(orc, gpu)
    sc = synth.make_scaffolds(1, 20000, 12, n_repeats=0)
    pile, _ = synth.simulate_reads(sc, 8, 6000, 1500, 0.13, 13)
    orc, gpu = run_both(pile, pile, 126, 500, self_block=1, join_mode=join_mode)
    assert_same(orc, gpu)


@pytest.mark.parametrize("wmax,xdrop", [(62, 300), (40, 600), (12, 300), (30, 900)])
def test_window_cap_variants_match_oracle(wmax, xdrop):
    """wmax > 30 runs the shared-memory two-slot kernel, wmax <= 30 the register-resident one; tight caps and
    wide x-drops exercise the window clamp."""
    ref, reads = small_case(17, cov=3)
    orc, gpu = run_both(ref, reads, 100, 500, wmax=wmax, xdrop=xdrop)
    assert len(orc[0]) > 30
    assert_same(orc, gpu)


@pytest.mark.parametrize("k", [16, 20, 31])
def test_long_kmers_use_the_wide_index_and_match_oracle(k):
    """k > 15 (damapper's default is 20): 16-byte {kmer, position} tuples, 3-word k-mer extraction, same results as the
    oracle -- with and without seed masks, for a reference mapping and for a pile self-alignment."""
    ref, reads = small_case(41 + k, cov=4, err=0.08)
    orc, gpu = run_both(ref, reads, 100, 500, k=k)
    assert len(orc[0]) > (5 if k == 31 else 40)
    assert_same(orc, gpu)
    amask = [[(0, int(ref.off[r + 1] - ref.off[r]) // 3)] for r in range(ref.nreads)]
    bmask = [[(37, 911)] for _ in range(reads.nreads)]
    orc_m, gpu_m = run_both(ref, reads, 100, 500, a_mask=amask, b_mask=bmask, k=k)
    assert_same(orc_m, gpu_m)
    assert orc_m[2]["nhits"] < orc[2]["nhits"]
    sc = synth.make_scaffolds(1, 20000, 43, n_repeats=0)
    pile, _ = synth.simulate_reads(sc, 8, 6000, 1500, 0.08, 44)
    orc, gpu = run_both(pile, pile, 126, 500, self_block=1, k=k)
    assert_same(orc, gpu)
    # an A block without a single valid k-mer (reads shorter than k): an empty index, no alignments, no crash
    tiny = synth.Block(np.arange(0, 60, 15, dtype=np.int64), np.random.default_rng(3).integers(0, 4, 45).astype(np.uint8))
    orc, gpu = run_both(tiny, reads, 100, 500, k=k)
    assert len(orc[0]) == 0
    assert_same(orc, gpu)


@pytest.mark.parametrize("k", [14, 20])
def test_resident_reference_index_gives_identical_results(k):
    """dn_block_index: the A-side index built once and reused over several read blocks; dropped by a mask change."""
    from dentist_b200 import dazzler
    ref, reads = small_case(61, cov=4)
    ga = dazzler.Block(ref.off, ref.bases)
    half = reads.nreads // 2
    blocks = [dazzler.Block(reads.off[:half + 1], reads.bases[:reads.off[half]]),
              dazzler.Block(reads.off[half:] - reads.off[half], reads.bases[reads.off[half]:])]
    plain = [dazzler.align_blocks(ga, gb, tspace=100, minlen=500, k=k) for gb in blocks]
    ga.index(k)
    for gb, want in zip(blocks, plain):
        for _ in range(2):
            got = dazzler.align_blocks(ga, gb, tspace=100, minlen=500, k=k)
            assert got[0].tobytes() == want[0].tobytes() and got[2].tobytes() == want[2].tobytes() and len(got[0]) > 20
    other = dazzler.align_blocks(ga, blocks[0], tspace=100, minlen=500, k=k + 1)         # another k: the index is ignored
    assert len(other[0]) > 20
    assert ga.maskDust(threshold=0.5) > 0                                                # a new mask drops the index
    masked = dazzler.align_blocks(ga, blocks[0], tspace=100, minlen=500, k=k)
    gm = dazzler.Block(ref.off, ref.bases); gm.maskDust(threshold=0.5)
    want = dazzler.align_blocks(gm, blocks[0], tspace=100, minlen=500, k=k)
    assert masked[0].tobytes() == want[0].tobytes() and masked[3]["hits"] < plain[0][3]["hits"]


def test_pile_self_alignment_matches_oracle():
    # processPileUps: daligner -s126 -l500 -e0.7 X X  (commandline.d:2886-2902)
    sc = synth.make_scaffolds(1, 25000, 21, n_repeats=0)
    reads, _ = synth.simulate_reads(sc, 12, 7000, 2000, 0.13, 22)
    orc, gpu = run_both(reads, reads, 126, 500, self_block=1)
    assert len(orc[0]) > 100
    assert_same(orc, gpu)


def test_ont_like_errors_and_long_reads():
    sc = synth.make_scaffolds(1, 200000, 31, n_repeats=0)
    ref, _ = synth.contigs_from(sc, synth.make_gaps(sc, 1, 32))
    reads, _ = synth.simulate_reads(sc, 3, 20000, 10000, 0.12, 33, mix=(0.25, 0.45, 0.30))
    orc, gpu = run_both(ref, reads, 100, 1000)
    assert len(orc[0]) > 10
    assert_same(orc, gpu)


def test_masked_kmers_are_not_seeded():
    ref, reads = small_case(7, cov=3)
    amask = [[(0, int(ref.off[r + 1] - ref.off[r]) // 2)] for r in range(ref.nreads)]
    bmask = [[(100, 400)] for _ in range(reads.nreads)]
    orc, gpu = run_both(ref, reads, 100, 500, a_mask=amask, b_mask=bmask)
    assert_same(orc, gpu)
    orc2, _ = run_both(ref, reads, 100, 500)
    assert orc[2]["nhits"] < orc2[2]["nhits"]


def test_edge_cases_empty_tiny_and_identical():
    from dentist_b200 import dazzler
    rng = np.random.default_rng(5)
    # reads shorter than k, empty read, identical reads (long exact slides crossing many tiles)
    g = rng.integers(0, 4, 3000, dtype=np.uint8)
    seqs = [g, g.copy(), g[:10], np.zeros(0, np.uint8), (3 - g)[::-1].copy(), g[500:2500].copy()]
    off = np.zeros(len(seqs) + 1, np.int64); off[1:] = np.cumsum([len(s) for s in seqs])
    blk = synth.Block(off, np.concatenate(seqs))
    orc, gpu = run_both(blk, blk, 126, 500, self_block=1)
    assert len(orc[0]) >= 6
    assert_same(orc, gpu)
    # no hits at all
    a = synth.Block(np.array([0, 2000]), rng.integers(0, 4, 2000, dtype=np.uint8))
    b = synth.Block(np.array([0, 2000]), rng.integers(0, 4, 2000, dtype=np.uint8))
    orc, gpu = run_both(a, b, 100, 500)
    assert len(gpu[0]) == 0
    assert_same(orc, gpu)
    # empty block
    e = dazzler.Block(np.array([0]), np.zeros(0, np.uint8))
    rec, _, _, _ = dazzler.align_blocks(e, e, tspace=100)
    assert len(rec) == 0


def test_bps_input_equals_byte_input():
    from dentist_b200 import dazzler
    ref, reads = small_case(9, cov=2)
    bps, boff = [], []
    o = 0
    for r in range(reads.nreads):
        p = synth.pack_2bit_dazz(reads.read(r)); boff.append(o); bps.append(p); o += len(p)
    gb1 = dazzler.Block(reads.off, reads.bases)
    gb2 = dazzler.Block(reads.off, bps=np.concatenate(bps), boff=np.array(boff))
    ga = dazzler.Block(ref.off, ref.bases)
    r1 = dazzler.align_blocks(ga, gb1, tspace=100, minlen=500)
    r2 = dazzler.align_blocks(ga, gb2, tspace=100, minlen=500)
    assert len(r1[0]) > 10
    assert r1[0].tobytes() == r2[0].tobytes() and np.array_equal(r1[2], r2[2])


def test_trace_invariants_at_scale():
    """Size-independent properties on a larger input (no oracle): tile counts, sums, order."""
    from dentist_b200 import dazzler
    from oracle import las
    sc = synth.make_scaffolds(4, 1000000, 41)
    ref, _ = synth.contigs_from(sc, synth.make_gaps(sc, 5, 42))
    reads, truth = synth.simulate_reads(sc, 5, 10000, 3000, 0.13, 43)
    ga, gb = dazzler.Block(ref.off, ref.bases), dazzler.Block(reads.off, reads.bases)
    rec, toff, tr, st = dazzler.align_blocks(ga, gb, tspace=100, minlen=1000)
    assert len(rec) > 1500
    nt = -(-rec["aepos"] // 100) - rec["abpos"] // 100
    assert np.array_equal(rec["tlen"], 2 * nt)
    ends = np.cumsum(rec["tlen"])
    assert np.array_equal(toff, ends - rec["tlen"])
    t = tr.reshape(-1, 2).astype(np.int64)
    idx = np.repeat(np.arange(len(rec)), nt)
    assert np.array_equal(np.bincount(idx, t[:, 0], len(rec)).astype(np.int64), rec["diffs"])
    assert np.array_equal(np.bincount(idx, t[:, 1], len(rec)).astype(np.int64), rec["bepos"] - rec["bbpos"])
    key = np.stack([rec["aread"], rec["bread"], rec["flags"] & 1, rec["abpos"]], 1).tolist()
    assert key == sorted(key)
    covered = np.zeros(reads.nreads); np.add.at(covered, rec["bread"], rec["bepos"] - rec["bbpos"])
    assert (covered / np.diff(reads.off) > 0.8).mean() > 0.9
    # idempotence: same input, same bytes
    rec2, _, tr2, _ = dazzler.align_blocks(ga, gb, tspace=100, minlen=1000)
    assert rec.tobytes() == rec2.tobytes() and np.array_equal(tr, tr2)


def test_long_ont_like_reads_at_scale():
    """configs[2]-like reads (20 kb +- 10 kb, 12 % ONT-like error, some > 50 kb): segments beyond the shared-memory
    classes, long traces and the record pool are exercised; checked through size-independent invariants."""
    from dentist_b200 import dazzler
    sc = synth.make_scaffolds(3, 1500000, 51)
    ref, meta = synth.contigs_from(sc, synth.make_gaps(sc, 3, 52))
    reads, truth = synth.simulate_reads(sc, 4, 20000, 10000, 0.12, 53, mix=(0.25, 0.45, 0.30))
    long_reads, _ = synth.simulate_reads(sc, 0.15, 90000, 5000, 0.12, 54, mix=(0.25, 0.45, 0.30))
    seqs = [reads.read(r) for r in range(reads.nreads)] + [long_reads.read(r) for r in range(long_reads.nreads)]
    off = np.zeros(len(seqs) + 1, np.int64); off[1:] = np.cumsum([len(s) for s in seqs])
    allr = synth.Block(off, np.concatenate(seqs))
    assert np.diff(off).max() > 80000
    ga, gb = dazzler.Block(ref.off, ref.bases), dazzler.Block(allr.off, allr.bases)
    rec, toff, tr, st = dazzler.align_blocks(ga, gb, tspace=100, minlen=1000)
    nt = -(-rec["aepos"] // 100) - rec["abpos"] // 100
    assert np.array_equal(rec["tlen"], 2 * nt)
    t = tr.reshape(-1, 2).astype(np.int64)
    idx = np.repeat(np.arange(len(rec)), nt)
    assert np.array_equal(np.bincount(idx, t[:, 0], len(rec)).astype(np.int64), rec["diffs"])
    assert np.array_equal(np.bincount(idx, t[:, 1], len(rec)).astype(np.int64), rec["bepos"] - rec["bbpos"])
    lens = np.diff(off)
    assert (rec["bepos"] <= lens[rec["bread"]]).all() and (rec["aepos"] <= np.diff(ref.off)[rec["aread"]]).all()
    covered = np.zeros(allr.nreads); np.add.at(covered, rec["bread"], rec["bepos"] - rec["bbpos"])
    assert (covered / lens > 0.8).mean() > 0.9
    err = rec["diffs"].sum() / (rec["aepos"] - rec["abpos"]).sum()
    assert 0.08 < err < 0.16
    # the same job through the sorted-merge join gives the same bytes
    rec2, _, tr2, _ = dazzler.align_blocks(ga, gb, tspace=100, minlen=1000, join_mode=1)
    assert rec.tobytes() == rec2.tobytes() and np.array_equal(tr, tr2)


def test_device_las_merge_equals_lasort_order():
    """dn_las_merge_device (what replaces LAmerge after the all-gatherv): segments split by read range and
    concatenated on the device merge back into exactly the single-GPU LAS."""
    import ctypes as C
    import torch
    from dentist_b200 import dazzler, _lib
    ref, reads = small_case(23, cov=4)
    ga, gb = dazzler.Block(ref.off, ref.bases), dazzler.Block(reads.off, reads.bases)
    las = dazzler.align(ga, gb, tspace=100, minlen=500)
    rec, toff, tr = las.rec, las.toff, las.trace
    cuts = [0, reads.nreads // 3, 2 * reads.nreads // 3, reads.nreads]
    parts_r, parts_t = [], []
    for lo, hi in zip(cuts[:-1][::-1], cuts[1:][::-1]):              # segments in a different order than the reads
        sel = np.flatnonzero((rec["bread"] >= lo) & (rec["bread"] < hi))
        parts_r.append(rec[sel]); parts_t.append(np.concatenate([tr[toff[i]:toff[i] + rec[i]["tlen"]] for i in sel]))
    drec = torch.from_numpy(np.frombuffer(np.concatenate(parts_r).tobytes(), np.uint8).copy()).cuda()
    dtr = torch.from_numpy(np.concatenate(parts_t).view(np.uint8).copy()).cuda()
    torch.cuda.synchronize()
    buf = _lib.LasBuf()
    _lib.check(_lib.lib().dn_las_merge_device(C.c_void_p(drec.data_ptr()), len(rec), C.c_void_p(dtr.data_ptr()), len(tr), 100,
                                              int(np.diff(ref.off).max()), int(np.diff(reads.off).max()), ref.nreads, reads.nreads, C.byref(buf)))
    m = dazzler.Las(buf)
    assert m.rec.tobytes() == rec.tobytes() and np.array_equal(m.toff, toff) and np.array_equal(m.trace, tr)
    # and the numpy merge used by the gloo tests agrees
    from dentist_b200 import sharding
    r2, o2, t2 = sharding.merge_las(parts_r, parts_t)
    assert r2.tobytes() == rec.tobytes() and np.array_equal(t2, tr)


def test_align_host_equals_resident_blocks():
    """dn_align_host (overlapped upload of B on a copy stream) returns what upload + dn_align_blocks returns: byte and
    .bps input, with masks, several calls in a row (recycled staging buffers)."""
    from dentist_b200 import dazzler
    ref, reads = small_case(71, cov=4)
    want = dazzler.align_blocks(dazzler.Block(ref.off, ref.bases), dazzler.Block(reads.off, reads.bases), tspace=100, minlen=500, k=20)
    assert len(want[0]) > 30
    parts, boff, o = [], [], 0
    for r in range(reads.nreads):
        q = synth.pack_2bit_dazz(reads.read(r)); boff.append(o); parts.append(q); o += len(q)
    bps, boff = np.concatenate(parts), np.array(boff, np.int64)
    for _ in range(3):
        got = dazzler.align_host(dazzler.HostBlock(ref.off, ref.bases), dazzler.HostBlock(reads.off, bps=bps, boff=boff), tspace=100, minlen=500, k=20)
        assert got[0].tobytes() == want[0].tobytes() and got[2].tobytes() == want[2].tobytes()
    # the chunked upload (the count pass follows the chunks of the arriving block), forced onto this small block: 2, 5 and 13 chunks,
    # .bps and byte input, k = 14 and k = 20 indexes
    import os
    os.environ["DN_UPLOAD_CHUNK_MIN_BYTES"] = "1"
    try:
        for chunks in ("2", "5", "13"):
            os.environ["DN_UPLOAD_CHUNKS"] = chunks
            got = dazzler.align_host(dazzler.HostBlock(ref.off, ref.bases), dazzler.HostBlock(reads.off, bps=bps, boff=boff), tspace=100, minlen=500, k=20)
            assert got[0].tobytes() == want[0].tobytes() and got[2].tobytes() == want[2].tobytes(), chunks
        want14 = dazzler.align_blocks(dazzler.Block(ref.off, ref.bases), dazzler.Block(reads.off, reads.bases), tspace=100, minlen=500, k=14)
        got = dazzler.align_host(dazzler.HostBlock(ref.off, ref.bases), dazzler.HostBlock(reads.off, reads.bases), tspace=100, minlen=500, k=14)
        assert got[0].tobytes() == want14[0].tobytes() and got[2].tobytes() == want14[2].tobytes()
    finally:
        del os.environ["DN_UPLOAD_CHUNK_MIN_BYTES"]; os.environ.pop("DN_UPLOAD_CHUNKS", None)
    amask = [[(0, 500)] for _ in range(ref.nreads)]; bmask = [[(100, 400)] for _ in range(reads.nreads)]
    want = dazzler.align_blocks(dazzler.Block(ref.off, ref.bases, mask=amask), dazzler.Block(reads.off, reads.bases, mask=bmask), tspace=100, minlen=500)
    got = dazzler.align_host(dazzler.HostBlock(ref.off, ref.bases, mask=amask), dazzler.HostBlock(reads.off, reads.bases, mask=bmask), tspace=100, minlen=500)
    assert got[0].tobytes() == want[0].tobytes() and got[2].tobytes() == want[2].tobytes()
    same = dazzler.HostBlock(reads.off, reads.bases)
    self_host = dazzler.align_host(same, same, tspace=126, minlen=500, self_block=1)
    g = dazzler.Block(reads.off, reads.bases)
    self_dev = dazzler.align_blocks(g, g, tspace=126, minlen=500, self_block=1)
    assert self_host[0].tobytes() == self_dev[0].tobytes()


def test_concurrent_callers_get_the_serial_results():
    """DENTIST calls the boundary from std.parallelism worker threads (processPileUps/package.d:153): four host threads
    hammer the same library (align, filters, QVs, consensus) and must each see exactly what a serial run returns."""
    import threading
    from dentist_b200 import dazzler
    cases = []
    for t in range(4):
        sc = synth.make_scaffolds(1, 20000, 300 + t, n_repeats=0)
        pile, _ = synth.simulate_reads(sc, 8, 6000, 1500, 0.13, 400 + t)
        cases.append(pile)

    def work(pile):
        g = dazzler.Block(pile.off, pile.bases)
        lens = np.diff(pile.off)
        las = dazzler.align(g, g, tspace=126, minlen=500, self_block=1)
        raw = las.rec.tobytes()
        las.filterLocalAlignments(0.3)
        las.chainLocalAlignments()
        q, _ = dazzler.computeQVs(lens, las, 4)
        las.filterPileUpAlignments(lens, lens, 126)
        las.forceFlat()
        cons = dazzler.getConsensus(g, las, [0, 1])
        g.free()
        return raw, q.tobytes(), las.rec.tobytes(), [c.tobytes() for c in cons]

    serial = [work(p) for p in cases]
    for _ in range(3):
        out = [None] * 4
        th = [threading.Thread(target=lambda i=i: out.__setitem__(i, work(cases[i]))) for i in range(4)]
        [t.start() for t in th]; [t.join() for t in th]
        assert out == serial


def _shards(reads, cuts):
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        yield lo, synth.Block(reads.off[lo:hi + 1] - reads.off[lo], reads.bases[reads.off[lo]:reads.off[hi]])


def test_shard_invariance():
    """align(ref, reads) == merge(align(ref, shard_i)): the rounds are counted per (bread, strand, aread) group, so
    what a read aligns to never depends on the reads it shares a block with.  4 shards and one-read shards."""
    from dentist_b200 import dazzler, sharding
    sc = synth.make_scaffolds(2, 150000, 81, n_repeats=2, repeat_copies=6)          # repeats: groups that go several rounds
    ref, _ = synth.contigs_from(sc, synth.make_gaps(sc, 2, 82))
    reads, _ = synth.simulate_reads(sc, 1.5, 7000, 2500, 0.13, 83)
    ga = dazzler.Block(ref.off, ref.bases)
    whole = dazzler.align(ga, dazzler.Block(reads.off, reads.bases), tspace=100, minlen=500)
    assert len(whole) > 150
    for cuts in (np.linspace(0, reads.nreads, 5).astype(int), np.arange(reads.nreads + 1)):
        recs, trs = [], []
        for lo, sh in _shards(reads, cuts):
            las = dazzler.align(ga, dazzler.Block(sh.off, sh.bases), tspace=100, minlen=500)
            r = las.rec.copy(); r["bread"] += lo
            recs.append(r); trs.append(las.trace.copy())
        mrec, mtoff, mtr = sharding.merge_las(recs, trs)
        assert mrec.tobytes() == whole.rec.tobytes() and np.array_equal(mtr, whole.trace) and np.array_equal(mtoff, whole.toff)


def test_config1_at_scale_matches_oracle():
    """BASELINE configs[1] (the bench workload, k = 20, -s100 -l1000) at scale 0.2: 2 Mbp assembly x 43 Mbp of reads,
    every record and trace point against the oracle (all host threads share the oracle's A index)."""
    import os
    import bench
    from dentist_b200 import dazzler
    from oracle import oracle
    ref, reads = bench.make_workload(0.2, 0)
    assert reads.total > 40e6
    la, tr, st = oracle.align(ref.off, ref.bases, reads.off, reads.bases, threads=os.cpu_count() or 1,
                              tspace=bench.PARAMS["tspace"], minlen=bench.PARAMS["minlen"], **bench.ORC)
    rec, toff, gtr, gst = dazzler.align_blocks(dazzler.Block(ref.off, ref.bases), dazzler.Block(reads.off, reads.bases), **bench.PARAMS)
    assert len(la) > 8000
    assert_same((la, tr, st), (rec, toff, gtr, gst))


@pytest.mark.parametrize("k", [14, 20])
def test_low_complexity_takes_every_index_build_path(k):
    """The bucket-counting index build (csrc/bucket.cu) orders buckets of 2, of <= 32 (thread-local insertion sort) and of
    <= 4096 tuples (CTA bitonic sort) in place and hands blocks with a larger bucket to the radix sort.  A reference with a
    tandem repeat (buckets of hundreds) and, in the second round, a 12 kb homopolymer run (one bucket of thousands: the
    radix fallback) must still give the oracle's alignments; -t caps what the repeats may seed."""
    rng = np.random.default_rng(77 + k)
    g = rng.integers(0, 4, 60000, dtype=np.uint8)
    unit = rng.integers(0, 4, 37, dtype=np.uint8)
    for with_run in (False, True):
        ref = g.copy()
        ref[10000:10000 + 37 * 60] = np.tile(unit, 60)                       # 60 copies of a 37-mer: every k-mer of it 60 times (minus the ends)
        if with_run:
            ref[30000:36000] = 0                                             # 6 000 x 'a': one k-mer ~6 000 times (> 4096: radix fallback)
        a = synth.Block(np.array([0, len(ref)]), ref)
        reads, _ = synth.simulate_reads([ref], 5, 5000, 1500, 0.12, 78 + k)
        for t in ((100,) if with_run else (100, 100000)):                    # the homopolymer run only under the frequency cap
            orc, gpu = run_both(a, reads, 100, 500, k=k, t=t)
            assert len(orc[0]) > 20
            assert_same(orc, gpu)
