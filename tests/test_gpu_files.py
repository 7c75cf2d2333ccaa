"""File-level drop-ins: getDalignment / getDamapping read DAZZ_DB files and write LAS files that the
reference's reader would accept (checked with the pinned oracle codec) and that equal the in-memory result."""
import numpy as np
import pytest

from dentist_b200 import synth
from tests import dbutil

pytestmark = pytest.mark.gpu


def _case(seed):
    sc = synth.make_scaffolds(1, 80000, seed, n_repeats=0)
    ref, _ = synth.contigs_from(sc, synth.make_gaps(sc, 1, seed + 1))
    reads, _ = synth.simulate_reads(sc, 4, 6000, 2000, 0.13, seed + 2)
    return ref, reads


def test_getDamapping_writes_both_las_files(tmp_path):
    from dentist_b200 import dazzler
    from oracle import las
    ref, reads = _case(51)
    dbutil.write_db(str(tmp_path / "ref.dam"), ref)
    dbutil.write_db(str(tmp_path / "reads.db"), reads)
    out = dazzler.getDamapping(str(tmp_path / "ref.dam"), str(tmp_path / "reads.db"), ["-C", "-T8", "-e0.7", "-l500"], str(tmp_path))
    assert out == str(tmp_path / "ref.reads.las")
    ts, rec, traces = las.decode(open(out, "rb").read())
    assert ts == 100 and len(rec) > 20
    ga, gb = dazzler.Block(ref.off, ref.bases), dazzler.Block(reads.off, reads.bases)
    mem = dazzler.align(ga, gb, tspace=100, minlen=500, k=20)                            # damapper's default k
    allrec, alltr = mem.rec.copy(), [t.copy() for t in mem.traces()]
    # damapper reports the best chain of every read only (no -n, commandline.d:2943-2955; dazzler.d:5920-5923)
    from oracle import chain_oracle
    keep = chain_oracle.keep_best_chains(allrec, chain_oracle.mapper_chain_flags(allrec))
    assert 20 < len(keep) <= len(allrec)
    mrec = allrec[keep]
    for f in ("aread", "bread", "abpos", "aepos", "bbpos", "bepos", "diffs", "tlen"):
        assert np.array_equal(rec[f], mrec[f]), f
    assert np.concatenate([t.reshape(-1) for t in traces]).tolist() == np.concatenate([alltr[i].reshape(-1) for i in keep]).tolist()
    assert sorted(set(rec["bread"].tolist())) == sorted(set(allrec["bread"].tolist()))     # every mapped read keeps exactly its best chain
    assert all(int(f) & las.BEST for f in rec["flags"] if int(f) & las.START)
    # -n<f>: chains within the fraction f of the best come back too; -n0.01 ~ everything
    (tmp_path / "n").mkdir()
    outn = dazzler.getDamapping(str(tmp_path / "ref.dam"), str(tmp_path / "reads.db"), ["-C", "-e0.7", "-l500", "-n0.01"], str(tmp_path / "n"))
    _, recn, _ = las.decode(open(outn, "rb").read())
    keepn = chain_oracle.keep_best_chains(allrec, chain_oracle.mapper_chain_flags(allrec), 0.01)
    assert len(recn) == len(keepn) >= len(rec) and np.array_equal(recn["abpos"], allrec[keepn]["abpos"])
    assert all(int(f) & (las.START | las.NEXT) for f in rec["flags"])              # chain flags set -> AlignmentChainPacker accepts it
    assert len(las.chains(rec)) == len(rec)
    # transposed file exists, is sorted by read and has uint8 traces
    ts2, rec2, tr2 = las.decode(open(str(tmp_path / "reads.ref.las"), "rb").read())
    assert len(rec2) > 20 and rec2["aread"].tolist() == sorted(rec2["aread"].tolist())
    for r, t in zip(rec2, tr2):
        assert int(t[:, 1].sum()) == r["bepos"] - r["bbpos"] and len(t) == las.num_tiles(int(r["abpos"]), int(r["aepos"]), 100)
    # the ABI's own reader agrees with the oracle reader
    ts3, rec3, toff3, tr3 = dazzler.read_las(out)
    assert ts3 == 100 and rec3.tobytes() == rec.tobytes()


def test_getDalignment_pile_with_dust_mask(tmp_path):
    """processPileUps: `daligner -T<n> -B -s126 -l500 -e0.7 -mdust X X` (commandline.d:2886-2902)."""
    from dentist_b200 import dazzler
    from oracle import las, oracle
    sc = synth.make_scaffolds(1, 20000, 61, n_repeats=0)
    pile, _ = synth.simulate_reads(sc, 8, 6000, 1500, 0.13, 62)
    db = str(tmp_path / "pile.db")
    dbutil.write_db(db, pile)
    mask = [[(200, 260), (1000, 1100)] for _ in range(pile.nreads)]
    dbutil.write_track(db, "dust", mask)
    out = dazzler.getDalignment(db, None, ["-T8", "-B", "-s126", "-l500", "-e0.7", "-mdust"], str(tmp_path))
    assert out.endswith("pile.pile.las")
    ts, rec, traces = las.decode(open(out, "rb").read())
    assert ts == 126
    m = np.zeros(pile.total, np.uint8)
    for r in range(pile.nreads):
        for b, e in mask[r]:
            m[pile.off[r] + b:pile.off[r] + e] = 1
    la, tr, _ = oracle.align(pile.off, pile.bases, pile.off, pile.bases, a_mask=m, b_mask=m, tspace=126, minlen=500, self=1)
    la, _, tr, _ = oracle.bridge(pile.off, pile.bases, pile.off, pile.bases, la, la["toff"].astype(np.int64), tr, 126)      # -B
    assert len(la) == len(rec) > 50
    for f in ("aread", "bread", "abpos", "aepos", "bbpos", "bepos", "diffs", "flags"):
        assert np.array_equal(rec[f], la[f]), f
    assert np.concatenate([t.reshape(-1) for t in traces]).tolist() == tr.tolist()
    assert not np.any(rec["aread"] == rec["bread"])


def test_errors_are_reported_not_swallowed(tmp_path):
    from dentist_b200 import dazzler
    with pytest.raises(dazzler.DnError, match="cannot (find|open)"):
        dazzler.getDalignment(str(tmp_path / "nope.db"), None, ["-s126"], str(tmp_path))
    ref, reads = _case(71)
    dbutil.write_db(str(tmp_path / "r.db"), reads)
    with pytest.raises(dazzler.DnError, match="unknown option"):
        dazzler.getDalignment(str(tmp_path / "r.db"), None, ["-Q3"], str(tmp_path))


def test_dbdust_writes_a_track_that_mdust_reads(tmp_path):
    from dentist_b200 import dazzler
    from oracle import dust
    import struct
    rng = np.random.default_rng(3)
    seqs = [rng.integers(0, 4, 2000, dtype=np.uint8) for _ in range(4)]
    seqs[1][300:420] = 3
    off = np.zeros(5, np.int64); off[1:] = np.cumsum([len(s) for s in seqs])
    blk = synth.Block(off, np.concatenate(seqs))
    db = str(tmp_path / "x.db")
    dbutil.write_db(db, blk)
    dazzler.dbdustFile(db, ["-w64", "-t2.0", "-m10"])
    anno = open(str(tmp_path / ".x.dust.anno"), "rb").read()
    data = open(str(tmp_path / ".x.dust.data"), "rb").read()
    n, size = struct.unpack_from("<ii", anno, 0)
    assert (n, size) == (4, 0)                                     # mask track header, dazzler.d:4943-4975
    offs = struct.unpack_from("<5q", anno, 8)
    got = [[struct.unpack_from("<ii", data, o) for o in range(offs[r], offs[r + 1], 8)] for r in range(4)]
    assert got == dust.dust_block(blk.off, blk.bases) and got[1] and not got[0]
    # -mdust now changes the seeds
    out = dazzler.getDalignment(db, None, ["-s126", "-l500", "-mdust"], str(tmp_path))
    assert out.endswith("x.x.las")


def test_dbdust_on_a_db_block_writes_the_block_level_track(tmp_path):
    """the workflow dusts DB blocks one by one (`DBdust X.<n>` writes .X.<n>.dust.*, merged later by Catrack;
    getMaskFiles dazzler.d:4870-4912).  dbdust on a block path must not overwrite the whole-DB track, and an alignment of
    that block with -mdust must find the block-level track."""
    from dentist_b200 import dazzler
    from oracle import dust
    import os
    import struct
    rng = np.random.default_rng(4)
    seqs = [rng.integers(0, 4, 2500, dtype=np.uint8) for _ in range(6)]
    seqs[4][300:460] = 2
    off = np.zeros(7, np.int64); off[1:] = np.cumsum([len(s) for s in seqs])
    blk = synth.Block(off, np.concatenate(seqs))
    db = str(tmp_path / "y.db")
    dbutil.write_db(db, blk, block_first=[3])                      # blocks: reads 0-2 and 3-5
    dbutil.write_track(db, "dust", [[] for _ in range(6)])         # a whole-DB track that must survive
    whole_before = open(str(tmp_path / ".y.dust.anno"), "rb").read()
    dazzler.dbdustFile(str(tmp_path / "y.2.db"))
    assert open(str(tmp_path / ".y.dust.anno"), "rb").read() == whole_before
    anno = open(str(tmp_path / ".y.2.dust.anno"), "rb").read(); data = open(str(tmp_path / ".y.2.dust.data"), "rb").read()
    n, size = struct.unpack_from("<ii", anno, 0)
    assert (n, size) == (3, 0)
    offs = struct.unpack_from("<4q", anno, 8)
    got = [[struct.unpack_from("<ii", data, o) for o in range(offs[r], offs[r + 1], 8)] for r in range(3)]
    assert got == dust.dust_block(blk.off, blk.bases)[3:] and got[1] and not got[0]
    # the block alignment with -mdust reads .y.2.dust (the whole-DB track is empty): fewer seed hits than without a mask
    sub = synth.Block(off[3:] - off[3], blk.bases[off[3]:])
    masked = dazzler.align_blocks(dazzler.Block(sub.off, sub.bases, mask=got), dazzler.Block(sub.off, sub.bases, mask=got), tspace=126, minlen=500, self_block=1, identity=1)
    out = dazzler.getDalignment(str(tmp_path / "y.2.db"), None, ["-s126", "-l500", "-mdust", "-I"], str(tmp_path))
    assert os.path.basename(out) == "y.2.y.2.las"
    _, rec, _, _ = dazzler.read_las(out)
    assert rec.tobytes() == masked[0].tobytes() and len(rec) >= 3


def test_getConsensus_file_form_on_the_reference_kat(tmp_path):
    """dazzler.d:4257-4299 through files: buildDamFile -> daligner -l15 -> filterPileUpAlignments -> getConsensus
    -> the consensus .dam holds exactly one read, equal to read 3."""
    import struct
    from dentist_b200 import dazzler
    from tests.test_pile_oracle import kat_block
    k, blk = kat_block()
    db = str(tmp_path / "kat.dam")
    dbutil.write_db(db, blk)
    las_path = dazzler.getDalignment(db, None, ["-l%d" % k["minlen"], "-s100"], str(tmp_path))
    ts, rec, toff, tr = dazzler.read_las(las_path)
    assert len(rec) == 6
    out = dazzler.getConsensusDb(db, las_path, 1)
    assert out == str(tmp_path / "kat-daccord-I0-0.dam")
    idx = open(str(tmp_path / ".kat-daccord-I0-0.idx"), "rb").read()
    bps = open(str(tmp_path / ".kat-daccord-I0-0.bps"), "rb").read()
    ureads, = struct.unpack_from("<i", idx, 0)
    rlen, = struct.unpack_from("<i", idx, 112 + 4)
    assert ureads == 1 and rlen == 1050
    codes = np.array([(b >> s) & 3 for b in bps for s in (6, 4, 2, 0)], np.uint8)[:rlen]
    assert np.array_equal(codes, blk.read(k["expected_read"]))
    with pytest.raises(dazzler.DnError, match="out of bounds"):
        dazzler.getConsensusDb(db, las_path, 9)


def test_cli_damapper_equals_the_library_call(tmp_path):
    """`dn-damapper -C -e0.7 R Q` run in a directory leaves the same R.Q.las / Q.R.las there as getDamapping."""
    import os
    import subprocess
    from dentist_b200 import dazzler
    ref, reads = _case(57)
    (tmp_path / "cli").mkdir(); (tmp_path / "api").mkdir()
    dbutil.write_db(str(tmp_path / "ref.dam"), ref)
    dbutil.write_db(str(tmp_path / "reads.db"), reads)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([os.path.join(root, "bin", "dn-damapper"), "-C", "-T4", "-e0.7", str(tmp_path / "ref.dam"), str(tmp_path / "reads.db")],
                       cwd=str(tmp_path / "cli"), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    dazzler.getDamapping(str(tmp_path / "ref.dam"), str(tmp_path / "reads.db"), ["-C", "-T4", "-e0.7"], str(tmp_path / "api"))
    for name in ("ref.reads.las", "reads.ref.las"):
        a = open(str(tmp_path / "cli" / name), "rb").read(); b = open(str(tmp_path / "api" / name), "rb").read()
        assert len(a) > 1000 and a == b, name
    r = subprocess.run([os.path.join(root, "bin", "dn-damapper"), "-C", str(tmp_path / "missing.dam"), str(tmp_path / "reads.db")],
                       cwd=str(tmp_path / "cli"), capture_output=True, text=True)
    assert r.returncode == 1 and "missing" in r.stderr


def test_computeQVs_file_form_writes_the_qual_track(tmp_path):
    """computeQVs(db, las, coverage) dazzler.d:3782-3792 on files; the track is what package.d:520-523 reads back."""
    import struct
    from dentist_b200 import dazzler
    from oracle import las, oracle
    sc = synth.make_scaffolds(1, 20000, 71, n_repeats=0)
    pile, _ = synth.simulate_reads(sc, 10, 6000, 1500, 0.13, 72)
    db = str(tmp_path / "pile.db")
    dbutil.write_db(db, pile)
    out = dazzler.getDalignment(db, None, ["-T8", "-s126", "-l500", "-e0.7"], str(tmp_path))
    ts, rec, traces = las.decode(open(out, "rb").read())
    assert ts == 126 and len(rec) > 50
    lens = np.diff(pile.off)
    toff = np.concatenate([[0], np.cumsum([2 * len(t) for t in traces])]).astype(np.int64)
    tr = np.concatenate([t.reshape(-1) for t in traces]).astype(np.uint16)
    for cov in (4, pile.nreads):
        dazzler.computeQVsDb(db, out, cov)
        qvs = dazzler.getIntrinsicQVs(db)
        oq, ooff = oracle.qv(lens, rec, toff[:-1], tr, 126, cov)
        assert len(qvs) == pile.nreads
        for r in range(pile.nreads):
            assert len(qvs[r]) == -(-int(lens[r]) // 126)
            assert np.array_equal(qvs[r], oq[ooff[r]:ooff[r + 1]]), r
    # the files themselves: int32 nreads, int32 0, nreads + 1 int64 offsets | one byte per tile
    anno = open(str(tmp_path / ".pile.qual.anno"), "rb").read()
    data = open(str(tmp_path / ".pile.qual.data"), "rb").read()
    n, sz = struct.unpack("<ii", anno[:8])
    offs = struct.unpack("<%dq" % (n + 1), anno[8:])
    assert n == pile.nreads and sz == 0 and offs[0] == 0 and offs[-1] == len(data) == sum(-(-int(l) // 126) for l in lens)
    assert max(data) <= 50
    # a missing track is an error, not an empty answer
    dbutil.write_db(str(tmp_path / "other.db"), pile)
    with pytest.raises(dazzler.DnError):
        dazzler.getIntrinsicQVs(str(tmp_path / "other.db"))
