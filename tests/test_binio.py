"""PileUpDb / InsertionDb codecs against the reference's own test data and size identities
(common/binio/pileupdb.d:430-464, 502-526; insertiondb.d:470-511, 832-860; binio/common.d:449-458)."""
import json
import os

import numpy as np
import pytest

from dentist_b200 import binio

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "binio_kat.json")))
CODE = {"a": 0, "c": 1, "g": 2, "t": 3}


def _seeded(d):
    return dict(id=d["id"], contigA=tuple(d["contigA"]), contigB=tuple(d["contigB"]), flags=d["flags"], tpd=d["tpd"], seed=d["seed"],
                las=[dict(ab=l["ab"], ae=l["ae"], bb=l["bb"], be=l["be"], diffs=l["diffs"], trace=np.array(l["trace"], np.uint16).reshape(-1, 2))
                     for l in d["las"]])


def _same_seeded(x, y):
    for k in ("id", "contigA", "contigB", "flags", "tpd", "seed"):
        assert x[k] == y[k], k
    assert len(x["las"]) == len(y["las"])
    for a, b in zip(x["las"], y["las"]):
        assert [a[k] for k in ("ab", "ae", "bb", "be", "diffs")] == [b[k] for k in ("ab", "ae", "bb", "be", "diffs")]
        assert np.array_equal(a["trace"], b["trace"])


def test_pileup_db_size_pointers_and_round_trip(tmp_path):
    c = KAT["pileupdb"]["counts"]
    piles = [[[_seeded(s) for s in ra] for ra in p] for p in KAT["pileupdb"]["data"]]
    assert len(piles) == c["numPileUps"] and sum(len(p) for p in piles) == c["numReadAlignments"]
    assert sum(len(ra) for p in piles for ra in p) == c["numSeededAlignments"]
    path = str(tmp_path / "pileups.db")
    binio.write_pileup_db(path, piles)
    total = 48 + 16 * c["numPileUps"] + 16 * c["numReadAlignments"] + 56 * c["numSeededAlignments"] + 40 * c["numLocalAlignments"] + 4 * c["numTracePoints"]
    assert os.path.getsize(path) == total                                         # pileupdb.d:440-458
    idx = np.fromfile(path, "<u8", 6)
    assert idx[0] == 48 and idx[1] == idx[0] + 16 * c["numPileUps"] and idx[2] == idx[1] + 16 * c["numReadAlignments"]   # :510-525
    assert idx[3] == idx[2] + 56 * c["numSeededAlignments"] and idx[4] == idx[3] + 40 * c["numLocalAlignments"] and idx[5] == total
    back = binio.read_pileup_db(path)
    assert [len(p) for p in back] == [len(p) for p in piles]
    for p, q in zip(piles, back):
        for ra, rb in zip(p, q):
            assert len(ra) == len(rb)
            for s, t in zip(ra, rb):
                _same_seeded(s, t)


def test_insertion_db_size_and_round_trip(tmp_path):
    c = KAT["insertiondb"]["counts"]
    ins = [dict(start=tuple(i["start"]), end=tuple(i["end"]), sequence=np.array([CODE[ch] for ch in i["sequence"]], np.uint8),
                contig_length=i["contig_length"], overlaps=[_seeded(s) for s in i["overlaps"]], read_ids=i["read_ids"])
           for i in KAT["insertiondb"]["data"]]
    assert len(ins) == c["numInsertions"] and sum(len(i["overlaps"]) for i in ins) == c["numOverlaps"]
    assert sum((len(i["sequence"]) + 3) // 4 for i in ins) == c["numCompressedBaseQuads"] and sum(len(i["read_ids"]) for i in ins) == c["numReadIds"]
    path = str(tmp_path / "insertions.db")
    binio.write_insertion_db(path, ins)
    total = (56 + 104 * c["numInsertions"] + c["numCompressedBaseQuads"] + 56 * c["numOverlaps"] + 40 * c["numLocalAlignments"] +
             4 * c["numTracePoints"] + 4 * c["numReadIds"])
    assert os.path.getsize(path) == total                                         # insertiondb.d:481-500
    back = binio.read_insertion_db(path)
    assert len(back) == len(ins)
    for a, b in zip(ins, back):
        assert a["start"] == b["start"] and a["end"] == b["end"] and a["contig_length"] == b["contig_length"]
        assert np.array_equal(a["sequence"], b["sequence"]) and list(a["read_ids"]) == list(b["read_ids"])
        for s, t in zip(a["overlaps"], b["overlaps"]):
            _same_seeded(s, t)


def test_compressed_sequence_kat():
    s = "atgccaactactttgaacgcgCCGCAAGGCACAGGTGCGCCT".lower()                       # binio/common.d:451
    q = binio.compress_sequence(np.array([CODE[ch] for ch in s], np.uint8))
    dentist = {"a": 0, "c": 1, "t": 2, "g": 3}                                        # CompressedBase, binio/common.d:325-331
    for i, ch in enumerate(s):
        assert (q[i // 4] >> (2 * (i % 4))) & 3 == dentist[ch]
    assert len(q) == (len(s) + 3) // 4 and q[0] == (0 | 2 << 2 | 3 << 4 | 1 << 6)
    assert np.array_equal(binio.decompress_sequence(q, len(s)), [CODE[ch] for ch in s])
    assert np.array_equal(binio.decompress_sequence(q, 10, base_offset=3), [CODE[ch] for ch in s[3:13]])


def test_truncated_files_raise(tmp_path):
    p = str(tmp_path / "bad.db")
    open(p, "wb").write(b"\0" * 20)
    with pytest.raises(binio.BinioError):
        binio.read_pileup_db(p)
    with pytest.raises(binio.BinioError):
        binio.read_insertion_db(p)
