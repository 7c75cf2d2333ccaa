"""A14: cropping slices (host logic, pinned by the reference's KATs cropper.d:552-647) and the device crop kernel."""
import json
import os

import numpy as np
import pytest

from dentist_b200 import pileups, synth

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cropper_kat.json")))


def test_get_cropping_slice_reference_kats():
    assert len(KAT) == 2 and sum(len(c["asserts"]) for c in KAT) == 5
    for c in KAT:
        for pos, b, e in c["asserts"]:
            got = pileups.get_cropping_slice(c["abpos"], c["aepos"], c["bbpos"], c["tspace"], c["trace"], c["complement"], c["seed"],
                                             c["read_len"], pos)
            assert got == (b, e), (c["seed"], pos)


def test_translate_trace_point_agrees_with_oracle_on_reference_trace():
    from oracle import las
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "las_golden.json")))["trace_kat"]
    for pos in range(g["abpos"], g["aepos"] + 1, 37):
        assert pileups.translate_trace_point(g["abpos"], g["aepos"], g["bbpos"], g["tspace"], g["trace"], pos) == \
            las.translate_trace_point(g["abpos"], g["aepos"], g["bbpos"], g["tspace"], g["trace"], pos, "floor")


@pytest.mark.gpu
def test_device_crop_equals_host_slicing():
    from dentist_b200 import dazzler
    sc = synth.make_scaffolds(1, 30000, 71, n_repeats=0)
    reads, _ = synth.simulate_reads(sc, 8, 6000, 1500, 0.13, 72)
    rng = np.random.default_rng(4)
    n = reads.nreads
    lens = np.diff(reads.off)
    sel = rng.permutation(n)[: n - 2]
    beg = np.array([int(rng.integers(0, lens[r] // 3)) for r in sel])
    end = np.array([int(lens[r] - rng.integers(0, lens[r] // 3)) for r in sel])
    beg[0], end[0] = 0, lens[sel[0]]                 # whole read
    beg[1], end[1] = 17, 17                          # empty slice
    grp = (np.arange(len(sel)) % 2).astype(np.int32)
    src = dazzler.Block(reads.off, reads.bases)
    cropped = dazzler.Block.crop(src, sel, beg, end, group=grp)
    seqs = [reads.read(r)[b:e] for r, b, e in zip(sel, beg, end)]
    off = np.zeros(len(seqs) + 1, np.int64); off[1:] = np.cumsum([len(s) for s in seqs])
    host = dazzler.Block(off, np.concatenate(seqs), group=grp)
    assert cropped.bases == host.bases == int(off[-1])
    a = dazzler.align(cropped, cropped, tspace=126, minlen=500, self_block=1)
    b = dazzler.align(host, host, tspace=126, minlen=500, self_block=1)
    assert len(a) > 50 and a.rec.tobytes() == b.rec.tobytes() and np.array_equal(a.trace, b.trace)
    assert (a.rec["flags"] & 1).any() and (~(a.rec["flags"] & 1).astype(bool)).any()      # both strands exercised
    with pytest.raises(dazzler.DnError, match="outside the read"):
        dazzler.Block.crop(src, [0], [5], [int(lens[0]) + 1])
