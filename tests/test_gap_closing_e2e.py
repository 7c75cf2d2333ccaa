"""The whole hot path on a synthetic assembly: map reads to contigs (dn_damap path) -> collect filters ->
pile-ups of gap-spanning reads -> crop at common trace points on the device -> batched processPileUp
(align, filter, chain, QVs, reference read, consensus) -> the consensus closes the gap: it contains the true gap
sequence with >= 97 % identity although every read carries 13 % errors."""
import numpy as np
import pytest

from dentist_b200 import pileups, synth

pytestmark = pytest.mark.gpu


def _identity(a, b):
    """1 - edit distance / len, banded, via numpy row DP (sequences ~1-3 kb)."""
    prev = np.arange(len(b) + 1)
    for i in range(1, len(a) + 1):
        sub = prev[:-1] + (a[i - 1] != b)
        cur = np.minimum(sub, prev[1:] + 1)
        cur = np.concatenate([[i], cur])
        for j in range(1, len(b) + 1):              # insertion chain
            if cur[j - 1] + 1 < cur[j]:
                cur[j] = cur[j - 1] + 1
        prev = cur
    return 1.0 - prev[-1] / max(len(a), len(b))


def test_gap_closing_end_to_end():
    from dentist_b200 import dazzler
    sc = synth.make_scaffolds(1, 300000, 901, n_repeats=0)
    gaps = synth.make_gaps(sc, 3, 902, min_len=300, max_len=900)
    ref, meta = synth.contigs_from(sc, gaps)
    reads, truth = synth.simulate_reads(sc, 25, 9000, 2500, 0.13, 903)
    alen, blen = np.diff(ref.off), np.diff(reads.off)
    ga, gb = dazzler.Block(ref.off, ref.bases), dazzler.Block(reads.off, reads.bases)
    las = dazzler.align(ga, gb, tspace=100, minlen=500)                      # ref vs reads (damapper role)
    las.chainMapper(reads.nreads)
    first, st, used = dazzler.collectFilter(las, alen, blen, max_alignment_error=0.3, proper_alignment_allowance=100, min_anchor_length=500)
    kept = first[st == 0]
    rec = las.rec
    traces = las.traces()
    crop_read, crop_beg, crop_end, crop_group = [], [], [], []
    gap_seqs = []
    for g, (gb_, ge_) in enumerate(gaps[0]):
        L, R = g, g + 1                                                     # flanking contigs of gap g
        back = {int(rec[i]["bread"]): i for i in kept if rec[i]["aread"] == L and rec[i]["aepos"] + 100 >= alen[L]}
        front = {int(rec[i]["bread"]): i for i in kept if rec[i]["aread"] == R and rec[i]["abpos"] <= 100}
        span = sorted(set(back) & set(front))
        span = [b for b in span if (rec[back[b]]["flags"] & 1) == (rec[front[b]]["flags"] & 1)]
        assert len(span) >= 6, (g, len(span))
        cL = pileups.common_trace_point([(int(rec[back[b]]["abpos"]), int(rec[back[b]]["aepos"])) for b in span], "back", 100, int(alen[L]))
        cR = pileups.common_trace_point([(int(rec[front[b]]["abpos"]), int(rec[front[b]]["aepos"])) for b in span], "front", 100, int(alen[R]))
        assert cL >= 0 and cR >= 0
        for b in span:
            i, j = back[b], front[b]
            comp = bool(rec[i]["flags"] & 1)
            lb, le = pileups.get_cropping_slice(int(rec[i]["abpos"]), int(rec[i]["aepos"]), int(rec[i]["bbpos"]), 100, traces[i], comp, "back", int(blen[b]), cL)
            rb, re_ = pileups.get_cropping_slice(int(rec[j]["abpos"]), int(rec[j]["aepos"]), int(rec[j]["bbpos"]), 100, traces[j], comp, "front", int(blen[b]), cR)
            beg, end = max(lb, rb), min(le, re_)                            # the part kept by both crops
            assert end - beg > (ge_ - gb_), (b, beg, end)
            crop_read.append(b); crop_beg.append(beg); crop_end.append(end); crop_group.append(g)
        # truth of the cropped region: from cL on the left contig to cR on the right contig
        s, lbeg = meta[L]; _, rbeg = meta[R]
        gap_seqs.append(sc[0][lbeg + cL:rbeg + cR])
    cropped = dazzler.Block.crop(gb, crop_read, crop_beg, crop_end, group=crop_group)
    class _B:                                                                # process_pileups wants host arrays for the block
        pass
    seqs = [reads.read(r)[b:e] for r, b, e in zip(crop_read, crop_beg, crop_end)]
    off = np.zeros(len(seqs) + 1, np.int64); off[1:] = np.cumsum([len(x) for x in seqs])
    host = synth.Block(off, np.concatenate(seqs))
    assert cropped.bases == host.total
    res = pileups.process_pileups(host, np.array(crop_group, np.int32), flanks=ref)
    assert len(res["consensus"]) == len(gaps[0])
    for g, cons in enumerate(res["consensus"]):
        t = gap_seqs[g]
        raw = host.read(res["reference_read"][g])
        best_c = max(_identity(cons, t), _identity((3 - cons)[::-1], t))
        best_r = max(_identity(raw, t), _identity((3 - raw)[::-1], t))
        assert best_r < 0.92 and best_c >= 0.97, (g, best_r, best_c)
    fl = res["flank_las"].rec
    for g in range(len(gaps[0])):                                            # each consensus anchors in both flanking contigs
        assert {g, g + 1} <= set(fl["aread"][fl["bread"] == g].tolist())
