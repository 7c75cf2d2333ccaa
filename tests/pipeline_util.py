"""`dentist collect | dentist process` in miniature, shared by the end-to-end tests and tools/run_example.py: map the
reads to the contigs (damapper role) -> collectPileUps' filters -> pile-ups -> PileUpDb -> batched processPileUps ->
InsertionDb, for ONE scaffold whose contigs and gaps alternate (contig g, gap g, contig g + 1)."""
import numpy as np

from dentist_b200 import binio, process


def identity(a, b):
    """1 - edit distance / len via numpy row DP (sequences of a few kb)."""
    prev = np.arange(len(b) + 1)
    for i in range(1, len(a) + 1):
        cur = np.minimum(prev[:-1] + (a[i - 1] != b), prev[1:] + 1)
        cur = np.concatenate([[i], cur])
        cur = np.minimum.accumulate(cur - np.arange(len(cur))) + np.arange(len(cur))     # insertion chain
        prev = cur
    return 1.0 - prev[-1] / max(len(a), len(b))


def build_pileups(las, alen, blen, kept, n_gaps, tspace=100):
    """The kept chains of every read become read alignments (collectReadAlignments, collectPileUps/pileups.d:821-888,
    with SeededAlignment.from seeds); read alignments with the same join -- gap (contig g end, contig g+1 begin) or an
    extension into that gap from either side -- form one pile-up (what bundling the scaffold graph's edges does,
    pileups.d:650-666)."""
    rec, toff, trace = las.rec, las.toff, las.trace
    by_read = {}
    for i in kept:
        j = i + 1
        while j < len(rec) and (int(rec[j]["flags"]) & 0x8):
            j += 1
        sub = slice(int(i), j)
        chain = binio.seeded_alignments_from_las(rec[sub], toff[sub], trace, alen, blen, tspace, lambda f, l: "front")[0]
        chain.pop("seed"); chain["id"] = int(i)
        by_read.setdefault(int(rec[i]["bread"]), []).append(chain)
    piles = [[] for _ in range(n_gaps)]
    for b in sorted(by_read):
        ras, reason = process.collect_read_alignments(by_read[b])
        for ra in ras:
            if process._is_gap(ra) and not process.is_in_order(ra):
                ra = [ra[1], ra[0]]                                   # ReadAlignment.getInOrder, base.d:2177-2183
            start, end = process.make_join(ra)
            if process._is_gap(ra):
                if start[1] == "end" and end[1] == "begin" and end[0] == start[0] + 1 and process._is_parallel(ra):
                    piles[start[0] - 1].append(ra)
            elif ra[0]["seed"] == "back" and ra[0]["contigA"][0] <= n_gaps:
                piles[ra[0]["contigA"][0] - 1].append(ra)             # extends contig g beyond its end, into gap g
            elif ra[0]["seed"] == "front" and ra[0]["contigA"][0] >= 2:
                piles[ra[0]["contigA"][0] - 2].append(ra)             # extends contig g+1 beyond its begin, into gap g
    return piles


def close_gaps(ref, reads, n_gaps, tmpdir, k=14, minlen=500, tspace=100):
    """ref / reads: synth.Block.  Returns (insertions read back from the InsertionDb, skipped, pile-ups)."""
    from dentist_b200 import dazzler
    alen, blen = np.diff(ref.off), np.diff(reads.off)
    ga, gb = dazzler.Block(ref.off, ref.bases), dazzler.Block(reads.off, reads.bases)
    las = dazzler.align(ga, gb, tspace=tspace, minlen=minlen, k=k)             # ref vs reads (damapper role)
    las.chainMapper(reads.nreads)
    first, st, used = dazzler.collectFilter(las, alen, blen, max_alignment_error=0.3, proper_alignment_allowance=tspace, min_anchor_length=500)
    piles = build_pileups(las, alen, blen, first[st == 0], n_gaps, tspace)
    db = str(tmpdir) + "/pileups.db"
    binio.write_pileup_db(db, piles)
    insertions, skipped = process.process_pileup_db(binio.read_pileup_db(db), reads, ref, proper_alignment_allowance=tspace)
    out = str(tmpdir) + "/insertions.db"
    binio.write_insertion_db(out, insertions)
    return binio.read_insertion_db(out), skipped, piles, dict(las=len(las), chains=len(first), kept=int((st == 0).sum()))


def gap_identity(ins, scaffold, meta, g):
    """Identity of an insertion's consensus, cut between its two anchors' inner ends, with the true sequence there."""
    left, right = ins["overlaps"]
    comp = bool(left["flags"] & 1)
    cons = ins["sequence"] if not comp else (3 - ins["sequence"])[::-1]
    _, lbeg = meta[g]; _, rbeg = meta[g + 1]
    t = scaffold[lbeg + left["las"][0]["ab"]:rbeg + right["las"][-1]["ae"]]
    return identity(cons, t), len(cons), len(t)
