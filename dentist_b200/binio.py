"""PileUpDb / InsertionDb codecs -- the two binary files either side of `dentist process`
(SURVEY §8f.4): `collect` writes pile-ups (common/binio/pileupdb.d), `process` writes insertions
(common/binio/insertiondb.d).  With them the batched pile-up path can be driven from a real pile-up DB and its
result handed to `dentist output` without DENTIST's own process step.

Layout (all little endian, D structs follow the C ABI):
  index            N x size_t block pointers                                   pileupdb.d:710-745 / insertiondb.d:738-790
  ArrayStorage     {size_t ptr (absolute file offset), size_t length}          binio/common.d:208-215
  SeededAlignment  56 B  {u64 id; u32 aId, aLen, bId, bLen; i8 flags; pad7; ArrayStorage las; u16 tpd; u8 seed; pad5}
                                                                              pileupdb.d:866-879
  LocalAlignment   40 B  {u32 ab, ae, bb, be, diffs; pad4; ArrayStorage tps}   pileupdb.d:882-892
  TracePoint        4 B  {u16 diffs, u16 bases}                                pileupdb.d:895-899
  Insertion       104 B  {ContigNode start, end (u64 id; u8 part; pad7); u8 baseOffset; pad7; u64 seqLen;
                          ArrayStorage quads; u64 contigLength; ArrayStorage overlaps; ArrayStorage readIds}
                                                                              insertiondb.d:987-997, scaffold.d:77-108
  CompressedBaseQuad 1 B, base i in bits 2i..2i+1, a0 c1 t2 g3                 binio/common.d:325-345 (bitfields: LSB first)
Flags are DENTIST's `Flag` (base.d:121-133): complement 1, disabled 2, alternateChain 4, chainContinuation 8, unchained 16.
"""
import numpy as np

ARR = np.dtype([("ptr", "<u8"), ("length", "<u8")])
SEEDED = np.dtype({"names": ["id", "contigAId", "contigALength", "contigBId", "contigBLength", "flags", "las_ptr", "las_len", "tpd", "seed"],
                   "formats": ["<u8", "<u4", "<u4", "<u4", "<u4", "i1", "<u8", "<u8", "<u2", "u1"],
                   "offsets": [0, 8, 12, 16, 20, 24, 32, 40, 48, 50], "itemsize": 56})
LOCAL = np.dtype({"names": ["ab", "ae", "bb", "be", "diffs", "tp_ptr", "tp_len"],
                  "formats": ["<u4", "<u4", "<u4", "<u4", "<u4", "<u8", "<u8"],
                  "offsets": [0, 4, 8, 12, 16, 24, 32], "itemsize": 40})
TP = np.dtype([("diffs", "<u2"), ("bases", "<u2")])
INSERTION = np.dtype({"names": ["start_id", "start_part", "end_id", "end_part", "base_offset", "seq_len", "seq_ptr", "seq_quads",
                                "contig_length", "ovl_ptr", "ovl_len", "rid_ptr", "rid_len"],
                      "formats": ["<u8", "u1", "<u8", "u1", "u1", "<u8", "<u8", "<u8", "<u8", "<u8", "<u8", "<u8", "<u8"],
                      "offsets": [0, 8, 16, 24, 32, 40, 48, 56, 64, 72, 80, 88, 96], "itemsize": 104})
PILEUP_INDEX_BYTES, INSERTION_INDEX_BYTES = 48, 56
CONTIG_PART = {"pre": 0, "begin": 1, "end": 2, "post": 3}
SEED = {"front": 0, "back": 1}
_DAZZ2DENTIST = np.array([0, 1, 3, 2], np.uint8)           # engine code a0 c1 g2 t3 <-> CompressedBase a0 c1 t2 g3 (involution)


class BinioError(Exception):
    """PileUpDbException / InsertionDbException (pileupdb.d:49-56, insertiondb.d:68-75)."""


def compress_sequence(bases):
    """engine base codes (a0 c1 g2 t3) -> CompressedBaseQuad bytes (CompressedSequence.from, binio/common.d:421-444)."""
    b = _DAZZ2DENTIST[np.asarray(bases, np.uint8)]
    pad = (-len(b)) % 4
    q = np.concatenate([b, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    return (q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)).astype(np.uint8)


def decompress_sequence(quads, length, base_offset=0):
    q = np.asarray(quads, np.uint8)
    b = np.stack([(q >> s) & 3 for s in (0, 2, 4, 6)], axis=1).reshape(-1)
    return _DAZZ2DENTIST[b[base_offset:base_offset + length]]


def _flatten_alignments(seeded):
    """seeded: list of dicts {id, contigA:(id,len), contigB:(id,len), flags, tpd, seed, las:[{ab,ae,bb,be,diffs,trace:(n,2)}]}
    -> (SEEDED array with las_len set, LOCAL array with tp_len set, TP array)"""
    sa = np.zeros(len(seeded), SEEDED)
    las, tps = [], []
    for i, s in enumerate(seeded):
        sa[i] = (s["id"], s["contigA"][0], s["contigA"][1], s["contigB"][0], s["contigB"][1], s.get("flags", 0), 0, len(s["las"]),
                 s.get("tpd", 0), SEED[s["seed"]] if isinstance(s["seed"], str) else s["seed"])
        for la in s["las"]:
            t = np.asarray(la["trace"], np.uint16).reshape(-1, 2)
            las.append((la["ab"], la["ae"], la["bb"], la["be"], la["diffs"], 0, len(t)))
            tps.append(t)
    la = np.array(las, LOCAL) if las else np.zeros(0, LOCAL)
    tp = np.zeros(sum(len(t) for t in tps), TP)
    if len(tp):
        cat = np.concatenate(tps)
        tp["diffs"], tp["bases"] = cat[:, 0], cat[:, 1]
    return sa, la, tp


def _link(lengths, ptr0, itemsize):
    """ArrayStorage pointers of consecutive children: ptr_i = ptr0 + itemsize * sum(lengths[:i])."""
    lengths = np.asarray(lengths, np.uint64)
    return np.uint64(ptr0) + np.uint64(itemsize) * (np.cumsum(lengths, dtype=np.uint64) - lengths)


def write_pileup_db(path, pileups):
    """writePileUpsDb (pileupdb.d:405-428).  pileups: list of pile-ups; a pile-up = list of read alignments; a read
    alignment = list of 1..2 seeded alignments (dicts, see _flatten_alignments)."""
    ras = [ra for p in pileups for ra in p]
    sa, la, tp = _flatten_alignments([s for ra in ras for s in ra])
    p_pu = PILEUP_INDEX_BYTES
    p_ra = p_pu + 16 * len(pileups)
    p_sa = p_ra + 16 * len(ras)
    p_la = p_sa + SEEDED.itemsize * len(sa)
    p_tp = p_la + LOCAL.itemsize * len(la)
    eof = p_tp + TP.itemsize * len(tp)
    pu = np.zeros(len(pileups), ARR); pu["length"] = [len(p) for p in pileups]; pu["ptr"] = _link(pu["length"], p_ra, 16)
    ra = np.zeros(len(ras), ARR); ra["length"] = [len(r) for r in ras]; ra["ptr"] = _link(ra["length"], p_sa, SEEDED.itemsize)
    sa["las_ptr"] = _link(sa["las_len"], p_la, LOCAL.itemsize)
    la["tp_ptr"] = _link(la["tp_len"], p_tp, TP.itemsize)
    with open(path, "wb") as f:
        f.write(np.array([p_pu, p_ra, p_sa, p_la, p_tp, eof], "<u8").tobytes())
        for block in (pu, ra, sa, la, tp):
            f.write(block.tobytes())
    return eof


def _nest_alignments(sa, la, tp, la_base, tp_base):
    out = []
    for s in sa:
        l0 = (int(s["las_ptr"]) - la_base) // LOCAL.itemsize
        las = []
        for r in la[l0:l0 + int(s["las_len"])]:
            t0 = (int(r["tp_ptr"]) - tp_base) // TP.itemsize
            t = tp[t0:t0 + int(r["tp_len"])]
            las.append(dict(ab=int(r["ab"]), ae=int(r["ae"]), bb=int(r["bb"]), be=int(r["be"]), diffs=int(r["diffs"]),
                            trace=np.stack([t["diffs"], t["bases"]], axis=1).astype(np.uint16)))
        out.append(dict(id=int(s["id"]), contigA=(int(s["contigAId"]), int(s["contigALength"])),
                        contigB=(int(s["contigBId"]), int(s["contigBLength"])), flags=int(s["flags"]), tpd=int(s["tpd"]),
                        seed="front" if s["seed"] == 0 else "back", las=las))
    return out


def read_pileup_db(path):
    """PileUpDb.parse + opIndex (pileupdb.d:119-253): the nested pile-ups of `write_pileup_db`."""
    raw = np.fromfile(path, np.uint8)
    if len(raw) < PILEUP_INDEX_BYTES:
        raise BinioError("pile-up DB truncated: no index")
    p_pu, p_ra, p_sa, p_la, p_tp, eof = (int(v) for v in raw[:PILEUP_INDEX_BYTES].view("<u8"))
    if not (p_pu == PILEUP_INDEX_BYTES <= p_ra <= p_sa <= p_la <= p_tp <= eof == len(raw)):
        raise BinioError("pile-up DB index is inconsistent with the file size")
    pu = raw[p_pu:p_ra].view(ARR); ra = raw[p_ra:p_sa].view(ARR)
    sa = raw[p_sa:p_la].view(SEEDED); la = raw[p_la:p_tp].view(LOCAL); tp = raw[p_tp:eof].view(TP)
    seeded = _nest_alignments(sa, la, tp, p_la, p_tp)
    out = []
    for p in pu:
        r0 = (int(p["ptr"]) - p_ra) // 16
        pile = []
        for r in ra[r0:r0 + int(p["length"])]:
            s0 = (int(r["ptr"]) - p_sa) // SEEDED.itemsize
            pile.append(seeded[s0:s0 + int(r["length"])])
        out.append(pile)
    return out


def write_insertion_db(path, insertions):
    """InsertionDbFileWriter.writeToFile (insertiondb.d:514-735).  insertions: list of dicts {start:(contigId, part),
    end:(contigId, part), sequence: engine base codes, contig_length, overlaps:[seeded alignment dicts], read_ids:[...]}."""
    quads = [compress_sequence(i["sequence"]) for i in insertions]
    sa, la, tp = _flatten_alignments([s for i in insertions for s in i["overlaps"]])
    rids = np.concatenate([np.asarray(i["read_ids"], "<u4") for i in insertions]) if insertions else np.zeros(0, "<u4")
    nq = sum(len(q) for q in quads)
    p_in = INSERTION_INDEX_BYTES
    p_q = p_in + INSERTION.itemsize * len(insertions)
    p_ov = p_q + nq
    p_la = p_ov + SEEDED.itemsize * len(sa)
    p_tp = p_la + LOCAL.itemsize * len(la)
    p_rid = p_tp + TP.itemsize * len(tp)
    eof = p_rid + 4 * len(rids)
    part = lambda p: CONTIG_PART[p] if isinstance(p, str) else int(p)
    ins = np.zeros(len(insertions), INSERTION)
    for k, i in enumerate(insertions):
        ins[k] = (i["start"][0], part(i["start"][1]), i["end"][0], part(i["end"][1]), 0, len(i["sequence"]), 0, len(quads[k]),
                  i.get("contig_length", 0), 0, len(i["overlaps"]), 0, len(i["read_ids"]))
    ins["seq_ptr"] = _link(ins["seq_quads"], p_q, 1)
    ins["ovl_ptr"] = _link(ins["ovl_len"], p_ov, SEEDED.itemsize)
    ins["rid_ptr"] = _link(ins["rid_len"], p_rid, 4)
    sa["las_ptr"] = _link(sa["las_len"], p_la, LOCAL.itemsize)
    la["tp_ptr"] = _link(la["tp_len"], p_tp, TP.itemsize)
    with open(path, "wb") as f:
        f.write(np.array([p_in, p_q, p_ov, p_la, p_tp, p_rid, eof], "<u8").tobytes())
        f.write(ins.tobytes())
        for q in quads:
            f.write(q.tobytes())
        for block in (sa, la, tp, rids):
            f.write(block.tobytes())
    return eof


def read_insertion_db(path):
    """InsertionDb.parse + opIndex (insertiondb.d:79-468)."""
    raw = np.fromfile(path, np.uint8)
    if len(raw) < INSERTION_INDEX_BYTES:
        raise BinioError("insertion DB truncated: no index")
    p_in, p_q, p_ov, p_la, p_tp, p_rid, eof = (int(v) for v in raw[:INSERTION_INDEX_BYTES].view("<u8"))
    if not (p_in == INSERTION_INDEX_BYTES <= p_q <= p_ov <= p_la <= p_tp <= p_rid <= eof == len(raw)):
        raise BinioError("insertion DB index is inconsistent with the file size")
    ins = raw[p_in:p_q].view(INSERTION)
    sa = raw[p_ov:p_la].view(SEEDED); la = raw[p_la:p_tp].view(LOCAL); tp = raw[p_tp:p_rid].view(TP)
    rids = raw[p_rid:eof].view("<u4")
    seeded = _nest_alignments(sa, la, tp, p_la, p_tp)
    names = {v: k for k, v in CONTIG_PART.items()}
    out = []
    for i in ins:
        q0 = int(i["seq_ptr"]); o0 = (int(i["ovl_ptr"]) - p_ov) // SEEDED.itemsize; r0 = (int(i["rid_ptr"]) - p_rid) // 4
        out.append(dict(start=(int(i["start_id"]), names[int(i["start_part"])]), end=(int(i["end_id"]), names[int(i["end_part"])]),
                        sequence=decompress_sequence(raw[q0:q0 + int(i["seq_quads"])], int(i["seq_len"]), int(i["base_offset"])),
                        contig_length=int(i["contig_length"]), overlaps=seeded[o0:o0 + int(i["ovl_len"])],
                        read_ids=rids[r0:r0 + int(i["rid_len"])].astype(np.uint32)))
    return out


def seeded_alignments_from_las(rec, toff, trace, alen, blen, tspace, seed_of):
    """LAS chains (START/NEXT/BEST flags; ids 0-based) -> seeded alignment dicts with DENTIST's flags and 1-based contig ids
    (dazzler.d:1731-1755).  seed_of(chain_first_record, chain_last_record) -> 'front' | 'back'."""
    COMP, START, NEXT, BEST = 0x1, 0x4, 0x8, 0x10
    out = []
    i, n, cid = 0, len(rec), 0
    while i < n:
        j = i + 1
        while j < n and (int(rec[j]["flags"]) & NEXT):
            j += 1
        f = int(rec[i]["flags"])
        flags = (1 if f & COMP else 0) | (4 if (f & START) and not (f & BEST) else 0)
        las = []
        for r in range(i, j):
            t = np.asarray(trace[int(toff[r]):int(toff[r]) + int(rec[r]["tlen"])], np.uint16).reshape(-1, 2)
            las.append(dict(ab=int(rec[r]["abpos"]), ae=int(rec[r]["aepos"]), bb=int(rec[r]["bbpos"]), be=int(rec[r]["bepos"]),
                            diffs=int(rec[r]["diffs"]), trace=t))
        a, b = int(rec[i]["aread"]), int(rec[i]["bread"])
        out.append(dict(id=cid, contigA=(a + 1, int(alen[a])), contigB=(b + 1, int(blen[b])), flags=flags, tpd=int(tspace),
                        seed=seed_of(rec[i], rec[j - 1]), las=las))
        cid += 1
        i = j
    return out
