"""Python mirror of the slice of `source/dentist/dazzler.d` that sits on the hot path, bound to the
C ABI.  Names and argument meaning follow the reference so the parity tests read like its own:

  getDalignment(dbA[, dbB], opts, outdir) -> las path     dazzler.d:3829-3844
  getDamapping(refDb, queryDb, opts, outdir) -> las path   dazzler.d:3855-3866
  getLasFile(dbA, dbB, outdir)                             dazzler.d:4345-4354

plus the in-memory hand-off: `Block` = a DAZZ_DB read block resident in HBM,
`align_blocks(A, B)` = the records of A.B.las without touching the file system.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import DnError  # noqa: F401  (re-export)


def init(device=0, tmpdir=None):
    _lib.check(_lib.lib().dn_init(int(device), tmpdir.encode() if tmpdir else None))


def launch_count():
    return int(_lib.lib().dn_launch_count())


def _block_desc(off, bases=None, bps=None, boff=None, mask=None, group=None):
    """dn_block_desc over caller-owned numpy arrays; returns (desc, arrays to keep alive, payload bytes)."""
    off = np.ascontiguousarray(off, dtype=np.int64)
    rlen = np.ascontiguousarray(np.diff(off), dtype=np.int32)
    d = _lib.BlockDesc()
    d.nreads = len(rlen)
    keep = [rlen]
    if bps is not None:
        data = np.ascontiguousarray(bps, dtype=np.uint8)
        bo = np.ascontiguousarray(boff, dtype=np.int64)
        d.format = 1
    else:
        data = np.ascontiguousarray(bases, dtype=np.uint8)
        bo = np.ascontiguousarray(off[:-1], dtype=np.int64)
        d.format = 0
    keep += [data, bo]
    d.rlen = rlen.ctypes.data
    d.boff = bo.ctypes.data
    d.data = data.ctypes.data
    d.data_bytes = data.nbytes
    if mask is not None:   # list of per-read interval lists [(b, e), ...]
        anno = np.zeros(len(rlen) + 1, np.int64)
        flat = []
        for r, iv in enumerate(mask):
            anno[r] = 4 * len(flat)
            for b, e in iv:
                flat += [b, e]
        anno[len(rlen)] = 4 * len(flat)
        md = np.ascontiguousarray(flat if flat else [0, 0], dtype=np.int32)
        keep += [anno, md]
        d.mask_anno = anno.ctypes.data
        d.mask_data = md.ctypes.data
    if group is not None:  # pile id per read: only reads of the same pile are compared
        grp = np.ascontiguousarray(group, dtype=np.int32)
        assert len(grp) == len(rlen)
        keep.append(grp)
        d.group = grp.ctypes.data
    return d, keep, int(data.nbytes + rlen.nbytes + bo.nbytes)


class HostBlock:
    """A sequence block still on the host (pinned or pageable numpy arrays): the input of `align_host`."""

    def __init__(self, off, bases=None, bps=None, boff=None, mask=None, group=None):
        self._desc, self._keep, self.h2d_bytes = _block_desc(off, bases, bps, boff, mask, group)
        self.nreads = int(self._desc.nreads)


class Block:
    """A sequence block resident in HBM (2-bit packed, both strands)."""

    def __init__(self, off, bases=None, bps=None, boff=None, mask=None, group=None):
        """Either `bases` (uint8 codes 0..3 concatenated, read r = bases[off[r]:off[r+1]]) or
        DAZZ_DB `.bps` bytes with per-read byte offsets `boff` and lengths diff(off)."""
        d, keep, nbytes = _block_desc(off, bases, bps, boff, mask, group)
        self._desc, self._keep = d, keep
        self._h = C.c_void_p()
        _lib.check(_lib.lib().dn_block_upload(C.byref(d), C.byref(self._h)))
        self.nreads = int(d.nreads)
        self.bases = int(_lib.lib().dn_block_bases(self._h))
        self.h2d_bytes = nbytes
        self.has_group = group is not None

    @classmethod
    def crop(cls, src, read, begin, end, group=None):
        """Cropped reads of `src` as a new resident block (cropper.d:383-421 without FASTA / DB files)."""
        read = np.ascontiguousarray(read, np.int32); begin = np.ascontiguousarray(begin, np.int32); end = np.ascontiguousarray(end, np.int32)
        grp = None if group is None else np.ascontiguousarray(group, np.int32)
        self = cls.__new__(cls)
        self._keep = []; self._desc = None
        self._h = C.c_void_p()
        _lib.check(_lib.lib().dn_block_crop(src._h, len(read), read.ctypes.data_as(C.c_void_p), begin.ctypes.data_as(C.c_void_p),
                                            end.ctypes.data_as(C.c_void_p), None if grp is None else grp.ctypes.data_as(C.c_void_p),
                                            C.byref(self._h)))
        self.nreads = len(read)
        self.bases = int(_lib.lib().dn_block_bases(self._h))
        self.h2d_bytes = int(read.nbytes * 3)
        self.has_group = group is not None
        return self

    def index(self, k):
        """Build and keep the k-mer index of this block (dn_block_index): later alignments with this block as A and
        the same k skip the index build.  k <= 0 drops it."""
        _lib.check(_lib.lib().dn_block_index(self._h, int(k)))

    def maskDust(self, window=64, threshold=2.0, minlen=10):
        """dbdust(db) followed by `-mdust` (package.d:476-481): DUST intervals join the block's own seed mask on the
        device.  Returns the number of masked bases."""
        m = C.c_int64(0)
        _lib.check(_lib.lib().dn_block_mask_dust(self._h, int(window), C.c_double(threshold), int(minlen), C.byref(m)))
        return int(m.value)

    def free(self):
        if self._h:
            _lib.lib().dn_block_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def make_params(**kw):
    p = _lib.AlignParams()
    _lib.lib().dn_align_params_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError("unknown alignment parameter %r" % k)
        setattr(p, k, v)
    return p


class _LasOwner:
    """Keeps a dn_las_buf alive while numpy views of its arrays exist (zero-copy hand-off)."""

    def __init__(self, buf):
        self.buf = buf

    def __del__(self):
        try:
            _lib.lib().dn_las_free(C.byref(self.buf))
        except Exception:
            pass


class _View(np.ndarray):
    _owner = None


def _wrap(ptr, nbytes, dtype, owner):
    if not nbytes:
        return np.zeros(0, dtype)
    raw = (C.c_uint8 * nbytes).from_address(C.addressof(ptr.contents))
    arr = np.frombuffer(raw, dtype=dtype).view(_View)
    arr._owner = owner
    return arr


def _take(buf):
    owner = _LasOwner(buf)
    n = int(buf.nrec)
    rec = _wrap(buf.rec, n * 40, _lib.REC_DTYPE, owner)
    toff = _wrap(buf.toff, n * 8, np.int64, owner)
    tr = _wrap(buf.trace, int(buf.ntrace) * 2, np.uint16, owner)
    st = {n_: getattr(buf.stats, n_) for n_, _ in _lib.AlignStats._fields_}
    return rec, toff, tr, int(buf.tspace), st


def align_host(a, b, **params):
    """dn_align_host: host blocks in, host LAS out -- upload, alignment and download in ONE call; the upload of `b`
    overlaps the upload and indexing of `a`.  Returns (records, trace offsets, trace, stats) like align_blocks."""
    p = make_params(**params)
    buf = _lib.LasBuf()
    _lib.check(_lib.lib().dn_align_host(C.byref(a._desc), C.byref(b._desc), C.byref(p), C.byref(buf)))
    rec, toff, tr, _, st = _take(buf)
    return rec, toff, tr, st


def align_blocks(a, b, **params):
    """All local alignments of resident block `a` vs `b`: (records, trace offsets, trace, stats)."""
    p = make_params(**params)
    buf = _lib.LasBuf()
    _lib.check(_lib.lib().dn_align_blocks(a._h, b._h, C.byref(p), C.byref(buf)))
    rec, toff, tr, _, st = _take(buf)
    return rec, toff, tr, st


def read_las(path):
    buf = _lib.LasBuf()
    _lib.check(_lib.lib().dn_las_read(path.encode(), C.byref(buf)))
    rec, toff, tr, tspace, _ = _take(buf)
    return tspace, rec, toff, tr


def write_las(path, tspace, rec, toff, trace):
    """Write caller-owned arrays as a LAS file (writer side of dazzler.d:1913-2170): rec = REC_DTYPE records,
    toff = trace offset (uint16 units) per record, trace = uint16 (diffs, bases) pairs."""
    rec = np.ascontiguousarray(rec, _lib.REC_DTYPE); toff = np.ascontiguousarray(toff, np.int64); trace = np.ascontiguousarray(trace, np.uint16)
    buf = _lib.LasBuf()
    buf.nrec = len(rec); buf.ntrace = len(trace); buf.tspace = int(tspace)
    buf.rec = C.cast(rec.ctypes.data, C.POINTER(_lib.LasRecord))
    buf.toff = C.cast(toff.ctypes.data, C.POINTER(C.c_int64))
    buf.trace = C.cast(trace.ctypes.data, C.POINTER(C.c_uint16))
    _lib.check(_lib.lib().dn_las_write(path.encode(), C.byref(buf)))


def _opts(opts):
    arr = (C.c_char_p * len(opts))(*[o.encode() for o in opts])
    return arr, len(opts)


def _db_name(db):
    b = os.path.basename(db)
    for ext in (".db", ".dam"):
        if b.endswith(ext):
            b = b[:-len(ext)]
    return b


def getLasFile(dbA, dbB, outdir):
    """dazzler.d:4345-4354"""
    return os.path.join(outdir, "%s.%s.las" % (_db_name(dbA), _db_name(dbB if dbB is not None else dbA)))


def getDalignment(dbA, dbB=None, opts=(), outdir="."):
    """dazzler.d:3829-3844 -- `daligner opts dbA dbB` run in outdir; returns the LAS path."""
    arr, n = _opts(list(opts))
    _lib.check(_lib.lib().dn_dalign(dbA.encode(), dbB.encode() if dbB else None, arr, n, outdir.encode()))
    return getLasFile(dbA, dbB, outdir)


def getDamapping(refDb, queryDb, opts=(), outdir="."):
    """dazzler.d:3855-3866 -- `damapper -C opts refDb queryDb`; returns outdir/<ref>.<query>.las."""
    arr, n = _opts(list(opts))
    _lib.check(_lib.lib().dn_damap(refDb.encode(), queryDb.encode(), arr, n, outdir.encode()))
    return getLasFile(refDb, queryDb, outdir)


# ---------------------------------------------------------------------------------------------
# per-pile stages (processPileUps/package.d:474-619) on in-memory LAS buffers


class Las:
    """An in-memory LAS (owner of a dn_las_buf).  `rec`, `toff`, `trace` are zero-copy views."""

    def __init__(self, buf):
        self._buf = buf
        self._owner = _LasOwner(buf)
        self.stats = {n_: getattr(buf.stats, n_) for n_, _ in _lib.AlignStats._fields_}
        self._refresh()

    def _refresh(self):
        b = self._buf
        n = int(b.nrec)
        self.rec = _wrap(b.rec, n * 40, _lib.REC_DTYPE, self._owner)
        self.toff = _wrap(b.toff, n * 8, np.int64, self._owner)
        self.trace = _wrap(b.trace, int(b.ntrace) * 2, np.uint16, self._owner)
        self.tspace = int(b.tspace)

    def __len__(self):
        return int(self._buf.nrec)

    def traces(self):
        return [self.trace[o:o + t].reshape(-1, 2) for o, t in zip(self.toff, self.rec["tlen"])]

    def filterLocalAlignments(self, max_err):
        """dazzler.d:3885-3899 with pred = averageErrorRate <= max_err (package.d:483-485)."""
        _lib.check(_lib.lib().dn_las_filter_error(C.byref(self._buf), C.c_double(max_err)))
        self._refresh()
        return self

    def filterPileUpAlignments(self, alen, blen, allowance):
        """dazzler.d:4043-4094 (forceFlat; records are already flat and sorted)."""
        alen = np.ascontiguousarray(alen, np.int32); blen = np.ascontiguousarray(blen, np.int32)
        _lib.check(_lib.lib().dn_las_filter_pileup(C.byref(self._buf), alen.ctypes.data_as(C.c_void_p), len(alen),
                                                    blen.ctypes.data_as(C.c_void_p), len(blen), int(allowance)))
        self._refresh()
        return self

    def chainLocalAlignments(self, max_indel=1000, max_chain_gap=10000, max_rel_overlap=0.3, min_rel_score=1.0, min_score=None):
        """dazzler.d:3995-4018 / chaining.d:122-334 with ChainingOptions (commandline.d:2820-2830)."""
        ms = self.tspace if min_score is None else min_score
        _lib.check(_lib.lib().dn_las_chain(C.byref(self._buf), int(max_indel), int(max_chain_gap), C.c_double(max_rel_overlap),
                                           C.c_double(min_rel_score), int(ms)))
        self._refresh()
        return self

    def forceFlat(self):
        """filterPileUpAlignments(..., Yes.forceFlat) (dazzler.d:4084-4093): drop the chain flags and sort the
        records in FlatLocalAlignment order (base.d:1787-1809)."""
        _lib.check(_lib.lib().dn_las_force_flat(C.byref(self._buf)))
        self._refresh()
        return self

    def chainMapper(self, nb_reads, max_indel=1000, max_gap=10000):
        """damapper-style START/NEXT/BEST flags (decoded at dazzler.d:1738-1755)."""
        _lib.check(_lib.lib().dn_las_chain_mapper(C.byref(self._buf), int(nb_reads), int(max_indel), int(max_gap)))
        self._refresh()
        return self

    def keepBestChains(self, nb_reads, n_frac=0.0):
        """What damapper reports (dazzler.d:5920-5923): the best chain per read, plus chains within `n_frac` of it (-n)."""
        _lib.check(_lib.lib().dn_las_keep_best_chains(C.byref(self._buf), int(nb_reads), C.c_double(n_frac)))
        self._refresh()
        return self

    def bridge(self, a, b, e=0.7):
        """`daligner -B` (dazzler.d:5823-5824): neighbouring records separated by a short gap become one.  Returns the
        number of bridges made."""
        nb = C.c_int64(0)
        _lib.check(_lib.lib().dn_las_bridge(a._h, b._h, C.byref(self._buf), int(round(6.0 / (1.0 - e))), C.byref(nb)))
        self._refresh()
        return int(nb.value)

    def transpose(self, a, b):
        """`damapper -C`'s second file (dazzler.d:5931-5936): the records of B.A.las for these records of A.B.las."""
        buf = _lib.LasBuf()
        _lib.check(_lib.lib().dn_las_transpose(a._h, b._h, C.byref(self._buf), C.byref(buf)))
        return Las(buf)

    def write(self, path):
        _lib.check(_lib.lib().dn_las_write(path.encode(), C.byref(self._buf)))


def align(a, b, **params):
    p = make_params(**params)
    buf = _lib.LasBuf()
    _lib.check(_lib.lib().dn_align_blocks(a._h, b._h, C.byref(p), C.byref(buf)))
    return Las(buf)


def computeQVs(rlen, las, coverage):
    """dazzler.d:3782-3792: returns (qv bytes, qoff[nreads+1])."""
    L = _lib.lib()
    rlen = np.ascontiguousarray(rlen, np.int32)
    qv = C.POINTER(C.c_uint8)(); qoff = C.POINTER(C.c_int64)()
    if np.ndim(coverage) == 0:
        _lib.check(L.dn_compute_qvs(rlen.ctypes.data_as(C.c_void_p), len(rlen), C.byref(las._buf), int(coverage),
                                    C.byref(qv), C.byref(qoff)))
    else:   # one coverage per read (batched pile-ups)
        cv = np.ascontiguousarray(coverage, np.int32)
        assert len(cv) == len(rlen)
        _lib.check(L.dn_compute_qvs_v(rlen.ctypes.data_as(C.c_void_p), len(rlen), C.byref(las._buf), 0,
                                      cv.ctypes.data_as(C.c_void_p), C.byref(qv), C.byref(qoff)))
    o = np.ctypeslib.as_array(qoff, shape=(len(rlen) + 1,)).copy()
    q = np.ctypeslib.as_array(qv, shape=(max(int(o[-1]), 1),))[:int(o[-1])].copy()
    L.dn_free(qv); L.dn_free(qoff)
    return q, o


class _SeqBuf(C.Structure):
    _fields_ = [("nseq", C.c_int32), ("off", C.POINTER(C.c_int64)), ("bases", C.POINTER(C.c_uint8))]


def getConsensus(block, las, reads):
    """dazzler.d:4213-4255 for each 0-based read id in `reads`; returns a list of base-code arrays."""
    L = _lib.lib()
    reads = np.ascontiguousarray(reads, np.int32)
    out = _SeqBuf()
    _lib.check(L.dn_consensus(block._h, C.byref(las._buf), reads.ctypes.data_as(C.c_void_p), len(reads), C.byref(out)))
    off = np.ctypeslib.as_array(out.off, shape=(len(reads) + 1,)).copy()
    tot = int(off[-1])
    bases = np.ctypeslib.as_array(out.bases, shape=(max(tot, 1),))[:tot].copy()
    L.dn_seq_free(C.byref(out))
    return [bases[off[i]:off[i + 1]] for i in range(len(reads))]


def dbdust(block, window=64, threshold=2.0, minlen=10):
    """dazzler.d:3815-3818 on a resident block: per-read list of masked (begin, end) intervals."""
    L = _lib.lib()
    anno = C.POINTER(C.c_int64)(); data = C.POINTER(C.c_int32)()
    _lib.check(L.dn_dust_block(block._h, int(window), C.c_double(threshold), int(minlen), C.byref(anno), C.byref(data)))
    a = np.ctypeslib.as_array(anno, shape=(block.nreads + 1,)).copy()
    n = int(a[-1]) // 4
    d = np.ctypeslib.as_array(data, shape=(max(n, 1),))[:n].copy()
    L.dn_free(anno); L.dn_free(data)
    return [[(int(d[i]), int(d[i + 1])) for i in range(int(a[r]) // 4, int(a[r + 1]) // 4, 2)] for r in range(block.nreads)]


def dbdustFile(dbFile, opts=()):
    """dazzler.d:3815-3818 on a DB file: writes the `dust` track next to it."""
    arr, n = _opts(list(opts))
    _lib.check(_lib.lib().dn_dbdust(dbFile.encode(), arr, n))


def computeQVsDb(dbFile, lasFile, coverage=0):
    """dazzler.d:3782-3792 on files (`DAScover`, `DASqv -c`): writes the `qual` track of dbFile."""
    _lib.check(_lib.lib().dn_compute_qvs_db(dbFile.encode(), lasFile.encode(), int(coverage)))


def getIntrinsicQVs(dbFile):
    """What getDbRecords(db, [readNumber, intrinsicQualityVector]) takes from `DBdump -r -i` (package.d:520-523,
    dazzler.d:2877-2898): the `qual` track of dbFile as one QV array per read."""
    L = _lib.lib()
    qv = C.POINTER(C.c_uint8)(); qoff = C.POINTER(C.c_int64)(); n = C.c_int32(0)
    _lib.check(L.dn_read_qvs_db(dbFile.encode(), C.byref(qv), C.byref(qoff), C.byref(n)))
    o = np.ctypeslib.as_array(qoff, shape=(n.value + 1,)).copy()
    q = np.ctypeslib.as_array(qv, shape=(max(int(o[-1]), 1),))[:int(o[-1])].copy()
    L.dn_free(qv); L.dn_free(qoff)
    return [q[int(o[r]):int(o[r + 1])] for r in range(n.value)]


def getConsensusDb(dbFile, filteredLasFile, readId, opts=()):
    """dazzler.d:4213-4238 (file form): returns the path of the consensus .dam; raises "empty consensus"."""
    arr, n = _opts(list(opts))
    out = C.create_string_buffer(4096)
    _lib.check(_lib.lib().dn_consensus_db(dbFile.encode(), filteredLasFile.encode(), int(readId), arr, n, out, 4096))
    return out.value.decode()


def collectFilter(las, alen, blen, repeat_mask=None, max_alignment_error=0.3, proper_alignment_allowance=100, min_anchor_length=500):
    """collectPileUps' filterAlignments (collectPileUps/package.d:129-160, filter.d:122-356) on a chained LAS.
    repeat_mask: per A contig a sorted list of disjoint (begin, end).  Returns (first record of each chain,
    status per chain, sorted list of B reads the filters consumed)."""
    L = _lib.lib()
    alen = np.ascontiguousarray(alen, np.int32); blen = np.ascontiguousarray(blen, np.int32)
    anno = data = None
    if repeat_mask is not None:
        anno = np.zeros(len(alen) + 1, np.int64); flat = []
        for r in range(len(alen)):
            anno[r] = 4 * len(flat)
            for b, e in repeat_mask[r] if r < len(repeat_mask) else []:
                flat += [b, e]
        anno[len(alen)] = 4 * len(flat)
        data = np.ascontiguousarray(flat if flat else [0, 0], np.int32)
    n = C.c_int64(0); first = C.POINTER(C.c_int32)(); st = C.POINTER(C.c_uint8)(); used = C.POINTER(C.c_uint8)()
    _lib.check(L.dn_collect_filter(C.byref(las._buf), alen.ctypes.data_as(C.c_void_p), len(alen), blen.ctypes.data_as(C.c_void_p), len(blen),
                                   None if anno is None else anno.ctypes.data_as(C.c_void_p), None if data is None else data.ctypes.data_as(C.c_void_p),
                                   C.c_double(max_alignment_error), int(proper_alignment_allowance), int(min_anchor_length),
                                   C.byref(n), C.byref(first), C.byref(st), C.byref(used)))
    nc = int(n.value)
    f = np.ctypeslib.as_array(first, shape=(max(nc, 1),))[:nc].copy()
    s_ = np.ctypeslib.as_array(st, shape=(max(nc, 1),))[:nc].copy()
    u = np.ctypeslib.as_array(used, shape=(max(len(blen), 1),))[:len(blen)].copy()
    L.dn_free(first); L.dn_free(st); L.dn_free(used)
    return f, s_, np.flatnonzero(u).tolist()


def _track(anno_p, data_p, n):
    L = _lib.lib()
    a = np.ctypeslib.as_array(anno_p, shape=(n + 1,)).copy()
    m = int(a[-1]) // 4
    d = np.ctypeslib.as_array(data_p, shape=(max(m, 1),))[:m].copy()
    L.dn_free(anno_p); L.dn_free(data_p)
    return [[(int(d[i]), int(d[i + 1])) for i in range(int(a[r]) // 4, int(a[r + 1]) // 4, 2)] for r in range(n)]


def maskRepetitiveRegions(las, alen, blen, coverage_bounds, improper_coverage_bounds=None, proper_alignment_allowance=100):
    """`dentist mask-repetitive-regions` (commands/maskRepetitiveRegions.d:135-232) on an in-memory LAS: per A contig the
    intervals whose chain coverage leaves coverage_bounds, united (reads alignments only) with those whose coverage by
    IMPROPER chains leaves improper_coverage_bounds."""
    L = _lib.lib()
    alen = np.ascontiguousarray(alen, np.int32); blen = np.ascontiguousarray(blen, np.int32)
    passes = [(coverage_bounds, 0)] + ([(improper_coverage_bounds, 1)] if improper_coverage_bounds is not None else [])
    masks = []
    for (lo, hi), improper in passes:
        anno = C.POINTER(C.c_int64)(); data = C.POINTER(C.c_int32)()
        _lib.check(L.dn_mask_coverage(C.byref(las._buf), alen.ctypes.data_as(C.c_void_p), len(alen), blen.ctypes.data_as(C.c_void_p), len(blen),
                                      C.c_double(lo), C.c_double(hi), improper, int(proper_alignment_allowance), C.byref(anno), C.byref(data)))
        masks.append(_track(anno, data, len(alen)))
    out = []
    for c in range(len(alen)):                                    # repetitiveRegions | repetitiveRegionsImproper (:209)
        merged = []
        for b, e in sorted(iv for m in masks for iv in m[c]):
            if merged and b <= merged[-1][1]:
                merged[-1] = (merged[-1][0], max(merged[-1][1], e))
            else:
                merged.append((b, e))
        out.append(merged)
    return out


def _track_arrays(mask, n):
    """per-sequence interval lists -> (anno, data) of the mask-track layout (dazzler.d:4943-5052)"""
    anno = np.zeros(n + 1, np.int64); flat = []
    for r in range(n):
        anno[r] = 4 * len(flat)
        for b, e in (mask[r] if r < len(mask) else []):
            flat += [b, e]
    anno[n] = 4 * len(flat)
    return anno, np.ascontiguousarray(flat if flat else [0, 0], np.int32)


def propagateMask(las, mask, na, blen):
    """`dentist propagate-mask` (commands/propagateMask.d:109-300) on an in-memory LAS with trace points: mask = per A
    contig a sorted list of disjoint (begin, end); returns the propagated mask per B read."""
    L = _lib.lib()
    blen = np.ascontiguousarray(blen, np.int32)
    manno, mdata = _track_arrays(mask, na)
    anno = C.POINTER(C.c_int64)(); data = C.POINTER(C.c_int32)()
    _lib.check(L.dn_propagate_mask(C.byref(las._buf), int(na), manno.ctypes.data_as(C.c_void_p), mdata.ctypes.data_as(C.c_void_p),
                                   blen.ctypes.data_as(C.c_void_p), len(blen), C.byref(anno), C.byref(data)))
    return _track(anno, data, len(blen))


def findReferenceReadCandidates(qv, qoff, group, npiles, bad_fraction=0.08):
    """processPileUps/package.d:518-568 for a batch: list (per pile) of read ids ranked by (numBadQVs, meanQV, readId)."""
    qv = np.ascontiguousarray(qv, np.uint8); qoff = np.ascontiguousarray(qoff, np.int64); group = np.ascontiguousarray(group, np.int32)
    rank = np.zeros(len(group), np.int32); poff = np.zeros(npiles + 1, np.int64)
    _lib.check(_lib.lib().dn_reference_read_candidates(qv.ctypes.data_as(C.c_void_p), qoff.ctypes.data_as(C.c_void_p),
                                                       group.ctypes.data_as(C.c_void_p), len(group), int(npiles), C.c_double(bad_fraction),
                                                       rank.ctypes.data_as(C.c_void_p), poff.ctypes.data_as(C.c_void_p)))
    return [rank[poff[p]:poff[p + 1]] for p in range(npiles)]


def block_add_mask(block, mask):
    """`-m<track>` for a resident block (dazzler.d:5842-5848): per-read interval lists join the seed-exclusion mask."""
    anno, data = _track_arrays(mask, block.nreads)
    _lib.check(_lib.lib().dn_block_add_mask(block._h, anno.ctypes.data_as(C.c_void_p), data.ctypes.data_as(C.c_void_p)))


class _HostLas:
    """Copied-out LAS of one pile-up's consensus-vs-flanks alignment (rec / toff / trace like `Las`)."""

    def __init__(self, buf):
        n = int(buf.nrec)
        self.rec = np.frombuffer(C.string_at(buf.rec, n * 40), _lib.REC_DTYPE).copy() if n else np.zeros(0, _lib.REC_DTYPE)
        self.toff = np.ctypeslib.as_array(buf.toff, shape=(n,)).copy() if n else np.zeros(0, np.int64)
        nt = int(buf.ntrace)
        self.trace = np.ctypeslib.as_array(buf.trace, shape=(nt,)).copy() if nt else np.zeros(0, np.uint16)
        self.tspace = int(buf.tspace)

    def __len__(self):
        return len(self.rec)

    def traces(self):
        return [self.trace[o:o + t].reshape(-1, 2) for o, t in zip(self.toff, self.rec["tlen"])]


class PileupBatch:
    """The dn_pileup_desc array of a batch, built once over host buffers: `run()` is then exactly the C call a D host
    makes (no Python work between the host buffers and dn_process_pileups)."""

    def __init__(self, ref, piles):
        self.ref, self.n = ref, len(piles)
        self.descs = (_lib.PileupDesc * max(self.n, 1))()
        self._keep = []
        self.bases_bytes = 0
        for i, p in enumerate(piles):
            rl = np.ascontiguousarray([len(r) for r in p["reads"]], np.int32)
            bs = np.ascontiguousarray(np.concatenate([np.asarray(r, np.uint8) for r in p["reads"]]) if len(rl) else np.zeros(0, np.uint8))
            fl = np.ascontiguousarray(p.get("flanks", ()), np.int32)
            self._keep += [rl, bs, fl]
            self.bases_bytes += int(bs.nbytes)
            d = self.descs[i]
            d.nreads = len(rl); d.rlen = rl.ctypes.data; d.bases = bs.ctypes.data
            if p.get("allowed") is not None:
                al = np.ascontiguousarray(p["allowed"], np.uint8); self._keep.append(al); d.allowed = al.ctypes.data
            d.nflanks = len(fl); d.flank_read = fl.ctypes.data
            if p.get("mask") is not None:
                anno, data = _track_arrays(p["mask"], len(fl)); self._keep += [anno, data]
                d.mask_anno = anno.ctypes.data; d.mask_data = data.ctypes.data

    def run(self, **params):
        L = _lib.lib()
        P = _lib.PileupParams()
        L.dn_pileup_params_default(C.byref(P))
        for k, v in params.items():
            if not hasattr(P, k):
                raise TypeError("unknown pile-up parameter %r" % k)
            setattr(P, k, v)
        outs = (_lib.InsertionOut * max(self.n, 1))()
        _lib.check(L.dn_process_pileups(self.ref._h if self.ref is not None else None, self.descs, self.n, C.byref(P), outs))
        return _Insertions(outs, self.n)


class _Insertions:
    """Owner of the dn_insertion_out array of one dn_process_pileups call."""

    def __init__(self, outs, n):
        self.outs, self.n = outs, n

    def consensus_bases(self):
        return sum(int(self.outs[i].cons_len) for i in range(self.n))

    def to_list(self):
        L = _lib.lib()
        res = []
        for i in range(self.n):
            o = self.outs[i]
            cons = np.ctypeslib.as_array(o.consensus, shape=(int(o.cons_len),)).copy() if o.cons_len else np.zeros(0, np.uint8)
            res.append(dict(status=int(o.status), reason=L.dn_pile_status_string(o.status).decode(), reference_read=int(o.reference_read),
                            ntries=int(o.ntries), consensus=cons, flank_las=_HostLas(o.flank_las)))
        return res

    def __del__(self):
        try:
            _lib.lib().dn_insertion_free(self.outs, self.n)
        except Exception:
            pass


def processPileUps(ref, piles, **params):
    """dn_process_pileups: the device part of PileUpProcessor.processPileUp (package.d:303-341) for a batch.
    ref: resident Block holding the flanking contigs (or None); piles: list of dicts with
      reads   = list of base-code arrays (the cropped reads, pile-up order)
      allowed = optional bool per read (allowedReferenceReadIds)
      flanks  = list of 0-based read ids in `ref` (croppingPositions order)
      mask    = optional per-flank list of (begin, end) intervals (repeat mask)
    params: fields of dn_pileup_params.  Returns one dict per pile-up: status, reason, reference_read (index in the
    pile-up or -1), ntries, consensus (base codes), flank_las (_HostLas: aread = index into `flanks`)."""
    return PileupBatch(ref, piles).run(**params).to_list()


# ---------------------------------------------------------------------------------------------
# multi-GPU (one process per GPU; the collective lives in the library: csrc/comm.cu)


def comm_init(rank, world, exchange=None):
    """Joins the library's NCCL communicator.  `exchange(id_bytes or None) -> id_bytes` moves the 128-byte id from
    rank 0 to the other ranks; the default uses torch.distributed's already initialised process group (plumbing)."""
    L = _lib.lib()
    idb = (C.c_uint8 * 128)()
    if rank == 0:
        _lib.check(L.dn_comm_get_id(idb))
    if exchange is None:
        import torch
        import torch.distributed as dist
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.tensor(list(bytes(idb)), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=0)
        raw = bytes(t.cpu().tolist())
    else:
        raw = exchange(bytes(idb) if rank == 0 else None)
    idb = (C.c_uint8 * 128).from_buffer_copy(raw)
    _lib.check(L.dn_comm_init(int(rank), int(world), idb))


def comm_shutdown():
    _lib.check(_lib.lib().dn_comm_shutdown())


def align_blocks_gather(a, b, bread_offset, root=0, **params):
    """dn_align_blocks_gather: every rank aligns its own resident read block `b` against `a`; the merged LAS arrives on
    `root` (every rank if root < 0).  Returns (records, trace offsets, trace, this rank's stats)."""
    p = make_params(**params)
    buf = _lib.LasBuf()
    _lib.check(_lib.lib().dn_align_blocks_gather(a._h, b._h, C.byref(p), int(bread_offset), int(root), C.byref(buf)))
    rec, toff, tr, _, st = _take(buf)
    return rec, toff, tr, st


def align_host_gather(a, b, bread_offset, root=0, **params):
    """dn_align_host_gather: host blocks in, merged host LAS out on `root` (upload + align + gather + merge + download)."""
    p = make_params(**params)
    buf = _lib.LasBuf()
    _lib.check(_lib.lib().dn_align_host_gather(C.byref(a._desc), C.byref(b._desc), C.byref(p), int(bread_offset), int(root), C.byref(buf)))
    rec, toff, tr, _, st = _take(buf)
    return rec, toff, tr, st


def comm_allgatherv(data):
    """dn_comm_allgatherv: bytes-like in, list of every rank's bytes out."""
    L = _lib.lib()
    buf = np.frombuffer(bytes(data), np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, np.uint8)
    world = int(L.dn_comm_size())
    counts = np.zeros(world, np.int64); out = C.c_void_p()
    _lib.check(L.dn_comm_allgatherv(buf.ctypes.data_as(C.c_void_p), int(buf.nbytes), C.byref(out), counts.ctypes.data_as(C.c_void_p)))
    tot = int(counts.sum())
    allb = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint8)), shape=(max(tot, 1),))[:tot].copy()
    L.dn_free(out)
    cuts = np.concatenate([[0], np.cumsum(counts)])
    return [allb[cuts[r]:cuts[r + 1]].tobytes() for r in range(world)]
