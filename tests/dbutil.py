"""Test helper: writes DAZZ_DB databases / mask tracks from Python (independent of csrc/dazzdb.cpp's writer).
Layout per thegenemyers/DAZZ_DB DB.h (stub text + .idx = 112-byte DAZZ_DB header + 40-byte DAZZ_READs + .bps);
track layout per dazzler.d:4943-5052."""
import os
import struct

import numpy as np

from dentist_b200 import synth

DB_BEST = 0x800


def write_db(path, block, cutoff=0, all_=1, flags=None, block_first=None):
    """block_first: first read of every DB block after the first (DBsplit); default one block."""
    d, base = os.path.split(path)
    dam = base.endswith(".dam")
    root = base[:-4] if dam else base[:-3]
    n = block.nreads
    bps, boffs, o = [], [], 0
    for r in range(n):
        p = synth.pack_2bit_dazz(block.read(r)); boffs.append(o); bps.append(p); o += len(p)
    lens = np.diff(block.off)
    hdr = struct.pack("<4i4f", n, n, cutoff, all_, .25, .25, .25, .25)
    hdr += struct.pack("<i4xq", int(lens.max()) if n else 0, int(lens.sum()))
    hdr += struct.pack("<5i4x", n, 1, 0, 0, 0)
    hdr += struct.pack("<Qi4xQQQ", 0, 0, 0, 0, 0)
    assert len(hdr) == 112
    with open(os.path.join(d, "." + root + ".idx"), "wb") as f:
        f.write(hdr)
        for r in range(n):
            fl = (flags[r] if flags is not None else DB_BEST) | 850
            f.write(struct.pack("<3i4xqqi4x", r, int(lens[r]), 0, boffs[r], 0 if dam else -1, fl))
    with open(os.path.join(d, "." + root + ".bps"), "wb") as f:
        f.write(np.concatenate(bps).tobytes() if bps else b"")
    if dam:
        open(os.path.join(d, "." + root + ".hdr"), "w").write(">%s\n" % root)
    firsts = [0] + list(block_first or []) + [n]
    with open(path, "w") as f:
        f.write("files = %9d\n  %9d %s %s\nblocks = %9d\nsize = %11d cutoff = %9d all = %1d\n"
                % (1, n, root, root, len(firsts) - 1, 200000000, cutoff, all_))
        for x in firsts:
            f.write(" %9d %9d\n" % (x, x))


def write_track(dbpath, name, intervals):
    d, base = os.path.split(dbpath)
    root = base.rsplit(".", 1)[0]
    anno = struct.pack("<ii", len(intervals), 0)
    data = b""
    offs = [0]
    for iv in intervals:
        for b, e in iv:
            data += struct.pack("<ii", b, e)
        offs.append(len(data))
    anno += struct.pack("<%dq" % len(offs), *offs)
    open(os.path.join(d, ".%s.%s.anno" % (root, name)), "wb").write(anno)
    open(os.path.join(d, ".%s.%s.data" % (root, name)), "wb").write(data)
