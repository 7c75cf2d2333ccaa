"""DUST low-complexity mask -- ORACLE (test infrastructure only) for dbdust() (dazzler.d:3815-3818 -> `DBdust`).
Parity unpinned: DAZZ_DB @ d22ae58d (DBdust.c) is absent; the reference only fixes the option names
(window -w, threshold -t, minimum -m: dazzler.d:3796-3806) and the track layout (dazzler.d:4943-5052).
Spec: windows of `w` bases at stride w/2; a window with l = len-2 >= 14 triplets is low-complexity when
10 * sum_t c_t*(c_t-1)/2 > round(10*threshold) * (l-1); union of such windows, merged, >= minlen."""
import numpy as np


def dust_read(seq, w=64, threshold=2.0, minlen=10):
    L = len(seq)
    stride = w // 2
    t10 = int(threshold * 10.0 + 0.5)
    flags = []
    for start in range(0, L, stride):
        n = min(w, L - start)
        l = n - 2
        f = False
        if l >= 14:
            s = seq[start:start + n].astype(np.int64)
            trip = s[:-2] * 16 + s[1:-1] * 4 + s[2:]
            c = np.bincount(trip, minlength=64)
            f = 10 * int((c * (c - 1) // 2).sum()) > t10 * (l - 1)
        flags.append(f)
    out = []
    q = 0
    while q < len(flags):
        if not flags[q]:
            q += 1
            continue
        e = q
        while e + 1 < len(flags) and flags[e + 1]:
            e += 1
        b0, e0 = q * stride, min(e * stride + w, L)
        if e0 - b0 >= minlen:
            out.append((b0, e0))
        q = e + 1
    return out


def dust_block(off, bases, **kw):
    return [dust_read(bases[off[r]:off[r + 1]], **kw) for r in range(len(off) - 1)]
