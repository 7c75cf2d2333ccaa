"""Seeded synthetic assemblies and read sets (replaces DAZZ_DB `simulator`, which the reference's
tests use -- tests/test-commands.sh:7-13, example/Makefile:13 -- and which is absent here).

Everything is numpy-vectorised so the 200 Mbp read set of BASELINE.json configs[1]
(10 Mbp assembly, 100 gaps, 20x 10 kb PacBio-like reads) is generated in seconds.
Bases are uint8 codes a=0 c=1 g=2 t=3 (DAZZ_DB convention).
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class Block:
    """A set of sequences: offsets[nreads+1] into the concatenated base array."""
    off: np.ndarray          # int64
    bases: np.ndarray        # uint8 0..3

    @property
    def nreads(self):
        return len(self.off) - 1

    def read(self, i):
        return self.bases[self.off[i]:self.off[i + 1]]

    @property
    def total(self):
        return int(self.off[-1])


def make_scaffolds(n, length, seed, repeat_len=2000, repeat_copies=20, n_repeats=5):
    rng = np.random.default_rng(seed)
    scaffolds = [rng.integers(0, 4, size=length, dtype=np.uint8) for _ in range(n)]
    # plant repeats (exercise masks / frequency cap)
    for _ in range(n_repeats):
        unit = rng.integers(0, 4, size=repeat_len, dtype=np.uint8)
        for _ in range(repeat_copies):
            s = scaffolds[int(rng.integers(0, n))]
            if len(s) <= repeat_len:
                continue
            p = int(rng.integers(0, len(s) - repeat_len))
            s[p:p + repeat_len] = unit
    return scaffolds


def make_gaps(scaffolds, gaps_per_scaffold, seed, min_len=100, max_len=5000, min_dist=20000):
    """Returns per scaffold a sorted list of (begin, end) gap intervals."""
    rng = np.random.default_rng(seed)
    out = []
    for s in scaffolds:
        n = len(s)
        gaps = []
        if gaps_per_scaffold > 0:
            slot = n // (gaps_per_scaffold + 1)
            for g in range(gaps_per_scaffold):
                glen = int(np.exp(rng.uniform(np.log(min_len), np.log(max_len))))
                centre = (g + 1) * slot
                jitter = max(1, (slot - min_dist - glen) // 2)
                b = centre + int(rng.integers(-jitter, jitter)) if jitter > 1 else centre
                b = max(min_dist // 2, min(n - glen - min_dist // 2, b))
                gaps.append((b, b + glen))
        out.append(gaps)
    return out


def contigs_from(scaffolds, gaps):
    """Split scaffolds at gaps -> Block of contigs + (scaffold, begin) per contig."""
    seqs, meta = [], []
    for si, (s, gl) in enumerate(zip(scaffolds, gaps)):
        p = 0
        for (b, e) in gl:
            seqs.append(s[p:b]); meta.append((si, p)); p = e
        seqs.append(s[p:]); meta.append((si, p))
    off = np.zeros(len(seqs) + 1, np.int64)
    off[1:] = np.cumsum([len(x) for x in seqs])
    return Block(off, np.concatenate(seqs) if seqs else np.zeros(0, np.uint8)), meta


def simulate_reads(scaffolds, coverage, mean_len, sd_len, err, seed, mix=(0.73, 0.20, 0.07), min_len=1000,
                   lognormal=True):
    """Sample reads uniformly from the scaffolds, both strands, with ins:del:sub errors at total rate `err`.
    Returns (Block, truth) with truth = int64 array [nreads, 4] = (scaffold, begin, end, strand)."""
    rng = np.random.default_rng(seed)
    total = sum(len(s) for s in scaffolds)
    target = int(total * coverage)
    n_est = max(1, int(target / mean_len * 1.3) + 8)
    if lognormal:
        sigma2 = np.log(1 + (sd_len / mean_len) ** 2)
        lens = rng.lognormal(np.log(mean_len) - sigma2 / 2, np.sqrt(sigma2), size=n_est)
    else:
        lens = rng.normal(mean_len, sd_len, size=n_est)
    lens = np.maximum(lens, min_len).astype(np.int64)
    csum = np.cumsum(lens)
    n = int(np.searchsorted(csum, target) + 1)
    lens = lens[:n]
    slen = np.array([len(s) for s in scaffolds], np.int64)
    sc = rng.choice(len(scaffolds), size=n, p=slen / slen.sum())
    lens = np.minimum(lens, slen[sc])
    beg = (rng.random(n) * (slen[sc] - lens + 1)).astype(np.int64)
    strand = rng.integers(0, 2, size=n)
    # gather true sequences
    soff = np.zeros(n + 1, np.int64); soff[1:] = np.cumsum(lens)
    src = np.empty(int(soff[-1]), np.uint8)
    for i in range(n):
        seg = scaffolds[sc[i]][beg[i]:beg[i] + lens[i]]
        if strand[i]:
            seg = (3 - seg)[::-1]
        src[soff[i]:soff[i + 1]] = seg
    # errors
    pi, pd, ps = (err * m for m in mix)
    u = rng.random(len(src), dtype=np.float32)
    dele = u < pd
    sub = (u >= pd) & (u < pd + ps)
    ins = rng.random(len(src), dtype=np.float32) < pi
    src = np.where(sub, (src + rng.integers(1, 4, size=len(src), dtype=np.uint8)) & 3, src).astype(np.uint8)
    cnt = ins.astype(np.int64) + (~dele).astype(np.int64)
    cend = np.cumsum(cnt)
    out = np.empty(int(cend[-1]) if len(cend) else 0, np.uint8)
    keep = ~dele
    out[cend[keep] - 1] = src[keep]
    out[(cend - cnt)[ins]] = rng.integers(0, 4, size=int(ins.sum()), dtype=np.uint8)
    off = np.zeros(n + 1, np.int64)
    off[1:] = cend[soff[1:] - 1]
    truth = np.stack([sc, beg, beg + lens, strand], axis=1).astype(np.int64)
    return Block(off, out), truth


def pack_2bit_dazz(bases):
    """DAZZ_DB .bps packing: 4 bases per byte, first base in the two most significant bits."""
    n = len(bases)
    pad = (-n) % 4
    b = np.concatenate([bases, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    return ((b[:, 0] << 6) | (b[:, 1] << 4) | (b[:, 2] << 2) | b[:, 3]).astype(np.uint8)


def make_pile_batch(scaffolds, gaps, seed, depth=18, anchor=1500, err=0.13, mix=(0.73, 0.20, 0.07)):
    """Cropped pile-ups as `processPileUps` sees them after `crop` (package.d:409, cropper.d:113): for
    every gap, `depth` noisy full-length copies (random strand) of the region gap +- anchor.
    Returns (Block of all cropped reads, pile id per read, list of (scaffold, begin, end) regions)."""
    rng = np.random.default_rng(seed)
    seqs, group, regions = [], [], []
    pid = 0
    for si, gl in enumerate(gaps):
        for (b, e) in gl:
            rb, re_ = max(0, b - anchor), min(len(scaffolds[si]), e + anchor)
            region = scaffolds[si][rb:re_]
            n = int(rng.integers(max(4, depth - 6), depth + 7))
            blk, _ = simulate_reads([region], n, len(region), 1, err, int(rng.integers(1 << 30)), mix=mix,
                                    min_len=len(region), lognormal=False)
            for r in range(blk.nreads):
                seqs.append(blk.read(r)); group.append(pid)
            regions.append((si, rb, re_)); pid += 1
    off = np.zeros(len(seqs) + 1, np.int64); off[1:] = np.cumsum([len(s) for s in seqs])
    return Block(off, np.concatenate(seqs)), np.array(group, np.int32), regions
