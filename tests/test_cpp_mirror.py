"""The C++ mirror of dazzler.d (include/dentist_b200.hpp) compiles against the C ABI; without a GPU every
compute call raises DazzlerCommandException (no CPU fallback); on a GPU it gives what the Python path gives."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M = (1 << 64) - 1


def build(tmp_path):
    exe = str(tmp_path / "replay")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "replay.cpp"),
                           "-L", os.path.join(ROOT, "dentist_b200"), "-ldentist_b200", "-Wl,-rpath," + os.path.join(ROOT, "dentist_b200"),
                           "-o", exe])
    return exe


def test_cpp_mirror_builds_and_fails_loudly_without_gpu(tmp_path):
    import torch
    exe = build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([exe, "7"], capture_output=True, text=True)
    assert r.returncode == 3 and "DazzlerCommandException" in r.stdout and "no CUDA device" in r.stdout


def _truth(seed, L):
    s = seed
    out = []
    for _ in range(L):
        s = (s * 6364136223846793005 + 1442695040888963407) & M
        out.append((s >> 33) & 3)
    return np.array(out, np.uint8)


def _pile(seed):
    s = seed
    def lcg():
        nonlocal s
        s = (s * 6364136223846793005 + 1442695040888963407) & M
        return s >> 33
    L, N = 3000, 8
    truth = [lcg() & 3 for _ in range(L)]
    reads = []
    for _ in range(N):
        out = []
        for i in range(L):
            u = lcg() % 100
            if u < 3:
                continue
            if u < 6:
                out.append(lcg() & 3)
            out.append((truth[i] + 1 + lcg() % 3) & 3 if u < 8 else truth[i])
        reads.append(np.array(out, np.uint8))
    return reads


@pytest.mark.gpu
def test_cpp_mirror_equals_python_path(tmp_path):
    from dentist_b200 import dazzler
    exe = build(tmp_path)
    r = subprocess.run([exe, "7"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = dict(ln.split(" ", 1) for ln in r.stdout.strip().split("\n"))
    reads = _pile(7)
    off = np.zeros(len(reads) + 1, np.int64); off[1:] = np.cumsum([len(x) for x in reads])
    g = dazzler.Block(off, np.concatenate(reads))
    assert int(got["dust"]) == g.maskDust()
    lens = np.diff(off)
    las = dazzler.align(g, g, tspace=126, minlen=500, self_block=1)
    assert int(got["raw"]) == len(las) > 20
    las.filterLocalAlignments(0.3)
    las.chainLocalAlignments()
    assert int(got["chained"]) == len(las)
    q, _ = dazzler.computeQVs(lens, las, 4)
    las.filterPileUpAlignments(lens, lens, 126)
    assert int(got["filtered"]) == len(las)
    assert int(got["qvsum"]) == int(q.astype(np.int64).sum())
    cov = dazzler.maskRepetitiveRegions(las, lens, lens, (0, 5))
    assert int(got["overcovered"]) == sum(e - b for c in cov for b, e in c) > 0
    prop = dazzler.propagateMask(las, [[(500, 900)]] + [[] for _ in reads[1:]], len(reads), lens)
    assert int(got["propagated"]) == sum(e - b for c in prop for b, e in c) > 0
    cons = dazzler.getConsensus(g, las, [0])[0]
    h = 1469598103934665603
    for b in cons.tolist():
        h = ((h ^ b) * 1099511628211) & M
    assert got["consensus"] == "%d %d" % (len(cons), h)
    # the batch entry point from C++ == the Python path (same C call), byte for byte
    L = 3000
    truth = _truth(7, L)
    ref = dazzler.Block(np.array([0, 1200, 2400]), np.concatenate([truth[:1200], truth[-1200:]]))
    allowed = [i == 3 for i in range(len(reads))]
    py = dazzler.processPileUps(ref, [dict(reads=reads, flanks=[0, 1]), dict(reads=reads, flanks=[0, 1], allowed=allowed)])
    lines = [ln.split(" ", 1)[1] for ln in r.stdout.strip().split("\n") if ln.startswith("batch ")]
    assert len(lines) == 2
    for o, ln in zip(py, lines):
        hc = hf = 1469598103934665603
        for b in o["consensus"].tolist():
            hc = ((hc ^ b) * 1099511628211) & M
        fl = o["flank_las"]
        for q, t in zip(fl.rec, fl.traces()):
            for v in (q["aread"], q["abpos"], q["aepos"], q["bbpos"], q["bepos"], q["diffs"], q["flags"], q["tlen"]):
                hf = ((hf ^ (int(v) & 0xffffffff)) * 1099511628211) & M
            for v in t.reshape(-1).tolist():
                hf = ((hf ^ v) * 1099511628211) & M
        assert ln == "%d|%d|%d|%d|%d|%d" % (o["status"], o["reference_read"], len(o["consensus"]), hc, len(fl), hf)
    assert py[0]["status"] == 0 and len(py[0]["flank_las"]) >= 2 and py[1]["reference_read"] in (3, -1)
