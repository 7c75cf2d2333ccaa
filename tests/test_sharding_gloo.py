"""N>1 plumbing on CPU: world_size-2 gloo run of the LAS all-gatherv + merge (what replaces LAmerge)."""
import os
import tempfile

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from dentist_b200 import sharding
from dentist_b200._lib import REC_DTYPE


def _fake_segment(seed, n, nreads_a=7, nreads_b=5):
    rng = np.random.default_rng(seed)
    rec = np.zeros(n, REC_DTYPE)
    rec["aread"] = rng.integers(0, nreads_a, n); rec["bread"] = rng.integers(0, nreads_b, n)
    rec["flags"] = rng.integers(0, 2, n); rec["abpos"] = rng.integers(0, 5000, n)
    rec["aepos"] = rec["abpos"] + rng.integers(100, 900, n)
    rec["bbpos"] = rng.integers(0, 5000, n); rec["bepos"] = rec["bbpos"] + rng.integers(100, 900, n)
    nt = -(-rec["aepos"] // 100) - rec["abpos"] // 100
    rec["tlen"] = 2 * nt
    order = np.lexsort((rec["abpos"], rec["flags"] & 1, rec["bread"], rec["aread"]))
    rec = rec[order]
    tr = rng.integers(0, 120, int(rec["tlen"].sum())).astype(np.uint16)
    ends = np.cumsum(rec["tlen"])
    rec["diffs"] = [int(tr[e - l:e:2].sum()) for e, l in zip(ends, rec["tlen"])]
    return rec, tr


def _worker(rank, world, initfile, outdir):
    dist.init_process_group("gloo", init_method="file://" + initfile, rank=rank, world_size=world)
    rec, tr = _fake_segment(100 + rank, 40 + 13 * rank)
    mrec, mtoff, mtr = sharding.gather_las(rec, tr, bread_offset=5 * rank, device="cpu")
    np.save(os.path.join(outdir, "rec%d.npy" % rank), mrec)
    np.save(os.path.join(outdir, "toff%d.npy" % rank), mtoff)
    np.save(os.path.join(outdir, "tr%d.npy" % rank), mtr)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_everything_once():
    for n, w in [(10, 1), (10, 3), (3, 8), (20070, 8)]:
        r = sharding.shard_ranges(n, w)
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert max(e - s for s, e in r) - min(e - s for s, e in r) <= 1


def test_gather_las_world2_gloo():
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, os.path.join(d, "init"), d), nprocs=2, join=True)
        got = [(np.load(os.path.join(d, "rec%d.npy" % r)), np.load(os.path.join(d, "toff%d.npy" % r)),
                np.load(os.path.join(d, "tr%d.npy" % r))) for r in range(2)]
    # both ranks hold the same merged LAS
    assert got[0][0].tobytes() == got[1][0].tobytes() and np.array_equal(got[0][2], got[1][2])
    # and it equals the single-process merge of the two segments
    segs = [_fake_segment(100 + r, 40 + 13 * r) for r in range(2)]
    for r, (rec, _) in enumerate(segs):
        rec["bread"] += 5 * r
    erec, etoff, etr = sharding.merge_las([s[0] for s in segs], [s[1] for s in segs])
    mrec, mtoff, mtr = got[0]
    assert mrec.tobytes() == erec.tobytes() and np.array_equal(mtoff, etoff) and np.array_equal(mtr, etr)
    assert len(mrec) == 40 + 53
    key = np.stack([mrec["aread"], mrec["bread"], mrec["flags"] & 1, mrec["abpos"]], 1).tolist()
    assert key == sorted(key)
    # every record still owns its own trace
    for i in range(len(mrec)):
        t = mtr[mtoff[i]:mtoff[i] + mrec[i]["tlen"]]
        assert int(t[0::2].sum()) == mrec[i]["diffs"]


def test_bench_deals_blocks_and_scaffolds_once():
    """bench.py's dealing of configs[2] read blocks and configs[3] scaffolds over the ranks: contiguous, complete, disjoint, balanced."""
    import bench
    for world in (1, 2, 3, 4, 8):
        for n in (8, 15, 200, 201):
            for deal in (bench.c3_blocks_of, bench.c4_scaffolds_of):
                parts = [list(deal(r, world, n)) for r in range(world)]
                flat = [x for p in parts for x in p]
                assert flat == list(range(n)), (deal.__name__, world, n)
                sizes = [len(p) for p in parts]
                assert max(sizes) - min(sizes) <= 1 or deal is bench.c3_blocks_of
