"""N > 1 on real GPUs: the in-library NCCL gather + placement merge (csrc/comm.cu) must give exactly the single-GPU
LAS.  Needs >= 2 visible GPUs (gpurun --gpus 2); skipped otherwise.  The CPU-side plumbing of the same exchange is
covered by tests/test_sharding_gloo.py (gloo, world size 2)."""
import os
import sys
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case():
    from dentist_b200 import synth
    sc = synth.make_scaffolds(2, 150000, 901, n_repeats=1, repeat_copies=4)
    ref, _ = synth.contigs_from(sc, synth.make_gaps(sc, 2, 902))
    reads, _ = synth.simulate_reads(sc, 3, 7000, 2500, 0.13, 903)
    return ref, reads


def _worker(rank, world, initfile, outdir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from dentist_b200 import dazzler, sharding, synth
    torch.cuda.set_device(rank)
    dazzler.init(rank)
    dist.init_process_group("nccl", init_method="file://" + initfile, rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    dazzler.comm_init(rank, world)
    ref, reads = _case()
    lo, hi = sharding.shard_ranges(reads.nreads, world)[rank]
    off = reads.off[lo:hi + 1] - reads.off[lo]; bases = reads.bases[reads.off[lo]:reads.off[hi]]
    ga = dazzler.Block(ref.off, ref.bases); gb = dazzler.Block(off, bases)
    for root in (0, -1, world - 1):
        rec, toff, tr, st = dazzler.align_blocks_gather(ga, gb, lo, root=root, tspace=100, minlen=500)
        got = root < 0 or root == rank
        assert (len(rec) > 0) == got and st["las"] > 0
        if got:
            np.save(os.path.join(outdir, "rec_%d_%d.npy" % (root, rank)), np.asarray(rec)); np.save(os.path.join(outdir, "tr_%d_%d.npy" % (root, rank)), np.asarray(tr))
            np.save(os.path.join(outdir, "toff_%d_%d.npy" % (root, rank)), np.asarray(toff))
    # host descriptors in, merged LAS out
    rec, toff, tr, st = dazzler.align_host_gather(dazzler.HostBlock(ref.off, ref.bases), dazzler.HostBlock(off, bases), lo, root=0, tspace=100, minlen=500)
    if rank == 0:
        np.save(os.path.join(outdir, "hrec.npy"), np.asarray(rec)); np.save(os.path.join(outdir, "htr.npy"), np.asarray(tr))
    parts = dazzler.comm_allgatherv(b"rank%d" % rank * (rank + 1))
    assert parts == [b"rank%d" % r * (r + 1) for r in range(world)]
    dist.barrier()
    dazzler.comm_shutdown()
    dist.destroy_process_group()


def test_in_library_gather_equals_single_gpu():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from dentist_b200 import dazzler
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, os.path.join(d, "init"), d), nprocs=world, join=True)
        ref, reads = _case()
        want = dazzler.align(dazzler.Block(ref.off, ref.bases), dazzler.Block(reads.off, reads.bases), tspace=100, minlen=500)
        assert len(want) > 100
        for root, ranks in ((0, [0]), (-1, [0, 1]), (1, [1])):
            for r in ranks:
                rec = np.load(os.path.join(d, "rec_%d_%d.npy" % (root, r))); tr = np.load(os.path.join(d, "tr_%d_%d.npy" % (root, r)))
                toff = np.load(os.path.join(d, "toff_%d_%d.npy" % (root, r)))
                assert rec.tobytes() == want.rec.tobytes() and np.array_equal(tr, want.trace) and np.array_equal(toff, want.toff)
        assert np.load(os.path.join(d, "hrec.npy")).tobytes() == want.rec.tobytes() and np.array_equal(np.load(os.path.join(d, "htr.npy")), want.trace)
