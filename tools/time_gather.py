"""Where the multi-GPU end-to-end step goes: dn_align_host_gather under torchrun, DN_TRACE=1 prints rank 0's split of the
exchange (counts + payload), the placement merge, the slice download and the wait for the other ranks' slices.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/time_gather.py [iterations]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
import bench
from dentist_b200 import dazzler, synth

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dazzler.init(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    dazzler.comm_init(rank, world)
ref, reads = bench.make_workload(1.0, rank)
def pinned(a):
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
def bps_of(blk):
    parts, boff, o = [], [], 0
    for r in range(blk.nreads):
        p = synth.pack_2bit_dazz(blk.read(r)); boff.append(o); parts.append(p); o += len(p)
    return pinned(np.concatenate(parts)), np.array(boff, np.int64)
rb, ro = bps_of(ref); qb, qo = bps_of(reads)
off = 0
if world > 1:
    cnts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(cnts, torch.tensor([reads.nreads], dtype=torch.int64, device=dev))
    off = int(sum(int(c.item()) for c in cnts[:rank]))
def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    barrier(); t0 = time.perf_counter()
    a = dazzler.HostBlock(ref.off, bps=rb, boff=ro); b = dazzler.HostBlock(reads.off, bps=qb, boff=qo)
    if world > 1:
        rec, toff, tr, st = dazzler.align_host_gather(a, b, off, root=0, **bench.PARAMS)
    else:
        rec, toff, tr, st = dazzler.align_host(a, b, **bench.PARAMS)
    t1 = time.perf_counter(); barrier(); t2 = time.perf_counter()
    print("rank %d iter %d: call %.2f ms (device ms_total %.2f), barrier after it %.2f ms" % (rank, it, (t1 - t0) * 1e3, st["ms_total"], (t2 - t1) * 1e3), flush=True)
if world > 1:
    dist.barrier(); dazzler.comm_shutdown(); dist.destroy_process_group()
