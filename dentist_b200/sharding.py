"""Multi-GPU plumbing: read blocks shard across ranks with no data-path collective (the reference's
Snakemake per-block fan-out, Snakefile:1143-1170); the per-rank LAS segments are merged by ONE
all-gatherv at the end (what `LAmerge` does through the file system, Snakefile:1173-1200).

torch.distributed is plumbing only: `nccl` over NVLink on the GPU box, `gloo` in the CPU tests.
"""
import numpy as np
import torch
import torch.distributed as dist

from ._lib import REC_DTYPE


def shard_ranges(n_units, world):
    """Contiguous, balanced ranges of read indices per rank (block order is preserved so that the
    concatenation of per-rank `Q.R.las` segments is already sorted by read)."""
    base, rem = divmod(n_units, world)
    out, s = [], 0
    for r in range(world):
        e = s + base + (1 if r < rem else 0)
        out.append((s, e))
        s = e
    return out


def _allgatherv_bytes(buf_u8, device, keep_on_device=False):
    """all-gather of variable-length byte buffers: sizes first, then one padded all_gather."""
    world = dist.get_world_size()
    n = torch.tensor([buf_u8.numel()], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    m = max(max(sizes), 1)
    pad = torch.zeros(m, dtype=torch.uint8, device=device)
    pad[:buf_u8.numel()] = buf_u8.to(device)
    outs = [torch.empty(m, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(outs, pad)
    if keep_on_device:
        return [o[:s] for o, s in zip(outs, sizes)], sizes
    return [o[:s].cpu().numpy() for o, s in zip(outs, sizes)], sizes


def gather_las(rec, trace, bread_offset, device="cpu", tspace=100, bounds=None, root=None):
    """All ranks contribute their LAS segment (records with rank-local `bread`, traces in record order); every
    rank gets the merged LAS in LAsort order (aread, bread, comp, abpos, ...), with `bread` shifted to the global
    read numbering.  Returns (records, trace offsets, trace).
    On CUDA the gathered bytes never leave the device before they are merged (dn_las_merge_device);
    `bounds` = (max_alen, max_blen, na_reads, nb_reads_total) sizes the sort keys.  On CPU (gloo tests) the
    merge is the numpy `merge_las`.  root=None: every rank merges (all ranks hold R.Q.las); root=r: all ranks take part
    in the all-gatherv but only rank r merges and downloads (the others return None) -- one merged file, as LAmerge writes."""
    rec = rec.copy()
    rec["bread"] += bread_offset
    rb = torch.from_numpy(np.frombuffer(rec.tobytes(), dtype=np.uint8).copy())
    tb = torch.from_numpy(np.frombuffer(np.ascontiguousarray(trace, dtype=np.uint16).tobytes(), dtype=np.uint8).copy())
    is_cuda = torch.device(device).type == "cuda"
    recs, rsz = _allgatherv_bytes(rb, device, keep_on_device=is_cuda)
    trs, tsz = _allgatherv_bytes(tb, device, keep_on_device=is_cuda)
    is_cuda_ = is_cuda
    if bounds is None and is_cuda_:      # key widths from the data: max over ranks of the coordinates / ids present
        loc = [int(rec[f].max()) + 1 if len(rec) else 1 for f in ("aepos", "bepos", "aread", "bread")]
        mx = torch.tensor(loc, dtype=torch.int64, device=device)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        bounds = tuple(int(v) for v in mx.tolist())
    if root is not None and dist.get_rank() != root:
        return None
    if not is_cuda:
        return merge_las([r.view(REC_DTYPE) for r in recs], [t.view(np.uint16) for t in trs])
    import ctypes as C
    from . import _lib, dazzler
    drec = torch.cat(recs) if recs else torch.zeros(0, dtype=torch.uint8, device=device)
    dtr = torch.cat(trs) if trs else torch.zeros(0, dtype=torch.uint8, device=device)
    torch.cuda.synchronize()
    buf = _lib.LasBuf()
    _lib.check(_lib.lib().dn_las_merge_device(C.c_void_p(drec.data_ptr()), sum(rsz) // 40, C.c_void_p(dtr.data_ptr()), sum(tsz) // 2,
                                              int(tspace), int(bounds[0]), int(bounds[1]), int(bounds[2]), int(bounds[3]), C.byref(buf)))
    las = dazzler.Las(buf)
    return las.rec, las.toff, las.trace


def merge_las(recs, traces):
    """k-way merge of LAS segments (each in LAsort order) into one (LAmerge semantics)."""
    toffs = []
    base = 0
    for r, t in zip(recs, traces):
        o = np.cumsum(r["tlen"], dtype=np.int64) - r["tlen"] + base
        toffs.append(o)
        base += len(t)
    rec = np.concatenate(recs) if recs else np.zeros(0, REC_DTYPE)
    tr = np.concatenate(traces) if traces else np.zeros(0, np.uint16)
    src = np.concatenate(toffs) if toffs else np.zeros(0, np.int64)
    if len(rec) == 0:
        return rec, np.zeros(0, np.int64), tr
    # base.d:1787-1809 order: (aread, bread, comp, abpos, aepos, bbpos, bepos, diffs); lexsort: last key first
    order = np.lexsort((rec["diffs"], rec["bepos"], rec["bbpos"], rec["aepos"], rec["abpos"],
                        rec["flags"] & 1, rec["bread"], rec["aread"]))
    rec = rec[order]
    src = src[order]
    tlen = rec["tlen"].astype(np.int64)
    dst = np.cumsum(tlen) - tlen
    idx = np.repeat(src - dst, tlen) + np.arange(int(tlen.sum()))
    return rec, dst, tr[idx]
