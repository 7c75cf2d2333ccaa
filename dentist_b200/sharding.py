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


def _gather_device(rec, trace, bread_offset, device):
    """CUDA path of gather_las: one H2D per array straight from the (pinned) result views, `bread` shifted on the
    device, ONE size exchange, two padded all_gather_into_tensor calls.  Returns (record bytes, trace bytes, sizes)."""
    world = dist.get_world_size()
    rb = torch.from_numpy(np.ascontiguousarray(rec).view(np.uint8).reshape(-1)).to(device, non_blocking=True)
    tb = torch.from_numpy(np.ascontiguousarray(trace, dtype=np.uint16).view(np.uint8).reshape(-1)).to(device, non_blocking=True)
    if len(rec) and bread_offset:
        rb.view(torch.int32).view(-1, 10)[:, 8] += int(bread_offset)          # dn_las_record.bread: int32 at byte 32
    n = torch.tensor([rb.numel(), tb.numel()], dtype=torch.int64, device=device)
    sizes = torch.empty(world * 2, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(sizes, n)
    sizes = sizes.view(world, 2).tolist()
    outs = []
    for col, buf in ((0, rb), (1, tb)):
        m = max(max(sz[col] for sz in sizes), 1)
        pad = torch.empty(m, dtype=torch.uint8, device=device)
        pad[:buf.numel()] = buf
        allb = torch.empty(world * m, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(allb, pad)
        outs.append((allb, m))
    return outs, sizes


def gather_las(rec, trace, bread_offset, device="cpu", tspace=100, bounds=None, root=None):
    """All ranks contribute their LAS segment (records with rank-local `bread`, traces in record order); every
    rank gets the merged LAS in LAsort order (aread, bread, comp, abpos, ...), with `bread` shifted to the global
    read numbering.  Returns (records, trace offsets, trace).
    On CUDA the gathered bytes never leave the device before they are merged (dn_las_merge_device);
    `bounds` = (max_alen, max_blen, na_reads, nb_reads_total) sizes the sort keys.  On CPU (gloo tests) the
    merge is the numpy `merge_las`.  root=None: every rank merges (all ranks hold R.Q.las); root=r: all ranks take part
    in the all-gatherv but only rank r merges and downloads (the others return None) -- one merged file, as LAmerge writes."""
    is_cuda = torch.device(device).type == "cuda"
    if not is_cuda:
        rec = rec.copy()
        rec["bread"] += bread_offset
        rb = torch.from_numpy(np.frombuffer(rec.tobytes(), dtype=np.uint8).copy())
        tb = torch.from_numpy(np.frombuffer(np.ascontiguousarray(trace, dtype=np.uint16).tobytes(), dtype=np.uint8).copy())
        recs, rsz = _allgatherv_bytes(rb, device)
        trs, tsz = _allgatherv_bytes(tb, device)
        if root is not None and dist.get_rank() != root:
            return None
        return merge_las([r.view(REC_DTYPE) for r in recs], [t.view(np.uint16) for t in trs])
    if bounds is None:                   # key widths from the data: max over ranks of the coordinates / ids present
        loc = [int(rec[f].max()) + 1 if len(rec) else 1 for f in ("aepos", "bepos", "aread")] + \
              [int(rec["bread"].max()) + 1 + int(bread_offset) if len(rec) else 1]
        mx = torch.tensor(loc, dtype=torch.int64, device=device)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        bounds = tuple(int(v) for v in mx.tolist())
    ((rall, rm), (tall, tm)), sizes = _gather_device(rec, trace, bread_offset, device)
    if root is not None and dist.get_rank() != root:
        return None
    import ctypes as C
    from . import _lib, dazzler
    world = len(sizes)
    drec = torch.cat([rall[r * rm:r * rm + sizes[r][0]] for r in range(world)])
    dtr = torch.cat([tall[r * tm:r * tm + sizes[r][1]] for r in range(world)])
    torch.cuda.synchronize()
    buf = _lib.LasBuf()
    _lib.check(_lib.lib().dn_las_merge_device(C.c_void_p(drec.data_ptr()), drec.numel() // 40, C.c_void_p(dtr.data_ptr()), dtr.numel() // 2,
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
