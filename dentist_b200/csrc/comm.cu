// comm.cu -- multi-GPU exchange inside the library: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// The reference fans the read blocks out as independent `damapper` jobs (Snakefile:1143-1170) and merges their LAS
// files with `LAmerge` through the file system (Snakefile:1173-1200).  Here every rank aligns ITS read block(s), the
// per-rank LAS segments go from HBM to HBM in ONE variable-size gather (counts by ncclAllGather, payload by grouped
// ncclSend / ncclRecv), and the merge is a placement pass: a read lives in exactly one block, so the segments' B read
// ranges are disjoint and the merged LAsort order is, per A read, the concatenation of the ranks' runs.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy a host such as PyTorch already loaded, DN_NCCL_LIB, or
// the system one), so the library has no link-time dependency on it and single-GPU users never touch it.
#include "api_internal.hpp"
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <atomic>
#include <chrono>
#include <new>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

using namespace dn;
using namespace dnapi;

namespace {

// ---- the slice of the NCCL ABI this file uses (nccl.h: ncclUniqueId is 128 opaque bytes, ncclUint8 = 1, ncclInt64 = 4)
struct NcclId { char internal[128]; };
typedef void *NcclComm;
struct Nccl {
    void *h = nullptr;
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
} N;
NcclComm g_comm = nullptr;
int g_rank = 0, g_world = 1;

bool load_nccl(std::string &err) {
    if (N.h) return true;
    const char *cands[] = {getenv("DN_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *c : cands) {
        if (!c || !*c) continue;
        N.h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
        if (N.h) break;
    }
    if (!N.h) { err = std::string("cannot load NCCL (set DN_NCCL_LIB): ") + (dlerror() ? dlerror() : ""); return false; }
#define SYM(field, name) *(void **)(&N.field) = dlsym(N.h, name); if (!N.field) { err = std::string("NCCL symbol missing: ") + name; N.h = nullptr; return false; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllGather, "ncclAllGather") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return true;
}

#define DN_NCCL(x) do { int r_ = (x); if (r_ != 0) throw dn::Error(std::string("NCCL error: ") + N.GetErrorString(r_)); } while (0)

__global__ void __launch_bounds__(256) k_shift_bread(dn_las_record *__restrict__ rec, int64_t n, int32_t shift) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rec[i].bread += shift;
}

// run[r * (na + 1) + a] = first record of segment r with aread >= a (a in [0, na]); one thread per (r, a)
__global__ void __launch_bounds__(256) k_seg_runs(const dn_las_record *__restrict__ rec, const int64_t *__restrict__ seg_beg, int world,
                                                  int64_t na, int32_t *__restrict__ run) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)world * (na + 1)) return;
    const int r = (int)(t / (na + 1)); const int64_t a = t % (na + 1);
    int64_t lo = seg_beg[r], hi = seg_beg[r + 1];
    while (lo < hi) { const int64_t m = (lo + hi) >> 1; if (rec[m].aread < a) lo = m + 1; else hi = m; }
    run[t] = (int32_t)(lo - seg_beg[r]);
}
__global__ void __launch_bounds__(256) k_seg_counts(const int32_t *__restrict__ run, int world, int64_t na, int32_t *__restrict__ cnt) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)world * na) return;
    const int64_t a = t / world; const int r = (int)(t % world);
    cnt[t] = run[(int64_t)r * (na + 1) + a + 1] - run[(int64_t)r * (na + 1) + a];
}
// record i of segment r goes to base[aread][r] + (its index inside the run); tl / src follow it
__global__ void __launch_bounds__(256) k_seg_place(const dn_las_record *__restrict__ rec, const int64_t *__restrict__ seg_beg, int world,
                                                   int64_t na, const int32_t *__restrict__ run, const int32_t *__restrict__ base,
                                                   int64_t n, dn_las_record *__restrict__ out, u32 *__restrict__ tl, int32_t *__restrict__ src) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int r = 0; while (r + 1 < world && seg_beg[r + 1] <= i) r++;
    const dn_las_record x = rec[i];
    const int64_t pos = (int64_t)base[(int64_t)x.aread * world + r] + ((i - seg_beg[r]) - run[(int64_t)r * (na + 1) + x.aread]);
    out[pos] = x; tl[pos] = (u32)x.tlen; src[pos] = (int32_t)i;
}
__global__ void __launch_bounds__(256) k_rec_tlen(const dn_las_record *__restrict__ rec, int64_t n, u32 *__restrict__ tl) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) tl[i] = (u32)rec[i].tlen;
}
// one warp per output record: its (diffs, bbases) pairs from the gathered trace buffer
__global__ void __launch_bounds__(256) k_seg_traces(const int32_t *__restrict__ src, int64_t n, const int64_t *__restrict__ src_toff,
                                                    const int64_t *__restrict__ dst_toff, const dn_las_record *__restrict__ out_rec,
                                                    const uint16_t *__restrict__ in, uint16_t *__restrict__ out) {
    const int64_t o = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; const int lane = threadIdx.x & 31;
    if (o >= n) return;
    const uint16_t *s = in + src_toff[src[o]];
    uint16_t *d = out + dst_toff[o];
    for (int q = lane; q < out_rec[o].tlen; q += 32) d[q] = s[q];
}

}  // namespace

namespace dn {

// the merged LAS of the gathered segments, still in HBM (arena memory)
struct MergedDev { dn_las_record *rec = nullptr; int64_t *toff = nullptr; uint16_t *trace = nullptr; int64_t n = 0, ntrace = 0; };

static void merge_segments_on_device(const dn_las_record *d_rec, const int64_t *h_seg_beg, int world, const uint16_t *d_trace, int64_t ntrace,
                                     int64_t na, MergedDev &M, cudaStream_t s) {
    const int64_t n = h_seg_beg[world];
    M = MergedDev(); M.n = n; M.ntrace = ntrace;
    if (n == 0) return;
    if (n >= (1ll << 31)) throw Error("too many records to merge");
    DBuf<int64_t> dseg(world + 1);
    DN_CUDA(cudaMemcpyAsync(dseg.p, h_seg_beg, sizeof(int64_t) * (world + 1), cudaMemcpyHostToDevice, s));
    const int64_t nrun = (int64_t)world * (na + 1), ncnt = (int64_t)world * na;
    DBuf<int32_t> run(nrun), cnt(ncnt + 1), base(ncnt + 1), src(n), tot32(1);
    DN_LAUNCH(k_seg_runs, (unsigned)((nrun + 255) / 256), 256, 0, s, d_rec, (const int64_t *)dseg.p, world, na, run.p);
    DN_LAUNCH(k_seg_counts, (unsigned)((ncnt + 255) / 256), 256, 0, s, (const int32_t *)run.p, world, na, cnt.p);
    exclusive_scan_i32(cnt.p, base.p, ncnt, tot32.p, s);
    DBuf<dn_las_record> orec(n); DBuf<u32> tl(n), stl(n); DBuf<int64_t> stoff(n), dtoff(n), tot(1);
    DN_LAUNCH(k_rec_tlen, (unsigned)((n + 255) / 256), 256, 0, s, d_rec, n, stl.p);
    exclusive_scan_u32_to_i64(stl.p, stoff.p, n, tot.p, s);
    DN_LAUNCH(k_seg_place, (unsigned)((n + 255) / 256), 256, 0, s, d_rec, (const int64_t *)dseg.p, world, na, (const int32_t *)run.p,
              (const int32_t *)base.p, n, orec.p, tl.p, src.p);
    exclusive_scan_u32_to_i64(tl.p, dtoff.p, n, tot.p, s);
    DBuf<uint16_t> otr((size_t)ntrace + 1);
    DN_LAUNCH(k_seg_traces, (unsigned)((n * 32 + 255) / 256), 256, 0, s, (const int32_t *)src.p, n, (const int64_t *)stoff.p,
              (const int64_t *)dtoff.p, (const dn_las_record *)orec.p, d_trace, otr.p);
    M.rec = orec.p; M.toff = dtoff.p; M.trace = otr.p;            // arena memory: stays valid until the next reset
}

void merge_segments_device(const dn_las_record *d_rec, const int64_t *h_seg_beg, int world, const uint16_t *d_trace, int64_t ntrace,
                           int64_t na, HostLas &out, cudaStream_t s) {
    const int64_t n = h_seg_beg[world];
    out = HostLas();
    out.rec = (dn_las_record *)hcache_alloc(sizeof(dn_las_record) * (size_t)(n + 1));
    out.toff = (int64_t *)hcache_alloc(sizeof(int64_t) * (size_t)(n + 1));
    out.trace = (uint16_t *)hcache_alloc(sizeof(uint16_t) * (size_t)(ntrace + 1));
    out.nrec = n; out.ntrace = ntrace;
    if (n == 0) return;
    MergedDev M;
    merge_segments_on_device(d_rec, h_seg_beg, world, d_trace, ntrace, na, M, s);
    DN_CUDA(cudaMemcpyAsync(out.rec, M.rec, sizeof(dn_las_record) * n, cudaMemcpyDeviceToHost, s));
    DN_CUDA(cudaMemcpyAsync(out.toff, M.toff, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, s));
    if (ntrace) DN_CUDA(cudaMemcpyAsync(out.trace, M.trace, sizeof(uint16_t) * ntrace, cudaMemcpyDeviceToHost, s));
    DN_CUDA(cudaStreamSynchronize(s));
}

}  // namespace dn

namespace {

// counts by ncclAllGather, payload by grouped send / recv: root >= 0: only the root receives; root < 0: every rank does
void gather_segments(const DevLas &mine, int root, std::vector<int64_t> &seg_beg, std::vector<int64_t> &tr_beg,
                     DBuf<dn_las_record> &allrec, DBuf<uint16_t> &alltr, cudaStream_t s) {
    const int W = g_world;
    DBuf<int64_t> dcnt(2), dall(2 * (size_t)W);
    const int64_t hc[2] = {mine.nrec, mine.ntrace};
    DN_CUDA(cudaMemcpyAsync(dcnt.p, hc, sizeof hc, cudaMemcpyHostToDevice, s));
    DN_NCCL(N.AllGather(dcnt.p, dall.p, 2, 4 /* ncclInt64 */, g_comm, s));
    std::vector<int64_t> h(2 * (size_t)W);
    DN_CUDA(cudaMemcpyAsync(h.data(), dall.p, sizeof(int64_t) * 2 * W, cudaMemcpyDeviceToHost, s));
    DN_CUDA(cudaStreamSynchronize(s));
    seg_beg.assign(W + 1, 0); tr_beg.assign(W + 1, 0);
    for (int r = 0; r < W; r++) { seg_beg[r + 1] = seg_beg[r] + h[2 * r]; tr_beg[r + 1] = tr_beg[r] + h[2 * r + 1]; }
    const bool recv = root < 0 || root == g_rank;
    if (recv) { allrec.alloc((size_t)seg_beg[W] + 1); alltr.alloc((size_t)tr_beg[W] + 1); }
    DN_NCCL(N.GroupStart());
    for (int r = 0; r < W; r++) {
        if (r == g_rank) continue;
        if (root < 0 || root == r) {             // r wants my segment
            if (mine.nrec) DN_NCCL(N.Send(mine.rec, sizeof(dn_las_record) * (size_t)mine.nrec, 1 /* ncclUint8 */, r, g_comm, s));
            if (mine.ntrace) DN_NCCL(N.Send(mine.trace, sizeof(uint16_t) * (size_t)mine.ntrace, 1, r, g_comm, s));
        }
        if (recv) {
            const int64_t nr = seg_beg[r + 1] - seg_beg[r], nt = tr_beg[r + 1] - tr_beg[r];
            if (nr) DN_NCCL(N.Recv(allrec.p + seg_beg[r], sizeof(dn_las_record) * (size_t)nr, 1, r, g_comm, s));
            if (nt) DN_NCCL(N.Recv(alltr.p + tr_beg[r], sizeof(uint16_t) * (size_t)nt, 1, r, g_comm, s));
        }
    }
    DN_NCCL(N.GroupEnd());
    if (recv) {
        if (mine.nrec) DN_CUDA(cudaMemcpyAsync(allrec.p + seg_beg[g_rank], mine.rec, sizeof(dn_las_record) * (size_t)mine.nrec, cudaMemcpyDeviceToDevice, s));
        if (mine.ntrace) DN_CUDA(cudaMemcpyAsync(alltr.p + tr_beg[g_rank], mine.trace, sizeof(uint16_t) * (size_t)mine.ntrace, cudaMemcpyDeviceToDevice, s));
    }
}

void to_out(HostLas &h, int tspace, dn_las_buf *out) {
    memset(out, 0, sizeof *out);
    out->nrec = h.nrec; out->ntrace = h.ntrace; out->tspace = tspace; out->stats = h.stats;
    out->rec = h.rec; out->toff = h.toff; out->trace = h.trace;
    h.rec = nullptr; h.toff = nullptr; h.trace = nullptr;
}

AlignParams params_of(const dn_align_params *p) {
    dn_align_params d; dn_align_params_default(&d);
    if (p) d = *p;
    AlignParams q;
    q.k = d.k; q.w = d.w; q.h = d.h; q.t = d.t; q.tspace = d.tspace; q.minlen = d.minlen;
    double e = d.e; if (e < 0.7) e = 0.7; if (e > 0.99) e = 0.99;
    q.cdiff = (int)(6.0 / (1.0 - e) + 0.5);
    q.xdrop = d.xdrop; q.wmax = d.wmax; q.rounds = d.rounds; q.poolmul = d.poolmul;
    q.self = (d.self_block && !d.identity) ? 1 : 0;
    q.join_mode = d.join_mode;
    return q;
}

// ---- sharded download (root >= 0, world > 1) ------------------------------------------------------------------
// The root would have to pull the whole merged LAS (world x one rank's result) through its ONE PCIe link.  Instead every
// rank receives all segments over NVLink (cheap), runs the same placement merge, and downloads only ITS 1/world slice of
// the merged arrays -- through its OWN PCIe link -- into a host segment shared by the ranks of the node (POSIX shared
// memory, page-locked by every rank).  The root's result points into that segment; two halves alternate, so a result
// stays valid until the second following gather call.  Any failure to set the segment up (no /dev/shm space, ...) is
// agreed on by all ranks at dn_comm_init and falls back to the root-only download.
struct ShmHeader { std::atomic<unsigned long long> arrivals[2]; };
struct Shm {
    bool ok = false; int fd = -1; uint8_t *base = nullptr; size_t size = 0, half = 0; std::string name; unsigned long long calls[2] = {0, 0};
    unsigned long long seq = 0;
} g_shm;
constexpr size_t SHM_HEADER = 4096;

// Results handed to the caller point into the segment and may be released (dn_las_free) long after dn_comm_shutdown: the
// mapping is therefore never unmapped before the process ends -- only the descriptor and the name go -- and every range
// ever mapped stays known to comm_owns_host_pointer.
std::vector<std::pair<const uint8_t *, size_t>> g_shm_ranges;
void shm_close() {
    if (g_shm.fd >= 0) close(g_shm.fd);
    if (!g_shm.name.empty() && g_rank == 0) shm_unlink(g_shm.name.c_str());
    g_shm = Shm();
}

// rank 0 creates the segment BEFORE it enters ncclCommInitRank; the others open it AFTER they left it
bool shm_create_or_open(const uint8_t *id, int rank, bool create) {
    unsigned long long h = 1469598103934665603ull;
    for (int i = 0; i < DN_COMM_ID_BYTES; i++) h = (h ^ id[i]) * 1099511628211ull;
    char nm[64]; snprintf(nm, sizeof nm, "/dn_b200_%016llx", h);
    g_shm.name = nm;
    const char *mb = getenv("DN_SHM_MB");
    const size_t half = (size_t)(mb ? atoll(mb) : 384) << 20;
    if (half == 0) return false;
    g_shm.half = half; g_shm.size = SHM_HEADER + 2 * half;
    g_shm.fd = shm_open(nm, create ? (O_CREAT | O_RDWR | O_TRUNC) : O_RDWR, 0600);
    if (g_shm.fd < 0) return false;
    if (create && (ftruncate(g_shm.fd, (off_t)g_shm.size) != 0 || posix_fallocate(g_shm.fd, 0, (off_t)g_shm.size) != 0)) return false;
    void *p = mmap(nullptr, g_shm.size, PROT_READ | PROT_WRITE, MAP_SHARED, g_shm.fd, 0);
    if (p == MAP_FAILED) return false;
    g_shm.base = (uint8_t *)p;
    g_shm_ranges.emplace_back(g_shm.base, g_shm.size);
    if (create) { ShmHeader *H = new (g_shm.base) ShmHeader; H->arrivals[0].store(0); H->arrivals[1].store(0); }
    if (cudaHostRegister(g_shm.base, g_shm.size, cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); return false; }
    (void)rank;
    return true;
}

bool shm_contains(const void *p) {
    for (const auto &r : g_shm_ranges) if ((const uint8_t *)p >= r.first && (const uint8_t *)p < r.first + r.second) return true;
    return false;
}

// after the local alignment: shift bread to the global numbering, gather, merge on the receiving ranks
void gather_and_merge(DevLas &mine, HostLas &h, int64_t bread_offset, int root, int64_t na_reads, cudaStream_t s) {
    if (mine.nrec && bread_offset)
        DN_LAUNCH(k_shift_bread, (unsigned)((mine.nrec + 255) / 256), 256, 0, s, mine.rec, mine.nrec, (int32_t)bread_offset);
    const bool want_shard = root >= 0 && g_world > 1 && g_shm.ok;
    std::vector<int64_t> seg_beg, tr_beg; DBuf<dn_las_record> allrec; DBuf<uint16_t> alltr;
    const bool trace = getenv("DN_TRACE") != nullptr && g_rank == 0;      // diagnostics only (adds stream syncs)
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    if (trace) cudaStreamSynchronize(s);
    const auto ts0 = now();
    gather_segments(mine, want_shard ? -1 : root, seg_beg, tr_beg, allrec, alltr, s);
    if (trace) cudaStreamSynchronize(s);
    const auto ts1 = now();
    const dn_align_stats st = h.stats;
    const int W = g_world;
    const int64_t n = seg_beg[W], nt = tr_beg[W];
    const size_t o_toff = ((size_t)n * sizeof(dn_las_record) + 255) & ~(size_t)255, o_tr = (o_toff + (size_t)n * 8 + 255) & ~(size_t)255;
    const bool shard = want_shard && o_tr + (size_t)nt * 2 + 256 <= g_shm.half;       // every rank knows the sizes: same decision everywhere
    if (shard) {
        const unsigned long long seq = g_shm.seq++; const int hf = (int)(seq & 1);
        uint8_t *dst = g_shm.base + SHM_HEADER + (size_t)hf * g_shm.half;
        MergedDev M;
        merge_segments_on_device(allrec.p, seg_beg.data(), W, alltr.p, nt, na_reads, M, s);
        if (trace) cudaStreamSynchronize(s);
        const auto ts2 = now();
        const int64_t r0 = n * g_rank / W, r1 = n * (g_rank + 1) / W, t0 = nt * g_rank / W, t1 = nt * (g_rank + 1) / W;
        if (r1 > r0) {
            DN_CUDA(cudaMemcpyAsync(dst + (size_t)r0 * sizeof(dn_las_record), M.rec + r0, sizeof(dn_las_record) * (size_t)(r1 - r0), cudaMemcpyDeviceToHost, s));
            DN_CUDA(cudaMemcpyAsync(dst + o_toff + (size_t)r0 * 8, M.toff + r0, 8 * (size_t)(r1 - r0), cudaMemcpyDeviceToHost, s));
        }
        if (t1 > t0) DN_CUDA(cudaMemcpyAsync(dst + o_tr + (size_t)t0 * 2, M.trace + t0, 2 * (size_t)(t1 - t0), cudaMemcpyDeviceToHost, s));
        DN_CUDA(cudaStreamSynchronize(s));
        const auto ts3 = now();
        ShmHeader *H = (ShmHeader *)g_shm.base;
        H->arrivals[hf].fetch_add(1, std::memory_order_release);
        g_shm.calls[hf]++;
        if (g_rank == root) {
            const unsigned long long want = g_shm.calls[hf] * (unsigned long long)W;
            while (H->arrivals[hf].load(std::memory_order_acquire) < want) { /* the other ranks' slices land within microseconds of ours */ }
            if (trace) fprintf(stderr, "[dn trace] gather: exchange %.3f ms (%lld records, %lld trace values from %d ranks), placement merge %.3f ms, "
                               "slice download %.3f ms, wait for the other slices %.3f ms\n", ms(ts0, ts1), (long long)n, (long long)nt, W, ms(ts1, ts2), ms(ts2, ts3), ms(ts3, now()));
            hcache_free(h.rec); hcache_free(h.toff); hcache_free(h.trace);
            h.rec = (dn_las_record *)dst; h.toff = (int64_t *)(dst + o_toff); h.trace = (uint16_t *)(dst + o_tr);
            h.nrec = n; h.ntrace = nt;
        }
    } else if (root < 0 || root == g_rank || want_shard) {
        // (want_shard but too large for the shared segment: every rank holds all segments; only the root merges and downloads)
        if (root < 0 || root == g_rank) {
            hcache_free(h.rec); hcache_free(h.toff); hcache_free(h.trace); h.rec = nullptr; h.toff = nullptr; h.trace = nullptr;
            merge_segments_device(allrec.p, seg_beg.data(), W, alltr.p, nt, na_reads, h, s);
        } else DN_CUDA(cudaStreamSynchronize(s));
    } else DN_CUDA(cudaStreamSynchronize(s));
    h.stats = st;                                 // the statistics stay this rank's own
}

}  // namespace

namespace dn { bool comm_owns_host_pointer(const void *p) { return shm_contains(p); } }

extern "C" {

int dn_comm_get_id(uint8_t *id) {
    if (!id) return fail(DN_ERR_INVALID, "null argument");
    std::string err;
    if (!load_nccl(err)) return fail(DN_ERR_INVALID, err);
    NcclId u; memset(&u, 0, sizeof u);
    if (int r = N.GetUniqueId(&u)) return fail(DN_ERR_CUDA, std::string("ncclGetUniqueId: ") + N.GetErrorString(r));
    memcpy(id, &u, DN_COMM_ID_BYTES);
    return DN_OK;
}

int dn_comm_init(int32_t rank, int32_t world, const uint8_t *id) {
    if (!id || world < 1 || rank < 0 || rank >= world) return fail(DN_ERR_INVALID, "bad rank / world / id");
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = ensure_device()) return rc;
    return guarded([&]() -> int {
        std::string err;
        if (!load_nccl(err)) return fail(DN_ERR_INVALID, err);
        cudaSetDevice(g_device);
        if (g_comm) { N.CommDestroy(g_comm); g_comm = nullptr; }
        shm_close();
        g_rank = rank; g_world = world;
        const bool use_shm = world > 1 && !getenv("DN_NO_SHM");
        bool mine_ok = false;
        if (use_shm && rank == 0) mine_ok = shm_create_or_open(id, rank, true);       // before the collective init: the others open it after
        NcclId u; memcpy(&u, id, DN_COMM_ID_BYTES);
        DN_NCCL(N.CommInitRank(&g_comm, world, u, rank));
        if (use_shm && rank != 0) mine_ok = shm_create_or_open(id, rank, false);
        if (world > 1) {
            // all ranks must agree: one rank without the segment switches everybody to the root-only download
            arena().reset();
            DBuf<int64_t> f(1), all(world);
            const int64_t v = mine_ok ? 1 : 0;
            DN_CUDA(cudaMemcpyAsync(f.p, &v, 8, cudaMemcpyHostToDevice, g_stream));
            DN_NCCL(N.AllGather(f.p, all.p, 1, 4, g_comm, g_stream));
            std::vector<int64_t> hv(world);
            DN_CUDA(cudaMemcpyAsync(hv.data(), all.p, 8 * (size_t)world, cudaMemcpyDeviceToHost, g_stream));
            DN_CUDA(cudaStreamSynchronize(g_stream));
            bool all_ok = use_shm;
            for (int64_t x : hv) all_ok = all_ok && x == 1;
            g_shm.ok = all_ok;
            // every rank holds its mapping now: the name can go, so that a crashed job leaves nothing behind in /dev/shm
            if (rank == 0 && !g_shm.name.empty()) { shm_unlink(g_shm.name.c_str()); }
            g_shm.name.clear();
            if (!all_ok) shm_close();
            if (!all_ok && use_shm && getenv("DN_TRACE")) fprintf(stderr, "[dn trace] shared host segment unavailable: root-only download\n");
        }
        return DN_OK;
    });
}

int dn_comm_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_comm) { if (g_stream) cudaStreamSynchronize(g_stream); N.CommDestroy(g_comm); g_comm = nullptr; }
    shm_close();
    g_rank = 0; g_world = 1;
    return DN_OK;
}
int64_t dn_comm_shared_segment_bytes(void) { return g_shm.ok ? (int64_t)g_shm.size : 0; }
int32_t dn_comm_rank(void) { return g_rank; }
int32_t dn_comm_size(void) { return g_world; }

int dn_align_blocks_gather(const dn_block *a, const dn_block *b, const dn_align_params *p, int64_t bread_offset, int32_t root, dn_las_buf *out) {
    if (!a || !b || !out) return fail(DN_ERR_INVALID, "null argument");
    if (!g_comm) return fail(DN_ERR_INVALID, "dn_comm_init has not been called");
    if (root >= g_world) return fail(DN_ERR_INVALID, "root out of range");
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = ensure_device()) return rc;
    return guarded([&] {
        cudaSetDevice(g_device);
        const AlignParams q = params_of(p);
        HostLas h; DevLas mine;
        align_blocks(a->b, b->b, q, h, g_stream, &mine);
        gather_and_merge(mine, h, bread_offset, root, a->b.nreads, g_stream);
        to_out(h, q.tspace, out);
        return DN_OK;
    });
}

int dn_align_host_gather(const dn_block_desc *a, const dn_block_desc *b, const dn_align_params *p, int64_t bread_offset, int32_t root, dn_las_buf *out) {
    if (!a || !b || !out) return fail(DN_ERR_INVALID, "null argument");
    if (!g_comm) return fail(DN_ERR_INVALID, "dn_comm_init has not been called");
    if (root >= g_world) return fail(DN_ERR_INVALID, "root out of range");
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = ensure_device()) return rc;
    return guarded([&] {
        cudaSetDevice(g_device);
        static cudaStream_t copy_stream = nullptr;
        if (!copy_stream) DN_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        static cudaStream_t pack_stream = nullptr;           // B's chunks are packed here as they arrive on the copy stream
        if (!pack_stream) DN_CUDA(cudaStreamCreateWithFlags(&pack_stream, cudaStreamNonBlocking));
        dn_block ba, bb;
        block_upload(*a, ba.b, g_stream, true);              // A first: its small copy must not queue behind B's
        block_upload(*b, bb.b, copy_stream, true, pack_stream);   // B's upload overlaps A's indexing; the count pass follows its chunks
        const AlignParams q = params_of(p);
        HostLas h; DevLas mine;
        try { align_blocks(ba.b, bb.b, q, h, g_stream, &mine); }
        catch (...) { cudaStreamSynchronize(copy_stream); cudaStreamSynchronize(pack_stream); throw; }
        cudaStreamSynchronize(copy_stream); cudaStreamSynchronize(pack_stream);
        gather_and_merge(mine, h, bread_offset, root, ba.b.nreads, g_stream);
        to_out(h, q.tspace, out);
        return DN_OK;
    });
}

// Variable-size all-gather of host byte buffers through HBM staging (e.g. the InsertionDb bytes of every rank's
// pile-up batch: what `merge-insertions` collects from files, commands/mergeInsertions.d).  *recv holds the ranks'
// buffers back to back, counts[r] their sizes; free with dn_free.
int dn_comm_allgatherv(const void *send, int64_t nbytes, void **recv, int64_t *counts) {
    if (nbytes < 0 || (nbytes && !send) || !recv || !counts) return fail(DN_ERR_INVALID, "null argument");
    if (!g_comm) return fail(DN_ERR_INVALID, "dn_comm_init has not been called");
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = ensure_device()) return rc;
    return guarded([&] {
        cudaSetDevice(g_device); arena().reset();
        cudaStream_t s = g_stream;
        const int W = g_world;
        DBuf<int64_t> dcnt(1), dall(W);
        DN_CUDA(cudaMemcpyAsync(dcnt.p, &nbytes, 8, cudaMemcpyHostToDevice, s));
        DN_NCCL(N.AllGather(dcnt.p, dall.p, 1, 4, g_comm, s));
        DN_CUDA(cudaMemcpyAsync(counts, dall.p, sizeof(int64_t) * W, cudaMemcpyDeviceToHost, s));
        DN_CUDA(cudaStreamSynchronize(s));
        std::vector<int64_t> beg(W + 1, 0);
        for (int r = 0; r < W; r++) beg[r + 1] = beg[r] + counts[r];
        DBuf<uint8_t> mine((size_t)nbytes + 1), all((size_t)beg[W] + 1);
        if (nbytes) DN_CUDA(cudaMemcpyAsync(mine.p, send, (size_t)nbytes, cudaMemcpyHostToDevice, s));
        DN_NCCL(N.GroupStart());
        for (int r = 0; r < W; r++) {
            if (r == g_rank) continue;
            if (nbytes) DN_NCCL(N.Send(mine.p, (size_t)nbytes, 1, r, g_comm, s));
            if (counts[r]) DN_NCCL(N.Recv(all.p + beg[r], (size_t)counts[r], 1, r, g_comm, s));
        }
        DN_NCCL(N.GroupEnd());
        if (nbytes) DN_CUDA(cudaMemcpyAsync(all.p + beg[g_rank], mine.p, (size_t)nbytes, cudaMemcpyDeviceToDevice, s));
        uint8_t *hbuf = (uint8_t *)hcache_alloc((size_t)beg[W] + 1);
        if (beg[W]) DN_CUDA(cudaMemcpyAsync(hbuf, all.p, (size_t)beg[W], cudaMemcpyDeviceToHost, s));
        DN_CUDA(cudaStreamSynchronize(s));
        *recv = hbuf;
        return DN_OK;
    });
}

}  // extern "C"
