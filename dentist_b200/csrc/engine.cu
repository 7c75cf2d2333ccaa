// engine.cu -- host orchestration of one block-vs-block alignment job on the current device.
// Everything between "blocks resident in HBM" and "sorted LAS records on the host" happens here;
// the only host-side arithmetic is the final ordering/serialisation of the (few) surviving records.
#include "seed.cuh"
#include <algorithm>
#include <string.h>
#include <stdlib.h>
#include <chrono>
#include <mutex>

namespace dn {

int ext_warps_per_cta();
int ext_ctas_per_sm();

namespace {

template <typename T> T d2h_scalar(const T *d, cudaStream_t s) {
    T v; DN_CUDA(cudaMemcpyAsync(&v, d, sizeof(T), cudaMemcpyDeviceToHost, s)); DN_CUDA(cudaStreamSynchronize(s)); return v;
}
int bits_for(uint64_t v) { int b = 0; while (v) { b++; v >>= 1; } return b; }   // bits needed to represent v

// DN_TRACE=1: print synchronised wall-clock per stage to stderr (diagnostics only; adds syncs)
struct Trace {
    bool on; cudaStream_t s; std::chrono::steady_clock::time_point t0;
    explicit Trace(cudaStream_t s_) : on(getenv("DN_TRACE") != nullptr), s(s_) { if (on) { cudaStreamSynchronize(s); t0 = std::chrono::steady_clock::now(); } }
    void mark(const char *what) {
        if (!on) return;
        cudaStreamSynchronize(s);
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[dn trace] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

struct Timer {
    cudaEvent_t a, b; cudaStream_t s;
    Timer(cudaStream_t s_) : s(s_) { cudaEventCreate(&a); cudaEventCreate(&b); }
    ~Timer() { cudaEventDestroy(a); cudaEventDestroy(b); }
    void start() { cudaEventRecord(a, s); }
    float stop() { cudaEventRecord(b, s); cudaEventSynchronize(b); float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }
};

}  // namespace

namespace {
struct HBlock { size_t cap; size_t pinned; size_t pad_[6]; };   // 64-byte header in front of every cached block
std::mutex g_hmu; std::vector<HBlock *> g_hfree;
}
// Result buffers are page-locked (D2H lands in them directly at PCIe speed) and recycled: both a fresh
// multi-MB malloc (page faults) and a cudaHostAlloc cost milliseconds.
void *hcache_alloc(size_t bytes) {
    if (bytes < 64) bytes = 64;
    {
        std::lock_guard<std::mutex> lk(g_hmu);
        int best = -1;
        for (int i = 0; i < (int)g_hfree.size(); i++)
            if (g_hfree[i]->cap >= bytes && g_hfree[i]->cap <= 4 * bytes + (1 << 20) && (best < 0 || g_hfree[i]->cap < g_hfree[best]->cap)) best = i;
        if (best >= 0) { HBlock *h = g_hfree[best]; g_hfree.erase(g_hfree.begin() + best); return (void *)(h + 1); }
    }
    size_t cap = bytes + (bytes >> 3);
    HBlock *h = nullptr; size_t pinned = 0;
    if (bytes >= (256 << 10)) {
        void *p = nullptr;
        if (cudaHostAlloc(&p, sizeof(HBlock) + cap, cudaHostAllocDefault) == cudaSuccess) { h = (HBlock *)p; pinned = 1; }
        else cudaGetLastError();
    }
    if (!h) h = (HBlock *)malloc(sizeof(HBlock) + cap);
    if (!h) throw std::bad_alloc();
    h->cap = cap; h->pinned = pinned;
    return (void *)(h + 1);
}
bool comm_owns_host_pointer(const void *p);           // comm.cu: results of the sharded multi-GPU download live in a shared segment
void hcache_free(void *p) {
    if (!p || comm_owns_host_pointer(p)) return;
    HBlock *h = (HBlock *)p - 1;
    std::lock_guard<std::mutex> lk(g_hmu);
    if (g_hfree.size() < 96) { g_hfree.push_back(h); return; }           // a pile-up batch cycles through ~40 result / staging blocks
    if (h->pinned) cudaFreeHost(h); else free(h);
}

// k > 15: the index entry packs into 8 bytes (kmer << pb | position) when the k-mer and the block's positions fit 64 bits
// together -- e.g. k = 20 against an assembly block of up to 16 M padded bases; larger blocks take 16-byte entries
static int packed_pos_bits(int64_t nA, int k) {
    if (k <= 15 || getenv("DN_NO_PACKED")) return 0;
    const int pb = bits_for((uint64_t)nA);
    return 2 * k + pb <= 64 ? pb : 0;
}

static int index_tbits(int64_t nA, int k, bool lookup) {
    static const int delta = getenv("DN_TBITS_DELTA") ? atoi(getenv("DN_TBITS_DELTA")) : 0;      // experiment: coarser / finer prefix table
    int tbits = bits_for((uint64_t)nA) + (lookup ? -1 + delta : 1);
    if (tbits < 16) tbits = 16;
    if (tbits > 2 * k) tbits = 2 * k;
    return tbits;
}

// dn_block_index: the A-side index of align_blocks (sorted k-mer tuples, prefix table, k-mer filter) built once and kept
// with the block -- a reference block is aligned against many read blocks (Snakefile:1143-1170 fan-out).
void block_build_index(DevBlock &A, int k, cudaStream_t s) {
    if (k < 4 || k > 31) throw Error("k must be in [4,31]");
    A.index.drop();
    const int64_t nA = A.total;
    if (nA == 0) return;
    const bool wide = k > 15;
    const int pb = packed_pos_bits(nA, k);
    DevBlock::Index &X = A.index;
    int64_t nI = nA;
    bool bucketed = false;                                   // index built by bucket counting (bucket.cu): the table comes with it
    auto table_geom = [&](int64_t n_index) {
        X.tbits = index_tbits(n_index, k, true);
        X.tbl.persistent(((size_t)1 << X.tbits) + 3);
    };
    if (!wide) {
        DBuf<u64> ta(nA), ta2;
        emit_tuples(A, false, k, 0u, ta.p, s);
        table_geom(nA);
        X.ta.persistent(nA);
        bucketed = build_index_u64(ta.p, X.ta.p, nA, 32, 2 * k - X.tbits, 1u << X.tbits, X.tbl.p, s);
        if (!bucketed) {
            ta2.alloc(nA);
            u64 *sa = radix_sort_u64(ta.p, ta2.p, nA, 32, 32 + 2 * k + 1, s);
            DN_CUDA(cudaMemcpyAsync(X.ta.p, sa, sizeof(u64) * nA, cudaMemcpyDeviceToDevice, s));
        }
    } else if (pb) {
        DBuf<u64> tp(nA), tp2;
        nI = emit_tuples_wide(A, k, tp.p, pb, s);
        table_geom(nI);
        X.ta.persistent(nI + 1);
        bucketed = build_index_u64(tp.p, X.ta.p, nI, pb, 2 * k - X.tbits, 1u << X.tbits, X.tbl.p, s);
        if (!bucketed) {
            tp2.alloc(nA);
            u64 *sp = radix_sort_u64(tp.p, tp2.p, nI, pb, pb + 2 * k, s);
            if (nI) DN_CUDA(cudaMemcpyAsync(X.ta.p, sp, sizeof(u64) * nI, cudaMemcpyDeviceToDevice, s));
        }
    } else {
        DBuf<ulonglong2> tw(nA), tw2(nA);
        nI = emit_tuples_wide(A, k, tw.p, 0, s);
        table_geom(nI);
        ulonglong2 *sw = radix_sort_rec16(tw.p, tw2.p, nI, 0, 0, 2 * k, s);
        X.tw.persistent(nI + 1);
        if (nI) DN_CUDA(cudaMemcpyAsync(X.tw.p, sw, sizeof(ulonglong2) * nI, cudaMemcpyDeviceToDevice, s));
    }
    X.n = nI; X.pb = pb;
    X.tbl_shift = bucketed ? 1 : 0;
    const int sh = 2 * k - X.tbits; const u32 nq = 1u << X.tbits;
    X.kbits_log2 = kbits_log2_for(nI);
    const int kshift = 32 - (X.kbits_log2 - 5);
    X.kbits.persistent((size_t)1 << (X.kbits_log2 - 5)); X.kbits.zero(s);
    if (!bucketed) X.tbl.zero(s);                            // an index without entries: every range is empty
    if (!wide) {
        if (!bucketed) DN_LAUNCH(k_prefix_table, (unsigned)((nI + 255) / 256), 256, 0, s, (const u64 *)X.ta.p, nI, sh, nq, X.tbl.p);
        DN_LAUNCH(k_kmer_bitmap, (unsigned)((nI + 255) / 256), 256, 0, s, (const u64 *)X.ta.p, nI, k, kshift, X.kbits.p);
    } else if (nI > 0 && pb) {
        if (!bucketed) DN_LAUNCH(k_prefix_table_p, (unsigned)((nI + 255) / 256), 256, 0, s, (const u64 *)X.ta.p, pb, nI, sh, nq, X.tbl.p);
        DN_LAUNCH(k_kmer_bitmap_p, (unsigned)((nI + 255) / 256), 256, 0, s, (const u64 *)X.ta.p, pb, nI, k, kshift, X.kbits.p);
    } else if (nI > 0) {
        DN_LAUNCH(k_prefix_table_w, (unsigned)((nI + 255) / 256), 256, 0, s, (const ulonglong2 *)X.tw.p, nI, sh, nq, X.tbl.p);
        DN_LAUNCH(k_kmer_bitmap_w, (unsigned)((nI + 255) / 256), 256, 0, s, (const ulonglong2 *)X.tw.p, nI, k, kshift, X.kbits.p);
    }
    DN_CUDA(cudaStreamSynchronize(s));
    X.k = k; X.valid = true;
}

void align_blocks(const DevBlock &A, const DevBlock &B, const AlignParams &P, HostLas &out, cudaStream_t s, DevLas *keep) {
    out = HostLas();
    if (keep) *keep = DevLas();
    arena().reset();
    if (P.k < 4 || P.k > 31) throw Error("k must be in [4,31]");
    if (P.wmax < 4 || P.wmax > 62) throw Error("wmax must be in [4,62]");
    if (P.w < 1 || P.w > 12 || P.tspace < 1 || P.tspace > 32767) throw Error("bad w / tspace");
    if (P.cdiff < 20 || P.xdrop < 1 || P.xdrop > 1000) throw Error("bad cdiff / xdrop");
    if (P.rounds < 1 || P.poolmul < 1) throw Error("bad rounds / poolmul");
    const unsigned long long launches0 = g_launches.load();
    Timer tt(s);
    tt.start();
    if (A.nreads == 0 || B.nreads == 0) {
        out.rec = (dn_las_record *)hcache_alloc(64); out.toff = (int64_t *)hcache_alloc(64); out.trace = (uint16_t *)hcache_alloc(64);
        out.stats.ms_total = tt.stop(); return;
    }
    if (A.ready) DN_CUDA(cudaStreamWaitEvent(s, A.ready, 0));
    const int k = P.k;
    const int64_t nA = A.total, nB = B.total;
    Trace tr(s);

    // ---- K1 + K2: tuples, radix sort by k-mer -------------------------------------------------
    // k <= 15: 8-byte tuples (kmer << 32 | position); k = 16..31: 16-byte {kmer, position} tuples, index lookup join only
    const bool wide = k > 15;
    if (wide && P.join_mode == 1) throw Error("the sorted-merge join (join_mode 1) supports k <= 15 only");
    DBuf<u64> ta, ta2; DBuf<ulonglong2> tw, tw2;
    u64 *sa = nullptr; ulonglong2 *sw = nullptr;
    const DevBlock::Index *cached = (A.index.valid && A.index.k == k && P.join_mode != 1) ? &A.index : nullptr;
    int64_t nI = nA;                                       // index entries: k > 15 keeps the valid positions only
    int pb = cached ? cached->pb : packed_pos_bits(nA, k);  // > 0: 8-byte packed entries kmer << pb | position (in `sa`)
    const bool lookup = cached || wide || P.join_mode == 2 || (P.join_mode == 0 && nA * 8 <= (2ll << 30));   // auto: index lookup unless the A index is huge
    DBuf<u32> tbl_own; const u32 *tblp = nullptr;
    bool bucketed = false;                                 // index built by bucket counting (bucket.cu): the prefix table comes with it
    int tbits = 0;
    if (cached) { sa = cached->ta.p; sw = cached->tw.p; nI = cached->n; tbits = cached->tbits; tblp = cached->tbl.p + cached->tbl_shift; }
    else if (!wide) {
        ta.alloc(nA);
        emit_tuples(A, false, k, 0u, ta.p, s);
        tbits = index_tbits(nA, k, lookup);
        tbl_own.alloc(((size_t)1 << tbits) + 3);
        if (lookup) { ta2.alloc(nA); bucketed = build_index_u64(ta.p, ta2.p, nA, 32, 2 * k - tbits, 1u << tbits, tbl_own.p, s); }
        if (bucketed) { sa = ta2.p; ta.release(); }
        else {
            if (!ta2.p) ta2.alloc(nA);
            sa = radix_sort_u64(ta.p, ta2.p, nA, 32, 32 + 2 * k + 1, s);
            if (sa == ta.p) ta2.release(); else ta.release();
        }
    } else if (pb) {
        ta.alloc(nA); ta2.alloc(nA);
        nI = emit_tuples_wide(A, k, ta.p, pb, s);
        tbits = index_tbits(nI, k, lookup);
        tbl_own.alloc(((size_t)1 << tbits) + 3);
        bucketed = build_index_u64(ta.p, ta2.p, nI, pb, 2 * k - tbits, 1u << tbits, tbl_own.p, s);
        if (bucketed) { sa = ta2.p; ta.release(); }
        else {
            sa = radix_sort_u64(ta.p, ta2.p, nI, pb, pb + 2 * k, s);
            if (sa == ta.p) ta2.release(); else ta.release();
        }
    } else {
        tw.alloc(nA); tw2.alloc(nA);
        nI = emit_tuples_wide(A, k, tw.p, 0, s);
        tbits = index_tbits(nI, k, lookup);
        tbl_own.alloc(((size_t)1 << tbits) + 3);
        sw = radix_sort_rec16(tw.p, tw2.p, nI, 0, 0, 2 * k, s);
        if (sw == tw.p) tw2.release(); else tw.release();
    }
    const bool wide16 = wide && !pb;
    tr.mark("A tuples + sort");
    const int npass_t = (2 * k + (wide ? 0 : 1) + 7) / 8;
    out.stats.tuples_a = A.total_real; out.stats.tuples_b = 2 * B.total_real;
    int64_t abytes = nA / 4 + (wide16 ? 16 : 8) * nA;                                                   // A: read packed, write tuples
    if (bucketed) abytes += 24 * nI + 12ll * ((int64_t)1 << tbits);                                       // count (1R), scatter (1R + 1W), table (zero, scan R + W)
    else if (!cached) abytes += (int64_t)npass_t * (wide16 ? 32 : 16) * nI + (wide16 ? 16 : 8) * nI;      // one-sweep radix: histograms (1R) + passes (1R + 1W)

    // hit-key geometry
    const int64_t bandw = 1ll << P.w;
    const int64_t maxlb = ((int64_t)B.maxlen + bandw - 1) / bandw * bandw;
    std::vector<int64_t> dbase(A.nreads + 1);
    { int64_t d = 0;
      for (int r = 0; r < A.nreads; r++) { dbase[r] = d; d = (d + A.h_len[r] + maxlb + bandw - 1) / bandw * bandw + bandw; }
      dbase[A.nreads] = d; }
    const int gdbits = bits_for((uint64_t)dbase[A.nreads] + 2 * bandw);
    const int keybits = gdbits + bits_for((uint64_t)2 * B.nreads);
    if (keybits > 62) throw Error("hit key does not fit 62 bits");
    DBuf<int64_t> d_dbase(A.nreads + 1);
    DN_CUDA(cudaMemcpyAsync(d_dbase.p, dbase.data(), sizeof(int64_t) * (A.nreads + 1), cudaMemcpyHostToDevice, s));
    const bool grouped = A.has_group && B.has_group;
    JoinGeom JG{A.chunk2read.p, A.off.p, d_dbase.p, B.chunk2read.p, B.off.p, nB, maxlb, gdbits, keybits, P.self,
                grouped ? A.group.p : nullptr, grouped ? B.group.p : nullptr, B.nreads};
    SeedGeom SG{d_dbase.p, A.nreads, gdbits};

    const int aposbits = bits_for((uint64_t)A.maxlen);
    bool segsorted = false, seg_in_hits2 = false;
    // B may still be uploading on the copy stream (dn_align_host): a chunked upload lets the count pass start behind the
    // metadata and follow the chunks; otherwise wait for the whole block
    if (!B.chunk_ready.empty() && lookup) { if (B.meta_ready) DN_CUDA(cudaStreamWaitEvent(s, B.meta_ready, 0)); }
    else if (B.ready) DN_CUDA(cudaStreamWaitEvent(s, B.ready, 0));
    // ---- K3: join ------------------------------------------------------------------------------
    const int sh = 2 * k - tbits; const u32 nq = 1u << tbits;
    if (!cached) {
        tblp = tbl_own.p + (bucketed ? 1 : 0);
        if (!bucketed) {
            if (!wide) DN_LAUNCH(k_prefix_table, (unsigned)((nA + 255) / 256), 256, 0, s, (const u64 *)sa, nA, sh, nq, tbl_own.p);
            else if (nI > 0 && pb) DN_LAUNCH(k_prefix_table_p, (unsigned)((nI + 255) / 256), 256, 0, s, (const u64 *)sa, pb, nI, sh, nq, tbl_own.p);
            else if (nI > 0) DN_LAUNCH(k_prefix_table_w, (unsigned)((nI + 255) / 256), 256, 0, s, (const ulonglong2 *)sw, nI, sh, nq, tbl_own.p);
            else tbl_own.zero(s);                                // no valid k-mer in A: every range is empty
        }
    }
    struct { const u32 *p; } tbl{tblp};
    DBuf<int64_t> dtotal(1);
    DBuf<unsigned long long> ninv(1); ninv.zero(s);
    DBuf<ulonglong2> hits, hits2;
    int64_t H = 0;
    abytes += 8 * nA + 4ll * nq;
    if (lookup) {
        // B's tuples are never materialised: hits come straight from the packed sequence
        const int64_t nwB = nB >> 4;
        const int nseg = 2 * B.nreads;
        DBuf<u32> kbits_own; struct { const u32 *p; } kbits{cached ? cached->kbits.p : nullptr};
        const int kblog = cached ? cached->kbits_log2 : kbits_log2_for(nI), kshift = 32 - (kblog - 5);
        if (!cached) {
            kbits_own.alloc((size_t)1 << (kblog - 5)); kbits_own.zero(s); kbits.p = kbits_own.p;
            if (!wide) DN_LAUNCH(k_kmer_bitmap, (unsigned)((nA + 255) / 256), 256, 0, s, (const u64 *)sa, nA, k, kshift, kbits_own.p);
            else if (nI > 0 && pb) DN_LAUNCH(k_kmer_bitmap_p, (unsigned)((nI + 255) / 256), 256, 0, s, (const u64 *)sa, pb, nI, k, kshift, kbits_own.p);
            else if (nI > 0) DN_LAUNCH(k_kmer_bitmap_w, (unsigned)((nI + 255) / 256), 256, 0, s, (const ulonglong2 *)sw, nI, k, kshift, kbits_own.p);
        }
        // pin the k-mer filter in the persisting part of L2 while the streaming lookups run
        {
            const size_t fbytes = kblog <= 28 ? ((size_t)1 << (kblog - 3)) : 0;       // only an L2-sized filter is pinned
            const size_t fset = l2_persisting_bytes(fbytes);
            cudaStreamAttrValue av; memset(&av, 0, sizeof av);
            av.accessPolicyWindow.base_ptr = (void *)kbits.p; av.accessPolicyWindow.num_bytes = fset >= fbytes ? fbytes : 0;
            av.accessPolicyWindow.hitRatio = 1.0f; av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &av); cudaGetLastError();
        }
        DBuf<int64_t> seg_beg((size_t)nseg + 1); DBuf<int32_t> seg_len((size_t)nseg + 1);
        // segments by size class, dealt out on the device right after the scan (ccnt[3] = segments beyond every class)
        const int scaps[3] = {2048, 8192, 16384};
        const bool seg_ok = gdbits + aposbits <= 63 && !getenv("DN_NO_SEGSORT");
        const bool radix = gdbits <= 32 && !getenv("DN_BITONIC");     // stable radix by gd: hits -> hits2
        DBuf<int32_t> clists((size_t)4 * nseg + 1); DBuf<u32> ccnt(4); ccnt.zero(s);
        u32 hcc[4] = {0, 0, 0, 0};
        // (a fused one-CTA-per-read count+reserve+emit kernel was measured at 2x the time of these two passes:
        //  the lookups are latency-bound and want full occupancy, which the per-read shared-memory state prevents)
        {
            DBuf<u32> wcnt(2 * nwB), wlist(2 * nwB), nlist(2); DBuf<int64_t> woff(2 * nwB);
            nlist.zero(s); DBuf<unsigned short> hitmask(2 * nwB);
            {   // one sweep over the forward words counts both strands; the complement strand's words collect by atomics
                DN_CUDA(cudaMemsetAsync(wcnt.p + nwB, 0, sizeof(u32) * nwB, s));
                DN_CUDA(cudaMemsetAsync(hitmask.p + nwB, 0, sizeof(unsigned short) * nwB, s));
                const u32 *mb = B.has_mask ? B.mask.p : nullptr;
                // a block that is still arriving (dn_align_host: chunked upload on the copy stream) is swept chunk by chunk, each
                // sweep behind its chunk's event; a resident block is one chunk
                const size_t nchunks = B.chunk_ready.empty() ? 1 : B.chunk_ready.size();
                for (size_t c = 0; c < nchunks; c++) {
                    int64_t w0 = 0, w1 = nwB;
                    if (!B.chunk_ready.empty()) {
                        DN_CUDA(cudaStreamWaitEvent(s, B.chunk_ready[c], 0));
                        w0 = B.chunk_word[c]; w1 = B.chunk_word[c + 1];
                    }
                    if (w1 <= w0) continue;
                    const unsigned grid = (unsigned)((w1 - w0 + 255) / 256);
                    if (!wide)
                        DN_LAUNCH(k_lookup_count, grid, 256, 0, s, (const u32 *)B.fwd.p, mb,
                                  (const int64_t *)B.off.p, (const int32_t *)B.len.p, (const int32_t *)B.chunk2read.p, nwB, k,
                                  (const u64 *)sa, (const u32 *)tbl.p, sh, P.t, (const u32 *)kbits.p, kshift, JG, wcnt.p, hitmask.p, wlist.p, nlist.p, w0, w1);
                    else if (pb)
                        DN_LAUNCH(k_lookup_count_p, grid, 256, 0, s, (const u32 *)B.fwd.p, mb,
                                  (const int64_t *)B.off.p, (const int32_t *)B.len.p, (const int32_t *)B.chunk2read.p, nwB, k,
                                  (const u64 *)sa, pb, (const u32 *)tbl.p, sh, P.t, (const u32 *)kbits.p, kshift, JG, wcnt.p, hitmask.p, wlist.p, nlist.p, w0, w1);
                    else
                        DN_LAUNCH(k_lookup_count_w, grid, 256, 0, s, (const u32 *)B.fwd.p, mb,
                                  (const int64_t *)B.off.p, (const int32_t *)B.len.p, (const int32_t *)B.chunk2read.p, nwB, k,
                                  (const ulonglong2 *)sw, (const u32 *)tbl.p, sh, P.t, (const u32 *)kbits.p, kshift, JG, wcnt.p, hitmask.p, wlist.p, nlist.p, w0, w1);
                }
                if (B.ready) DN_CUDA(cudaStreamWaitEvent(s, B.ready, 0));       // everything of B from here on
            }
            exclusive_scan_u32_to_i64(wcnt.p, woff.p, 2 * nwB, dtotal.p, s);
            // radix variant writes to the other buffer: singletons are copied too (minimum length 1)
            launch_seg_offsets(woff.p, B.off.p, B.nreads, nwB, dtotal.p, seg_beg.p, seg_len.p, radix ? 1 : 2, scaps, clists.p, ccnt.p, s);
            u32 nl[2] = {0, 0};
            DN_CUDA(cudaMemcpyAsync(&H, dtotal.p, 8, cudaMemcpyDeviceToHost, s));            // ONE drain: hit total, word lists, class counts
            DN_CUDA(cudaMemcpyAsync(nl, nlist.p, 8, cudaMemcpyDeviceToHost, s));
            DN_CUDA(cudaMemcpyAsync(hcc, ccnt.p, 16, cudaMemcpyDeviceToHost, s));
            DN_CUDA(cudaStreamSynchronize(s));
            if (H >= (1ll << 31) - 4096) throw Error("too many seed hits for one block pair (>= 2^31); use smaller blocks or lower -t");
            hits.alloc((size_t)H + 1); hits2.alloc((size_t)H + 1);
            if (H > 0)
                for (int st = 0; st < 2; st++) {
                    const u32 *mb = B.has_mask ? (st ? B.mask_rc.p : B.mask.p) : nullptr;
                    if (nl[st] == 0) continue;
                    if (!wide)
                        DN_LAUNCH(k_lookup_emit, (unsigned)((nl[st] + 255) / 256), 256, 0, s, (const u32 *)(st ? B.rc.p : B.fwd.p), mb,
                                  (const int64_t *)B.off.p, (const int32_t *)B.len.p, (const int32_t *)B.chunk2read.p, (int64_t)nl[st], k,
                                  (const u64 *)sa, (const u32 *)tbl.p, sh, P.t, (const unsigned short *)(hitmask.p + st * nwB), (const u32 *)(wcnt.p + st * nwB),
                                  (const int64_t *)(woff.p + st * nwB), st, JG, hits.p, (const u32 *)(wlist.p + st * nwB));
                    else if (pb)
                        DN_LAUNCH(k_lookup_emit_p, (unsigned)((nl[st] + 255) / 256), 256, 0, s, (const u32 *)(st ? B.rc.p : B.fwd.p), mb,
                                  (const int64_t *)B.off.p, (const int32_t *)B.len.p, (const int32_t *)B.chunk2read.p, (int64_t)nl[st], k,
                                  (const u64 *)sa, pb, (const u32 *)tbl.p, sh, P.t, (const unsigned short *)(hitmask.p + st * nwB), (const u32 *)(wcnt.p + st * nwB),
                                  (const int64_t *)(woff.p + st * nwB), st, JG, hits.p, (const u32 *)(wlist.p + st * nwB));
                    else
                        DN_LAUNCH(k_lookup_emit_w, (unsigned)((nl[st] + 255) / 256), 256, 0, s, (const u32 *)(st ? B.rc.p : B.fwd.p), mb,
                                  (const int64_t *)B.off.p, (const int32_t *)B.len.p, (const int32_t *)B.chunk2read.p, (int64_t)nl[st], k,
                                  (const ulonglong2 *)sw, (const u32 *)tbl.p, sh, P.t, (const unsigned short *)(hitmask.p + st * nwB), (const u32 *)(wcnt.p + st * nwB),
                                  (const int64_t *)(woff.p + st * nwB), st, JG, hits.p, (const u32 *)(wlist.p + st * nwB));
                }
            abytes += 2 * (nB / 4) * 2 + 2 * nwB * (4 + 4 + 8 + 8 + 4) + 16 * H;   // packed B read twice per strand, word counts/offsets, hits
        }
        if (H >= (1ll << 31) - 4096) throw Error("too many seed hits for one block pair (>= 2^31); use smaller blocks or lower -t");
        {
            cudaStreamAttrValue av; memset(&av, 0, sizeof av);            // release the L2 set-aside
            cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &av); cudaGetLastError();
            cudaCtxResetPersistingL2Cache(); cudaGetLastError();
        }
        // every (strand, read) segment is contiguous and, per diagonal, already in apos order: sort inside the segments only
        if (H > 0 && seg_ok && hcc[3] <= 256) {
            if (getenv("DN_TRACE"))
                fprintf(stderr, "[dn trace] two-pass join, segments %d, classes %u/%u/%u, oversized %u\n", nseg, hcc[0], hcc[1], hcc[2], hcc[3]);
            for (int c = 0; c < 3; c++) {
                if (hcc[c] == 0) continue;
                const int32_t *lst = clists.p + (size_t)c * nseg;
                if (radix) launch_segsort_radix(hits.p, hits2.p, seg_beg.p, seg_len.p, lst, (int)hcc[c], scaps[c], gdbits, s);
                else launch_segsort(hits.p, seg_beg.p, seg_len.p, lst, (int)hcc[c], scaps[c], gdbits, aposbits, s);
            }
            ulonglong2 *sorted = radix ? hits2.p : hits.p;
            if (hcc[3] > 0) {
                // the few oversized segments: gather, radix sort by (bs, gd, apos), copy back segment by segment (ascending bs)
                std::vector<int32_t> big(hcc[3]); std::vector<int64_t> hbeg(nseg); std::vector<int32_t> hlen(nseg);
                DN_CUDA(cudaMemcpyAsync(big.data(), clists.p + (size_t)3 * nseg, sizeof(int32_t) * hcc[3], cudaMemcpyDeviceToHost, s));
                DN_CUDA(cudaMemcpyAsync(hbeg.data(), seg_beg.p, sizeof(int64_t) * nseg, cudaMemcpyDeviceToHost, s));
                DN_CUDA(cudaMemcpyAsync(hlen.data(), seg_len.p, sizeof(int32_t) * nseg, cudaMemcpyDeviceToHost, s));
                DN_CUDA(cudaStreamSynchronize(s));
                std::sort(big.begin(), big.end());
                int64_t nbig = 0; for (int i : big) nbig += hlen[i];
                DBuf<ulonglong2> t1(nbig), t2(nbig);
                int64_t o = 0;
                for (int i : big) { DN_CUDA(cudaMemcpyAsync(t1.p + o, hits.p + hbeg[i], 16 * (int64_t)hlen[i], cudaMemcpyDeviceToDevice, s)); o += hlen[i]; }
                ulonglong2 *r = radix_sort_rec16(t1.p, t2.p, nbig, 1, 0, aposbits, s);
                r = radix_sort_rec16(r, r == t1.p ? t2.p : t1.p, nbig, 0, 0, keybits, s);
                o = 0;
                for (int i : big) { DN_CUDA(cudaMemcpyAsync(sorted + hbeg[i], r + o, 16 * (int64_t)hlen[i], cudaMemcpyDeviceToDevice, s)); o += hlen[i]; }
            }
            segsorted = true; seg_in_hits2 = radix;
            abytes += 32 * H + 12ll * nseg;
        }
    } else {
        DBuf<u64> tb(2 * nB), tb2(2 * nB);
        emit_tuples(B, false, k, 0u, tb.p, s);
        emit_tuples(B, true, k, (u32)nB, tb.p + nB, s);
        u64 *sb = radix_sort_u64(tb.p, tb2.p, 2 * nB, 32, 32 + 2 * k + 1, s);
        DBuf<u32> cnt(2 * nB), start(2 * nB);
        DN_LAUNCH(k_join_count, (unsigned)((2 * nB + 255) / 256), 256, 0, s, (const u64 *)sa, (const u32 *)tbl.p, sh,
                  (const u64 *)sb, 2 * nB, P.t, cnt.p, start.p);
        DBuf<int64_t> hoff(2 * nB);
        exclusive_scan_u32_to_i64(cnt.p, hoff.p, 2 * nB, dtotal.p, s);
        H = d2h_scalar(dtotal.p, s);
        if (H >= (1ll << 31) - 4096) throw Error("too many seed hits for one block pair (>= 2^31); use smaller blocks or lower -t");
        hits.alloc((size_t)H + 1); hits2.alloc((size_t)H + 1);
        if (H > 0)
            DN_LAUNCH(k_join_emit, (unsigned)((2 * nB + 255) / 256), 256, 0, s, (const u64 *)sa, (const u64 *)sb, 2 * nB,
                      (const u32 *)cnt.p, (const u32 *)start.p, (const int64_t *)hoff.p, JG, hits.p, ninv.p);
        abytes += 2 * (nB / 4) + 8 * 2 * nB + (int64_t)npass_t * 24 * 2 * nB + 2 * 8 * 2 * nB + 16 * 2 * nB + 16 * H;
    }
    const int64_t ninvalid = lookup ? 0 : (int64_t)d2h_scalar(ninv.p, s);      // the lookup join never emits a self / cross-pile pair
    tr.mark("join");
    ta.release(); ta2.release(); tw.release(); tw2.release(); tbl_own.release();

    // ---- hit sort: by apos, then stably by (bread, strand, aread, diagonal) -------------------
    ulonglong2 *hs = seg_in_hits2 ? hits2.p : hits.p, *ho = seg_in_hits2 ? hits.p : hits2.p;
    if (!segsorted) {
        hs = radix_sort_rec16(hits.p, hits2.p, H, 1, 0, aposbits, s);
        ho = hs == hits.p ? hits2.p : hits.p;
        hs = radix_sort_rec16(hs, ho, H, 0, 0, keybits + 1, s);
        ho = hs == hits.p ? hits2.p : hits.p;
        const int npass_h = (aposbits + 7) / 8 + (keybits + 1 + 7) / 8;
        abytes += (int64_t)npass_h * 48 * H;
    }
    int64_t n = H - ninvalid;                   // invalid (self) hits sorted to the end
    tr.mark("hit sort");
    out.stats.hits = n;

    // ---- rounds of band filter -> seeds -> extension -> retirement ----------------------------
    ExtGeom EG{A.fwd.p, A.rc.p, B.fwd.p, B.rc.p, A.off.p, B.off.p, A.len.p, B.len.p,
               P.tspace, P.cdiff, P.xdrop, P.wmax, P.poolmul, (u32)((1ull << 32) / (u32)P.tspace + 1), B.nreads,
               (u32)A.fwd.n, (u32)B.fwd.n, 0};
    {
        long long span = (long long)B.maxlen + B.maxlen / 2 + 64; if (A.maxlen < span) span = A.maxlen;
        if ((unsigned long long)span >= (1ull << 32) / (unsigned)P.tspace) throw Error("reads too long for the tile arithmetic");
    }
    const int64_t pool_stride = (int64_t)P.poolmul * ([&] { long long sp = (long long)B.maxlen + B.maxlen / 2 + 64;
                                                           if (A.maxlen < sp) sp = A.maxlen; return sp; }() / P.tspace + 4);
    std::vector<DBuf<Cand>> round_cands; std::vector<DBuf<uint16_t>> round_traces;
    std::vector<int32_t> round_beg{0};
    std::vector<int64_t> round_ntr;
    float ms_ext = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ext_ev;      // brackets of the k_extend launches, read after the last sync
    int64_t ext_bytes = 0;
    DBuf<int32_t> dtot32(2);          // 8-byte aligned (arena allocations are 512-byte aligned)

    for (int round = 0; round < P.rounds && n > 0; round++) {
        DBuf<int32_t> bflag(n), bidx(n), covsum(n);
        // the packed totals land as {total cover, bands} in dtot32[0..1] (little endian: low word first)
        if (n < (1ll << 30)) launch_cover_scan((const ulonglong2 *)hs, n, k, P.w, bflag.p, bidx.p, covsum.p, (unsigned long long *)dtot32.p, s);
        else {
            DBuf<int32_t> cov(n);
            DN_LAUNCH(k_hit_cover, (unsigned)((n + 255) / 256), 256, 0, s, (const ulonglong2 *)hs, n, k, P.w, cov.p, bflag.p);
            exclusive_scan_i32(bflag.p, bidx.p, n, dtot32.p + 1, s);
            exclusive_scan_i32(cov.p, covsum.p, n, dtot32.p, s);
        }
        int32_t two[2];
        DN_CUDA(cudaMemcpyAsync(two, dtot32.p, 8, cudaMemcpyDeviceToHost, s)); DN_CUDA(cudaStreamSynchronize(s));
        const int32_t nbands = two[1];
        DBuf<int32_t> bfirst((size_t)nbands + 2), cstart(nbands), cidx(nbands);
        DBuf<u64> bkey((size_t)nbands + 1);
        DBuf<uint8_t> hot(nbands);
        DN_LAUNCH(k_band_table, (unsigned)((n + 255) / 256), 256, 0, s, (const ulonglong2 *)hs, n, P.w, (const int32_t *)bflag.p,
                  (const int32_t *)bidx.p, bfirst.p, bkey.p, nbands);
        DN_LAUNCH(k_band_hot, (nbands + 255) / 256, 256, 0, s, (const int32_t *)bfirst.p, (const u64 *)bkey.p,
                  (const int32_t *)covsum.p, (const int32_t *)dtot32.p, n, nbands, P.h, hot.p, cstart.p);
        DBuf<int32_t> dseeds(1);
        exclusive_scan_i32(cstart.p, cidx.p, nbands, dseeds.p, s);
        const int32_t nseeds = d2h_scalar(dseeds.p, s);
        abytes += 16 * n + 8 * n + 3 * 12 * n + 16ll * nbands;
        tr.mark("band filter");
        if (nseeds == 0) break;
        DBuf<Seed> seeds(nseeds); DBuf<uint8_t> consumed(n); consumed.zero(s);
        DN_LAUNCH(k_seeds, (nbands + 255) / 256, 256, 0, s, (const ulonglong2 *)hs, (const int32_t *)bfirst.p, (const u64 *)bkey.p,
                  (const uint8_t *)hot.p, (const int32_t *)cstart.p, (const int32_t *)cidx.p, nbands, SG, seeds.p, consumed.p);
        out.stats.seeds += nseeds; out.stats.extensions += 2ll * nseeds;

        // ---- K5: extension
        DBuf<int64_t> tile_off;
        int64_t ntile_cap;
        {   // one tile capacity for every task when that stays affordable: no per-task caps, no scan, no host round trip
            long long sp = (long long)B.maxlen + B.maxlen / 2 + 64; if (A.maxlen < sp) sp = A.maxlen;
            const int64_t capmax = sp / P.tspace + 3;
            if (2ll * nseeds * capmax * (int64_t)sizeof(int2) <= (1ll << 30)) {
                EG.tile_stride = capmax;                              // task t owns tiles [t * capmax, (t + 1) * capmax)
                ntile_cap = 2ll * nseeds * capmax;
            } else {
                EG.tile_stride = 0;
                tile_off.alloc(2 * (size_t)nseeds);
                DBuf<u32> caps(2 * (size_t)nseeds);
                launch_task_caps(seeds.p, nseeds, EG, caps.p, s);
                exclusive_scan_u32_to_i64(caps.p, tile_off.p, 2 * (size_t)nseeds, dtotal.p, s);
                ntile_cap = d2h_scalar(dtotal.p, s);
            }
        }
        DBuf<int2> tiles((size_t)ntile_cap + 1); DBuf<ExtOut> outs(2 * (size_t)nseeds);
        const int wpc = ext_warps_per_cta();
        int ctas = sm_count() * ext_ctas_per_sm();    // 48 registers: 4 CTAs x 9 warps per SM (measured optimum; DN_EXT_CTAS overrides)
        { int64_t need = (2ll * nseeds + wpc - 1) / wpc; if (ctas > need) ctas = (int)need;
          const int64_t budget = 24ll << 30;   // bytes of HBM for trace-record pools
          int64_t maxc = budget / (pool_stride * 16 * wpc); if (maxc < 1) maxc = 1;
          if (ctas > maxc) ctas = (int)maxc; }
        DBuf<int4> pool((size_t)ctas * wpc * pool_stride);
        DBuf<int> counter(1); counter.zero(s);
        DBuf<int> order; 
        if (!getenv("DN_EXT_NO_ORDER")) {              // tasks handed out longest-expected-first (no effect on any result)
            DBuf<int> oscr(128); order.alloc(2 * (size_t)nseeds);
            launch_task_order(seeds.p, nseeds, EG, oscr.p, order.p, s);
        }
        tr.mark("seeds + ext setup");
        {   // events bracket exactly the k_extend launch; no host sync here
            cudaEvent_t ea, eb; cudaEventCreate(&ea); cudaEventCreate(&eb);
            cudaEventRecord(ea, s);
            launch_extend(seeds.p, nseeds, EG, tile_off.p, tiles.p, outs.p, pool.p, pool_stride, ctas * wpc, counter.p, order.p, s);
            cudaEventRecord(eb, s);
            ext_ev.emplace_back(ea, eb);
        }
        tr.mark("k_extend");

        // ---- K6: candidates, traces, retirement.  The kept alignments are counted on the device; their buffers are sized by
        // upper bounds (every seed kept; every tile slot used), so combine -> traces -> retire -> compaction offsets run without
        // the host, which reads the three totals in ONE drain at the end of the round.
        DBuf<Cand> cand_all(nseeds); DBuf<int32_t> valid(nseeds), vidx(nseeds); DBuf<u32> ntl(nseeds); DBuf<int64_t> toff(nseeds);
        launch_combine(seeds.p, nseeds, EG, P.minlen, tile_off.p, tiles.p, outs.p, cand_all.p, valid.p, ntl.p, s);
        exclusive_scan_i32(valid.p, vidx.p, nseeds, dtot32.p, s);
        exclusive_scan_u32_to_i64(ntl.p, toff.p, nseeds, dtotal.p, s);
        DBuf<Cand> rc((size_t)nseeds + 1); DBuf<uint16_t> rtr((size_t)2 * ntile_cap + 2);
        launch_write_traces(seeds.p, nseeds, EG, tile_off.p, tiles.p, outs.p, cand_all.p, valid.p, vidx.p, toff.p, rc.p, rtr.p, s);
        DBuf<int32_t> keep(n), kidx(n); DBuf<int2> crange((size_t)nbands + 1);
        launch_retire((const ulonglong2 *)hs, n, consumed.p, rc.p, dtot32.p, SG, P.w, bflag.p, bidx.p, hot.p, bfirst.p, nbands, crange.p, keep.p, s);
        exclusive_scan_i32(keep.p, kidx.p, n, dtot32.p + 1, s);
        int32_t two2[2]; int64_t ntr = 0;
        DN_CUDA(cudaMemcpyAsync(two2, dtot32.p, 8, cudaMemcpyDeviceToHost, s));
        DN_CUDA(cudaMemcpyAsync(&ntr, dtotal.p, 8, cudaMemcpyDeviceToHost, s));
        DN_CUDA(cudaStreamSynchronize(s));
        const int32_t nvalid = two2[0]; const int64_t n_new = two2[1];
        if (nvalid == 0) {                                            // no group kept anything: every group is finished (spec item 7)
            if (getenv("DN_TRACE")) fprintf(stderr, "[dn trace] round %d: %lld hits, %d bands, %d seeds, no candidate >= minlen: stop\n", round, (long long)n, nbands, nseeds);
            break;
        }
        launch_compact_hits((const ulonglong2 *)hs, n, keep.p, kidx.p, ho, s);      // same stream: the next round is ordered behind it
        std::swap(hs, ho);
        abytes += 24 * n + 16 * n_new;
        if (getenv("DN_TRACE")) fprintf(stderr, "[dn trace] round %d: %lld hits, %d bands, %d seeds, %d candidates >= minlen, %lld hits left\n", round, (long long)(n), nbands, nseeds, nvalid, (long long)n_new);
        n = n_new;
        round_beg.push_back(round_beg.back() + nvalid);
        round_ntr.push_back(2 * ntr);
        round_cands.push_back(std::move(rc)); round_traces.push_back(std::move(rtr));
        tr.mark("combine + retire");
    }

    // ---- duplicate removal, LAsort ordering and trace gather on the device; one download ---------
    const int ncand = round_beg.back();
    const int nrounds = (int)round_cands.size();
    if (nrounds > 16) throw Error("more than 16 rounds");
    int64_t aligned = 0, tot = 0;
    if (ncand > 0) {
        DBuf<Cand> all(ncand); DBuf<int32_t> rb(nrounds + 1); DBuf<uint8_t> drop(ncand);
        for (int r = 0; r < nrounds; r++)
            if (round_beg[r + 1] > round_beg[r])
                DN_CUDA(cudaMemcpyAsync(all.p + round_beg[r], round_cands[r].p, sizeof(Cand) * (round_beg[r + 1] - round_beg[r]),
                                        cudaMemcpyDeviceToDevice, s));
        DN_CUDA(cudaMemcpyAsync(rb.p, round_beg.data(), sizeof(int32_t) * (nrounds + 1), cudaMemcpyHostToDevice, s));
        launch_dedupe(all.p, ncand, rb.p, nrounds, drop.p, s);
        // LAsort order (base.d:1787-1809): (aread, bread, comp, abpos, aepos, bbpos, bepos, diffs), candidate index last
        FinalBits fb{bits_for((uint64_t)A.maxlen), bits_for((uint64_t)B.maxlen), bits_for((uint64_t)A.nreads), bits_for((uint64_t)B.nreads), B.nreads};
        DBuf<ulonglong2> it1(ncand), it2(ncand); DBuf<unsigned long long> ctr(3); ctr.zero(s);
        ulonglong2 *cur = it1.p, *oth = it2.p;
        const int pbits = 1 + fb.nra + fb.nrb + 1 + fb.na;            // (dropped, aread, bread, comp, abpos) in one key
        if (pbits <= 64 && !getenv("DN_FINAL_4SORTS")) {
            // one stable sort on the packed leading fields + a fix-up of the rare ties on the remaining ones: 6 passes instead of 16
            launch_final_setkey(all.p, drop.p, cur, ncand, 4, fb, ctr.p, s);
            ulonglong2 *res = radix_sort_rec16(cur, oth, ncand, 0, 0, pbits, s);
            if (res != cur) { oth = cur; cur = res; }
            launch_final_fixup(all.p, cur, ncand, s);
        } else {
            const int fbits[4] = {bits_for((uint64_t)A.maxlen + B.maxlen), 2 * fb.nb, 2 * fb.na + 1, fb.nra + fb.nrb + 1};
            for (int f = 0; f < 4; f++) {
                launch_final_setkey(all.p, drop.p, cur, ncand, f, fb, ctr.p, s);
                ulonglong2 *res = radix_sort_rec16(cur, oth, ncand, 0, 0, fbits[f], s);
                if (res != cur) { oth = cur; cur = res; }
            }
        }
        // records and traces are laid out for all ncand items (the dropped duplicates sort to the end and get no record);
        // the buffers are sized by the bounds the host already knows, so the counters travel with the result: ONE drain
        int64_t trbound = 0; for (int64_t v : round_ntr) trbound += v;
        DBuf<dn_las_record> drec(ncand); DBuf<u32> tl(ncand); DBuf<int64_t> dtoff(ncand);
        launch_final_records(all.p, cur, ncand, B.nreads, drec.p, tl.p, ctr.p, s);
        exclusive_scan_u32_to_i64(tl.p, dtoff.p, ncand, dtotal.p, s);
        FinalGeom FG; memset(&FG, 0, sizeof FG); FG.nrounds = nrounds;
        for (int r = 0; r < nrounds; r++) { FG.round_trace[r] = round_traces[r].p; FG.round_beg[r] = round_beg[r]; }
        FG.round_beg[nrounds] = round_beg[nrounds];
        DBuf<uint16_t> dtr((size_t)trbound + 1);
        launch_final_traces(all.p, cur, ncand, ctr.p, dtoff.p, FG, dtr.p, s);
        tr.mark("final records + traces");
        unsigned long long hctr[3];
        if (!keep) {
            out.rec = (dn_las_record *)hcache_alloc(sizeof(dn_las_record) * ((size_t)ncand + 1));
            out.toff = (int64_t *)hcache_alloc(sizeof(int64_t) * ((size_t)ncand + 1));
            out.trace = (uint16_t *)hcache_alloc(sizeof(uint16_t) * ((size_t)trbound + 1));
            tr.mark("host buffer alloc");
            DN_CUDA(cudaMemcpyAsync(out.rec, drec.p, sizeof(dn_las_record) * ncand, cudaMemcpyDeviceToHost, s));
            DN_CUDA(cudaMemcpyAsync(out.toff, dtoff.p, sizeof(int64_t) * ncand, cudaMemcpyDeviceToHost, s));
            if (trbound) DN_CUDA(cudaMemcpyAsync(out.trace, dtr.p, sizeof(uint16_t) * trbound, cudaMemcpyDeviceToHost, s));
        }
        DN_CUDA(cudaMemcpyAsync(hctr, ctr.p, sizeof hctr, cudaMemcpyDeviceToHost, s));
        DN_CUDA(cudaMemcpyAsync(&tot, dtotal.p, 8, cudaMemcpyDeviceToHost, s));
        DN_CUDA(cudaStreamSynchronize(s));
        const int nkeep = ncand - (int)hctr[0];
        if (getenv("DN_TRACE")) fprintf(stderr, "[dn trace] candidates %d, kept after duplicate removal %d\n", ncand, nkeep);
        out.nrec = nkeep;
        aligned = (int64_t)hctr[1]; ext_bytes = (int64_t)hctr[2];
        if (keep) { keep->rec = drec.p; keep->toff = dtoff.p; keep->trace = dtr.p; keep->nrec = nkeep; keep->ntrace = tot; }
    }
    if (keep) { out.nrec = 0; tot = keep->ntrace; }          // the records stay in HBM: the host result holds the statistics only
    if (!out.rec) out.rec = (dn_las_record *)hcache_alloc(64);
    if (!out.toff) out.toff = (int64_t *)hcache_alloc(64);
    if (!out.trace) out.trace = (uint16_t *)hcache_alloc(64);
    out.ntrace = keep ? 0 : tot;
    tr.mark("dedupe + order + download");
    out.stats.las = keep ? keep->nrec : out.nrec; out.stats.aligned_bases = aligned; out.stats.trace_points = tot / 2;
    out.stats.algo_bytes_seed = abytes; out.stats.algo_bytes_extend = ext_bytes;
    out.stats.ms_total = tt.stop();                          // syncs the stream: the extension brackets are complete
    for (auto &e : ext_ev) { float ms = 0; cudaEventElapsedTime(&ms, e.first, e.second); ms_ext += ms; cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    out.stats.ms_extend = ms_ext; out.stats.ms_seed = out.stats.ms_total - ms_ext;      // seed = everything but the k_extend launches
    out.stats.launches = g_launches.load() - launches0;
}

}  // namespace dn
