// engine.cuh -- internal interfaces of the alignment engine (device-resident blocks, parameters,
// stage entry points).  The public boundary is include/dentist_b200.h.
#pragma once
#include "common.cuh"
#include "../../include/dentist_b200.h"

namespace dn {

// A sequence block resident in HBM.
//  * every read starts at a multiple of 64 bases (16 B of 2-bit data) in one concatenated,
//    padded coordinate space ("g" coordinates); `off[r]` is that start, off[nreads] the total
//  * fwd / rc: 2-bit codes, 16 bases per u32, base p of the block in bits 2*(p&15) of word p>>4
//    (LSB-first so unaligned windows are a funnel shift); rc holds every read reverse-complemented
//    in place (same offsets) -- as daligner complements a block once instead of per alignment
//  * chunk2read[g>>10] = read whose padded extent contains g (O(1) position -> read lookup)
//  * mask / mask_rc: optional bit per g coordinate (1 = masked: no k-mer may touch it)
struct DevBlock {
    int nreads = 0;
    int64_t total = 0;            // padded total bases
    int64_t total_real = 0;       // sum of read lengths
    int maxlen = 0;
    std::vector<int64_t> h_off;
    std::vector<int32_t> h_len;
    DBuf<u32> fwd, rc;
    DBuf<int64_t> off;
    DBuf<int32_t> len;
    DBuf<int32_t> chunk2read;
    DBuf<u32> mask, mask_rc;
    bool has_mask = false;
    DBuf<int32_t> group;          // optional pile id per read
    bool has_group = false;
    // asynchronous upload (dn_align_host): the copy stream records `ready`, the first consumer's stream waits on it;
    // the staging buffers live until the block dies
    cudaEvent_t ready = nullptr;
    // chunked upload: chunk c = words [chunk_word[c], chunk_word[c + 1]) of fwd / rc, packed when chunk_ready[c] fires;
    // meta_ready = offsets, lengths, chunk2read are on the device
    cudaEvent_t meta_ready = nullptr;
    std::vector<cudaEvent_t> chunk_ready; std::vector<int64_t> chunk_word;
    DBuf<uint8_t> up_raw; DBuf<int64_t> up_boff;
    DevBlock() {}
    DevBlock(const DevBlock &) = delete; DevBlock &operator=(const DevBlock &) = delete;
    ~DevBlock() { if (ready) cudaEventDestroy(ready); if (meta_ready) cudaEventDestroy(meta_ready); for (cudaEvent_t e : chunk_ready) cudaEventDestroy(e); }
    // optional resident k-mer index (dn_block_index): the sorted tuple list, prefix table and k-mer filter that
    // align_blocks otherwise rebuilds for the A block on every call.  Invalidated when the seed mask changes.
    struct Index {
        int k = 0, tbits = 0, kbits_log2 = 27;
        int tbl_shift = 0;             // the prefix table starts at tbl.p + tbl_shift (1: built by bucket counting, bucket.cu)
        int pb = 0;                    // > 0: k > 15 with 8-byte packed entries kmer << pb | position in `ta`
        int64_t n = 0;                 // index entries (k > 15: valid positions only)
        DBuf<u64> ta; DBuf<ulonglong2> tw; DBuf<u32> tbl, kbits;
        bool valid = false;
        void drop() { ta.release(); tw.release(); tbl.release(); kbits.release(); valid = false; k = 0; pb = 0; }
    } index;
};
void block_build_index(DevBlock &A, int k, cudaStream_t s);

struct AlignParams {
    int k = 14, w = 6, h = 35, t = 32, tspace = 100, minlen = 500, cdiff = 20, xdrop = 300, wmax = 30,
        rounds = 3, self = 0, poolmul = 64, join_mode = 0;   // join_mode: 0 auto, 1 sorted merge, 2 index lookup
};

// Host result in its final (C ABI) form; arrays come from the recycling host cache (hcache_*).
struct HostLas {
    dn_las_record *rec = nullptr;       // sorted in LAsort order
    int64_t *toff = nullptr;            // trace offset (uint16 units) per record
    uint16_t *trace = nullptr;          // (diffs, bbases) pairs
    int64_t nrec = 0, ntrace = 0;
    dn_align_stats stats{};
};

// Recycling host allocator for result buffers: fresh multi-MB mallocs page-fault for milliseconds.
void *hcache_alloc(size_t bytes);
void hcache_free(void *p);

void block_upload(const dn_block_desc &d, DevBlock &out, cudaStream_t s, bool async = false, cudaStream_t pack_stream = nullptr);
void block_add_mask(DevBlock &B, const int64_t *anno_h, const int32_t *data_h, cudaStream_t s);
void block_crop(const DevBlock &src, int n, const int32_t *read, const int32_t *begin, const int32_t *end, const int32_t *group,
                DevBlock &out, cudaStream_t s);
void force_flat_device(dn_las_record *h_rec, int64_t *h_toff, int64_t n, cudaStream_t s);
void merge_las_device(const dn_las_record *d_rec, int64_t n, const uint16_t *d_trace, int64_t ntrace, int64_t max_alen, int64_t max_blen,
                      int64_t na_reads, int64_t nb_reads, HostLas &out, cudaStream_t s, bool reset_arena = true);
// The same result left in HBM (arena memory: valid until the next call resets the arena) -- the input of the multi-GPU
// gather, which must not bounce through the host.
struct DevLas { dn_las_record *rec = nullptr; int64_t *toff = nullptr; uint16_t *trace = nullptr; int64_t nrec = 0, ntrace = 0; };
// keep != nullptr: records / traces stay on the device (out gets the statistics only)
void align_blocks(const DevBlock &A, const DevBlock &B, const AlignParams &P, HostLas &out, cudaStream_t s, DevLas *keep = nullptr);
// LAmerge for the all-gathered segments of `world` ranks (segment r = records [seg_beg[r], seg_beg[r+1]) in LAsort order,
// B reads of rank r all below those of rank r+1): one pass of run offsets + placement instead of a full re-sort
void merge_segments_device(const dn_las_record *d_rec, const int64_t *h_seg_beg, int world, const uint16_t *d_trace, int64_t ntrace,
                           int64_t na_reads, HostLas &out, cudaStream_t s);

}  // namespace dn
