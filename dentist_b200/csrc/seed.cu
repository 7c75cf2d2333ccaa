// seed.cu -- block upload (2-bit pack + reverse complement), k-mer tuple emit, sorted-list join
// into seed hits, diagonal-band filter and seed selection.  Stages K0-K4 of DESIGN.md; the
// device-side replacement for daligner's tuple sort / merge / band filter that DENTIST reaches
// through dazzler.d:6131-6170.
#include "engine.cuh"
#include "seed.cuh"
#include "scan.cuh"

namespace dn {

// ------------------------------------------------------------------------- K0: upload / pack

__global__ void __launch_bounds__(128) k_pack(const uint8_t *__restrict__ data, const int64_t *__restrict__ boff,
                                              const int32_t *__restrict__ len, const int64_t *__restrict__ off,
                                              int format, u32 *__restrict__ fwd, u32 *__restrict__ rc, int r0) {
    const int r = r0 + blockIdx.x;
    const int L = len[r];
    const uint8_t *src = data + boff[r];
    const int64_t w0 = off[r] >> 4;
    const int nw = (int)((off[r + 1] - off[r]) >> 4);
    for (int w = threadIdx.x; w < nw; w += blockDim.x) {
        u32 f = 0, c = 0;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            int p = w * 16 + j;
            if (p < L) {
                u32 bf, br;
                int q = L - 1 - p;
                if (format == DN_SEQ_BYTES) { bf = src[p] & 3u; br = 3u - (src[q] & 3u); }
                else {
                    bf = (src[p >> 2] >> (6 - 2 * (p & 3))) & 3u;
                    br = 3u - ((src[q >> 2] >> (6 - 2 * (q & 3))) & 3u);
                }
                f |= bf << (2 * j); c |= br << (2 * j);
            }
        }
        fwd[w0 + w] = f; rc[w0 + w] = c;
    }
}

__global__ void k_chunk2read(const int64_t *__restrict__ off, int nreads, int32_t *__restrict__ c2r) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nreads) return;
    int64_t c0 = (off[r] + 1023) >> 10, c1 = (off[r + 1] + 1023) >> 10;
    for (int64_t c = c0; c < c1; c++) c2r[c] = r;
}

// one thread per mask interval: set bits in the forward and the mirrored (rc) bit arrays
__global__ void k_mask_bits(const int64_t *__restrict__ anno, const int32_t *__restrict__ mdata, int nreads,
                            const int64_t *__restrict__ off, const int32_t *__restrict__ len,
                            u32 *__restrict__ mask, u32 *__restrict__ mask_rc) {
    int r = blockIdx.x;
    int64_t i0 = anno[r] / 8, i1 = anno[r + 1] / 8;      // (begin,end) int32 pairs
    const int L = len[r];
    for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        int b = mdata[2 * i], e = mdata[2 * i + 1];
        if (b < 0) b = 0;
        if (e > L) e = L;
        for (int p = b; p < e; p++) {
            int64_t g = off[r] + p, gr = off[r] + (L - 1 - p);
            atomicOr(&mask[g >> 5], 1u << (g & 31));
            atomicOr(&mask_rc[gr >> 5], 1u << (gr & 31));
        }
    }
}

// ORs the intervals of a mask track (anno = byte offsets into int32 (begin, end) pairs, dazzler.d:4943-5052) into the
// block's seed-exclusion bits; several tracks (-mdust -mrep) accumulate.
void block_add_mask(DevBlock &B, const int64_t *anno_h, const int32_t *data_h, cudaStream_t s) {
    const int64_t nbytes = anno_h[B.nreads];
    if (nbytes <= 0) return;
    B.index.drop();                                        // masked k-mers leave the index: rebuild on demand
    if (!B.has_mask) {
        const size_t mw = (size_t)(B.total >> 5) + 4;
        B.mask.persistent(mw); B.mask_rc.persistent(mw); B.mask.zero(s); B.mask_rc.zero(s);
        B.has_mask = true;
    }
    DBuf<int64_t> anno; anno.persistent(B.nreads + 1); DBuf<int32_t> md; md.persistent((size_t)nbytes / 4 + 2);
    DN_CUDA(cudaMemcpyAsync(anno.p, anno_h, sizeof(int64_t) * (B.nreads + 1), cudaMemcpyHostToDevice, s));
    DN_CUDA(cudaMemcpyAsync(md.p, data_h, nbytes, cudaMemcpyHostToDevice, s));
    DN_LAUNCH(k_mask_bits, B.nreads, 64, 0, s, (const int64_t *)anno.p, (const int32_t *)md.p, B.nreads,
              (const int64_t *)B.off.p, (const int32_t *)B.len.p, B.mask.p, B.mask_rc.p);
    DN_CUDA(cudaStreamSynchronize(s));
}

// async = true: no host sync at the end; `B.ready` is recorded and the staging buffers stay with the block
// (the caller keeps the host arrays alive and untouched until the block was consumed -- dn_align_host does).
// pack_stream != nullptr (async only): the raw bytes travel in chunks on `s`, every chunk is packed on `pack_stream` as soon as it
// has arrived and announces itself through B.chunk_ready[c] -- the consumer's first pass over the block (the lookup join's count
// pass) follows the chunks instead of waiting for the whole upload.
void block_upload(const dn_block_desc &d, DevBlock &B, cudaStream_t s, bool async, cudaStream_t pack_stream) {
    if (d.nreads < 0 || (d.nreads > 0 && (!d.rlen || !d.boff || !d.data))) throw Error("dn_block_desc: null field");
    B.nreads = d.nreads;
    B.h_len.assign(d.rlen, d.rlen + d.nreads);
    B.h_off.resize(d.nreads + 1);
    int64_t g = 0; B.maxlen = 0; B.total_real = 0;
    bool monotonic = true;
    for (int r = 0; r < d.nreads; r++) {
        if (d.rlen[r] < 0) throw Error("negative read length");
        B.h_off[r] = g; g += ((int64_t)d.rlen[r] + 63) / 64 * 64;
        if (d.rlen[r] > B.maxlen) B.maxlen = d.rlen[r];
        B.total_real += d.rlen[r];
        int64_t need = d.format == DN_SEQ_BPS ? ((int64_t)d.rlen[r] + 3) / 4 : d.rlen[r];
        if (d.boff[r] < 0 || d.boff[r] + need > d.data_bytes) throw Error("dn_block_desc: read outside data");
        if (r > 0 && d.boff[r] < d.boff[r - 1]) monotonic = false;
    }
    B.h_off[d.nreads] = g; B.total = g;
    if (2 * g >= (1ll << 32) - 1024) throw Error("block too large: 2*padded bases must be < 2^32");
    size_t nwords = (size_t)(g >> 4) + 8;
    // (read per call: tests force the chunked path on small blocks)
    const int want_chunks = getenv("DN_UPLOAD_CHUNKS") ? atoi(getenv("DN_UPLOAD_CHUNKS")) : 4;
    const int64_t min_bytes = getenv("DN_UPLOAD_CHUNK_MIN_BYTES") ? atoll(getenv("DN_UPLOAD_CHUNK_MIN_BYTES")) : (8 << 20);
    const bool chunked = async && pack_stream && want_chunks > 1 && want_chunks <= 64 && monotonic && d.nreads >= 2 * want_chunks &&
                         !(d.mask_anno && d.mask_data) && d.data_bytes >= min_bytes;
    cudaStream_t ps = chunked ? pack_stream : s;              // the stream the device-side work of the upload runs on
    B.fwd.persistent(nwords); B.rc.persistent(nwords);
    B.fwd.zero(ps); B.rc.zero(ps);
    B.off.persistent(d.nreads + 1); B.len.persistent(d.nreads > 0 ? d.nreads : 1);
    B.chunk2read.persistent((size_t)(g >> 10) + 2);
    B.chunk2read.zero(ps);
    DN_CUDA(cudaMemcpyAsync(B.off.p, B.h_off.data(), sizeof(int64_t) * (d.nreads + 1), cudaMemcpyHostToDevice, s));
    if (d.nreads == 0) { DN_CUDA(cudaStreamSynchronize(s)); return; }
    DN_CUDA(cudaMemcpyAsync(B.len.p, B.h_len.data(), sizeof(int32_t) * d.nreads, cudaMemcpyHostToDevice, s));
    DBuf<uint8_t> &raw = B.up_raw; raw.persistent((size_t)d.data_bytes + 16);
    DBuf<int64_t> &boff = B.up_boff; boff.persistent(d.nreads);
    DN_CUDA(cudaMemcpyAsync(boff.p, d.boff, sizeof(int64_t) * d.nreads, cudaMemcpyHostToDevice, s));
    B.has_group = d.group != nullptr;
    if (d.group) {
        B.group.persistent(d.nreads);
        DN_CUDA(cudaMemcpyAsync(B.group.p, d.group, sizeof(int32_t) * d.nreads, cudaMemcpyHostToDevice, s));
    }
    B.has_mask = false;
    auto event = [](cudaEvent_t &e) { if (!e) DN_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); return e; };
    if (chunked) {
        cudaEvent_t meta_copied = nullptr; event(meta_copied);
        DN_CUDA(cudaEventRecord(meta_copied, s));
        DN_CUDA(cudaStreamWaitEvent(ps, meta_copied, 0));
        DN_LAUNCH(k_chunk2read, (d.nreads + 255) / 256, 256, 0, ps, (const int64_t *)B.off.p, d.nreads, B.chunk2read.p);
        DN_CUDA(cudaEventRecord(event(B.meta_ready), ps));
        const int C = want_chunks;
        for (cudaEvent_t e : B.chunk_ready) cudaEventDestroy(e);
        B.chunk_ready.assign(C, nullptr); B.chunk_word.assign(C + 1, 0);
        std::vector<cudaEvent_t> copied(C, nullptr);
        for (int c = 0; c < C; c++) {
            const int r0 = (int)((int64_t)d.nreads * c / C), r1 = (int)((int64_t)d.nreads * (c + 1) / C);
            const int64_t b0 = c == 0 ? 0 : d.boff[r0], b1 = c == C - 1 ? d.data_bytes : d.boff[r1];
            if (b1 > b0) DN_CUDA(cudaMemcpyAsync(raw.p + b0, (const uint8_t *)d.data + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, s));
            DN_CUDA(cudaEventRecord(event(copied[c]), s));
            DN_CUDA(cudaStreamWaitEvent(ps, copied[c], 0));
            if (r1 > r0)
                DN_LAUNCH(k_pack, r1 - r0, 128, 0, ps, (const uint8_t *)raw.p, (const int64_t *)boff.p, (const int32_t *)B.len.p,
                          (const int64_t *)B.off.p, d.format, B.fwd.p, B.rc.p, r0);
            DN_CUDA(cudaEventRecord(event(B.chunk_ready[c]), ps));
            B.chunk_word[c] = B.h_off[r0] >> 4; B.chunk_word[c + 1] = B.h_off[r1] >> 4;
        }
        DN_CUDA(cudaEventRecord(event(B.ready), ps));
        // the helper events may be destroyed once recorded and waited on: the driver keeps what the streams still need
        cudaEventDestroy(meta_copied); for (cudaEvent_t e : copied) cudaEventDestroy(e);
        return;
    }
    DN_CUDA(cudaMemcpyAsync(raw.p, d.data, d.data_bytes, cudaMemcpyHostToDevice, s));
    DN_LAUNCH(k_pack, d.nreads, 128, 0, s, (const uint8_t *)raw.p, (const int64_t *)boff.p, (const int32_t *)B.len.p,
              (const int64_t *)B.off.p, d.format, B.fwd.p, B.rc.p, 0);
    DN_LAUNCH(k_chunk2read, (d.nreads + 255) / 256, 256, 0, s, (const int64_t *)B.off.p, d.nreads, B.chunk2read.p);
    if (d.mask_anno && d.mask_data) block_add_mask(B, d.mask_anno, d.mask_data, s);
    if (async) {
        DN_CUDA(cudaEventRecord(event(B.ready), s));
        return;
    }
    DN_CUDA(cudaStreamSynchronize(s));
    raw.release(); boff.release();
}

// ------------------------------------------------------------------------- crop
// Cropped pile-up reads straight from the resident read block (cropper.d:383-421 builds a FASTA and a
// fresh DB instead): output read i = bases [begin, end) of source read `read[i]`.  Both strands are
// shifted copies: fwd_out = fwd_src + begin, rc_out = rc_src + (L - end).  One thread per output word.
__global__ void __launch_bounds__(128) k_crop(const u32 *__restrict__ sfwd, const u32 *__restrict__ src_rc,
                                              const int64_t *__restrict__ soff, const int32_t *__restrict__ slen,
                                              const int32_t *__restrict__ read, const int32_t *__restrict__ begin,
                                              const int32_t *__restrict__ len, const int64_t *__restrict__ off,
                                              u32 *__restrict__ fwd, u32 *__restrict__ rc) {
    const int r = blockIdx.x;
    const int L = len[r], s = read[r];
    const int64_t gf = soff[s] + begin[r], gr = soff[s] + (slen[s] - (begin[r] + L));
    const int64_t w0 = off[r] >> 4;
    const int nw = (int)((off[r + 1] - off[r]) >> 4);
    for (int w = threadIdx.x; w < nw; w += blockDim.x) {
        const int p = w * 16, n = min(16, L - p);
        u32 f = 0, c = 0;
        if (n > 0) {
            const int64_t a = gf + p, b = gr + p;
            f = __funnelshift_r(sfwd[a >> 4], sfwd[(a >> 4) + 1], (int)(a & 15) << 1);
            c = __funnelshift_r(src_rc[b >> 4], src_rc[(b >> 4) + 1], (int)(b & 15) << 1);
            if (n < 16) { const u32 m = (1u << (2 * n)) - 1u; f &= m; c &= m; }
        }
        fwd[w0 + w] = f; rc[w0 + w] = c;
    }
}

void block_crop(const DevBlock &S, int n, const int32_t *read, const int32_t *begin, const int32_t *end, const int32_t *group,
                DevBlock &B, cudaStream_t s) {
    B.nreads = n; B.h_len.resize(n); B.h_off.resize(n + 1);
    int64_t g = 0; B.maxlen = 0; B.total_real = 0;
    for (int r = 0; r < n; r++) {
        if (read[r] < 0 || read[r] >= S.nreads) throw Error("crop: read id out of bounds");
        if (begin[r] < 0 || end[r] < begin[r] || end[r] > S.h_len[read[r]]) throw Error("crop: slice outside the read");
        const int L = end[r] - begin[r];
        B.h_len[r] = L; B.h_off[r] = g; g += ((int64_t)L + 63) / 64 * 64;
        if (L > B.maxlen) B.maxlen = L;
        B.total_real += L;
    }
    B.h_off[n] = g; B.total = g;
    const size_t nwords = (size_t)(g >> 4) + 8;
    B.fwd.persistent(nwords); B.rc.persistent(nwords); B.fwd.zero(s); B.rc.zero(s);
    B.off.persistent(n + 1); B.len.persistent(n > 0 ? n : 1); B.chunk2read.persistent((size_t)(g >> 10) + 2); B.chunk2read.zero(s);
    DN_CUDA(cudaMemcpyAsync(B.off.p, B.h_off.data(), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, s));
    B.has_mask = false; B.has_group = group != nullptr;
    if (n == 0) { DN_CUDA(cudaStreamSynchronize(s)); return; }
    DN_CUDA(cudaMemcpyAsync(B.len.p, B.h_len.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
    DBuf<int32_t> dread, dbeg; dread.persistent(n); dbeg.persistent(n);
    DN_CUDA(cudaMemcpyAsync(dread.p, read, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
    DN_CUDA(cudaMemcpyAsync(dbeg.p, begin, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
    if (group) { B.group.persistent(n); DN_CUDA(cudaMemcpyAsync(B.group.p, group, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s)); }
    DN_LAUNCH(k_crop, n, 128, 0, s, (const u32 *)S.fwd.p, (const u32 *)S.rc.p, (const int64_t *)S.off.p, (const int32_t *)S.len.p,
              (const int32_t *)dread.p, (const int32_t *)dbeg.p, (const int32_t *)B.len.p, (const int64_t *)B.off.p, B.fwd.p, B.rc.p);
    DN_LAUNCH(k_chunk2read, (n + 255) / 256, 256, 0, s, (const int64_t *)B.off.p, n, B.chunk2read.p);
    DN_CUDA(cudaStreamSynchronize(s));
}

// ------------------------------------------------------------------------- K1: k-mer tuples

// One thread per 16-base word of the padded block: emits 16 tuples (kmer << 32 | payload).
// Positions that cannot start a k-mer (read end, padding, masked) get kmer 0xffffffff and sort last.
__global__ void __launch_bounds__(256) k_tuples(const u32 *__restrict__ seq, const u32 *__restrict__ maskbits,
                                                const int64_t *__restrict__ off, const int32_t *__restrict__ len,
                                                const int32_t *__restrict__ c2r, int64_t nwords, int k, u32 payload_base,
                                                u64 *__restrict__ out) {
    int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= nwords) return;
    const int64_t g0 = wi << 4;
    int r = c2r[g0 >> 10];
    while (off[r + 1] <= g0) r++;
    const int p0 = (int)(g0 - off[r]);
    const int L = len[r];
    const u64 v = ((u64)seq[wi + 1] << 32) | seq[wi];
    const u64 kmask = (1ull << (2 * k)) - 1ull;
    u64 mwin = 0;
    if (maskbits) {
        int64_t mw = g0 >> 5;
        mwin = (((u64)maskbits[mw + 1] << 32) | maskbits[mw]) >> (g0 & 31);
    }
    const u64 mk = (1ull << k) - 1ull;
    ulonglong2 *o2 = reinterpret_cast<ulonglong2 *>(out + g0);
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
        u64 t[2];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            int jj = j + q;
            bool valid = (p0 + jj + k <= L) && (((mwin >> jj) & mk) == 0);
            u64 km = valid ? ((v >> (2 * jj)) & kmask) : 0xffffffffull;
            t[q] = (km << 32) | (u64)(payload_base + (u32)(g0 + jj));
        }
        o2[j >> 1] = make_ulonglong2(t[0], t[1]);
    }
}

void emit_tuples(const DevBlock &B, bool rc, int k, u32 payload_base, u64 *out, cudaStream_t s) {
    int64_t nwords = B.total >> 4;
    if (nwords == 0) return;
    const u32 *mb = B.has_mask ? (rc ? B.mask_rc.p : B.mask.p) : nullptr;
    DN_LAUNCH(k_tuples, (unsigned)((nwords + 255) / 256), 256, 0, s, (const u32 *)(rc ? B.rc.p : B.fwd.p), mb,
              (const int64_t *)B.off.p, (const int32_t *)B.len.p, (const int32_t *)B.chunk2read.p, nwords, k, payload_base, out);
}

// k = 16..31: the k-mer needs up to 62 bits, so an A tuple is 16 bytes {kmer, global position}; invalid positions get
// kmer ~0 and sort last.  A k-mer starting at base jj <= 15 of the word ends before base 46: three packed words.
__device__ __forceinline__ u64 wide_kmer(u64 v, u32 w2, int jj, u64 kmask) {
    const u64 x = jj ? ((v >> (2 * jj)) | ((u64)w2 << (64 - 2 * jj))) : v;
    return x & kmask;
}

// Only positions that can start a k-mer become tuples (no sentinel entries to sort): pass 1 counts them per word,
// pass 2 writes them behind the scanned offsets.  The sort key is then exactly the 2k k-mer bits.
__device__ __forceinline__ u32 wide_valid_mask(const u32 *__restrict__ maskbits, const int64_t *__restrict__ off,
                                               const int32_t *__restrict__ len, const int32_t *__restrict__ c2r, int64_t wi, int k) {
    const int64_t g0 = wi << 4;
    int r = c2r[g0 >> 10];
    while (off[r + 1] <= g0) r++;
    const int p0 = (int)(g0 - off[r]);
    const int L = len[r];
    const u64 mk = (1ull << k) - 1ull;
    u64 mwin = 0;
    if (maskbits) { int64_t mw = g0 >> 5; mwin = (((u64)maskbits[mw + 1] << 32) | maskbits[mw]) >> (g0 & 31); }
    u32 valid = 0;
#pragma unroll
    for (int jj = 0; jj < 16; jj++)
        if ((p0 + jj + k <= L) && (((mwin >> jj) & mk) == 0)) valid |= 1u << jj;
    return valid;
}

__global__ void __launch_bounds__(256) k_tuples_wide_count(const u32 *__restrict__ maskbits, const int64_t *__restrict__ off,
                                                           const int32_t *__restrict__ len, const int32_t *__restrict__ c2r,
                                                           int64_t nwords, int k, u32 *__restrict__ cnt) {
    int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= nwords) return;
    cnt[wi] = (u32)__popc(wide_valid_mask(maskbits, off, len, c2r, wi, k));
}

// PB = 0: 16-byte {kmer, position} tuples; PB > 0 (2k + PB <= 64): 8-byte tuples kmer << PB | position
template <bool PACKED>
__global__ void __launch_bounds__(256) k_tuples_wide(const u32 *__restrict__ seq, const u32 *__restrict__ maskbits,
                                                     const int64_t *__restrict__ off, const int32_t *__restrict__ len,
                                                     const int32_t *__restrict__ c2r, int64_t nwords, int k,
                                                     const int64_t *__restrict__ woff, void *__restrict__ out_, int pb) {
    int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= nwords) return;
    u32 valid = wide_valid_mask(maskbits, off, len, c2r, wi, k);
    if (!valid) return;
    const int64_t g0 = wi << 4;
    const u64 v = ((u64)seq[wi + 1] << 32) | seq[wi];
    const u32 w2 = seq[wi + 2];
    const u64 kmask = (1ull << (2 * k)) - 1ull;
    int64_t o = woff[wi];
    while (valid) {
        const int jj = __ffs(valid) - 1; valid &= valid - 1;
        if (PACKED) reinterpret_cast<u64 *>(out_)[o++] = (wide_kmer(v, w2, jj, kmask) << pb) | (u64)(u32)(g0 + jj);
        else reinterpret_cast<ulonglong2 *>(out_)[o++] = make_ulonglong2(wide_kmer(v, w2, jj, kmask), (u64)(u32)(g0 + jj));
    }
}

// returns the number of tuples written (a host sync: the caller sizes the sort and the index with it)
int64_t emit_tuples_wide(const DevBlock &B, int k, void *out, int pb, cudaStream_t s) {
    int64_t nwords = B.total >> 4;
    if (nwords == 0) return 0;
    const u32 *mb = B.has_mask ? (const u32 *)B.mask.p : nullptr;
    DBuf<u32> cnt(nwords); DBuf<int64_t> woff(nwords), tot(1);
    DN_LAUNCH(k_tuples_wide_count, (unsigned)((nwords + 255) / 256), 256, 0, s, mb, (const int64_t *)B.off.p, (const int32_t *)B.len.p,
              (const int32_t *)B.chunk2read.p, nwords, k, cnt.p);
    exclusive_scan_u32_to_i64(cnt.p, woff.p, nwords, tot.p, s);
    if (pb > 0)
        DN_LAUNCH(k_tuples_wide<true>, (unsigned)((nwords + 255) / 256), 256, 0, s, (const u32 *)B.fwd.p, mb,
                  (const int64_t *)B.off.p, (const int32_t *)B.len.p, (const int32_t *)B.chunk2read.p, nwords, k, (const int64_t *)woff.p, out, pb);
    else
        DN_LAUNCH(k_tuples_wide<false>, (unsigned)((nwords + 255) / 256), 256, 0, s, (const u32 *)B.fwd.p, mb,
                  (const int64_t *)B.off.p, (const int32_t *)B.len.p, (const int32_t *)B.chunk2read.p, nwords, k, (const int64_t *)woff.p, out, 0);
    int64_t n = 0;
    DN_CUDA(cudaMemcpyAsync(&n, tot.p, sizeof n, cudaMemcpyDeviceToHost, s)); DN_CUDA(cudaStreamSynchronize(s));
    return n;
}

// ------------------------------------------------------------------------- K3: join

// The sorted A index behind one interface: 8-byte tuples (kmer << 32 | position) for k <= 15, 16-byte tuples
// {kmer, position} for k = 16..31.
struct Idx32 {
    const u64 *t;
    typedef u32 key_t;
    static constexpr bool wide = false;
    __device__ __forceinline__ u32 key(int64_t i) const { return (u32)(t[i] >> 32); }
    __device__ __forceinline__ u32 pos(int64_t i) const { return (u32)t[i]; }
    __device__ __forceinline__ static bool invalid(u32 km) { return km == 0xffffffffu; }
    __device__ __forceinline__ static u32 fold(u32 km) { return km; }
};
struct Idx64 {
    const ulonglong2 *t;
    typedef u64 key_t;
    static constexpr bool wide = true;
    __device__ __forceinline__ u64 key(int64_t i) const { return t[i].x; }
    __device__ __forceinline__ u32 pos(int64_t i) const { return (u32)t[i].y; }
    __device__ __forceinline__ static bool invalid(u64 km) { return km == ~0ull; }
    __device__ __forceinline__ static u32 fold(u64 km) { return (u32)(km ^ (km >> 31)); }
};

// k = 16..31 in 8 bytes when 2k + pb <= 64 (pb = bits of the block's padded size): kmer << pb | position.  Half the
// bytes of Idx64 in the sort, the table build and every index walk of the lookup join.
struct IdxP {
    const u64 *t; int pb;
    typedef u64 key_t;
    static constexpr bool wide = true;
    __device__ __forceinline__ u64 key(int64_t i) const { return t[i] >> pb; }
    __device__ __forceinline__ u32 pos(int64_t i) const { return (u32)(t[i] & ((1ull << pb) - 1ull)); }
    __device__ __forceinline__ static bool invalid(u64) { return false; }
    __device__ __forceinline__ static u32 fold(u64 km) { return (u32)(km ^ (km >> 31)); }
};

// tbl[q] = first index i in the sorted A list with min(kmer_i >> sh, nq) >= q, q in [0, nq]
template <class IDX>
__device__ __forceinline__ void prefix_table_body(IDX ta, int64_t na, int sh, u32 nq, u32 *__restrict__ tbl) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= na) return;
    u64 q = (u64)ta.key(i) >> sh; if (q > nq) q = nq;
    int64_t qp;
    if (i == 0) qp = -1;
    else { u64 t = (u64)ta.key(i - 1) >> sh; if (t > nq) t = nq; qp = (int64_t)t; }
    for (int64_t x = qp + 1; x <= (int64_t)q; x++) tbl[x] = (u32)i;
    if (i == na - 1) for (int64_t x = (int64_t)q + 1; x <= (int64_t)nq; x++) tbl[x] = (u32)na;
}
__global__ void __launch_bounds__(256) k_prefix_table(const u64 *__restrict__ ta, int64_t na, int sh, u32 nq, u32 *__restrict__ tbl) {
    prefix_table_body(Idx32{ta}, na, sh, nq, tbl);
}
__global__ void __launch_bounds__(256) k_prefix_table_w(const ulonglong2 *__restrict__ ta, int64_t na, int sh, u32 nq, u32 *__restrict__ tbl) {
    prefix_table_body(Idx64{ta}, na, sh, nq, tbl);
}

__global__ void __launch_bounds__(256) k_prefix_table_p(const u64 *__restrict__ ta, int pb, int64_t na, int sh, u32 nq, u32 *__restrict__ tbl) {
    prefix_table_body(IdxP{ta, pb}, na, sh, nq, tbl);
}

__device__ __forceinline__ void a_range(const u64 *__restrict__ ta, const u32 *__restrict__ tbl, int sh, u32 km, u32 &s, u32 &e) {
    u32 q = km >> sh;
    u32 lo = tbl[q], hi = tbl[q + 1];
    // lower bound of km in [lo,hi)
    u32 a = lo, b = hi;
    while (a < b) { u32 m = (a + b) >> 1; if ((u32)(ta[m] >> 32) < km) a = m + 1; else b = m; }
    s = a;
    b = hi;
    while (a < b) { u32 m = (a + b) >> 1; if ((u32)(ta[m] >> 32) <= km) a = m + 1; else b = m; }
    e = a;
}

__global__ void __launch_bounds__(256) k_join_count(const u64 *__restrict__ ta, const u32 *__restrict__ tbl, int sh,
                                                    const u64 *__restrict__ tb, int64_t nb, int tcap,
                                                    u32 *__restrict__ cnt, u32 *__restrict__ start) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nb) return;
    u32 km = (u32)(tb[j] >> 32);
    u32 c = 0, s = 0;
    if (km != 0xffffffffu) {
        u32 e; a_range(ta, tbl, sh, km, s, e);
        c = e - s;
        if (c > (u32)tcap) c = 0;
    }
    cnt[j] = c; start[j] = s;
}

__device__ __forceinline__ int read_of(const int32_t *__restrict__ c2r, const int64_t *__restrict__ off, int64_t g) {
    int r = c2r[g >> 10];
    while (off[r + 1] <= g) r++;
    return r;
}

__device__ __forceinline__ bool pair_ok(const JoinGeom &G, int ar, int br) {
    return !((G.self && ar == br) || (G.a_group && G.a_group[ar] != G.b_group[br]));
}

// hit record: .x = key = (bs << gdbits) | gd   (invalid: 1 << keybits), .y = apos | bpos << 32
__global__ void __launch_bounds__(256) k_join_emit(const u64 *__restrict__ ta, const u64 *__restrict__ tb, int64_t nb,
                                                   const u32 *__restrict__ cnt, const u32 *__restrict__ start,
                                                   const int64_t *__restrict__ hoff, JoinGeom G,
                                                   ulonglong2 *__restrict__ hits, unsigned long long *__restrict__ ninvalid) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nb) return;
    u32 c = cnt[j];
    if (c == 0) return;
    u32 pb = (u32)tb[j];
    int strand = pb >= G.nbp ? 1 : 0;
    int64_t gb = strand ? (int64_t)pb - G.nbp : (int64_t)pb;
    int br = read_of(G.b_c2r, G.b_off, gb);
    int bpos = (int)(gb - G.b_off[br]);
    u64 bs = (u64)strand * G.nb_reads + br;
    int64_t o = hoff[j];
    u32 s = start[j];
    u32 ninv = 0;
    for (u32 x = 0; x < c; x++) {
        int64_t ga = (int64_t)(u32)ta[s + x];
        int ar = read_of(G.a_c2r, G.a_off, ga);
        int apos = (int)(ga - G.a_off[ar]);
        u64 key;
        if ((G.self && ar == br) || (G.a_group && G.a_group[ar] != G.b_group[br])) { key = 1ull << G.keybits; ninv++; }
        else { u64 gd = (u64)(G.a_dbase[ar] + apos - bpos + G.maxlb); key = (bs << G.gdbits) | gd; }
        hits[o + x] = make_ulonglong2(key, (u64)(u32)apos | ((u64)(u32)bpos << 32));
    }
    if (ninv) atomicAdd(ninvalid, (unsigned long long)ninv);
}

// ------------------------------------------------------------------------- K3': index lookup join
// When the A index (sorted tuple list + prefix table) is L2-sized, B's tuples are never
// materialised or sorted: one thread per 16-base word of B recomputes its 16 k-mers, looks each up
// in the A index and (pass 1) counts / (pass 2) emits the hits.  Output-identical to the
// sorted-merge join because the hit list is totally ordered by the sort that follows.

// k-mer presence filter (blocked Bloom: one 32-bit word per k-mer, three bits inside it; >= 12 bits per indexed position,
// 16 MB and pinned in L2 for a 10 Mbp block).  Most read k-mers carry a sequencing error and occur nowhere in A; the filter
// rejects them with ONE sector read instead of the table + list walk.

// reverse complement of a k-mer value (bases LSB first, 2 bits each, complement = 3 - base = ~base)
__device__ __forceinline__ u32 rc_kmer(u32 x, int k) {
    u32 r = __brev(x);
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
    return (~r) >> (32 - 2 * k);
}
__device__ __forceinline__ u64 rc_kmer(u64 x, int k) {
    u64 r = __brevll(x);
    r = ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
    return (~r) >> (64 - 2 * k);
}
// One filter word serves both strands: the WORD is chosen by the canonical k-mer min(x, rc(x)), the three bits inside it
// by the k-mer as it stands.  A probe of B position p then answers "is x in A?" (forward strand) and "is rc(x) in A?"
// (complement strand, position L-k-p) with a single 32-byte sector instead of two.
__device__ __forceinline__ u32 kbit_mask(u32 f) {
    const u32 h = f * 0x85EBCA6Bu;
    return (1u << (h >> 27)) | (1u << ((h >> 22) & 31u)) | (1u << ((h >> 17) & 31u));
}

template <class IDX>
__device__ __forceinline__ void kmer_bitmap_body(IDX ta, int64_t na, int k, int kshift, u32 *__restrict__ bits) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= na) return;
    const typename IDX::key_t km = ta.key(i);
    if (IDX::invalid(km)) return;
    if (i > 0 && ta.key(i - 1) == km) return;                    // one atomic per distinct k-mer
    const typename IDX::key_t y = rc_kmer(km, k), c = km < y ? km : y;
    atomicOr(&bits[(IDX::fold(c) * 0x9E3779B1u) >> kshift], kbit_mask(IDX::fold(km)));
}
__global__ void __launch_bounds__(256) k_kmer_bitmap(const u64 *__restrict__ ta, int64_t na, int k, int kshift, u32 *__restrict__ bits) {
    kmer_bitmap_body(Idx32{ta}, na, k, kshift, bits);
}
__global__ void __launch_bounds__(256) k_kmer_bitmap_w(const ulonglong2 *__restrict__ ta, int64_t na, int k, int kshift, u32 *__restrict__ bits) {
    kmer_bitmap_body(Idx64{ta}, na, k, kshift, bits);
}

__global__ void __launch_bounds__(256) k_kmer_bitmap_p(const u64 *__restrict__ ta, int pb, int64_t na, int k, int kshift, u32 *__restrict__ bits) {
    kmer_bitmap_body(IdxP{ta, pb}, na, k, kshift, bits);
}

struct WordKmers { u64 v; u64 mwin; u32 w2; int p0, L, r; };

template <bool WIDE>
__device__ __forceinline__ WordKmers load_word(const u32 *__restrict__ seq, const u32 *__restrict__ maskbits,
                                               const int64_t *__restrict__ off, const int32_t *__restrict__ len,
                                               const int32_t *__restrict__ c2r, int64_t wi) {
    WordKmers w;
    const int64_t g0 = wi << 4;
    int r = c2r[g0 >> 10];
    while (off[r + 1] <= g0) r++;
    w.r = r; w.p0 = (int)(g0 - off[r]); w.L = len[r];
    // streaming (evict-first) loads: the packed reads pass through once per sweep and must not push the index out of L2
    w.v = ((u64)__ldcs(seq + wi + 1) << 32) | __ldcs(seq + wi);
    w.w2 = WIDE ? __ldcs(seq + wi + 2) : 0u;
    w.mwin = 0;
    if (maskbits) { int64_t mw = g0 >> 5; w.mwin = (((u64)maskbits[mw + 1] << 32) | maskbits[mw]) >> (g0 & 31); }
    return w;
}

template <class IDX>
__device__ __forceinline__ typename IDX::key_t kmer_at(const WordKmers &w, int jj, u64 kmask) {
    if (IDX::wide) return (typename IDX::key_t)wide_kmer(w.v, w.w2, jj, kmask);
    return (typename IDX::key_t)((w.v >> (2 * jj)) & kmask);
}

template <class IDX>
__device__ __forceinline__ void a_range_fwd(IDX ta, const u32 *__restrict__ tbl, int sh, typename IDX::key_t km, int tcap, u32 &s, u32 &c) {
    const u32 q = (u32)((u64)km >> sh);
    u32 a = tbl[q], hi = tbl[q + 1], b = hi;
    while (a < b) { u32 m = (a + b) >> 1; if (ta.key(m) < km) a = m + 1; else b = m; }
    s = a; c = 0;
    while (a < hi && ta.key(a) == km) { a++; if (++c > (u32)tcap) { c = 0; break; } }
}

// Count pass, BOTH strands in one sweep over the forward words.  Phase 1: every thread probes the k-mer filter once per
// position (16 loads in flight); the word answers for the k-mer x (forward strand, position p) and for rc(x) (complement
// strand, position L-k-p).  Phase 2: the candidates (about 1 in 8 probes) are dealt out evenly over the lanes of the warp,
// so the index walks -- dependent, mostly DRAM-missing loads -- run with all lanes busy.  Forward-strand counts collect in
// shared memory (the word belongs to this warp), complement-strand counts go to their word of the mirrored numbering with
// global atomics (wcnt / hitmask of strand 1 start zeroed).  ncu before: one launch per strand, 340 M L2 sectors each,
// L2-sector bound.
template <class IDX>
__device__ __forceinline__ void lookup_count_body(const u32 *__restrict__ seq, const u32 *__restrict__ maskbits,
                                                  const int64_t *__restrict__ off, const int32_t *__restrict__ len,
                                                  const int32_t *__restrict__ c2r, int64_t nwords, int k,
                                                  IDX ta, const u32 *__restrict__ tbl, int sh, int tcap,
                                                  const u32 *__restrict__ kbits, int kshift, const JoinGeom &G, u32 *__restrict__ wcnt,
                                                  unsigned short *__restrict__ hitmask,
                                                  u32 *__restrict__ wlist, u32 *__restrict__ nlist, int64_t w0, int64_t w1) {
    // words [w0, w1) of the block's nwords: the pass can follow a read block that is still arriving chunk by chunk
    typedef typename IDX::key_t key_t;
    __shared__ u32 s_total[8][32], s_hm[8][32];
    const u32 FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t wi = w0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool inrange = wi < w1;
    const u64 kmask = (1ull << (2 * k)) - 1ull, mk = (1ull << k) - 1ull;
    const bool restricted = G.self || G.a_group;
    WordKmers w; w.v = 0; w.mwin = 0; w.w2 = 0; w.p0 = 0; w.L = 0; w.r = 0;
    u32 present = 0;                                  // bit jj: forward strand, bit 16 + jj: complement strand
    if (inrange) {
        w = load_word<IDX::wide>(seq, maskbits, off, len, c2r, wi);
#pragma unroll
        for (int jj = 0; jj < 16; jj++)
            if ((w.p0 + jj + k <= w.L) && (((w.mwin >> jj) & mk) == 0)) {
                const key_t x = kmer_at<IDX>(w, jj, kmask), y = rc_kmer(x, k), c = x < y ? x : y;
                const u32 W = kbits[(IDX::fold(c) * 0x9E3779B1u) >> kshift];
                const u32 mx = kbit_mask(IDX::fold(x)), my = kbit_mask(IDX::fold(y));
                if ((W & mx) == mx) present |= 1u << jj;
                if ((W & my) == my) present |= 1u << (16 + jj);
            }
    }
    s_total[warp][lane] = 0; s_hm[warp][lane] = 0;
    const int cnt = __popc(present);
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
    const int E = inc - cnt, T = __shfl_sync(FULL, inc, 31);
    __syncwarp();
    u32 *wcnt1 = wcnt + nwords; u32 *hm1 = reinterpret_cast<u32 *>(hitmask + nwords);       // strand 1 arrays (nwords is even)
    for (int base = 0; base < T; base += 32) {
        const int q = base + lane;
        int owner = 0;                                   // last lane whose exclusive prefix is <= q
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
            const int mid = owner + step;
            const int Em = __shfl_sync(FULL, E, mid & 31);
            if (mid < 32 && Em <= q) owner = mid;
        }
        const u32 pm = __shfl_sync(FULL, present, owner);
        const int Eo = __shfl_sync(FULL, E, owner);
        WordKmers o;
        o.v = ((u64)__shfl_sync(FULL, (u32)(w.v >> 32), owner) << 32) | __shfl_sync(FULL, (u32)w.v, owner);
        o.w2 = __shfl_sync(FULL, w.w2, owner);
        o.r = __shfl_sync(FULL, w.r, owner);
        o.p0 = __shfl_sync(FULL, w.p0, owner);
        o.L = __shfl_sync(FULL, w.L, owner);
        if (q < T) {
            const int b = __fns(pm, 0, q - Eo + 1), jj = b & 15, strand = b >> 4;
            key_t km = kmer_at<IDX>(o, jj, kmask);
            if (strand) km = rc_kmer(km, k);
            u32 s, c; a_range_fwd(ta, tbl, sh, km, tcap, s, c);
            u32 add = c;
            if (restricted) {                     // self pairs / pairs across pile-ups are never emitted
                add = 0;
                for (u32 x = 0; x < c; x++) {
                    const int ar = read_of(G.a_c2r, G.a_off, (int64_t)ta.pos(s + x));
                    add += pair_ok(G, ar, o.r) ? 1u : 0u;
                }
            }
            if (add) {
                if (!strand) { atomicAdd(&s_total[warp][owner], add); atomicOr(&s_hm[warp][owner], 1u << jj); }
                else {
                    // the same k-mer seen from the complement strand: position L - k - p of the mirrored read
                    const int64_t g1 = off[o.r] + (o.L - k - (o.p0 + jj));
                    const int64_t w1 = g1 >> 4;
                    const u32 old = atomicAdd(&wcnt1[w1], add);
                    atomicOr(&hm1[w1 >> 1], (1u << (int)(g1 & 15)) << ((w1 & 1) ? 16 : 0));
                    if (old == 0) wlist[nwords + atomicAdd(nlist + 1, 1u)] = (u32)w1;
                }
            }
        }
    }
    __syncwarp();
    const u32 total = s_total[warp][lane], hm = s_hm[warp][lane];
    if (inrange) { __stcs(wcnt + wi, total); hitmask[wi] = (unsigned short)hm; }
    if (hm) {
        // words with hits go on a compact list (any order: their output offsets come from the scan of wcnt), so the
        // emit pass runs dense instead of with 1 live thread in 9 (ncu: 3.7 active threads per warp before)
        const u32 act = __activemask();
        const int leader = __ffs(act) - 1;
        u32 base = 0;
        if (lane == leader) base = atomicAdd(nlist, (u32)__popc(act));
        base = __shfl_sync(act, base, leader);
        wlist[base + __popc(act & ((1u << lane) - 1u))] = (u32)wi;
    }
}

__global__ void __launch_bounds__(256) k_lookup_count(const u32 *__restrict__ seq, const u32 *__restrict__ maskbits,
                                                      const int64_t *__restrict__ off, const int32_t *__restrict__ len,
                                                      const int32_t *__restrict__ c2r, int64_t nwords, int k,
                                                      const u64 *__restrict__ ta, const u32 *__restrict__ tbl, int sh, int tcap,
                                                      const u32 *__restrict__ kbits, int kshift, JoinGeom G, u32 *__restrict__ wcnt,
                                                      unsigned short *__restrict__ hitmask,
                                                      u32 *__restrict__ wlist, u32 *__restrict__ nlist, int64_t w0, int64_t w1) {
    lookup_count_body(seq, maskbits, off, len, c2r, nwords, k, Idx32{ta}, tbl, sh, tcap, kbits, kshift, G, wcnt, hitmask, wlist, nlist, w0, w1);
}
__global__ void __launch_bounds__(256) k_lookup_count_w(const u32 *__restrict__ seq, const u32 *__restrict__ maskbits,
                                                        const int64_t *__restrict__ off, const int32_t *__restrict__ len,
                                                        const int32_t *__restrict__ c2r, int64_t nwords, int k,
                                                        const ulonglong2 *__restrict__ ta, const u32 *__restrict__ tbl, int sh, int tcap,
                                                        const u32 *__restrict__ kbits, int kshift, JoinGeom G, u32 *__restrict__ wcnt,
                                                        unsigned short *__restrict__ hitmask,
                                                        u32 *__restrict__ wlist, u32 *__restrict__ nlist, int64_t w0, int64_t w1) {
    lookup_count_body(seq, maskbits, off, len, c2r, nwords, k, Idx64{ta}, tbl, sh, tcap, kbits, kshift, G, wcnt, hitmask, wlist, nlist, w0, w1);
}

__global__ void __launch_bounds__(256) k_lookup_count_p(const u32 *__restrict__ seq, const u32 *__restrict__ maskbits,
                                                        const int64_t *__restrict__ off, const int32_t *__restrict__ len,
                                                        const int32_t *__restrict__ c2r, int64_t nwords, int k,
                                                        const u64 *__restrict__ ta, int pb, const u32 *__restrict__ tbl, int sh, int tcap,
                                                        const u32 *__restrict__ kbits, int kshift, JoinGeom G, u32 *__restrict__ wcnt,
                                                        unsigned short *__restrict__ hitmask,
                                                        u32 *__restrict__ wlist, u32 *__restrict__ nlist, int64_t w0, int64_t w1) {
    lookup_count_body(seq, maskbits, off, len, c2r, nwords, k, IdxP{ta, pb}, tbl, sh, tcap, kbits, kshift, G, wcnt, hitmask, wlist, nlist, w0, w1);
}

// Emit pass over the compact list of words with hits.  Like the count pass it deals the (word, position) units of a warp's 32
// words out evenly over the lanes (ncu before: 8 live threads per warp -- one thread walked all hit positions of its word while
// the lanes of hit-free positions idled), walks the index once per unit, and places the unit's hits behind those of the earlier
// positions of the same word by a segmented warp scan (units of a word are consecutive), so a word's hits stay ordered by bpos.
template <class IDX>
__device__ __forceinline__ void lookup_emit_body(const u32 *__restrict__ seq, const u32 *__restrict__ maskbits,
                                                 const int64_t *__restrict__ off, const int32_t *__restrict__ len,
                                                 const int32_t *__restrict__ c2r, int64_t nwords, int k,
                                                 IDX ta, const u32 *__restrict__ tbl, int sh, int tcap,
                                                 const unsigned short *__restrict__ hitmask, const u32 *__restrict__ wcnt,
                                                 const int64_t *__restrict__ woff, int strand,
                                                 const JoinGeom &G, ulonglong2 *__restrict__ hits, const u32 *__restrict__ wlist) {
    __shared__ u32 s_done[8][32];                         // hits of each lane's word already placed by earlier batches
    const u32 FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t li = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool inrange = li < nwords;                     // nwords = length of the list of words with hits
    const u64 kmask = (1ull << (2 * k)) - 1ull;
    const bool restricted = G.self || G.a_group;
    WordKmers w; w.v = 0; w.mwin = 0; w.w2 = 0; w.p0 = 0; w.L = 0; w.r = 0;
    u32 present = 0; long long wo = 0;
    if (inrange) {
        const int64_t wi = wlist[li];
        w = load_word<IDX::wide>(seq, maskbits, off, len, c2r, wi);
        present = hitmask[wi];                            // from the count pass: no filter probes, no fruitless lookups
        wo = __ldcs((const long long *)woff + wi);
    }
    s_done[warp][lane] = 0;
    const int cnt = __popc(present);
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
    const int E = inc - cnt, T = __shfl_sync(FULL, inc, 31);
    __syncwarp();
    for (int base = 0; base < T; base += 32) {
        const int q = base + lane;
        int owner = 0;                                    // last lane whose exclusive prefix is <= q
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
            const int mid = owner + step;
            const int Em = __shfl_sync(FULL, E, mid & 31);
            if (mid < 32 && Em <= q) owner = mid;
        }
        const u32 pm = __shfl_sync(FULL, present, owner);
        const int Eo = __shfl_sync(FULL, E, owner);
        WordKmers o;
        o.v = ((u64)__shfl_sync(FULL, (u32)(w.v >> 32), owner) << 32) | __shfl_sync(FULL, (u32)w.v, owner);
        o.w2 = __shfl_sync(FULL, w.w2, owner);
        o.r = __shfl_sync(FULL, w.r, owner);
        o.p0 = __shfl_sync(FULL, w.p0, owner);
        const long long owo = ((long long)__shfl_sync(FULL, (int)(wo >> 32), owner) << 32) | (u32)__shfl_sync(FULL, (int)(u32)wo, owner);
        u32 s = 0, c = 0, add = 0; int jj = 0;
        if (q < T) {
            jj = __fns(pm, 0, q - Eo + 1);
            const typename IDX::key_t km = kmer_at<IDX>(o, jj, kmask);
            a_range_fwd(ta, tbl, sh, km, tcap, s, c);
            add = c;
            if (restricted) {
                add = 0;
                for (u32 x = 0; x < c; x++) add += pair_ok(G, read_of(G.a_c2r, G.a_off, (int64_t)ta.pos(s + x)), o.r) ? 1u : 0u;
            }
        }
        // segmented exclusive scan of `add` over the units of one word
        u32 sc = add;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(FULL, sc, d); if (lane >= d) sc += t; }
        const u32 excl = sc - add;
        const int head_lane = Eo > base ? Eo - base : 0;  // first unit of this word inside the batch
        const u32 head = __shfl_sync(FULL, excl, head_lane);
        if (q < T) {
            int64_t at = owo + s_done[warp][owner] + (excl - head);
            const u64 bs = (u64)strand * G.nb_reads + o.r;
            const int bpos = o.p0 + jj;
            for (u32 x = 0; x < c; x++) {
                const int64_t ga = (int64_t)ta.pos(s + x);
                const int ar = read_of(G.a_c2r, G.a_off, ga);
                if (!pair_ok(G, ar, o.r)) continue;
                const int apos = (int)(ga - G.a_off[ar]);
                const u64 gd = (u64)(G.a_dbase[ar] + apos - bpos + G.maxlb);
                __stcs(reinterpret_cast<ulonglong2 *>(hits + at), make_ulonglong2((bs << G.gdbits) | gd, (u64)(u32)apos | ((u64)(u32)bpos << 32)));
                at++;
            }
        }
        __syncwarp();
        // the last unit of a word inside this batch books the word's progress for the next batch
        const int cnt_o = __popc(pm);
        if (q < T && (lane == 31 || q == T - 1 || q == Eo + cnt_o - 1)) s_done[warp][owner] += sc - head;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256) k_lookup_emit(const u32 *__restrict__ seq, const u32 *__restrict__ maskbits,
                                                     const int64_t *__restrict__ off, const int32_t *__restrict__ len,
                                                     const int32_t *__restrict__ c2r, int64_t nwords, int k,
                                                     const u64 *__restrict__ ta, const u32 *__restrict__ tbl, int sh, int tcap,
                                                     const unsigned short *__restrict__ hitmask, const u32 *__restrict__ wcnt,
                                                     const int64_t *__restrict__ woff, int strand,
                                                     JoinGeom G, ulonglong2 *__restrict__ hits, const u32 *__restrict__ wlist) {
    lookup_emit_body(seq, maskbits, off, len, c2r, nwords, k, Idx32{ta}, tbl, sh, tcap, hitmask, wcnt, woff, strand, G, hits, wlist);
}
__global__ void __launch_bounds__(256) k_lookup_emit_w(const u32 *__restrict__ seq, const u32 *__restrict__ maskbits,
                                                       const int64_t *__restrict__ off, const int32_t *__restrict__ len,
                                                       const int32_t *__restrict__ c2r, int64_t nwords, int k,
                                                       const ulonglong2 *__restrict__ ta, const u32 *__restrict__ tbl, int sh, int tcap,
                                                       const unsigned short *__restrict__ hitmask, const u32 *__restrict__ wcnt,
                                                       const int64_t *__restrict__ woff, int strand,
                                                       JoinGeom G, ulonglong2 *__restrict__ hits, const u32 *__restrict__ wlist) {
    lookup_emit_body(seq, maskbits, off, len, c2r, nwords, k, Idx64{ta}, tbl, sh, tcap, hitmask, wcnt, woff, strand, G, hits, wlist);
}

__global__ void __launch_bounds__(256, 6) k_lookup_emit_p(const u32 *__restrict__ seq, const u32 *__restrict__ maskbits,
                                                       const int64_t *__restrict__ off, const int32_t *__restrict__ len,
                                                       const int32_t *__restrict__ c2r, int64_t nwords, int k,
                                                       const u64 *__restrict__ ta, int pb, const u32 *__restrict__ tbl, int sh, int tcap,
                                                       const unsigned short *__restrict__ hitmask, const u32 *__restrict__ wcnt,
                                                       const int64_t *__restrict__ woff, int strand,
                                                       JoinGeom G, ulonglong2 *__restrict__ hits, const u32 *__restrict__ wlist) {
    lookup_emit_body(seq, maskbits, off, len, c2r, nwords, k, IdxP{ta, pb}, tbl, sh, tcap, hitmask, wcnt, woff, strand, G, hits, wlist);
}

// ------------------------------------------------------------------------- K4: band filter

// per hit: covered-base contribution and band-start flag
__global__ void __launch_bounds__(256) k_hit_cover(const ulonglong2 *__restrict__ hits, int64_t n, int k, int w,
                                                   int32_t *__restrict__ cov, int32_t *__restrict__ bflag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ulonglong2 h = hits[i];
    int c = k, f = 1;
    if (i > 0) {
        ulonglong2 p = hits[i - 1];
        int da = (int)(u32)h.y - (int)(u32)p.y;
        if (p.x == h.x && da < k) c = da;
        f = (p.x >> w) != (h.x >> w);
    }
    cov[i] = c; bflag[i] = f;
}

// band table: for every band start i (bflag), bfirst[bidx] = i, bkey[bidx] = key >> w
__global__ void __launch_bounds__(256) k_band_table(const ulonglong2 *__restrict__ hits, int64_t n, int w,
                                                    const int32_t *__restrict__ bflag, const int32_t *__restrict__ bidx,
                                                    int32_t *__restrict__ bfirst, u64 *__restrict__ bkey, int32_t nbands) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (bflag[i]) { int b = bidx[i]; bfirst[b] = (int32_t)i; bkey[b] = hits[i].x >> w; }
    if (i == n - 1) bfirst[nbands] = (int32_t)n;
}

// pass(q) = covered bases of the band pair (q, q+1) >= h.
// hot[q] = pass(q) || (adjacent(q-1,q) && pass(q-1)); cstart[q] = hot[q] && !(hot[q-1] && adjacent(q-1,q))
__global__ void __launch_bounds__(256) k_band_hot(const int32_t *__restrict__ bfirst, const u64 *__restrict__ bkey,
                                                  const int32_t *__restrict__ covsum, const int32_t *__restrict__ d_total_cov,
                                                  int64_t nhits, int32_t nbands, int h, uint8_t *__restrict__ hot, int32_t *__restrict__ cstart) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nbands) return;
    const int32_t total_cov = *d_total_cov;
    auto cs = [&](int32_t i) { return i >= nhits ? total_cov : covsum[i]; };
    auto pass = [&](int x) -> bool {
        int sc = cs(bfirst[x + 1]) - cs(bfirst[x]);
        const bool adj = (x + 1 < nbands) && (bkey[x + 1] == bkey[x] + 1);
        if (adj) sc += cs(bfirst[x + 2]) - cs(bfirst[x + 1]);
        return sc >= h;
    };
    auto is_hot = [&](int x) -> bool {
        if (pass(x)) return true;
        return x > 0 && bkey[x - 1] + 1 == bkey[x] && pass(x - 1);
    };
    bool hq = is_hot(q);
    hot[q] = hq;
    bool cont = hq && q > 0 && bkey[q - 1] + 1 == bkey[q] && is_hot(q - 1);
    cstart[q] = hq && !cont;
}

// k_hit_cover fused into the scan that consumes it: the load computes (band flag, covered bases) of hit i from hits i - 1
// and i, the scan runs over the packed pair (flag << 32 | cover), the store splits the prefix into bidx / covsum.  The
// per-hit cover never exists in memory.  Totals: packed (bands << 32 | cover) through *d_total.
namespace {
struct CoverScan {
    const ulonglong2 *hits; int k, w; int32_t *bflag, *bidx, *covsum;
    __device__ __forceinline__ unsigned long long load(size_t i) const {
        const ulonglong2 h = hits[i];
        int c = k, f = 1;
        if (i > 0) {
            const ulonglong2 p = hits[i - 1];
            const int da = (int)(u32)h.y - (int)(u32)p.y;
            if (p.x == h.x && da < k) c = da;
            f = (p.x >> w) != (h.x >> w);
        }
        bflag[i] = f;
        return ((unsigned long long)f << 32) | (unsigned long long)(u32)c;
    }
    __device__ __forceinline__ void store(size_t i, unsigned long long v) const { bidx[i] = (int32_t)(v >> 32); covsum[i] = (int32_t)(u32)v; }
};
}
void launch_cover_scan(const ulonglong2 *hits, int64_t n, int k, int w, int32_t *bflag, int32_t *bidx, int32_t *covsum,
                       unsigned long long *d_total, cudaStream_t s) {
    scan_chained<unsigned long long>(CoverScan{hits, k, w, bflag, bidx, covsum}, (size_t)n, d_total, s);
}

// one thread per cluster start: walk to the end of the run, pick the median hit as seed
__global__ void __launch_bounds__(256) k_seeds(const ulonglong2 *__restrict__ hits, const int32_t *__restrict__ bfirst,
                                               const u64 *__restrict__ bkey, const uint8_t *__restrict__ hot,
                                               const int32_t *__restrict__ cstart, const int32_t *__restrict__ cidx,
                                               int32_t nbands, SeedGeom G, Seed *__restrict__ seeds,
                                               uint8_t *__restrict__ consumed) {
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nbands || !cstart[q]) return;
    int e = q;
    while (e + 1 < nbands && hot[e + 1] && bkey[e + 1] == bkey[e] + 1) e++;
    int32_t f = bfirst[q], l = bfirst[e + 1];
    int32_t m = f + (l - f - 1) / 2;
    ulonglong2 h = hits[m];
    u64 gd = h.x & ((1ull << G.gdbits) - 1ull);
    u32 bs = (u32)(h.x >> G.gdbits);
    // aread = last r with dbase[r] <= gd
    int lo = 0, hi = G.na;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if ((u64)G.a_dbase[mid] <= gd) lo = mid; else hi = mid; }
    Seed s; s.a = lo; s.bs = (int32_t)bs; s.apos = (int32_t)(u32)h.y; s.bpos = (int32_t)(h.y >> 32);
    seeds[cidx[q]] = s;
    consumed[m] = 1;
}

}  // namespace dn
