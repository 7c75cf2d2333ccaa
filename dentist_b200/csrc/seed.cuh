// seed.cuh -- shared structs between the seeding and extension stages.
#pragma once
#include <stdlib.h>
#include "engine.cuh"

namespace dn {

// k-mer presence filter (blocked Bloom, 3 bits per k-mer): >= 12 bits per indexed position, 2^27 bits (16 MB, pinned in
// L2) at least, 2^32 at most (a 100 Mbp reference block gets 2^31 bits = 256 MB; a 16 MB filter would be ~90 % full there)
inline int kbits_log2_for(int64_t n_index) {
    static const int per = getenv("DN_KBITS_PER") ? atoi(getenv("DN_KBITS_PER")) : 12;      // filter bits per indexed position (at least)
    int b = 27; while (b < 32 && (1ll << b) < (long long)per * n_index) b++; return b;
}

struct Seed { int32_t a, bs, apos, bpos; };

// geometry handed to the join: how a (aread, apos, bread, strand, bpos) hit becomes a sort key
struct JoinGeom {
    const int32_t *a_c2r; const int64_t *a_off; const int64_t *a_dbase;
    const int32_t *b_c2r; const int64_t *b_off;
    int64_t nbp;          // padded bases of B: payloads >= nbp are the complement strand
    int64_t maxlb;        // max B read length rounded up to the band width
    int gdbits, keybits, self;
    const int32_t *a_group, *b_group;   // optional: pairs with different group ids are dropped
    int nb_reads;                       // hit/seed/candidate `bs` = strand * nb_reads + bread (strand-major)
};
struct SeedGeom { const int64_t *a_dbase; int na; int gdbits; };

// one candidate local alignment (device + host layout)
struct Cand {
    int32_t a, bs, ab, ae, bb, be, diffs, nt, dmin, dmax;
    int64_t toff;          // offset of its joined trace (uint16 pairs -> 2*nt elements) in the round's trace buffer
};

struct ExtOut { int32_t i_end, j_end, d_end, ntiles; };

void emit_tuples(const DevBlock &B, bool rc, int k, u32 payload_base, u64 *out, cudaStream_t s);
// bucket.cu: the ordered 8-byte index + its prefix table (C + 1, C = nq + 3 words) without a full sort; false = take the radix path
bool build_index_u64(const u64 *tuples, u64 *out, int64_t n, int key_shift, int sh, u32 nq, u32 *C, cudaStream_t s);
int64_t emit_tuples_wide(const DevBlock &B, int k, void *out, int pb, cudaStream_t s);   // k = 16..31: tuples of the valid forward positions (pb = 0: 16-byte {kmer, position}; pb > 0: 8-byte kmer << pb | position); returns their number

// kernels defined in seed.cu
__global__ void k_prefix_table(const u64 *ta, int64_t na, int sh, u32 nq, u32 *tbl);
__global__ void k_join_count(const u64 *ta, const u32 *tbl, int sh, const u64 *tb, int64_t nb, int tcap, u32 *cnt, u32 *start);
__global__ void k_join_emit(const u64 *ta, const u64 *tb, int64_t nb, const u32 *cnt, const u32 *start, const int64_t *hoff,
                            JoinGeom G, ulonglong2 *hits, unsigned long long *ninvalid);
__global__ void k_lookup_count(const u32 *seq, const u32 *maskbits, const int64_t *off, const int32_t *len, const int32_t *c2r,
                               int64_t nwords, int k, const u64 *ta, const u32 *tbl, int sh, int tcap, const u32 *kbits, int kshift, JoinGeom G,
                               u32 *wcnt, unsigned short *hitmask, u32 *wlist, u32 *nlist, int64_t w0, int64_t w1);
__global__ void k_kmer_bitmap(const u64 *ta, int64_t na, int k, int kshift, u32 *bits);
__global__ void k_lookup_emit(const u32 *seq, const u32 *maskbits, const int64_t *off, const int32_t *len, const int32_t *c2r,
                              int64_t nwords, int k, const u64 *ta, const u32 *tbl, int sh, int tcap, const unsigned short *hitmask,
                              const u32 *wcnt, const int64_t *woff, int strand, JoinGeom G, ulonglong2 *hits, const u32 *wlist);
// k = 16..31 variants over 16-byte {kmer, position} index entries
__global__ void k_prefix_table_w(const ulonglong2 *ta, int64_t na, int sh, u32 nq, u32 *tbl);
__global__ void k_kmer_bitmap_w(const ulonglong2 *ta, int64_t na, int k, int kshift, u32 *bits);
__global__ void k_lookup_count_w(const u32 *seq, const u32 *maskbits, const int64_t *off, const int32_t *len, const int32_t *c2r,
                                 int64_t nwords, int k, const ulonglong2 *ta, const u32 *tbl, int sh, int tcap, const u32 *kbits, int kshift, JoinGeom G,
                                 u32 *wcnt, unsigned short *hitmask, u32 *wlist, u32 *nlist, int64_t w0, int64_t w1);
__global__ void k_lookup_emit_w(const u32 *seq, const u32 *maskbits, const int64_t *off, const int32_t *len, const int32_t *c2r,
                                int64_t nwords, int k, const ulonglong2 *ta, const u32 *tbl, int sh, int tcap, const unsigned short *hitmask,
                                const u32 *wcnt, const int64_t *woff, int strand, JoinGeom G, ulonglong2 *hits, const u32 *wlist);
__global__ void k_hit_cover(const ulonglong2 *hits, int64_t n, int k, int w, int32_t *cov, int32_t *bflag);
__global__ void k_band_table(const ulonglong2 *hits, int64_t n, int w, const int32_t *bflag, const int32_t *bidx,
                             int32_t *bfirst, u64 *bkey, int32_t nbands);
__global__ void k_band_hot(const int32_t *bfirst, const u64 *bkey, const int32_t *covsum, const int32_t *d_total_cov, int64_t nhits,
                           int32_t nbands, int h, uint8_t *hot, int32_t *cstart);
// hit cover + band flags fused into their prefix sums (one launch); *d_total = bands << 32 | total cover
void launch_cover_scan(const ulonglong2 *hits, int64_t n, int k, int w, int32_t *bflag, int32_t *bidx, int32_t *covsum,
                       unsigned long long *d_total, cudaStream_t s);
__global__ void k_seeds(const ulonglong2 *hits, const int32_t *bfirst, const u64 *bkey, const uint8_t *hot,
                        const int32_t *cstart, const int32_t *cidx, int32_t nbands, SeedGeom G, Seed *seeds, uint8_t *consumed);

// k = 16..31 with 2k + pb <= 64: 8-byte packed index entries (kmer << pb | position)
__global__ void k_prefix_table_p(const u64 *ta, int pb, int64_t na, int sh, u32 nq, u32 *tbl);
__global__ void k_kmer_bitmap_p(const u64 *ta, int pb, int64_t na, int k, int kshift, u32 *bits);
__global__ void k_lookup_count_p(const u32 *seq, const u32 *maskbits, const int64_t *off, const int32_t *len, const int32_t *c2r,
                                 int64_t nwords, int k, const u64 *ta, int pb, const u32 *tbl, int sh, int tcap, const u32 *kbits, int kshift, JoinGeom G,
                                 u32 *wcnt, unsigned short *hitmask, u32 *wlist, u32 *nlist, int64_t w0, int64_t w1);
__global__ void k_lookup_emit_p(const u32 *seq, const u32 *maskbits, const int64_t *off, const int32_t *len, const int32_t *c2r,
                                int64_t nwords, int k, const u64 *ta, int pb, const u32 *tbl, int sh, int tcap, const unsigned short *hitmask,
                                const u32 *wcnt, const int64_t *woff, int strand, JoinGeom G, ulonglong2 *hits, const u32 *wlist);

// segmented hit sort (segsort.cu)
void launch_seg_offsets(const int64_t *woff, const int64_t *b_off, int nr, int64_t nwB, const int64_t *dH, int64_t *seg_beg, int32_t *seg_len,
                        int minlen, const int *caps, int32_t *lists, u32 *counts, cudaStream_t s);
void launch_segsort(ulonglong2 *hits, const int64_t *seg_beg, const int32_t *seg_len, const int32_t *seglist, int nseg, int cap, int gdbits,
                    int aposbits, cudaStream_t s);

void launch_segsort_radix(const ulonglong2 *in, ulonglong2 *out, const int64_t *seg_beg, const int32_t *seg_len, const int32_t *seglist,
                          int nseg, int cap, int gdbits, cudaStream_t s);

// extension stage (extend.cu)
struct ExtGeom {
    const u32 *a_fwd, *a_rc, *b_fwd, *b_rc;
    const int64_t *a_off, *b_off; const int32_t *a_len, *b_len;
    int ts, cdiff, xdrop, wmax, poolmul; u32 ts_magic;
    int nb_reads;
    u32 a_words, b_words;           // packed words allocated per strand array (bounds of the staged bulk copies)
    int64_t tile_stride;            // > 0: task t owns tiles [t * tile_stride, (t + 1) * tile_stride) and no offset array exists
};
void launch_task_caps(const Seed *seeds, int nseeds, ExtGeom G, u32 *caps, cudaStream_t s);
void launch_extend(const Seed *seeds, int nseeds, ExtGeom G, const int64_t *tile_off, int2 *tiles, ExtOut *outs,
                   int4 *pool, int64_t pool_stride, int nwarps_total, int *counter, const int *order, cudaStream_t s);
void launch_task_order(const Seed *seeds, int nseeds, ExtGeom G, int *scratch /* 128 ints */, int *order /* 2 * nseeds */, cudaStream_t s);
void launch_combine(const Seed *seeds, int nseeds, ExtGeom G, int minlen, const int64_t *tile_off, const int2 *tiles,
                    const ExtOut *outs, Cand *cand_all, int32_t *valid, u32 *ntl, cudaStream_t s);
void launch_write_traces(const Seed *seeds, int nseeds, ExtGeom G, const int64_t *tile_off, const int2 *tiles,
                         const ExtOut *outs, const Cand *cand_all, const int32_t *valid, const int32_t *vidx,
                         const int64_t *toff, Cand *cand_out, uint16_t *trace, cudaStream_t s);
void launch_retire(const ulonglong2 *hits, int64_t n, const uint8_t *consumed, const Cand *rc, const int32_t *d_nrc, SeedGeom G, int w,
                   const int32_t *bflag, const int32_t *bidx, const uint8_t *hot, const int32_t *bfirst, int32_t nbands, int2 *crange,
                   int32_t *keep, cudaStream_t s);
void launch_compact_hits(const ulonglong2 *hits, int64_t n, const int32_t *keep, const int32_t *kidx, ulonglong2 *out, cudaStream_t s);
struct FinalBits { int na, nb, nra, nrb, nb_reads; };       // bits of: A coordinate, B coordinate, A read id, B read id
struct FinalGeom { const uint16_t *round_trace[16]; int32_t round_beg[17]; int nrounds; };
void launch_final_setkey(const Cand *c, const uint8_t *drop, ulonglong2 *items, int n, int field, FinalBits fb,
                         unsigned long long *ndrop, cudaStream_t s);
// ncand items, of which the first ncand - ctr[0] are kept (ctr[0] = dropped duplicates, still on the device)
void launch_final_records(const Cand *c, const ulonglong2 *items, int ncand, int nb_reads, dn_las_record *rec, u32 *tl, unsigned long long *ctr, cudaStream_t s);
void launch_final_traces(const Cand *c, const ulonglong2 *items, int ncand, const unsigned long long *ctr, const int64_t *toff, FinalGeom G, uint16_t *out, cudaStream_t s);
void launch_final_fixup(const Cand *c, ulonglong2 *items, int n, cudaStream_t s);
void launch_dedupe(const Cand *cands, int ncand, const int32_t *round_beg, int nrounds, uint8_t *drop, cudaStream_t s);

}  // namespace dn
