/*
 * dentist_b200.h -- C ABI of the B200-native alignment + consensus engine that replaces the
 * external-tool calls in DENTIST's `source/dentist/dazzler.d` (the drop-in boundary).
 *
 * Every entry point returns 0 on success and non-zero on error; the message is available from
 * dn_last_error() (thread-local).  On the D side a non-zero return is turned into the
 * `DazzlerCommandException` that `executeWrapper` throws today (dazzler.d:6551-6591), which
 * processPileUps already maps to "skip this pile-up" (processPileUps/package.d:351-373).
 * Entry points are re-entrant: DENTIST calls them concurrently from std.parallelism workers
 * (processPileUps/package.d:153).  No torch types, plain pointers and sizes only.
 *
 * There is NO CPU fallback: every compute entry point fails with DN_ERR_NO_DEVICE when no CUDA
 * device is usable.
 */
#ifndef DENTIST_B200_H
#define DENTIST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DN_OK 0
#define DN_ERR_INVALID 1
#define DN_ERR_NO_DEVICE 2
#define DN_ERR_CUDA 3
#define DN_ERR_IO 4
#define DN_ERR_EMPTY 5     /* e.g. "empty consensus" -- dazzler.d:4232-4235 */

/* ---- lifecycle ------------------------------------------------------------------------- */

/* Select the CUDA device this process uses (one process per GPU).  `tmpdir` may be NULL.
 * Replaces the start-up tool discovery of `assertExternalToolsAvailable` (commandline.d:339-352). */
int dn_init(int device, const char *tmpdir);
int dn_shutdown(void);
const char *dn_last_error(void);
const char *dn_version(void);
/* Number of CUDA kernels this library has launched so far in this process. */
uint64_t dn_launch_count(void);

/* ---- data model ------------------------------------------------------------------------ */

/* One LAS record, byte-for-byte the 40 bytes DENTIST reads from / writes to a .las file:
 * DazzlerOverlap[8..48)  (dazzler.d:1717-1725, 1988-2016, 2146-2150).  aread/bread are 0-based
 * (DENTIST adds 1, dazzler.d:1731-1734).  flags: COMP 0x1, START 0x4, NEXT 0x8, BEST 0x10,
 * ELIM 0x20 (dazzler.d:1991-1998). */
typedef struct dn_las_record {
    int32_t tlen, diffs, abpos, bbpos, aepos, bepos;
    uint32_t flags;
    int32_t aread, bread;
    int32_t pad_;
} dn_las_record;

#define DN_LAS_COMP 0x1u
#define DN_LAS_START 0x4u
#define DN_LAS_NEXT 0x8u
#define DN_LAS_BEST 0x10u
#define DN_LAS_ELIM 0x20u

/* Sequence block as it sits in a DAZZ_DB (`.bps` + `.idx`) or in memory.
 * format DN_SEQ_BYTES: one base per byte, codes a=0 c=1 g=2 t=3 (what DAZZ_DB loads in memory);
 * format DN_SEQ_BPS:   DAZZ_DB .bps -- 4 bases per byte, first base in the two most significant
 *                      bits, read r starts at byte boff[r] and spans ceil(rlen[r]/4) bytes.
 * mask_*: optional mask track in the reference's track layout (dazzler.d:4943-5052):
 *   mask_anno = nreads+1 int64 BYTE offsets into mask_data, mask_data = int32 (begin,end) pairs. */
#define DN_SEQ_BYTES 0
#define DN_SEQ_BPS 1
typedef struct dn_block_desc {
    int32_t nreads;
    int32_t format;
    const int32_t *rlen;      /* [nreads] */
    const int64_t *boff;      /* [nreads] byte offset of each read in `data` */
    const void *data;
    int64_t data_bytes;
    const int64_t *mask_anno; /* may be NULL */
    const int32_t *mask_data; /* may be NULL */
    const int32_t *group;     /* may be NULL; [nreads] pile id: only reads of equal group are compared (a whole
                                 batch of pile-ups = one block = one launch; package.d:153 runs them one by one) */
} dn_block_desc;

/* Parameters of the local aligner; the subset of daligner/damapper flags DENTIST passes
 * (pileUpAlignmentOptions commandline.d:2886-2902, postConsensusAlignmentOptions :2918-2935,
 * refVsReadsAlignmentOptions :2943-2955) plus the tool defaults DENTIST leaves alone. */
typedef struct dn_align_params {
    int32_t k;          /* -k  k-mer length (4..31; > 15 uses 16-byte tuples) default 14  */
    int32_t w;          /* -w  log2 of the diagonal band width         default 6   */
    int32_t h;          /* -h  bases covered by k-mer hits in a band   default 35  */
    int32_t t;          /* -t  ignore k-mers occurring more often in A default 32  */
    int32_t tspace;     /* -s  trace point spacing                     default 100 */
    int32_t minlen;     /* -l  minimum local alignment length          default 1000*/
    double e;           /* -e  average correlation rate (>= .7)        default .7  */
    int32_t identity;   /* -I  keep read-vs-itself pairs               default 0   */
    int32_t self_block; /* A and B are the same block (daligner X X): skip aread == bread unless -I */
    int32_t rounds;     /* seed/extend rounds                          default 3   */
    int32_t xdrop;      /* extension x-drop                            default 300 */
    int32_t wmax;       /* live diagonals per wave (<= 62)             default 30  */
    int32_t poolmul;    /* trace record pool multiplier                default 64  */
    int32_t join_mode;  /* 0 auto | 1 sort both tuple lists and merge | 2 look B k-mers up in the sorted A index */
} dn_align_params;
void dn_align_params_default(dn_align_params *p);

typedef struct dn_align_stats {
    int64_t tuples_a, tuples_b, hits, seeds, extensions, las, aligned_bases, trace_points;
    int64_t algo_bytes_seed;      /* algorithmic HBM bytes of the seeding stages   */
    int64_t algo_bytes_extend;    /* algorithmic HBM bytes of the wave extension              */
    float ms_seed, ms_extend, ms_total;   /* CUDA-event times on the engine's stream: total, the k_extend launches, and seed = total - extend */
    uint64_t launches;
} dn_align_stats;

typedef struct dn_las_buf {
    int64_t nrec;
    dn_las_record *rec;       /* LAsort order = FlatLocalAlignment.opCmp (base.d:1787-1809)   */
    int64_t *toff;            /* [nrec] offset of each record's trace in `trace` (uint16 units) */
    int64_t ntrace;
    uint16_t *trace;          /* (diffs, bbases) pairs, widened to uint16 like dazzler.d:1816-1834 */
    int32_t tspace;
    dn_align_stats stats;
} dn_las_buf;
void dn_las_free(dn_las_buf *buf);

/* ---- resident blocks + in-memory alignment (no files on the path) -------- */

typedef struct dn_block dn_block;   /* opaque: a sequence block resident in HBM */
/* H2D copy + on-device 2-bit packing / reverse complement.  Replaces daligner's DB load. */
int dn_block_upload(const dn_block_desc *desc, dn_block **out);
/* Cropped pile-up reads without leaving HBM (replaces getCroppedReadAsFasta + buildDbFile, cropper.d:171, 383-421):
 * output read i = bases [begin[i], end[i]) of read[i] of `src` (forward coordinates, as getCroppingSlice
 * cropper.d:503-550 returns them); `group` (optional) = pile-up id per output read. */
int dn_block_crop(const dn_block *src, int32_t n, const int32_t *read, const int32_t *begin, const int32_t *end, const int32_t *group,
                  dn_block **out);
void dn_block_free(dn_block *blk);
int64_t dn_block_bases(const dn_block *blk);
/* Optional: build the k-mer index of a block once (sorted tuples, prefix table, k-mer filter) and keep it with the
 * block; every later dn_align_blocks(blk, B, p) with p->k == k then skips the A-side index build -- one reference
 * block is aligned against many read blocks (Snakefile:1143-1170).  Results are identical with and without it.
 * k <= 0 drops the index; changing the block's seed mask (dn_block_mask_dust) drops it too. */
int dn_block_index(dn_block *blk, int32_t k);

/* All local alignments of block A vs block B (both strands of B).  What `daligner A B` computes
 * for DENTIST (dazzler.d:6131-6140); result = the records of A.B.las in LAsort order. */
int dn_align_blocks(const dn_block *a, const dn_block *b, const dn_align_params *p, dn_las_buf *out);
/* Same, straight from host descriptors (upload + align + download). */
int dn_align_host(const dn_block_desc *a, const dn_block_desc *b, const dn_align_params *p, dn_las_buf *out);

/* LAmerge (Snakefile:863-873, 1173-1200) for device-resident segments: `d_rec` = nrec concatenated 40-byte records,
 * `d_trace` = their uint16 traces concatenated in the same order (both DEVICE pointers, e.g. the receive
 * buffer of the all-gatherv); returns one LAS in LAsort order.  max_*len / n*_reads bound the key widths. */
int dn_las_merge_device(const void *d_rec, int64_t nrec, const void *d_trace, int64_t ntrace, int32_t tspace, int64_t max_alen,
                        int64_t max_blen, int64_t na_reads, int64_t nb_reads, dn_las_buf *out);

/* ---- multi-GPU: one process per GPU, NCCL inside the library ---------------------------------
 * Read blocks shard across the ranks with no data-path collective (the Snakemake fan-out, Snakefile:1143-1170);
 * the per-rank LAS segments are exchanged HBM to HBM in ONE variable-size gather and merged by placement
 * (what LAmerge does through files, Snakefile:1173-1200).  NCCL is bound at run time (libnccl.so.2 / DN_NCCL_LIB). */
#define DN_COMM_ID_BYTES 128
/* Rank 0 creates the communicator id (ncclGetUniqueId) and hands its 128 bytes to the other ranks by whatever the host
 * has (a file in tmpdir, MPI, torch.distributed ...); every rank then joins with dn_comm_init on its own device. */
int dn_comm_get_id(uint8_t *id);
int dn_comm_init(int32_t rank, int32_t world, const uint8_t *id);
int dn_comm_shutdown(void);
/* > 0: the ranks of this node share a page-locked host segment of that many bytes and a gather with root >= 0 downloads
 * the merged LAS in slices, one per rank and PCIe link (the root's result then points into the segment and stays valid
 * until the second following gather call); 0: the root downloads everything itself.  DN_SHM_MB sizes one half (default 384),
 * DN_NO_SHM=1 disables it. */
int64_t dn_comm_shared_segment_bytes(void);
int32_t dn_comm_rank(void);
int32_t dn_comm_size(void);
/* dn_align_blocks on every rank's own read block `b` (B read numbers local to the block) + gather + merge:
 * `bread_offset` = global number of the block's first read (blocks must be dealt out in rank order: a rank's reads
 * all come before the next rank's).  root >= 0: that rank receives the merged A.B.las, the others get an empty LAS with
 * their own statistics; root < 0: every rank receives it. */
int dn_align_blocks_gather(const dn_block *a, const dn_block *b, const dn_align_params *p, int64_t bread_offset, int32_t root, dn_las_buf *out);
/* the same from host descriptors (upload + align + gather + merge + download on the receiving ranks) */
int dn_align_host_gather(const dn_block_desc *a, const dn_block_desc *b, const dn_align_params *p, int64_t bread_offset, int32_t root, dn_las_buf *out);
/* Variable-size all-gather of host byte buffers (e.g. every rank's InsertionDb bytes, commands/mergeInsertions.d):
 * *recv = the ranks' buffers back to back (free with dn_free), counts[r] = size of rank r's. */
int dn_comm_allgatherv(const void *send, int64_t nbytes, void **recv, int64_t *counts /* [world] */);

/* Serialise to the LAS wire format (dazzler.d:1913-2170: int64 novl, int32 tspace, 40-byte records,
 * uint8 traces iff tspace <= 125). */
int dn_las_write(const char *path, const dn_las_buf *buf);
int dn_las_read(const char *path, dn_las_buf *out);

/* ---- per-pile stages of processPileUps (commands/processPileUps/package.d:474-619) -------------- */

/* filterLocalAlignments!(la => la.averageErrorRate <= maxAlignmentError)  dazzler.d:3885-3899,
 * package.d:483-485: keeps records with diffs/(aepos-abpos) <= max_err (fp64 compare). In place,
 * order preserving; trace storage is left untouched (toff still points into it). */
int dn_las_filter_error(dn_las_buf *las, double max_err);
/* filterPileUpAlignments(..., Yes.forceFlat) / isValidPileUpAlignment  dazzler.d:4084-4141: keeps
 * aread != bread and (left-anchored and right-proper) or (right-anchored and left-proper) within
 * `allowance` (= trace spacing, commandline.d:2325-2332).  alen/blen = read lengths of the A/B DB. */
int dn_las_filter_pileup(dn_las_buf *las, const int32_t *alen, int32_t na, const int32_t *blen, int32_t nb, int32_t allowance);
/* computeQVs(db, las, coverage)  dazzler.d:3782-3792 (`DAScover`, `DASqv -c`) and
 * computeIntrinsicQV  dazzler.d:4303-4308 (`computeintrinsicqv -d`): one QV byte in [0,50] per
 * trace-spacing tile of every read; *qv has qoff[nreads] bytes, read r owns [qoff[r], qoff[r+1]).
 * `las` must be sorted by A read.  Free both arrays with dn_free. */
int dn_compute_qvs(const int32_t *rlen, int32_t nreads, const dn_las_buf *las, int32_t coverage, uint8_t **qv, int64_t **qoff);
/* same with one coverage per read (a batch of pile-ups in one block: package.d:498-501 per pile) */
int dn_compute_qvs_v(const int32_t *rlen, int32_t nreads, const dn_las_buf *las, int32_t coverage, const int32_t *cov_per_read,
                     uint8_t **qv, int64_t **qoff);
void dn_free(void *p);
/* getConsensus(db, las, readId, opts)  dazzler.d:4213-4255 (`daccord -f -I<i>,<i>`): full-length
 * consensus of each listed read (0-based) over the local alignments in `las` that have it as A read.
 * Several reads = several piles in one launch.  Sequences come back as base codes 0..3. */
typedef struct dn_seq_buf { int32_t nseq; int64_t *off; uint8_t *bases; } dn_seq_buf;
int dn_consensus(const dn_block *db, const dn_las_buf *las, const int32_t *reads, int32_t nreads, dn_seq_buf *out);
void dn_seq_free(dn_seq_buf *buf);

/* chainLocalAlignments(db, las, chainingOptions)  dazzler.d:3995-4018 -> common/alignments/chaining.d:122-334:
 * per (A,B) pair the best chain(s) of local alignments; the records are rewritten chain by chain with
 * START(+BEST)/NEXT flags as writeAlignmentChain does (dazzler.d:2037-2083); ELIM records are dropped.
 * Options = ChainingOptions (commandline.d:2820-2830; defaults 1000, 10000, 0.3, 1.0, tspace).
 * `las` must be ordered by (aread, bread). */
int dn_las_chain(dn_las_buf *las, int32_t max_indel, int32_t max_chain_gap, double max_rel_overlap, double min_rel_score,
                 int32_t min_score);

/* filterPileUpAlignments(..., Yes.forceFlat) tail (dazzler.d:4084-4093): clear the chain flags and sort the
 * records in FlatLocalAlignment order (base.d:1787-1809). */
int dn_las_force_flat(dn_las_buf *las);
/* findReferenceReadCandidates (processPileUps/package.d:518-568) for a batch of pile-ups (host logic):
 * rank[pile_off[p] .. pile_off[p+1]) = reads of pile p ordered by (numBadQVs, meanQV, readId).  group[r] < 0 = read r is
 * not in allowedReferenceReadIds (package.d:456-468): it neither enters the QV histogram nor the ranking. */
int dn_reference_read_candidates(const uint8_t *qv, const int64_t *qoff, const int32_t *group, int32_t nreads, int32_t npiles,
                                 double bad_fraction, int32_t *rank, int64_t *pile_off);

/* ---- the whole per-pile-up device path in ONE call (SURVEY §8b "fused fast path") ------------------- */

/* One pile-up of a batch as PileUpProcessor holds it after crop() (processPileUps/package.d:411-426):
 * its cropped reads (cropper.d:97-380; read i <-> pileUp[i]), which of them are in allowedReferenceReadIds
 * (package.d:456-468), and its flanking contigs (croppingPositions order) with their part of the repeat mask
 * (reduceRepeatMaskToFlankingContigs / adjustRepeatMaskToMakeMappingPossible, package.d:399-454). */
typedef struct dn_pileup_desc {
    int32_t nreads;
    const int32_t *rlen;        /* [nreads] cropped read lengths */
    const uint8_t *bases;       /* base codes 0..3 of the cropped reads, concatenated in read order */
    const uint8_t *allowed;     /* [nreads] != 0: member of allowedReferenceReadIds; NULL = all */
    int32_t nflanks;            /* flanking contigs, 1 (extension) or 2 (gap); 0 = no post-consensus alignment */
    const int32_t *flank_read;  /* [nflanks] 0-based read index of each flanking contig in the `ref` block */
    const int64_t *mask_anno;   /* optional repeat mask of the flanking contigs, track layout (dazzler.d:4943-5052): */
    const int32_t *mask_data;   /*   nflanks+1 int64 byte offsets into int32 (begin, end) pairs; NULL = none */
} dn_pileup_desc;

/* The options of `dentist process` that reach the tools (commandline.d: maxAlignmentError :1808, minAnchorLength :2036,
 * properAlignmentAllowance :2317-2332, badFraction :1101, chainingOptions :2820-2830; minQVCoverage dazzler.d:3771,
 * forceLargeTracePointType dazzler.d:154). */
typedef struct dn_pileup_params {
    double max_alignment_error;          /* 0.3: filterLocalAlignments, and -e(1 - err) of the pile alignment */
    int32_t min_anchor_length;           /* 500: -l of the pile alignment */
    int32_t tspace;                      /* 126: -s of the pile and flank alignments (<= 128) */
    int32_t proper_alignment_allowance;  /* 126: filterPileUpAlignments */
    double bad_fraction;                 /* 0.08 */
    int32_t min_qv_coverage;             /* 4 */
    int32_t dust;                        /* != 0: dbdust(croppedDb) + -mdust (package.d:476-481) */
    int32_t max_indel, max_chain_gap;    /* 1000, 10000 */
    double max_rel_overlap, min_rel_score; /* 0.3, 1.0 */
    int32_t min_score;                   /* 0 = tspace */
    int32_t k, flank_k;                  /* 14, 14: daligner's -k for the pile / flank alignments */
    int32_t bridge;                      /* 1: -B of both alignments (commandline.d:2886-2902, 2918-2935), see dn_las_bridge */
} dn_pileup_params;
void dn_pileup_params_default(dn_pileup_params *p);

#define DN_PILE_OK 0
#define DN_PILE_EMPTY_ALIGNMENT 1        /* "empty pileup alignment"                    package.d:487-490 */
#define DN_PILE_EMPTY_AFTER_FILTER 2     /* "empty pileup alignment after filtering"    package.d:512-515 */
#define DN_PILE_NO_REFERENCE_READ 3      /* "no valid reference read found"             package.d:331-339 */
const char *dn_pile_status_string(int32_t status);   /* the reference's message for a status */

/* Per pile-up result: what processPileUp holds when it reaches getInsertionAlignment() (package.d:341-345). */
typedef struct dn_insertion_out {
    int32_t status;             /* DN_PILE_*: != 0 => the D host logs `pileUpSkipped` with dn_pile_status_string() */
    int32_t reference_read;     /* referenceReadIdx: index in the pile-up of the read that was corrected, -1 if none */
    int32_t ntries;             /* reference read candidates tried (consensus failures retry the next one, :307-329) */
    int64_t cons_len;
    uint8_t *consensus;         /* base codes 0..3 of the consensus (read 1 of consensusDb, package.d:783) */
    dn_las_buf flank_las;       /* postConsensusAlignment before filterContainedAlignmentChains: aread = index into the
                                   pile-up's flank list, bread = 0 (the consensus), LAsort order, traces included */
} dn_insertion_out;
void dn_insertion_free(dn_insertion_out *out, int32_t n);

/* computeQVs + findReferenceReadCandidates + selectReferenceRead/computeConsensus (with retry) +
 * alignConsensusToFlankingContigs (package.d:303-341) for n pile-ups at once: every alignment / QV / consensus step is
 * ONE grouped launch sequence over all pile-ups instead of >= 12 forked tools per pile-up.  `ref` = the resident
 * reference block the flank ids refer to (may be NULL when no pile-up has flanks).  `out` = caller array of n results;
 * a pile-up that fails is reported in out[i].status, it does not fail the call.  Free with dn_insertion_free. */
int dn_process_pileups(const dn_block *ref, const dn_pileup_desc *piles, int32_t n, const dn_pileup_params *p, dn_insertion_out *out);

/* OR a mask track (track layout, dazzler.d:4943-5052: nreads+1 int64 byte offsets, int32 (begin, end) pairs) into a
 * resident block's seed-exclusion mask -- what `-m<track>` does for daligner / damapper (dazzler.d:5842-5848). */
int dn_block_add_mask(dn_block *blk, const int64_t *mask_anno, const int32_t *mask_data);

/* damapper-style chain flags: START on the first record of a chain, NEXT on continuations, BEST on the
 * top-scoring chain of every B read -- the flags DENTIST decodes at dazzler.d:1738-1755 and packs into
 * AlignmentChains at :708-743.  `las` must be in LAsort order (as dn_align_blocks returns it). */
int dn_las_chain_mapper(dn_las_buf *las, int32_t nb_reads, int32_t max_indel, int32_t max_gap);

/* `damapper -C` (dazzler.d:5931-5936): "Y.X.las ... contains all the same matches as in X.Y.las" with the reads' roles
 * swapped.  `las` = records of A.B.las (traces required), a / b = the resident blocks they refer to; out = the records of
 * B.A.las in LAsort order: coordinates mirrored (complemented alignments: each read in its own frame), trace points
 * re-laid every tspace bases of the new A read along the per-tile alignment path, chain flags cleared. */
int dn_las_transpose(const dn_block *a, const dn_block *b, const dn_las_buf *las, dn_las_buf *out);

/* What `daligner -B` adds (dazzler.d:5823-5824, "bridge consecutive aligned segments into one if possible"; DENTIST passes
 * it for the pile and the flank alignments, commandline.d:2886-2902, 2918-2935): neighbouring records of one (aread, bread,
 * comp) whose gap is short (<= 128 A bases, <= 250 B bases, diagonal drift within the error budget) become one record; the
 * bridge is a unit-cost global alignment of the gap, the trace points of the merged record follow the concatenated path.
 * `las` in LAsort order with traces, a / b = the resident blocks; cdiff = round(6 / (1 - e)) (20 at -e0.7).  In place;
 * *nbridged (may be NULL) = bridges made.  DALIGNER's own rule is not in the reference: specification in oracle/. */
int dn_las_bridge(const dn_block *a, const dn_block *b, dn_las_buf *las, int32_t cdiff, int64_t *nbridged);

/* What damapper reports: per mapped read its best chain and, with -n<f>, every chain scoring at least the fraction f of
 * the best (dazzler.d:5920-5923).  n_frac <= 0: the BEST chains only (DENTIST passes no -n, commandline.d:2943-2955).
 * `las` must carry the flags of dn_las_chain_mapper.  In place, order preserving. */
int dn_las_keep_best_chains(dn_las_buf *las, int32_t nb_reads, double n_frac);

/* The alignment filters of collectPileUps (commands/collectPileUps/filter.d:122-356, order of package.d:129-141:
 * LQ, Improper, WeaklyAnchored, Contained, Ambiguous, Redundant) over the AlignmentChains of a chained
 * ref-vs-reads LAS (`las` in LAsort order with START/NEXT flags).  mask_* = repeat mask on the A contigs in
 * the track layout (may be NULL).  Returns per chain the index of its first record and a status byte
 * (0 kept, 1 LQ, 2 improper, 3 weakly anchored, 4 contained, 5 ambiguous read, 6 redundant read, 7 disabled on
 * input) and per B read whether the filters consumed it (removed from `unusedReads`).  Free with dn_free. */
int dn_collect_filter(const dn_las_buf *las, const int32_t *alen, int32_t na, const int32_t *blen, int32_t nb,
                      const int64_t *mask_anno, const int32_t *mask_data, double max_err, int32_t allowance, int32_t min_anchor,
                      int64_t *nchains, int32_t **chain_first, uint8_t **chain_status, uint8_t **read_used);

/* BadAlignmentCoverageAssessor (commands/maskRepetitiveRegions.d:258-420) over the chains of a ref-vs-reads or self
 * LAS (`alignmentIntervals`, :186-205): the parts of every A contig whose alignment coverage is < lower or > upper,
 * in the mask-track layout (*anno = na+1 int64 byte offsets, *data = int32 (begin, end) pairs; free with dn_free).
 * improper_only != 0 counts only chains that are not isProper(allowance) (the second pass of assessRepeatStructure,
 * :160-176).  An empty LAS gives an empty mask (:349-350). */
int dn_mask_coverage(const dn_las_buf *las, const int32_t *alen, int32_t na, const int32_t *blen, int32_t nb,
                     double lower, double upper, int32_t improper_only, int32_t allowance, int64_t **anno, int32_t **data);

/* `dentist propagate-mask` (commands/propagateMask.d:109-300): every local alignment of `las` (trace points required)
 * carries the parts of its A contig's mask (mask_anno / mask_data: track layout, sorted disjoint intervals) over to its
 * B read -- begin rounded down, end rounded up to a trace point (base.d:185-242), mirrored for complement alignments;
 * the union per B read returns in the track layout (*anno = nb+1 offsets, *data = pairs; free with dn_free). */
int dn_propagate_mask(const dn_las_buf *las, int32_t na, const int64_t *mask_anno, const int32_t *mask_data,
                      const int32_t *blen, int32_t nb, int64_t **anno, int32_t **data);

/* dbdust(db, opts)  dazzler.d:3815-3818 (`DBdust -w -t -m`): low-complexity intervals of every read of a
 * resident block in the reference's mask-track layout (dazzler.d:4943-5052): *anno = nreads+1 int64
 * byte offsets into *data, *data = int32 (begin,end) pairs.  Free both with dn_free.  Feed them to
 * dn_block_desc.mask_* (what `-mdust` does). */
int dn_dust_block(const dn_block *blk, int32_t window, double threshold, int32_t minlen, int64_t **anno, int32_t **data);

/* dbdust + `-mdust` in one step (processPileUps/package.d:476-481, 655-665): DUST intervals of the resident block are
 * OR-ed into its own seed-exclusion mask; nothing returns to the host except the number of masked bases (may be NULL). */
int dn_block_mask_dust(dn_block *blk, int32_t window, double threshold, int32_t minlen, int64_t *masked_bases);

/* ---- file-level drop-ins for dazzler.d ---------------------------------------------------- */

/* getDalignment(dbA[, dbB], opts, outdir)  dazzler.d:3829-3844 / dalign() :6131-6140.
 * dbB == NULL => self comparison.  Writes outdir/<A>.<B>.las; `opts` are daligner flags
 * ("-s126", "-l500", "-e0.7", "-k14", "-T8", "-B", "-A", "-I", "-mdust", ...). */
int dn_dalign(const char *dbA, const char *dbB, const char *const *opts, int nopts, const char *outdir);
/* dbdust(dbFile, dbdustOptions)  dazzler.d:3815-3818: writes the `dust` track (.<db>.dust.anno/.data). */
int dn_dbdust(const char *db, const char *const *opts, int nopts);
/* getConsensus(dbFile, filteredLasFile, readId, options)  dazzler.d:4213-4255: consensus of read `read_id_1based`
 * (DENTIST ids are 1-based; daccord's -I is 0-based, :4225-4227) over the alignments in `las`; writes
 * <dir>/<db>-daccord-I<i>-<i>.dam (+ hidden .idx/.bps/.hdr) holding exactly one read and returns its path in
 * out_db.  DN_ERR_EMPTY ("empty consensus", :4232-4235) when nothing comes back. */
int dn_consensus_db(const char *db, const char *las, uint32_t read_id_1based, const char *const *opts, int nopts, char *out_db, size_t cap);
/* computeQVs(dbFile, lasFile, coverage)  dazzler.d:3782-3792 (`DAScover -v`, `DASqv -v -c<coverage>`, :6142-6156) on
 * files: writes the `qual` track of `db` (.<db>.qual.anno/.data: int32 nreads, int32 0, nreads+1 int64 byte offsets; one QV
 * byte in [0,50] per trace-spacing tile).  coverage == 0 stands for DAScover's own estimate (DENTIST never passes it,
 * package.d:498-503). */
int dn_compute_qvs_db(const char *db, const char *las, uint32_t coverage);
/* the read DENTIST does afterwards -- getDbRecords(db, [readNumber, intrinsicQualityVector]) = `DBdump -r -i`,
 * package.d:520-523, decoded by DbRecord.fromQVChar dazzler.d:2877-2898 -- without the text round trip: read r owns
 * qv[qoff[r] .. qoff[r+1]).  Free both with dn_free. */
int dn_read_qvs_db(const char *db, uint8_t **qv, int64_t **qoff, int32_t *nreads);
/* getDamapping(refDb, queryDb, opts, outdir)  dazzler.d:3855-3866 / damapper() :6163-6170. */
int dn_damap(const char *refDb, const char *queryDb, const char *const *opts, int nopts, const char *outdir);

#ifdef __cplusplus
}
#endif
#endif
