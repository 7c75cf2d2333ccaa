// pile.cuh -- device entry points of the per-pile stages (pile.cu).
#pragma once
#include "engine.cuh"

namespace dn {

struct ConsTask { int32_t la, ap, alen, bp, bb; };
struct ConsGeom {
    const u32 *fwd, *rc; const int64_t *off; const int32_t *len;
    const int64_t *vote_off;      // [ntargets+1] first vote column of each target read (L+1 columns each)
};

// `daligner -B` bridge between two neighbouring records: rows = A[ga, ga + n), columns = B[gb, gb + m) of the strand array,
// a0 = A coordinate of row 0 (for the trace-point grid), nrow grid rows in (0, n] whose crossings go to rows[row_off ...]
struct BridgeTask { int64_t ga, gb; int32_t n, m, a0, comp, nrow; int64_t row_off; };
void launch_bridge(const BridgeTask *tasks, int64_t ntasks, const u32 *a_fwd, const u32 *b_fwd, const u32 *b_rc, int ts, u32 *scratch /* cons_vote_threads() * 2048 */,
                   int32_t *total, int2 *rows, cudaStream_t s);

// transposition (damapper -C): A and B blocks of the alignment
struct TrGeom { const u32 *a_fwd, *b_fwd, *b_rc; const int64_t *a_off, *b_off; const int32_t *a_len, *b_len; };
void launch_tr_tiles(const ConsTask *tasks, int64_t ntasks, const dn_las_record *rec, const int64_t *toff, const uint16_t *trace,
                     const int64_t *la_task0, TrGeom G, int ts, int kmax, u32 *scratch, int4 *cross, int32_t *ncross, int32_t *tcost, cudaStream_t s);
void launch_tr_assemble(const dn_las_record *rec, int64_t nla, const int64_t *la_task0, TrGeom G, int kmax, const int4 *cross,
                        const int32_t *ncross, const int32_t *tcost, const int64_t *out_toff, dn_las_record *out, uint16_t *out_trace,
                        int32_t *status, cudaStream_t s);

// mode 0: keep diffs/(aepos-abpos) <= max_err; mode 1: isValidPileUpAlignment(allowance). Order preserving.
void las_filter_device(const dn_las_record *rec, const int64_t *toff, int64_t n, int mode, double max_err, const int32_t *alen,
                       const int32_t *blen, int allowance, dn_las_record *orec, int64_t *otoff, int64_t *n_out, cudaStream_t s);
void qv_device(const int32_t *rlen, int nreads, const dn_las_record *rec, int64_t nla, const int64_t *toff, const uint16_t *trace,
               int ts, int cov, const int32_t *cov_per_read, const int64_t *qoff, uint8_t *qv, cudaStream_t s);
// damapper-style chain flags (START/NEXT/BEST) on records in LAsort order
void mapper_chain_device(dn_las_record *rec, int64_t n, int nb_reads, int max_indel, int max_gap, cudaStream_t s);
void launch_cons_tasks(const dn_las_record *rec, const int64_t *toff, const uint16_t *trace, const int32_t *vla, int nvla,
                       const int64_t *task_off, int ts, ConsTask *tasks, cudaStream_t s);
int cons_vote_threads();
void launch_cons_vote(const ConsTask *tasks, int64_t ntasks, const dn_las_record *rec, const int32_t *la_target, ConsGeom G,
                      u32 *scratch, int32_t *cnt, int32_t *ins, int32_t *insn, int32_t *cov, cudaStream_t s);
void launch_cons_count(ConsGeom G, const int32_t *targets, int ntargets, int64_t ncols, const int32_t *cnt, const int32_t *ins,
                       const int32_t *insn, const int32_t *cov, int32_t *nemit, uint8_t *sym, cudaStream_t s);
void launch_cons_write(int64_t ncols, const int32_t *nemit, const int32_t *eoff, const uint8_t *sym, uint8_t *out, cudaStream_t s);

}  // namespace dn
