// dentist_b200.hpp -- C++ host-side mirror of the slice of `source/dentist/dazzler.d` that sits on the
// hot path.  The reference host is D; no D toolchain exists in this image, so this header restates the
// D interface (same names, argument meaning and error behaviour) on top of the C ABI in
// dentist_b200.h.  A non-zero return becomes DazzlerCommandException, exactly what the D wrappers
// throw today when an external tool fails (dazzler.d:199-206, 6586-6591).
#pragma once
#include "dentist_b200.h"

#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace dentist {
namespace dazzler {

struct DazzlerCommandException : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline void enforce(int rc) {
    if (rc != DN_OK) throw DazzlerCommandException(dn_last_error());
}

namespace detail {
inline std::vector<const char *> cstrs(const std::vector<std::string> &v) {
    std::vector<const char *> out;
    for (const auto &s : v) out.push_back(s.c_str());
    return out;
}
inline std::string baseName(const std::string &db) {
    std::string b = db.substr(db.find_last_of('/') == std::string::npos ? 0 : db.find_last_of('/') + 1);
    for (const char *ext : {".db", ".dam"}) {
        std::string e(ext);
        if (b.size() > e.size() && b.compare(b.size() - e.size(), e.size(), e) == 0) b.resize(b.size() - e.size());
    }
    return b;
}
}  // namespace detail

/// dazzler.d:4345-4354
inline std::string getLasFile(const std::string &dbA, const std::string &dbB, const std::string &baseDirectory) {
    return baseDirectory + "/" + detail::baseName(dbA) + "." + detail::baseName(dbB.empty() ? dbA : dbB) + ".las";
}

/// dazzler.d:3829-3844 -- `daligner <opts> dbA [dbB]` in outdir; returns the LAS path.
inline std::string getDalignment(const std::string &dbA, const std::string &dbB, const std::vector<std::string> &dalignerOpts,
                                 const std::string &outdir) {
    auto o = detail::cstrs(dalignerOpts);
    enforce(dn_dalign(dbA.c_str(), dbB.empty() ? nullptr : dbB.c_str(), o.data(), (int)o.size(), outdir.c_str()));
    return getLasFile(dbA, dbB, outdir);
}
inline std::string getDalignment(const std::string &dbA, const std::vector<std::string> &dalignerOpts, const std::string &outdir) {
    return getDalignment(dbA, std::string(), dalignerOpts, outdir);
}

/// dazzler.d:3855-3866 -- `damapper -C <opts> refDb queryDb`; returns outdir/<ref>.<query>.las.
inline std::string getDamapping(const std::string &refDb, const std::string &queryDb, const std::vector<std::string> &damapperOpts,
                                const std::string &outdir) {
    auto o = detail::cstrs(damapperOpts);
    enforce(dn_damap(refDb.c_str(), queryDb.c_str(), o.data(), (int)o.size(), outdir.c_str()));
    return getLasFile(refDb, queryDb, outdir);
}

/// dazzler.d:3815-3818 -- writes the `dust` track of dbFile.
inline void dbdust(const std::string &dbFile, const std::vector<std::string> &dbdustOptions) {
    auto o = detail::cstrs(dbdustOptions);
    enforce(dn_dbdust(dbFile.c_str(), o.data(), (int)o.size()));
}

/// An in-memory LAS (what getAlignments, dazzler.d:431-480, would parse from the file).
class Las {
  public:
    Las() { buf_ = dn_las_buf{}; }
    Las(const Las &) = delete;
    Las &operator=(const Las &) = delete;
    Las(Las &&o) noexcept : buf_(o.buf_) { o.buf_ = dn_las_buf{}; }
    ~Las() { dn_las_free(&buf_); }
    dn_las_buf *raw() { return &buf_; }
    const dn_las_buf *raw() const { return &buf_; }
    int64_t size() const { return buf_.nrec; }
    bool lasEmpty() const { return buf_.nrec == 0; }                       // dazzler.d:220-225
    const dn_las_record &operator[](int64_t i) const { return buf_.rec[i]; }
    const uint16_t *trace(int64_t i) const { return buf_.trace + buf_.toff[i]; }
    int tracePointDistance() const { return buf_.tspace; }

    /// filterLocalAlignments!(la => la.averageErrorRate <= maxAlignmentError)  dazzler.d:3885-3899
    void filterLocalAlignments(double maxAlignmentError) { enforce(dn_las_filter_error(&buf_, maxAlignmentError)); }
    /// filterPileUpAlignments(db, las, properAlignmentAllowance, Yes.forceFlat)  dazzler.d:4043-4094
    void filterPileUpAlignments(const std::vector<int32_t> &aLengths, const std::vector<int32_t> &bLengths, int32_t allowance) {
        enforce(dn_las_filter_pileup(&buf_, aLengths.data(), (int32_t)aLengths.size(), bLengths.data(), (int32_t)bLengths.size(), allowance));
    }
    /// chainLocalAlignments(db, las, chainingOptions)  dazzler.d:3995-4018, defaults commandline.d:2820-2830
    void chainLocalAlignments(int32_t maxIndel = 1000, int32_t maxChainGap = 10000, double maxRelOverlap = 0.3, double minRelScore = 1.0,
                              int32_t minScore = -1) {
        enforce(dn_las_chain(&buf_, maxIndel, maxChainGap, maxRelOverlap, minRelScore, minScore < 0 ? buf_.tspace : minScore));
    }
    void writeAlignments(const std::string &lasFile) const { enforce(dn_las_write(lasFile.c_str(), &buf_)); }   // dazzler.d:1913-1960

  private:
    dn_las_buf buf_;
};

/// A DAZZ_DB read block resident in HBM.
class Block {
  public:
    explicit Block(const dn_block_desc &d) { enforce(dn_block_upload(&d, &h_)); }
    Block(const Block &) = delete;
    Block &operator=(const Block &) = delete;
    ~Block() { dn_block_free(h_); }
    const dn_block *raw() const { return h_; }
    /// dbdust(db) + `-mdust` (processPileUps/package.d:476-481): returns the number of masked bases.
    int64_t maskDust(int32_t window = 64, double threshold = 2.0, int32_t minLength = 10) {
        int64_t m = 0;
        enforce(dn_block_mask_dust(h_, window, threshold, minLength, &m));
        return m;
    }

  private:
    dn_block *h_ = nullptr;
};

/// In-memory getDalignment: the records of A.B.las without the file system.
inline Las align(const Block &a, const Block &b, const dn_align_params &p) {
    Las l;
    enforce(dn_align_blocks(a.raw(), b.raw(), &p, l.raw()));
    return l;
}

/// What `-B` adds to a daligner call (dazzler.d:5823-5824; pileUpAlignmentOptions / postConsensusAlignmentOptions pass it):
/// neighbouring local alignments of a read pair separated by a short gap become one.  Returns the number of bridges made.
inline int64_t bridge(const Block &a, const Block &b, Las &las, double averageCorrelationRate = 0.7) {
    int64_t n = 0;
    enforce(dn_las_bridge(a.raw(), b.raw(), las.raw(), (int32_t)(6.0 / (1.0 - averageCorrelationRate) + 0.5), &n));
    return n;
}

/// computeQVs(db, las, coverage)  dazzler.d:3782-3792 -> one QV byte per trace-spacing tile of every read.
inline std::vector<std::vector<uint8_t>> computeQVs(const std::vector<int32_t> &readLengths, const Las &las, uint32_t coverage) {
    uint8_t *qv = nullptr; int64_t *off = nullptr;
    enforce(dn_compute_qvs(readLengths.data(), (int32_t)readLengths.size(), las.raw(), (int32_t)coverage, &qv, &off));
    std::vector<std::vector<uint8_t>> out(readLengths.size());
    for (size_t r = 0; r < out.size(); r++) out[r].assign(qv + off[r], qv + off[r + 1]);
    dn_free(qv); dn_free(off);
    return out;
}

/// computeQVs(dbFile, lasFile, coverage)  dazzler.d:3782-3792, file form: writes the `qual` track of dbFile.
inline void computeQVs(const std::string &dbFile, const std::string &lasFile, uint32_t coverage = 0) {
    enforce(dn_compute_qvs_db(dbFile.c_str(), lasFile.c_str(), coverage));
}
/// the intrinsicQualityVector column of getDbRecords (`DBdump -r -i`, package.d:520-523) read from the `qual` track
inline std::vector<std::vector<uint8_t>> getIntrinsicQVs(const std::string &dbFile) {
    uint8_t *qv = nullptr; int64_t *off = nullptr; int32_t n = 0;
    enforce(dn_read_qvs_db(dbFile.c_str(), &qv, &off, &n));
    std::vector<std::vector<uint8_t>> out((size_t)n);
    for (size_t r = 0; r < out.size(); r++) out[r].assign(qv + off[r], qv + off[r + 1]);
    dn_free(qv); dn_free(off);
    return out;
}

/// A mask as DENTIST's ReferenceRegion restricted to one DB: per contig a sorted list of disjoint [begin, end).
using Mask = std::vector<std::vector<std::pair<int32_t, int32_t>>>;

namespace detail {
inline Mask takeTrack(int64_t *anno, int32_t *data, size_t n) {
    Mask out(n);
    for (size_t r = 0; r < n; r++)
        for (int64_t i = anno[r] / 4; i < anno[r + 1] / 4; i += 2) out[r].emplace_back(data[i], data[i + 1]);
    dn_free(anno); dn_free(data);
    return out;
}
}  // namespace detail

/// BadAlignmentCoverageAssessor(lower, upper)(alignmentIntervals(improperOnly), contigIntervals)
/// commands/maskRepetitiveRegions.d:135-232, 258-420
inline Mask maskCoverage(const Las &las, const std::vector<int32_t> &aLengths, const std::vector<int32_t> &bLengths, double lowerLimit,
                         double upperLimit, bool improperOnly = false, int32_t properAlignmentAllowance = 0) {
    int64_t *anno = nullptr; int32_t *data = nullptr;
    enforce(dn_mask_coverage(las.raw(), aLengths.data(), (int32_t)aLengths.size(), bLengths.data(), (int32_t)bLengths.size(), lowerLimit,
                             upperLimit, improperOnly ? 1 : 0, properAlignmentAllowance, &anno, &data));
    return detail::takeTrack(anno, data, aLengths.size());
}

/// MaskPropagator  commands/propagateMask.d:109-300: the A-contig mask carried over to the B reads through the trace points.
inline Mask propagateMask(const Las &las, const Mask &inputMask, const std::vector<int32_t> &bLengths) {
    std::vector<int64_t> manno(inputMask.size() + 1, 0); std::vector<int32_t> mdata;
    for (size_t c = 0; c < inputMask.size(); c++) {
        manno[c] = 4 * (int64_t)mdata.size();
        for (auto &iv : inputMask[c]) { mdata.push_back(iv.first); mdata.push_back(iv.second); }
    }
    manno[inputMask.size()] = 4 * (int64_t)mdata.size();
    mdata.push_back(0); mdata.push_back(0);
    int64_t *anno = nullptr; int32_t *data = nullptr;
    enforce(dn_propagate_mask(las.raw(), (int32_t)inputMask.size(), manno.data(), mdata.data(), bLengths.data(), (int32_t)bLengths.size(), &anno, &data));
    return detail::takeTrack(anno, data, bLengths.size());
}

/// getConsensus(db, las, readId, opts)  dazzler.d:4213-4238; readId is 1-based like in DENTIST.
/// Throws "empty consensus" like the reference when nothing comes back.
inline std::vector<uint8_t> getConsensus(const Block &db, const Las &las, size_t readId) {
    int32_t r = (int32_t)readId - 1;
    dn_seq_buf s{};
    enforce(dn_consensus(db.raw(), las.raw(), &r, 1, &s));
    std::vector<uint8_t> out(s.bases + s.off[0], s.bases + s.off[1]);
    dn_seq_free(&s);
    if (out.empty()) throw std::runtime_error("empty consensus");            // dazzler.d:4232-4235
    return out;
}

/// One pile-up after crop() and the result PileUpProcessor needs to go on with getInsertionAlignment()
/// (commands/processPileUps/package.d:283-374).
struct PileUp {
    std::vector<std::vector<uint8_t>> croppedReads;       // base codes 0..3
    std::vector<uint8_t> allowedReferenceRead;            // empty = all
    std::vector<int32_t> flankingContigs;                 // 0-based read ids in the reference block
};
struct PileUpResult {
    int32_t status = 0; std::string reason; int32_t referenceReadIdx = -1;
    std::vector<uint8_t> consensus;
    std::vector<dn_las_record> postConsensusAlignment; std::vector<std::vector<uint16_t>> traces;
};

/// processPileUps' device path for a whole batch in ONE call (dn_process_pileups).  A failing pile-up comes back with
/// the reference's reason (`pileUpSkipped`), it does not throw.
inline std::vector<PileUpResult> processPileUps(const Block &refDb, const std::vector<PileUp> &pileUps, const dn_pileup_params *params = nullptr) {
    const size_t n = pileUps.size();
    std::vector<dn_pileup_desc> d(n); std::vector<std::vector<int32_t>> rlen(n); std::vector<std::vector<uint8_t>> bases(n);
    for (size_t i = 0; i < n; i++) {
        for (const auto &r : pileUps[i].croppedReads) { rlen[i].push_back((int32_t)r.size()); bases[i].insert(bases[i].end(), r.begin(), r.end()); }
        d[i] = dn_pileup_desc{};
        d[i].nreads = (int32_t)rlen[i].size(); d[i].rlen = rlen[i].data(); d[i].bases = bases[i].data();
        d[i].allowed = pileUps[i].allowedReferenceRead.empty() ? nullptr : pileUps[i].allowedReferenceRead.data();
        d[i].nflanks = (int32_t)pileUps[i].flankingContigs.size(); d[i].flank_read = pileUps[i].flankingContigs.data();
    }
    std::vector<dn_insertion_out> o(n);
    enforce(dn_process_pileups(refDb.raw(), d.data(), (int32_t)n, params, o.data()));
    std::vector<PileUpResult> out(n);
    for (size_t i = 0; i < n; i++) {
        out[i].status = o[i].status; out[i].reason = dn_pile_status_string(o[i].status); out[i].referenceReadIdx = o[i].reference_read;
        if (o[i].cons_len) out[i].consensus.assign(o[i].consensus, o[i].consensus + o[i].cons_len);
        const dn_las_buf &l = o[i].flank_las;
        for (int64_t x = 0; x < l.nrec; x++) {
            out[i].postConsensusAlignment.push_back(l.rec[x]);
            out[i].traces.emplace_back(l.trace + l.toff[x], l.trace + l.toff[x] + l.rec[x].tlen);
        }
    }
    dn_insertion_free(o.data(), (int32_t)n);
    return out;
}

}  // namespace dazzler
}  // namespace dentist
