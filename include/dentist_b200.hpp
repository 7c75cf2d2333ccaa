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

  private:
    dn_block *h_ = nullptr;
};

/// In-memory getDalignment: the records of A.B.las without the file system.
inline Las align(const Block &a, const Block &b, const dn_align_params &p) {
    Las l;
    enforce(dn_align_blocks(a.raw(), b.raw(), &p, l.raw()));
    return l;
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

}  // namespace dazzler
}  // namespace dentist
