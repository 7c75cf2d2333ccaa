// dazzdb.hpp -- minimal native reader/writer for DAZZ_DB databases (.db/.dam stub + hidden
// .idx/.bps) and interval tracks (.anno/.data), so that the engine takes the same inputs the
// external tools take from DENTIST (dazzler.d:137-140, 171-182, 4383-4480, 4943-5052).
// The .idx/.bps layout follows thegenemyers/DAZZ_DB @ d22ae58d (DB.h); it is NOT in the reference
// tree and could not be verified against real files here (SURVEY Appendix A.1, items marked *).
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/dentist_b200.h"

namespace dn {

struct HostDb {
    std::string name;                    // base name used in <A>.<B>.las
    std::vector<int32_t> rlen;
    std::vector<int64_t> boff;
    std::vector<uint8_t> bps;            // DAZZ .bps bytes of the selected reads' file
    std::vector<int64_t> mask_anno;      // nreads+1 byte offsets (empty = no mask)
    std::vector<int32_t> mask_data;
    dn_block_desc desc() const {
        dn_block_desc d{};
        d.nreads = (int32_t)rlen.size(); d.format = DN_SEQ_BPS; d.rlen = rlen.data(); d.boff = boff.data();
        d.data = bps.data(); d.data_bytes = (int64_t)bps.size();
        d.mask_anno = mask_anno.empty() ? nullptr : mask_anno.data();
        d.mask_data = mask_anno.empty() ? nullptr : mask_data.data();
        return d;
    }
};

bool read_dazz_db(const std::string &path, const std::vector<std::string> &mask_tracks, HostDb &out, std::string &err);
// writes <dir>/<name>.db|.dam + .<name>.idx + .<name>.bps (+ .hdr for .dam) with one block, all reads kept
bool write_dazz_db(const std::string &path, const std::vector<std::vector<uint8_t>> &reads, std::string &err);
bool write_mask_track(const std::string &dbpath, const std::string &track, const std::vector<std::vector<int32_t>> &intervals, std::string &err);
// per-read byte tracks (`qual` of DASqv, `inqual` of computeintrinsicqv): one QV byte per trace-spacing tile
bool write_byte_track(const std::string &dbpath, const std::string &track, const uint8_t *data, const int64_t *off, int32_t nreads, std::string &err);
bool read_byte_track(const std::string &dbpath, const std::string &track, std::vector<int64_t> &off, std::vector<uint8_t> &data, std::string &err);

}  // namespace dn
