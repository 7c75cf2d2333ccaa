#include "dazzdb.hpp"
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

namespace dn {
namespace {

#pragma pack(push, 1)
struct IdxHeader {            // DAZZ_DB struct as written at the head of .idx (LP64, natural alignment)
    int32_t ureads, treads, cutoff, allarr;
    float freq[4];
    int32_t maxlen, pad0;
    int64_t totlen;
    int32_t nreads, trimmed, part, ufirst, tfirst, pad1;
    uint64_t path; int32_t loaded, pad2; uint64_t bases, reads, tracks;
};
struct IdxRead {              // DAZZ_READ
    int32_t origin, rlen, fpulse, pad0;
    int64_t boff, coff;
    int32_t flags, pad1;
};
#pragma pack(pop)
static_assert(sizeof(IdxHeader) == 112, "DAZZ_DB header is 112 bytes on LP64");
static_assert(sizeof(IdxRead) == 40, "DAZZ_READ is 40 bytes on LP64");
constexpr int DB_BEST = 0x0800;

struct PathParts { std::string dir, root, ext; int block; };

bool exists(const std::string &p) { FILE *f = fopen(p.c_str(), "rb"); if (!f) return false; fclose(f); return true; }

bool split_path(const std::string &path, PathParts &pp, std::string &err) {
    size_t sl = path.find_last_of('/');
    pp.dir = sl == std::string::npos ? "." : path.substr(0, sl);
    std::string base = sl == std::string::npos ? path : path.substr(sl + 1);
    pp.ext = ""; pp.block = 0;
    if (base.size() > 3 && base.compare(base.size() - 3, 3, ".db") == 0) { pp.ext = ".db"; base.resize(base.size() - 3); }
    else if (base.size() > 4 && base.compare(base.size() - 4, 4, ".dam") == 0) { pp.ext = ".dam"; base.resize(base.size() - 4); }
    // optional block suffix  root.<n>
    size_t dot = base.find_last_of('.');
    if (dot != std::string::npos && dot + 1 < base.size()) {
        bool dig = true; for (size_t i = dot + 1; i < base.size(); i++) if (base[i] < '0' || base[i] > '9') dig = false;
        if (dig) {
            std::string root = base.substr(0, dot);
            if (exists(pp.dir + "/" + root + ".db") || exists(pp.dir + "/" + root + ".dam")) { pp.block = atoi(base.c_str() + dot + 1); base = root; }
        }
    }
    pp.root = base;
    if (pp.ext.empty()) {
        if (exists(pp.dir + "/" + base + ".db")) pp.ext = ".db";
        else if (exists(pp.dir + "/" + base + ".dam")) pp.ext = ".dam";
        else { err = "cannot find DB stub for " + path; return false; }
    }
    return true;
}

bool slurp(const std::string &p, std::vector<uint8_t> &out) {
    FILE *f = fopen(p.c_str(), "rb"); if (!f) return false;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? n : 0);
    bool ok = n <= 0 || fread(out.data(), 1, n, f) == (size_t)n;
    fclose(f); return ok;
}
}  // namespace

bool read_dazz_db(const std::string &path, const std::vector<std::string> &mask_tracks, HostDb &out, std::string &err) {
    PathParts pp;
    if (!split_path(path, pp, err)) return false;
    const std::string stub = pp.dir + "/" + pp.root + pp.ext;
    FILE *f = fopen(stub.c_str(), "r");
    if (!f) { err = "cannot open " + stub; return false; }
    char line[4096]; int nblocks = 0, cutoff = 0, all = 1; long long bsize = 0;
    std::vector<int> ufirst, tfirst;
    bool in_blocks = false;
    while (fgets(line, sizeof line, f)) {
        int a, b;
        if (sscanf(line, "blocks = %d", &nblocks) == 1) continue;
        if (sscanf(line, "size = %lld cutoff = %d all = %d", &bsize, &cutoff, &all) == 3) { in_blocks = true; continue; }
        if (in_blocks && sscanf(line, " %d %d", &a, &b) == 2) { ufirst.push_back(a); tfirst.push_back(b); }
    }
    fclose(f);
    std::vector<uint8_t> idx, bps;
    if (!slurp(pp.dir + "/." + pp.root + ".idx", idx) || idx.size() < sizeof(IdxHeader)) { err = "cannot read .idx of " + stub; return false; }
    if (!slurp(pp.dir + "/." + pp.root + ".bps", bps)) { err = "cannot read .bps of " + stub; return false; }
    IdxHeader h; memcpy(&h, idx.data(), sizeof h);
    if (h.ureads < 0 || idx.size() < sizeof(IdxHeader) + (size_t)h.ureads * sizeof(IdxRead)) { err = "truncated .idx of " + stub; return false; }
    const IdxRead *rd = (const IdxRead *)(idx.data() + sizeof(IdxHeader));
    if (!in_blocks) { cutoff = h.cutoff < 0 ? 0 : h.cutoff; all = h.allarr & 1; }
    int u0 = 0, u1 = h.ureads;
    if (pp.block > 0) {
        if (pp.block > nblocks || (int)ufirst.size() < pp.block + 1) { err = "block out of range in " + path; return false; }
        u0 = ufirst[pp.block - 1]; u1 = ufirst[pp.block];
    }
    out.name = pp.root + (pp.block > 0 ? "." + std::to_string(pp.block) : "");
    out.rlen.clear(); out.boff.clear();
    std::vector<int> uid;                   // untrimmed id of every kept read
    int tcount_before = 0;
    for (int u = 0; u < h.ureads; u++) {
        bool kept = rd[u].rlen >= cutoff && (all || (rd[u].flags & DB_BEST));
        if (u < u0) { tcount_before += kept; continue; }
        if (u >= u1) break;
        if (!kept) continue;
        out.rlen.push_back(rd[u].rlen); out.boff.push_back(rd[u].boff); uid.push_back(u);
    }
    out.bps.swap(bps);
    // mask tracks
    out.mask_anno.clear(); out.mask_data.clear();
    if (!mask_tracks.empty()) {
        std::vector<std::vector<int32_t>> iv(out.rlen.size());
        bool any = false;
        for (const std::string &t : mask_tracks) {
            std::vector<uint8_t> anno, data;
            // a block DB takes its block-level track .<root>.<block>.<track> (what DBdust X.<block> writes, before Catrack
            // merges the blocks; getMaskFiles dazzler.d:4870-4912) when there is one, else the whole-DB track
            std::string base = pp.dir + "/." + pp.root + "." + t;
            bool block_track = false;
            if (pp.block > 0) {
                const std::string bb = pp.dir + "/." + pp.root + "." + std::to_string(pp.block) + "." + t;
                if (exists(bb + ".anno")) { base = bb; block_track = true; }
            }
            if (!slurp(base + ".anno", anno) || anno.size() < 8) continue;        // a missing track masks nothing
            slurp(base + ".data", data);
            int32_t tn, tsz; memcpy(&tn, anno.data(), 4); memcpy(&tsz, anno.data() + 4, 4);
            if (tsz != 0 || anno.size() < 8 + (size_t)(tn + 1) * 8) continue;
            const int64_t *o = (const int64_t *)(anno.data() + 8);
            for (size_t r = 0; r < uid.size(); r++) {
                int64_t id = block_track ? (tn == u1 - u0 ? (int64_t)uid[r] - u0 : (int64_t)r)
                                         : (tn == h.ureads) ? uid[r] : (int64_t)tcount_before + (int64_t)r;
                if (id < 0 || id >= tn) continue;
                for (int64_t b = o[id]; b + 8 <= o[id + 1] && b + 8 <= (int64_t)data.size(); b += 8) {
                    int32_t be[2]; memcpy(be, data.data() + b, 8);
                    iv[r].push_back(be[0]); iv[r].push_back(be[1]); any = true;
                }
            }
        }
        if (any) {
            out.mask_anno.resize(out.rlen.size() + 1);
            int64_t off = 0;
            for (size_t r = 0; r < iv.size(); r++) { out.mask_anno[r] = off; out.mask_data.insert(out.mask_data.end(), iv[r].begin(), iv[r].end()); off += 4 * (int64_t)iv[r].size(); }
            out.mask_anno[iv.size()] = off;
        }
    }
    return true;
}

bool write_dazz_db(const std::string &path, const std::vector<std::vector<uint8_t>> &reads, std::string &err) {
    size_t sl = path.find_last_of('/');
    std::string dir = sl == std::string::npos ? "." : path.substr(0, sl);
    std::string base = sl == std::string::npos ? path : path.substr(sl + 1);
    bool dam = base.size() > 4 && base.compare(base.size() - 4, 4, ".dam") == 0;
    std::string root = base.substr(0, base.size() - (dam ? 4 : 3));
    IdxHeader h; memset(&h, 0, sizeof h);
    std::vector<IdxRead> rd(reads.size());
    std::vector<uint8_t> bps;
    int64_t tot = 0; int maxlen = 0; double fr[4] = {0, 0, 0, 0};
    for (size_t r = 0; r < reads.size(); r++) {
        memset(&rd[r], 0, sizeof(IdxRead));
        rd[r].origin = (int)r; rd[r].rlen = (int)reads[r].size(); rd[r].fpulse = 0; rd[r].boff = (int64_t)bps.size();
        rd[r].coff = dam ? 0 : -1; rd[r].flags = DB_BEST | 850;
        for (size_t i = 0; i < reads[r].size(); i += 4) {
            uint8_t b = 0;
            for (int j = 0; j < 4; j++) { uint8_t c = i + j < reads[r].size() ? (reads[r][i + j] & 3) : 0; b |= c << (6 - 2 * j); }
            bps.push_back(b);
        }
        for (uint8_t c : reads[r]) fr[c & 3] += 1;
        tot += (int64_t)reads[r].size(); if ((int)reads[r].size() > maxlen) maxlen = (int)reads[r].size();
    }
    h.ureads = h.treads = (int)reads.size(); h.cutoff = 0; h.allarr = 1; h.maxlen = maxlen; h.totlen = tot;
    for (int i = 0; i < 4; i++) h.freq[i] = tot ? (float)(fr[i] / tot) : 0.25f;
    h.nreads = (int)reads.size(); h.trimmed = 1;
    FILE *f = fopen((dir + "/" + root + (dam ? ".dam" : ".db")).c_str(), "w");
    if (!f) { err = "cannot write stub"; return false; }
    fprintf(f, "files = %9d\n  %9d %s %s\nblocks = %9d\nsize = %11lld cutoff = %9d all = %1d\n %9d %9d\n %9d %9d\n",
            1, (int)reads.size(), root.c_str(), root.c_str(), 1, 200000000ll, 0, 1, 0, 0, (int)reads.size(), (int)reads.size());
    fclose(f);
    f = fopen((dir + "/." + root + ".idx").c_str(), "wb"); if (!f) { err = "cannot write .idx"; return false; }
    fwrite(&h, sizeof h, 1, f); if (!rd.empty()) fwrite(rd.data(), sizeof(IdxRead), rd.size(), f); fclose(f);
    f = fopen((dir + "/." + root + ".bps").c_str(), "wb"); if (!f) { err = "cannot write .bps"; return false; }
    if (!bps.empty()) fwrite(bps.data(), 1, bps.size(), f);
    fclose(f);
    if (dam) { f = fopen((dir + "/." + root + ".hdr").c_str(), "wb"); if (f) { fprintf(f, ">%s\n", root.c_str()); fclose(f); } }
    return true;
}

bool write_mask_track(const std::string &dbpath, const std::string &track, const std::vector<std::vector<int32_t>> &iv, std::string &err) {
    PathParts pp;
    if (!split_path(dbpath, pp, err)) return false;
    // a block DB (X.<n>) gets the block-level track .<root>.<n>.<track>, like DBdust X.<n>; the whole-DB track is not touched
    std::string base = pp.dir + "/." + pp.root + (pp.block > 0 ? "." + std::to_string(pp.block) : "") + "." + track;
    FILE *fa = fopen((base + ".anno").c_str(), "wb"), *fd = fopen((base + ".data").c_str(), "wb");
    if (!fa || !fd) { if (fa) fclose(fa); if (fd) fclose(fd); err = "cannot write track " + base; return false; }
    int32_t n = (int32_t)iv.size(), sz = 0;
    bool ok = fwrite(&n, 4, 1, fa) == 1 && fwrite(&sz, 4, 1, fa) == 1;
    int64_t off = 0;
    for (size_t r = 0; ok && r <= iv.size(); r++) {
        ok = fwrite(&off, 8, 1, fa) == 1;
        if (r < iv.size()) { if (!iv[r].empty()) ok = ok && fwrite(iv[r].data(), 4, iv[r].size(), fd) == iv[r].size(); off += 4 * (int64_t)iv[r].size(); }
    }
    if (fclose(fa) != 0) ok = false;
    if (fclose(fd) != 0) ok = false;
    if (!ok) { err = "write failed for track " + base; return false; }
    return true;
}

// Byte-valued per-read track (DASqv's `qual`, computeintrinsicqv's `inqual`): the same .anno/.data pair as an interval
// track -- int32 nreads, int32 size (= 0: variable length), nreads + 1 int64 byte offsets -- with one QV byte per
// trace-spacing tile in .data (what `DBdump -i` prints as letters, dazzler.d:2877-2898).
bool write_byte_track(const std::string &dbpath, const std::string &track, const uint8_t *data, const int64_t *off, int32_t n, std::string &err) {
    PathParts pp;
    if (!split_path(dbpath, pp, err)) return false;
    std::string base = pp.dir + "/." + pp.root + (pp.block > 0 ? "." + std::to_string(pp.block) : "") + "." + track;
    FILE *fa = fopen((base + ".anno").c_str(), "wb"), *fd = fopen((base + ".data").c_str(), "wb");
    if (!fa || !fd) { if (fa) fclose(fa); if (fd) fclose(fd); err = "cannot write track " + base; return false; }
    int32_t sz = 0;
    bool ok = fwrite(&n, 4, 1, fa) == 1 && fwrite(&sz, 4, 1, fa) == 1;
    ok = ok && fwrite(off, 8, (size_t)n + 1, fa) == (size_t)n + 1;
    if (ok && off[n] > 0) ok = fwrite(data, 1, (size_t)off[n], fd) == (size_t)off[n];
    if (fclose(fa) != 0) ok = false;
    if (fclose(fd) != 0) ok = false;
    if (!ok) { err = "write failed for track " + base; return false; }
    return true;
}

bool read_byte_track(const std::string &dbpath, const std::string &track, std::vector<int64_t> &off, std::vector<uint8_t> &data, std::string &err) {
    PathParts pp;
    if (!split_path(dbpath, pp, err)) return false;
    std::string base = pp.dir + "/." + pp.root + (pp.block > 0 ? "." + std::to_string(pp.block) : "") + "." + track;
    std::vector<uint8_t> anno;
    if (!slurp(base + ".anno", anno) || anno.size() < 8) { err = "cannot read track " + base; return false; }
    int32_t n, sz; memcpy(&n, anno.data(), 4); memcpy(&sz, anno.data() + 4, 4);
    if (n < 0 || sz != 0 || anno.size() < 8 + ((size_t)n + 1) * 8) { err = "malformed track " + base; return false; }
    off.resize((size_t)n + 1); memcpy(off.data(), anno.data() + 8, ((size_t)n + 1) * 8);
    data.clear(); slurp(base + ".data", data);
    for (int32_t r = 0; r < n; r++) if (off[r] < 0 || off[r] > off[r + 1]) { err = "malformed track " + base; return false; }
    if ((int64_t)data.size() < off[n]) { err = "track data shorter than its index: " + base; return false; }
    return true;
}

}  // namespace dn
