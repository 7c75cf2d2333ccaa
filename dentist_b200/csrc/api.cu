// api.cu -- the C ABI (include/dentist_b200.h): lifecycle, resident blocks, in-memory alignment,
// LAS serialisation and the file-level drop-ins for dazzler.d's getDalignment / getDamapping.
#include "api_internal.hpp"
#include "dazzdb.hpp"
#include <string.h>
#include <memory>
#include <chrono>
#include <stdlib.h>

using namespace dn;
using namespace dnapi;

namespace dnapi {
thread_local std::string t_err;
thread_local bool g_trusted_las = false;
std::mutex g_mu;
int g_device = -1;
cudaStream_t g_stream = nullptr;

int fail(int code, const std::string &m) { t_err = m; return code; }

static void setup_device() {
    if (!g_stream) cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking);
}

int ensure_device() {
    if (g_device >= 0) return DN_OK;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(DN_ERR_NO_DEVICE, "no CUDA device available (dentist_b200 has no CPU fallback)");
    if (cudaSetDevice(0) != cudaSuccess) return fail(DN_ERR_NO_DEVICE, "cudaSetDevice(0) failed");
    g_device = 0;
    setup_device();
    return DN_OK;
}

}  // namespace dnapi

namespace {
AlignParams to_internal(const dn_align_params *p) {
    dn_align_params d; dn_align_params_default(&d);
    if (p) d = *p;
    AlignParams q;
    q.k = d.k; q.w = d.w; q.h = d.h; q.t = d.t; q.tspace = d.tspace; q.minlen = d.minlen;
    double e = d.e; if (e < 0.7) e = 0.7; if (e > 0.99) e = 0.99;
    q.cdiff = (int)(6.0 / (1.0 - e) + 0.5);                   // S = 3*(i+j) - C*d drifts to zero at 1-e diffs/column
    q.xdrop = d.xdrop; q.wmax = d.wmax; q.rounds = d.rounds; q.poolmul = d.poolmul;
    q.self = (d.self_block && !d.identity) ? 1 : 0;
    q.join_mode = d.join_mode;
    return q;
}

void to_buf(HostLas &h, int tspace, dn_las_buf *out) {
    memset(out, 0, sizeof *out);
    out->nrec = h.nrec; out->ntrace = h.ntrace; out->tspace = tspace; out->stats = h.stats;
    out->rec = h.rec; out->toff = h.toff; out->trace = h.trace;        // ownership moves to the caller (dn_las_free)
    h.rec = nullptr; h.toff = nullptr; h.trace = nullptr;
}
}  // namespace

extern "C" {

int dn_init(int device, const char *tmpdir) {
    (void)tmpdir;
    std::lock_guard<std::mutex> lk(g_mu);
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return fail(DN_ERR_NO_DEVICE, "no CUDA device available (dentist_b200 has no CPU fallback)");
    if (device < 0) device = 0;
    if (device >= n) return fail(DN_ERR_INVALID, "device index out of range");
    if (cudaSetDevice(device) != cudaSuccess) return fail(DN_ERR_CUDA, "cudaSetDevice failed");
    g_device = device;
    setup_device();
    return DN_OK;
}
int dn_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_stream) { cudaStreamSynchronize(g_stream); cudaStreamDestroy(g_stream); g_stream = nullptr; }
    arena().destroy(); dcache_destroy();
    g_device = -1;
    return DN_OK;
}
const char *dn_last_error(void) { return t_err.c_str(); }
const char *dn_version(void) { return "dentist_b200 0.1.0 (sm_100a)"; }
uint64_t dn_launch_count(void) { return dn::g_launches.load(); }

void dn_align_params_default(dn_align_params *p) {
    memset(p, 0, sizeof *p);
    p->k = 14; p->w = 6; p->h = 35; p->t = 32; p->tspace = 100; p->minlen = 1000; p->e = 0.7;
    p->identity = 0; p->self_block = 0; p->rounds = 3; p->xdrop = 300; p->wmax = 30; p->poolmul = 64;
}

void dn_las_free(dn_las_buf *b) {
    if (!b) return;
    hcache_free(b->rec); hcache_free(b->toff); hcache_free(b->trace);
    memset(b, 0, sizeof *b);
}

int dn_block_upload(const dn_block_desc *desc, dn_block **out) {
    if (!desc || !out) return fail(DN_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = ensure_device()) return rc;
    return guarded([&] {
        cudaSetDevice(g_device);
        dn_block *b = new dn_block();
        try { block_upload(*desc, b->b, g_stream); } catch (...) { delete b; throw; }
        *out = b; return DN_OK;
    });
}
int dn_block_crop(const dn_block *src, int32_t n, const int32_t *read, const int32_t *begin, const int32_t *end, const int32_t *group,
                  dn_block **out) {
    if (!src || !out || n < 0 || (n && (!read || !begin || !end))) return fail(DN_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = ensure_device()) return rc;
    return guarded([&] {
        cudaSetDevice(g_device);
        dn_block *b = new dn_block();
        try { block_crop(src->b, n, read, begin, end, group, b->b, g_stream); } catch (...) { delete b; throw; }
        *out = b; return DN_OK;
    });
}
void dn_block_free(dn_block *blk) {
    if (!blk) return;
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_device >= 0) cudaSetDevice(g_device);
    delete blk;
}
int64_t dn_block_bases(const dn_block *blk) { return blk ? blk->b.total_real : 0; }

int dn_block_index(dn_block *blk, int32_t k) {
    if (!blk) return fail(DN_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = ensure_device()) return rc;
    return guarded([&] {
        cudaSetDevice(g_device); arena().reset();
        if (k <= 0) blk->b.index.drop();
        else block_build_index(blk->b, k, g_stream);
        return DN_OK;
    });
}

int dn_align_blocks(const dn_block *a, const dn_block *b, const dn_align_params *p, dn_las_buf *out) {
    if (!a || !b || !out) return fail(DN_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = ensure_device()) return rc;
    return guarded([&] {
        cudaSetDevice(g_device);
        AlignParams q = to_internal(p);
        HostLas h;
        align_blocks(a->b, b->b, q, h, g_stream);
        to_buf(h, q.tspace, out);
        return DN_OK;
    });
}

int dn_align_host(const dn_block_desc *a, const dn_block_desc *b, const dn_align_params *p, dn_las_buf *out) {
    if (!a || !b || !out) return fail(DN_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = ensure_device()) return rc;
    return guarded([&] {
        cudaSetDevice(g_device);
        static cudaStream_t copy_stream = nullptr;
        if (!copy_stream) DN_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        static cudaStream_t pack_stream = nullptr;           // B's chunks are packed here as they arrive on the copy stream
        if (!pack_stream) DN_CUDA(cudaStreamCreateWithFlags(&pack_stream, cudaStreamNonBlocking));
        const bool same = (a == b);
        std::unique_ptr<dn_block> ba(new dn_block()), bb(same ? nullptr : new dn_block());
        // B's (large) host->device copy and packing run on the copy stream while A is indexed on the engine's stream;
        // align_blocks waits for B right before the join
        const bool trace = getenv("DN_TRACE") != nullptr;
        auto now = [] { return std::chrono::steady_clock::now(); };
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        // A first: both uploads share the host->device copy engine, and A's small copy must not queue behind B's
        const auto t0 = now();
        block_upload(*a, ba->b, g_stream, true);             // no host sync: the alignment follows on the same stream
        const auto t1 = now();
        if (!same) block_upload(*b, bb->b, copy_stream, true, pack_stream);
        const auto t2 = now();
        AlignParams q = to_internal(p);
        HostLas h;
        try { align_blocks(ba->b, same ? ba->b : bb->b, q, h, g_stream); }
        catch (...) { cudaStreamSynchronize(copy_stream); cudaStreamSynchronize(pack_stream); throw; }
        const auto t3 = now();
        cudaStreamSynchronize(copy_stream); cudaStreamSynchronize(pack_stream);
        to_buf(h, q.tspace, out);
        if (trace) fprintf(stderr, "[dn trace] align_host: A upload %.3f ms, B upload (host side) %.3f ms, align %.3f ms (device %.3f), tail %.3f ms\n",
                           ms(t0, t1), ms(t1, t2), ms(t2, t3), h.stats.ms_total, ms(t3, now()));
        return DN_OK;
    });
}

int dn_las_merge_device(const void *d_rec, int64_t nrec, const void *d_trace, int64_t ntrace, int32_t tspace, int64_t max_alen,
                        int64_t max_blen, int64_t na_reads, int64_t nb_reads, dn_las_buf *out) {
    if (!out || nrec < 0 || ntrace < 0 || (nrec && !d_rec) || (ntrace && !d_trace)) return fail(DN_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(g_mu);
    if (int rc = ensure_device()) return rc;
    return guarded([&] {
        cudaSetDevice(g_device);
        cudaDeviceSynchronize();                         // the inputs come from another stream (e.g. NCCL via torch)
        HostLas h;
        merge_las_device((const dn_las_record *)d_rec, nrec, (const uint16_t *)d_trace, ntrace, max_alen, max_blen, na_reads, nb_reads, h, g_stream);
        to_buf(h, tspace, out);
        return DN_OK;
    });
}

int dn_las_write(const char *path, const dn_las_buf *buf) {
    if (!path || !buf) return fail(DN_ERR_INVALID, "null argument");
    return guarded([&] {
        FILE *f = fopen(path, "wb");
        if (!f) return fail(DN_ERR_IO, std::string("cannot open for writing: ") + path);
        int64_t novl = buf->nrec; int32_t ts = buf->tspace;
        bool ok = fwrite(&novl, 8, 1, f) == 1 && fwrite(&ts, 4, 1, f) == 1;
        const bool large = ts > 125;                     // TRACE_XOVR, dazzler.d:2019-2025
        std::vector<uint8_t> small;
        for (int64_t i = 0; ok && i < buf->nrec; i++) {
            ok = fwrite(&buf->rec[i], 40, 1, f) == 1;
            const uint16_t *t = buf->trace + buf->toff[i]; const int tl = buf->rec[i].tlen;
            if (large) ok = ok && fwrite(t, 2, tl, f) == (size_t)tl;
            else {
                small.resize(tl);
                for (int x = 0; x < tl; x++) {
                    // a value that does not fit the 1-byte trace of tspace <= 125 would silently break sum(bbases) == bepos - bbpos
                    if (t[x] > 255) { fclose(f); remove(path); return fail(DN_ERR_INVALID, "trace value > 255 cannot be written with trace spacing <= 125 (use -s126 or larger)"); }
                    small[x] = (uint8_t)t[x];
                }
                ok = ok && fwrite(small.data(), 1, tl, f) == (size_t)tl;
            }
        }
        if (fclose(f) != 0 || !ok) return fail(DN_ERR_IO, std::string("write failed: ") + path);
        return DN_OK;
    });
}

int dn_las_read(const char *path, dn_las_buf *out) {
    if (!path || !out) return fail(DN_ERR_INVALID, "null argument");
    return guarded([&] {
        FILE *f = fopen(path, "rb");
        if (!f) return fail(DN_ERR_IO, std::string("cannot open: ") + path);
        int64_t novl = 0; int32_t ts = 0;
        if (fread(&novl, 8, 1, f) != 1 || fread(&ts, 4, 1, f) != 1 || novl < 0) { fclose(f); return fail(DN_ERR_IO, "error reading LAS file: unexpected end of file; expected header"); }
        std::vector<dn_las_record> rec(novl); std::vector<int64_t> toff(novl); std::vector<uint16_t> trace;
        const bool large = ts > 125;
        std::vector<uint8_t> small;
        for (int64_t i = 0; i < novl; i++) {
            if (fread(&rec[i], 40, 1, f) != 1) { fclose(f); return fail(DN_ERR_IO, "error reading LAS file: unexpected end of file; expected overlapHead"); }
            const int tl = rec[i].tlen;
            if (tl < 0 || (tl & 1)) { fclose(f); return fail(DN_ERR_IO, "illegal value for tlen: must be multiple of 2"); }
            toff[i] = (int64_t)trace.size();
            size_t o = trace.size(); trace.resize(o + tl);
            bool ok;
            if (large) ok = fread(trace.data() + o, 2, tl, f) == (size_t)tl;
            else { small.resize(tl); ok = fread(small.data(), 1, tl, f) == (size_t)tl; for (int x = 0; x < tl; x++) trace[o + x] = small[x]; }
            if (!ok) { fclose(f); return fail(DN_ERR_IO, "error reading LAS file: unexpected end of file; expected tracePoints"); }
        }
        fclose(f);
        HostLas h; h.nrec = novl; h.ntrace = (int64_t)trace.size();
        h.rec = (dn_las_record *)hcache_alloc(sizeof(dn_las_record) * (novl + 1));
        h.toff = (int64_t *)hcache_alloc(sizeof(int64_t) * (novl + 1));
        h.trace = (uint16_t *)hcache_alloc(sizeof(uint16_t) * (trace.size() + 1));
        if (novl) { memcpy(h.rec, rec.data(), sizeof(dn_las_record) * novl); memcpy(h.toff, toff.data(), sizeof(int64_t) * novl); }
        if (!trace.empty()) memcpy(h.trace, trace.data(), sizeof(uint16_t) * trace.size());
        to_buf(h, ts, out);
        return DN_OK;
    });
}

// ---- file-level drop-ins ------------------------------------------------------------------

static int parse_opts(const char *const *opts, int nopts, dn_align_params *p, bool *asym, std::vector<std::string> *masks, double *best_frac = nullptr,
                      bool *bridge = nullptr) {
    for (int i = 0; i < nopts; i++) {
        const char *o = opts[i];
        if (!o || o[0] != '-' || !o[1]) return fail(DN_ERR_INVALID, std::string("bad option: ") + (o ? o : "(null)"));
        const char *v = o + 2;
        switch (o[1]) {
            case 'k': p->k = atoi(v); break;
            case 'w': p->w = atoi(v); break;
            case 'h': p->h = atoi(v); break;
            case 't': p->t = atoi(v); break;
            case 's': p->tspace = atoi(v); break;
            case 'l': p->minlen = atoi(v); break;
            case 'e': p->e = atof(v); break;
            case 'I': p->identity = 1; break;
            case 'A': *asym = true; break;
            case 'm': masks->push_back(v); break;
            case 'n': if (best_frac) *best_frac = atof(v); break;                 // damapper: also report chains within this fraction of the best (dazzler.d:5920-5923)
            case 'B': if (bridge) *bridge = true; break;                               // bridge consecutive aligned segments (dazzler.d:5823-5824)
            case 'T': case 'M': case 'v': case 'b': case 'a': case 'C': case 'N': case 'z': case 'P': case 'p': break;   // accepted, no effect on a GPU
            default: return fail(DN_ERR_INVALID, std::string("unknown option: ") + o);
        }
    }
    if (p->k > 31) p->k = 31;                      /* damapper's default k = 20 is honoured (16-byte tuples) */
    return DN_OK;
}

static int align_files(const char *dbA, const char *dbB, const char *const *opts, int nopts, const char *outdir, bool mapper) {
    if (!dbA || !outdir) return fail(DN_ERR_INVALID, "null argument");
    return guarded([&]() -> int {
        dn_align_params p; dn_align_params_default(&p);
        if (mapper) { p.minlen = 1000; p.k = 20; }       /* damapper's own defaults (DENTIST passes neither -k nor -l, commandline.d:2943-2955) */
        bool asym = false, bridge = false; std::vector<std::string> masks; double best_frac = 0.0;     // no -n: the best chain of every read only
        if (int rc = parse_opts(opts, nopts, &p, &asym, &masks, &best_frac, &bridge)) return rc;
        const bool self = (dbB == nullptr) || std::string(dbA) == std::string(dbB);
        HostDb A, B;
        std::string err;
        if (!read_dazz_db(dbA, masks, A, err)) return fail(DN_ERR_IO, err);
        if (!self && !read_dazz_db(dbB, masks, B, err)) return fail(DN_ERR_IO, err);
        HostDb &Bx = self ? A : B;
        p.self_block = self ? 1 : 0;
        dn_block_desc da = A.desc(), db = Bx.desc();
        struct Blk { dn_block *b = nullptr; ~Blk() { if (b) dn_block_free(b); } } ga, gb;
        struct LasG { dn_las_buf l; LasG() { memset(&l, 0, sizeof l); } ~LasG() { dn_las_free(&l); } } ab, ba;
        if (int rc = dn_block_upload(&da, &ga.b)) return rc;
        if (!self) { if (int rc = dn_block_upload(&db, &gb.b)) return rc; }
        const dn_block *pb = self ? ga.b : gb.b;
        if (int rc = dn_align_blocks(ga.b, pb, &p, &ab.l)) return rc;
        if (bridge && !mapper) { if (int rc = dn_las_bridge(ga.b, pb, &ab.l, (int32_t)(6.0 / (1.0 - p.e) + 0.5), nullptr)) return rc; }   // daligner -B
        if (mapper) {
            if (int rc = dn_las_chain_mapper(&ab.l, (int32_t)Bx.rlen.size(), 1000, 10000)) return rc;
            if (int rc = dn_las_keep_best_chains(&ab.l, (int32_t)Bx.rlen.size(), best_frac)) return rc;
        }
        std::string pa = std::string(outdir) + "/" + A.name + "." + Bx.name + ".las";
        if (int rc = dn_las_write(pa.c_str(), &ab.l)) return rc;
        if (!self && (!asym || mapper)) {
            // the second file holds "all the same matches" with the reads' roles swapped (damapper -C, dazzler.d:5931-5936;
            // daligner without -A, :5797-5801): the transposition of the first, not a second alignment
            if (int rc = dn_las_transpose(ga.b, gb.b, &ab.l, &ba.l)) return rc;
            if (mapper) { if (int rc = dn_las_chain_mapper(&ba.l, (int32_t)A.rlen.size(), 1000, 10000)) return rc; }
            std::string pbn = std::string(outdir) + "/" + Bx.name + "." + A.name + ".las";
            if (int rc = dn_las_write(pbn.c_str(), &ba.l)) return rc;
        }
        return DN_OK;
    });
}

int dn_dbdust(const char *db, const char *const *opts, int nopts) {
    if (!db) return fail(DN_ERR_INVALID, "null argument");
    return guarded([&]() -> int {
        int window = 64, minlen = 10; double thr = 2.0;                 // DBdust defaults (dazzler.d:3796-3806)
        for (int i = 0; i < nopts; i++) {
            const char *o = opts[i];
            if (!o || o[0] != '-' || !o[1]) return fail(DN_ERR_INVALID, "bad option");
            if (o[1] == 'w') window = atoi(o + 2); else if (o[1] == 't') thr = atof(o + 2); else if (o[1] == 'm') minlen = atoi(o + 2);
            else if (o[1] == 'b') continue; else return fail(DN_ERR_INVALID, std::string("unknown option: ") + o);
        }
        HostDb D; std::string err;
        if (!read_dazz_db(db, {}, D, err)) return fail(DN_ERR_IO, err);
        dn_block_desc d = D.desc();
        dn_block *blk = nullptr;
        if (int rc = dn_block_upload(&d, &blk)) return rc;
        int64_t *anno = nullptr; int32_t *data = nullptr;
        int rc = dn_dust_block(blk, window, thr, minlen, &anno, &data);
        dn_block_free(blk);
        if (rc) return rc;
        std::vector<std::vector<int32_t>> iv(D.rlen.size());
        for (size_t r = 0; r < iv.size(); r++) iv[r].assign(data + anno[r] / 4, data + anno[r + 1] / 4);
        hcache_free(anno); hcache_free(data);
        if (!write_mask_track(db, "dust", iv, err)) return fail(DN_ERR_IO, err);
        return DN_OK;
    });
}

int dn_consensus_db(const char *db, const char *las, uint32_t read_id_1based, const char *const *opts, int nopts, char *out_db, size_t cap) {
    (void)opts; (void)nopts;                                   // daccord options (-t, -w, -a, -k ...) have no effect here
    if (!db || !las || !out_db || read_id_1based == 0) return fail(DN_ERR_INVALID, "null argument");
    return guarded([&]() -> int {
        HostDb D; std::string err;
        if (!read_dazz_db(db, {}, D, err)) return fail(DN_ERR_IO, err);
        if (read_id_1based > D.rlen.size()) return fail(DN_ERR_INVALID, "read id out of bounds");
        dn_las_buf L; memset(&L, 0, sizeof L);
        if (int rc = dn_las_read(las, &L)) return rc;
        if (L.nrec == 0) { dn_las_free(&L); return fail(DN_ERR_EMPTY, "empty pre-consensus alignment"); }   // dazzler.d:4246-4249
        dn_block_desc d = D.desc();
        dn_block *blk = nullptr;
        int rc = dn_block_upload(&d, &blk);
        if (rc) { dn_las_free(&L); return rc; }
        const int32_t r = (int32_t)read_id_1based - 1;                       // daccord -I<i>,<i> is 0-based, dazzler.d:4225-4227
        dn_seq_buf s; memset(&s, 0, sizeof s);
        rc = dn_consensus(blk, &L, &r, 1, &s);
        dn_block_free(blk); dn_las_free(&L);
        if (rc) return rc;
        const int64_t n = s.off[1] - s.off[0];
        if (n <= 0) { dn_seq_free(&s); return fail(DN_ERR_EMPTY, "empty consensus"); }                     // dazzler.d:4232-4235
        // <db dir>/<db root>-daccord-I<i>-<i>.dam  (dazzler.d:6187-6220)
        std::string p(db);
        size_t sl = p.find_last_of('/'); std::string dir = sl == std::string::npos ? "." : p.substr(0, sl);
        std::string out = dir + "/" + D.name + "-daccord-I" + std::to_string(r) + "-" + std::to_string(r) + ".dam";
        std::vector<std::vector<uint8_t>> reads(1);
        reads[0].assign(s.bases + s.off[0], s.bases + s.off[1]);
        dn_seq_free(&s);
        if (!write_dazz_db(out, reads, err)) return fail(DN_ERR_IO, err);
        if (out.size() + 1 > cap) return fail(DN_ERR_INVALID, "output path buffer too small");
        memcpy(out_db, out.c_str(), out.size() + 1);
        return DN_OK;
    });
}

// computeQVs(dbFile, lasFile, coverage)  dazzler.d:3782-3792: `DAScover` + `DASqv -c<coverage>` on files.  Writes the `qual`
// track of `db`; DENTIST reads it back per read through `DBdump -r -i` (package.d:520-523) -- dn_read_qvs_db is that read.
int dn_compute_qvs_db(const char *db, const char *las, uint32_t coverage) {
    if (!db || !las) return fail(DN_ERR_INVALID, "null argument");
    return guarded([&]() -> int {
        HostDb D; std::string err;
        if (!read_dazz_db(db, {}, D, err)) return fail(DN_ERR_IO, err);
        dn_las_buf L; memset(&L, 0, sizeof L);
        if (int rc = dn_las_read(las, &L)) return rc;
        int64_t cov = coverage;
        if (cov == 0) {
            // no coverage given: DAScover's estimate stands in as the mean depth of the alignments over the DB (>= 1);
            // DENTIST always passes a positive coverage (package.d:498-503)
            int64_t span = 0, tot = 0;
            for (int64_t i = 0; i < L.nrec; i++) span += L.rec[i].aepos - L.rec[i].abpos;
            for (int32_t l : D.rlen) tot += l;
            cov = tot > 0 ? span / tot : 0; if (cov < 1) cov = 1;
        }
        uint8_t *qv = nullptr; int64_t *qoff = nullptr;
        int rc = dn_compute_qvs(D.rlen.data(), (int32_t)D.rlen.size(), &L, (int32_t)cov, &qv, &qoff);
        dn_las_free(&L);
        if (rc) return rc;
        const bool ok = write_byte_track(db, "qual", qv, qoff, (int32_t)D.rlen.size(), err);
        hcache_free(qv); hcache_free(qoff);
        return ok ? DN_OK : fail(DN_ERR_IO, err);
    });
}

int dn_read_qvs_db(const char *db, uint8_t **qv, int64_t **qoff, int32_t *nreads) {
    if (!db || !qv || !qoff || !nreads) return fail(DN_ERR_INVALID, "null argument");
    return guarded([&]() -> int {
        std::vector<int64_t> off; std::vector<uint8_t> data; std::string err;
        if (!read_byte_track(db, "qual", off, data, err)) return fail(DN_ERR_IO, err);
        const int32_t n = (int32_t)off.size() - 1;
        int64_t *ho = (int64_t *)hcache_alloc(sizeof(int64_t) * ((size_t)n + 1));
        uint8_t *hq = (uint8_t *)hcache_alloc((size_t)off[n] + 1);
        memcpy(ho, off.data(), sizeof(int64_t) * ((size_t)n + 1));
        if (off[n]) memcpy(hq, data.data(), (size_t)off[n]);
        *qv = hq; *qoff = ho; *nreads = n;
        return DN_OK;
    });
}

int dn_dalign(const char *dbA, const char *dbB, const char *const *opts, int nopts, const char *outdir) {
    return align_files(dbA, dbB, opts, nopts, outdir, false);
}
int dn_damap(const char *refDb, const char *queryDb, const char *const *opts, int nopts, const char *outdir) {
    return align_files(refDb, queryDb, opts, nopts, outdir, true);
}

}  // extern "C"
