// api_internal.hpp -- state shared by the C-ABI translation units.
#pragma once
#include "engine.cuh"
#include <mutex>
#include <string>

struct dn_block { dn::DevBlock b; };

namespace dnapi {
extern std::mutex g_mu;            // serialises device work: entry points are re-entrant, the GPU queue is one
extern cudaStream_t g_stream;
extern int g_device;
int fail(int code, const std::string &m);
int ensure_device();
// Set by dn_process_pileups around its stage calls: the LAS handed from stage to stage there was produced by this library
// in the same call, so the per-record / per-trace-point host validation of caller-supplied LAS is skipped.
extern thread_local bool g_trusted_las;
struct TrustedLas { bool prev; TrustedLas() : prev(g_trusted_las) { g_trusted_las = true; } ~TrustedLas() { g_trusted_las = prev; } };
template <typename F> int guarded(F &&f) {
    try { return f(); }
    catch (const dn::Error &e) { return fail(DN_ERR_CUDA, e.what()); }
    catch (const std::bad_alloc &) { return fail(DN_ERR_INVALID, "out of host memory"); }
    catch (const std::exception &e) { return fail(DN_ERR_INVALID, e.what()); }
}
}  // namespace dnapi
