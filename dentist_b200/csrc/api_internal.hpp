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
template <typename F> int guarded(F &&f) {
    try { return f(); }
    catch (const dn::Error &e) { return fail(DN_ERR_CUDA, e.what()); }
    catch (const std::bad_alloc &) { return fail(DN_ERR_INVALID, "out of host memory"); }
    catch (const std::exception &e) { return fail(DN_ERR_INVALID, e.what()); }
}
}  // namespace dnapi
