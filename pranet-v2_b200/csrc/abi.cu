// Library-wide pieces of the C ABI: version, thread-local error string, launch counter.
#include <mutex>
#include <unordered_set>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "pv2_common.cuh"

namespace pv2 {
static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("PV2_PDL"); return !(e && e[0] == '0'); }();
    return on;
}
int tune_int(const char* name, int def) {
    const char* e = getenv(name);
    return (e && e[0]) ? atoi(e) : def;
}
// Every pv2 kernel asks for the SAME L1 / shared-memory split (maximum shared memory).  An SM has to drain before its carve-out
// can change, and the head alternates 100-220 KB tensor-core kernels with small element-wise ones on up to twelve concurrent
// chains: with per-kernel default carve-outs the SMs kept reconfiguring and the chains serialised (~40 us between two dependent
// 10 us kernels of a chain).  The element-wise kernels stream and do not miss the L1.  PV2_CARVEOUT=-1 leaves the driver default.
// Exception (`streaming`): kernels that are pure HBM streams keep the driver default -- with the maximum-shared split the optimizer tail
// ran at 167 us instead of 147 us (5.1 vs 5.8 TB/s; measured, same box, alternating).  PV2_CARVEOUT_STREAM=100 treats them like the rest.
void prefer_max_shared(const void* kernel, bool streaming) {
    static const int pct_all = [] { const char* e = getenv("PV2_CARVEOUT"); return (e && e[0]) ? atoi(e) : 100; }();
    static const int pct_str = [] { const char* e = getenv("PV2_CARVEOUT_STREAM"); return (e && e[0]) ? atoi(e) : -1; }();
    const int pct = streaming ? pct_str : pct_all;
    if (pct < 0) return;
    static std::mutex mu;
    static std::unordered_set<const void*> done;
    std::lock_guard<std::mutex> lk(mu);
    if (done.insert(kernel).second) (void)cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
}  // namespace pv2

extern "C" int pv2_version(void) { return 100; }
extern "C" const char* pv2_last_error(void) { return pv2::g_err; }
extern "C" unsigned long long pv2_launch_count(void) { return pv2::g_launches.load(std::memory_order_relaxed); }
