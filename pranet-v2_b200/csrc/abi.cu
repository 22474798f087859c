// Library-wide pieces of the C ABI: version, thread-local error string, launch counter.
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
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
}  // namespace pv2

extern "C" int pv2_version(void) { return 100; }
extern "C" const char* pv2_last_error(void) { return pv2::g_err; }
extern "C" unsigned long long pv2_launch_count(void) { return pv2::g_launches.load(std::memory_order_relaxed); }
