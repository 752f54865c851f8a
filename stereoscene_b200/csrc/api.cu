// Library-level entry points of the C ABI: version, error text, launch counter.
#include <cstdio>
#include <cstring>
#include "common.cuh"

namespace ss {

thread_local char g_last_error[256] = "";
std::atomic<long long> g_launches{0};

int set_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s", where, cudaGetErrorString(e));
    return SS_ERR_CUDA;
}

// ---- per-kernel launch census (ss_kernel_census): check_launch() passes the kernel's name (a string literal) ----
constexpr int kMaxFamilies = 64;
static const char* g_family_name[kMaxFamilies];
static std::atomic<long long> g_family_count[kMaxFamilies];
static std::atomic<int> g_families{0};

void count_kernel(const char* name) {
    const int n = g_families.load(std::memory_order_acquire);
    for (int i = 0; i < n; ++i)
        if (g_family_name[i] == name || strcmp(g_family_name[i], name) == 0) { g_family_count[i].fetch_add(1, std::memory_order_relaxed); return; }
    static std::atomic_flag lock = ATOMIC_FLAG_INIT;
    while (lock.test_and_set(std::memory_order_acquire)) {}
    int m = g_families.load(std::memory_order_acquire), hit = -1;
    for (int i = 0; i < m; ++i)
        if (strcmp(g_family_name[i], name) == 0) hit = i;
    if (hit < 0 && m < kMaxFamilies) { g_family_name[m] = name; g_family_count[m].store(0); hit = m; g_families.store(m + 1, std::memory_order_release); }
    if (hit >= 0) g_family_count[hit].fetch_add(1, std::memory_order_relaxed);
    lock.clear(std::memory_order_release);
}

int set_arg_error(const char* msg) {
    snprintf(g_last_error, sizeof(g_last_error), "%s", msg);
    return SS_ERR_INVALID_ARGUMENT;
}

}  // namespace ss

extern "C" int ss_abi_version(void) { return SS_ABI_VERSION; }
extern "C" const char* ss_last_error_string(void) { return ss::g_last_error; }
extern "C" long long ss_launch_count(void) { return ss::g_launches.load(); }
extern "C" int ss_kernel_census(char* buf, size_t cap) {
    const int n = ss::g_families.load(std::memory_order_acquire);
    size_t off = 0;
    if (buf && cap) buf[0] = 0;
    for (int i = 0; i < n && buf && off + 1 < cap; ++i) {
        const int w = snprintf(buf + off, cap - off, "%s=%lld\n", ss::g_family_name[i], ss::g_family_count[i].load());
        if (w < 0) break;
        off += (size_t)w < cap - off ? (size_t)w : cap - off - 1;
    }
    return n;
}
