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

int set_arg_error(const char* msg) {
    snprintf(g_last_error, sizeof(g_last_error), "%s", msg);
    return SS_ERR_INVALID_ARGUMENT;
}

}  // namespace ss

extern "C" int ss_abi_version(void) { return SS_ABI_VERSION; }
extern "C" const char* ss_last_error_string(void) { return ss::g_last_error; }
extern "C" long long ss_launch_count(void) { return ss::g_launches.load(); }
