// Library-level entry points of libb200splat.so (version, error strings, launch counter).
#include <atomic>

#include "common.cuh"

static std::atomic<long long> g_b2s_launches{0};

void b2s_count_launch(int n) { g_b2s_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" int b2s_version(void) { return 200; }

extern "C" long long b2s_launch_count(void) { return g_b2s_launches.load(std::memory_order_relaxed); }

extern "C" const char *b2s_error_string(int code) {
    switch (code) {
        case B2S_OK: return "ok";
        case B2S_ERR_ARG: return "b200splat: invalid argument";
        case B2S_ERR_UNSUPPORTED: return "b200splat: unsupported configuration (tile size / channels / degree)";
        case B2S_ERR_WORKSPACE: return "b200splat: workspace too small";
        default: break;
    }
    if (code <= -1000) return cudaGetErrorString((cudaError_t)(-(code + 1000)));
    return "b200splat: unknown error";
}
