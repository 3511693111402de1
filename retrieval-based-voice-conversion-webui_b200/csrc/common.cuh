// Common host/device helpers for librvcb200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <stdexcept>

namespace rvcb {

void set_last_error(const std::string& s);

struct Error : std::runtime_error {
    explicit Error(const std::string& s) : std::runtime_error(s) {}
};

#define RVCB_CHECK(cond, msg)                                                                 \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            throw ::rvcb::Error(std::string(__FILE__) + ":" + std::to_string(__LINE__) +      \
                                ": " + (msg));                                                \
        }                                                                                     \
    } while (0)

#define CUDA_CHECK(expr)                                                                      \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            throw ::rvcb::Error(std::string(__FILE__) + ":" + std::to_string(__LINE__) +      \
                                ": CUDA error: " + cudaGetErrorString(_e) + " in " #expr);    \
        }                                                                                     \
    } while (0)

#define KERNEL_CHECK() CUDA_CHECK(cudaGetLastError())

// launch counter (bench.py reports gpu_launches from it)
extern unsigned long long g_launch_count;
inline void count_launch(int n = 1) { g_launch_count += n; }
// Upper bound on the CTAs of a persistent grid (0 = all SMs).  The Python front doors lower it to half the machine while the two
// independent front branches (RMVPE on one stream, HuBERT + retrieval on another) are being launched, so that kernels of the two
// streams run side by side on disjoint SMs instead of taking turns at the whole chip (rvcb_set_grid_cap).
extern int g_grid_cap;
inline int sm_budget(int sms) { return (g_grid_cap > 0 && g_grid_cap < sms) ? g_grid_cap : sms; }

enum Act : int {
    ACT_NONE = 0,
    ACT_RELU = 1,
    ACT_GELU = 2,      // exact erf
    ACT_LRELU = 3,     // slope = act param
    ACT_TANH = 4,
    ACT_SIGMOID = 5,
};

__device__ __forceinline__ float apply_act(float v, int act, float p) {
    switch (act) {
        case ACT_RELU: return fmaxf(v, 0.f);
        case ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
        case ACT_LRELU: return v > 0.f ? v : v * p;
        case ACT_TANH: return tanhf(v);
        case ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        default: return v;
    }
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long ceil_div_l(long a, long b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

// ---------------------------------------------------------------------------------------
// Device bump arena: every handle owns one; activations of one forward pass are carved
// from it and the pointer is reset at the start of each call (no cudaMalloc on the path).
// ---------------------------------------------------------------------------------------
struct Arena {
    char* base = nullptr;
    size_t cap = 0, off = 0, high = 0;
    unsigned generation = 0;          // bumped on every growth (diagnostics / tests)
    char* retired[24] = {};           // outgrown blocks, kept alive until the handle dies
    int n_retired = 0;
    // Growth never frees: CUDA graphs captured by the Python front doors (Pipeline / rtrvc.RVC) bake arena addresses
    // into their kernel nodes, and a graph for an earlier, shorter shape may be replayed after a longer utterance made
    // the arena grow.  The outgrown block stays allocated (its layout for that shape is still self-consistent: the
    // arena is reset at the start of every call), growth is geometric (>= 1.5x) so the retired blocks sum to < 2x cap.
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        size_t want = bytes;
        if (cap) want = bytes > cap + cap / 2 ? bytes : cap + cap / 2;
        char* nb = nullptr;
        CUDA_CHECK(cudaMalloc(&nb, want));
        if (base) {
            if (n_retired < 24) retired[n_retired++] = base;
            else CUDA_CHECK(cudaFree(base));          // unreachable in practice: 24 geometric growths = 16000x
        }
        base = nb;
        cap = want;
        ++generation;
    }
    void reset() { off = 0; }
    template <typename T>
    T* alloc(size_t n) {
        size_t bytes = (n * sizeof(T) + 1023) & ~size_t(1023);
        RVCB_CHECK(off + bytes <= cap, "arena overflow: need " + std::to_string(off + bytes) + " have " + std::to_string(cap));
        T* p = reinterpret_cast<T*>(base + off);
        off += bytes;
        if (off > high) high = off;
        return p;
    }
    ~Arena() {
        if (base) cudaFree(base);
        for (int i = 0; i < n_retired; ++i) cudaFree(retired[i]);
    }
};

template <typename T>
T* dev_upload(const T* host, size_t n) {
    T* d = nullptr;
    CUDA_CHECK(cudaMalloc(&d, n * sizeof(T)));
    CUDA_CHECK(cudaMemcpy(d, host, n * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}

}  // namespace rvcb
