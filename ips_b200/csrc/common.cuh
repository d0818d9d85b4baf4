// Shared helpers for the ips_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

namespace ipsb {

// thread-local message returned by ipsb_last_error()
char* err_buf();
int fail(const char* fmt, ...);

#define IPSB_CUDA(expr)                                                              \
    do {                                                                             \
        cudaError_t e__ = (expr);                                                    \
        if (e__ != cudaSuccess)                                                      \
            return ipsb::fail("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
    } while (0)

#define IPSB_LAUNCH_CHECK()                                                          \
    do {                                                                             \
        cudaError_t e__ = cudaGetLastError();                                        \
        if (e__ != cudaSuccess)                                                      \
            return ipsb::fail("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
    } while (0)

#define IPSB_REQUIRE(cond, ...)                                                      \
    do {                                                                             \
        if (!(cond)) return ipsb::fail(__VA_ARGS__);                                 \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
int sm_count();
// Persistent-kernel grid: the fewest CTAs that finish `units` work items in the same number of rounds as `max_ctas`
// would (192 units on 148 SMs take 2 rounds either way -> 96 CTAs); the SMs left free run the other lane's kernels.
static inline int balanced_grid(int64_t units, int max_ctas) {
    if (units <= max_ctas) return (int)units;
    const int64_t rounds = (units + max_ctas - 1) / max_ctas;
    return (int)((units + rounds - 1) / rounds);
}

// TMA-fed tcgen05 implicit GEMM (umma_conv_tma.cu)
int conv_tma(const void* x, const void* w, const float* scale, const float* shift, const void* res, void* y,
             int64_t P, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad, int relu,
             bool out_f32, cudaStream_t st, bool in_pf = false, bool out_pf = false);
// stride-1 3x3 convolution on padded-flat activations (umma_conv_halo.cu)
int conv3x3_halo(const void* x, const void* w, const float* scale, const float* shift, const void* res, void* y,
                 int64_t P, int H, int W, int Cin, int Cout, int relu, cudaStream_t st);
int conv_stem_s2d(const void* x, const void* w, const float* scale, const float* shift, void* y, int64_t P, int H, int W,
                  int Cout, int relu, cudaStream_t st);
int conv_stem_tma(const void* x, const void* w, const float* scale, const float* shift, void* y,
                  int64_t P, int H, int W, int Cout, int relu, cudaStream_t st, bool out_f32 = false);

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// 128-bit streaming load / store (read-once data: bypass L1 allocation)
__device__ __forceinline__ int4 ld_stream16(const void* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream16(void* p, const int4& v) {
    asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

}  // namespace ipsb
