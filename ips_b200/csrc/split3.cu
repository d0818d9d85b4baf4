// "bf16x3" precision: near-fp32 arithmetic on the bf16 tensor cores.
//
// Every fp32 operand is split into two bf16 values, v = hi + lo with hi = bf16(v), lo = bf16(v - hi) (16-17 significant
// bits together), and a product x*w is evaluated as xh*wh + xh*wl + xl*wh: three bf16 MMAs whose products are exact and are
// accumulated in fp32 (the dropped xl*wl term is below 2^-17 of |x||w|).  Activations live in HBM in a "tri" layout: per
// pixel 3C bf16 channels [hi | hi | lo]; with the weights laid out as [wh | wl | wh] along the input channels ONE launch of the
// fp32-output tcgen05 kernel (umma_conv_tma.cu) with Cin' = 3C accumulates all three products in tensor memory.  The stem
// (4 input channels, special frame layout) runs as three launches on (hi, lo) frames.  The kernels below add partial
// results, apply the residual / ReLU / max-pool in fp32 and split the result again.
// Used for index-stable selection at tensor-core speed; reference arithmetic: fp32 conv2d / linear of
// architecture/ips_net.py:17-60.
#include "common.cuh"
#include "../../include/ips_b200.h"

namespace {

using bf16 = __nv_bfloat16;

__device__ __forceinline__ void split(float v, bf16& h, bf16& l) {
    h = __float2bfloat16_rn(v);
    l = __float2bfloat16_rn(v - __bfloat162float(h));
}

// (rows, C, H, W) fp32 -> two zero-bordered (n_rows, Hp, Wp, 4) bf16 frames (image at row pt, column pl)
__global__ void stage_padded_split_kernel(const float* __restrict__ src, int64_t first_row, int64_t n_rows, int C, int H, int W,
                                          int Hp, int Wp, int pt, int pl, bf16* __restrict__ hi, bf16* __restrict__ lo) {
    const int64_t total = n_rows * Hp * Wp;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int xp = (int)(i % Wp);
        const int64_t t = i / Wp;
        const int yp = (int)(t % Hp);
        const int64_t r = t / Hp;
        const int x = xp - pl, y = yp - pt;
        __align__(8) bf16 h[4], l[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) { h[c] = __float2bfloat16_rn(0.f); l[c] = h[c]; }
        if (x >= 0 && x < W && y >= 0 && y < H) {
            const float* sp = src + ((first_row + r) * C) * (int64_t)H * W + (int64_t)y * W + x;
            for (int c = 0; c < C && c < 4; ++c) split(__ldg(sp + (int64_t)c * H * W), h[c], l[c]);
        }
        *reinterpret_cast<uint2*>(hi + i * 4) = *reinterpret_cast<const uint2*>(h);
        *reinterpret_cast<uint2*>(lo + i * 4) = *reinterpret_cast<const uint2*>(l);
    }
}

struct Sum3Params {
    const float* a; const float* b; const float* c;        // partial results (b, c may be null)
    const float* res_f32; const bf16* res_hi; const bf16* res_lo;
    bf16* out_hi; bf16* out_lo; float* out_f32;
    int64_t n4;                                            // float4 groups
    int relu;
    int tri_C;                                             // > 0: out_hi / res_hi are tri tensors (pixels, 3C) = [hi | hi | lo]
};

// offset (in elements) of the hi copy of float4 group i in a tri tensor with C channels (lo at + 2C, second hi at + C)
__device__ __forceinline__ int64_t tri_off(int64_t i, int C) {
    const int64_t e = i * 4, px = e / C;
    return px * 3 * C + (e - px * C);
}

__global__ void __launch_bounds__(256) sum3_split_kernel(Sum3Params p) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < p.n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<const float4*>(p.a)[i];
        if (p.b) { const float4 u = reinterpret_cast<const float4*>(p.b)[i]; v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w; }
        if (p.c) { const float4 u = reinterpret_cast<const float4*>(p.c)[i]; v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w; }
        if (p.res_f32) { const float4 u = reinterpret_cast<const float4*>(p.res_f32)[i]; v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w; }
        if (p.res_hi) {
            const int64_t ro = p.tri_C ? tri_off(i, p.tri_C) : i * 4;
            const uint2 hh = *reinterpret_cast<const uint2*>(p.res_hi + ro);
            const uint2 ll = *reinterpret_cast<const uint2*>((p.tri_C ? p.res_hi + 2 * p.tri_C : p.res_lo) + ro);
            v.x += __uint_as_float(hh.x << 16) + __uint_as_float(ll.x << 16);
            v.y += __uint_as_float(hh.x & 0xffff0000u) + __uint_as_float(ll.x & 0xffff0000u);
            v.z += __uint_as_float(hh.y << 16) + __uint_as_float(ll.y << 16);
            v.w += __uint_as_float(hh.y & 0xffff0000u) + __uint_as_float(ll.y & 0xffff0000u);
        }
        if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        if (p.out_f32) reinterpret_cast<float4*>(p.out_f32)[i] = v;
        if (p.out_hi) {
            __align__(8) bf16 h[4], l[4];
            split(v.x, h[0], l[0]); split(v.y, h[1], l[1]); split(v.z, h[2], l[2]); split(v.w, h[3], l[3]);
            if (p.tri_C) {
                bf16* d = p.out_hi + tri_off(i, p.tri_C);
                *reinterpret_cast<uint2*>(d) = *reinterpret_cast<const uint2*>(h);
                *reinterpret_cast<uint2*>(d + p.tri_C) = *reinterpret_cast<const uint2*>(h);
                *reinterpret_cast<uint2*>(d + 2 * p.tri_C) = *reinterpret_cast<const uint2*>(l);
            } else {
                reinterpret_cast<uint2*>(p.out_hi)[i] = *reinterpret_cast<const uint2*>(h);
                reinterpret_cast<uint2*>(p.out_lo)[i] = *reinterpret_cast<const uint2*>(l);
            }
        }
    }
}

// (P, H, W, C) fp32 partial sums a + b + c -> ReLU -> max_pool2d(3, 2, 1) -> (P, Ho, Wo, C) hi / lo; 4 channels per thread
__global__ void __launch_bounds__(256) sum3_maxpool_split_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                 const float* __restrict__ c, int64_t P, int H, int W, int C, int Ho, int Wo,
                                                                 int relu, bf16* __restrict__ out_hi, bf16* __restrict__ out_lo, int tri) {
    const int C4 = C >> 2;
    const int64_t total = P * Ho * Wo * C4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        int64_t t = i / C4;
        const int ox = (int)(t % Wo); t /= Wo;
        const int oy = (int)(t % Ho);
        const int64_t pp = t / Ho;
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        for (int dy = -1; dy <= 1; ++dy) {
            const int y = 2 * oy + dy;
            if (y < 0 || y >= H) continue;
            for (int dx = -1; dx <= 1; ++dx) {
                const int x = 2 * ox + dx;
                if (x < 0 || x >= W) continue;
                const int64_t j = ((pp * H + y) * W + x) * C4 + c4;
                float4 v = reinterpret_cast<const float4*>(a)[j];
                const float4 u = reinterpret_cast<const float4*>(b)[j], w = reinterpret_cast<const float4*>(c)[j];
                v.x += u.x + w.x; v.y += u.y + w.y; v.z += u.z + w.z; v.w += u.w + w.w;
                m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
            }
        }
        if (relu) { m.x = fmaxf(m.x, 0.f); m.y = fmaxf(m.y, 0.f); m.z = fmaxf(m.z, 0.f); m.w = fmaxf(m.w, 0.f); }
        __align__(8) bf16 h[4], l[4];
        split(m.x, h[0], l[0]); split(m.y, h[1], l[1]); split(m.z, h[2], l[2]); split(m.w, h[3], l[3]);
        if (tri) {                                            // (pixels, 3C) = [hi | hi | lo]
            bf16* d = out_hi + tri_off(i, C);
            *reinterpret_cast<uint2*>(d) = *reinterpret_cast<const uint2*>(h);
            *reinterpret_cast<uint2*>(d + C) = *reinterpret_cast<const uint2*>(h);
            *reinterpret_cast<uint2*>(d + 2 * C) = *reinterpret_cast<const uint2*>(l);
        } else {
            reinterpret_cast<uint2*>(out_hi)[i] = *reinterpret_cast<const uint2*>(h);
            reinterpret_cast<uint2*>(out_lo)[i] = *reinterpret_cast<const uint2*>(l);
        }
    }
}

int grid_for(int64_t n, int per) {
    int64_t g = (n + per - 1) / per;
    const int64_t cap = (int64_t)ipsb::sm_count() * 16;
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}

}  // namespace

extern "C" {

int ipsb_stage_patches_padded_split(const float* src, int64_t first_row, int64_t n_rows, int C, int H, int W,
                                    void* dst_hi, void* dst_lo, void* stream) {
    IPSB_REQUIRE(src && dst_hi && dst_lo && n_rows > 0 && C >= 1 && C <= 4, "stage_padded_split: bad arguments (C=%d)", C);
    const int Hp = H + 6, Wp = W + 6;
    const int64_t total = n_rows * Hp * Wp;
    stage_padded_split_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, first_row, n_rows, C, H, W, Hp, Wp, 3, 4,
                                                                                      (bf16*)dst_hi, (bf16*)dst_lo);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_sum3_split(const float* a, const float* b, const float* c, const float* res_f32, const void* res_hi, const void* res_lo,
                    int relu, void* out_hi, void* out_lo, float* out_f32, int64_t n, int tri_C, void* stream) {
    IPSB_REQUIRE(a && n > 0 && n % 4 == 0, "sum3_split: n=%lld must be a positive multiple of 4", (long long)n);
    IPSB_REQUIRE(tri_C >= 0 && tri_C % 4 == 0 && (tri_C == 0 || n % tri_C == 0), "sum3_split: tri layout with C=%d", tri_C);
    IPSB_REQUIRE(tri_C > 0 || ((res_hi == nullptr) == (res_lo == nullptr) && (out_hi == nullptr) == (out_lo == nullptr)),
                 "sum3_split: hi / lo buffers come in pairs");
    IPSB_REQUIRE(out_hi || out_f32, "sum3_split: an output is required");
    Sum3Params p{a, b, c, res_f32, (const bf16*)res_hi, (const bf16*)res_lo, (bf16*)out_hi, (bf16*)out_lo, out_f32, n / 4, relu, tri_C};
    sum3_split_kernel<<<grid_for(p.n4, 256), 256, 0, (cudaStream_t)stream>>>(p);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_sum3_maxpool_split(const float* a, const float* b, const float* c, int64_t P, int H, int W, int C, int relu,
                            void* out_hi, void* out_lo, int tri, void* stream) {
    IPSB_REQUIRE(a && b && c && out_hi && (out_lo || tri) && P > 0 && C % 4 == 0, "sum3_maxpool_split: bad arguments");
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const int64_t total = P * Ho * Wo * (C / 4);
    sum3_maxpool_split_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(a, b, c, P, H, W, C, Ho, Wo, relu,
                                                                                      (bf16*)out_hi, (bf16*)out_lo, tri);
    IPSB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
