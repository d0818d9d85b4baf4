// HBM-bound data movement: patch staging (NCHW fp32 -> channels-last), row gathers,
// pooling, LayerNorm rows.  All 128-bit where alignment allows.
#include <stdlib.h>
#include "common.cuh"
#include "pf.cuh"
#include "../../include/ips_b200.h"

namespace {

using bf16 = __nv_bfloat16;

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }

// one thread per output pixel: C strided plane reads (coalesced over pixels), one
// Cpad-wide channels-last store.
template <typename T, int CPAD>
__global__ void stage_kernel(const float* __restrict__ src, const int64_t* __restrict__ row_idx,
                             int64_t first_row, int64_t n_rows, int C, int HW, T* __restrict__ dst) {
    const int64_t total = n_rows * HW;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / HW;
        const int px = (int)(i - r * HW);
        const int64_t srow = row_idx ? row_idx[r] : first_row + r;
        const float* s = src + srow * (int64_t)C * HW + px;
        __align__(16) T out[CPAD];
#pragma unroll
        for (int c = 0; c < CPAD; ++c) out[c] = from_f32<T>(c < C ? __ldg(s + (int64_t)c * HW) : 0.f);
        T* d = dst + i * CPAD;
        if (sizeof(T) * CPAD == 8) {
            *reinterpret_cast<uint2*>(d) = *reinterpret_cast<const uint2*>(out);
        } else if (sizeof(T) * CPAD == 16) {
            *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(out);
        } else {
#pragma unroll
            for (int c = 0; c < CPAD; ++c) d[c] = out[c];
        }
    }
}

// same, into a zero-bordered (Hp, Wp) frame with the image at (pt, pl): feeds the TMA stem.
// One thread = four consecutive frame pixels of one row (pl % 4 == 0, W % 4 == 0): C float4 plane loads
// (coalesced along x) and one 32-byte channels-last store; border quads are written as zeros.
template <int CPAD>
__global__ void stage_padded_kernel(const float* __restrict__ src, const int64_t* __restrict__ row_idx,
                                    int64_t first_row, int64_t n_rows, int C, int H, int W, int Hp, int Wp,
                                    int pt, int pl, bool vec4, bf16* __restrict__ dst) {
    const int qpr = (Wp + 3) >> 2;                              // quads per frame row
    const int64_t total = n_rows * Hp * qpr;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t fr = i / qpr;                              // frame row index = r * Hp + y'
        const int qx = (int)(i - fr * qpr);
        const int64_t r = fr / Hp;
        const int y = (int)(fr - r * Hp) - pt;
        const int x0 = qx * 4 - pl;                              // image x of the quad's first pixel
        __align__(16) bf16 out[4][CPAD];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int c = 0; c < CPAD; ++c) out[k][c] = __float2bfloat16_rn(0.f);
        if (y >= 0 && y < H && x0 + 3 >= 0 && x0 < W) {
            const int64_t srow = row_idx ? row_idx[r] : first_row + r;
            const float* sp = src + srow * (int64_t)C * H * W + (int64_t)y * W + x0;
            if (vec4 && x0 >= 0 && x0 + 3 < W) {                 // aligned interior quad: one float4 per plane
                for (int c = 0; c < C; ++c) {
                    const float4 v = *reinterpret_cast<const float4*>(sp + (int64_t)c * H * W);
                    out[0][c] = __float2bfloat16_rn(v.x); out[1][c] = __float2bfloat16_rn(v.y);
                    out[2][c] = __float2bfloat16_rn(v.z); out[3][c] = __float2bfloat16_rn(v.w);
                }
            } else {
                for (int k = 0; k < 4; ++k)
                    if (x0 + k >= 0 && x0 + k < W)
                        for (int c = 0; c < C; ++c) out[k][c] = __float2bfloat16_rn(__ldg(sp + (int64_t)c * H * W + k));
            }
        }
        bf16* d = dst + (fr * Wp + qx * 4) * CPAD;
        const int npx = min(4, Wp - qx * 4);
        if (npx == 4) {
            *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(&out[0][0]);
            *reinterpret_cast<uint4*>(d + 2 * CPAD) = *reinterpret_cast<const uint4*>(&out[2][0]);
        } else {
            for (int k = 0; k < npx; ++k) *reinterpret_cast<uint2*>(d + k * CPAD) = *reinterpret_cast<const uint2*>(&out[k][0]);
        }
    }
}

// space-to-depth staging for the stem: one thread per 2x2 input block.  Frame pixel (Y', X') of patch r holds 16 bf16:
// channel (dy*2+dx)*4 + c = in(2Y'+dy-4, 2X'+dx-4, c) (zero outside the image / for c >= C); Ys x Wp frame pixels per patch.
// The source is either a tensor of patches (geo.n_cols == 0: patch r starts at r*C*H*W, rows of W floats) or whole images
// (B, C, img_h, img_w) cut into a grid of patches on the fly (ipsb_image_geo: patch n of an image = row n / n_cols,
// column n % n_cols of the grid, at stride (sh, sw) -- the reference's CPU `unfold`, data/*/..._dataset.py).
struct SrcGeo {
    int64_t img_stride, chan_stride, row_stride;    // floats between images / channels / pixel rows
    int n_per_image, n_cols, sh, sw;
    int pair_ok;                                    // 8-byte loads of aligned pixel pairs are safe
};
__device__ __forceinline__ const float* patch_base(const float* src, const SrcGeo& g, int64_t srow, int C, int H, int W) {
    if (g.n_cols == 0) return src + srow * (int64_t)C * H * W;
    const int64_t b = srow / g.n_per_image;
    const int n = (int)(srow - b * g.n_per_image);
    const int pr = n / g.n_cols, pc = n - pr * g.n_cols;
    return src + b * g.img_stride + (int64_t)pr * g.sh * g.row_stride + (int64_t)pc * g.sw;
}

// VEC2: one thread stages TWO horizontally adjacent frame pixels from one 16-byte load per (row, channel).
template <bool VEC2>
__global__ void stage_s2d_kernel(const float* __restrict__ src, const SrcGeo geo, const int64_t* __restrict__ row_idx, int64_t first_row,
                                 int64_t n_rows, int C, int H, int W, int Ys, int Wp, bf16* __restrict__ dst) {
    const int Wt = VEC2 ? (Wp + 1) / 2 : Wp;              // threads per frame row
    const int64_t total = n_rows * Ys * Wt;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / (Ys * Wt);
        const int rem = (int)(i - r * (Ys * Wt));
        const int Yp = rem / Wt, Xt = rem - Yp * Wt;
        const int Xp = VEC2 ? 2 * Xt : Xt;
        __align__(16) bf16 out[VEC2 ? 32 : 16];
#pragma unroll
        for (int k = 0; k < (VEC2 ? 32 : 16); ++k) out[k] = __float2bfloat16_rn(0.f);
        const int y0 = 2 * Yp - 4, x0 = 2 * Xp - 4;
        const int xs = VEC2 ? 4 : 2;                      // input columns covered by this thread
        if (y0 + 1 >= 0 && y0 < H && x0 + xs - 1 >= 0 && x0 < W) {
            const int64_t srow = row_idx ? row_idx[r] : first_row + r;
            const float* sp = patch_base(src, geo, srow, C, H, W);
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                const int y = y0 + dy;
                if (y < 0 || y >= H) continue;
#pragma unroll
                for (int c = 0; c < 4; ++c) {             // compile-time indices keep `out` in registers
                    if (c >= C) break;
                    const float* rowp = sp + (int64_t)c * geo.chan_stride + (int64_t)y * geo.row_stride;
                    if (VEC2) {
                        float v[4] = {0.f, 0.f, 0.f, 0.f};
                        if (x0 >= 0 && x0 + 3 < W) {      // x0 is a multiple of 4: aligned quad
                            const float4 q = *reinterpret_cast<const float4*>(rowp + x0);
                            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (x0 + j >= 0 && x0 + j < W) v[j] = rowp[x0 + j];
                        }
                        out[(dy * 2 + 0) * 4 + c] = __float2bfloat16_rn(v[0]);
                        out[(dy * 2 + 1) * 4 + c] = __float2bfloat16_rn(v[1]);
                        out[16 + (dy * 2 + 0) * 4 + c] = __float2bfloat16_rn(v[2]);
                        out[16 + (dy * 2 + 1) * 4 + c] = __float2bfloat16_rn(v[3]);
                    } else if (geo.pair_ok && x0 >= 0 && x0 + 1 < W) {         // x0 is even: aligned pair
                        const float2 v = *reinterpret_cast<const float2*>(rowp + x0);
                        out[(dy * 2 + 0) * 4 + c] = __float2bfloat16_rn(v.x);
                        out[(dy * 2 + 1) * 4 + c] = __float2bfloat16_rn(v.y);
                    } else {
                        if (x0 >= 0 && x0 < W) out[(dy * 2 + 0) * 4 + c] = __float2bfloat16_rn(rowp[x0]);
                        if (x0 + 1 >= 0 && x0 + 1 < W) out[(dy * 2 + 1) * 4 + c] = __float2bfloat16_rn(rowp[x0 + 1]);
                    }
                }
            }
        }
        uint4* d = reinterpret_cast<uint4*>(dst + ((r * Ys + Yp) * Wp + Xp) * 16);
        d[0] = *reinterpret_cast<const uint4*>(&out[0]);
        d[1] = *reinterpret_cast<const uint4*>(&out[8]);
        if (VEC2 && Xp + 1 < Wp) {
            d[2] = *reinterpret_cast<const uint4*>(&out[16]);
            d[3] = *reinterpret_cast<const uint4*>(&out[24]);
        }
    }
}

// out[b, m, c, y, x] = img[b, c, pr*sh + y, pc*sw + x] for patch n = idx[b, m] (idx == null: n = m, i.e. patchify);
// one block per (patch, channel), threads along the pixels of the patch
__global__ void gather_patches_image_kernel(const float* __restrict__ img, const SrcGeo geo, const int64_t* __restrict__ idx,
                                            int M, int C, int H, int W, float* __restrict__ out) {
    const int64_t pm = blockIdx.x;                       // b * M + m
    const int c = blockIdx.y;
    const int64_t b = pm / M;
    const int64_t n = idx ? idx[pm] : pm - b * M;
    float* d = out + (pm * C + c) * (int64_t)H * W;
    if (n < 0) {                                         // not owned (sharded assembly): zero patch
        for (int i = threadIdx.x; i < H * W; i += blockDim.x) d[i] = 0.f;
        return;
    }
    const float* s = patch_base(img, geo, b * geo.n_per_image + n, C, H, W) + (int64_t)c * geo.chan_stride;
    for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
        const int y = i / W, x = i - y * W;
        d[i] = s[(int64_t)y * geo.row_stride + x];
    }
}

// im2col of a channels-last bf16 tensor for the weight-gradient GEMM: out[m, (r*kw+s)*C + c] = x[p, oy*stride+r-pad,
// ox*stride+s-pad, c] (zero outside), m = (p*Ho + oy)*Wo + ox; columns >= kh*kw*C are zero.  One thread per 16-byte chunk.
__global__ void im2col_bf16_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int64_t M, int H, int W, int C, int Ho, int Wo,
                                   int kh, int kw, int stride, int pad, int Kp) {
    const int chunks = Kp / 8;
    const int K = kh * kw * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M * chunks; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / chunks;
        const int k0 = (int)(i - m * chunks) * 8;
        const int ox = (int)(m % Wo);
        const int64_t t = m / Wo;
        const int oy = (int)(t % Ho);
        const int64_t p = t / Ho;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (C % 8 == 0) {
            if (k0 < K) {
                const int tap = k0 / C, c0 = k0 - tap * C;
                const int r = tap / kw, sx = tap - r * kw;
                const int iy = oy * stride + r - pad, ix = ox * stride + sx - pad;
                if (iy >= 0 && iy < H && ix >= 0 && ix < W)
                    v = *reinterpret_cast<const uint4*>(x + ((p * H + iy) * W + ix) * (int64_t)C + c0);
            }
        } else {
            __align__(16) bf16 e[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = k0 + j;
                float val = 0.f;
                if (k < K) {
                    const int tap = k / C, c = k - tap * C;
                    const int r = tap / kw, sx = tap - r * kw;
                    const int iy = oy * stride + r - pad, ix = ox * stride + sx - pad;
                    if (iy >= 0 && iy < H && ix >= 0 && ix < W) val = __bfloat162float(x[((p * H + iy) * W + ix) * (int64_t)C + c]);
                }
                e[j] = __float2bfloat16_rn(val);
            }
            v = *reinterpret_cast<const uint4*>(e);
        }
        *reinterpret_cast<uint4*>(out + m * Kp + k0) = v;
    }
}

// dst row (b,m) <- src row; grid.x = rows, grid.y = 16 KB segments of a row
__global__ void gather_rows16_kernel(const unsigned char* __restrict__ src, int64_t batch_stride_rows,
                                     const int64_t* __restrict__ idx, int M, int64_t row_bytes,
                                     unsigned char* __restrict__ dst) {
    const int64_t row = blockIdx.x;
    const int64_t b = row / M;
    const int64_t sel = idx[row];
    const int64_t srow = b * batch_stride_rows + (sel < 0 ? 0 : sel);
    const unsigned char* s = src + srow * row_bytes;
    unsigned char* d = dst + row * row_bytes;
    const int64_t seg_lo = (int64_t)blockIdx.y * 16384;
    const int64_t seg_hi = min(seg_lo + 16384, row_bytes);
    int64_t off = seg_lo + (int64_t)threadIdx.x * 16;
    const int64_t step = (int64_t)blockDim.x * 16;
    // 4 independent 128-bit loads in flight per thread
    int4 v[4];
    bool ok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        ok[j] = off + j * step < seg_hi;
        if (ok[j]) v[j] = sel < 0 ? make_int4(0, 0, 0, 0) : ipsb::ld_stream16(s + off + j * step);   // idx < 0: zero row
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (ok[j]) ipsb::st_stream16(d + off + j * step, v[j]);
}

__global__ void gather_rows4_kernel(const uint32_t* __restrict__ src, int64_t batch_stride_rows,
                                    const int64_t* __restrict__ idx, int M, int64_t row_words,
                                    uint32_t* __restrict__ dst) {
    const int64_t row = blockIdx.x;
    const int64_t b = row / M;
    const int64_t sel = idx[row];
    const uint32_t* s = src + (b * batch_stride_rows + (sel < 0 ? 0 : sel)) * row_words;
    uint32_t* d = dst + row * row_words;
    for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < row_words;
         i += (int64_t)gridDim.y * blockDim.x)
        d[i] = sel < 0 ? 0u : s[i];
}

// max_pool2d(3, 2, 1) channels-last, VEC channels per thread (16 bytes)
template <typename T, int VEC>
__global__ void maxpool_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t P, int H, int W, int C,
                               int Ho, int Wo, int oG0, int oWp, int oSp, int iWp, int iSp) {
    const int cv = C / VEC;
    const int64_t total = P * Ho * Wo * cv;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int c0 = (int)(i % cv) * VEC;
        int64_t t = i / cv;
        const int ox = (int)(t % Wo); t /= Wo;
        const int oy = (int)(t % Ho);
        const int64_t p = t / Ho;
        float m[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) m[k] = -INFINITY;
        for (int dy = 0; dy < 3; ++dy) {
            const int iy = oy * 2 - 1 + dy;
            if (iy < 0 || iy >= H) continue;
            for (int dx = 0; dx < 3; ++dx) {
                const int ix = ox * 2 - 1 + dx;
                if (ix < 0 || ix >= W) continue;
                const T* s = x + (p * iSp + (int64_t)iy * iWp + ix) * (int64_t)C + c0;
                __align__(16) T v[VEC];
                *reinterpret_cast<uint4*>(v) = *reinterpret_cast<const uint4*>(s);
#pragma unroll
                for (int k = 0; k < VEC; ++k) m[k] = fmaxf(m[k], to_f32<T>(v[k]));
            }
        }
        __align__(16) T o[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) o[k] = from_f32<T>(m[k]);
        *reinterpret_cast<uint4*>(y + (oG0 + p * oSp + (int64_t)oy * oWp + ox) * (int64_t)C + c0) = *reinterpret_cast<const uint4*>(o);
    }
}

// Backward of max_pool2d(3, 2, 1), channels-last fp32, 4 channels per thread, gather form (deterministic, no atomics): an
// input pixel receives dy of every window (at most 2 x 2) whose maximum it is.  The maximum of a window is its FIRST largest
// element in row-major scan order (strict >), torch's rule (the forward of nn.MaxPool2d keeps that index).
__global__ void maxpool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int64_t P, int H,
                                   int W, int C, int Ho, int Wo) {
    const int cv = C / 4;
    const int64_t total = P * H * W * cv;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c0 = (int)(i % cv) * 4;
        int64_t t = i / cv;
        const int ix = (int)(t % W); t /= W;
        const int iy = (int)(t % H);
        const int64_t p = t / H;
        const float4 me = *reinterpret_cast<const float4*>(x + ((p * H + iy) * W + ix) * (int64_t)C + c0);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int oy0 = max(0, iy / 2), oy1 = min(Ho - 1, (iy + 1) / 2);      // windows with 2 oy - 1 <= iy <= 2 oy + 1
        const int ox0 = max(0, ix / 2), ox1 = min(Wo - 1, (ix + 1) / 2);
        for (int oy = oy0; oy <= oy1; ++oy)
            for (int ox = ox0; ox <= ox1; ++ox) {
                // is (iy, ix) the first maximum of window (oy, ox)?  per channel: nothing before it is >= it, nothing after is > it
                bool w0 = true, w1 = true, w2 = true, w3 = true;
                for (int r = 0; r < 3; ++r) {
                    const int y = oy * 2 - 1 + r;
                    if (y < 0 || y >= H) continue;
                    for (int q = 0; q < 3; ++q) {
                        const int xx = ox * 2 - 1 + q;
                        if (xx < 0 || xx >= W || (y == iy && xx == ix)) continue;
                        const float4 v = *reinterpret_cast<const float4*>(x + ((p * H + y) * W + xx) * (int64_t)C + c0);
                        const bool before = (y < iy) || (y == iy && xx < ix);
                        if (before) { w0 &= !(v.x >= me.x); w1 &= !(v.y >= me.y); w2 &= !(v.z >= me.z); w3 &= !(v.w >= me.w); }
                        else { w0 &= !(v.x > me.x); w1 &= !(v.y > me.y); w2 &= !(v.z > me.z); w3 &= !(v.w > me.w); }
                    }
                }
                const float4 g = *reinterpret_cast<const float4*>(dy + ((p * Ho + oy) * Wo + ox) * (int64_t)C + c0);
                if (w0) acc.x += g.x;
                if (w1) acc.y += g.y;
                if (w2) acc.z += g.z;
                if (w3) acc.w += g.w;
            }
        *reinterpret_cast<float4*>(dx + ((p * H + iy) * W + ix) * (int64_t)C + c0) = acc;
    }
}

// y = relu(a + b); backward mask = y > 0 (residual tail of a BasicBlock), 4 floats per thread
__global__ void add_relu_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ y, int64_t n4) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 u = a[i], v = b[i];
        y[i] = make_float4(fmaxf(u.x + v.x, 0.f), fmaxf(u.y + v.y, 0.f), fmaxf(u.z + v.z, 0.f), fmaxf(u.w + v.w, 0.f));
    }
}
__global__ void relu_bwd_kernel(const float4* __restrict__ y, const float4* __restrict__ dy, float4* __restrict__ dx, int64_t n4) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 u = y[i], g = dy[i];
        dx[i] = make_float4(u.x > 0.f ? g.x : 0.f, u.y > 0.f ? g.y : 0.f, u.z > 0.f ? g.z : 0.f, u.w > 0.f ? g.w : 0.f);
    }
}

template <typename T>
__global__ void avgpool_kernel(const T* __restrict__ x, float* __restrict__ y, int64_t P, int HW, int C) {
    const int64_t total = P * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / C;
        const int c = (int)(i - p * C);
        const T* s = x + p * (int64_t)HW * C + c;
        float acc = 0.f;
        for (int k = 0; k < HW; ++k) acc += to_f32<T>(s[(int64_t)k * C]);
        y[i] = acc / (float)HW;
    }
}

__global__ void avgpool_pf_kernel(const bf16* __restrict__ x, float* __restrict__ y, int64_t P, int Sp, int HW, int C) {
    const int64_t total = P * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / C;
        const int c = (int)(i - p * C);
        const bf16* s = x + p * (int64_t)Sp * C + c;
        float acc = 0.f;
        for (int k = 0; k < Sp; ++k) acc += __bfloat162float(s[(int64_t)k * C]);
        y[i] = acc / (float)HW;
    }
}

// one 256-thread block per row; two-pass mean / biased variance in fp32
template <typename T, bool kNorm>
__global__ void rows_kernel(const float* __restrict__ x, T* __restrict__ y, int F, float eps) {
    __shared__ float red[8];
    __shared__ float stat[2];
    const float* xr = x + (int64_t)blockIdx.x * F;
    T* yr = y + (int64_t)blockIdx.x * F;
    const int tid = threadIdx.x;
    float mean = 0.f, rstd = 1.f;
    if (kNorm) {
        float s = 0.f;
        for (int i = tid; i < F; i += blockDim.x) s += xr[i];
        s = ipsb::warp_sum(s);
        if ((tid & 31) == 0) red[tid >> 5] = s;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
            stat[0] = t / (float)F;
        }
        __syncthreads();
        mean = stat[0];
        float q = 0.f;
        for (int i = tid; i < F; i += blockDim.x) { const float d = xr[i] - mean; q += d * d; }
        q = ipsb::warp_sum(q);
        if ((tid & 31) == 0) red[tid >> 5] = q;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
            stat[1] = rsqrtf(t / (float)F + eps);
        }
        __syncthreads();
        rstd = stat[1];
    }
    for (int i = tid; i < F; i += blockDim.x) yr[i] = from_f32<T>((xr[i] - mean) * rstd);
}

// warp per row, the whole row in registers (V float4 per lane, all loads issued up front): one pass over HBM
template <typename T, bool kNorm, int V>
__global__ void __launch_bounds__(256) rows_warp_kernel(const float* __restrict__ x, T* __restrict__ y, int64_t rows, float eps) {
    constexpr int F = V * 128;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + row * F);
    float4 v[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int4 raw = ipsb::ld_stream16(xr + i * 32 + lane);
        v[i] = *reinterpret_cast<const float4*>(&raw);
    }
    float mean = 0.f, rstd = 1.f;
    if (kNorm) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        mean = ipsb::warp_sum(s) / (float)F;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
        rstd = rsqrtf(ipsb::warp_sum(q) / (float)F + eps);
    }
    T* yr = y + row * F;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        __align__(8) T o[4] = {from_f32<T>((v[i].x - mean) * rstd), from_f32<T>((v[i].y - mean) * rstd),
                               from_f32<T>((v[i].z - mean) * rstd), from_f32<T>((v[i].w - mean) * rstd)};
        if (sizeof(T) == 2) *reinterpret_cast<uint2*>(yr + (i * 32 + lane) * 4) = *reinterpret_cast<const uint2*>(o);
        else *reinterpret_cast<float4*>(yr + (i * 32 + lane) * 4) = *reinterpret_cast<const float4*>(o);
    }
}

// bf16 feature bags (features stored as bf16 halve the bytes of the CAMELYON path): warp per row, the row
// (V x 256 bf16) in registers, LayerNorm statistics in fp32 on the exactly-upcast values -> bf16 GEMM operand
template <bool kNorm, int V>
__global__ void __launch_bounds__(256) rows_warp_bf16in_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int64_t rows, float eps) {
    constexpr int F = V * 256;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int4* xr = reinterpret_cast<const int4*>(x + row * F);
    float v[V][8];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int4 raw = ipsb::ld_stream16(xr + i * 32 + lane);
        const bf16* e = reinterpret_cast<const bf16*>(&raw);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = __bfloat162float(e[j]);
    }
    float mean = 0.f, rstd = 1.f;
    if (kNorm) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) s += v[i][j];
        mean = ipsb::warp_sum(s) / (float)F;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; q += d * d; }
        rstd = rsqrtf(ipsb::warp_sum(q) / (float)F + eps);
    }
    int4* yr = reinterpret_cast<int4*>(y + row * F);
#pragma unroll
    for (int i = 0; i < V; ++i) {
        __align__(16) bf16 o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = __float2bfloat16_rn((v[i][j] - mean) * rstd);
        yr[i * 32 + lane] = *reinterpret_cast<const int4*>(o);
    }
}

// generic width: one block per row, fp32 two-pass statistics
template <bool kNorm>
__global__ void rows_bf16in_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int F, float eps) {
    __shared__ float red[8];
    __shared__ float stat[2];
    const bf16* xr = x + (int64_t)blockIdx.x * F;
    bf16* yr = y + (int64_t)blockIdx.x * F;
    const int tid = threadIdx.x;
    float mean = 0.f, rstd = 1.f;
    if (kNorm) {
        float s = 0.f;
        for (int i = tid; i < F; i += blockDim.x) s += __bfloat162float(xr[i]);
        s = ipsb::warp_sum(s);
        if ((tid & 31) == 0) red[tid >> 5] = s;
        __syncthreads();
        if (tid == 0) { float t = 0.f; for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w]; stat[0] = t / (float)F; }
        __syncthreads();
        mean = stat[0];
        float q = 0.f;
        for (int i = tid; i < F; i += blockDim.x) { const float d = __bfloat162float(xr[i]) - mean; q += d * d; }
        q = ipsb::warp_sum(q);
        if ((tid & 31) == 0) red[tid >> 5] = q;
        __syncthreads();
        if (tid == 0) { float t = 0.f; for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w]; stat[1] = rsqrtf(t / (float)F + eps); }
        __syncthreads();
        rstd = stat[1];
    }
    for (int i = tid; i < F; i += blockDim.x) yr[i] = __float2bfloat16_rn((__bfloat162float(xr[i]) - mean) * rstd);
}

// out[c] = sum_r x[r, c] (bias gradients, BatchNorm statistics); one block per 32 columns, fixed order
__global__ void colsum_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out, int64_t rows, int cols) {
    __shared__ float part[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
    const int R = gridDim.y;                      // row chunks: out is (R, cols) partials when R > 1
    const int64_t per = (rows + R - 1) / R, r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
    float s = 0.f;
    if (c < cols)
        for (int64_t r = r0 + w; r < r1; r += 8) s += y ? x[r * cols + c] * y[r * cols + c] : x[r * cols + c];
    part[w][threadIdx.x & 31] = s;
    __syncthreads();
    if (w == 0 && c < cols) {
        float t = 0.f;
        for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x];
        out[(int64_t)blockIdx.y * cols + c] = t;
    }
}

__global__ void conv_weight_layouts_kernel(const float* __restrict__ w, int Cout, int Cin, int kh, int kw, bf16* __restrict__ w_nk,
                                           bf16* __restrict__ w_t, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        // i indexes w_nk (writes coalesced): co, r, s, ci
        const int ci = (int)(i % Cin);
        int64_t t = i / Cin;
        const int sx = (int)(t % kw); t /= kw;
        const int r = (int)(t % kh);
        const int co = (int)(t / kh);
        const bf16 v = __float2bfloat16_rn(w[(((int64_t)co * Cin + ci) * kh + r) * kw + sx]);
        w_nk[i] = v;
        if (w_t) w_t[((int64_t)ci * kh * kw + (int64_t)(kh - 1 - r) * kw + (kw - 1 - sx)) * Cout + co] = v;
    }
}

// weight gradient from the TN GEMM (K', Cout) fp32, k = (r*kw + s)*C + ci  ->  (Cout, Cin, kh, kw) fp32 (nn.Conv2d layout),
// ci < Cin <= C (the stem's staged input carries a zero 4th channel)
__global__ void wgrad_to_oihw_kernel(const float* __restrict__ dw, int Cout, int Cin, int C, int kh, int kw, float* __restrict__ out,
                                     int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int sx = (int)(i % kw);
        int64_t t = i / kw;
        const int r = (int)(t % kh); t /= kh;
        const int ci = (int)(t % Cin);
        const int co = (int)(t / Cin);
        out[i] = dw[((int64_t)(r * kw + sx) * C + ci) * Cout + co];
    }
}

__global__ void cast_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, int64_t n4) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        __align__(8) bf16 o[4] = {__float2bfloat16_rn(v.x), __float2bfloat16_rn(v.y), __float2bfloat16_rn(v.z), __float2bfloat16_rn(v.w)};
        reinterpret_cast<uint2*>(y)[i] = *reinterpret_cast<const uint2*>(o);
    }
}

int grid_for(int64_t total, int threads) {
    int64_t g = ipsb::ceil_div(total, threads);
    const int64_t cap = (int64_t)ipsb::sm_count() * 16;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

extern "C" {

int ipsb_stage_patches(const float* src, const int64_t* row_idx, int64_t first_row, int64_t n_rows,
                       int C, int H, int W, int Cpad, int dt, void* dst, void* stream) {
    IPSB_REQUIRE(n_rows > 0 && C > 0 && C <= Cpad, "stage: bad shape rows=%lld C=%d Cpad=%d", (long long)n_rows, C, Cpad);
    IPSB_REQUIRE(Cpad == 4, "stage: only Cpad=4 is built (C=%d)", C);
    const int64_t total = n_rows * H * W;
    const int g = grid_for(total, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (dt == IPSB_BF16)
        stage_kernel<bf16, 4><<<g, 256, 0, st>>>(src, row_idx, first_row, n_rows, C, H * W, (bf16*)dst);
    else if (dt == IPSB_F32)
        stage_kernel<float, 4><<<g, 256, 0, st>>>(src, row_idx, first_row, n_rows, C, H * W, (float*)dst);
    else
        return ipsb::fail("stage: unknown dtype %d", dt);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_stage_patches_padded(const float* src, const int64_t* row_idx, int64_t first_row, int64_t n_rows,
                              int C, int H, int W, int pad_top, int pad_left, int Hp, int Wp, void* dst, void* stream) {
    IPSB_REQUIRE(n_rows > 0 && C > 0 && C <= 4 && Hp >= H + pad_top && Wp >= W + pad_left, "stage_padded: bad shape");
    IPSB_REQUIRE(Wp % 2 == 0, "stage_padded: frame width %d must be even", Wp);
    const bool vec4 = W % 4 == 0 && pad_left % 4 == 0 && ((uintptr_t)src % 16 == 0);
    const int64_t total = n_rows * Hp * ((Wp + 3) / 4);
    stage_padded_kernel<4><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        src, row_idx, first_row, n_rows, C, H, W, Hp, Wp, pad_top, pad_left, vec4, (bf16*)dst);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_gather_rows(const void* src, int64_t src_batch_stride_rows, const int64_t* idx,
                     int B, int M, int64_t row_bytes, void* dst, void* stream) {
    IPSB_REQUIRE(B > 0 && M > 0 && row_bytes > 0 && row_bytes % 4 == 0, "gather_rows: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = row_bytes % 16 == 0 && ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0);
    if (vec) {
        dim3 grid((unsigned)(B * (int64_t)M), (unsigned)ipsb::ceil_div(row_bytes, 16384));
        gather_rows16_kernel<<<grid, 256, 0, st>>>((const unsigned char*)src, src_batch_stride_rows, idx, M,
                                                   row_bytes, (unsigned char*)dst);
    } else {
        const int64_t words = row_bytes / 4;
        dim3 grid((unsigned)(B * (int64_t)M), (unsigned)(words > 4096 ? 8 : 1));
        gather_rows4_kernel<<<grid, 256, 0, st>>>((const uint32_t*)src, src_batch_stride_rows, idx, M, words,
                                                  (uint32_t*)dst);
    }
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_maxpool3x3s2(const void* x, void* y, int64_t P, int H, int W, int C, int dt, void* stream) {
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (dt == IPSB_BF16) {
        IPSB_REQUIRE(C % 8 == 0, "maxpool: C=%d not a multiple of 8", C);
        const int64_t total = P * Ho * Wo * (C / 8);
        maxpool_kernel<bf16, 8><<<grid_for(total, 256), 256, 0, st>>>((const bf16*)x, (bf16*)y, P, H, W, C, Ho, Wo, 0, Wo, Ho * Wo, W, H * W);
    } else if (dt == IPSB_F32) {
        IPSB_REQUIRE(C % 4 == 0, "maxpool: C=%d not a multiple of 4", C);
        const int64_t total = P * Ho * Wo * (C / 4);
        maxpool_kernel<float, 4><<<grid_for(total, 256), 256, 0, st>>>((const float*)x, (float*)y, P, H, W, C, Ho, Wo, 0, Wo, Ho * Wo, W, H * W);
    } else {
        return ipsb::fail("maxpool: unknown dtype %d", dt);
    }
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_maxpool3x3s2_bwd_f32(const float* x, const float* dy, float* dx, int64_t P, int H, int W, int C, void* stream) {
    IPSB_REQUIRE(x && dy && dx && P > 0 && C % 4 == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)dy % 16 == 0) && ((uintptr_t)dx % 16 == 0),
                 "maxpool_bwd: C=%d must be a multiple of 4, pointers 16-byte aligned", C);
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    maxpool_bwd_kernel<<<grid_for(P * H * W * (C / 4), 256), 256, 0, (cudaStream_t)stream>>>(x, dy, dx, P, H, W, C, Ho, Wo);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_add_relu_f32(const float* a, const float* b, float* y, int64_t n, void* stream) {
    IPSB_REQUIRE(a && b && y && n > 0 && n % 4 == 0 && ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0) && ((uintptr_t)y % 16 == 0),
                 "add_relu: n=%lld must be a multiple of 4, pointers 16-byte aligned", (long long)n);
    add_relu_kernel<<<grid_for(n / 4, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)a, (const float4*)b, (float4*)y, n / 4);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_relu_bwd_f32(const float* y, const float* dy, float* dx, int64_t n, void* stream) {
    IPSB_REQUIRE(y && dy && dx && n > 0 && n % 4 == 0 && ((uintptr_t)y % 16 == 0) && ((uintptr_t)dy % 16 == 0) && ((uintptr_t)dx % 16 == 0),
                 "relu_bwd: n=%lld must be a multiple of 4, pointers 16-byte aligned", (long long)n);
    relu_bwd_kernel<<<grid_for(n / 4, 256), 256, 0, (cudaStream_t)stream>>>((const float4*)y, (const float4*)dy, (float4*)dx, n / 4);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_maxpool3x3s2_pf_strided(const void* x, void* y, int64_t P, int H, int W, int C, int in_Wp, int in_Sp, void* stream) {
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    IPSB_REQUIRE(C % 8 == 0 && in_Wp >= W && in_Sp >= H * in_Wp - (in_Wp - W), "maxpool_pf: bad arguments (C=%d)", C);
    const pf::Geo g = pf::make(P, Ho, Wo);
    const int64_t total = P * Ho * Wo * (C / 8);
    maxpool_kernel<bf16, 8><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, P, H, W, C,
                                                                                   Ho, Wo, g.G0, g.Wp, g.Sp, in_Wp, in_Sp);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_maxpool3x3s2_pf(const void* x, void* y, int64_t P, int H, int W, int C, void* stream) {
    return ipsb_maxpool3x3s2_pf_strided(x, y, P, H, W, C, W, H * W, stream);
}

int ipsb_stage_patches_s2d(const float* src, const int64_t* row_idx, int64_t first_row, int64_t n_rows, int C, int H, int W,
                           void* dst, void* stream) {
    IPSB_REQUIRE(n_rows > 0 && C > 0 && C <= 4 && H % 2 == 0 && W % 2 == 0 && ((uintptr_t)src % 8 == 0),
                 "stage_s2d: needs C <= 4 and even H, W");
    const int Ys = H / 2 + 3, Wp = W / 2 + 3;
    static const bool no_tma = getenv("IPSB_STAGE_NO_TMA") != nullptr;
    if (!no_tma && ((uintptr_t)dst % 16 == 0) && ipsb_stage_tma_ok(src, C, H, W))      // copy-engine form (stage_tma.cu)
        return ipsb_stage_patches_s2d_tma(src, row_idx, first_row, n_rows, C, H, W, dst, stream);
    const SrcGeo geo{0, (int64_t)H * W, (int64_t)W, 0, 0, 0, 0, 1};
    if (W % 4 == 0 && ((uintptr_t)src % 16 == 0))       // rows and patches start 16-byte aligned: two frame pixels per thread
        stage_s2d_kernel<true><<<grid_for(n_rows * Ys * ((Wp + 1) / 2), 256), 256, 0, (cudaStream_t)stream>>>(
            src, geo, row_idx, first_row, n_rows, C, H, W, Ys, Wp, (bf16*)dst);
    else
        stage_s2d_kernel<false><<<grid_for(n_rows * Ys * Wp, 256), 256, 0, (cudaStream_t)stream>>>(src, geo, row_idx, first_row, n_rows,
                                                                                                  C, H, W, Ys, Wp, (bf16*)dst);
    IPSB_LAUNCH_CHECK();
    return 0;
}

static int make_src_geo(const ipsb_image_geo* g, int C, int H, int W, const float* img, SrcGeo* out) {
    IPSB_REQUIRE(g && g->img_h >= H && g->img_w >= W && g->stride_h > 0 && g->stride_w > 0, "image geometry: bad sizes");
    const int n_rows_grid = (g->img_h - H) / g->stride_h + 1, n_cols = (g->img_w - W) / g->stride_w + 1;
    IPSB_REQUIRE(g->n_per_image == n_rows_grid * n_cols, "image geometry: n_per_image=%d but the grid has %d x %d patches",
                 g->n_per_image, n_rows_grid, n_cols);
    out->chan_stride = (int64_t)g->img_h * g->img_w;
    out->img_stride = out->chan_stride * C;
    out->row_stride = g->img_w;
    out->n_per_image = g->n_per_image; out->n_cols = n_cols; out->sh = g->stride_h; out->sw = g->stride_w;
    out->pair_ok = (g->img_w % 2 == 0) && (g->stride_w % 2 == 0) && ((uintptr_t)img % 8 == 0);
    return 0;
}

int ipsb_stage_image_s2d(const float* img, const ipsb_image_geo* g, int64_t first_row, int64_t n_rows, int C, int H, int W,
                         void* dst, void* stream) {
    IPSB_REQUIRE(n_rows > 0 && C > 0 && C <= 4 && H % 2 == 0 && W % 2 == 0, "stage_image_s2d: needs C <= 4 and even H, W");
    SrcGeo geo;
    if (int rc = make_src_geo(g, C, H, W, img, &geo)) return rc;
    const int Ys = H / 2 + 3, Wp = W / 2 + 3;
    const bool quad = (g->img_w % 4 == 0) && (g->stride_w % 4 == 0) && ((uintptr_t)img % 16 == 0);
    if (quad)
        stage_s2d_kernel<true><<<grid_for(n_rows * Ys * ((Wp + 1) / 2), 256), 256, 0, (cudaStream_t)stream>>>(
            img, geo, nullptr, first_row, n_rows, C, H, W, Ys, Wp, (bf16*)dst);
    else
        stage_s2d_kernel<false><<<grid_for(n_rows * Ys * Wp, 256), 256, 0, (cudaStream_t)stream>>>(img, geo, nullptr, first_row, n_rows,
                                                                                                  C, H, W, Ys, Wp, (bf16*)dst);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_gather_patches_image(const float* img, const ipsb_image_geo* g, const int64_t* idx, int B, int M, int C, int H, int W,
                              float* out, void* stream) {
    IPSB_REQUIRE(B > 0 && M > 0 && C > 0, "gather_patches_image: bad sizes");
    SrcGeo geo;
    if (int rc = make_src_geo(g, C, H, W, img, &geo)) return rc;
    IPSB_REQUIRE(idx != nullptr || M == g->n_per_image, "gather_patches_image: patchify needs M == n_per_image");
    dim3 grid((unsigned)(B * (int64_t)M), (unsigned)C);
    gather_patches_image_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, geo, idx, M, C, H, W, out);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_im2col_bf16(const void* x, void* out, int64_t P, int H, int W, int C, int kh, int kw, int stride, int pad, int Kp, void* stream) {
    const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
    IPSB_REQUIRE(P > 0 && Ho > 0 && Wo > 0 && Kp % 8 == 0 && Kp >= kh * kw * C, "im2col: bad sizes (Kp=%d, K=%d)", Kp, kh * kw * C);
    const int64_t M = P * Ho * Wo;
    im2col_bf16_kernel<<<grid_for(M * (Kp / 8), 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)out, M, H, W, C, Ho, Wo, kh, kw,
                                                                                     stride, pad, Kp);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_avgpool_pf(const void* x, float* y, int64_t P, int H, int W, int C, void* stream) {
    // pad rows are zero: the mean over the patch's Sp rows divided by H*W pixels
    const pf::Geo g = pf::make(P, H, W);
    avgpool_pf_kernel<<<grid_for(P * C, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x + (int64_t)g.G0 * C, y, P, g.Sp, H * W, C);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int64_t ipsb_pf_rows(int64_t P, int H, int W) { return pf::make(P, H, W).rows; }

int ipsb_avgpool(const void* x, float* y, int64_t P, int HW, int C, int dt, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int g = grid_for(P * C, 256);
    if (dt == IPSB_BF16)
        avgpool_kernel<bf16><<<g, 256, 0, st>>>((const bf16*)x, y, P, HW, C);
    else if (dt == IPSB_F32)
        avgpool_kernel<float><<<g, 256, 0, st>>>((const float*)x, y, P, HW, C);
    else
        return ipsb::fail("avgpool: unknown dtype %d", dt);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_layernorm_rows_f32(const float* x, float* y, int64_t rows, int F, float eps, void* stream) {
    IPSB_REQUIRE(rows > 0 && F > 0, "layernorm: bad shape");
    rows_kernel<float, true><<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(x, y, F, eps);
    IPSB_LAUNCH_CHECK();
    return 0;
}

/* out[c] = sum_r x[r,c] * (y ? y[r,c] : 1) */
int ipsb_colsum_f32(const float* x, const float* y, float* out, float* scratch, int64_t rows, int cols, void* stream) {
    IPSB_REQUIRE(rows > 0 && cols > 0, "colsum: bad shape");
    if (rows <= 4096 || scratch == nullptr) {
        colsum_kernel<<<dim3((cols + 31) / 32, 1), 256, 0, (cudaStream_t)stream>>>(x, y, out, rows, cols);
    } else {                                     // long columns: 64 row chunks -> partials -> second pass over the 64 partial rows
        colsum_kernel<<<dim3((cols + 31) / 32, 64), 256, 0, (cudaStream_t)stream>>>(x, y, scratch, rows, cols);
        colsum_kernel<<<dim3((cols + 31) / 32, 1), 256, 0, (cudaStream_t)stream>>>(scratch, nullptr, out, 64, cols);
    }
    IPSB_LAUNCH_CHECK();
    return 0;
}

// Both bf16 operand layouts of a conv weight for the grad-mode step, one launch (replaces five elementwise launches):
//   w_nk[co, (r*kw + s)*Cin + ci]            = w[co, ci, r, s]                      forward / weight-gradient operand
//   w_t [ci, (r*kw + s)*Cout + co]           = w[co, ci, kh-1-r, kw-1-s]            input-gradient operand (flipped, transposed)
int ipsb_conv_weight_layouts(const float* w, int Cout, int Cin, int kh, int kw, void* w_nk, void* w_t, void* stream) {
    IPSB_REQUIRE(w && w_nk && Cout > 0 && Cin > 0 && kh > 0 && kw > 0, "conv_weight_layouts: bad arguments");
    const int64_t n = (int64_t)Cout * Cin * kh * kw;
    conv_weight_layouts_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(w, Cout, Cin, kh, kw, (bf16*)w_nk, (bf16*)w_t, n);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_wgrad_to_oihw(const float* dw_kc, int Cout, int Cin, int C, int kh, int kw, float* out, void* stream) {
    IPSB_REQUIRE(dw_kc && out && Cout > 0 && Cin > 0 && Cin <= C && kh > 0 && kw > 0, "wgrad_to_oihw: bad arguments");
    const int64_t n = (int64_t)Cout * Cin * kh * kw;
    wgrad_to_oihw_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(dw_kc, Cout, Cin, C, kh, kw, out, n);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_cast_bf16(const float* x, void* y, int64_t n, void* stream) {
    IPSB_REQUIRE(n > 0 && n % 4 == 0 && ((uintptr_t)x % 16 == 0), "cast_bf16: n=%lld must be a multiple of 4", (long long)n);
    cast_bf16_kernel<<<grid_for(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(x, (bf16*)y, n / 4);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_rows_to_bf16(const float* x, void* y, int64_t rows, int F, int layernorm, float eps, void* stream) {
    IPSB_REQUIRE(rows > 0 && F > 0, "rows_to_bf16: bad shape");
    if (F == 2048 && ((uintptr_t)x % 16 == 0)) {          // CAMELYON feature width: warp-per-row single-pass kernel
        const unsigned g = (unsigned)((rows + 7) / 8);
        if (layernorm) rows_warp_kernel<bf16, true, 16><<<g, 256, 0, (cudaStream_t)stream>>>(x, (bf16*)y, rows, eps);
        else rows_warp_kernel<bf16, false, 16><<<g, 256, 0, (cudaStream_t)stream>>>(x, (bf16*)y, rows, eps);
        IPSB_LAUNCH_CHECK();
        return 0;
    }
    if (layernorm)
        rows_kernel<bf16, true><<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(x, (bf16*)y, F, eps);
    else
        rows_kernel<bf16, false><<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(x, (bf16*)y, F, eps);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_rows_bf16_to_bf16(const void* x, void* y, int64_t rows, int F, int layernorm, float eps, void* stream) {
    IPSB_REQUIRE(rows > 0 && F > 0, "rows_bf16_to_bf16: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    if (F == 2048 && ((uintptr_t)x % 16 == 0)) {
        const unsigned g = (unsigned)((rows + 7) / 8);
        if (layernorm) rows_warp_bf16in_kernel<true, 8><<<g, 256, 0, st>>>((const bf16*)x, (bf16*)y, rows, eps);
        else rows_warp_bf16in_kernel<false, 8><<<g, 256, 0, st>>>((const bf16*)x, (bf16*)y, rows, eps);
    } else if (layernorm) {
        rows_bf16in_kernel<true><<<(unsigned)rows, 256, 0, st>>>((const bf16*)x, (bf16*)y, F, eps);
    } else {
        rows_bf16in_kernel<false><<<(unsigned)rows, 256, 0, st>>>((const bf16*)x, (bf16*)y, F, eps);
    }
    IPSB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
