// Forward/backward kernels of the non-GEMM operators of the grad-mode step (fp32):
//   BatchNorm1d in batch-statistics mode (+ReLU)      architecture/ips_net.py:58-59 under net.train()
//   LayerNorm with affine                              architecture/transformer.py:107,130
//   multi-head cross-attention core with dropout mask  architecture/transformer.py:29-41,98
// The GEMMs around them are ipsb_gemm_bf16 / ipsb_gemm_f32.
#include "common.cuh"
#include "../../include/ips_b200.h"

namespace {

// ---- column reductions over rows, two stages (deterministic): grid (cols/32, R row chunks) -> partials -> final ----
constexpr int kRowChunks = 64;
constexpr int kRowChunks4 = 256;      // float4 kernels: more, smaller chunks (scratch: 2 * 256 * cols floats)

// partial[(q * R + chunk) * cols + c] for q in {0,1}:
//   MODE 0: sum x                      MODE 1: sum (x - mean)^2
//   MODE 2: sum g, sum g * xhat  with g = dy * (y > 0 if relu), xhat = (x - mean) * rstd
template <int MODE>
__global__ void col_partial_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ y,
                                   const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ partial,
                                   int64_t rows, int cols, int relu) {
    __shared__ float p0[8][33], p1[8][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, c = blockIdx.x * 32 + lane;
    const int R = gridDim.y, chunk = blockIdx.y;
    const int64_t per = (rows + R - 1) / R, r0 = chunk * per, r1 = min(rows, r0 + per);
    float a = 0.f, b = 0.f;
    if (c < cols) {
        const float m = (MODE >= 1) ? mean[c] : 0.f;
        const float rs = (MODE == 2) ? rstd[c] : 0.f;
        for (int64_t r = r0 + w; r < r1; r += 8) {
            const int64_t i = r * cols + c;
            if (MODE == 0) a += x[i];
            else if (MODE == 1) { const float d = x[i] - m; a += d * d; }
            else {
                const float g = (relu && !(y[i] > 0.f)) ? 0.f : dy[i];
                a += g;
                b += g * (x[i] - m) * rs;
            }
        }
    }
    p0[w][lane] = a; p1[w][lane] = b;
    __syncthreads();
    if (w == 0 && c < cols) {
        float t0 = 0.f, t1 = 0.f;
        for (int k = 0; k < 8; ++k) { t0 += p0[k][lane]; t1 += p1[k][lane]; }
        partial[(int64_t)chunk * cols + c] = t0;
        if (MODE == 2) partial[((int64_t)R + chunk) * cols + c] = t1;
    }
}

// float4 version for cols in {4 .. 1024} with (cols / 4) dividing 256: a block reads 256 / (cols / 4) whole rows per
// iteration (fully coalesced 4 KB), every thread owns one group of 4 columns; same partial layout as above.
template <int MODE>
__global__ void __launch_bounds__(256) col_partial4_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                           const float* __restrict__ y, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, float* __restrict__ partial,
                                                           int64_t rows, int cols, int relu) {
    __shared__ float4 p0[256], p1[256];
    const int tpr = cols >> 2, rpi = 256 / tpr;                  // threads per row, rows per iteration
    const int slot = threadIdx.x / tpr, cg = threadIdx.x - slot * tpr;
    const int R = gridDim.x, chunk = blockIdx.x;
    const int64_t per = (rows + R - 1) / R, r0 = chunk * per, r1 = min(rows, r0 + per);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    float4 m = a, rs = a;
    if (MODE >= 1) m = reinterpret_cast<const float4*>(mean)[cg];
    if (MODE == 2) rs = reinterpret_cast<const float4*>(rstd)[cg];
    for (int64_t r = r0 + slot; r < r1; r += rpi) {
        const int64_t i = r * tpr + cg;
        const float4 xv = reinterpret_cast<const float4*>(x)[i];
        if (MODE == 0) { a.x += xv.x; a.y += xv.y; a.z += xv.z; a.w += xv.w; }
        else if (MODE == 1) {
            const float dx = xv.x - m.x, dy_ = xv.y - m.y, dz = xv.z - m.z, dw = xv.w - m.w;
            a.x += dx * dx; a.y += dy_ * dy_; a.z += dz * dz; a.w += dw * dw;
        } else {
            float4 g = reinterpret_cast<const float4*>(dy)[i];
            if (relu) {
                const float4 yv = reinterpret_cast<const float4*>(y)[i];
                if (!(yv.x > 0.f)) g.x = 0.f;
                if (!(yv.y > 0.f)) g.y = 0.f;
                if (!(yv.z > 0.f)) g.z = 0.f;
                if (!(yv.w > 0.f)) g.w = 0.f;
            }
            a.x += g.x; a.y += g.y; a.z += g.z; a.w += g.w;
            b.x += g.x * (xv.x - m.x) * rs.x; b.y += g.y * (xv.y - m.y) * rs.y;
            b.z += g.z * (xv.z - m.z) * rs.z; b.w += g.w * (xv.w - m.w) * rs.w;
        }
    }
    p0[threadIdx.x] = a; p1[threadIdx.x] = b;
    __syncthreads();
    if (slot == 0) {                                             // fixed order over the row slots: deterministic
        float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0;
        for (int k = 0; k < rpi; ++k) {
            const float4 u = p0[k * tpr + cg], v = p1[k * tpr + cg];
            t0.x += u.x; t0.y += u.y; t0.z += u.z; t0.w += u.w;
            t1.x += v.x; t1.y += v.y; t1.z += v.z; t1.w += v.w;
        }
        reinterpret_cast<float4*>(partial + (int64_t)chunk * cols)[cg] = t0;
        if (MODE == 2) reinterpret_cast<float4*>(partial + ((int64_t)R + chunk) * cols)[cg] = t1;
    }
}


// out[q * cols + c] = scale * sum_chunk partial[(q * R + chunk) * cols + c]
__global__ void col_final_kernel(const float* __restrict__ partial, float* __restrict__ out, int cols, int R, int nq, float scale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq * cols) return;
    const int q = i / cols, c = i - q * cols;
    float t = 0.f;
    for (int k = 0; k < R; ++k) t += partial[((int64_t)q * R + k) * cols + c];
    out[i] = t * scale;
}

// y = act(gamma * (x - mean) * rstd + beta)
__global__ void bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                                const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ y,
                                int64_t n, int cols, int relu) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % cols);
        float v = (x[i] - mean[c]) * rstd[c] * gamma[c] + beta[c];
        if (relu) v = fmaxf(v, 0.f);
        y[i] = v;
    }
}

__global__ void bn_apply4_kernel(const float4* __restrict__ x, const float4* __restrict__ mean, const float4* __restrict__ rstd,
                                 const float4* __restrict__ gamma, const float4* __restrict__ beta, float4* __restrict__ y,
                                 int64_t n4, int tpr, int relu) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % tpr);
        const float4 xv = x[i], m = mean[c], r = rstd[c], g = gamma[c], b = beta[c];
        float4 v = make_float4((xv.x - m.x) * r.x * g.x + b.x, (xv.y - m.y) * r.y * g.y + b.y, (xv.z - m.z) * r.z * g.z + b.z,
                               (xv.w - m.w) * r.w * g.w + b.w);
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        y[i] = v;
    }
}

__global__ void bn_bwd_apply4_kernel(const float4* __restrict__ dy, const float4* __restrict__ x, const float4* __restrict__ y,
                                     const float4* __restrict__ mean, const float4* __restrict__ rstd, const float4* __restrict__ gamma,
                                     const float4* __restrict__ sums, float4* __restrict__ dx, int64_t rows, int tpr, int relu) {
    const int64_t n4 = rows * tpr;
    const float inv = 1.f / (float)rows;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % tpr);
        float4 g = dy[i];
        const float4 xv = x[i], m = mean[c], r = rstd[c], ga = gamma[c], s0 = sums[c], s1 = sums[tpr + c];
        if (relu) {
            const float4 yv = y[i];
            if (!(yv.x > 0.f)) g.x = 0.f;
            if (!(yv.y > 0.f)) g.y = 0.f;
            if (!(yv.z > 0.f)) g.z = 0.f;
            if (!(yv.w > 0.f)) g.w = 0.f;
        }
        float4 o;
        o.x = ga.x * r.x * (g.x - s0.x * inv - (xv.x - m.x) * r.x * s1.x * inv);
        o.y = ga.y * r.y * (g.y - s0.y * inv - (xv.y - m.y) * r.y * s1.y * inv);
        o.z = ga.z * r.z * (g.z - s0.z * inv - (xv.z - m.z) * r.z * s1.z * inv);
        o.w = ga.w * r.w * (g.w - s0.w * inv - (xv.w - m.w) * r.w * s1.w * inv);
        dx[i] = o;
    }
}

// dx = gamma * rstd * (g - sum_g / R - xhat * sum_gx / R)
__global__ void bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ y,
                                    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                                    const float* __restrict__ sums, float* __restrict__ dx, int64_t rows, int cols, int relu) {
    const int64_t n = rows * cols;
    const float inv = 1.f / (float)rows;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % cols);
        const float g = (relu && !(y[i] > 0.f)) ? 0.f : dy[i];
        const float xh = (x[i] - mean[c]) * rstd[c];
        dx[i] = gamma[c] * rstd[c] * (g - sums[c] * inv - xh * sums[cols + c] * inv);
    }
}

// ---- LayerNorm backward: one block per row; dgamma / dbeta are column sums done by the caller ------
__global__ void layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                                     float* __restrict__ dx, float* __restrict__ xhat_out, int D, float eps) {
    __shared__ float red[8];
    __shared__ float stat[4];
    const int row = blockIdx.x, tid = threadIdx.x, nw = blockDim.x >> 5;
    const float* xr = x + (int64_t)row * D;
    const float* gr = dy + (int64_t)row * D;
    auto block_sum = [&](float v) {
        v = ipsb::warp_sum(v);
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = v;
        __syncthreads();
        float t = 0.f;
        for (int w = 0; w < nw; ++w) t += red[w];
        return t;
    };
    float s = 0.f;
    for (int i = tid; i < D; i += blockDim.x) s += xr[i];
    const float mean = block_sum(s) / (float)D;
    float q = 0.f;
    for (int i = tid; i < D; i += blockDim.x) { const float d = xr[i] - mean; q += d * d; }
    const float rstd = rsqrtf(block_sum(q) / (float)D + eps);
    float a = 0.f, b = 0.f;
    for (int i = tid; i < D; i += blockDim.x) {
        const float g = gr[i] * (gamma ? gamma[i] : 1.f), xh = (xr[i] - mean) * rstd;
        a += g; b += g * xh;
    }
    const float ma = block_sum(a) / (float)D;
    const float mb = block_sum(b) / (float)D;
    (void)stat;
    for (int i = tid; i < D; i += blockDim.x) {
        const float g = gr[i] * (gamma ? gamma[i] : 1.f), xh = (xr[i] - mean) * rstd;
        dx[(int64_t)row * D + i] = rstd * (g - ma - xh * mb);
        if (xhat_out) xhat_out[(int64_t)row * D + i] = xh;
    }
}

// ---- cross-attention core, training ------------------------------------------------------------------
// one CTA per (b, h).  probabilities p (B,H,T,M) are written for the backward; `mask` (same shape, 0/1 floats) and
// `keep_scale` = 1/(1-p_drop) implement nn.Dropout on the attention weights.
__global__ void __launch_bounds__(256)
attn_train_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                      const float* __restrict__ mask, float keep_scale, float* __restrict__ prob, float* __restrict__ out,
                      int M, int H, int Dk, int Dv, int T) {
    extern __shared__ float sm[];
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int tid = threadIdx.x, nthreads = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nthreads >> 5;
    float* qs = sm;                 // [Dk]
    float* red = qs + Dk;           // [nw]
    float* accs = red + nw;         // [nw][Dv]
    const float* kb = k + (int64_t)b * M * H * Dk + h * Dk;
    const float* vb = v + (int64_t)b * M * H * Dv + h * Dv;
    for (int t = 0; t < T; ++t) {
        float* pr = prob + (((int64_t)b * H + h) * T + t) * M;
        const float* mk = mask ? mask + (((int64_t)b * H + h) * T + t) * M : nullptr;
        __syncthreads();
        for (int i = tid; i < Dk; i += nthreads) qs[i] = q[(int64_t)t * H * Dk + h * Dk + i];
        __syncthreads();
        // pass 1: logits -> prob buffer, running max
        float mx = -INFINITY;
        for (int m = warp; m < M; m += nw) {
            float s = 0.f;
            for (int d = lane; d < Dk; d += 32) s = fmaf(qs[d], kb[(int64_t)m * H * Dk + d], s);
            s = ipsb::warp_sum(s);
            if (lane == 0) pr[m] = s;
            mx = fmaxf(mx, s);
        }
        if (lane == 0) red[warp] = mx;
        __syncthreads();
        mx = red[0];
        for (int w = 1; w < nw; ++w) mx = fmaxf(mx, red[w]);
        __syncthreads();
        float den = 0.f;
        for (int m = tid; m < M; m += nthreads) { const float e = expf(pr[m] - mx); pr[m] = e; den += e; }
        den = ipsb::warp_sum(den);
        if (lane == 0) red[warp] = den;
        __syncthreads();
        den = 0.f;
        for (int w = 0; w < nw; ++w) den += red[w];
        __syncthreads();
        for (int m = tid; m < M; m += nthreads) pr[m] = pr[m] / den;
        __syncthreads();
        // pass 2: out = sum_m p~ v
        float a0 = 0.f, a1 = 0.f;
        for (int m = warp; m < M; m += nw) {
            const float pt = pr[m] * (mk ? mk[m] * keep_scale : 1.f);
            if (lane < Dv) a0 = fmaf(pt, vb[(int64_t)m * H * Dv + lane], a0);
            if (lane + 32 < Dv) a1 = fmaf(pt, vb[(int64_t)m * H * Dv + lane + 32], a1);
        }
        if (lane < Dv) accs[warp * Dv + lane] = a0;
        if (lane + 32 < Dv) accs[warp * Dv + lane + 32] = a1;
        __syncthreads();
        if (tid < Dv) {
            float a = 0.f;
            for (int w = 0; w < nw; ++w) a += accs[w * Dv + tid];
            out[((int64_t)b * T + t) * H * Dv + h * Dv + tid] = a;
        }
    }
}

// backward: dq_part (B,T,H*Dk) per batch element (summed over b by the caller), dk (B,M,H*Dk), dv (B,M,H*Dv)
__global__ void __launch_bounds__(256)
attn_train_bwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                      const float* __restrict__ mask, float keep_scale, const float* __restrict__ prob,
                      const float* __restrict__ dout, float* __restrict__ dq_part, float* __restrict__ dk, float* __restrict__ dv,
                      int M, int H, int Dk, int Dv, int T) {
    extern __shared__ float sm[];
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int tid = threadIdx.x, nthreads = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nthreads >> 5;
    float* qs = sm;                 // [Dk]
    float* dos = qs + Dk;           // [Dv]
    float* red = dos + Dv;          // [nw]
    float* accs = red + nw;         // [nw][Dk]
    const float* kb = k + (int64_t)b * M * H * Dk + h * Dk;
    const float* vb = v + (int64_t)b * M * H * Dv + h * Dv;
    float* dkb = dk + (int64_t)b * M * H * Dk + h * Dk;
    float* dvb = dv + (int64_t)b * M * H * Dv + h * Dv;
    for (int t = 0; t < T; ++t) {
        const float* pr = prob + (((int64_t)b * H + h) * T + t) * M;
        const float* mk = mask ? mask + (((int64_t)b * H + h) * T + t) * M : nullptr;
        __syncthreads();
        for (int i = tid; i < Dk; i += nthreads) qs[i] = q[(int64_t)t * H * Dk + h * Dk + i];
        for (int i = tid; i < Dv; i += nthreads) dos[i] = dout[((int64_t)b * T + t) * H * Dv + h * Dv + i];
        __syncthreads();
        // pass 1: dot = sum_m p_m dp_m with dp_m = (dout . v_m) * mask_m * keep_scale
        float dot = 0.f;
        for (int m = warp; m < M; m += nw) {
            float s = 0.f;
            for (int d = lane; d < Dv; d += 32) s = fmaf(dos[d], vb[(int64_t)m * H * Dv + d], s);
            s = ipsb::warp_sum(s);
            dot += pr[m] * s * (mk ? mk[m] * keep_scale : 1.f);
        }
        if (lane == 0) red[warp] = dot;
        __syncthreads();
        dot = 0.f;
        for (int w = 0; w < nw; ++w) dot += red[w];
        // pass 2: dv, dk, dq
        float dq0 = 0.f, dq1 = 0.f;
        for (int m = warp; m < M; m += nw) {
            float s = 0.f;
            for (int d = lane; d < Dv; d += 32) s = fmaf(dos[d], vb[(int64_t)m * H * Dv + d], s);
            s = ipsb::warp_sum(s);
            const float ms = mk ? mk[m] * keep_scale : 1.f;
            const float p = pr[m];
            const float ds = p * (s * ms - dot);
            const float pt = p * ms;
            for (int d = lane; d < Dv; d += 32) {
                const float add = pt * dos[d];
                dvb[(int64_t)m * H * Dv + d] = (t == 0 ? 0.f : dvb[(int64_t)m * H * Dv + d]) + add;
            }
            for (int d = lane; d < Dk; d += 32) {
                const float add = ds * qs[d];
                dkb[(int64_t)m * H * Dk + d] = (t == 0 ? 0.f : dkb[(int64_t)m * H * Dk + d]) + add;
            }
            if (lane < Dk) dq0 = fmaf(ds, kb[(int64_t)m * H * Dk + lane], dq0);
            if (lane + 32 < Dk) dq1 = fmaf(ds, kb[(int64_t)m * H * Dk + lane + 32], dq1);
        }
        if (lane < Dk) accs[warp * Dk + lane] = dq0;
        if (lane + 32 < Dk) accs[warp * Dk + lane + 32] = dq1;
        __syncthreads();
        if (tid < Dk) {
            float a = 0.f;
            for (int w = 0; w < nw; ++w) a += accs[w * Dk + tid];
            dq_part[((int64_t)b * T + t) * H * Dk + h * Dk + tid] = a;
        }
    }
}

int grid_for(int64_t n) {
    int64_t g = (n + 255) / 256;
    const int64_t cap = (int64_t)ipsb::sm_count() * 16;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

extern "C" {

// float4 kernels need 16-byte aligned rows and a column-group count that divides the block
static bool fast4(int cols, const void* x) { return cols % 4 == 0 && cols <= 1024 && 256 % (cols / 4) == 0 && ((uintptr_t)x % 16 == 0); }

/* scratch: 2 * 256 * cols floats */
int ipsb_bn_stats_f32(const float* x, float* mean, float* var, float* scratch, int64_t rows, int cols, void* stream) {
    IPSB_REQUIRE(rows > 0 && cols > 0 && scratch != nullptr, "bn_stats: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (fast4(cols, x)) {                                         // float4 path: 256 row chunks
        const int R = (int)(rows < kRowChunks4 ? rows : kRowChunks4);
        col_partial4_kernel<0><<<R, 256, 0, st>>>(x, nullptr, nullptr, nullptr, nullptr, scratch, rows, cols, 0);
        col_final_kernel<<<(cols + 255) / 256, 256, 0, st>>>(scratch, mean, cols, R, 1, 1.f / (float)rows);
        col_partial4_kernel<1><<<R, 256, 0, st>>>(x, nullptr, nullptr, mean, nullptr, scratch, rows, cols, 0);
        col_final_kernel<<<(cols + 255) / 256, 256, 0, st>>>(scratch, var, cols, R, 1, 1.f / (float)rows);
        IPSB_LAUNCH_CHECK();
        return 0;
    }
    dim3 grid((cols + 31) / 32, kRowChunks);
    col_partial_kernel<0><<<grid, 256, 0, st>>>(x, nullptr, nullptr, nullptr, nullptr, scratch, rows, cols, 0);
    col_final_kernel<<<(cols + 255) / 256, 256, 0, st>>>(scratch, mean, cols, kRowChunks, 1, 1.f / (float)rows);
    col_partial_kernel<1><<<grid, 256, 0, st>>>(x, nullptr, nullptr, mean, nullptr, scratch, rows, cols, 0);
    col_final_kernel<<<(cols + 255) / 256, 256, 0, st>>>(scratch, var, cols, kRowChunks, 1, 1.f / (float)rows);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_bn_apply_f32(const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta, float* y,
                      int64_t rows, int cols, int relu, void* stream) {
    if (cols % 4 == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) && ((uintptr_t)mean % 16 == 0) &&
        ((uintptr_t)rstd % 16 == 0) && ((uintptr_t)gamma % 16 == 0) && ((uintptr_t)beta % 16 == 0)) {
        bn_apply4_kernel<<<grid_for(rows * (cols / 4)), 256, 0, (cudaStream_t)stream>>>(
            (const float4*)x, (const float4*)mean, (const float4*)rstd, (const float4*)gamma, (const float4*)beta, (float4*)y,
            rows * (cols / 4), cols / 4, relu);
        IPSB_LAUNCH_CHECK();
        return 0;
    }
    bn_apply_kernel<<<grid_for(rows * cols), 256, 0, (cudaStream_t)stream>>>(x, mean, rstd, gamma, beta, y, rows * cols, cols, relu);
    IPSB_LAUNCH_CHECK();
    return 0;
}

/* sums (2*cols): [sum g, sum g*xhat]  (= dbeta, dgamma);  dx as in the BatchNorm backward */
int ipsb_bn_backward_f32(const float* dy, const float* x, const float* y, const float* mean, const float* rstd, const float* gamma,
                         float* sums, float* dx, float* scratch, int64_t rows, int cols, int relu, void* stream) {
    IPSB_REQUIRE(rows > 0 && cols > 0 && scratch != nullptr, "bn_backward: bad arguments");
    if (fast4(cols, x) && ((uintptr_t)dy % 16 == 0) && ((uintptr_t)y % 16 == 0) && ((uintptr_t)dx % 16 == 0) &&
        ((uintptr_t)mean % 16 == 0) && ((uintptr_t)rstd % 16 == 0) && ((uintptr_t)gamma % 16 == 0) && ((uintptr_t)sums % 16 == 0)) {
        cudaStream_t st = (cudaStream_t)stream;
        const int R = (int)(rows < kRowChunks4 ? rows : kRowChunks4);
        col_partial4_kernel<2><<<R, 256, 0, st>>>(x, dy, y, mean, rstd, scratch, rows, cols, relu);
        col_final_kernel<<<(2 * cols + 255) / 256, 256, 0, st>>>(scratch, sums, cols, R, 2, 1.f);
        bn_bwd_apply4_kernel<<<grid_for(rows * (cols / 4)), 256, 0, st>>>(
            (const float4*)dy, (const float4*)x, (const float4*)y, (const float4*)mean, (const float4*)rstd, (const float4*)gamma,
            (const float4*)sums, (float4*)dx, rows, cols / 4, relu);
        IPSB_LAUNCH_CHECK();
        return 0;
    }
    dim3 grid((cols + 31) / 32, kRowChunks);
    col_partial_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(x, dy, y, mean, rstd, scratch, rows, cols, relu);
    col_final_kernel<<<(2 * cols + 255) / 256, 256, 0, (cudaStream_t)stream>>>(scratch, sums, cols, kRowChunks, 2, 1.f);
    IPSB_LAUNCH_CHECK();
    bn_bwd_apply_kernel<<<grid_for(rows * cols), 256, 0, (cudaStream_t)stream>>>(dy, x, y, mean, rstd, gamma, sums, dx, rows, cols, relu);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_layernorm_backward_f32(const float* dy, const float* x, const float* gamma, float* dx, float* xhat, int64_t rows, int D,
                                float eps, void* stream) {
    IPSB_REQUIRE(rows > 0 && D > 0, "layernorm_backward: bad shape");
    layernorm_bwd_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(dy, x, gamma, dx, xhat, D, eps);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_attention_train_fwd_f32(const float* q_scaled, const float* k, const float* v, const float* mask, float keep_scale,
                                 float* prob, float* out, int B, int M, int H, int Dk, int Dv, int T, void* stream) {
    IPSB_REQUIRE(B > 0 && M > 0 && Dk <= 64 && Dv <= 64, "attention_train_fwd: bad shape (Dk, Dv <= 64)");
    const size_t smem = (size_t)(Dk + 8 + 8 * Dv) * sizeof(float);
    attn_train_fwd_kernel<<<B * H, 256, smem, (cudaStream_t)stream>>>(q_scaled, k, v, mask, keep_scale, prob, out, M, H, Dk, Dv, T);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_attention_train_bwd_f32(const float* q_scaled, const float* k, const float* v, const float* mask, float keep_scale,
                                 const float* prob, const float* dout, float* dq_part, float* dk, float* dv,
                                 int B, int M, int H, int Dk, int Dv, int T, void* stream) {
    IPSB_REQUIRE(B > 0 && M > 0 && Dk <= 64 && Dv <= 64, "attention_train_bwd: bad shape (Dk, Dv <= 64)");
    const size_t smem = (size_t)(Dk + Dv + 8 + 8 * Dk) * sizeof(float);
    attn_train_bwd_kernel<<<B * H, 256, smem, (cudaStream_t)stream>>>(q_scaled, k, v, mask, keep_scale, prob, dout, dq_part, dk, dv,
                                                                     M, H, Dk, Dv, T);
    IPSB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
