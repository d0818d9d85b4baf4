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


// Batch statistics in ONE pass over HBM: a block sums its row chunk, derives the chunk's column means, and takes the
// squared deviations from them in a second sweep over the same chunk (L2 / L1 hot: a chunk is rows / 256 rows);
// partial[chunk] = column sums, partial[R + chunk] = column M2 around the chunk mean.  col_stats_final_kernel merges the
// chunks with the parallel-variance formula  M2 = sum_k M2_k + n_k (mean_k - mean)^2  -- as accurate as the two-pass form.
__global__ void __launch_bounds__(256) col_stats4_kernel(const float* __restrict__ x, float* __restrict__ partial, int64_t rows, int cols) {
    __shared__ float4 p0[256];
    __shared__ float4 cmean[256];
    const int tpr = cols >> 2, rpi = 256 / tpr;
    const int slot = threadIdx.x / tpr, cg = threadIdx.x - slot * tpr;
    const int R = gridDim.x, chunk = blockIdx.x;
    const int64_t per = (rows + R - 1) / R, r0 = chunk * per, r1 = min(rows, r0 + per);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t r = r0 + slot; r < r1; r += rpi) {
        const float4 xv = reinterpret_cast<const float4*>(x)[r * tpr + cg];
        a.x += xv.x; a.y += xv.y; a.z += xv.z; a.w += xv.w;
    }
    p0[threadIdx.x] = a;
    __syncthreads();
    if (slot == 0) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < rpi; ++k) { const float4 u = p0[k * tpr + cg]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
        reinterpret_cast<float4*>(partial + (int64_t)chunk * cols)[cg] = t;
        const float inv = 1.f / (float)max((int64_t)1, r1 - r0);
        cmean[cg] = make_float4(t.x * inv, t.y * inv, t.z * inv, t.w * inv);
    }
    __syncthreads();
    const float4 m = cmean[cg];
    a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t r = r0 + slot; r < r1; r += rpi) {
        const float4 xv = reinterpret_cast<const float4*>(x)[r * tpr + cg];
        const float dx = xv.x - m.x, dy = xv.y - m.y, dz = xv.z - m.z, dw = xv.w - m.w;
        a.x += dx * dx; a.y += dy * dy; a.z += dz * dz; a.w += dw * dw;
    }
    p0[threadIdx.x] = a;
    __syncthreads();
    if (slot == 0) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < rpi; ++k) { const float4 u = p0[k * tpr + cg]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
        reinterpret_cast<float4*>(partial + ((int64_t)R + chunk) * cols)[cg] = t;
    }
}

// merges the chunk statistics; optionally also rstd and the running statistics (bn_finalize) in the same launch.
// A block owns 32 columns, 8 thread groups stride over the chunks; two sweeps (total -> mean, then the merged M2).
__global__ void __launch_bounds__(256) col_stats_final_kernel(const float* __restrict__ partial, int64_t rows, int cols, int R,
                                                              float* __restrict__ mean, float* __restrict__ var, float momentum,
                                                              float unbias, float eps, float* __restrict__ rstd,
                                                              float* __restrict__ running_mean, float* __restrict__ running_var) {
    __shared__ float red[8][33];
    __shared__ float mean_s[32];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    const int64_t per = (rows + R - 1) / R;
    float t = 0.f;
    if (c < cols)
        for (int k = g; k < R; k += 8) t += partial[(int64_t)k * cols + c];
    red[g][lane] = t;
    __syncthreads();
    if (g == 0) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) a += red[j][lane];
        mean_s[lane] = a / (float)rows;
    }
    __syncthreads();
    const float m = mean_s[lane];
    float m2 = 0.f;
    if (c < cols)
        for (int k = g; k < R; k += 8) {
            const int64_t n_k = min(rows, (k + 1) * per) - min(rows, k * per);
            if (n_k <= 0) continue;
            const float d = partial[(int64_t)k * cols + c] / (float)n_k - m;
            m2 += partial[((int64_t)R + k) * cols + c] + (float)n_k * d * d;
        }
    __syncthreads();
    red[g][lane] = m2;
    __syncthreads();
    if (g == 0 && c < cols) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) a += red[j][lane];
        const float v = a / (float)rows;
        mean[c] = m;
        var[c] = v;
        if (rstd) rstd[c] = rsqrtf(v + eps);
        if (running_mean) running_mean[c] = running_mean[c] * (1.f - momentum) + m * momentum;
        if (running_var) running_var[c] = running_var[c] * (1.f - momentum) + v * (momentum * unbias);
    }
}

// out[q * cols + c] = scale * sum_chunk partial[(q * R + chunk) * cols + c]
// A block owns 32 columns of one q: 8 thread groups stride over the chunks (coalesced 128-byte rows of the partial table,
// independent loads), the groups are summed in a fixed order (deterministic).  (One thread per column walking all R
// chunks was a chain of R dependent L2 round trips: 20 us per launch, sixty launches per train step.)
__global__ void __launch_bounds__(256) col_final_kernel(const float* __restrict__ partial, float* __restrict__ out, int cols, int R, int nq,
                                                        float scale) {
    __shared__ float red[8][33];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int cblocks = (cols + 31) / 32;
    const int q = blockIdx.x / cblocks, c = (blockIdx.x - q * cblocks) * 32 + lane;
    float t = 0.f;
    if (c < cols)
        for (int k = g; k < R; k += 8) t += partial[((int64_t)q * R + k) * cols + c];
    red[g][lane] = t;
    __syncthreads();
    if (g == 0 && c < cols) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) a += red[j][lane];
        out[(int64_t)q * cols + c] = a * scale;
    }
}

__global__ void bn_finalize_kernel(const float* __restrict__ mean, const float* __restrict__ var, int cols, float momentum, float unbias,
                                   float eps, float* __restrict__ rstd, float* __restrict__ running_mean, float* __restrict__ running_var) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols) return;
    const float m = mean[j], v = var[j];
    rstd[j] = rsqrtf(v + eps);
    if (running_mean) running_mean[j] = running_mean[j] * (1.f - momentum) + m * momentum;
    if (running_var) running_var[j] = running_var[j] * (1.f - momentum) + v * (momentum * unbias);
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
                                     const float4* __restrict__ sums, float4* __restrict__ dx, int64_t rows, int tpr, int relu,
                                     int64_t rows_total) {
    const int64_t n4 = rows * tpr;
    const float inv = 1.f / (float)rows_total;      // rows of the whole (possibly multi-rank) batch the sums were taken over
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
                                    const float* __restrict__ sums, float* __restrict__ dx, int64_t rows, int cols, int relu,
                                    int64_t rows_total) {
    const int64_t n = rows * cols;
    const float inv = 1.f / (float)rows_total;
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
// softmax(q k^T) (dropout) v per (batch, head, query token), parallel over the M selected patches: a CTA per (b, h, t)
// owns the softmax statistics, CTAs per (b, h, chunk of 256 patches) do the per-patch work (one warp per patch).
// probabilities p (B,H,T,M) are kept for the backward; `mask` (same shape, 0/1 floats) and `keep_scale` = 1/(1-p_drop)
// implement nn.Dropout on the attention weights.  Every reduction has a fixed order (deterministic).
constexpr int kAttnChunk = 256;

// logits s[b,h,t,m] = q_t . k_m -> prob buffer; grid (B*H, chunks), warp per patch
__global__ void __launch_bounds__(256)
attn_logits_kernel(const float* __restrict__ q, const float* __restrict__ k, float* __restrict__ prob, int M, int H, int Dk, int T) {
    extern __shared__ float sm[];                       // [T][Dk]
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < T * Dk; i += blockDim.x) sm[i] = q[(int64_t)(i / Dk) * H * Dk + h * Dk + (i % Dk)];
    __syncthreads();
    const int m0 = blockIdx.y * kAttnChunk;
    const float* kb = k + (int64_t)b * M * H * Dk + h * Dk;
    for (int m = m0 + warp; m < min(M, m0 + kAttnChunk); m += 8) {
        const float k0 = lane < Dk ? kb[(int64_t)m * H * Dk + lane] : 0.f;
        const float k1 = lane + 32 < Dk ? kb[(int64_t)m * H * Dk + lane + 32] : 0.f;
        for (int t = 0; t < T; ++t) {
            float s = (lane < Dk ? sm[t * Dk + lane] * k0 : 0.f) + (lane + 32 < Dk ? sm[t * Dk + lane + 32] * k1 : 0.f);
            s = ipsb::warp_sum(s);
            if (lane == 0) prob[(((int64_t)b * H + h) * T + t) * M + m] = s;
        }
    }
}

// in-place softmax over M of one (b,h,t) row; one CTA per row
__global__ void __launch_bounds__(256) attn_softmax_kernel(float* __restrict__ prob, int M) {
    __shared__ float red[8];
    float* pr = prob + (int64_t)blockIdx.x * M;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float mx = -INFINITY;
    for (int m = tid; m < M; m += 256) mx = fmaxf(mx, pr[m]);
    mx = ipsb::warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float den = 0.f;
    for (int m = tid; m < M; m += 256) { const float e = expf(pr[m] - mx); pr[m] = e; den += e; }
    den = ipsb::warp_sum(den);
    if (lane == 0) red[warp] = den;
    __syncthreads();
    den = 0.f;
    for (int w = 0; w < 8; ++w) den += red[w];
    const float inv = 1.f / den;
    for (int m = tid; m < M; m += 256) pr[m] *= inv;
}

// partial out[(b,h,t), chunk, :] = sum_{m in chunk} p~_m v_m; grid (B*H, chunks); part (B*H*T, chunks, Dv)
__global__ void __launch_bounds__(256)
attn_pv_kernel(const float* __restrict__ v, const float* __restrict__ mask, float keep_scale, const float* __restrict__ prob,
               float* __restrict__ part, int M, int H, int Dv, int T, int chunks) {
    extern __shared__ float sm[];                       // [8][T][Dv] per-warp accumulators
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.y * kAttnChunk;
    const float* vb = v + (int64_t)b * M * H * Dv + h * Dv;
    for (int t = 0; t < T; ++t) {
        const int64_t row = (((int64_t)b * H + h) * T + t) * M;
        float a0 = 0.f, a1 = 0.f;
        for (int m = m0 + warp; m < min(M, m0 + kAttnChunk); m += 8) {
            const float pt = prob[row + m] * (mask ? mask[row + m] * keep_scale : 1.f);
            if (lane < Dv) a0 = fmaf(pt, vb[(int64_t)m * H * Dv + lane], a0);
            if (lane + 32 < Dv) a1 = fmaf(pt, vb[(int64_t)m * H * Dv + lane + 32], a1);
        }
        if (lane < Dv) sm[(warp * T + t) * Dv + lane] = a0;
        if (lane + 32 < Dv) sm[(warp * T + t) * Dv + lane + 32] = a1;
    }
    __syncthreads();
    for (int i = tid; i < T * Dv; i += 256) {
        float a = 0.f;
        for (int w = 0; w < 8; ++w) a += sm[w * T * Dv + i];
        const int t = i / Dv, d = i - t * Dv;
        part[((((int64_t)b * H + h) * T + t) * chunks + blockIdx.y) * Dv + d] = a;
    }
}

// out[b,t,h*Dv+d] = sum_chunk part; one thread per output element
__global__ void attn_out_kernel(const float* __restrict__ part, float* __restrict__ out, int B, int H, int Dv, int T, int chunks) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * H * T * Dv) return;
    const int d = i % Dv, t = (i / Dv) % T, h = (i / (Dv * T)) % H, b = i / (Dv * T * H);
    float a = 0.f;
    for (int c = 0; c < chunks; ++c) a += part[((((int64_t)b * H + h) * T + t) * chunks + c) * Dv + d];
    out[((int64_t)b * T + t) * H * Dv + h * Dv + d] = a;
}

// backward: grid (B*H, chunks), warp per patch.  With dp_m = (dout . v_m) mask_m keep_scale the softmax backward needs
// dot_t = sum_m p_m dp_m = dout_t . out_t (out = the forward result), so no reduction over M precedes the per-patch work.
// dk (B,M,H*Dk), dv (B,M,H*Dv) are written once per patch; dq partials land in dq_part (B*chunks, T*H*Dk), summed by the caller.
__global__ void __launch_bounds__(256)
attn_train_bwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                      const float* __restrict__ mask, float keep_scale, const float* __restrict__ prob,
                      const float* __restrict__ out, const float* __restrict__ dout, float* __restrict__ dq_part,
                      float* __restrict__ dk, float* __restrict__ dv, int M, int H, int Dk, int Dv, int T, int chunks) {
    extern __shared__ float sm[];
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* qs = sm;                      // [T][Dk]
    float* dos = qs + T * Dk;            // [T][Dv]
    float* dots = dos + T * Dv;          // [T]
    float* accs = dots + T;              // [8][T][Dk]
    for (int i = tid; i < T * Dk; i += 256) qs[i] = q[(int64_t)(i / Dk) * H * Dk + h * Dk + (i % Dk)];
    for (int i = tid; i < T * Dv; i += 256) dos[i] = dout[((int64_t)b * T + i / Dv) * H * Dv + h * Dv + (i % Dv)];
    __syncthreads();
    if (warp < T || T > 8) {
        for (int t = warp; t < T; t += 8) {
            float s = 0.f;
            for (int d = lane; d < Dv; d += 32) s = fmaf(dos[t * Dv + d], out[((int64_t)b * T + t) * H * Dv + h * Dv + d], s);
            s = ipsb::warp_sum(s);
            if (lane == 0) dots[t] = s;
        }
    }
    __syncthreads();
    const int m0 = blockIdx.y * kAttnChunk;
    const float* kb = k + (int64_t)b * M * H * Dk + h * Dk;
    const float* vb = v + (int64_t)b * M * H * Dv + h * Dv;
    float* dkb = dk + (int64_t)b * M * H * Dk + h * Dk;
    float* dvb = dv + (int64_t)b * M * H * Dv + h * Dv;
    for (int i = lane; i < T * Dk; i += 32) accs[warp * T * Dk + i] = 0.f;
    __syncwarp();
    for (int m = m0 + warp; m < min(M, m0 + kAttnChunk); m += 8) {
        const float v0 = lane < Dv ? vb[(int64_t)m * H * Dv + lane] : 0.f;
        const float v1 = lane + 32 < Dv ? vb[(int64_t)m * H * Dv + lane + 32] : 0.f;
        const float k0 = lane < Dk ? kb[(int64_t)m * H * Dk + lane] : 0.f;
        const float k1 = lane + 32 < Dk ? kb[(int64_t)m * H * Dk + lane + 32] : 0.f;
        float dv0 = 0.f, dv1 = 0.f, dk0 = 0.f, dk1 = 0.f;
        for (int t = 0; t < T; ++t) {
            const int64_t pi = (((int64_t)b * H + h) * T + t) * M + m;
            float s = (lane < Dv ? dos[t * Dv + lane] * v0 : 0.f) + (lane + 32 < Dv ? dos[t * Dv + lane + 32] * v1 : 0.f);
            s = ipsb::warp_sum(s);
            const float ms = mask ? mask[pi] * keep_scale : 1.f;
            const float p = prob[pi];
            const float ds = p * (s * ms - dots[t]);
            const float pt = p * ms;
            if (lane < Dv) dv0 = fmaf(pt, dos[t * Dv + lane], dv0);
            if (lane + 32 < Dv) dv1 = fmaf(pt, dos[t * Dv + lane + 32], dv1);
            if (lane < Dk) { dk0 = fmaf(ds, qs[t * Dk + lane], dk0); accs[(warp * T + t) * Dk + lane] += ds * k0; }
            if (lane + 32 < Dk) { dk1 = fmaf(ds, qs[t * Dk + lane + 32], dk1); accs[(warp * T + t) * Dk + lane + 32] += ds * k1; }
        }
        if (lane < Dv) dvb[(int64_t)m * H * Dv + lane] = dv0;
        if (lane + 32 < Dv) dvb[(int64_t)m * H * Dv + lane + 32] = dv1;
        if (lane < Dk) dkb[(int64_t)m * H * Dk + lane] = dk0;
        if (lane + 32 < Dk) dkb[(int64_t)m * H * Dk + lane + 32] = dk1;
    }
    __syncthreads();
    for (int i = tid; i < T * Dk; i += 256) {
        float a = 0.f;
        for (int w = 0; w < 8; ++w) a += accs[w * T * Dk + i];
        const int t = i / Dk, d = i - t * Dk;
        dq_part[((int64_t)b * chunks + blockIdx.y) * T * H * Dk + (int64_t)t * H * Dk + h * Dk + d] = a;
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
        col_final_kernel<<<(cols + 31) / 32, 256, 0, st>>>(scratch, mean, cols, R, 1, 1.f / (float)rows);
        col_partial4_kernel<1><<<R, 256, 0, st>>>(x, nullptr, nullptr, mean, nullptr, scratch, rows, cols, 0);
        col_final_kernel<<<(cols + 31) / 32, 256, 0, st>>>(scratch, var, cols, R, 1, 1.f / (float)rows);
        IPSB_LAUNCH_CHECK();
        return 0;
    }
    dim3 grid((cols + 31) / 32, kRowChunks);
    col_partial_kernel<0><<<grid, 256, 0, st>>>(x, nullptr, nullptr, nullptr, nullptr, scratch, rows, cols, 0);
    col_final_kernel<<<(cols + 31) / 32, 256, 0, st>>>(scratch, mean, cols, kRowChunks, 1, 1.f / (float)rows);
    col_partial_kernel<1><<<grid, 256, 0, st>>>(x, nullptr, nullptr, mean, nullptr, scratch, rows, cols, 0);
    col_final_kernel<<<(cols + 31) / 32, 256, 0, st>>>(scratch, var, cols, kRowChunks, 1, 1.f / (float)rows);
    IPSB_LAUNCH_CHECK();
    return 0;
}

/* Batch statistics + (optionally: rstd, rstd != NULL) + (optionally: running statistics) in two launches; x is read from HBM
 * once.  scratch: 2 * 256 * cols floats.  Column counts outside the float4 form fall back to the two-pass kernels. */
int ipsb_bn_stats_finalize_f32(const float* x, float* mean, float* var, float* rstd, float* running_mean, float* running_var,
                               float momentum, float unbias, float eps, float* scratch, int64_t rows, int cols, void* stream) {
    IPSB_REQUIRE(rows > 0 && cols > 0 && scratch != nullptr && mean && var, "bn_stats_finalize: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (fast4(cols, x)) {
        const int R = (int)(rows < kRowChunks4 ? rows : kRowChunks4);
        col_stats4_kernel<<<R, 256, 0, st>>>(x, scratch, rows, cols);
        col_stats_final_kernel<<<(cols + 31) / 32, 256, 0, st>>>(scratch, rows, cols, R, mean, var, momentum, unbias, eps, rstd,
                                                                  running_mean, running_var);
        IPSB_LAUNCH_CHECK();
        return 0;
    }
    if (int rc = ipsb_bn_stats_f32(x, mean, var, scratch, rows, cols, stream)) return rc;
    if (rstd || running_mean || running_var) {
        IPSB_REQUIRE(rstd != nullptr, "bn_stats_finalize: rstd required with running statistics");
        return ipsb_bn_finalize_f32(mean, var, cols, momentum, unbias, eps, rstd, running_mean, running_var, stream);
    }
    return 0;
}

/* tail of the batch statistics: rstd = 1/sqrt(var + eps) and the running-statistics update of nn.BatchNorm
 * (running_var takes the unbiased variance: unbias = rows / (rows - 1)), one launch instead of six elementwise ones */
int ipsb_bn_finalize_f32(const float* mean, const float* var, int cols, float momentum, float unbias, float eps, float* rstd,
                         float* running_mean, float* running_var, void* stream) {
    IPSB_REQUIRE(mean && var && rstd && cols > 0, "bn_finalize: bad arguments");
    bn_finalize_kernel<<<(cols + 255) / 256, 256, 0, (cudaStream_t)stream>>>(mean, var, cols, momentum, unbias, eps, rstd, running_mean, running_var);
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

/* phase 1: sums (2*cols) = [sum g, sum g*xhat] over this rank's rows (= dbeta, dgamma) */
int ipsb_bn_backward_sums_f32(const float* dy, const float* x, const float* y, const float* mean, const float* rstd,
                              float* sums, float* scratch, int64_t rows, int cols, int relu, void* stream) {
    IPSB_REQUIRE(rows > 0 && cols > 0 && scratch != nullptr, "bn_backward: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (fast4(cols, x) && ((uintptr_t)dy % 16 == 0) && ((uintptr_t)y % 16 == 0) && ((uintptr_t)mean % 16 == 0) &&
        ((uintptr_t)rstd % 16 == 0)) {
        const int R = (int)(rows < kRowChunks4 ? rows : kRowChunks4);
        col_partial4_kernel<2><<<R, 256, 0, st>>>(x, dy, y, mean, rstd, scratch, rows, cols, relu);
        col_final_kernel<<<2 * ((cols + 31) / 32), 256, 0, st>>>(scratch, sums, cols, R, 2, 1.f);
    } else {
        dim3 grid((cols + 31) / 32, kRowChunks);
        col_partial_kernel<2><<<grid, 256, 0, st>>>(x, dy, y, mean, rstd, scratch, rows, cols, relu);
        col_final_kernel<<<2 * ((cols + 31) / 32), 256, 0, st>>>(scratch, sums, cols, kRowChunks, 2, 1.f);
    }
    IPSB_LAUNCH_CHECK();
    return 0;
}

/* phase 2: dx from sums taken over rows_total rows (the all-reduced sums of a synchronised BatchNorm, or phase 1's) */
int ipsb_bn_backward_apply_f32(const float* dy, const float* x, const float* y, const float* mean, const float* rstd, const float* gamma,
                               const float* sums, float* dx, int64_t rows, int cols, int64_t rows_total, int relu, void* stream) {
    IPSB_REQUIRE(rows > 0 && cols > 0 && rows_total >= rows, "bn_backward_apply: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (cols % 4 == 0 && ((uintptr_t)dy % 16 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) && ((uintptr_t)dx % 16 == 0) &&
        ((uintptr_t)mean % 16 == 0) && ((uintptr_t)rstd % 16 == 0) && ((uintptr_t)gamma % 16 == 0) && ((uintptr_t)sums % 16 == 0))
        bn_bwd_apply4_kernel<<<grid_for(rows * (cols / 4)), 256, 0, st>>>(
            (const float4*)dy, (const float4*)x, (const float4*)y, (const float4*)mean, (const float4*)rstd, (const float4*)gamma,
            (const float4*)sums, (float4*)dx, rows, cols / 4, relu, rows_total);
    else
        bn_bwd_apply_kernel<<<grid_for(rows * cols), 256, 0, st>>>(dy, x, y, mean, rstd, gamma, sums, dx, rows, cols, relu, rows_total);
    IPSB_LAUNCH_CHECK();
    return 0;
}

/* sums (2*cols): [sum g, sum g*xhat]  (= dbeta, dgamma);  dx as in the BatchNorm backward */
int ipsb_bn_backward_f32(const float* dy, const float* x, const float* y, const float* mean, const float* rstd, const float* gamma,
                         float* sums, float* dx, float* scratch, int64_t rows, int cols, int relu, void* stream) {
    if (int rc = ipsb_bn_backward_sums_f32(dy, x, y, mean, rstd, sums, scratch, rows, cols, relu, stream)) return rc;
    return ipsb_bn_backward_apply_f32(dy, x, y, mean, rstd, gamma, sums, dx, rows, cols, rows, relu, stream);
}

int ipsb_layernorm_backward_f32(const float* dy, const float* x, const float* gamma, float* dx, float* xhat, int64_t rows, int D,
                                float eps, void* stream) {
    IPSB_REQUIRE(rows > 0 && D > 0, "layernorm_backward: bad shape");
    layernorm_bwd_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(dy, x, gamma, dx, xhat, D, eps);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int64_t ipsb_attention_chunks(int M) { return (M + kAttnChunk - 1) / kAttnChunk; }

/* scratch: B*H*T*chunks*Dv floats (chunks = ipsb_attention_chunks(M)) */
int ipsb_attention_train_fwd_f32(const float* q_scaled, const float* k, const float* v, const float* mask, float keep_scale,
                                 float* prob, float* out, float* scratch, int B, int M, int H, int Dk, int Dv, int T, void* stream) {
    IPSB_REQUIRE(B > 0 && M > 0 && Dk <= 64 && Dv <= 64 && scratch != nullptr, "attention_train_fwd: bad arguments (Dk, Dv <= 64)");
    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = (M + kAttnChunk - 1) / kAttnChunk;
    dim3 grid(B * H, chunks);
    attn_logits_kernel<<<grid, 256, (size_t)T * Dk * 4, st>>>(q_scaled, k, prob, M, H, Dk, T);
    attn_softmax_kernel<<<B * H * T, 256, 0, st>>>(prob, M);
    attn_pv_kernel<<<grid, 256, (size_t)8 * T * Dv * 4, st>>>(v, mask, keep_scale, prob, scratch, M, H, Dv, T, chunks);
    attn_out_kernel<<<(B * H * T * Dv + 255) / 256, 256, 0, st>>>(scratch, out, B, H, Dv, T, chunks);
    IPSB_LAUNCH_CHECK();
    return 0;
}

/* out: the forward result (B,T,H*Dv); dq_part: (B*chunks, T*H*Dk), to be summed over its rows by the caller */
int ipsb_attention_train_bwd_f32(const float* q_scaled, const float* k, const float* v, const float* mask, float keep_scale,
                                 const float* prob, const float* out, const float* dout, float* dq_part, float* dk, float* dv,
                                 int B, int M, int H, int Dk, int Dv, int T, void* stream) {
    IPSB_REQUIRE(B > 0 && M > 0 && Dk <= 64 && Dv <= 64 && out != nullptr, "attention_train_bwd: bad arguments (Dk, Dv <= 64)");
    const int chunks = (M + kAttnChunk - 1) / kAttnChunk;
    const size_t smem = (size_t)(T * Dk + T * Dv + T + 8 * T * Dk) * sizeof(float);
    IPSB_REQUIRE(smem <= 48 * 1024, "attention_train_bwd: T=%d too large", T);
    attn_train_bwd_kernel<<<dim3(B * H, chunks), 256, smem, (cudaStream_t)stream>>>(q_scaled, k, v, mask, keep_scale, prob, out, dout,
                                                                                     dq_part, dk, dv, M, H, Dk, Dv, T, chunks);
    IPSB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
