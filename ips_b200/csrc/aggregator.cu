// Forward kernels of the cross-attention aggregator and the heads (inference / no-grad):
//   architecture/transformer.py:85-109 (softmax(Q K^T / sqrt(D_k)) V per head and query token),
//   :104-107 and :127-130 (residual + LayerNorm), architecture/ips_net.py:72-81 (softmax / sigmoid heads).
// The projections themselves are the library's GEMM kernels.
#include "common.cuh"
#include "../../include/ips_b200.h"

namespace {

// one CTA per (b, h): T query tokens against M keys / values of head h; online softmax in fp32
__global__ void __launch_bounds__(256)
cross_attention_kernel(const float* __restrict__ q,   // (T, H*Dk), already divided by sqrt(Dk)
                       const float* __restrict__ k,   // (B, M, H*Dk)
                       const float* __restrict__ v,   // (B, M, H*Dv)
                       float* __restrict__ out,       // (B, T, H*Dv)
                       int M, int H, int Dk, int Dv, int T) {
    extern __shared__ float sm[];                     // qs[T*Dk], red[...]
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int tid = threadIdx.x, nthreads = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    float* qs = sm;                                   // [T][Dk]
    float* wmax = qs + T * Dk;                        // [nwarps]
    float* wsum = wmax + nwarps;                      // [nwarps]
    float* wacc = wsum + nwarps;                      // [nwarps][Dv]
    for (int i = tid; i < T * Dk; i += nthreads) qs[i] = q[(i / Dk) * H * Dk + h * Dk + (i % Dk)];
    __syncthreads();
    const float* kb = k + (int64_t)b * M * H * Dk + h * Dk;
    const float* vb = v + (int64_t)b * M * H * Dv + h * Dv;
    for (int t = 0; t < T; ++t) {
        // each warp walks keys m = warp, warp+nwarps, ...; lanes split the Dk / Dv dimension
        float mx = -INFINITY, den = 0.f;
        float acc0 = 0.f, acc1 = 0.f;                 // Dv <= 64: lane owns columns lane, lane+32
        for (int m = warp; m < M; m += nwarps) {
            float s = 0.f;
            for (int d = lane; d < Dk; d += 32) s = fmaf(qs[t * Dk + d], kb[(int64_t)m * H * Dk + d], s);
            s = ipsb::warp_sum(s);
            const float nm = fmaxf(mx, s);
            const float corr = expf(mx - nm), p = expf(s - nm);
            den = den * corr + p;
            acc0 = acc0 * corr + p * (lane < Dv ? vb[(int64_t)m * H * Dv + lane] : 0.f);
            acc1 = acc1 * corr + p * (lane + 32 < Dv ? vb[(int64_t)m * H * Dv + lane + 32] : 0.f);
            mx = nm;
        }
        if (lane == 0) { wmax[warp] = mx; wsum[warp] = den; }
        __syncthreads();
        float gmx = -INFINITY;
        for (int w = 0; w < nwarps; ++w) gmx = fmaxf(gmx, wmax[w]);
        const float scale = (mx == -INFINITY) ? 0.f : expf(mx - gmx);
        if (lane < Dv) wacc[warp * Dv + lane] = acc0 * scale;
        if (lane + 32 < Dv) wacc[warp * Dv + lane + 32] = acc1 * scale;
        __syncthreads();
        if (tid < Dv) {
            float gden = 0.f, a = 0.f;
            for (int w = 0; w < nwarps; ++w) {
                gden += (wmax[w] == -INFINITY) ? 0.f : wsum[w] * expf(wmax[w] - gmx);
                a += wacc[w * Dv + tid];
            }
            out[((int64_t)b * T + t) * H * Dv + h * Dv + tid] = a / gden;
        }
        __syncthreads();
    }
}

// y = LayerNorm(x + r) * gamma + beta, rows of D; r has r_rows rows and is broadcast (row % r_rows)
__global__ void residual_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ r, int r_rows,
                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                          float* __restrict__ y, int D, float eps) {
    __shared__ float red[8];
    __shared__ float stat[2];
    const int row = blockIdx.x, tid = threadIdx.x;
    const float* xr = x + (int64_t)row * D;
    const float* rr = r ? r + (int64_t)(row % r_rows) * D : nullptr;
    float s = 0.f;
    for (int i = tid; i < D; i += blockDim.x) s += xr[i] + (rr ? rr[i] : 0.f);
    s = ipsb::warp_sum(s);
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) { float t = 0.f; for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w]; stat[0] = t / (float)D; }
    __syncthreads();
    const float mean = stat[0];
    float qv = 0.f;
    for (int i = tid; i < D; i += blockDim.x) { const float d = xr[i] + (rr ? rr[i] : 0.f) - mean; qv += d * d; }
    qv = ipsb::warp_sum(qv);
    if ((tid & 31) == 0) red[tid >> 5] = qv;
    __syncthreads();
    if (tid == 0) { float t = 0.f; for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w]; stat[1] = rsqrtf(t / (float)D + eps); }
    __syncthreads();
    const float rstd = stat[1];
    for (int i = tid; i < D; i += blockDim.x)
        y[(int64_t)row * D + i] = (xr[i] + (rr ? rr[i] : 0.f) - mean) * rstd * gamma[i] + beta[i];
}

// rows of n logits -> softmax (act 0) or sigmoid (act 1); one thread per row (n_class is small)
__global__ void head_activation_kernel(const float* __restrict__ zl, float* __restrict__ y, int rows, int n, int act) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const float* zr = zl + (int64_t)row * n;
    float* yr = y + (int64_t)row * n;
    if (act == 1) {
        for (int i = 0; i < n; ++i) yr[i] = 1.f / (1.f + expf(-zr[i]));
        return;
    }
    float mx = -INFINITY;
    for (int i = 0; i < n; ++i) mx = fmaxf(mx, zr[i]);
    float den = 0.f;
    for (int i = 0; i < n; ++i) den += expf(zr[i] - mx);
    for (int i = 0; i < n; ++i) yr[i] = expf(zr[i] - mx) / den;
}

// Head activation + loss + its gradient in one pass (get_preds, ips_net.py:157-166, + compute_loss,
// training/iterative.py:83-98).  act 0: p = softmax(z), loss_b = -log(p[y_b] + eps) (NLLLoss of log(p + eps), mean over
// the batch); act 1: p = sigmoid(z), BCELoss(p, t) with PyTorch's log clamp at -100, mean over all B*n elements.
// One block; warp w owns rows w, w + nwarps, ...; per-warp partial sums are combined in a fixed order (deterministic).
__global__ void head_loss_kernel(const float* __restrict__ zl, const int64_t* __restrict__ cls, const float* __restrict__ tgt,
                                 int rows, int n, int act, float eps, float* __restrict__ loss, float* __restrict__ dz,
                                 float* __restrict__ probs) {
    __shared__ float part[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    float acc = 0.f;
    for (int row = warp; row < rows; row += nwarps) {
        const float* zr = zl + (int64_t)row * n;
        if (act == 0) {
            float mx = -INFINITY;
            for (int i = lane; i < n; i += 32) mx = fmaxf(mx, zr[i]);
            mx = ipsb::warp_max(mx);
            float den = 0.f;
            for (int i = lane; i < n; i += 32) den += expf(zr[i] - mx);
            den = ipsb::warp_sum(den);
            const int y = (int)cls[row];
            const float py = expf(zr[y] - mx) / den;
            acc += -logf(py + eps);
            const float g = py / (py + eps) / (float)rows;              // d loss / d z_j = g * (p_j - [j == y])
            for (int i = lane; i < n; i += 32) {
                const float pj = expf(zr[i] - mx) / den;
                if (probs) probs[(int64_t)row * n + i] = pj;
                if (dz) dz[(int64_t)row * n + i] = g * (pj - (i == y ? 1.f : 0.f));
            }
        } else {
            float a = 0.f;
            for (int i = lane; i < n; i += 32) {
                const float p = 1.f / (1.f + expf(-zr[i]));
                const float t = tgt[(int64_t)row * n + i];
                const float lp = fmaxf(logf(p), -100.f), lq = fmaxf(logf(1.f - p), -100.f);
                a += -(t * lp + (1.f - t) * lq);
                if (probs) probs[(int64_t)row * n + i] = p;
                if (dz) {                                                // clamped branches have zero slope
                    const float gp = (logf(p) > -100.f ? t * (1.f - p) : 0.f) - (logf(1.f - p) > -100.f ? (1.f - t) * p : 0.f);
                    dz[(int64_t)row * n + i] = -gp / ((float)rows * (float)n);
                }
            }
            acc += ipsb::warp_sum(a) / (float)n;
        }
    }
    if (lane == 0) part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int w = 0; w < nwarps; ++w) tot += part[w];
        *loss = tot / (float)rows;
    }
}

__global__ void add_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = a[i] + b[i];
}

}  // namespace

extern "C" {

int ipsb_cross_attention_f32(const float* q_scaled, const float* k, const float* v, float* out,
                             int B, int M, int H, int Dk, int Dv, int T, void* stream) {
    IPSB_REQUIRE(B > 0 && M > 0 && H > 0 && Dk > 0 && Dv > 0 && Dv <= 64 && T > 0, "cross_attention: bad shape (Dv <= 64)");
    const int threads = 256, nwarps = threads / 32;
    const size_t smem = (size_t)(T * Dk + 2 * nwarps + nwarps * Dv) * sizeof(float);
    cross_attention_kernel<<<B * H, threads, smem, (cudaStream_t)stream>>>(q_scaled, k, v, out, M, H, Dk, Dv, T);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_residual_layernorm_f32(const float* x, const float* r, int r_rows, const float* gamma, const float* beta,
                                float* y, int64_t rows, int D, float eps, void* stream) {
    IPSB_REQUIRE(rows > 0 && D > 0 && (r == nullptr || r_rows > 0), "residual_layernorm: bad shape");
    residual_layernorm_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(x, r, r_rows, gamma, beta, y, D, eps);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_head_activation_f32(const float* logits, float* y, int rows, int n, int act, void* stream) {
    IPSB_REQUIRE(rows > 0 && n > 0 && (act == 0 || act == 1), "head_activation: bad arguments");
    head_activation_kernel<<<(rows + 127) / 128, 128, 0, (cudaStream_t)stream>>>(logits, y, rows, n, act);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_head_loss_f32(const float* logits, const int64_t* class_idx, const float* targets, int rows, int n, int act, float eps,
                       float* loss, float* dlogits, float* probs, void* stream) {
    IPSB_REQUIRE(rows > 0 && n > 0 && loss != nullptr && (act == 0 ? class_idx != nullptr : (act == 1 && targets != nullptr)),
                 "head_loss: bad arguments");
    head_loss_kernel<<<1, 512, 0, (cudaStream_t)stream>>>(logits, class_idx, targets, rows, n, act, eps, loss, dlogits, probs);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_add_f32(const float* a, const float* b, float* y, int64_t n, void* stream) {
    IPSB_REQUIRE(n > 0, "add: bad size");
    int64_t g = (n + 255) / 256;
    if (g > (int64_t)ipsb::sm_count() * 16) g = (int64_t)ipsb::sm_count() * 16;
    add_rows_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(a, b, y, n);
    IPSB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
