// fp32 CUDA-core path ("exact" mode): implicit-GEMM convolution and dense linear
// layers on one 64x64x16 register-tiled engine, plus the score basis / logit kernels.
// Accumulation is a single fp32 chain per output in ascending k, the closest a GPU
// gets to the reference's fp32 CPU operators; used for index-exact parity and as the
// on-device yardstick for the tcgen05 kernels.
#include "common.cuh"
#include "../../include/ips_b200.h"

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

struct ConvGeom {
    int H, W, Cin, Ho, Wo, kh, kw, stride, pad;
};

// A-tile loaders: each of 256 threads fetches one float4 = 4 consecutive k of one row.
struct DenseA {
    const float* a;
    int K;
    __device__ __forceinline__ float4 load(int64_t row, int64_t M, int k) const {
        if (row >= M || k >= K) return make_float4(0.f, 0.f, 0.f, 0.f);
        return *reinterpret_cast<const float4*>(a + row * K + k);
    }
};
// A stored (K, M): the contraction index is the slow one (weight-gradient products)
struct DenseAT {
    const float* a;
    int K;
    int64_t lda;          // = M
    __device__ __forceinline__ float4 load(int64_t row, int64_t M, int k) const {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row >= M) return v;
        if (k + 0 < K) v.x = a[(int64_t)(k + 0) * lda + row];
        if (k + 1 < K) v.y = a[(int64_t)(k + 1) * lda + row];
        if (k + 2 < K) v.z = a[(int64_t)(k + 2) * lda + row];
        if (k + 3 < K) v.w = a[(int64_t)(k + 3) * lda + row];
        return v;
    }
};
struct ConvA {
    const float* x;
    ConvGeom g;
    int K;
    __device__ __forceinline__ float4 load(int64_t row, int64_t M, int k) const {
        if (row >= M || k >= K) return make_float4(0.f, 0.f, 0.f, 0.f);
        const int tap = k / g.Cin, c = k - tap * g.Cin;
        const int r = tap / g.kw, s = tap - r * g.kw;
        const int hw = g.Ho * g.Wo;
        const int64_t p = row / hw;
        const int rem = (int)(row - p * hw);
        const int oy = rem / g.Wo, ox = rem - oy * g.Wo;
        const int iy = oy * g.stride - g.pad + r, ix = ox * g.stride - g.pad + s;
        if (iy < 0 || iy >= g.H || ix < 0 || ix >= g.W) return make_float4(0.f, 0.f, 0.f, 0.f);
        return *reinterpret_cast<const float4*>(x + ((p * g.H + iy) * g.W + ix) * (int64_t)g.Cin + c);
    }
};

// kWNK: weights stored (N,K) (nn.Linear) instead of (K,N) (conv, k-major rows)
template <class ALoader, bool kWNK>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(ALoader A, const float* __restrict__ w, const float* __restrict__ scale,
                 const float* __restrict__ shift, const float* __restrict__ res, float* __restrict__ y,
                 int64_t M, int N, int K, int relu) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int a_row = tid >> 2, a_kq = (tid & 3) * 4;
    for (int k0 = 0; k0 < K; k0 += BK) {
        const float4 av = A.load(m0 + a_row, M, k0 + a_kq);
        As[a_kq + 0][a_row] = av.x;
        As[a_kq + 1][a_row] = av.y;
        As[a_kq + 2][a_row] = av.z;
        As[a_kq + 3][a_row] = av.w;
        if (kWNK) {
            const int n = n0 + a_row, k = k0 + a_kq;
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < N && k < K) bv = *reinterpret_cast<const float4*>(w + (int64_t)n * K + k);
            Bs[a_kq + 0][a_row] = bv.x;
            Bs[a_kq + 1][a_row] = bv.y;
            Bs[a_kq + 2][a_row] = bv.z;
            Bs[a_kq + 3][a_row] = bv.w;
        } else {
            const int k = k0 + (tid >> 4), n = n0 + (tid & 15) * 4;
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < K && n < N) bv = *reinterpret_cast<const float4*>(w + (int64_t)k * N + n);
            *reinterpret_cast<float4*>(&Bs[tid >> 4][(tid & 15) * 4]) = bv;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (scale) v *= scale[n];
            if (shift) v += shift[n];
            if (res) v += res[m * N + n];
            if (relu) v = fmaxf(v, 0.f);
            y[m * N + n] = v;
        }
    }
}

// U[d, h*T+t] = sum_j k_w[h*Dk+j, d] * (q_w q_t)[h*Dk+j] / sqrt(Dk); one block per (h,t)
__global__ void score_basis_kernel(const float* __restrict__ q_tok, const float* __restrict__ q_w,
                                   const float* __restrict__ k_w, float* __restrict__ U,
                                   int D, int H, int Dk, int T) {
    extern __shared__ float qp[];   // Dk projected+scaled query entries of this head/token
    const int h = blockIdx.x / T, t = blockIdx.x % T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const float temp = sqrtf((float)Dk);
    for (int j = warp; j < Dk; j += nwarps) {
        const float* wr = q_w + (int64_t)(h * Dk + j) * D;
        const float* qr = q_tok + (int64_t)t * D;
        float s = 0.f;
        for (int d = lane; d < D; d += 32) s = fmaf(wr[d], qr[d], s);
        s = ipsb::warp_sum(s);
        if (lane == 0) qp[j] = s / temp;
    }
    __syncthreads();
    for (int d = tid; d < D; d += blockDim.x) {
        float s = 0.f;
        for (int j = 0; j < Dk; ++j) s = fmaf(k_w[(int64_t)(h * Dk + j) * D + d], qp[j], s);
        U[(int64_t)d * (H * T) + blockIdx.x] = s;
    }
}

// z[row, ht] = emb[row,:] . U[:,ht] (+ add).  One warp per row: lane l owns the float4 groups l, l+32, ... of the row
// (coalesced 512-byte reads), accumulates 8 columns at a time against U^T staged in shared memory and the warp reduces
// with shuffles.  The summation order of a row is fixed by D alone, so a row's logits do not depend on how rows are
// chunked or sharded.
__global__ void __launch_bounds__(256)
logits_kernel(const float* __restrict__ emb, const float* __restrict__ U, const float* __restrict__ add_tab,
              const int64_t* __restrict__ add_idx, float* __restrict__ z, int64_t rows, int D, int HT) {
    extern __shared__ __align__(16) float Ut[];   // [HT][D]
    for (int i = threadIdx.x; i < D * HT; i += blockDim.x) {
        const int d = i / HT, c = i - d * HT;
        Ut[c * D + d] = U[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int n4 = D >> 2;
    for (int64_t row = (int64_t)blockIdx.x * wpb + warp; row < rows; row += (int64_t)gridDim.x * wpb) {
        const float4* e = reinterpret_cast<const float4*>(emb + row * D);
        for (int c0 = 0; c0 < HT; c0 += 8) {
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
            for (int d4 = lane; d4 < n4; d4 += 32) {
                const float4 ev = e[d4];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (c0 + j < HT) {
                        const float4 uv = reinterpret_cast<const float4*>(Ut + (c0 + j) * D)[d4];
                        acc[j] = fmaf(ev.x, uv.x, acc[j]);
                        acc[j] = fmaf(ev.y, uv.y, acc[j]);
                        acc[j] = fmaf(ev.z, uv.z, acc[j]);
                        acc[j] = fmaf(ev.w, uv.w, acc[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
            }
            if (lane < 8 && c0 + lane < HT) {
                float s = acc[0];
#pragma unroll
                for (int j = 1; j < 8; ++j) s = (lane == j) ? acc[j] : s;
                const int c = c0 + lane;
                if (add_tab) s += add_tab[(add_idx ? add_idx[row] : row) * HT + c];
                z[row * HT + c] = s;
            }
        }
    }
}

}  // namespace

extern "C" {

int ipsb_conv_f32(const float* x, const float* w, const float* scale, const float* shift,
                  const float* res, float* y, int64_t P, int H, int W, int Cin, int Cout,
                  int kh, int kw, int stride, int pad, int relu, void* stream) {
    IPSB_REQUIRE(P > 0 && Cin % 4 == 0 && Cout % 4 == 0, "conv_f32: Cin=%d / Cout=%d must be multiples of 4", Cin, Cout);
    ConvGeom g{H, W, Cin, (H + 2 * pad - kh) / stride + 1, (W + 2 * pad - kw) / stride + 1, kh, kw, stride, pad};
    const int K = kh * kw * Cin;
    const int64_t M = P * g.Ho * g.Wo;
    ConvA A{x, g, K};
    dim3 grid((unsigned)ipsb::ceil_div(M, BM), (unsigned)ipsb::ceil_div(Cout, BN));
    gemm_simt_kernel<ConvA, false><<<grid, 256, 0, (cudaStream_t)stream>>>(A, w, scale, shift, res, y, M, Cout, K, relu);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_linear_f32(const float* a, const float* w, const float* scale, const float* shift,
                    float* y, int64_t M, int N, int K, int relu, void* stream) {
    IPSB_REQUIRE(M > 0 && N > 0 && K % 4 == 0, "linear_f32: K=%d must be a multiple of 4", K);
    DenseA A{a, K};
    dim3 grid((unsigned)ipsb::ceil_div(M, BM), (unsigned)ipsb::ceil_div(N, BN));
    gemm_simt_kernel<DenseA, true><<<grid, 256, 0, (cudaStream_t)stream>>>(A, w, scale, shift, nullptr, y, M, N, K, relu);
    IPSB_LAUNCH_CHECK();
    return 0;
}

// fp32 CUDA-core GEMM in the three layouts of the train step (any M, N, K with K % 4 == 0 for modes 0/1, N % 4 == 0 for
// modes 1/2): mode 0 (NT) a (M,K) b (N,K); mode 1 (NN) a (M,K) b (K,N); mode 2 (TN) a (K,M) b (K,N).  c = a-op b + shift.
int ipsb_gemm_f32(int mode, const float* a, const float* b, const float* shift, float* c, int64_t M, int N, int64_t K,
                  int relu, void* stream) {
    IPSB_REQUIRE(mode >= 0 && mode <= 2 && M > 0 && N > 0 && K > 0 && K < (1ll << 31), "gemm_f32: bad arguments");
    dim3 grid((unsigned)ipsb::ceil_div(M, BM), (unsigned)ipsb::ceil_div(N, BN));
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0) {
        IPSB_REQUIRE(K % 4 == 0, "gemm_f32 NT: K=%lld must be a multiple of 4", (long long)K);
        DenseA A{a, (int)K};
        gemm_simt_kernel<DenseA, true><<<grid, 256, 0, st>>>(A, b, nullptr, shift, nullptr, c, M, N, (int)K, relu);
    } else if (mode == 1) {
        IPSB_REQUIRE(K % 4 == 0 && N % 4 == 0, "gemm_f32 NN: K=%lld and N=%d must be multiples of 4", (long long)K, N);
        DenseA A{a, (int)K};
        gemm_simt_kernel<DenseA, false><<<grid, 256, 0, st>>>(A, b, nullptr, shift, nullptr, c, M, N, (int)K, relu);
    } else {
        IPSB_REQUIRE(N % 4 == 0, "gemm_f32 TN: N=%d must be a multiple of 4", N);
        DenseAT A{a, (int)K, M};
        gemm_simt_kernel<DenseAT, false><<<grid, 256, 0, st>>>(A, b, nullptr, shift, nullptr, c, M, N, (int)K, relu);
    }
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_score_basis(const float* q_tok, const float* q_w, const float* k_w, float* U,
                     int D, int H, int Dk, int T, void* stream) {
    IPSB_REQUIRE(D > 0 && H > 0 && Dk > 0 && T > 0, "score_basis: bad shape");
    score_basis_kernel<<<H * T, 256, Dk * sizeof(float), (cudaStream_t)stream>>>(q_tok, q_w, k_w, U, D, H, Dk, T);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_logits(const float* emb, const float* U, const float* add_tab, const int64_t* add_idx,
                float* z, int64_t rows, int D, int HT, void* stream) {
    IPSB_REQUIRE(rows > 0 && D % 4 == 0 && HT > 0 && HT <= 32, "logits: bad shape rows=%lld D=%d HT=%d", (long long)rows, D, HT);
    const size_t smem = (size_t)HT * D * sizeof(float);
    IPSB_REQUIRE(smem <= 200 * 1024, "logits: D*HT too large for shared memory");
    IPSB_CUDA(cudaFuncSetAttribute(logits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t g = ipsb::ceil_div(rows, 8);              // 8 warps = 8 rows per block pass
    const int64_t cap = (int64_t)ipsb::sm_count() * 4;
    if (g > cap) g = cap;
    logits_kernel<<<(unsigned)g, 256, smem, (cudaStream_t)stream>>>(emb, U, add_tab, add_idx, z, rows, D, HT);
    IPSB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
