// Feature projector + score projection in ONE kernel: features -> logits, nothing else reaches HBM.
//
// Reference (architecture/ips_net.py:54-60 and architecture/transformer.py:71-83, eval mode inside `ips`):
//     emb = ReLU(BatchNorm1d(Linear(LayerNorm(x))))            x: (rows, K) fp32 features, emb: (rows, N)
//     z   = emb @ U                                            U = k_w^T (q_w q) / sqrt(D_k), (N, H*T)      
// Round 1 ran four passes over HBM for this (LayerNorm + cast, GEMM, logits; 205 MB bf16 copy + 102 MB embeddings written
// and re-read per 50 k-patch slide).  Here a CTA owns 128 rows at a time:
//
//   8 worker warps    K loop: stream the fp32 (or bf16) rows from HBM with 256-bit loads -- a ring of eight register buffers
//                     per thread keeps seven loads (57 KB per CTA) in flight --, round to bf16 straight into the
//                     128B-swizzled K-major A stage in shared memory and accumulate each row's sum / sum of squares of the
//                     ROUNDED values: LayerNorm is applied algebraically in the epilogue,
//                         LN(x) W^T = rstd * (x W^T - mean * colsum(W))
//                     so the tensor cores consume the raw features and no normalised copy exists anywhere;
//                     tile end: the same warps drain the accumulator: tcgen05.ld -> rstd * (a_n * acc - mean * b_n) + shift_n
//                     -> ReLU -> 8 running dot products with U in registers (fp32) -> 32 bytes of logits per row to HBM
//                     (their loads for the next tile are already in flight);
//   TMA warp          streams W (N, K) bf16 in (256 x 64) boxes;
//   MMA warp          tcgen05.mma 128 x 256 x 16, two accumulators side by side: all 512 TMEM columns hold the 128 x N tile.
//
// (Measured and dropped: L2 prefetches ahead of the register loads -- per k block or row-contiguous a quarter tile ahead --
// made the kernel 20-35 % slower.)
// HBM traffic = the features once (8 192 B / patch fp32) + 32 B / patch of logits; W (2 MB) is re-streamed from L2 per tile.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "umma.cuh"
#include "../../include/ips_b200.h"

namespace {

using bf16 = __nv_bfloat16;
constexpr int TILE_M = 128;
constexpr int BK = 64;
constexpr int SA = 3;                         // A stages (128 rows x 64 bf16 = 16 KB)
constexpr int SW = 4;                         // W stages (256 rows x 64 bf16 = 32 KB)
constexpr int A_BYTES = TILE_M * 128;
constexpr int W_BYTES = 256 * 128;
constexpr int HTP = 8;                        // logit columns kept per row (H*T <= 8)
constexpr int N_WORK_WARPS = 8;                // converters during the K loop, epilogue at the end of a tile
                                              // (10 warps: the register file leaves ~168 registers per thread for the load pipeline)
constexpr int N_WORKERS = 32 * N_WORK_WARPS;
constexpr int THREADS = 32 * (2 + N_WORK_WARPS);

struct ProjParams {
    const void* x;            // (rows, K) fp32 or bf16
    const float* table;       // (N, 12): a_n, b_n, shift_n, 0, U[n][0..7]
    float* z;                 // (rows, HT)
    int64_t rows;
    int K, N, HT, KS, halves; // KS = K / 64, halves = N / 256
    int tiles;
    float eps;
    // scan-order form: tile row g of the flat (batches x rows_per_batch) sequence reads feature row
    // (g / rows_per_batch) * rows_per_batch + perm[(g / rows_per_batch) * perm_stride + g % rows_per_batch]; z stays in tile order
    const int64_t* perm;
    int64_t perm_stride, rows_per_batch;
    int* tile_flags;          // [tiles] set to 1 (release) once a tile's logits are in z, or null
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void ldg256(const float* p, float (&v)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&p);
}

template <bool IN_BF16>
__global__ void __launch_bounds__(THREADS, 1)
projector_logits_kernel(const __grid_constant__ CUtensorMap tmW, const ProjParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw0 = umma::smem_u32(smem_raw);
    const uint32_t smem0 = (raw0 + 1023u) & ~1023u;
    const uint32_t a0 = smem0;                                   // SA x 16 KB
    const uint32_t w0 = a0 + SA * A_BYTES;                       // SW x 32 KB
    const uint32_t tab0 = w0 + SW * W_BYTES;                     // N x 48 B
    const uint32_t stat0 = tab0 + (uint32_t)p.N * 48u;           // 128 x (mean, rstd)
    const uint32_t zb0 = stat0 + TILE_M * 8u;                    // 2 (tile parity) x 128 x 8 floats
    const uint32_t bar0 = zb0 + 2u * TILE_M * HTP * 4u;
    auto w_full = [&](int s) { return bar0 + 8u * s; };
    auto w_empty = [&](int s) { return bar0 + 8u * (SW + s); };
    auto a_full = [&](int s) { return bar0 + 8u * (2 * SW + s); };
    auto a_empty = [&](int s) { return bar0 + 8u * (2 * SW + SA + s); };
    const uint32_t t_full = bar0 + 8u * (2 * SW + 2 * SA);
    const uint32_t t_empty = t_full + 8u;
    const uint32_t tmem_slot = t_full + 16u;
    auto gptr = [&](uint32_t saddr) { return smem_raw + (saddr - raw0); };
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gptr(tmem_slot));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < SW; ++s) { umma::mbar_init(w_full(s), 1); umma::mbar_init(w_empty(s), 1); }
        for (int s = 0; s < SA; ++s) { umma::mbar_init(a_full(s), N_WORK_WARPS); umma::mbar_init(a_empty(s), 1); }
        umma::mbar_init(t_full, 1);
        umma::mbar_init(t_empty, N_WORKERS);
        umma::fence_barrier_init();
    }
    if (warp == 1) umma::tmem_alloc(tmem_slot, 512);
    // column table -> shared memory (read by the epilogue as broadcast 128-bit loads)
    {
        float4* dst = reinterpret_cast<float4*>(gptr(tab0));
        const float4* src = reinterpret_cast<const float4*>(p.table);
        for (int i = tid; i < p.N * 3; i += THREADS) dst[i] = __ldg(src + i);
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int n_my = (blockIdx.x < (unsigned)p.tiles) ? (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const uint32_t total = (uint32_t)n_my * (uint32_t)p.KS;      // (tile, k block) pairs of this CTA

    if (warp == 0) {
        // ---------------- W producer
        if (lane == 0) {
            uint32_t wi = 0;
            for (uint32_t it = 0; it < total; ++it) {
                const int kb = (int)(it % (uint32_t)p.KS);
                for (int h = 0; h < p.halves; ++h, ++wi) {
                    const int st = wi % SW;
                    umma::mbar_wait(w_empty(st), ((wi / SW) & 1) ^ 1);
                    umma::mbar_expect_tx(w_full(st), W_BYTES);
                    tma_load_2d(w0 + st * W_BYTES, &tmW, w_full(st), kb * BK, h * 256);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---------------- MMA issuer: the whole warp runs the loop, the elected lane issues
        const uint32_t leader = umma::elect_one();
        constexpr uint32_t idesc = umma::idesc_bf16_f32(TILE_M, 256);
        uint32_t wi = 0;
        for (uint32_t it = 0; it < total; ++it) {
            const int kb = (int)(it % (uint32_t)p.KS);
            const uint32_t tcount = it / (uint32_t)p.KS;
            if (kb == 0) {                                      // the epilogue has drained the previous tile
                umma::mbar_wait(t_empty, (tcount & 1) ^ 1);
                umma::tc_fence_after();
            }
            const int sa = it % SA;
            umma::mbar_wait(a_full(sa), (it / SA) & 1);
            umma::tc_fence_after();
            const uint32_t a_addr = a0 + sa * A_BYTES;
            for (int h = 0; h < p.halves; ++h, ++wi) {
                const int st = wi % SW;
                umma::mbar_wait(w_full(st), (wi / SW) & 1);
                umma::tc_fence_after();
                const uint32_t w_addr = w0 + st * W_BYTES;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                    umma::mma_bf16_w(tmem_base + (uint32_t)h * 256u, umma::smem_desc_sw128(a_addr) + 2u * k,
                                     umma::smem_desc_sw128(w_addr) + 2u * k, idesc, (kb > 0) || (k != 0), leader);
                umma::mma_commit_w(w_empty(st), leader);
            }
            umma::mma_commit_w(a_empty(sa), leader);
            if (kb == p.KS - 1) umma::mma_commit_w(t_full, leader);
        }
        __syncwarp();
    } else {
        // ---------------- workers: K loop = HBM -> registers -> bf16 -> swizzled A stage (+ row statistics of the rounded
        //                  values); tile end = epilogue on the accumulator
        const int cw = warp - 2;                                 // 0..7: rows [16 cw, 16 cw + 16) of the tile in the K loop
        const int sub = lane >> 3, chunk = lane & 7;             // 8 lanes cover one row's 64-element k block
        const int q = warp & 3;                                  // TMEM lane quarter this warp may read
        const int grp = cw >> 2;                                 // epilogue: column half [grp * N/2, +N/2)
        const int erow = q * 32 + lane;                          // epilogue: accumulator row of this thread
        const int cols_per_grp = p.N / 2;
        const int col0 = grp * cols_per_grp;
        const float4* tab = reinterpret_cast<const float4*>(gptr(tab0));
        float2* stats = reinterpret_cast<float2*>(gptr(stat0));
        const uint32_t total_q = total * 4u;                     // one "quarter" = one row of this thread's four per k block
        // The load stream runs seven quarters ahead of the conversion.  Its position is kept incrementally (no integer
        // divisions in the loop): pointer to this thread's chunk of row (16 cw + sub) of the stream's tile at its k block,
        // the k block index inside the tile, and which of the thread's four rows exist (ragged last tile).
        const int esz = IN_BF16 ? 2 : 4;
        const int64_t row_bytes = (int64_t)p.K * esz;
        const char* const ld_base = reinterpret_cast<const char*>(p.x) + chunk * 8 * esz;
        int ld_t = 0, ld_kb = 0;
        uint32_t ld_issued = 0;
        const char* ld_ptr = ld_base;                            // ld_base + k block offset
        int ld_row[4];                                           // feature rows of this thread's four tile rows (-1: outside)
        auto ld_enter_tile = [&]() {
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)ld_t * gridDim.x;
            const int64_t g0 = tile * TILE_M + cw * 16 + sub;    // tile rows g0, g0 + 4, g0 + 8, g0 + 12
            ld_ptr = ld_base;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t g = g0 + 4 * j;
                int r = -1;
                if (g < p.rows) {
                    if (p.perm) {
                        const int64_t bb = g / p.rows_per_batch, s = g - bb * p.rows_per_batch;
                        r = (int)(bb * p.rows_per_batch + __ldg(p.perm + bb * p.perm_stride + s));
                    } else {
                        r = (int)g;
                    }
                }
                ld_row[j] = r;
            }
        };
        ld_enter_tile();
        // quarter J of the stream's current k block: one 256-bit (fp32) / 128-bit (bf16) load
        auto load = [&](const int J, float (&buf)[8]) {
            if (ld_issued < total_q) {
                if (ld_row[J] >= 0) {
                    const char* src = ld_ptr + (int64_t)ld_row[J] * row_bytes;
                    if (IN_BF16) {
                        const int4 qv = ipsb::ld_stream16(src);
                        const uint32_t u[4] = {(uint32_t)qv.x, (uint32_t)qv.y, (uint32_t)qv.z, (uint32_t)qv.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            buf[2 * j] = __uint_as_float(u[j] << 16);
                            buf[2 * j + 1] = __uint_as_float(u[j] & 0xffff0000u);
                        }
                    } else {
                        ldg256(reinterpret_cast<const float*>(src), buf);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) buf[j] = 0.f;
                }
                ++ld_issued;
                if (J == 3) {                                    // next k block of the stream
                    ld_ptr += BK * esz;
                    if (++ld_kb == p.KS) { ld_kb = 0; ++ld_t; ld_enter_tile(); }
                }
            }
        };
        float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
        // convert quarter qi (row J of the thread, known at compile time) while seven later quarters stay in flight
        uint32_t cv_sa = 0, cv_ph = 1;                           // A stage of the k block being converted, parity to wait for
        auto conv = [&](const int J, float (&cur)[8], float (&fut)[8]) {
            load((J + 3) & 3, fut);                              // quarter (qi + 7): row (J + 7) % 4
            const uint32_t sa = cv_sa;
            if (J == 0) umma::mbar_wait(a_empty(sa), cv_ph);
            const int rl = cw * 16 + J * 4 + sub;
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                w[j] = pack_bf16x2(cur[2 * j], cur[2 * j + 1]);
                const float lo = __uint_as_float(w[j] << 16), hi = __uint_as_float(w[j] & 0xffff0000u);
                s1[J] += lo + hi;
                s2[J] = fmaf(lo, lo, fmaf(hi, hi, s2[J]));
            }
            const uint32_t dst = a0 + sa * A_BYTES + (uint32_t)rl * 128u + (((uint32_t)chunk ^ (uint32_t)(rl & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
            if (J == 3) {                                        // the warp's 16 rows of this k block are in the stage
                umma::fence_proxy_async();                       // generic-proxy stores -> visible to tcgen05.mma
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(a_full(sa));
                if (++cv_sa == SA) { cv_sa = 0; cv_ph ^= 1u; }
            }
        };
        // ================= end of a tile: row statistics, then the epilogue =================
        auto epilogue = [&](uint32_t tcount) {
            const int64_t tile = (int64_t)blockIdx.x + (int64_t)tcount * gridDim.x;
            const int par = (int)(tcount & 1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float a = s1[i], c = s2[i];
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) {
                    a += __shfl_xor_sync(0xffffffffu, a, o);
                    c += __shfl_xor_sync(0xffffffffu, c, o);
                }
                if (chunk == 0) {
                    const float mean = a / (float)p.K;
                    const float var = fmaxf(c / (float)p.K - mean * mean, 0.f);
                    stats[cw * 16 + i * 4 + sub] = make_float2(mean, rsqrtf(var + p.eps));
                }
                s1[i] = 0.f; s2[i] = 0.f;
            }
            umma::named_bar_sync(1, N_WORKERS);                  // statistics of all 128 rows are in shared memory
            const float2 st = stats[erow];
            const float k_mu = -st.x * st.y, rstd = st.y;        // e = relu(acc * (rstd a_n) + (shift_n - rstd mean b_n))
            umma::mbar_wait(t_full, tcount & 1);
            umma::tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col0;
            float zacc[HTP];
#pragma unroll
            for (int j = 0; j < HTP; ++j) zacc[j] = 0.f;
#pragma unroll 1
            for (int c0 = 0; c0 < cols_per_grp; c0 += 32) {
                uint32_t v[32];
                umma::tmem_ld32(t_row + (uint32_t)c0, v);
                umma::tmem_ld_wait();
                if (c0 + 32 >= cols_per_grp) {                   // accumulator fully read by this thread
                    umma::tc_fence_before();
                    umma::mbar_arrive(t_empty);
                }
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int n = col0 + c0 + c;
                    const float4 pr = tab[3 * n], u0 = tab[3 * n + 1], u1 = tab[3 * n + 2];
                    const float e = fmaxf(fmaf(__uint_as_float(v[c]), rstd * pr.x, fmaf(k_mu, pr.y, pr.z)), 0.f);
                    zacc[0] = fmaf(e, u0.x, zacc[0]); zacc[1] = fmaf(e, u0.y, zacc[1]);
                    zacc[2] = fmaf(e, u0.z, zacc[2]); zacc[3] = fmaf(e, u0.w, zacc[3]);
                    zacc[4] = fmaf(e, u1.x, zacc[4]); zacc[5] = fmaf(e, u1.y, zacc[5]);
                    zacc[6] = fmaf(e, u1.z, zacc[6]); zacc[7] = fmaf(e, u1.w, zacc[7]);
                }
            }
            // combine the two column halves in a fixed order (deterministic) and store the row's logits
            float* zb = reinterpret_cast<float*>(gptr(zb0)) + ((size_t)par * TILE_M + erow) * HTP;
            if (grp == 1) {
                reinterpret_cast<float4*>(zb)[0] = make_float4(zacc[0], zacc[1], zacc[2], zacc[3]);
                reinterpret_cast<float4*>(zb)[1] = make_float4(zacc[4], zacc[5], zacc[6], zacc[7]);
            }
            umma::named_bar_sync(1, N_WORKERS);
            if (grp == 0) {
                const float4 o0 = reinterpret_cast<const float4*>(zb)[0], o1 = reinterpret_cast<const float4*>(zb)[1];
                const float r[HTP] = {zacc[0] + o0.x, zacc[1] + o0.y, zacc[2] + o0.z, zacc[3] + o0.w,
                                      zacc[4] + o1.x, zacc[5] + o1.y, zacc[6] + o1.z, zacc[7] + o1.w};
                const int64_t grow = tile * TILE_M + erow;
                if (grow < p.rows) {
                    float* dst = p.z + grow * p.HT;
                    if (p.HT == HTP) {
                        reinterpret_cast<float4*>(dst)[0] = make_float4(r[0], r[1], r[2], r[3]);
                        reinterpret_cast<float4*>(dst)[1] = make_float4(r[4], r[5], r[6], r[7]);
                    } else {
#pragma unroll
                        for (int j = 0; j < HTP; ++j)
                            if (j < p.HT) dst[j] = r[j];
                    }
                }
                if (p.tile_flags) {                              // publish the tile to the selection loop running beside us
                    __threadfence();
                    umma::named_bar_sync(2, N_WORKERS / 2);
                    if (erow == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.tile_flags + tile), "r"(1) : "memory");
                }
            }
        };
        float b0[8], b1[8], b2[8], b3[8], b4[8], b5[8], b6[8], b7[8];
        load(0, b0); load(1, b1); load(2, b2); load(3, b3); load(0, b4); load(1, b5); load(2, b6);
        for (int t = 0; t < n_my; ++t) {                         // KS is even: the ring of eight buffers lines up with the tiles
            for (int kb = 0; kb < p.KS; kb += 2) {
                conv(0, b0, b7);
                conv(1, b1, b0);
                conv(2, b2, b1);
                conv(3, b3, b2);
                conv(0, b4, b3);
                conv(1, b5, b4);
                conv(2, b6, b5);
                conv(3, b7, b6);
            }
            epilogue((uint32_t)t);
        }
    }

    umma::tc_fence_before();
    __syncthreads();
    if (warp == 1) umma::tmem_dealloc(tmem_base, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

}  // namespace

extern "C" {

int ipsb_projector_preload(void) {
    cudaFuncAttributes fa;
    IPSB_CUDA(cudaFuncGetAttributes(&fa, projector_logits_kernel<true>));
    IPSB_CUDA(cudaFuncGetAttributes(&fa, projector_logits_kernel<false>));
    return 0;
}

int ipsb_projector_logits(const void* x, int x_is_bf16, const void* w_bf16, const float* table, float* z,
                          int64_t rows, int K, int N, int HT, float eps, void* stream) {
    return ipsb_projector_logits_scan(x, x_is_bf16, w_bf16, table, z, rows, K, N, HT, eps, nullptr, 0, rows, nullptr, 0, stream);
}

int ipsb_projector_logits_scan(const void* x, int x_is_bf16, const void* w_bf16, const float* table, float* z,
                               int64_t rows, int K, int N, int HT, float eps,
                               const int64_t* perm, int64_t perm_batch_stride, int64_t rows_per_batch,
                               int* tile_flags, int max_ctas, void* stream) {
    IPSB_REQUIRE(x && w_bf16 && table && z && rows > 0, "projector_logits: null argument");
    IPSB_REQUIRE(rows < (1ll << 31), "projector_logits: %lld rows", (long long)rows);
    IPSB_REQUIRE(perm == nullptr || (rows_per_batch > 0 && rows % rows_per_batch == 0), "projector_logits: rows_per_batch");
    IPSB_REQUIRE(K % (2 * BK) == 0 && K >= 4 * BK, "projector_logits: K=%d must be a multiple of %d, at least %d", K, 2 * BK, 4 * BK);
    IPSB_REQUIRE(N == 256 || N == 512, "projector_logits: N=%d (256 or 512 supported: the 128 x N tile lives in TMEM)", N);
    IPSB_REQUIRE(HT >= 1 && HT <= HTP, "projector_logits: H*T=%d exceeds %d", HT, HTP);
    IPSB_REQUIRE(((uintptr_t)x % 32) == 0 && ((uintptr_t)w_bf16 % 16) == 0 && ((uintptr_t)table % 16) == 0 && ((uintptr_t)z % 16) == 0,
                 "projector_logits: misaligned pointer");
    EncodeTiledFn enc = encode_fn();
    IPSB_REQUIRE(enc != nullptr, "projector_logits: cuTensorMapEncodeTiled not available from the driver");
    alignas(64) CUtensorMap tmW;
    {
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
        cuuint64_t strides[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {BK, 256};
        cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_bf16), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        IPSB_REQUIRE(r == CUDA_SUCCESS, "projector_logits: cuTensorMapEncodeTiled(W) failed with %d", (int)r);
    }
    ProjParams p;
    p.x = x; p.table = table; p.z = z; p.rows = rows; p.K = K; p.N = N; p.HT = HT; p.KS = K / BK; p.halves = N / 256;
    p.tiles = (int)((rows + TILE_M - 1) / TILE_M);
    p.eps = eps;
    p.perm = perm; p.perm_stride = perm_batch_stride; p.rows_per_batch = rows_per_batch > 0 ? rows_per_batch : rows;
    p.tile_flags = tile_flags;
    const size_t smem = 1024 + (size_t)SA * A_BYTES + (size_t)SW * W_BYTES + (size_t)N * 48 + TILE_M * 8 + 2 * TILE_M * HTP * 4 +
                        8 * (2 * SW + 2 * SA + 2) + 64;
    IPSB_REQUIRE(smem <= 227 * 1024, "projector_logits: %zu bytes of shared memory", smem);
    int grid = ipsb::sm_count();
    if (max_ctas > 0 && max_ctas < grid) grid = max_ctas;       // the selection loop's cluster keeps its SMs
    if (p.tiles < grid) grid = p.tiles;
    if (x_is_bf16) {
        auto kern = projector_logits_kernel<true>;
        IPSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, THREADS, smem, (cudaStream_t)stream>>>(tmW, p);
    } else {
        auto kern = projector_logits_kernel<false>;
        IPSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, THREADS, smem, (cudaStream_t)stream>>>(tmW, p);
    }
    IPSB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
