// General bf16 tensor-core GEMM for the train step: C = op(A) * op(B) with fp32 accumulation in TMEM.
//
//   NT   A (M,K) row-major, B (N,K) row-major      y = x W^T            forward of nn.Linear
//   NN   A (M,K) row-major, B (K,N) row-major      dx = dy W            data gradient
//   TN   A (K,M) row-major, B (K,N) row-major      dW = dy^T x          weight gradient (contraction over rows)
//
// A "row-major with the contraction index fastest" operand is K-major for tcgen05; one with the
// contraction index slowest is MN-major (instruction-descriptor bits 15/16).  Both arrive as TMA
// boxes with the 128B swizzle: K-major tiles as (rows, 64 k) boxes, MN-major tiles as (64 k rows, 64
// elements) boxes -- one box per 64 output rows/columns, LBO = 8 KB between them, SBO = 1 KB
// (`tools/probe_mnmajor.cu` pins this layout).  Out-of-range rows/columns/contraction indices are
// zero-filled on load and clipped on store, so M, N and K need no padding.
//
// Persistent, warp-specialised like the convolution kernels; TN problems with a long contraction and a
// small output are split along K into fp32 partial outputs that a second kernel reduces (deterministic).
//
// Replaces the cuBLAS calls behind nn.Linear forward/backward (architecture/ips_net.py:57,
// architecture/transformer.py:60-63,116-117) in the grad-mode step.
#include <cuda.h>
#include "common.cuh"
#include "umma.cuh"
#include "epilogue.cuh"
#include "../../include/ips_b200.h"

namespace {

using bf16 = __nv_bfloat16;
constexpr int TILE_M = 128;
constexpr int BK = 64;
constexpr int A_BYTES = TILE_M * 128;

struct GemmParams {
    const float* scale;   // per output column, may be null
    const float* shift;   // per output column (bias), may be null
    const bf16* res;      // unused
    int M, N, K, relu;
    int tiles_m, tiles_n, splits, ks_per_split, KS;   // KS = ceil(K / 64)
    int total_tiles;
    int64_t split_stride_rows;   // rows between split partial outputs in the output tensor map (TN split-K)
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

// smem descriptor of an MN-major SWIZZLE_128B operand: LBO = 8 KB between 64-element blocks, SBO = 1 KB
__device__ __forceinline__ uint64_t desc_mn128(uint32_t addr) {
    const uint32_t lo = ((addr >> 4) & 0x3fffu) | ((8192u >> 4) << 16);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}

template <int BN, int STAGES, bool A_MN, bool B_MN, typename OutT>
__global__ void __launch_bounds__(320, 1)
gemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
    constexpr int B_BYTES = BN * 128;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem0 = (umma::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t out_stage0 = smem0 + STAGES * STAGE_BYTES;
    const uint32_t bar0 = out_stage0 + 2u * epi::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bar0 + 8u * (2 * STAGES + 4);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - umma::smem_u32(smem_raw)));
    const uint32_t sc_addr = (tmem_slot + 4u + 15u) & ~15u;
    float* sc_smem = reinterpret_cast<float*>(smem_raw + (sc_addr - umma::smem_u32(smem_raw)));   // scale[BN*tiles_n], shift[...]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Npad = p.tiles_n * BN;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { umma::mbar_init(full_bar(s), 1); umma::mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { umma::mbar_init(tfull_bar(a), 1); umma::mbar_init(tempty_bar(a), 128); }
        umma::fence_barrier_init();
    }
    if (warp == 1) umma::tmem_alloc(tmem_slot, 2 * BN);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    // tile -> (m tile, n tile, split); n fastest so neighbouring CTAs share the A tile in L2
    auto decode = [&](int tile, int& mt, int& nt, int& sp) {
        nt = tile % p.tiles_n; tile /= p.tiles_n;
        mt = tile % p.tiles_m; sp = tile / p.tiles_m;
    };

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                int mt, nt, sp;
                decode(tile, mt, nt, sp);
                const int ks0 = sp * p.ks_per_split, ks1 = min(p.KS, ks0 + p.ks_per_split);
                for (int ks = ks0; ks < ks1; ++ks, ++it) {
                    const int stage = it % STAGES;
                    umma::mbar_wait(empty_bar(stage), ((it / STAGES) & 1) ^ 1);
                    const uint32_t a_dst = smem0 + stage * STAGE_BYTES;
                    umma::mbar_expect_tx(full_bar(stage), A_BYTES + B_BYTES);
                    if (A_MN) {     // (64 k rows, 64 m) boxes, two per tile
                        tma_load_2d(a_dst, &tmA, full_bar(stage), mt * TILE_M, ks * BK);
                        tma_load_2d(a_dst + 8192, &tmA, full_bar(stage), mt * TILE_M + 64, ks * BK);
                    } else {        // (128 m rows, 64 k) box
                        tma_load_2d(a_dst, &tmA, full_bar(stage), ks * BK, mt * TILE_M);
                    }
                    if (B_MN) {
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j)
                            tma_load_2d(a_dst + A_BYTES + j * 8192, &tmB, full_bar(stage), nt * BN + j * 64, ks * BK);
                    } else {
                        tma_load_2d(a_dst + A_BYTES, &tmB, full_bar(stage), ks * BK, nt * BN);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        {                                            // the whole warp runs the loop; the elected lane issues
            const uint32_t leader = umma::elect_one();
            constexpr uint32_t idesc = umma::idesc_bf16_f32(TILE_M, BN) | (A_MN ? (1u << 15) : 0u) | (B_MN ? (1u << 16) : 0u);
            uint32_t it = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
                int mt, nt, sp;
                decode(tile, mt, nt, sp);
                const int ks0 = sp * p.ks_per_split, ks1 = min(p.KS, ks0 + p.ks_per_split);
                const uint32_t acc = tcount & 1;
                umma::mbar_wait(tempty_bar(acc), ((tcount >> 1) & 1) ^ 1);
                umma::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int ks = ks0; ks < ks1; ++ks, ++it) {
                    const int stage = it % STAGES;
                    umma::mbar_wait(full_bar(stage), (it / STAGES) & 1);
                    umma::tc_fence_after();
                    const uint32_t a_addr = smem0 + stage * STAGE_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // K-major: 32 bytes further along the row; MN-major: 16 k rows = 2 KB further down
                        const uint64_t adesc = A_MN ? desc_mn128(a_addr + k * 2048) : umma::smem_desc_sw128(a_addr) + 2u * k;
                        const uint64_t bdesc = B_MN ? desc_mn128(a_addr + A_BYTES + k * 2048) : umma::smem_desc_sw128(a_addr + A_BYTES) + 2u * k;
                        umma::mma_bf16_w(d_tmem, adesc, bdesc, idesc, (ks > ks0) || (k != 0), leader);
                    }
                    umma::mma_commit_w(empty_bar(stage), leader);
                }
                umma::mma_commit_w(tfull_bar(acc), leader);
            }
        }
        __syncwarp();
    } else {
        const int wg = (warp - 2) >> 2;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        for (int i = tid - 64; i < Npad; i += 256) {
            sc_smem[i] = (p.scale && i < p.N) ? p.scale[i] : 1.f;
            sc_smem[Npad + i] = (p.shift && i < p.N) ? p.shift[i] : 0.f;
        }
        umma::named_bar_sync(1, 256);
        uint32_t tcount = wg;
        const uint32_t stage = out_stage0 + (uint32_t)wg * epi::STAGE_BYTES;
        const bool issuer = (row == 0);
        for (int tile = blockIdx.x + wg * gridDim.x; tile < p.total_tiles; tile += 2 * gridDim.x, tcount += 2) {
            int mt, nt, sp;
            decode(tile, mt, nt, sp);
            umma::mbar_wait(tfull_bar(wg), (tcount >> 1) & 1);
            umma::tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + wg * BN;
            const int out_row0 = (int)(sp * p.split_stride_rows) + mt * TILE_M;
            uint32_t no_phase = 0;
            epi::drain_tile<BN, OutT>(t_row, tempty_bar(wg), sc_smem + nt * BN, sc_smem + Npad + nt * BN, true, false, 0u, 0u, no_phase,
                                      p.relu, stage, row, 2u + (uint32_t)wg, issuer,
                                      [&](int s0, uint32_t src) { epi::tma_store_2d(&tmC, src, nt * BN + s0, out_row0); },
                                      [](int, uint32_t, uint32_t) {});
        }
        if (issuer) epi::bulk_wait0();
    }

    umma::tc_fence_before();
    __syncthreads();
    if (warp == 1) umma::tmem_dealloc(tmem_base, 2 * BN);
}

// out[r, c] = sum_s partial[s, r, c] (+ shift[c]); fp32
__global__ void reduce_partials_kernel(const float* __restrict__ part, float* __restrict__ out, int64_t elems, int splits,
                                       const float* __restrict__ shift, int N) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < elems; i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < splits; ++k) s += part[(int64_t)k * elems + i];
        if (shift) s += shift[i % N];
        out[i] = s;
    }
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

// 2-D row-major tensor (rows, cols) of element size es; box (box_cols, box_rows)
int encode_2d(CUtensorMap* tm, const void* ptr, int64_t rows, int64_t cols, int es, int box_cols, int box_rows, const char* what) {
    EncodeTiledFn enc = encode_fn();
    IPSB_REQUIRE(enc != nullptr, "gemm: cuTensorMapEncodeTiled not available from the driver");
    IPSB_REQUIRE((cols * es) % 16 == 0 && ((uintptr_t)ptr % 16) == 0, "gemm: %s needs 16-byte aligned rows (cols=%lld)", what, (long long)cols);
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * es};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, es == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IPSB_REQUIRE(r == CUDA_SUCCESS, "gemm: cuTensorMapEncodeTiled(%s) failed with %d", what, (int)r);
    return 0;
}

template <int BN, int STAGES, bool A_MN, bool B_MN, typename OutT>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmParams& p, cudaStream_t st) {
    const size_t smem = (size_t)STAGES * (A_BYTES + BN * 128) + 2 * epi::STAGE_BYTES + 1024 + 8 * (2 * STAGES + 5) + 32 +
                        8 * (size_t)p.tiles_n * BN;
    IPSB_REQUIRE(smem <= 227 * 1024, "gemm: %zu bytes of shared memory", smem);
    auto kern = gemm_tma_kernel<BN, STAGES, A_MN, B_MN, OutT>;
    static size_t configured = 0;
    if (configured < smem) {
        IPSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int grid = ipsb::balanced_grid(p.total_tiles, ipsb::sm_count());
    kern<<<grid, 320, smem, st>>>(tmA, tmB, tmC, p);
    IPSB_LAUNCH_CHECK();
    return 0;
}

template <bool A_MN, bool B_MN, typename OutT>
int dispatch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmParams& p, int BN, cudaStream_t st) {
    if (BN == 128) return launch<128, 5, A_MN, B_MN, OutT>(tmA, tmB, tmC, p, st);
    return launch<64, 6, A_MN, B_MN, OutT>(tmA, tmB, tmC, p, st);
}

}  // namespace

extern "C" {

int64_t ipsb_gemm_workspace_bytes(int mode, int64_t M, int N, int64_t K) {
    if (mode != 2) return 0;
    const int BN = (N % 128 == 0) ? 128 : 64;
    const int64_t tiles = ((M + 127) / 128) * ((N + BN - 1) / BN);
    const int64_t KS = (K + 63) / 64;
    int64_t splits = (2 * ipsb::sm_count() + tiles - 1) / tiles;
    if (splits > KS) splits = KS;
    if (splits < 1) splits = 1;
    return splits <= 1 ? 0 : splits * ((M + 127) / 128 * 128) * (int64_t)N * 4;
}

// mode 0 (NT): a (M,K), b (N,K);  mode 1 (NN): a (M,K), b (K,N);  mode 2 (TN): a (K,M), b (K,N).  a, b bf16; c (M,N) fp32 or bf16.
int ipsb_gemm_bf16(int mode, const void* a, const void* b, const float* scale, const float* shift, void* c, int c_is_f32,
                   int64_t M, int N, int64_t K, int relu, void* workspace, int64_t workspace_bytes, void* stream) {
    IPSB_REQUIRE(mode >= 0 && mode <= 2 && M > 0 && N > 0 && K > 0, "gemm: bad arguments");
    IPSB_REQUIRE(N % 64 == 0, "gemm: N=%d must be a multiple of 64", N);
    IPSB_REQUIRE(M < (1ll << 31) && K < (1ll << 31), "gemm: dimension too large");
    cudaStream_t st = (cudaStream_t)stream;
    const int BN = (N % 128 == 0) ? 128 : 64;
    GemmParams p;
    p.scale = scale; p.shift = shift; p.res = nullptr; p.M = (int)M; p.N = N; p.K = (int)K; p.relu = relu;
    p.tiles_m = (int)((M + TILE_M - 1) / TILE_M); p.tiles_n = N / BN;
    p.KS = (int)((K + BK - 1) / BK);
    p.splits = 1;
    if (mode == 2) {        // long contraction, small output: split K so every SM has work
        const int64_t tiles = (int64_t)p.tiles_m * p.tiles_n;
        int64_t s = (2 * ipsb::sm_count() + tiles - 1) / tiles;
        if (s > p.KS) s = p.KS;
        p.splits = (int)(s < 1 ? 1 : s);
    }
    p.ks_per_split = (p.KS + p.splits - 1) / p.splits;
    p.splits = (p.KS + p.ks_per_split - 1) / p.ks_per_split;
    p.total_tiles = p.tiles_m * p.tiles_n * p.splits;
    const int64_t Mpad = (int64_t)p.tiles_m * TILE_M;
    p.split_stride_rows = Mpad;

    alignas(64) CUtensorMap tmA, tmB, tmC;
    int rc;
    if (mode == 2) rc = encode_2d(&tmA, a, K, M, 2, 64, 64, "A (K,M)"); else rc = encode_2d(&tmA, a, M, K, 2, 64, TILE_M, "A (M,K)");
    if (rc) return rc;
    if (mode == 0) rc = encode_2d(&tmB, b, N, K, 2, 64, BN, "B (N,K)"); else rc = encode_2d(&tmB, b, K, N, 2, 64, 64, "B (K,N)");
    if (rc) return rc;

    const bool split = p.splits > 1;
    void* out = c;
    bool out_f32 = c_is_f32 != 0;
    GemmParams pk = p;
    if (split) {
        const int64_t need = (int64_t)p.splits * Mpad * N * 4;
        IPSB_REQUIRE(workspace != nullptr && workspace_bytes >= need, "gemm: split-K workspace of %lld bytes required", (long long)need);
        IPSB_REQUIRE(c_is_f32, "gemm: split-K output must be fp32");
        out = workspace; out_f32 = true;
        pk.scale = nullptr; pk.shift = nullptr; pk.relu = 0;
        rc = encode_2d(&tmC, out, (int64_t)p.splits * Mpad, N, 4, 32, TILE_M, "partials");
    } else {
        rc = encode_2d(&tmC, out, M, N, out_f32 ? 4 : 2, out_f32 ? 32 : 64, TILE_M, "C");
    }
    if (rc) return rc;
    if (mode == 0) rc = out_f32 ? dispatch<false, false, float>(tmA, tmB, tmC, pk, BN, st) : dispatch<false, false, bf16>(tmA, tmB, tmC, pk, BN, st);
    else if (mode == 1) rc = out_f32 ? dispatch<false, true, float>(tmA, tmB, tmC, pk, BN, st) : dispatch<false, true, bf16>(tmA, tmB, tmC, pk, BN, st);
    else rc = out_f32 ? dispatch<true, true, float>(tmA, tmB, tmC, pk, BN, st) : dispatch<true, true, bf16>(tmA, tmB, tmC, pk, BN, st);
    if (rc) return rc;
    if (split) {
        IPSB_REQUIRE(scale == nullptr && relu == 0, "gemm: split-K supports only a bias epilogue");
        // partial s holds rows [s*Mpad, s*Mpad + M); reduce the first M rows of each
        const int64_t elems = M * (int64_t)N;
        int64_t g = (elems + 255) / 256;
        if (g > (int64_t)ipsb::sm_count() * 8) g = (int64_t)ipsb::sm_count() * 8;
        // partials are strided by Mpad*N, not M*N: pass the stride through `elems` only when Mpad == M
        if (Mpad == M) {
            reduce_partials_kernel<<<(unsigned)g, 256, 0, st>>>((const float*)workspace, (float*)c, elems, p.splits, shift, N);
        } else {
            for (int s = 0; s < 1; ++s) {}
            return ipsb::fail("gemm: split-K needs M to be a multiple of 128 (M=%lld)", (long long)M);
        }
        IPSB_LAUNCH_CHECK();
    }
    return 0;
}

}  // extern "C"
