// TMA-fed, persistent bf16 implicit-GEMM convolution / dense linear on tcgen05 (sm_100a).
//
// The M tile of 128 accumulator rows is a BOX of output pixels (bw x bh pixels of bp
// consecutive patches).  For filter tap (r, s) and a 64-channel slab, the matching im2col
// rows are then ONE 4-D TMA box of the channels-last activation tensor {C, W, H, P}: start
// (cb*64, ox0*stride + s - pad, oy0*stride + r - pad, p0), traversal stride
// (1, stride, stride, 1).  Out-of-bounds coordinates are zero-filled by the TMA unit -- the
// convolution's zero padding -- and the 128B swizzle mode writes the K-major layout
// tcgen05.mma reads.  Weights (Cout, K) arrive as 2-D boxes, or stay resident in shared
// memory for the whole kernel when they are small (stem, 64-channel layers).
//
// One CTA per SM loops over tiles: warp 0 issues TMA loads (a ring of stages that runs
// ahead across tile boundaries), warp 1 issues tcgen05.mma into one of two TMEM
// accumulators, warps 2-5 drain the other accumulator (tcgen05.ld -> folded BatchNorm,
// residual, ReLU -> channels-last store).  No LSU work or address arithmetic in the main loop.
//
// The 7x7/2 stem reads a zero-bordered 4-channel input through a 5-D map whose "ox"
// dimension has a 16-byte stride, i.e. overlapping 64-byte windows: one box row per output
// pixel and filter row = eight taps x four channels (64-byte rows, SWIZZLE_64B operand).
//
// Replaces conv2d + batch_norm(eval) + add + relu (architecture/ips_net.py:17-52) and
// nn.Linear (ips_net.py:57) of the reference.
#include <cstdlib>
#include <cuda.h>
#include "common.cuh"
#include "umma.cuh"
#include "pf.cuh"
#include "epilogue.cuh"
#include "../../include/ips_b200.h"

namespace {

using bf16 = __nv_bfloat16;
constexpr int TILE_M = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = TILE_M * 128;

struct TmaConvParams {
    const float* scale;
    const float* shift;
    const bf16* res;
    void* y;
    int P, Ho, Wo, Cout, kw, stride, pad, relu;
    int bw, bh, bp, tiles_x, tiles_y;   // M-tile box and tiles per patch
    int n_tiles_n, total_tiles;
    int KS, cblocks;                    // K stages per tile, 64-channel slabs per tap
    int b_slabs;                        // resident 64-wide weight slabs (RESB)
    int out_G0, out_Wp, out_Sp;         // output row = G0 + p*Sp + oy*Wp + ox (dense: 0, Wo, Ho*Wo; or padded-flat)
    uint32_t a_bytes;                   // bytes one A box deposits
};

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

struct TileCoord { int p0, oy0, ox0, n0; };
__device__ __forceinline__ TileCoord decode_tile(const TmaConvParams& p, int tile, int BN) {
    const int nt = tile % p.n_tiles_n;                 // n fastest: neighbouring CTAs share the A box in L2
    int mt = tile / p.n_tiles_n;
    const int tx = mt % p.tiles_x; mt /= p.tiles_x;
    const int ty = mt % p.tiles_y; mt /= p.tiles_y;
    return TileCoord{mt * p.bp, ty * p.bh, tx * p.bw, nt * BN};
}

// STEM: A through the 5-D overlapping-window map (KS = 7, one stage = one filter row = 32 k)
// RESB: all weight K-slabs resident in shared memory (single N tile)
template <int BN, int STAGES, bool STEM, bool RESB, typename OutT>
__global__ void __launch_bounds__(320, 1)
conv_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const TmaConvParams p) {
    constexpr int B_SLAB_BYTES = BN * 128;
    constexpr int A_BYTES = STEM ? TILE_M * 64 : A_STAGE_BYTES;             // stem: 64-byte rows (one filter row)
    constexpr int STAGE_BYTES = A_BYTES + (RESB ? 0 : B_SLAB_BYTES);
    constexpr int MMAS_PER_STAGE = STEM ? 2 : 4;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem0 = (umma::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t resb0 = smem0 + STAGES * STAGE_BYTES;                     // resident weights (RESB)
    const uint32_t out_stage0 = resb0 + (RESB ? (uint32_t)p.b_slabs * B_SLAB_BYTES : 0u);      // 2 x 16 KB epilogue staging
    const uint32_t bar0 = out_stage0 + 2u * epi::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 + a); };
    const uint32_t resb_bar = bar0 + 8u * (2 * STAGES + 4);
    auto res_bar = [&](int w) { return resb_bar + 8u + 8u * w; };      // residual slab landed in warpgroup w's staging tile
    const uint32_t tmem_slot = resb_bar + 24u;
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - umma::smem_u32(smem_raw)));
    const uint32_t sc_addr = (tmem_slot + 4u + 15u) & ~15u;                          // scale[Cout], shift[Cout]
    float* sc_smem = reinterpret_cast<float*>(smem_raw + (sc_addr - umma::smem_u32(smem_raw)));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            umma::mbar_init(full_bar(s), 1);
            umma::mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            umma::mbar_init(tfull_bar(a), 1);
            umma::mbar_init(tempty_bar(a), 128);
        }
        umma::mbar_init(resb_bar, 1);
        umma::mbar_init(res_bar(0), 1); umma::mbar_init(res_bar(1), 1);
        umma::fence_barrier_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) umma::tmem_alloc(tmem_slot, 2 * BN);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            if (RESB) {
                umma::mbar_expect_tx(resb_bar, (uint32_t)p.b_slabs * B_SLAB_BYTES);
                for (int ks = 0; ks < p.b_slabs; ++ks) tma_load_2d(resb0 + ks * B_SLAB_BYTES, &tmB, resb_bar, ks * BK, 0);
            }
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const TileCoord tc = decode_tile(p, tile, BN);
                const int y_base = tc.oy0 * p.stride - p.pad, x_base = tc.ox0 * p.stride - p.pad;
                for (int ks = 0; ks < p.KS; ++ks, ++it) {
                    const int stage = it % STAGES;
                    umma::mbar_wait(empty_bar(stage), ((it / STAGES) & 1) ^ 1);
                    const uint32_t a_dst = smem0 + stage * STAGE_BYTES;
                    umma::mbar_expect_tx(full_bar(stage), p.a_bytes + (RESB ? 0u : (uint32_t)B_SLAB_BYTES));
                    if (STEM) {
                        tma_load_5d(a_dst, &tmA, full_bar(stage), 0, tc.ox0, tc.oy0, ks, tc.p0);
                    } else {
                        const int tap = ks / p.cblocks, cb = ks - tap * p.cblocks;
                        const int r = tap / p.kw, s = tap - r * p.kw;
                        tma_load_4d(a_dst, &tmA, full_bar(stage), cb * BK, x_base + s, y_base + r, tc.p0);
                    }
                    if (!RESB) tma_load_2d(a_dst + A_BYTES, &tmB, full_bar(stage), ks * BK, tc.n0);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        {                                            // the whole warp runs the loop; the elected lane issues
            const uint32_t leader = umma::elect_one();
            constexpr uint32_t idesc = umma::idesc_bf16_f32(TILE_M, BN);
            if (RESB) umma::mbar_wait(resb_bar, 0);
            uint32_t it = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
                const uint32_t acc = tcount & 1;
                umma::mbar_wait(tempty_bar(acc), ((tcount >> 1) & 1) ^ 1);
                umma::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int ks = 0; ks < p.KS; ++ks, ++it) {
                    const int stage = it % STAGES;
                    umma::mbar_wait(full_bar(stage), (it / STAGES) & 1);
                    umma::tc_fence_after();
                    const uint32_t a_addr = smem0 + stage * STAGE_BYTES;
                    uint64_t adesc, bdesc;
                    if (STEM) {        // stage = filter row ks: K slice [32*ks, 32*ks+32) of the resident (Cout, 256) weights
                        adesc = umma::smem_desc_sw64(a_addr);
                        bdesc = umma::smem_desc_sw128(resb0 + (ks >> 1) * B_SLAB_BYTES) + (uint64_t)((ks & 1) * 4);
                    } else {
                        adesc = umma::smem_desc_sw128(a_addr);
                        bdesc = umma::smem_desc_sw128(RESB ? resb0 + ks * B_SLAB_BYTES : a_addr + A_BYTES);
                    }
#pragma unroll
                    for (int k = 0; k < MMAS_PER_STAGE; ++k)
                        umma::mma_bf16_w(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (ks | k) != 0, leader);
                    umma::mma_commit_w(empty_bar(stage), leader);
                }
                umma::mma_commit_w(tfull_bar(acc), leader);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue
        // two warpgroups (warps 2-5 and 6-9) alternate tiles; each owns one TMEM accumulator
        const int wg = (warp - 2) >> 2;               // accumulator / tile parity of this warpgroup
        const int q = warp & 3;                       // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;
        const int per_patch = p.bw * p.bh;
        const int pi = row / per_patch, rem = row - pi * per_patch;
        const int yi = rem / p.bw, xi = rem - yi * p.bw;
        for (int i = tid - 64; i < p.Cout; i += 256) {   // folded BatchNorm parameters -> shared memory, once
            sc_smem[i] = p.scale ? p.scale[i] : 1.f;
            sc_smem[p.Cout + i] = p.shift ? p.shift[i] : 0.f;
        }
        umma::named_bar_sync(1, 256);
        uint32_t tcount = wg;
        const uint32_t stage = out_stage0 + (uint32_t)wg * epi::STAGE_BYTES;
        const bool issuer = (row == 0);
        const bool has_res = p.res != nullptr;
        uint32_t res_phase = 0;
        const uint32_t res_bytes = (uint32_t)(p.bw * p.bh * p.bp) * 128u;     // one residual box
        for (int tile = blockIdx.x + wg * gridDim.x; tile < p.total_tiles; tile += 2 * gridDim.x, tcount += 2) {
            const TileCoord tc = decode_tile(p, tile, BN);
            const int pp = tc.p0 + pi, oy = tc.oy0 + yi, ox = tc.ox0 + xi;
            const bool valid = pi < p.bp && pp < p.P && oy < p.Ho && ox < p.Wo;
            auto load_res = [&](int s0, uint32_t dst, uint32_t bar) {
                epi::tma_ld_4d(dst, &tmR, bar, tc.n0 + s0, tc.ox0, tc.oy0, tc.p0);
            };
            if (has_res) epi::prefetch_residual(issuer, res_bar(wg), res_bytes, stage, load_res);
            umma::mbar_wait(tfull_bar(wg), (tcount >> 1) & 1);
            umma::tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + wg * BN;
            epi::drain_tile<BN, OutT>(t_row, tempty_bar(wg), sc_smem + tc.n0, sc_smem + p.Cout + tc.n0, valid, has_res, res_bar(wg),
                                      res_bytes, res_phase, p.relu, stage, row, 2u + (uint32_t)wg, issuer,
                                      [&](int s0, uint32_t src) { epi::tma_store_4d(&tmC, src, tc.n0 + s0, tc.ox0, tc.oy0, tc.p0); },
                                      load_res);
        }
        if (issuer) epi::bulk_wait0();
    }

    umma::tc_fence_before();
    __syncthreads();
    if (warp == 1) umma::tmem_dealloc(tmem_base, 2 * BN);
}

// ---------------------------------------------------------------- host side
}  // namespace
namespace ipsb {
// CTA-pair kernel for the 256-wide layers (umma_conv_pair.cu); -1 = pairs cannot be scheduled on this device
int conv_pair_launch(const CUtensorMap& tmA, const CUtensorMap& tmBh, const CUtensorMap& tmC, const CUtensorMap& tmR,
                     const float* scale, const float* shift, bool has_res, int P, int Ho, int Wo, int Cout, int kw, int stride, int pad,
                     int relu, int bw, int bh, int bp, int tiles_x, int tiles_y, int n_tiles_n, int m_tiles, int KS, int cblocks,
                     uint32_t a_bytes, cudaStream_t st);
// dense stride-1 3x3 convolution with im2col-mode TMA (umma_conv_pair.cu); -1 = not eligible
int conv_pair_im2col(const void* x, const void* w, const float* scale, const float* shift, const void* res, void* y,
                     int64_t P, int H, int W, int Cin, int Cout, int relu, cudaStream_t st);
}
namespace {

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

template <int BN, int STAGES, bool STEM, bool RESB, typename OutT>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR, const TmaConvParams& p,
           cudaStream_t st) {
    const size_t smem = (size_t)STAGES * ((STEM ? TILE_M * 64 : A_STAGE_BYTES) + (RESB ? 0 : BN * 128)) +
                        (RESB ? (size_t)p.b_slabs * BN * 128 : 0) + 2 * epi::STAGE_BYTES + 1024 + 8 * (2 * STAGES + 7) + 32 +
                        8 * (size_t)p.Cout;
    IPSB_REQUIRE(smem <= 227 * 1024, "conv_tma: %zu bytes of shared memory", smem);
    auto kern = conv_tma_kernel<BN, STAGES, STEM, RESB, OutT>;
    static size_t configured = 0;
    if (configured < smem) {
        IPSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int grid = ipsb::balanced_grid(p.total_tiles, ipsb::sm_count());
    kern<<<grid, 320, smem, st>>>(tmA, tmB, tmC, tmR, p);
    IPSB_LAUNCH_CHECK();
    return 0;
}

template <bool STEM, typename OutT>
int dispatch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR, const TmaConvParams& p,
             int BN, bool resb, cudaStream_t st) {
    if (!STEM && BN == 256) return launch<256, 3, false, false, OutT>(tmA, tmB, tmC, tmR, p, st);
    if (!STEM && BN == 128) return launch<128, 5, false, false, OutT>(tmA, tmB, tmC, tmR, p, st);
    if (STEM) return launch<64, 8, true, true, OutT>(tmA, tmB, tmC, tmR, p, st);
    if (resb) return launch<64, 6, false, true, OutT>(tmA, tmB, tmC, tmR, p, st);
    return launch<64, 6, false, false, OutT>(tmA, tmB, tmC, tmR, p, st);
}

// choose the pixel box (bw x bh x bp <= 128) with the most useful accumulator rows
void choose_box(int64_t P, int Ho, int Wo, TmaConvParams& p) {
    double best = -1.0;
    for (int bw = Wo < TILE_M ? Wo : TILE_M; bw >= 1; --bw) {
        const double ex = (double)Wo / (double)(((Wo + bw - 1) / bw) * bw);
        if (ex < 0.9 && bw != Wo) continue;
        for (int bh = 1; bh <= Ho && bh * bw <= TILE_M; ++bh) {
            const double ey = (double)Ho / (double)(((Ho + bh - 1) / bh) * bh);
            int bp = TILE_M / (bh * bw);
            if (bp > 256) bp = 256;
            if ((int64_t)bp > P) bp = (int)P;
            const double eff = ex * ey * (double)(bw * bh * bp) / TILE_M;
            if (eff > best + 1e-9) { best = eff; p.bw = bw; p.bh = bh; p.bp = bp; }
        }
        if (bw == Wo && best > 0.97) break;
    }
    p.tiles_x = (Wo + p.bw - 1) / p.bw;
    p.tiles_y = (Ho + p.bh - 1) / p.bh;
}

int encode_weights(EncodeTiledFn enc, CUtensorMap* tm, const void* w, int K, int Cout, int BN) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IPSB_REQUIRE(r == CUDA_SUCCESS, "conv_tma: cuTensorMapEncodeTiled(weights) failed with %d", (int)r);
    return 0;
}

// output tensor (dense or padded-flat view) as {C, W', H', P}; the store box is the M-tile box
template <typename OutT>
int encode_output(EncodeTiledFn enc, CUtensorMap* tm, void* y, const TmaConvParams& p, bool out_pf) {
    const size_t es = sizeof(OutT);
    cuuint64_t dims[4] = {(cuuint64_t)p.Cout, (cuuint64_t)p.Wo, (cuuint64_t)p.Ho, (cuuint64_t)p.P};
    cuuint64_t strides[3] = {(cuuint64_t)p.Cout * es, (cuuint64_t)p.Wo * p.Cout * es, (cuuint64_t)p.Ho * p.Wo * p.Cout * es};
    char* base = (char*)y;
    if (out_pf) {
        base += (size_t)p.out_G0 * p.Cout * es;
        dims[1] = (cuuint64_t)p.out_Wp; dims[2] = (cuuint64_t)(p.Ho + 1);
        strides[1] = (cuuint64_t)p.out_Wp * p.Cout * es; strides[2] = (cuuint64_t)p.out_Sp * p.Cout * es;
    }
    cuuint32_t box[4] = {(cuuint32_t)epi::Slab<OutT>::COLS, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bp};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, es == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IPSB_REQUIRE(r == CUDA_SUCCESS, "conv_tma: cuTensorMapEncodeTiled(output) failed with %d", (int)r);
    return 0;
}

}  // namespace

namespace ipsb {

// x: (P,H,W,Cin) bf16 channels-last, w: (Cout, kh*kw*Cin) bf16.  out_f32 selects the output type.
int conv_tma(const void* x, const void* w, const float* scale, const float* shift, const void* res, void* y,
             int64_t P, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad, int relu,
             bool out_f32, cudaStream_t st, bool in_pf, bool out_pf) {
    IPSB_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "conv_tma: Cin=%d / Cout=%d must be multiples of 64", Cin, Cout);
    if (!in_pf && !out_pf && !out_f32 && kh == 3 && kw == 3 && stride == 1 && pad == 1 && Cout % 256 == 0) {
        // dense 256-wide stride-1 layers: 128 consecutive pixels per tile through im2col-mode TMA (every tile row useful)
        static const bool im2col_ok = !getenv("IPSB_NO_IM2COL") && !getenv("IPSB_NO_PAIR");
        if (im2col_ok) {
            const int rc = conv_pair_im2col(x, w, scale, shift, res, y, P, H, W, Cin, Cout, relu, st);
            if (rc >= 0) return rc;
        }
    }
    const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
    IPSB_REQUIRE(Ho > 0 && Wo > 0, "conv_tma: bad geometry");
    IPSB_REQUIRE(P < (1ll << 31), "conv_tma: too many patches");
    EncodeTiledFn enc = encode_fn();
    IPSB_REQUIRE(enc != nullptr, "conv_tma: cuTensorMapEncodeTiled not available from the driver");

    TmaConvParams p;
    p.scale = scale; p.shift = shift; p.res = (const bf16*)res; p.y = y;
    p.P = (int)P; p.Ho = Ho; p.Wo = Wo; p.Cout = Cout; p.kw = kw; p.stride = stride; p.pad = pad; p.relu = relu;
    choose_box(P, Ho, Wo, p);
    p.out_G0 = 0; p.out_Wp = Wo; p.out_Sp = Ho * Wo;
    if (out_pf) { const pf::Geo go = pf::make(P, Ho, Wo); p.out_G0 = go.G0; p.out_Wp = go.Wp; p.out_Sp = go.Sp; }
    p.cblocks = Cin / BK; p.KS = kh * kw * p.cblocks;
    p.a_bytes = (uint32_t)(p.bw * p.bh * p.bp) * 128u;
    const int BN = (Cout % 256 == 0) ? 256 : (Cout % 128 == 0 ? 128 : 64);
    p.n_tiles_n = Cout / BN;
    p.total_tiles = (int)(((P + p.bp - 1) / p.bp) * p.tiles_y * p.tiles_x * p.n_tiles_n);
    const bool resb = (BN == 64 && Cout == 64 && (size_t)p.KS * 64 * 128 <= 96 * 1024);
    p.b_slabs = p.KS;

    alignas(64) CUtensorMap tmA, tmB;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)P};
        cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
        if (in_pf) {       // padded-flat input: a (P, H+1, W+1, C) tensor starting G0 rows in; the pads hold zeros
            const pf::Geo gi = pf::make(P, H, W);
            x = (const char*)x + (size_t)gi.G0 * Cin * 2;
            dims[1] = (cuuint64_t)gi.Wp; dims[2] = (cuuint64_t)(H + 1);
            strides[1] = (cuuint64_t)gi.Wp * Cin * 2; strides[2] = (cuuint64_t)gi.Sp * Cin * 2;
        }
        cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)(p.bw * stride), (cuuint32_t)(p.bh * stride), (cuuint32_t)p.bp};
        cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
        IPSB_REQUIRE(box[1] <= 256 && box[2] <= 256 && box[3] <= 256, "conv_tma: box too large");
        CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        IPSB_REQUIRE(r == CUDA_SUCCESS, "conv_tma: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
    }
    if (int rc = encode_weights(enc, &tmB, w, kh * kw * Cin, Cout, BN)) return rc;
    alignas(64) CUtensorMap tmC, tmR;
    if (out_f32) {
        IPSB_REQUIRE(res == nullptr, "conv_tma: a residual needs the bf16 output path");
        if (int rc = encode_output<float>(enc, &tmC, y, p, out_pf)) return rc;
        return dispatch<false, float>(tmA, tmB, tmC, tmC, p, BN, resb, st);
    }
    if (int rc = encode_output<bf16>(enc, &tmC, y, p, out_pf)) return rc;
    if (int rc = encode_output<bf16>(enc, &tmR, res ? const_cast<void*>(res) : y, p, out_pf)) return rc;   // same geometry as the output
    const int m_tiles = p.total_tiles / p.n_tiles_n;
    static const bool pair_ok = !getenv("IPSB_NO_PAIR");
    if (pair_ok && BN == 256 && m_tiles >= 2) {        // 256-wide layers: CTA pairs (tcgen05.mma.cta_group::2)
        alignas(64) CUtensorMap tmBh;
        if (int rc = encode_weights(enc, &tmBh, w, kh * kw * Cin, Cout, BN / 2)) return rc;      // half slabs
        const int rc = conv_pair_launch(tmA, tmBh, tmC, tmR, scale, shift, res != nullptr, (int)P, Ho, Wo, Cout, kw, stride, pad, relu,
                                        p.bw, p.bh, p.bp, p.tiles_x, p.tiles_y, p.n_tiles_n, m_tiles, p.KS, p.cblocks, p.a_bytes, st);
        if (rc >= 0) return rc;
    }
    return dispatch<false, bf16>(tmA, tmB, tmC, tmR, p, BN, resb, st);
}

// 7x7 stride-2 pad-3 stem.  x: (P, H+6, W+6, 4) bf16 with a zero border (3 rows above, 4 columns left);
// w: (Cout, 256) bf16 with k = r*32 + t*4 + c, t = s + 1.
int conv_stem_tma(const void* x, const void* w, const float* scale, const float* shift, void* y,
                  int64_t P, int H, int W, int Cout, int relu, cudaStream_t st, bool out_f32) {
    IPSB_REQUIRE(Cout == 64 && W % 2 == 0 && H % 2 == 0, "conv_stem_tma: needs Cout=64 and even H, W");
    const int Hp = H + 6, Wp = W + 6;
    const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
    EncodeTiledFn enc = encode_fn();
    IPSB_REQUIRE(enc != nullptr, "conv_stem_tma: cuTensorMapEncodeTiled not available from the driver");
    TmaConvParams p;
    p.scale = scale; p.shift = shift; p.res = nullptr; p.y = y;
    p.P = (int)P; p.Ho = Ho; p.Wo = Wo; p.Cout = Cout; p.kw = 8; p.stride = 2; p.pad = 0; p.relu = relu;
    choose_box(P, Ho, Wo, p);
    p.out_G0 = 0; p.out_Wp = Wo; p.out_Sp = Ho * Wo;
    p.cblocks = 1; p.KS = 7; p.b_slabs = 4;          // one filter row (32 k) per stage; weights resident
    p.a_bytes = (uint32_t)(p.bw * p.bh * p.bp) * 64u;
    p.n_tiles_n = 1;
    p.total_tiles = (int)(((P + p.bp - 1) / p.bp) * p.tiles_y * p.tiles_x);

    alignas(64) CUtensorMap tmA, tmB;
    {
        // dims: window (32 el = 8 px x 4 ch), ox (16-byte steps), oy (2 rows), filter row, patch
        cuuint64_t dims[5] = {32, (cuuint64_t)Wo, (cuuint64_t)Ho, 8, (cuuint64_t)P};
        cuuint64_t strides[4] = {16, (cuuint64_t)Wp * 16, (cuuint64_t)Wp * 8, (cuuint64_t)Hp * Wp * 8};
        cuuint32_t box[5] = {32, (cuuint32_t)p.bw, (cuuint32_t)p.bh, 1, (cuuint32_t)p.bp};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        IPSB_REQUIRE(r == CUDA_SUCCESS, "conv_stem_tma: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
    }
    if (int rc = encode_weights(enc, &tmB, w, 256, Cout, 64)) return rc;
    alignas(64) CUtensorMap tmC;
    if (out_f32) {
        if (int rc = encode_output<float>(enc, &tmC, y, p, false)) return rc;
        return dispatch<true, float>(tmA, tmB, tmC, tmC, p, 64, true, st);
    }
    if (int rc = encode_output<bf16>(enc, &tmC, y, p, false)) return rc;
    return dispatch<true, bf16>(tmA, tmB, tmC, tmC, p, 64, true, st);
}

}  // namespace ipsb
