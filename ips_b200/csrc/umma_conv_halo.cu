// Stride-1 3x3 convolution on padded-flat (PF) activations: shifted-window implicit GEMM.
//
// An M tile is 128 CONSECUTIVE PF rows g0 .. g0+127 (pixels and pad rows alike).  The block
// of input rows [g0 - Wp - 1, g0 + 128 + Wp + 1) is fetched ONCE per 64-channel slab with a
// single 2-D TMA box; the im2col operand of filter tap (r, s) is that block viewed from row
// offset r*Wp + s, which for a 128B-swizzled K-major operand is nothing but a different
// descriptor start address (see pf.cuh).  Activations are therefore read from L2/HBM once
// per layer instead of nine times, and the main loop is MMA-bound.
//
//   warp 0   TMA producer: A blocks (ring) + weight slabs (ring, or resident when they fit)
//   warp 1   tcgen05.mma issuer, two TMEM accumulators
//   warps 2-9 epilogue (two warpgroups alternating tiles): TMEM -> folded BN, residual, ReLU -> PF rows
//            (pad rows written as zero)
//
// Replaces conv2d(3x3, stride 1) + batch_norm(eval) + add + relu of the reference's BasicBlocks
// (architecture/ips_net.py:17-52).
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "umma.cuh"
#include "pf.cuh"
#include "epilogue.cuh"
#include "../../include/ips_b200.h"

namespace {

using bf16 = __nv_bfloat16;
constexpr int TILE_M = 128;
constexpr int BK = 64;

struct HaloParams {
    const float* scale;
    const float* shift;
    const bf16* res;      // PF, same geometry as the output, or null
    bf16* y;              // PF output
    int P, H, W, Wp, Sp, G0, Cout, relu;
    int cblocks;          // Cin / 64
    int total_tiles;
    int a_rows;           // rows per TMA box of the A block (single tile: 130 + 2*Wp)
    int n_boxes;          // boxes per A block (2 when two tiles share a block)
    uint32_t a_slot_bytes;
    int64_t pix_rows;     // P * Sp: rows [G0, G0 + pix_rows) hold pixels / pads of real patches
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

// K-major operand with 32-byte rows (16 bf16): 8-row atoms of 256 bytes, SWIZZLE_32B (tools/probe_sw32.cu)
__device__ __forceinline__ uint64_t desc_sw32(uint32_t addr) {
    const uint32_t lo = ((addr >> 4) & 0x3fffu) | (1u << 16);
    const uint32_t hi = (256u >> 4) | (1u << 14) | (6u << 29);
    return ((uint64_t)hi << 32) | lo;
}

// epilogue warpgroups: the stem (16 short MMAs per tile) is epilogue-bound with two, so it gets four
template <bool S2D> struct EpiWgs { static constexpr int N = S2D ? 4 : 2; };

template <int BN, int SA, int SB, bool RESB, bool PAIR, bool S2D>
__global__ void __launch_bounds__(64 + 128 * EpiWgs<S2D>::N, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const HaloParams p) {
    constexpr int B_SLAB_BYTES = BN * 128;
    constexpr int NWG = EpiWgs<S2D>::N;              // epilogue warpgroups
    constexpr int NACC = PAIR ? 2 : NWG;             // TMEM accumulator stages (pairs: two stages of two tiles)
    static_assert(!PAIR || NWG == 2, "paired tiles use two warpgroups");
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem0 = (umma::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b0 = smem0 + SA * p.a_slot_bytes;                                   // weight ring or resident slabs
    const uint32_t out_stage0 = b0 + (RESB ? (uint32_t)(S2D ? 4 : 9 * p.cblocks) : (uint32_t)SB) * B_SLAB_BYTES;   // NWG x 16 KB epilogue staging
    const uint32_t bar0 = out_stage0 + (uint32_t)NWG * epi::STAGE_BYTES;
    auto a_full = [&](int s) { return bar0 + 8u * s; };
    auto a_empty = [&](int s) { return bar0 + 8u * (SA + s); };
    auto b_full = [&](int s) { return bar0 + 8u * (2 * SA + s); };
    auto b_empty = [&](int s) { return bar0 + 8u * (2 * SA + SB + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * SA + 2 * SB + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * SA + 2 * SB + NACC + a); };
    const uint32_t resb_bar = bar0 + 8u * (2 * SA + 2 * SB + 2 * NACC);
    auto res_bar = [&](int w) { return resb_bar + 8u + 8u * w; };      // residual slab landed in warpgroup w's staging tile
    const uint32_t tmem_slot = resb_bar + 8u + 8u * NWG;
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - umma::smem_u32(smem_raw)));
    const uint32_t sc_addr = (tmem_slot + 4u + 15u) & ~15u;
    float* sc_smem = reinterpret_cast<float*>(smem_raw + (sc_addr - umma::smem_u32(smem_raw)));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < SA; ++s) { umma::mbar_init(a_full(s), 1); umma::mbar_init(a_empty(s), 1); }
        for (int s = 0; s < SB; ++s) { umma::mbar_init(b_full(s), 1); umma::mbar_init(b_empty(s), 1); }
        for (int a = 0; a < NACC; ++a) { umma::mbar_init(tfull_bar(a), 1); umma::mbar_init(tempty_bar(a), PAIR ? 256 : 128); }
        umma::mbar_init(resb_bar, 1);
        for (int w = 0; w < NWG; ++w) umma::mbar_init(res_bar(w), 1);
        umma::fence_barrier_init();
    }
    constexpr int MT = PAIR ? 2 : 1;                 // M tiles computed per pipeline step (they share every weight slab)
    if (warp == 1) umma::tmem_alloc(tmem_slot, NACC * MT * BN);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    // S2D (7x7/2 stem as a 4x4 stride-1 convolution on the 2x2 space-to-depth input): 16 taps of 16 channels,
    // A rows of 32 bytes (SWIZZLE_32B), the (Cout, 256) weights resident as four 64-wide slabs
    constexpr int ROW_BYTES = S2D ? 32 : 128;
    constexpr int NTAPS = S2D ? 16 : 9;
    const int n_slabs = S2D ? 4 : 9 * p.cblocks;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            if (RESB) {
                umma::mbar_expect_tx(resb_bar, (uint32_t)n_slabs * B_SLAB_BYTES);
                for (int ks = 0; ks < n_slabs; ++ks) tma_load_2d(b0 + ks * B_SLAB_BYTES, &tmB, resb_bar, ks * BK, 0);
            }
            uint32_t ia = 0, ib = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {     // PAIR: `tile` counts pairs of M tiles
                for (int cb = 0; cb < p.cblocks; ++cb, ++ia) {
                    const int sa = ia % SA;
                    umma::mbar_wait(a_empty(sa), ((ia / SA) & 1) ^ 1);
                    umma::mbar_expect_tx(a_full(sa), (uint32_t)(p.a_rows * p.n_boxes) * ROW_BYTES);
                    for (int bx = 0; bx < p.n_boxes; ++bx)
                        tma_load_2d(smem0 + sa * p.a_slot_bytes + (uint32_t)(bx * p.a_rows) * ROW_BYTES, &tmA, a_full(sa), cb * BK,
                                    tile * (MT * TILE_M) + bx * p.a_rows);
                    if (!RESB) {
                        for (int tap = 0; tap < NTAPS; ++tap, ++ib) {
                            const int sb = ib % SB;
                            umma::mbar_wait(b_empty(sb), ((ib / SB) & 1) ^ 1);
                            umma::mbar_expect_tx(b_full(sb), B_SLAB_BYTES);
                            tma_load_2d(b0 + sb * B_SLAB_BYTES, &tmB, b_full(sb), (tap * p.cblocks + cb) * BK, 0);
                        }
                    }
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
            uint32_t ia = 0, ib = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
                const uint32_t acc = tcount % NACC;
                umma::mbar_wait(tempty_bar(acc), ((tcount / NACC) & 1) ^ 1);
                umma::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (MT * BN);
                for (int cb = 0; cb < p.cblocks; ++cb, ++ia) {
                    const int sa = ia % SA;
                    umma::mbar_wait(a_full(sa), (ia / SA) & 1);
                    umma::tc_fence_after();
                    const uint32_t a_base = smem0 + sa * p.a_slot_bytes;
                    for (int tap = 0; tap < NTAPS; ++tap) {
                        if (S2D) {       // tap (a, b) of the 4x4 window: rows shifted by a*Wp + b, K slice tap*16 of the weights
                            const int ta = tap >> 2, tb = tap & 3;
                            const uint64_t adesc = desc_sw32(a_base + (uint32_t)(ta * p.Wp + tb) * 32u);
                            const uint64_t bdesc = umma::smem_desc_sw128(b0 + ta * B_SLAB_BYTES) + (uint64_t)(2 * tb);
                            umma::mma_bf16_w(d_tmem, adesc, bdesc, idesc, tap != 0, leader);
                            continue;
                        }
                        const int r = tap / 3, s = tap - 3 * r;
                        uint32_t b_addr;
                        int sb = 0;
                        if (RESB) {
                            b_addr = b0 + (tap * p.cblocks + cb) * B_SLAB_BYTES;
                        } else {
                            sb = ib % SB;
                            umma::mbar_wait(b_full(sb), (ib / SB) & 1);
                            umma::tc_fence_after();
                            b_addr = b0 + sb * B_SLAB_BYTES;
                        }
                        const uint64_t adesc = umma::smem_desc_sw128(a_base + (uint32_t)(r * p.Wp + s) * 128u);
                        const uint64_t bdesc = umma::smem_desc_sw128(b_addr);
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {            // both tiles of a pair reuse the weight slab
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k)
                                umma::mma_bf16_w(d_tmem + mt * BN, adesc + (uint64_t)(mt * (TILE_M * 128 / 16)) + 2u * k, bdesc + 2u * k,
                                                 idesc, (cb | tap | k) != 0, leader);
                        }
                        if (!RESB) { umma::mma_commit_w(b_empty(sb), leader); ++ib; }
                    }
                    umma::mma_commit_w(a_empty(sa), leader);
                }
                umma::mma_commit_w(tfull_bar(acc), leader);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue
        // two warpgroups (warps 2-5 and 6-9) alternate tiles; each owns one TMEM accumulator
        const int wg = (warp - 2) >> 2;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        for (int i = tid - 64; i < p.Cout; i += 128 * NWG) {
            sc_smem[i] = p.scale ? p.scale[i] : 1.f;
            sc_smem[p.Cout + i] = p.shift ? p.shift[i] : 0.f;
        }
        umma::named_bar_sync(1, 128 * NWG);
        // single tiles: the warpgroups alternate tiles (accumulator = tile parity); pairs: both warpgroups work on
        // every pair, warpgroup w draining the w-th tile of the pair
        uint32_t tcount = PAIR ? 0 : wg;
        const uint32_t stage = out_stage0 + (uint32_t)wg * epi::STAGE_BYTES;
        const bool issuer = (row == 0);
        const bool has_res = p.res != nullptr;
        uint32_t res_phase = 0;
        for (int step = blockIdx.x + (PAIR ? 0 : wg * gridDim.x); step < p.total_tiles;
             step += (PAIR ? 1 : NWG) * gridDim.x, tcount += (PAIR ? 1 : NWG)) {
            const int tile = PAIR ? 2 * step + wg : step;            // M tile (128 PF rows) this warpgroup drains
            const uint32_t acc = PAIR ? (tcount & 1) : (uint32_t)wg;
            const int rel = tile * TILE_M + row;                     // row index relative to G0 (fits int32: checked on the host)
            const bool in_range = rel < (int)p.pix_rows;             // rows of real patches (pixels or their pads)
            const int rem = rel % p.Sp;
            const int yy = rem / p.Wp, xx = rem - yy * p.Wp;
            const bool pixel = in_range && (S2D || (yy < p.H && xx < p.W));   // pad rows are stored as zeros (S2D: no pads)
            const int g0 = p.G0 + tile * TILE_M;
            auto load_res = [&](int s0, uint32_t dst, uint32_t bar) { epi::tma_ld_2d(dst, &tmR, bar, s0, g0); };
            if (has_res) epi::prefetch_residual(issuer, res_bar(wg), epi::STAGE_BYTES, stage, load_res);
            umma::mbar_wait(tfull_bar(acc), (tcount / NACC) & 1);
            umma::tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (PAIR ? acc * 2 * BN + wg * BN : wg * BN);
            epi::drain_tile<BN, bf16>(t_row, tempty_bar(acc), sc_smem, sc_smem + p.Cout, pixel, has_res, res_bar(wg), epi::STAGE_BYTES, res_phase,
                                      p.relu, stage, row, 2u + (uint32_t)wg, issuer,
                                      [&](int s0, uint32_t src) { epi::tma_store_2d(&tmC, src, s0, g0); }, load_res);
        }
        if (issuer) epi::bulk_wait0();
    }

    umma::tc_fence_before();
    __syncthreads();
    if (warp == 1) umma::tmem_dealloc(tmem_base, NACC * MT * BN);
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

template <int BN, int SA, int SB, bool RESB, bool PAIR, bool S2D = false>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR, const HaloParams& p,
           cudaStream_t st) {
    const size_t smem = (size_t)SA * p.a_slot_bytes + (size_t)(RESB ? (S2D ? 4 : 9 * p.cblocks) : SB) * BN * 128 + EpiWgs<S2D>::N * epi::STAGE_BYTES + 1024 +
                        8 * (2 * SA + 2 * SB + 3 * EpiWgs<S2D>::N + 1) + 32 + 8 * (size_t)p.Cout;
    IPSB_REQUIRE(smem <= 227 * 1024, "conv_halo: %zu bytes of shared memory", smem);
    auto kern = conv_halo_kernel<BN, SA, SB, RESB, PAIR, S2D>;
    static size_t configured = 0;
    if (configured < smem) {
        IPSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int grid = ipsb::balanced_grid(p.total_tiles, ipsb::sm_count());
    kern<<<grid, 64 + 128 * EpiWgs<S2D>::N, smem, st>>>(tmA, tmB, tmC, tmR, p);
    IPSB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

namespace ipsb {

// x, res, y: PF(H, W) tensors with Cin / Cout / Cout channels; w: (Cout, 9*Cin) bf16, k = (r*3+s)*Cin + c.
int conv3x3_halo(const void* x, const void* w, const float* scale, const float* shift, const void* res, void* y,
                 int64_t P, int H, int W, int Cin, int Cout, int relu, cudaStream_t st) {
    IPSB_REQUIRE(Cin % 64 == 0 && (Cout == 64 || Cout == 128), "conv3x3_halo: Cin=%d Cout=%d not supported", Cin, Cout);
    EncodeTiledFn enc = encode_fn();
    IPSB_REQUIRE(enc != nullptr, "conv3x3_halo: cuTensorMapEncodeTiled not available from the driver");
    const pf::Geo g = pf::make(P, H, W);
    HaloParams p;
    p.scale = scale; p.shift = shift; p.res = (const bf16*)res; p.y = (bf16*)y;
    p.P = (int)P; p.H = H; p.W = W; p.Wp = g.Wp; p.Sp = g.Sp; p.G0 = g.G0; p.Cout = Cout; p.relu = relu;
    p.cblocks = Cin / BK;
    p.pix_rows = P * (int64_t)g.Sp;
    IPSB_REQUIRE(p.pix_rows + TILE_M < (1ll << 31), "conv3x3_halo: too many rows");
    p.total_tiles = (int)((p.pix_rows + TILE_M - 1) / TILE_M);
    const bool pair = (Cout == 128) && getenv("IPSB_HALO_NOPAIR") == nullptr;     // streamed weights: two M tiles share every slab
    if (pair) {
        p.total_tiles = (p.total_tiles + 1) / 2;                 // pipeline steps = pairs of tiles
        p.n_boxes = 2;
        p.a_rows = (TILE_M + g.Wp + 1 + 7) / 8 * 8;              // two boxes cover 256 + 2*Wp + 2 rows (8-row aligned halves)
    } else {
        p.n_boxes = 1;
        p.a_rows = TILE_M + 2 * g.Wp + 2;
    }
    IPSB_REQUIRE(p.a_rows <= 256, "conv3x3_halo: width %d too large for one TMA box", W);
    p.a_slot_bytes = (uint32_t)((p.a_rows * p.n_boxes * 128 + 1023) / 1024 * 1024);

    alignas(64) CUtensorMap tmA, tmB;
    {
        cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)g.rows};
        cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
        cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)p.a_rows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        IPSB_REQUIRE(r == CUDA_SUCCESS, "conv3x3_halo: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
    }
    {
        const int K = 9 * Cin;
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
        cuuint64_t strides[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)Cout};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        IPSB_REQUIRE(r == CUDA_SUCCESS, "conv3x3_halo: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
    }
    alignas(64) CUtensorMap tmC, tmR;
    for (int i = 0; i < 2; ++i) {   // output (and residual) rows [G0 + pix_rows) as a 2-D tensor; rows past the last patch are clipped
        void* base = (i == 0) ? y : const_cast<void*>(res ? res : y);
        cuuint64_t dims[2] = {(cuuint64_t)Cout, (cuuint64_t)(g.G0 + p.pix_rows)};
        cuuint64_t strides[1] = {(cuuint64_t)Cout * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)TILE_M};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(i == 0 ? &tmC : &tmR, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        IPSB_REQUIRE(r == CUDA_SUCCESS, "conv3x3_halo: cuTensorMapEncodeTiled(output) failed with %d", (int)r);
    }
    if (Cout == 64) {
        if ((size_t)9 * p.cblocks * 64 * 128 <= 80 * 1024) return launch<64, 4, 1, true, false>(tmA, tmB, tmC, tmR, p, st);
        return launch<64, 4, 6, false, false>(tmA, tmB, tmC, tmR, p, st);
    }
    if (pair) return launch<128, 3, 4, false, true>(tmA, tmB, tmC, tmR, p, st);
    return launch<128, 4, 5, false, false>(tmA, tmB, tmC, tmR, p, st);
}


// 7x7/2 pad-3 stem on the space-to-depth frame of ipsb_stage_patches_s2d: x (P * Sp rows, 16) bf16 with Sp = (H/2+3)*(W/2+3);
// w (Cout=64, 256) bf16, k = (a*4+b)*16 + (dy*2+dx)*4 + c; y (P * Sp rows, 64) bf16 in the same "wide" row order
// (output (oy, ox) of patch p at row p*Sp + oy*(W/2+3) + ox; the other rows hold don't-care values).
int conv_stem_s2d(const void* x, const void* w, const float* scale, const float* shift, void* y, int64_t P, int H, int W,
                  int Cout, int relu, cudaStream_t st) {
    IPSB_REQUIRE(Cout == 64 && H % 2 == 0 && W % 2 == 0, "conv_stem_s2d: needs Cout=64 and even H, W");
    EncodeTiledFn enc = encode_fn();
    IPSB_REQUIRE(enc != nullptr, "conv_stem_s2d: cuTensorMapEncodeTiled not available from the driver");
    const int Ho = H / 2, Wo = W / 2, Wp = Wo + 3, Sp = (Ho + 3) * Wp;
    HaloParams p;
    p.scale = scale; p.shift = shift; p.res = nullptr; p.y = (bf16*)y;
    p.P = (int)P; p.H = Ho; p.W = Wo; p.Wp = Wp; p.Sp = Sp; p.G0 = 0; p.Cout = Cout; p.relu = relu;
    p.cblocks = 1;
    p.pix_rows = P * (int64_t)Sp;
    IPSB_REQUIRE(p.pix_rows + TILE_M < (1ll << 31), "conv_stem_s2d: too many rows");
    p.total_tiles = (int)((p.pix_rows + TILE_M - 1) / TILE_M);
    p.n_boxes = 2;
    p.a_rows = ((TILE_M + 3 * Wp + 3 + 1) / 2 + 7) / 8 * 8;         // two boxes cover 128 + 3*Wp + 3 rows
    IPSB_REQUIRE(p.a_rows <= 256, "conv_stem_s2d: width %d too large", W);
    p.a_slot_bytes = (uint32_t)((p.a_rows * 2 * 32 + 1023) / 1024 * 1024);
    alignas(64) CUtensorMap tmA, tmB, tmC;
    {
        cuuint64_t dims[2] = {16, (cuuint64_t)p.pix_rows};
        cuuint64_t strides[1] = {32};
        cuuint32_t box[2] = {16, (cuuint32_t)p.a_rows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        IPSB_REQUIRE(r == CUDA_SUCCESS, "conv_stem_s2d: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
    }
    {
        cuuint64_t dims[2] = {256, (cuuint64_t)Cout};
        cuuint64_t strides[1] = {512};
        cuuint32_t box[2] = {64, 64};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        IPSB_REQUIRE(r == CUDA_SUCCESS, "conv_stem_s2d: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)Cout, (cuuint64_t)p.pix_rows};
        cuuint64_t strides[1] = {(cuuint64_t)Cout * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)TILE_M};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tmC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, y, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        IPSB_REQUIRE(r == CUDA_SUCCESS, "conv_stem_s2d: cuTensorMapEncodeTiled(output) failed with %d", (int)r);
    }
    return launch<64, 6, 1, true, false, true>(tmA, tmB, tmC, tmC, p, st);
}

}  // namespace ipsb
