// Implicit-GEMM convolution for the 256-wide layers on CTA PAIRS: tcgen05.mma.cta_group::2 (M = 256 across two SMs).
//
// Same tiling as conv_tma_kernel (umma_conv_tma.cu): M tile = box of output pixels bw x bh x bp, one 4-D TMA box per
// filter tap, N tile = 256 output channels, K stage = 64 input channels of one tap.  A cluster of two CTAs works on
// two M tiles under the SAME N tile:
//   * each CTA loads its own A box (16 KB) and only HALF of the weight slab (128 of the 256 rows, 16 KB) per stage;
//     both CTAs' loads complete on the LEADER's (cluster rank 0) full barrier (cp.async.bulk.tensor ... .cta_group::2);
//   * the leader's MMA warp issues one tcgen05.mma.cta_group::2 (M256 x N256 x K16) per 16 k: every SM reads its 4 KB of
//     A and 4 KB of B per 128 tensor-clocks instead of 4 + 8 KB, and fills 32 KB instead of 48 KB per stage.  The
//     single-CTA kernel needs ~190 B/clk of shared-memory bandwidth at full tensor rate, the pair ~128 B/clk;
//   * tcgen05.commit.multicast frees the stage / publishes the accumulator in both CTAs; each CTA's two epilogue
//     warpgroups drain its own 128 TMEM lanes (BN + residual + ReLU -> TMA store) and release the accumulator with a
//     remote arrive on the leader's barrier.
// 5 stages of 32 KB per CTA, 2 x 256 TMEM columns (double-buffered accumulators) in each CTA.
//
// Two kernels share this pipeline: conv_pair_kernel (pixel-box M tiles, any stride, padded-flat or dense tensors) and
// conv_pair_im2col_kernel further down (dense stride-1 3x3 layers: 128 consecutive output pixels per tile through
// im2col-mode TMA, every tile row useful whatever the map size).
#include <cstdlib>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "umma.cuh"
#include "pf.cuh"
#include "epilogue.cuh"

namespace {

using bf16 = __nv_bfloat16;
constexpr int TILE_M = 128;
constexpr int BK = 64;
constexpr int BN = 256;
constexpr int A_BYTES = TILE_M * 128;          // 128 pixels x 64 channels bf16
constexpr int BH_BYTES = (BN / 2) * 128;       // this CTA's half of the weight slab
constexpr int STAGE_BYTES = A_BYTES + BH_BYTES;
constexpr int STAGES = 5;

struct PairParams {
    const float* scale;
    const float* shift;
    int has_res;
    int P, Ho, Wo, Cout, kw, stride, pad, relu;
    int bw, bh, bp, tiles_x, tiles_y;
    int n_tiles_n, total_pairs;
    int KS, cblocks;
    uint32_t a_bytes;
};

__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

struct TileCoord { int p0, oy0, ox0, n0; };
// unit = pair of M tiles (2*mu, 2*mu + 1) under one N tile, n fastest; CTA `rank` takes M tile 2*mu + rank
// (an M tile past the end decodes to patches >= P: loads are zero-filled, stores clipped)
__device__ __forceinline__ TileCoord decode_unit(const PairParams& p, int unit, int rank) {
    const int nt = unit % p.n_tiles_n;
    int mt = 2 * (unit / p.n_tiles_n) + rank;
    const int tx = mt % p.tiles_x; mt /= p.tiles_x;
    const int ty = mt % p.tiles_y; mt /= p.tiles_y;
    return TileCoord{mt * p.bp, ty * p.bh, tx * p.bw, nt * BN};
}

__global__ void __launch_bounds__(320, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const PairParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem0 = (umma::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t out_stage0 = smem0 + STAGES * STAGE_BYTES;                 // 2 x 16 KB epilogue staging
    const uint32_t bar0 = out_stage0 + 2u * epi::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };                     // used in the leader CTA only
    auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 + a); };  // used in the leader CTA only
    auto res_bar = [&](int w) { return bar0 + 8u * (2 * STAGES + 4 + w); };
    const uint32_t tmem_slot = bar0 + 8u * (2 * STAGES + 6);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - umma::smem_u32(smem_raw)));
    const uint32_t sc_addr = (tmem_slot + 4u + 15u) & ~15u;                   // scale[Cout], shift[Cout]
    float* sc_smem = reinterpret_cast<float*>(smem_raw + (sc_addr - umma::smem_u32(smem_raw)));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)umma::cluster_ctarank();
    const int unit0 = blockIdx.x >> 1, ustride = gridDim.x >> 1;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            umma::mbar_init(full_bar(s), 1);
            umma::mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            umma::mbar_init(tfull_bar(a), 1);
            umma::mbar_init(tempty_bar(a), 256);        // 128 epilogue threads of each CTA
        }
        umma::mbar_init(res_bar(0), 1); umma::mbar_init(res_bar(1), 1);
        umma::fence_barrier_init();
    }
    if (warp == 1) umma::tmem_alloc2(tmem_slot, 2 * BN);
    umma::tc_fence_before();
    __syncthreads();
    umma::cluster_sync_all();                           // the peer's barriers exist before anything signals them
    umma::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            uint32_t it = 0;
            for (int unit = unit0; unit < p.total_pairs; unit += ustride) {
                const TileCoord tc = decode_unit(p, unit, rank);
                const int y_base = tc.oy0 * p.stride - p.pad, x_base = tc.ox0 * p.stride - p.pad;
                for (int ks = 0; ks < p.KS; ++ks, ++it) {
                    const int stage = it % STAGES;
                    umma::mbar_wait(empty_bar(stage), ((it / STAGES) & 1) ^ 1);
                    const uint32_t a_dst = smem0 + stage * STAGE_BYTES;
                    const uint32_t lead_full = umma::mapa(full_bar(stage), 0);
                    if (rank == 0) umma::mbar_expect_tx(full_bar(stage), 2u * (p.a_bytes + (uint32_t)BH_BYTES));   // both CTAs' bytes
                    const int tap = ks / p.cblocks, cb = ks - tap * p.cblocks;
                    const int r = tap / p.kw, s = tap - r * p.kw;
                    tma2_load_4d(a_dst, &tmA, lead_full, cb * BK, x_base + s, y_base + r, tc.p0);
                    tma2_load_2d(a_dst + A_BYTES, &tmB, lead_full, ks * BK, tc.n0 + rank * (BN / 2));
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA; warp-uniform, elected lane)
        if (rank == 0) {
            const uint32_t leader = umma::elect_one();
            constexpr uint32_t idesc = umma::idesc_bf16_f32(2 * TILE_M, BN);
            uint32_t it = 0, tcount = 0;
            for (int unit = unit0; unit < p.total_pairs; unit += ustride, ++tcount) {
                const uint32_t acc = tcount & 1;
                umma::mbar_wait(tempty_bar(acc), ((tcount >> 1) & 1) ^ 1);
                umma::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int ks = 0; ks < p.KS; ++ks, ++it) {
                    const int stage = it % STAGES;
                    umma::mbar_wait(full_bar(stage), (it / STAGES) & 1);
                    umma::tc_fence_after();
                    const uint32_t a_addr = smem0 + stage * STAGE_BYTES;
                    const uint64_t adesc = umma::smem_desc_sw128(a_addr);
                    const uint64_t bdesc = umma::smem_desc_sw128(a_addr + A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma::mma2_bf16_w(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (ks | k) != 0, leader);
                    umma::mma2_commit_mc_w(empty_bar(stage), (uint16_t)0x3, leader);      // frees the stage in both CTAs
                }
                umma::mma2_commit_mc_w(tfull_bar(acc), (uint16_t)0x3, leader);             // accumulator ready in both CTAs
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue (both CTAs, own 128 TMEM lanes)
        const int wg = (warp - 2) >> 2;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int per_patch = p.bw * p.bh;
        const int pi = row / per_patch, rem = row - pi * per_patch;
        const int yi = rem / p.bw, xi = rem - yi * p.bw;
        for (int i = tid - 64; i < p.Cout; i += 256) {
            sc_smem[i] = p.scale ? p.scale[i] : 1.f;
            sc_smem[p.Cout + i] = p.shift ? p.shift[i] : 0.f;
        }
        umma::named_bar_sync(1, 256);
        uint32_t tcount = wg;
        const uint32_t stage = out_stage0 + (uint32_t)wg * epi::STAGE_BYTES;
        const bool issuer = (row == 0);
        const bool has_res = p.has_res != 0;
        uint32_t res_phase = 0;
        const uint32_t res_bytes = (uint32_t)(p.bw * p.bh * p.bp) * 128u;
        const uint32_t lead_tempty = umma::mapa(tempty_bar(wg), 0);
        for (int unit = unit0 + wg * ustride; unit < p.total_pairs; unit += 2 * ustride, tcount += 2) {
            const TileCoord tc = decode_unit(p, unit, rank);
            const int pp = tc.p0 + pi, oy = tc.oy0 + yi, ox = tc.ox0 + xi;
            const bool valid = pi < p.bp && pp < p.P && oy < p.Ho && ox < p.Wo;
            auto load_res = [&](int s0, uint32_t dst, uint32_t bar) {
                epi::tma_ld_4d(dst, &tmR, bar, tc.n0 + s0, tc.ox0, tc.oy0, tc.p0);
            };
            if (has_res) epi::prefetch_residual(issuer, res_bar(wg), res_bytes, stage, load_res);
            umma::mbar_wait(tfull_bar(wg), (tcount >> 1) & 1);
            umma::tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + wg * BN;
            epi::drain_tile<BN, bf16, true>(t_row, lead_tempty, sc_smem + tc.n0, sc_smem + p.Cout + tc.n0, valid, has_res, res_bar(wg),
                                            res_bytes, res_phase, p.relu, stage, row, 2u + (uint32_t)wg, issuer,
                                            [&](int s0, uint32_t src) { epi::tma_store_4d(&tmC, src, tc.n0 + s0, tc.ox0, tc.oy0, tc.p0); },
                                            load_res);
        }
        if (issuer) epi::bulk_wait0();
    }

    umma::tc_fence_before();
    __syncthreads();
    umma::cluster_sync_all();                           // no CTA leaves (or frees TMEM) while its peer may still signal it
    if (warp == 1) umma::tmem_dealloc2(tmem_base, 2 * BN);
}

// ---- im2col variant: stride-1 3x3 / pad 1 convolutions on DENSE (P, H, W, C) activations ------------------------------------
// M tile = 128 CONSECUTIVE output pixels of the flattened (p, y, x) order -- every tile row is useful whatever the map size
// (7x7 maps fill 98 of 128 rows of a pixel box).  One im2col-mode TMA load (cuTensorMapEncodeIm2col: corners -pad and
// pad - (k-1), 64 channels x 128 pixels; start {c, x - 1, y - 1, p}, offsets {s, r}) delivers the tap-(r, s) operand of the
// tile across image rows and images with the padding zero-filled (tools/probe_im2col.cu).  Output and residual are plain
// 2-D (pixels, channels) tensors.  Everything else is the CTA-pair pipeline above.
struct Im2colParams {
    const float* scale;
    const float* shift;
    int has_res, relu;
    int H, W, Cout, cblocks, KS;
    int n_tiles_n, total_units;      // units = pairs of M tiles x N tiles (n fastest)
    int M;                           // P * H * W output pixels
};

__device__ __forceinline__ void tma2_im2col_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h, int n,
                                               int off_w, int off_h) {
    const uint16_t a = (uint16_t)off_w, b = (uint16_t)off_h;
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(a), "h"(b)
        : "memory");
}

__global__ void __launch_bounds__(320, 1)
conv_pair_im2col_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const Im2colParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem0 = (umma::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t out_stage0 = smem0 + STAGES * STAGE_BYTES;
    const uint32_t bar0 = out_stage0 + 2u * epi::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 + a); };
    auto res_bar = [&](int w) { return bar0 + 8u * (2 * STAGES + 4 + w); };
    const uint32_t tmem_slot = bar0 + 8u * (2 * STAGES + 6);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - umma::smem_u32(smem_raw)));
    const uint32_t sc_addr = (tmem_slot + 4u + 15u) & ~15u;
    float* sc_smem = reinterpret_cast<float*>(smem_raw + (sc_addr - umma::smem_u32(smem_raw)));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)umma::cluster_ctarank();
    const int unit0 = blockIdx.x >> 1, ustride = gridDim.x >> 1;
    const int HW = p.H * p.W;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { umma::mbar_init(full_bar(s), 1); umma::mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { umma::mbar_init(tfull_bar(a), 1); umma::mbar_init(tempty_bar(a), 256); }
        umma::mbar_init(res_bar(0), 1); umma::mbar_init(res_bar(1), 1);
        umma::fence_barrier_init();
    }
    if (warp == 1) umma::tmem_alloc2(tmem_slot, 2 * BN);
    umma::tc_fence_before();
    __syncthreads();
    umma::cluster_sync_all();
    umma::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            uint32_t it = 0;
            for (int unit = unit0; unit < p.total_units; unit += ustride) {
                const int nt = unit % p.n_tiles_n;
                const int m0 = (2 * (unit / p.n_tiles_n) + rank) * TILE_M;       // first output pixel of my M tile
                const int n = m0 / HW, rem = m0 - n * HW;
                const int py = rem / p.W, qx = rem - py * p.W;
                for (int ks = 0; ks < p.KS; ++ks, ++it) {
                    const int stage = it % STAGES;
                    umma::mbar_wait(empty_bar(stage), ((it / STAGES) & 1) ^ 1);
                    const uint32_t a_dst = smem0 + stage * STAGE_BYTES;
                    const uint32_t lead_full = umma::mapa(full_bar(stage), 0);
                    if (rank == 0) umma::mbar_expect_tx(full_bar(stage), 2u * (uint32_t)STAGE_BYTES);
                    const int tap = ks / p.cblocks, cb = ks - tap * p.cblocks;
                    const int r = tap / 3, s = tap - 3 * r;
                    tma2_im2col_4d(a_dst, &tmA, lead_full, cb * BK, qx - 1, py - 1, n, s, r);
                    tma2_load_2d(a_dst + A_BYTES, &tmB, lead_full, ks * BK, nt * BN + rank * (BN / 2));
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA)
        if (rank == 0) {
            const uint32_t leader = umma::elect_one();
            constexpr uint32_t idesc = umma::idesc_bf16_f32(2 * TILE_M, BN);
            uint32_t it = 0, tcount = 0;
            for (int unit = unit0; unit < p.total_units; unit += ustride, ++tcount) {
                const uint32_t acc = tcount & 1;
                umma::mbar_wait(tempty_bar(acc), ((tcount >> 1) & 1) ^ 1);
                umma::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int ks = 0; ks < p.KS; ++ks, ++it) {
                    const int stage = it % STAGES;
                    umma::mbar_wait(full_bar(stage), (it / STAGES) & 1);
                    umma::tc_fence_after();
                    const uint32_t a_addr = smem0 + stage * STAGE_BYTES;
                    const uint64_t adesc = umma::smem_desc_sw128(a_addr);
                    const uint64_t bdesc = umma::smem_desc_sw128(a_addr + A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma::mma2_bf16_w(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (ks | k) != 0, leader);
                    umma::mma2_commit_mc_w(empty_bar(stage), (uint16_t)0x3, leader);
                }
                umma::mma2_commit_mc_w(tfull_bar(acc), (uint16_t)0x3, leader);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue (both CTAs, own 128 TMEM lanes)
        const int wg = (warp - 2) >> 2;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        for (int i = tid - 64; i < p.Cout; i += 256) {
            sc_smem[i] = p.scale ? p.scale[i] : 1.f;
            sc_smem[p.Cout + i] = p.shift ? p.shift[i] : 0.f;
        }
        umma::named_bar_sync(1, 256);
        uint32_t tcount = wg;
        const uint32_t stage = out_stage0 + (uint32_t)wg * epi::STAGE_BYTES;
        const bool issuer = (row == 0);
        const bool has_res = p.has_res != 0;
        uint32_t res_phase = 0;
        const uint32_t lead_tempty = umma::mapa(tempty_bar(wg), 0);
        for (int unit = unit0 + wg * ustride; unit < p.total_units; unit += 2 * ustride, tcount += 2) {
            const int n0 = (unit % p.n_tiles_n) * BN;
            const int m0 = (2 * (unit / p.n_tiles_n) + rank) * TILE_M;
            const bool valid = m0 + row < p.M;
            auto load_res = [&](int s0, uint32_t dst, uint32_t bar) { epi::tma_ld_2d(dst, &tmR, bar, n0 + s0, m0); };
            if (has_res) epi::prefetch_residual(issuer, res_bar(wg), (uint32_t)epi::STAGE_BYTES, stage, load_res);
            umma::mbar_wait(tfull_bar(wg), (tcount >> 1) & 1);
            umma::tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + wg * BN;
            epi::drain_tile<BN, bf16, true>(t_row, lead_tempty, sc_smem + n0, sc_smem + p.Cout + n0, valid, has_res, res_bar(wg),
                                            (uint32_t)epi::STAGE_BYTES, res_phase, p.relu, stage, row, 2u + (uint32_t)wg, issuer,
                                            [&](int s0, uint32_t src) { epi::tma_store_2d(&tmC, src, n0 + s0, m0); }, load_res);
        }
        if (issuer) epi::bulk_wait0();
    }

    umma::tc_fence_before();
    __syncthreads();
    umma::cluster_sync_all();
    if (warp == 1) umma::tmem_dealloc2(tmem_base, 2 * BN);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*,
                                   const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
void* driver_fn(const char* name) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return ptr;
}

}  // namespace

namespace ipsb {

// Launch on clusters of two CTAs.  The tensor maps are built by conv_tma (umma_conv_tma.cu): tmBh boxes HALF slabs
// (64 k x 128 rows).  Returns -1 (no error set) when the device cannot co-schedule a CTA pair.
int conv_pair_launch(const CUtensorMap& tmA, const CUtensorMap& tmBh, const CUtensorMap& tmC, const CUtensorMap& tmR,
                     const float* scale, const float* shift, bool has_res, int P, int Ho, int Wo, int Cout, int kw, int stride, int pad,
                     int relu, int bw, int bh, int bp, int tiles_x, int tiles_y, int n_tiles_n, int m_tiles, int KS, int cblocks,
                     uint32_t a_bytes, cudaStream_t st) {
    PairParams p;
    p.scale = scale; p.shift = shift; p.has_res = has_res ? 1 : 0;
    p.P = P; p.Ho = Ho; p.Wo = Wo; p.Cout = Cout; p.kw = kw; p.stride = stride; p.pad = pad; p.relu = relu;
    p.bw = bw; p.bh = bh; p.bp = bp; p.tiles_x = tiles_x; p.tiles_y = tiles_y; p.n_tiles_n = n_tiles_n;
    p.total_pairs = ((m_tiles + 1) / 2) * n_tiles_n;
    p.KS = KS; p.cblocks = cblocks; p.a_bytes = a_bytes;
    const size_t smem = (size_t)STAGES * STAGE_BYTES + 2 * epi::STAGE_BYTES + 1024 + 8 * (2 * STAGES + 7) + 32 + 8 * (size_t)Cout;
    IPSB_REQUIRE(smem <= 227 * 1024, "conv_pair: %zu bytes of shared memory", smem);
    static size_t configured = 0;
    static int max_clusters = 0;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem; cfg.stream = st; cfg.attrs = attr; cfg.numAttrs = 1;
    if (configured < smem) {
        IPSB_CUDA(cudaFuncSetAttribute(conv_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cfg.gridDim = dim3(2 * (ipsb::sm_count() / 2));
        IPSB_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, conv_pair_kernel, &cfg));
        configured = smem;
    }
    if (max_clusters <= 0) return -1;
    const int cap = max_clusters < ipsb::sm_count() / 2 ? max_clusters : ipsb::sm_count() / 2;
    const int clusters = ipsb::balanced_grid(p.total_pairs, cap);
    cfg.gridDim = dim3(2 * clusters);
    IPSB_CUDA(cudaLaunchKernelEx(&cfg, conv_pair_kernel, tmA, tmBh, tmC, tmR, p));
    IPSB_LAUNCH_CHECK();
    return 0;
}

// Stride-1 3x3 / pad 1 convolution, dense (P,H,W,Cin) bf16 -> dense (P,H,W,Cout) bf16 (+ residual of the output's shape), on
// CTA pairs with im2col-mode TMA.  Returns -1 (no error set) when the shape is not eligible or pairs cannot be scheduled:
// the caller then uses the pixel-box kernels.  Needs an even number of full 128-pixel tiles (P*H*W % 256 == 0).
int conv_pair_im2col(const void* x, const void* w, const float* scale, const float* shift, const void* res, void* y,
                     int64_t P, int H, int W, int Cin, int Cout, int relu, cudaStream_t st) {
    const int64_t M = P * H * W;
    if (Cin % 64 != 0 || Cout % BN != 0 || M % (2 * TILE_M) != 0 || M >= (1ll << 31) || H > 256 || W > 256) return -1;
    static EncodeTiledFn encT = (EncodeTiledFn)driver_fn("cuTensorMapEncodeTiled");
    static EncodeIm2colFn encI = (EncodeIm2colFn)driver_fn("cuTensorMapEncodeIm2col");
    if (!encT || !encI) return -1;
    Im2colParams p;
    p.scale = scale; p.shift = shift; p.has_res = res ? 1 : 0; p.relu = relu;
    p.H = H; p.W = W; p.Cout = Cout; p.cblocks = Cin / BK; p.KS = 9 * p.cblocks;
    p.n_tiles_n = Cout / BN; p.M = (int)M;
    p.total_units = (int)(M / (2 * TILE_M)) * p.n_tiles_n;
    alignas(64) CUtensorMap tmA, tmB, tmC, tmR;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)P};
        cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
        int lower[2] = {-1, -1}, upper[2] = {-1, -1};                 // pad 1; upper corner = pad - (k - 1)
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encI(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, lower, upper, BK, TILE_M, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        IPSB_REQUIRE(r == CUDA_SUCCESS, "conv_pair_im2col: cuTensorMapEncodeIm2col failed with %d", (int)r);
    }
    {
        const int K = 9 * Cin;
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
        cuuint64_t strides[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {BK, BN / 2}, estr[2] = {1, 1};
        CUresult r = encT(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        IPSB_REQUIRE(r == CUDA_SUCCESS, "conv_pair_im2col: cuTensorMapEncodeTiled(weights) failed with %d", (int)r);
    }
    for (int i = 0; i < 2; ++i) {                                     // output and residual: (M, Cout) row-major
        cuuint64_t dims[2] = {(cuuint64_t)Cout, (cuuint64_t)M};
        cuuint64_t strides[1] = {(cuuint64_t)Cout * 2};
        cuuint32_t box[2] = {64, TILE_M}, estr[2] = {1, 1};
        void* base = (i == 0) ? y : const_cast<void*>(res ? res : y);
        CUresult r = encT(i == 0 ? &tmC : &tmR, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        IPSB_REQUIRE(r == CUDA_SUCCESS, "conv_pair_im2col: cuTensorMapEncodeTiled(output) failed with %d", (int)r);
    }
    const size_t smem = (size_t)STAGES * STAGE_BYTES + 2 * epi::STAGE_BYTES + 1024 + 8 * (2 * STAGES + 7) + 32 + 8 * (size_t)Cout;
    IPSB_REQUIRE(smem <= 227 * 1024, "conv_pair_im2col: %zu bytes of shared memory", smem);
    static size_t configured = 0;
    static int max_clusters = 0;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem; cfg.stream = st; cfg.attrs = attr; cfg.numAttrs = 1;
    if (configured < smem) {
        IPSB_CUDA(cudaFuncSetAttribute(conv_pair_im2col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cfg.gridDim = dim3(2 * (ipsb::sm_count() / 2));
        IPSB_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, conv_pair_im2col_kernel, &cfg));
        configured = smem;
    }
    if (max_clusters <= 0) return -1;
    const int cap = max_clusters < ipsb::sm_count() / 2 ? max_clusters : ipsb::sm_count() / 2;
    const int clusters = ipsb::balanced_grid(p.total_units, cap);
    cfg.gridDim = dim3(2 * clusters);
    IPSB_CUDA(cudaLaunchKernelEx(&cfg, conv_pair_im2col_kernel, tmA, tmB, tmC, tmR, p));
    IPSB_LAUNCH_CHECK();
    return 0;
}

}  // namespace ipsb
