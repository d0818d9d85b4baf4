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

}  // namespace ipsb
