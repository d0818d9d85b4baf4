// 7x7 stride-2 stem + folded BatchNorm + ReLU + 3x3 stride-2 max-pool in ONE kernel (tcgen05 / TMEM / TMA, sm_100a).
//
// Replaces conv1 -> bn1 -> relu -> maxpool of the truncated ResNet (architecture/ips_net.py:17-39 of the reference)
// for the no-grad selection pass.  The unfused chain wrote the 64-channel stem output (the largest activation of the
// encoder: 360 MB per 1024 traffic patches) to HBM and read it back for the pool; here it never leaves the SM.
//
// Input is the space-to-depth frame of ipsb_stage_patches_s2d (16 bf16 per frame pixel, row pitch Wp = W/2 + 3,
// Sp = (H/2 + 3) * Wp frame pixels per patch), on which the stem is a 4x4 stride-1 convolution whose 16 taps are
// row-shifted views of one block of frame rows (SWIZZLE_32B, see umma_conv_halo.cu).  Stem output pixel (y, x) of a
// patch sits at flat frame index y*Wp + x, so 128 consecutive frame rows form one M tile.
//
// Work unit = GROUP: R pooled rows of one patch.  They need stem rows 2*pr0-1 .. 2*(pr0+R-1)+1, i.e. the (2R+1)*Wp
// consecutive flat rows starting one image row above -> NT tiles of 128 rows.  Per group:
//   producer warp    TMA-loads each tile's block of frame rows (128 + 3*Wp + 3 rows of 32 bytes)
//   MMA warp         16 tcgen05.mma (M=128, N=64, K=16) per tile into one of eight TMEM accumulators
//   4 epilogue WGs   tile t of the CTA goes to warpgroup t%4: tcgen05.ld -> BN + ReLU -> bf16 -> the group's staging
//                    buffer in shared memory (NT x 16 KB, 128-byte rows, XOR-swizzled chunks); after a barrier all
//                    512 threads max-pool out of shared memory and write the padded-flat layer-1 input with
//                    16-byte stores; a second barrier frees the buffer.
// Garbage rows (the 3 pad columns of the frame pitch, the row above the first / below the last image row, frame rows
// of the neighbouring patch) are computed but never read by the pool, which only visits 0 <= y < Ho, 0 <= x < Wo.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "umma.cuh"
#include "pf.cuh"
#include "epilogue.cuh"
#include "../../include/ips_b200.h"

namespace {

using bf16 = __nv_bfloat16;
constexpr int TILE_M = 128;
constexpr int BN = 64;
constexpr int NWG = 4;                     // epilogue warpgroups
constexpr int NACC = 8;                    // TMEM accumulator stages (all 512 columns): the MMA warp runs a group ahead of the pool
constexpr int B_SLAB_BYTES = BN * 128;     // (64, 64) bf16 slab of the (64, 256) weights
constexpr int THREADS = 64 + 128 * NWG;

struct StemPoolParams {
    const float* scale;
    const float* shift;
    bf16* y;                 // padded-flat (rows, 64) layer-1 input
    int P, Ho, Wo, Wp, Sp;   // stem output size, frame pitch / rows per patch
    int Hq, Wq;              // pooled size
    int oG0, oWp, oSp;       // padded-flat geometry of y
    int R, NT, GPP;          // pooled rows per group, tiles per group, groups per patch
    int total_groups;
    int a_rows, n_boxes;     // TMA boxes of one tile's block of frame rows
    uint32_t a_slot_bytes;
    int relu;
    uint32_t per_row_magic;  // ceil(2^32 / (Wq * 8)): idx / per_row by multiply-high
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

// K-major operand with 32-byte rows (16 bf16): 8-row atoms of 256 bytes, SWIZZLE_32B
__device__ __forceinline__ uint64_t desc_sw32(uint32_t addr) {
    const uint32_t lo = ((addr >> 4) & 0x3fffu) | (1u << 16);
    const uint32_t hi = (256u >> 4) | (1u << 14) | (6u << 29);
    return ((uint64_t)hi << 32) | lo;
}

__device__ __forceinline__ uint32_t bmax2(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}

template <int SA>
__global__ void __launch_bounds__(THREADS, 1)
stem_pool_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const StemPoolParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem0 = (umma::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b0 = smem0 + SA * p.a_slot_bytes;                 // four resident weight slabs
    const uint32_t stage0 = b0 + 4u * B_SLAB_BYTES;                  // NT x 16 KB stem rows of the current group
    const uint32_t bar0 = stage0 + (uint32_t)p.NT * epi::STAGE_BYTES;
    auto a_full = [&](int s) { return bar0 + 8u * s; };
    auto a_empty = [&](int s) { return bar0 + 8u * (SA + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * SA + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * SA + NACC + a); };
    const uint32_t resb_bar = bar0 + 8u * (2 * SA + 2 * NACC);
    const uint32_t tmem_slot = resb_bar + 8u;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - umma::smem_u32(smem_raw)));
    const uint32_t sc_addr = (tmem_slot + 4u + 15u) & ~15u;
    float* sc_smem = reinterpret_cast<float*>(smem_raw + (sc_addr - umma::smem_u32(smem_raw)));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < SA; ++s) { umma::mbar_init(a_full(s), 1); umma::mbar_init(a_empty(s), 1); }
        for (int a = 0; a < NACC; ++a) { umma::mbar_init(tfull_bar(a), 1); umma::mbar_init(tempty_bar(a), 128); }
        umma::mbar_init(resb_bar, 1);
        umma::fence_barrier_init();
    }
    if (warp == 1) umma::tmem_alloc(tmem_slot, NACC * BN);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            umma::mbar_expect_tx(resb_bar, 4u * B_SLAB_BYTES);
            for (int ks = 0; ks < 4; ++ks) tma_load_2d(b0 + ks * B_SLAB_BYTES, &tmB, resb_bar, ks * 64, 0);
            uint32_t ia = 0;
            for (int grp = blockIdx.x; grp < p.total_groups; grp += gridDim.x) {
                const int pp = grp / p.GPP, gi = grp - pp * p.GPP;
                const int f0 = pp * p.Sp + (2 * gi * p.R - 1) * p.Wp;          // first frame row of the group (may be < 0: zero fill)
                for (int j = 0; j < p.NT; ++j, ++ia) {
                    const int sa = ia % SA;
                    umma::mbar_wait(a_empty(sa), ((ia / SA) & 1) ^ 1);
                    umma::mbar_expect_tx(a_full(sa), (uint32_t)(p.a_rows * p.n_boxes) * 32u);
                    for (int bx = 0; bx < p.n_boxes; ++bx)
                        tma_load_2d(smem0 + sa * p.a_slot_bytes + (uint32_t)(bx * p.a_rows) * 32u, &tmA, a_full(sa), 0,
                                    f0 + j * TILE_M + bx * p.a_rows);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (warp-uniform, elected lane issues)
        const uint32_t leader = umma::elect_one();
        constexpr uint32_t idesc = umma::idesc_bf16_f32(TILE_M, BN);
        umma::mbar_wait(resb_bar, 0);
        uint32_t tc = 0;
        for (int grp = blockIdx.x; grp < p.total_groups; grp += gridDim.x) {
            for (int j = 0; j < p.NT; ++j, ++tc) {
                const uint32_t acc = tc % NACC;
                umma::mbar_wait(tempty_bar(acc), ((tc / NACC) & 1) ^ 1);
                const int sa = tc % SA;
                umma::mbar_wait(a_full(sa), (tc / SA) & 1);
                umma::tc_fence_after();
                const uint32_t a_base = smem0 + sa * p.a_slot_bytes;
                const uint32_t d_tmem = tmem_base + acc * BN;
#pragma unroll
                for (int tap = 0; tap < 16; ++tap) {     // tap (a, b) of the 4x4 window: rows shifted by a*Wp + b
                    const int ta = tap >> 2, tb = tap & 3;
                    const uint64_t adesc = desc_sw32(a_base + (uint32_t)(ta * p.Wp + tb) * 32u);
                    const uint64_t bdesc = umma::smem_desc_sw128(b0 + ta * B_SLAB_BYTES) + (uint64_t)(2 * tb);
                    umma::mma_bf16_w(d_tmem, adesc, bdesc, idesc, tap != 0, leader);
                }
                umma::mma_commit_w(a_empty(sa), leader);
                umma::mma_commit_w(tfull_bar(acc), leader);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue + pool
        const int wg = (warp - 2) >> 2;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int et = tid - 64;                                      // 0 .. 511
        for (int i = et; i < BN; i += 128 * NWG) {
            sc_smem[i] = p.scale ? p.scale[i] : 1.f;
            sc_smem[BN + i] = p.shift ? p.shift[i] : 0.f;
        }
        umma::named_bar_sync(1, 128 * NWG);
        uint32_t tc0 = 0;                                             // CTA-local index of the group's first tile
        const int per_row = p.Wq * 8;                                 // 16-byte chunk tasks per pooled row
        for (int grp = blockIdx.x; grp < p.total_groups; grp += gridDim.x, tc0 += p.NT) {
            const int pp = grp / p.GPP, gi = grp - pp * p.GPP;
            const int pr0 = gi * p.R;
            for (int j = 0; j < p.NT; ++j) {
                const uint32_t tc = tc0 + j;
                if ((int)(tc % NWG) != wg) continue;
                const uint32_t acc = tc % NACC;
                umma::mbar_wait(tfull_bar(acc), (tc / NACC) & 1);
                umma::tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
                const uint32_t row_smem = stage0 + (uint32_t)(j * TILE_M + row) * 128u;
#pragma unroll
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t v[32];
                    umma::tmem_ld32(t_row + (uint32_t)c0, v);
                    umma::tmem_ld_wait();
                    if (c0 + 32 >= BN) {                              // accumulator fully read: hand it back to the MMA warp
                        umma::tc_fence_before();
                        umma::mbar_arrive(tempty_bar(acc));
                    }
                    epi::stage32<bf16>(v, sc_smem + c0, sc_smem + BN + c0, true, false, p.relu, row_smem, row, c0 / 8);
                }
            }
            umma::named_bar_sync(1, 128 * NWG);                       // the group's stem rows are staged
            const int rg = min(p.R, p.Hq - pr0);
            for (int idx = et; idx < rg * per_row; idx += 128 * NWG) {
                const int pr = (int)(((uint64_t)(uint32_t)idx * p.per_row_magic) >> 32);      // idx / per_row
                const int rem = idx - pr * per_row;
                const int px = rem >> 3, ch = rem & 7;
                uint32_t m0 = 0xff80ff80u, m1 = m0, m2 = m0, m3 = m0;  // -inf
                // window rows / columns outside the image are clamped onto the border pixel: a duplicate does not change a max
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    const int yy = min(max(2 * (pr0 + pr) - 1 + dy, 0), p.Ho - 1);
                    const int lrow = (yy - (2 * pr0 - 1)) * p.Wp;
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const int xx = min(max(2 * px - 1 + dx, 0), p.Wo - 1);
                        const uint32_t l = (uint32_t)(lrow + xx);
                        uint32_t a, b, c, d;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
                                     : "r"(stage0 + l * 128u + (((uint32_t)ch ^ (l & 7u)) << 4)));
                        m0 = bmax2(m0, a); m1 = bmax2(m1, b); m2 = bmax2(m2, c); m3 = bmax2(m3, d);
                    }
                }
                const int64_t orow = (int64_t)p.oG0 + (int64_t)pp * p.oSp + (int64_t)(pr0 + pr) * p.oWp + px;
                *reinterpret_cast<uint4*>(p.y + orow * BN + ch * 8) = make_uint4(m0, m1, m2, m3);
            }
            umma::named_bar_sync(1, 128 * NWG);                       // staging buffer free for the next group
        }
    }

    umma::tc_fence_before();
    __syncthreads();
    if (warp == 1) umma::tmem_dealloc(tmem_base, NACC * BN);
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

// frame: s2d frame of ipsb_stage_patches_s2d for P patches of HxW; w: (64, 256) bf16 (mode-4 packing of
// ipsb_conv_bf16_umma); y: padded-flat (ipsb_pf_rows(P, Hq, Wq), 64) bf16 whose pad rows are already zero.
int ipsb_stem_pool_s2d(const void* frame, const void* w, const float* scale, const float* shift, void* y, int64_t P, int H, int W,
                       int relu, void* stream) {
    IPSB_REQUIRE(P > 0 && H % 2 == 0 && W % 2 == 0 && H >= 4 && W >= 4, "stem_pool: needs even H, W >= 4");
    EncodeTiledFn enc = encode_fn();
    IPSB_REQUIRE(enc != nullptr, "stem_pool: cuTensorMapEncodeTiled not available from the driver");
    StemPoolParams p;
    p.scale = scale; p.shift = shift; p.y = (bf16*)y; p.relu = relu;
    p.P = (int)P; p.Ho = H / 2; p.Wo = W / 2; p.Wp = p.Wo + 3; p.Sp = (p.Ho + 3) * p.Wp;
    p.Hq = (p.Ho - 1) / 2 + 1; p.Wq = (p.Wo - 1) / 2 + 1;
    const pf::Geo g = pf::make(P, p.Hq, p.Wq);
    p.oG0 = g.G0; p.oWp = g.Wp; p.oSp = g.Sp;
    const int64_t frame_rows = P * (int64_t)p.Sp;
    IPSB_REQUIRE(frame_rows + 8 * TILE_M < (1ll << 31), "stem_pool: too many frame rows");
    p.n_boxes = 2;
    p.a_rows = ((TILE_M + 3 * p.Wp + 3 + 1) / 2 + 7) / 8 * 8;       // two boxes cover 128 + 3*Wp + 3 rows
    IPSB_REQUIRE(p.a_rows <= 256, "stem_pool: width %d too large", W);
    p.a_slot_bytes = (uint32_t)((p.a_rows * 2 * 32 + 1023) / 1024 * 1024);
    constexpr int SA = 4;
    // pooled rows per group: fewest tiles per patch with the staging buffer within shared memory
    const size_t fixed = (size_t)SA * p.a_slot_bytes + 4 * B_SLAB_BYTES + 2048 + 8 * (2 * SA + 2 * NACC + 2) + 8 * BN;
    const int nt_max = (int)((227 * 1024 - fixed) / epi::STAGE_BYTES);
    int best_cost = 1 << 30;
    p.R = 0;
    for (int R = 1; R <= p.Hq; ++R) {
        const int nt = ((2 * R + 1) * p.Wp + TILE_M - 1) / TILE_M;
        if (nt > nt_max) break;
        const int gpp = (p.Hq + R - 1) / R;
        if (gpp * nt <= best_cost) { best_cost = gpp * nt; p.R = R; p.NT = nt; p.GPP = gpp; }
    }
    IPSB_REQUIRE(p.R > 0, "stem_pool: width %d does not fit the staging buffer", W);
    p.total_groups = (int)(P * p.GPP);
    p.per_row_magic = (uint32_t)(((1ull << 32) + (uint64_t)(p.Wq * 8) - 1) / (uint64_t)(p.Wq * 8));
    alignas(64) CUtensorMap tmA, tmB;
    {
        cuuint64_t dims[2] = {16, (cuuint64_t)frame_rows};
        cuuint64_t strides[1] = {32};
        cuuint32_t box[2] = {16, (cuuint32_t)p.a_rows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(frame), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        IPSB_REQUIRE(r == CUDA_SUCCESS, "stem_pool: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
    }
    {
        cuuint64_t dims[2] = {256, (cuuint64_t)BN};
        cuuint64_t strides[1] = {512};
        cuuint32_t box[2] = {64, 64};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        IPSB_REQUIRE(r == CUDA_SUCCESS, "stem_pool: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
    }
    const size_t smem = fixed + (size_t)p.NT * epi::STAGE_BYTES;
    auto kern = stem_pool_kernel<SA>;
    static size_t configured = 0;
    if (configured < smem) {
        IPSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const int grid = ipsb::balanced_grid(p.total_groups, ipsb::sm_count());
    kern<<<grid, THREADS, smem, (cudaStream_t)stream>>>(tmA, tmB, p);
    IPSB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
