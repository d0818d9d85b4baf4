// TMA-staged patch -> space-to-depth frame (north_star (1): "TMA-staged unfold/gather").
//
// The reference unfolds patches on the CPU and hands (rows, C, H, W) fp32 tensors to the encoder (ips_net.py:204-209);
// the stem kernel wants, per patch, a frame of (H/2+3) x (W/2+3) pixels of 16 bf16 with
//     frame(Y', X')[(dy*2+dx)*4 + c] = in(c, 2Y'+dy-4, 2X'+dx-4)        (zero outside the patch, 4th channel zero)
// (DESIGN.md section 2).  stage_s2d_kernel (layout.cu) does this with 16-byte loads per thread and reaches 0.54-0.57 of
// the HBM peak.  Here the copy engine does the addressing:
//   * load:  ONE 4-D tensor-map box per work item (a band of YB frame rows of one patch): x in [-4, W+2+pad), y in
//     [2Y0-4, 2Y0-4+2YB), all C channels, one patch -- out-of-bounds elements (the conv's zero padding on all four sides)
//     are zero-filled by the TMA unit, so the kernel has no boundary code at all; a patch index array (gathered rows)
//     is just another coordinate;
//   * the 256 threads turn the fp32 box into frame pixels (6 x 64-bit shared loads, 2 x 128-bit shared stores each);
//   * store: the band of the frame is contiguous in HBM -> one cp.async.bulk shared -> global per item.
// Persistent CTAs, a ring of ST load stages (mbarrier complete_tx) and two output buffers (bulk-group waits): 60-80 KB of
// loads in flight per SM without a single register holding data.
#include <cuda.h>
#include "common.cuh"
#include "umma.cuh"
#include "../../include/ips_b200.h"

namespace {

using bf16 = __nv_bfloat16;
constexpr int YB = 8;                  // frame rows per work item
constexpr int ST = 4;                  // load stages
constexpr int NT = 256;

struct StageParams {
    const int64_t* row_idx;            // gathered source rows or null
    int64_t first_row, n_rows;
    int C, H, W, Ys, Wp, box_w, bands;
    bf16* dst;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(NT) stage_s2d_tma_kernel(const __grid_constant__ CUtensorMap tm, const StageParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw0 = umma::smem_u32(smem_raw);
    const uint32_t base = (raw0 + 127u) & ~127u;
    const uint32_t in_bytes = (uint32_t)(p.C * 2 * YB * p.box_w * 4);
    const uint32_t in_pitch = (in_bytes + 127u) & ~127u;
    const uint32_t out_bytes = (uint32_t)(YB * p.Wp * 32);
    const uint32_t out_pitch = (out_bytes + 127u) & ~127u;
    const uint32_t in0 = base, out0 = in0 + ST * in_pitch, bar0 = out0 + 2 * out_pitch;
    unsigned char* gen0 = smem_raw + (base - raw0);
    const int tid = threadIdx.x;
    const int64_t items = p.n_rows * p.bands;

    if (tid == 0) {
        for (int s = 0; s < ST; ++s) umma::mbar_init(bar0 + 8u * s, 1);
        umma::fence_barrier_init();
    }
    __syncthreads();

    auto issue = [&](int64_t item, int stage) {            // thread 0 only
        const int64_t r = item / p.bands;
        const int band = (int)(item - r * p.bands);
        const int64_t srow = p.row_idx ? p.row_idx[r] : p.first_row + r;
        umma::mbar_expect_tx(bar0 + 8u * stage, in_bytes);
        tma_load_4d(in0 + stage * in_pitch, &tm, bar0 + 8u * stage, -4, 2 * band * YB - 4, 0, (int)srow);
    };
    const int64_t first = blockIdx.x, step = gridDim.x;
    if (tid == 0)
        for (int s = 0; s < ST; ++s)
            if (first + s * step < items) issue(first + s * step, s);

    int k = 0;
    for (int64_t item = first; item < items; item += step, ++k) {
        const int stage = k % ST, ob = k & 1;
        umma::mbar_wait(bar0 + 8u * stage, (k / ST) & 1);
        const int64_t r = item / p.bands;
        const int band = (int)(item - r * p.bands);
        const int rows_here = min(YB, p.Ys - band * YB);
        const float* tile = reinterpret_cast<const float*>(gen0 + stage * in_pitch);
        unsigned char* outb = gen0 + ST * in_pitch + ob * out_pitch;
        const int npix = rows_here * p.Wp;
        for (int i = tid; i < npix; i += NT) {
            const int yl = i / p.Wp, xp = i - yl * p.Wp;
            __align__(16) bf16 o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = __float2bfloat16_rn(0.f);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (c >= p.C) break;
#pragma unroll
                for (int dy = 0; dy < 2; ++dy) {
                    const float2 v = *reinterpret_cast<const float2*>(tile + ((size_t)c * 2 * YB + 2 * yl + dy) * p.box_w + 2 * xp);
                    o[(dy * 2 + 0) * 4 + c] = __float2bfloat16_rn(v.x);
                    o[(dy * 2 + 1) * 4 + c] = __float2bfloat16_rn(v.y);
                }
            }
            uint4* d = reinterpret_cast<uint4*>(outb + (size_t)i * 32);
            d[0] = *reinterpret_cast<const uint4*>(&o[0]);
            d[1] = *reinterpret_cast<const uint4*>(&o[8]);
        }
        umma::fence_proxy_async();                          // frame band in shared memory -> visible to the bulk store
        if (tid == 0) bulk_wait_read<0>();                  // the other output buffer's previous store has left shared memory
        __syncthreads();
        if (tid == 0) {
            bf16* g = p.dst + ((r * p.Ys + (int64_t)band * YB) * p.Wp) * 16;
            bulk_store(g, out0 + ob * out_pitch, (uint32_t)npix * 32u);
            bulk_commit();
            const int64_t nxt = item + (int64_t)ST * step;   // the stage just consumed is free (all threads passed the barrier)
            if (nxt < items) issue(nxt, stage);
        }
    }
    if (tid == 0) bulk_wait_read<0>();                      // shared memory must outlive the stores
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

static size_t stage_smem_bytes(int C, int W) {
    const int box_w = (W + 6 + 3) / 4 * 4, Wp = W / 2 + 3;
    const size_t in_pitch = ((size_t)C * 2 * YB * box_w * 4 + 127) / 128 * 128;
    const size_t out_pitch = ((size_t)YB * Wp * 32 + 127) / 128 * 128;
    return 128 + ST * in_pitch + 2 * out_pitch + 8 * ST + 64;
}

// 1 when the TMA form applies: 16-byte aligned source, row pitch and channel / patch pitches multiples of 16 bytes, box <= 256
// elements wide, the stage ring inside shared memory
int ipsb_stage_tma_ok(const float* src, int C, int H, int W) {
    return src != nullptr && ((uintptr_t)src % 16) == 0 && C >= 1 && C <= 4 && H % 2 == 0 && W % 4 == 0 && W + 8 <= 256 &&
           H >= 2 && stage_smem_bytes(C, W) <= 227 * 1024 && encode_fn() != nullptr;
}

// Same contract as ipsb_stage_patches_s2d (layout.cu).  max_row: an upper bound (exclusive) of the source rows that may be
// addressed (first_row + n_rows, or the length of the tensor row_idx points into); 0 = unknown.
int ipsb_stage_patches_s2d_tma(const float* src, const int64_t* row_idx, int64_t first_row, int64_t n_rows, int C, int H, int W,
                               void* dst, void* stream) {
    IPSB_REQUIRE(n_rows > 0 && ipsb_stage_tma_ok(src, C, H, W), "stage_s2d_tma: needs a 16-byte aligned source, C <= 4, even H, W %% 4 == 0");
    IPSB_REQUIRE(((uintptr_t)dst % 16) == 0, "stage_s2d_tma: misaligned destination");
    StageParams p;
    p.row_idx = row_idx; p.first_row = first_row; p.n_rows = n_rows;
    p.C = C; p.H = H; p.W = W; p.Ys = H / 2 + 3; p.Wp = W / 2 + 3;
    p.box_w = (W + 6 + 3) / 4 * 4;
    p.bands = (p.Ys + YB - 1) / YB;
    p.dst = (bf16*)dst;
    alignas(64) CUtensorMap tm;
    {
        // (x, y, c, patch); the patch extent is a bound for the address check only (rows may be gathered from anywhere below it)
        cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)(1u << 30)};
        cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
        cuuint32_t box[4] = {(cuuint32_t)p.box_w, (cuuint32_t)(2 * YB), (cuuint32_t)C, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        const CUresult r = encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(src), dims, strides, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        IPSB_REQUIRE(r == CUDA_SUCCESS, "stage_s2d_tma: cuTensorMapEncodeTiled failed with %d", (int)r);
    }
    const size_t smem = stage_smem_bytes(C, W);
    static size_t attr_set = 0;
    if (smem > attr_set) {
        IPSB_CUDA(cudaFuncSetAttribute(stage_s2d_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = smem;
    }
    const int64_t items = n_rows * p.bands;
    const int per_sm = (int)((227 * 1024) / (smem + 1024));
    int64_t grid = (int64_t)ipsb::sm_count() * (per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm));
    if (grid > items) grid = items;
    stage_s2d_tma_kernel<<<(unsigned)grid, NT, smem, (cudaStream_t)stream>>>(tm, p);
    IPSB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
