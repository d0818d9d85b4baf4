// bf16 implicit-GEMM convolution / dense linear on the 5th-gen tensor cores.
//
//   D[m, n] = sum_k A[m, k] * Wt[n, k]        m = output pixel (patch, oy, ox) flattened
//                                              n = output channel, k = (r, s, c), c fastest
//
// One CTA = one 128 x BN output tile.  Warps 0-3 gather the im2col rows of A and the
// weight rows straight from channels-last HBM into 128B-swizzled K-major shared memory
// (cp.async with zero fill for the halo), one elected thread of warp 4 issues
// tcgen05.mma (M=128, N=BN, K=16) into a TMEM accumulator, tcgen05.commit recycles the
// stages; afterwards warps 0-3 read the accumulator back with tcgen05.ld and apply the
// folded BatchNorm (scale, shift), the residual and ReLU before the channels-last store.
// Two CTAs are resident per SM so one tile's epilogue overlaps the next tile's mainloop.
//
// Replaces conv2d + batch_norm(eval) + add + relu of the reference's truncated ResNet
// (architecture/ips_net.py:17-52) and nn.Linear (ips_net.py:57).
#include "common.cuh"
#include "umma.cuh"
#include "../../include/ips_b200.h"

namespace {

using bf16 = __nv_bfloat16;

constexpr int TILE_M = 128;
constexpr int BK = 64;                    // bf16 elements per stage row = 128 bytes
constexpr int A_STAGE_BYTES = TILE_M * 128;

struct ConvParams {
    const bf16* x;
    const bf16* w;       // (Cout, Kp) K-major
    const float* scale;  // per output channel, may be null
    const float* shift;
    const bf16* res;     // (M, Cout) or null
    void* y;             // (M, Cout) bf16 or fp32
    int64_t M;           // P * Ho * Wo
    int H, W, Cin, Ho, Wo, Cout, kh, kw, stride, pad, relu;
    int Kp;              // padded K (multiple of 64)
};

template <typename T> struct OutStore;
template <> struct OutStore<bf16> {
    // 8 consecutive channels
    static __device__ __forceinline__ void store8(void* base, int64_t off, const float (&v)[8]) {
        __align__(16) __nv_bfloat162 h[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(base) + off) = *reinterpret_cast<const uint4*>(h);
    }
};
template <> struct OutStore<float> {
    static __device__ __forceinline__ void store8(void* base, int64_t off, const float (&v)[8]) {
        float4* d = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off);
        d[0] = make_float4(v[0], v[1], v[2], v[3]);
        d[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
};

// MODE 0: Cin % 64 == 0, one stage = one filter tap x 64 channels
// MODE 1: 7x7 stride-2 stem on 4-channel input, one stage = two filter rows x 8 taps x 4 ch;
//         the 8 taps are columns ix0-1 .. ix0+6 (k = r*32 + (s+1)*4 + c) so every 16-byte chunk is aligned
template <int BN, int STAGES, int MODE, typename OutT>
__global__ void __launch_bounds__(160)
conv_umma_kernel(const ConvParams p) {
    constexpr int B_STAGE_BYTES = BN * 128;
    constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    constexpr int LAG = STAGES - 1;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem0 = (umma::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar0 = smem0 + STAGES * STAGE_BYTES;   // full[S], empty[S], accum, then tmem slot
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
    const uint32_t accum_bar = bar0 + 8u * (2 * STAGES);
    const uint32_t tmem_slot = accum_bar + 8u;
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - umma::smem_u32(smem_raw)));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t m0 = (int64_t)blockIdx.x * TILE_M;
    const int n0 = blockIdx.y * BN;
    const int KS = p.Kp / BK;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            umma::mbar_init(full_bar(s), 128);
            umma::mbar_init(empty_bar(s), 1);
        }
        umma::mbar_init(accum_bar, 1);
        umma::fence_barrier_init();
    }
    if (warp == 4) umma::tmem_alloc(tmem_slot, BN);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp < 4) {
        // ------------------------------------------------------------ producer
        // Lane l owns the 16-byte chunk j = l & 7 of eight tile rows (warp*32 + i*4 + (l >> 3)):
        // eight consecutive lanes fetch one full 128-byte line, four lines per instruction.
        const uint32_t j = (uint32_t)(lane & 7);
        const int rsub = lane >> 3;
        const int cin8 = p.Cin >> 3;                       // 16-byte chunks per input pixel (mode 0)
        uint32_t img_chunk[8];                             // chunk offset of the row's patch image
        int iy0[8], ix0[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t mr = m0 + warp * 32 + i * 4 + rsub;
            if (mr < p.M) {
                const int hw = p.Ho * p.Wo;
                const int64_t pimg = mr / hw;
                const int rem = (int)(mr - pimg * hw);
                const int oy = rem / p.Wo, ox = rem - oy * p.Wo;
                iy0[i] = oy * p.stride - p.pad;
                ix0[i] = ox * p.stride - p.pad - (MODE == 1 ? 1 : 0);   // stem: taps shifted by one, 8th tap first
                img_chunk[i] = (uint32_t)(pimg * (int64_t)p.H * p.W * (MODE == 0 ? cin8 : 1) / (MODE == 0 ? 1 : 2));
            } else {
                iy0[i] = -(1 << 28);                       // never in bounds -> zero fill
                ix0[i] = 0;
                img_chunk[i] = 0;
            }
        }
        const int cblocks = (MODE == 0) ? p.Cin / BK : 1;
        const uint4* xq = reinterpret_cast<const uint4*>(p.x);
        const uint4* wq = reinterpret_cast<const uint4*>(p.w);
        const uint32_t kp8 = (uint32_t)(p.Kp >> 3);

        for (int ks = 0; ks < KS; ++ks) {
            const int stage = ks % STAGES, use = ks / STAGES;
            umma::mbar_wait(empty_bar(stage), (use & 1) ^ 1);
            const uint32_t a_base = smem0 + stage * STAGE_BYTES;
            int r, s_or_q, cb = 0;
            if (MODE == 0) {
                const int tap = ks / cblocks;
                cb = ks - tap * cblocks;
                r = tap / p.kw;
                s_or_q = tap - r * p.kw;
            } else {
                r = 2 * ks + (int)(j >> 2);                // two filter rows per stage
                s_or_q = 2 * (int)(j & 3);                 // pixel pair within the 8 (shifted) taps
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = warp * 32 + i * 4 + rsub;
                const int iy = iy0[i] + r, ix = ix0[i] + s_or_q;
                const bool ok = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W && (MODE == 0 || r < p.kh);
                uint32_t off = 0;
                if (ok) {
                    off = (MODE == 0) ? img_chunk[i] + (uint32_t)(iy * p.W + ix) * (uint32_t)cin8 + (uint32_t)cb * 8u + j
                                      : img_chunk[i] + (uint32_t)((iy * p.W + ix) >> 1);
                }
                umma::cp_async16(a_base + (uint32_t)row * 128u + ((j ^ (uint32_t)(row & 7)) << 4), xq + off, ok ? 16u : 0u);
            }
            // weight rows n0 .. n0+BN of this K slab
            const uint32_t b_base = a_base + A_STAGE_BYTES;
#pragma unroll
            for (int i = 0; i < BN / 16; ++i) {
                const int n = warp * (BN / 4) + i * 4 + rsub;
                const uint32_t off = (uint32_t)(n0 + n) * kp8 + (uint32_t)ks * 8u + j;
                umma::cp_async16(b_base + (uint32_t)n * 128u + ((j ^ (uint32_t)(n & 7)) << 4), wq + off, 16u);
            }
            umma::cp_async_commit();
            if (ks >= LAG) {
                umma::cp_async_wait<LAG>();
                umma::fence_proxy_async();
                umma::mbar_arrive(full_bar((ks - LAG) % STAGES));
            }
        }
        umma::cp_async_wait<0>();
        umma::fence_proxy_async();
        for (int ks = (KS > LAG ? KS - LAG : 0); ks < KS; ++ks) umma::mbar_arrive(full_bar(ks % STAGES));

        // this thread's output row in the epilogue
        const int64_t m = m0 + tid;
        const bool valid = m < p.M;

        // ------------------------------------------------------------ epilogue
        umma::mbar_wait(accum_bar, 0);
        umma::tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            umma::tmem_ld32(t_row + (uint32_t)c0, v);
            umma::tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int n = n0 + c0 + g * 8;
                    float o[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float a = __uint_as_float(v[g * 8 + i]);
                        if (p.scale) a *= __ldg(p.scale + n + i);
                        if (p.shift) a += __ldg(p.shift + n + i);
                        o[i] = a;
                    }
                    if (p.res) {
                        const uint4 rv = *reinterpret_cast<const uint4*>(p.res + m * p.Cout + n);
                        const bf16* rb = reinterpret_cast<const bf16*>(&rv);
#pragma unroll
                        for (int i = 0; i < 8; ++i) o[i] += __bfloat162float(rb[i]);
                    }
                    if (p.relu) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) o[i] = fmaxf(o[i], 0.f);
                    }
                    OutStore<OutT>::store8(p.y, m * p.Cout + n, o);
                }
            }
        }
    } else {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma::idesc_bf16_f32(TILE_M, BN);
            for (int ks = 0; ks < KS; ++ks) {
                const int stage = ks % STAGES, use = ks / STAGES;
                umma::mbar_wait(full_bar(stage), use & 1);
                umma::tc_fence_after();
                const uint32_t a_addr = smem0 + stage * STAGE_BYTES;
                const uint64_t adesc = umma::smem_desc_sw128(a_addr);
                const uint64_t bdesc = umma::smem_desc_sw128(a_addr + A_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                    umma::mma_bf16(tmem_base, adesc + 2u * k, bdesc + 2u * k, idesc, (ks | k) != 0);
                umma::mma_commit(empty_bar(stage));
            }
            umma::mma_commit(accum_bar);
        }
        __syncwarp();
    }

    umma::tc_fence_before();
    __syncthreads();
    if (warp == 4) umma::tmem_dealloc(tmem_base, BN);
}

template <int BN, int STAGES, int MODE, typename OutT>
int launch(const ConvParams& p, cudaStream_t st) {
    constexpr size_t smem = (size_t)STAGES * (A_STAGE_BYTES + BN * 128) + 1024 /*align*/ + 8 * (2 * STAGES + 1) + 16;
    auto kern = conv_umma_kernel<BN, STAGES, MODE, OutT>;
    IPSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ipsb::ceil_div(p.M, TILE_M), (unsigned)(p.Cout / BN));
    kern<<<grid, 160, smem, st>>>(p);
    IPSB_LAUNCH_CHECK();
    return 0;
}

template <int MODE, typename OutT>
int dispatch_bn(const ConvParams& p, cudaStream_t st) {
    if (p.Cout % 128 == 0) return launch<128, 3, MODE, OutT>(p, st);
    if (p.Cout % 64 == 0) return launch<64, 4, MODE, OutT>(p, st);
    return ipsb::fail("umma: Cout=%d must be a multiple of 64", p.Cout);
}

}  // namespace

extern "C" {

int ipsb_conv_bf16_umma(const void* x, const void* w, const float* scale, const float* shift,
                        const void* res, void* y, int64_t P, int H, int W, int Cin, int Cout,
                        int kh, int kw, int stride, int pad, int relu, int mode, void* stream) {
    ConvParams p;
    p.x = (const bf16*)x; p.w = (const bf16*)w; p.scale = scale; p.shift = shift;
    p.res = (const bf16*)res; p.y = y;
    p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.kh = kh; p.kw = kw; p.stride = stride; p.pad = pad; p.relu = relu;
    p.Ho = (H + 2 * pad - kh) / stride + 1;
    p.Wo = (W + 2 * pad - kw) / stride + 1;
    p.M = P * p.Ho * p.Wo;
    IPSB_REQUIRE(P > 0 && p.Ho > 0 && p.Wo > 0, "conv_umma: bad geometry");
    if (mode == 0)      // TMA-fed path: the M tile is a box of output pixels
        return ipsb::conv_tma(x, w, scale, shift, res, y, P, H, W, Cin, Cout, kh, kw, stride, pad, relu, false,
                              (cudaStream_t)stream);
    if (mode == 4) {    // space-to-depth stem: x is the s2d frame of ipsb_stage_patches_s2d for HxW images, y the wide row order
        IPSB_REQUIRE(Cin == 16 && kh == 7 && kw == 7 && stride == 2, "conv_umma mode 4 is the 7x7/2 stem on the s2d frame");
        return ipsb::conv_stem_s2d(x, w, scale, shift, y, P, H, W, Cout, relu, (cudaStream_t)stream);
    }
    if (mode == 3) {    // TMA-fed stem: x is the zero-bordered (P, H, W, 4) frame, image = (H-6) x (W-6)
        IPSB_REQUIRE(Cin == 4 && kh == 7 && kw == 7 && stride == 2, "conv_umma mode 3 is the 7x7/2 stem");
        return ipsb::conv_stem_tma(x, w, scale, shift, y, P, H - 6, W - 6, Cout, relu, (cudaStream_t)stream);
    }
    if (mode == 2) {    // cp.async gather path (any geometry with Cin % 64 == 0); kept as a cross-check
        IPSB_REQUIRE(Cin % 64 == 0, "conv_umma mode 2: Cin=%d must be a multiple of 64", Cin);
        p.Kp = kh * kw * Cin;
        return dispatch_bn<0, bf16>(p, (cudaStream_t)stream);
    }
    if (mode == 1) {
        IPSB_REQUIRE(Cin == 4 && kh == 7 && kw == 7 && stride == 2 && pad == 3 && W % 2 == 0,
                     "conv_umma mode 1 is the 7x7/2 pad-3 stem on 4-channel input of even width");
        p.Kp = 256;
        return dispatch_bn<1, bf16>(p, (cudaStream_t)stream);
    }
    return ipsb::fail("conv_umma: unknown mode %d", mode);
}

int ipsb_conv_bf16_pf(const void* x, const void* w, const float* scale, const float* shift, const void* res, void* y,
                      int64_t P, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad, int relu,
                      int in_pf, int out_pf, void* stream) {
    if (in_pf && out_pf && kh == 3 && kw == 3 && stride == 1 && pad == 1 && (Cout == 64 || Cout == 128))
        return ipsb::conv3x3_halo(x, w, scale, shift, res, y, P, H, W, Cin, Cout, relu, (cudaStream_t)stream);
    return ipsb::conv_tma(x, w, scale, shift, res, y, P, H, W, Cin, Cout, kh, kw, stride, pad, relu, false,
                          (cudaStream_t)stream, in_pf != 0, out_pf != 0);
}

// fp32-output convolution on dense bf16 activations: one of the three partial products of the bf16x3 precision (split3.cu)
int ipsb_conv_bf16_f32out(const void* x, const void* w, const float* scale, const float* shift, float* y, int64_t P, int H, int W,
                          int Cin, int Cout, int kh, int kw, int stride, int pad, int mode, void* stream) {
    IPSB_REQUIRE(x && w && y && P > 0, "conv_bf16_f32out: null argument");
    if (mode == 3) {    // TMA-fed stem on the zero-bordered (P, H, W, 4) frame, image = (H-6) x (W-6)
        IPSB_REQUIRE(Cin == 4 && kh == 7 && kw == 7 && stride == 2, "conv_bf16_f32out mode 3 is the 7x7/2 stem");
        return ipsb::conv_stem_tma(x, w, scale, shift, y, P, H - 6, W - 6, Cout, 0, (cudaStream_t)stream, true);
    }
    IPSB_REQUIRE(mode == 0, "conv_bf16_f32out: mode %d", mode);
    return ipsb::conv_tma(x, w, scale, shift, nullptr, y, P, H, W, Cin, Cout, kh, kw, stride, pad, 0, true, (cudaStream_t)stream);
}

int ipsb_linear_bf16_umma(const void* a, const void* w, const float* scale, const float* shift,
                          float* y, int64_t M, int N, int K, int relu, void* stream) {
    IPSB_REQUIRE(M > 0 && K % 64 == 0 && N % 64 == 0, "linear_umma: K=%d, N=%d must be multiples of 64", K, N);
    return ipsb::conv_tma(a, w, scale, shift, nullptr, y, M, 1, 1, K, N, 1, 1, 1, 0, relu, true, (cudaStream_t)stream);
}

}  // extern "C"
