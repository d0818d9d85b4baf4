// Probe: im2col-mode TMA (cuTensorMapEncodeIm2col + cp.async.bulk.tensor.4d...im2col) for a 3x3 / pad 1 / stride 1
// convolution on a dense (N, H, W, C) bf16 tensor: does ONE load deliver the tap-(r,s) operand of 128 CONSECUTIVE output
// pixels (crossing image rows and images, zero-filled padding)?  That would let 7x7 feature maps fill 128 of 128 tile
// rows instead of 98.  Hypotheses for the start coordinate of output pixel (q, p, n): (A) tensor space {q - pad, p - pad, n},
// (B) output space {q, p, n}.  Prints which one matches the host gather for several base pixels and taps.
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../ips_b200/csrc/umma.cuh"

constexpr int NI = 6, H = 7, W = 7, C = 64, PIX = 128;

__global__ void probe(const __grid_constant__ CUtensorMap tm, int cw, int ch, int cn, int offw, int offh, __nv_bfloat16* out) {
    extern __shared__ unsigned char raw[];
    const uint32_t s0 = (umma::smem_u32(raw) + 1023u) & ~1023u;
    unsigned char* base = raw + (s0 - umma::smem_u32(raw));
    const uint32_t bar = s0 + PIX * 128;
    const int tid = threadIdx.x;
    for (int i = tid; i < PIX * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x7fc07fc0u;   // NaN fill
    if (tid == 0) { umma::mbar_init(bar, 1); umma::fence_barrier_init(); }
    umma::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
        umma::mbar_expect_tx(bar, PIX * 128);
        const uint16_t ow = (uint16_t)offw, oh = (uint16_t)offh;
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
            ::"r"(s0), "l"(&tm), "r"(bar), "r"(0), "r"(cw), "r"(ch), "r"(cn), "h"(ow), "h"(oh)
            : "memory");
    }
    umma::mbar_wait(bar, 0);
    // de-swizzle (SWIZZLE_128B: 16-byte chunk j of row r sits at chunk j ^ (r & 7))
    for (int i = tid; i < PIX * 8; i += blockDim.x) {
        const int r = i >> 3, j = i & 7;
        *reinterpret_cast<uint4*>(out + r * C + j * 8) = *reinterpret_cast<const uint4*>(base + r * 128 + ((j ^ (r & 7)) << 4));
    }
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*,
                                   const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        printf("cuTensorMapEncodeIm2col not available\n");
        return 1;
    }
    EncodeIm2colFn enc = (EncodeIm2colFn)fp;
    const int n_el = NI * H * W * C;
    __nv_bfloat16* hx = new __nv_bfloat16[n_el];
    for (int n = 0; n < NI; ++n) for (int h = 0; h < H; ++h) for (int w = 0; w < W; ++w) for (int c = 0; c < C; ++c) {
        const int id = 1 + (n * H + h) * W + w;                        // 1 .. 294 (ids < 256 exact; parity of c flips the sign)
        hx[((n * H + h) * W + w) * C + c] = __float2bfloat16((float)((id % 251) * ((c & 1) ? -1 : 1)) + (float)(c / 32) * 0.0f);
    }
    __nv_bfloat16 *dx, *dout;
    cudaMalloc(&dx, n_el * 2); cudaMalloc(&dout, PIX * C * 2);
    cudaMemcpy(dx, hx, n_el * 2, cudaMemcpyHostToDevice);
    alignas(64) CUtensorMap tm;
    cuuint64_t dims[4] = {C, W, H, NI};
    cuuint64_t strides[3] = {C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    int lower[2] = {-1, -1}, upper[2] = {-1, -1};                     // pad 1; upper = pad - (k - 1)
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dx, dims, strides, lower, upper, C, PIX, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode im2col -> %d\n", (int)r);
    if (r != CUDA_SUCCESS) return 1;
    const int smem = PIX * 128 + 2048;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    __nv_bfloat16* ho = new __nv_bfloat16[PIX * C];
    const int bases[] = {0, 30, 100};
    const int taps[][2] = {{0, 0}, {1, 1}, {2, 2}, {2, 0}};
    for (int hyp = 0; hyp < 2; ++hyp) {
        int total_bad = 0;
        for (int m0 : bases) for (auto& t : taps) {
            const int n = m0 / (H * W), p = (m0 / W) % H, qx = m0 % W;
            const int sh = hyp == 0 ? -1 : 0;                           // (A) tensor space, (B) output space
            cudaMemset(dout, 0, PIX * C * 2);
            probe<<<1, 128, smem>>>(tm, qx + sh, p + sh, n, t[1], t[0], dout);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("hyp %c m0=%d tap=(%d,%d): CUDA error %s\n", 'A' + hyp, m0, t[0], t[1], cudaGetErrorString(e)); return 1; }
            cudaMemcpy(ho, dout, PIX * C * 2, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int j = 0; j < PIX; ++j) {
                const int m = m0 + j;
                const int nn = m / (H * W), pp = (m / W) % H, qq = m % W;
                const int ih = pp + t[0] - 1, iw = qq + t[1] - 1;
                for (int c = 0; c < C; ++c) {
                    float ref = 0.f;
                    if (nn < NI && ih >= 0 && ih < H && iw >= 0 && iw < W) ref = __bfloat162float(hx[((nn * H + ih) * W + iw) * C + c]);
                    const float got = __bfloat162float(ho[j * C + c]);
                    if (!(got == ref)) ++bad;
                }
            }
            printf("hyp %c  base pixel %3d  tap (%d,%d): %d mismatches of %d\n", 'A' + hyp, m0, t[0], t[1], bad, PIX * C);
            total_bad += bad;
        }
        printf("hypothesis %c: %s\n", 'A' + hyp, total_bad == 0 ? "MATCHES" : "does not match");
    }
    return 0;
}
