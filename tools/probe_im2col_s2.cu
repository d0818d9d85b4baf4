// Probe: im2col-mode TMA with traversal stride 2 (elementStrides = {1,2,2,1}) for the stride-2 3x3 / pad 1 and 1x1 / pad 0
// convolutions at the head of layers 2-4 (13x13 -> 7x7): one load = one tap of 128 consecutive OUTPUT pixels?
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../ips_b200/csrc/umma.cuh"

constexpr int NI = 6, H = 13, W = 13, C = 64, PIX = 128, HO = 7, WO = 7;

__global__ void probe(const __grid_constant__ CUtensorMap tm, int cw, int ch, int cn, int offw, int offh, __nv_bfloat16* out) {
    extern __shared__ unsigned char raw[];
    const uint32_t s0 = (umma::smem_u32(raw) + 1023u) & ~1023u;
    unsigned char* base = raw + (s0 - umma::smem_u32(raw));
    const uint32_t bar = s0 + PIX * 128;
    const int tid = threadIdx.x;
    for (int i = tid; i < PIX * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x7fc07fc0u;
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
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return 1;
    EncodeIm2colFn enc = (EncodeIm2colFn)fp;
    const int n_el = NI * H * W * C;
    __nv_bfloat16* hx = new __nv_bfloat16[n_el];
    for (int i = 0; i < NI * H * W; ++i) for (int c = 0; c < C; ++c) hx[i * C + c] = __float2bfloat16((float)((1 + i) % 251) * ((c & 1) ? -1.f : 1.f));
    __nv_bfloat16 *dx, *dout;
    cudaMalloc(&dx, n_el * 2); cudaMalloc(&dout, PIX * C * 2);
    cudaMemcpy(dx, hx, n_el * 2, cudaMemcpyHostToDevice);
    const int smem = PIX * 128 + 2048;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    __nv_bfloat16* ho = new __nv_bfloat16[PIX * C];
    struct Cfg { int k, pad; } cfgs[] = {{3, 1}, {1, 0}};
    for (auto& cf : cfgs) {
        alignas(64) CUtensorMap tm;
        cuuint64_t dims[4] = {C, W, H, NI};
        cuuint64_t strides[3] = {C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        int lower[2] = {-cf.pad, -cf.pad}, upper[2] = {cf.pad - (cf.k - 1), cf.pad - (cf.k - 1)};
        cuuint32_t estr[4] = {1, 2, 2, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dx, dims, strides, lower, upper, C, PIX, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("k=%d pad=%d stride 2: encode -> %d\n", cf.k, cf.pad, (int)r);
        if (r != CUDA_SUCCESS) continue;
        const int bases[] = {0, 30, 100};
        int total_bad = 0;
        for (int m0 : bases) for (int t = 0; t < cf.k * cf.k; t += (cf.k == 3 ? 4 : 1)) {
            const int tr = t / cf.k, ts = t % cf.k;
            const int n = m0 / (HO * WO), p = (m0 / WO) % HO, qx = m0 % WO;
            cudaMemset(dout, 0, PIX * C * 2);
            probe<<<1, 128, smem>>>(tm, 2 * qx - cf.pad, 2 * p - cf.pad, n, ts, tr, dout);      // start = input position of the output pixel
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("  m0=%d tap=(%d,%d): CUDA error %s\n", m0, tr, ts, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(ho, dout, PIX * C * 2, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int j = 0; j < PIX; ++j) {
                const int m = m0 + j;
                const int nn = m / (HO * WO), pp = (m / WO) % HO, qq = m % WO;
                const int ih = 2 * pp + tr - cf.pad, iw = 2 * qq + ts - cf.pad;
                for (int c = 0; c < C; ++c) {
                    float ref = 0.f;
                    if (nn < NI && ih >= 0 && ih < H && iw >= 0 && iw < W) ref = __bfloat162float(hx[((nn * H + ih) * W + iw) * C + c]);
                    if (!(__bfloat162float(ho[j * C + c]) == ref)) ++bad;
                }
            }
            printf("  base pixel %3d tap (%d,%d): %d mismatches of %d\n", m0, tr, ts, bad, PIX * C);
            total_bad += bad;
        }
        printf("k=%d stride 2: %s\n", cf.k, total_bad == 0 ? "MATCHES" : "does not match");
    }
    return 0;
}
