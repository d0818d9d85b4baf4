// Probe: does tcgen05.mma accept a SWIZZLE_128B K-major A operand whose start address is shifted by a
// number of 128-byte rows that is not a multiple of 8 (i.e. not 1024-byte aligned)?  Variants: descriptor
// base_offset field = 0, or = (start_addr >> 7) & 7.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o probe_shift probe_shift.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../ips_b200/csrc/umma.cuh"

constexpr int ROWS = 256, N = 64, K = 64;

__global__ void probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int off, int use_base_offset, int lbo) {
    extern __shared__ unsigned char raw[];
    const uint32_t s0 = (umma::smem_u32(raw) + 1023u) & ~1023u;
    unsigned char* base = raw + (s0 - umma::smem_u32(raw));
    const uint32_t a_addr = s0, b_addr = s0 + ROWS * 128, bar = b_addr + N * 128, slot = bar + 8;
    const int tid = threadIdx.x;
    // fill swizzled: chunk j of row r at r*128 + ((j ^ (r & 7)) << 4)
    for (int i = tid; i < ROWS * 8; i += blockDim.x) {
        const int r = i >> 3, j = i & 7;
        *reinterpret_cast<uint4*>(base + r * 128 + ((j ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + r * K + j * 8);
    }
    for (int i = tid; i < N * 8; i += blockDim.x) {
        const int r = i >> 3, j = i & 7;
        *reinterpret_cast<uint4*>(base + ROWS * 128 + r * 128 + ((j ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * K + j * 8);
    }
    if (tid == 0) { umma::mbar_init(bar, 1); umma::fence_barrier_init(); }
    if (tid < 32) umma::tmem_alloc(slot, 64);
    umma::fence_proxy_async();
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(base + (slot - s0));
    if (tid == 0) {
        const uint32_t start = a_addr + off * 128;
        uint64_t ad = umma::smem_desc_sw128(start);
        if (use_base_offset) ad |= (uint64_t)((start >> 7) & 7) << 49;
        if (lbo) ad = (ad & ~(0x3fffull << 16));   // LBO field = 0 variant
        const uint64_t bd = umma::smem_desc_sw128(b_addr);
        for (int k = 0; k < 4; ++k) umma::mma_bf16(tmem, ad + 2u * k, bd + 2u * k, umma::idesc_bf16_f32(128, N), k != 0);
        umma::mma_commit(bar);
    }
    __syncthreads();
    umma::mbar_wait(bar, 0);
    umma::tc_fence_after();
    if (tid < 128) {
        const int warp = tid >> 5;
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
            umma::tmem_ld_wait();
            for (int i = 0; i < 32; ++i) D[tid * N + c0 + i] = __uint_as_float(v[i]);
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (tid < 32) umma::tmem_dealloc(tmem, 64);
}

int main() {
    __nv_bfloat16 *hA = new __nv_bfloat16[ROWS * K], *hB = new __nv_bfloat16[N * K];
    for (int r = 0; r < ROWS; ++r) for (int k = 0; k < K; ++k) hA[r * K + k] = __float2bfloat16((float)((r * 7 + k * 3) % 13 - 6));
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) hB[n * K + k] = __float2bfloat16((float)((n * 5 + k) % 7 - 3));
    __nv_bfloat16 *dA, *dB; float* dD;
    cudaMalloc(&dA, ROWS * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dD, 128 * N * 4);
    cudaMemcpy(dA, hA, ROWS * K * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, N * K * 2, cudaMemcpyHostToDevice);
    const int smem = ROWS * 128 + N * 128 + 2048;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float* hD = new float[128 * N];
    const int offs[] = {0, 8, 1, 3, 27, 29, 56};
    for (int off : offs) for (int ubo = 0; ubo < 2; ++ubo) {
        cudaMemset(dD, 0, 128 * N * 4);
        probe<<<1, 128, smem>>>(dA, dB, dD, off, ubo, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("off=%d base_offset=%d: CUDA error %s\n", off, ubo, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hD, dD, 128 * N * 4, cudaMemcpyDeviceToHost);
        int bad = 0; int first = -1;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
            float ref = 0;
            for (int k = 0; k < K; ++k) ref += __bfloat162float(hA[(m + off) * K + k]) * __bfloat162float(hB[n * K + k]);
            if (hD[m * N + n] != ref) { if (first < 0) first = m; ++bad; }
        }
        printf("off=%2d base_offset_field=%d : %s (%d wrong, first bad row %d)\n", off, ubo, bad ? "MISMATCH" : "exact", bad, first);
    }
    return 0;
}
