// Probe: tcgen05.mma with MN-major (a_major = b_major = 1) SWIZZLE_128B operands, the layout a TMA box
// [K rows][64 MN elements] deposits.  D[m][n] = sum_k At[k][m] * Bt[k][n]  (a "TN" GEMM: the weight-gradient shape).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o probe_mnmajor probe_mnmajor.cu
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../ips_b200/csrc/umma.cuh"

constexpr int M = 128, N = 64, K = 64;   // one stage: K = 64 contraction rows

// MN-major SW128 descriptor: LBO = bytes between 64-element MN blocks, SBO = bytes between 8-row K groups
__device__ uint64_t desc_mn(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    const uint32_t lo = ((addr >> 4) & 0x3fffu) | (((lbo_bytes >> 4) & 0x3fffu) << 16);
    const uint32_t hi = ((sbo_bytes >> 4) & 0x3fffu) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}

__global__ void probe(const __nv_bfloat16* At, const __nv_bfloat16* Bt, float* D, int variant) {
    extern __shared__ unsigned char raw[];
    const uint32_t s0 = (umma::smem_u32(raw) + 1023u) & ~1023u;
    unsigned char* base = raw + (s0 - umma::smem_u32(raw));
    // A: two blocks (m 0..63, m 64..127), each [K rows][128 bytes], swizzled chunk ^= (row & 7)
    const uint32_t a_addr = s0, b_addr = s0 + 2 * K * 128, bar = b_addr + K * 128, slot = bar + 8;
    const int tid = threadIdx.x;
    for (int i = tid; i < 2 * K * 8; i += blockDim.x) {
        const int blk = i / (K * 8), r = (i / 8) % K, j = i & 7;
        *reinterpret_cast<uint4*>(base + blk * K * 128 + r * 128 + ((j ^ (r & 7)) << 4)) =
            *reinterpret_cast<const uint4*>(At + r * M + blk * 64 + j * 8);
    }
    for (int i = tid; i < K * 8; i += blockDim.x) {
        const int r = i / 8, j = i & 7;
        *reinterpret_cast<uint4*>(base + 2 * K * 128 + r * 128 + ((j ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(Bt + r * N + j * 8);
    }
    if (tid == 0) { umma::mbar_init(bar, 1); umma::fence_barrier_init(); }
    if (tid < 32) umma::tmem_alloc(slot, 64);
    umma::fence_proxy_async();
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(base + (slot - s0));
    if (tid == 0) {
        // idesc: bf16 x bf16 -> f32, a_major = 1 (bit 15), b_major = 1 (bit 16)
        const uint32_t idesc = umma::idesc_bf16_f32(M, N) | (1u << 15) | (1u << 16);
        for (int k = 0; k < K / 16; ++k) {
            uint64_t ad, bd;
            if (variant == 0) {          // LBO = block stride, SBO = 1024 (8-row groups contiguous)
                ad = desc_mn(a_addr + k * 16 * 128, K * 128, 1024);
                bd = desc_mn(b_addr + k * 16 * 128, K * 128, 1024);
            } else {                     // swapped roles
                ad = desc_mn(a_addr + k * 16 * 128, 1024, K * 128);
                bd = desc_mn(b_addr + k * 16 * 128, 1024, K * 128);
            }
            umma::mma_bf16(tmem, ad, bd, idesc, k != 0);
        }
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
    __nv_bfloat16 *hA = new __nv_bfloat16[K * M], *hB = new __nv_bfloat16[K * N];
    for (int k = 0; k < K; ++k) for (int m = 0; m < M; ++m) hA[k * M + m] = __float2bfloat16((float)((k * 7 + m * 3) % 13 - 6));
    for (int k = 0; k < K; ++k) for (int n = 0; n < N; ++n) hB[k * N + n] = __float2bfloat16((float)((n * 5 + k) % 7 - 3));
    __nv_bfloat16 *dA, *dB; float* dD;
    cudaMalloc(&dA, K * M * 2); cudaMalloc(&dB, K * N * 2); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, hA, K * M * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, K * N * 2, cudaMemcpyHostToDevice);
    const int smem = 2 * K * 128 + K * 128 + 2048;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float* hD = new float[M * N];
    for (int variant = 0; variant < 2; ++variant) {
        cudaMemset(dD, 0, M * N * 4);
        probe<<<1, 128, smem>>>(dA, dB, dD, variant);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
            float ref = 0;
            for (int k = 0; k < K; ++k) ref += __bfloat162float(hA[k * M + m]) * __bfloat162float(hB[k * N + n]);
            bad += (hD[m * N + n] != ref);
        }
        printf("MN-major variant %d (%s): %s (%d wrong of %d)\n", variant, variant == 0 ? "LBO=block stride, SBO=1024" : "LBO=1024, SBO=block stride",
               bad ? "MISMATCH" : "exact", bad, M * N);
    }
    return 0;
}
