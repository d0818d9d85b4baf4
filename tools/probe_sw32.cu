// Probe: K-major SWIZZLE_32B A operand (rows of 32 bytes = 16 bf16) with a start address shifted by an arbitrary
// number of rows, one tcgen05.mma (K = 16).  B: K-major SWIZZLE_128B slab, K slice selected by +32 bytes.
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../ips_b200/csrc/umma.cuh"

constexpr int ROWS = 512, N = 64, KB = 64;

__device__ uint64_t desc_sw32(uint32_t addr) {
    const uint32_t lo = ((addr >> 4) & 0x3fffu) | (1u << 16);
    const uint32_t hi = (256u >> 4) | (1u << 14) | (6u << 29);     // SBO = 256 B (8 rows of 32 B), SWIZZLE_32B
    return ((uint64_t)hi << 32) | lo;
}

__global__ void probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int off, int kslice) {
    extern __shared__ unsigned char raw[];
    const uint32_t s0 = (umma::smem_u32(raw) + 1023u) & ~1023u;
    unsigned char* base = raw + (s0 - umma::smem_u32(raw));
    const uint32_t a_addr = s0, b_addr = s0 + ROWS * 32, bar = b_addr + N * 128, slot = bar + 8;
    const int tid = threadIdx.x;
    // A: row r = 32 bytes = chunks j in {0,1}; SW32: chunk ^= (addr bit 7) = (r >> 2) & 1
    for (int i = tid; i < ROWS * 2; i += blockDim.x) {
        const int r = i >> 1, j = i & 1;
        *reinterpret_cast<uint4*>(base + r * 32 + ((j ^ ((r >> 2) & 1)) << 4)) = *reinterpret_cast<const uint4*>(A + r * 16 + j * 8);
    }
    for (int i = tid; i < N * 8; i += blockDim.x) {
        const int r = i >> 3, j = i & 7;
        *reinterpret_cast<uint4*>(base + ROWS * 32 + r * 128 + ((j ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * KB + j * 8);
    }
    if (tid == 0) { umma::mbar_init(bar, 1); umma::fence_barrier_init(); }
    if (tid < 32) umma::tmem_alloc(slot, 64);
    umma::fence_proxy_async();
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(base + (slot - s0));
    if (tid == 0) {
        umma::mma_bf16(tmem, desc_sw32(a_addr + off * 32), umma::smem_desc_sw128(b_addr) + 2u * kslice, umma::idesc_bf16_f32(128, N), 0);
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
    __nv_bfloat16 *hA = new __nv_bfloat16[ROWS * 16], *hB = new __nv_bfloat16[N * KB];
    for (int r = 0; r < ROWS; ++r) for (int k = 0; k < 16; ++k) hA[r * 16 + k] = __float2bfloat16((float)((r * 7 + k * 3) % 13 - 6));
    for (int n = 0; n < N; ++n) for (int k = 0; k < KB; ++k) hB[n * KB + k] = __float2bfloat16((float)((n * 5 + k) % 7 - 3));
    __nv_bfloat16 *dA, *dB; float* dD;
    cudaMalloc(&dA, ROWS * 32); cudaMalloc(&dB, N * KB * 2); cudaMalloc(&dD, 128 * N * 4);
    cudaMemcpy(dA, hA, ROWS * 32, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, N * KB * 2, cudaMemcpyHostToDevice);
    const int smem = ROWS * 32 + N * 128 + 2048;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float* hD = new float[128 * N];
    const int offs[] = {0, 8, 1, 3, 4, 53, 162, 215};
    for (int off : offs) for (int ks = 0; ks < 4; ks += 3) {
        cudaMemset(dD, 0, 128 * N * 4);
        probe<<<1, 128, smem>>>(dA, dB, dD, off, ks);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("off=%d: CUDA error %s\n", off, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hD, dD, 128 * N * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
            float ref = 0;
            for (int k = 0; k < 16; ++k) ref += __bfloat162float(hA[(m + off) * 16 + k]) * __bfloat162float(hB[n * KB + ks * 16 + k]);
            bad += (hD[m * N + n] != ref);
        }
        printf("SW32 A shifted by %3d rows, B k-slice %d: %s (%d wrong)\n", off, ks, bad ? "MISMATCH" : "exact", bad);
    }
    return 0;
}
