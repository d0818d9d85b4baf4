// Prototype (not part of the library): layer-3 style 3x3 / pad 1 / stride 1 convolution 256 -> 256 on DENSE 7x7 maps with
// the CTA-pair kernel of ips_b200/csrc/umma_conv_pair.cu, but M tiles = 128 CONSECUTIVE output pixels fed by im2col-mode
// TMA (tools/probe_im2col.cu) instead of 7x7x2 pixel boxes (98 of 128 tile rows).  Checks sampled outputs against a host
// reference and prints the time per launch.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lcuda.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../ips_b200/csrc/umma.cuh"
#include "../ips_b200/csrc/epilogue.cuh"

using bf16 = __nv_bfloat16;
constexpr int TILE_M = 128, BK = 64, BN = 256, A_BYTES = TILE_M * 128, BH_BYTES = (BN / 2) * 128, STAGE_BYTES = A_BYTES + BH_BYTES, STAGES = 5;
constexpr int HH = 7, WW = 7, CIN = 256, COUT = 256, KS = 9 * (CIN / BK);

struct Params { const float* scale; const float* shift; int M, total_pairs, relu; };

__device__ __forceinline__ void tma2_im2col(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h, int n, int ow, int oh) {
    const uint16_t a = (uint16_t)ow, b = (uint16_t)oh;
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(a), "h"(b) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

__global__ void __launch_bounds__(320, 1)
pair_im2col_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                   const Params p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem0 = (umma::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t out_stage0 = smem0 + STAGES * STAGE_BYTES;
    const uint32_t bar0 = out_stage0 + 2u * epi::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bar0 + 8u * (2 * STAGES + 4);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - umma::smem_u32(smem_raw)));
    const uint32_t sc_addr = (tmem_slot + 4u + 15u) & ~15u;
    float* sc_smem = reinterpret_cast<float*>(smem_raw + (sc_addr - umma::smem_u32(smem_raw)));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)umma::cluster_ctarank();
    const int unit0 = blockIdx.x >> 1, ustride = gridDim.x >> 1;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { umma::mbar_init(full_bar(s), 1); umma::mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { umma::mbar_init(tfull_bar(a), 1); umma::mbar_init(tempty_bar(a), 256); }
        umma::fence_barrier_init();
    }
    if (warp == 1) umma::tmem_alloc2(tmem_slot, 2 * BN);
    umma::tc_fence_before();
    __syncthreads();
    umma::cluster_sync_all();
    umma::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int unit = unit0; unit < p.total_pairs; unit += ustride) {
                const int m0 = (2 * unit + rank) * TILE_M;                       // first output pixel of my M tile
                const int n = m0 / (HH * WW), rem = m0 - n * (HH * WW);
                const int py = rem / WW, qx = rem - py * WW;
                for (int ks = 0; ks < KS; ++ks, ++it) {
                    const int stage = it % STAGES;
                    umma::mbar_wait(empty_bar(stage), ((it / STAGES) & 1) ^ 1);
                    const uint32_t a_dst = smem0 + stage * STAGE_BYTES;
                    const uint32_t lead_full = umma::mapa(full_bar(stage), 0);
                    if (rank == 0) umma::mbar_expect_tx(full_bar(stage), 2u * (uint32_t)STAGE_BYTES);
                    const int tap = ks / (CIN / BK), cb = ks - tap * (CIN / BK);
                    const int r = tap / 3, s = tap - 3 * r;
                    tma2_im2col(a_dst, &tmA, lead_full, cb * BK, qx - 1, py - 1, n, s, r);
                    tma2_load_2d(a_dst + A_BYTES, &tmB, lead_full, ks * BK, rank * (BN / 2));
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (rank == 0) {
            const uint32_t leader = umma::elect_one();
            constexpr uint32_t idesc = umma::idesc_bf16_f32(2 * TILE_M, BN);
            uint32_t it = 0, tcount = 0;
            for (int unit = unit0; unit < p.total_pairs; unit += ustride, ++tcount) {
                const uint32_t acc = tcount & 1;
                umma::mbar_wait(tempty_bar(acc), ((tcount >> 1) & 1) ^ 1);
                umma::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int ks = 0; ks < KS; ++ks, ++it) {
                    const int stage = it % STAGES;
                    umma::mbar_wait(full_bar(stage), (it / STAGES) & 1);
                    umma::tc_fence_after();
                    const uint32_t a_addr = smem0 + stage * STAGE_BYTES;
                    const uint64_t adesc = umma::smem_desc_sw128(a_addr), bdesc = umma::smem_desc_sw128(a_addr + A_BYTES);
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
        const int wg = (warp - 2) >> 2, q = warp & 3, row = q * 32 + lane;
        for (int i = tid - 64; i < COUT; i += 256) { sc_smem[i] = p.scale[i]; sc_smem[COUT + i] = p.shift[i]; }
        umma::named_bar_sync(1, 256);
        uint32_t tcount = wg, res_phase = 0;
        const uint32_t stage = out_stage0 + (uint32_t)wg * epi::STAGE_BYTES;
        const bool issuer = (row == 0);
        const uint32_t lead_tempty = umma::mapa(tempty_bar(wg), 0);
        for (int unit = unit0 + wg * ustride; unit < p.total_pairs; unit += 2 * ustride, tcount += 2) {
            const int m0 = (2 * unit + rank) * TILE_M;
            const bool valid = m0 + row < p.M;
            umma::mbar_wait(tfull_bar(wg), (tcount >> 1) & 1);
            umma::tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + wg * BN;
            epi::drain_tile<BN, bf16, true>(t_row, lead_tempty, sc_smem, sc_smem + COUT, valid, false, 0u, 0u, res_phase, p.relu, stage, row,
                                            2u + (uint32_t)wg, issuer,
                                            [&](int s0, uint32_t src) { epi::tma_store_2d(&tmC, src, s0, m0); },
                                            [](int, uint32_t, uint32_t) {});
        }
        if (issuer) epi::bulk_wait0();
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::cluster_sync_all();
    if (warp == 1) umma::tmem_dealloc2(tmem_base, 2 * BN);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char** argv) {
    const int P = argc > 1 ? atoi(argv[1]) : 1536;
    const int M = P * HH * WW, K = 9 * CIN;
    void *f1 = nullptr, *f2 = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f1, cudaEnableDefault, &q);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f2, cudaEnableDefault, &q);
    typedef CUresult (*EncI)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                             cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                             CUtensorMapFloatOOBfill);
    typedef CUresult (*EncT)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    EncI encI = (EncI)f1; EncT encT = (EncT)f2;
    if (!encI || !encT) { printf("driver entry points missing\n"); return 1; }

    std::vector<bf16> hx((size_t)M * CIN), hw((size_t)COUT * K);
    std::vector<float> hs(COUT), hh(COUT);
    uint32_t seed = 12345u;
    auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return ((seed >> 9) & 0xffff) / 65536.0f - 0.5f; };
    for (auto& v : hx) v = __float2bfloat16(rnd());
    for (auto& v : hw) v = __float2bfloat16(rnd() * 0.05f);
    for (int c = 0; c < COUT; ++c) { hs[c] = 0.5f + (c % 7) * 0.1f; hh[c] = (c % 5) * 0.02f - 0.04f; }
    bf16 *dx, *dw, *dy; float *ds, *dh;
    CK(cudaMalloc(&dx, hx.size() * 2)); CK(cudaMalloc(&dw, hw.size() * 2)); CK(cudaMalloc(&dy, (size_t)M * COUT * 2));
    CK(cudaMalloc(&ds, COUT * 4)); CK(cudaMalloc(&dh, COUT * 4));
    CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ds, hs.data(), COUT * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dh, hh.data(), COUT * 4, cudaMemcpyHostToDevice));

    alignas(64) CUtensorMap tmA, tmB, tmC;
    {
        cuuint64_t dims[4] = {CIN, WW, HH, (cuuint64_t)P};
        cuuint64_t strides[3] = {CIN * 2, (cuuint64_t)WW * CIN * 2, (cuuint64_t)HH * WW * CIN * 2};
        int lower[2] = {-1, -1}, upper[2] = {-1, -1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encI(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dx, dims, strides, lower, upper, BK, TILE_M, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode A im2col -> %d\n", (int)r); return 1; }
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)K, COUT}; cuuint64_t strides[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {BK, BN / 2}, estr[2] = {1, 1};
        CUresult r = encT(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dw, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode B -> %d\n", (int)r); return 1; }
    }
    {
        cuuint64_t dims[2] = {COUT, (cuuint64_t)M}; cuuint64_t strides[1] = {COUT * 2};
        cuuint32_t box[2] = {64, TILE_M}, estr[2] = {1, 1};
        CUresult r = encT(&tmC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dy, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode C -> %d\n", (int)r); return 1; }
    }
    Params p{ds, dh, M, ((M + TILE_M - 1) / TILE_M + 1) / 2, 1};
    const size_t smem = (size_t)STAGES * STAGE_BYTES + 2 * epi::STAGE_BYTES + 1024 + 8 * (2 * STAGES + 5) + 32 + 8 * COUT;
    CK(cudaFuncSetAttribute(pair_im2col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem; cfg.stream = 0; cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.gridDim = dim3(148);
    int max_clusters = 0;
    CK(cudaOccupancyMaxActiveClusters(&max_clusters, pair_im2col_kernel, &cfg));
    int clusters = p.total_pairs < max_clusters ? p.total_pairs : max_clusters;
    const int rounds = (p.total_pairs + clusters - 1) / clusters;
    clusters = (p.total_pairs + rounds - 1) / rounds;
    cfg.gridDim = dim3(2 * clusters);
    printf("P=%d M=%d tiles=%d pairs=%d clusters=%d (max %d) smem=%zu\n", P, M, (M + 127) / 128, p.total_pairs, clusters, max_clusters, smem);
    CK(cudaLaunchKernelEx(&cfg, pair_im2col_kernel, tmA, tmB, tmC, p));
    CK(cudaDeviceSynchronize());
    std::vector<bf16> hy((size_t)M * COUT);
    CK(cudaMemcpy(hy.data(), dy, hy.size() * 2, cudaMemcpyDeviceToHost));
    int bad = 0; double worst = 0;
    for (int t = 0; t < 4000; ++t) {
        seed = seed * 1664525u + 1013904223u;
        const int m = (int)(seed % (uint32_t)M), co = (int)((seed >> 11) % COUT);
        const int n = m / 49, py = (m % 49) / 7, qx = m % 7;
        double acc = 0;
        for (int r = 0; r < 3; ++r) for (int s = 0; s < 3; ++s) {
            const int ih = py + r - 1, iw = qx + s - 1;
            if (ih < 0 || ih >= HH || iw < 0 || iw >= WW) continue;
            const bf16* xr = &hx[((size_t)(n * HH + ih) * WW + iw) * CIN];
            const bf16* wr = &hw[(size_t)co * K + (r * 3 + s) * CIN];
            for (int c = 0; c < CIN; ++c) acc += (double)__bfloat162float(xr[c]) * (double)__bfloat162float(wr[c]);
        }
        double ref = acc * hs[co] + hh[co];
        if (ref < 0) ref = 0;
        const double got = __bfloat162float(hy[(size_t)m * COUT + co]);
        const double err = fabs(got - ref);
        if (err > 1e-2 + 1e-2 * fabs(ref)) ++bad;
        if (err > worst) worst = err;
    }
    printf("check: %d of 4000 sampled outputs off (worst abs err %.4f)\n", bad, worst);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) cudaLaunchKernelEx(&cfg, pair_im2col_kernel, tmA, tmB, tmC, p);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) cudaLaunchKernelEx(&cfg, pair_im2col_kernel, tmA, tmB, tmC, p);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double us = ms / 20 * 1e3, tf = 2.0 * M * COUT * K / (us * 1e-6) / 1e12;
    printf("im2col pair kernel: %.1f us per launch, %.1f TFLOP/s (product kernel, 7x7x2 boxes: ~56 us per 1024 patches isolated)\n", us, tf);
    return bad ? 2 : 0;
}
