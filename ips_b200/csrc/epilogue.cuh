// Shared epilogue of the tcgen05 convolution kernels.
//
// One warpgroup (128 threads, thread = accumulator row) drains one TMEM accumulator in slabs
// of 128 bytes per row (64 bf16 / 32 fp32 channels): tcgen05.ld -> folded BatchNorm from shared
// memory, residual, ReLU -> 128B-swizzled staging tile in shared memory -> ONE TMA tensor store
// per slab.  A residual arrives the same way in reverse: one TMA load drops the residual slab INTO the
// staging tile (same box, same swizzle), each thread reads its own 16-byte chunks and overwrites them
// with the result (a warp's direct 16-byte loads of 32 different residual rows cost the 64-channel
// layers 30 us per launch).  The load for a tile's first slab is issued before the warpgroup waits
// for its accumulator, so its latency hides behind the other warpgroup's tile.  Stores therefore leave the SM as full 128-byte lines issued by the TMA unit
// (a warp's direct 16-byte stores to 32 different rows were measured at < 2 TB/s chip-wide and
// bound the 64-channel layers).  The TMA unit clips rows outside the tensor, so ragged tiles
// need no masking.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include "umma.cuh"

namespace epi {

using bf16 = __nv_bfloat16;
constexpr int STAGE_BYTES = 128 * 128;        // 128 rows x 128 bytes

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_ld_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_ld_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <typename T> struct Slab;
template <> struct Slab<bf16> { static constexpr int COLS = 64; };
template <> struct Slab<float> { static constexpr int COLS = 32; };

// 32 accumulator columns of one row -> staging row (swizzled 16-byte chunks)
// has_res: the staging row already holds the residual slab (TMA-loaded); it is read and overwritten in place
template <typename OutT>
__device__ __forceinline__ void stage32(const uint32_t (&v)[32], const float* sc, const float* sh, bool pixel,
                                        bool has_res, int relu, uint32_t row_smem, int row, int chunk0);

// packed helpers: two fp32 lanes per instruction (FFMA2 / FADD2), ReLU on the rounded bf16 pair (HMNMX2)
__device__ __forceinline__ uint64_t fma2(uint32_t v0, uint32_t v1, float s0, float s1, float h0, float h1) {
    uint64_t o;
    asm("{\n\t.reg .b64 a, b, c;\n\t"
        "mov.b64 a, {%1, %2};\n\tmov.b64 b, {%3, %4};\n\tmov.b64 c, {%5, %6};\n\t"
        "fma.rn.f32x2 %0, a, b, c;\n\t}"
        : "=l"(o)
        : "r"(v0), "r"(v1), "r"(__float_as_uint(s0)), "r"(__float_as_uint(s1)), "r"(__float_as_uint(h0)), "r"(__float_as_uint(h1)));
    return o;
}

template <>
__device__ __forceinline__ void stage32<bf16>(const uint32_t (&v)[32], const float* sc, const float* sh, bool pixel,
                                              bool has_res, int relu, uint32_t row_smem, int row, int chunk0) {
    const float4* sc4 = reinterpret_cast<const float4*>(sc);
    const float4* sh4 = reinterpret_cast<const float4*>(sh);
#pragma unroll
    for (int g = 0; g < 4; ++g) {                 // 8 channels = one 16-byte chunk
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        const uint32_t chunk_addr = row_smem + (((uint32_t)(chunk0 + g) ^ (uint32_t)(row & 7)) << 4);
        if (pixel) {
            const float4 s0 = sc4[2 * g], s1 = sc4[2 * g + 1], h0 = sh4[2 * g], h1 = sh4[2 * g + 1];
            const float s[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
            const float h[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
            uint4 rv = make_uint4(0u, 0u, 0u, 0u);
            if (has_res)
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rv.x), "=r"(rv.y), "=r"(rv.z), "=r"(rv.w) : "r"(chunk_addr));
            const uint32_t rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint64_t o = fma2(v[g * 8 + 2 * i], v[g * 8 + 2 * i + 1], s[2 * i], s[2 * i + 1], h[2 * i], h[2 * i + 1]);
                float lo = __uint_as_float((uint32_t)o), hi = __uint_as_float((uint32_t)(o >> 32));
                if (has_res) {                     // bf16 -> fp32 is a 16-bit shift
                    lo += __uint_as_float(rr[i] << 16);
                    hi += __uint_as_float(rr[i] & 0xffff0000u);
                }
                __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
                uint32_t u = *reinterpret_cast<uint32_t*>(&p);
                if (relu) asm("max.bf16x2 %0, %0, %1;" : "+r"(u) : "r"(0u));   // ReLU after rounding == rounding after ReLU
                w[i] = u;
            }
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                     ::"r"(chunk_addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
    }
}

template <>
__device__ __forceinline__ void stage32<float>(const uint32_t (&v)[32], const float* sc, const float* sh, bool pixel,
                                               bool has_res, int relu, uint32_t row_smem, int row, int chunk0) {
    (void)has_res;
#pragma unroll
    for (int g = 0; g < 8; ++g) {                 // 4 channels = one 16-byte chunk
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        if (pixel) {
            const float4 s = reinterpret_cast<const float4*>(sc)[g], h = reinterpret_cast<const float4*>(sh)[g];
            o[0] = fmaf(__uint_as_float(v[g * 4 + 0]), s.x, h.x);
            o[1] = fmaf(__uint_as_float(v[g * 4 + 1]), s.y, h.y);
            o[2] = fmaf(__uint_as_float(v[g * 4 + 2]), s.z, h.z);
            o[3] = fmaf(__uint_as_float(v[g * 4 + 3]), s.w, h.w);
            if (relu) {
#pragma unroll
                for (int i = 0; i < 4; ++i) o[i] = fmaxf(o[i], 0.f);
            }
        }
        const uint32_t chunk = (uint32_t)(chunk0 + g);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                     ::"r"(row_smem + ((chunk ^ (uint32_t)(row & 7)) << 4)), "r"(__float_as_uint(o[0])), "r"(__float_as_uint(o[1])),
                       "r"(__float_as_uint(o[2])), "r"(__float_as_uint(o[3])) : "memory");
    }
}

// Issued by the warpgroup's store thread before it waits for the accumulator: the residual of the tile's first slab
// lands in the staging tile.  load(slab_col0, dst, bar) must issue the TMA load of a 128-row x 128-byte box.
template <class Load>
__device__ __forceinline__ void prefetch_residual(bool issuer, uint32_t res_bar, uint32_t res_bytes, uint32_t stage, const Load& load) {
    if (issuer) {
        bulk_wait_read0();                             // the previous store has finished reading the staging tile
        umma::mbar_expect_tx(res_bar, res_bytes);
        load(0, stage, res_bar);
    }
}

// Drain one accumulator (BN fp32 columns of this thread's row) through the staging tile.
//   t_row      TMEM address of this thread's lane + the accumulator's first column
//   sc / sh    shared-memory scale / shift for the BN columns of this tile
//   pixel      row carries a real output (else zeros are staged)
//   has_res    a residual is added (prefetch_residual was called for this tile); res_bar / res_phase: its mbarrier,
//              res_bytes: bytes one residual box deposits (rows of the M-tile box x 128)
//   stage      this warpgroup's 16 KB staging tile (1024-byte aligned), bar_id: its named barrier
//   issue(slab_col0, stage) is called by ONE thread per slab and must issue the TMA store,
//   load(slab_col0, stage, bar) the TMA load of the residual slab.
//   CLUSTER_RELEASE: tempty_bar is a shared::cluster address (the leader CTA's barrier of a CTA pair)
template <int BN, typename OutT, bool CLUSTER_RELEASE = false, class Issue, class Load>
__device__ __forceinline__ void drain_tile(uint32_t t_row, uint32_t tempty_bar, const float* sc, const float* sh, bool pixel,
                                           bool has_res, uint32_t res_bar, uint32_t res_bytes, uint32_t& res_phase, int relu, uint32_t stage, int row,
                                           uint32_t bar_id, bool issuer, const Issue& issue, const Load& load) {
    constexpr int COLS = Slab<OutT>::COLS;
    const uint32_t row_smem = stage + (uint32_t)row * 128u;
#pragma unroll 1
    for (int s0 = 0; s0 < BN; s0 += COLS) {
        if (has_res) {                                 // residual slab has landed (implies the staging tile was free)
            umma::mbar_wait(res_bar, res_phase);
            res_phase ^= 1u;
        } else {
            if (issuer) bulk_wait_read0();             // previous store has finished reading the staging tile
            umma::named_bar_sync(bar_id, 128);
        }
#pragma unroll
        for (int c0 = 0; c0 < COLS; c0 += 32) {
            uint32_t v[32];
            umma::tmem_ld32(t_row + (uint32_t)(s0 + c0), v);
            umma::tmem_ld_wait();
            if (s0 + c0 + 32 >= BN) {                  // accumulator fully read: hand it back to the MMA warp
                umma::tc_fence_before();
                if (CLUSTER_RELEASE) umma::mbar_arrive_cluster(tempty_bar);
                else                 umma::mbar_arrive(tempty_bar);
            }
            stage32<OutT>(v, sc + s0 + c0, sh + s0 + c0, pixel, has_res, relu, row_smem, row, c0 * (int)sizeof(OutT) / 16);
        }
        umma::fence_proxy_async();                     // generic-proxy smem writes -> visible to the TMA unit
        umma::named_bar_sync(bar_id, 128);
        if (issuer) {
            issue(s0, stage);
            bulk_commit();
            if (has_res && s0 + COLS < BN) {           // next slab's residual (latency exposed; budget is the other warpgroup's tile)
                bulk_wait_read0();
                umma::mbar_expect_tx(res_bar, res_bytes);
                load(s0 + COLS, stage, res_bar);
            }
        }
    }
}

}  // namespace epi
