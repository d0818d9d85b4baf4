// Sequence-sharded selection: data exchange between the ranks of one node through NVLink peer memory.
//
// north_star / SURVEY.md 8e: each rank scans its slice of the patch axis and keeps a local top-M; the R*M candidates
// (H*T logits + index each) are merged in one global re-score; the winning patches are assembled on the ranks that
// need them.  The reference (architecture/ips_net.py:218-262) is single-GPU; round 1 did the exchange with two NCCL
// all-gathers, torch glue and a 41 MB all-reduce of zero-padded rows.  Here every transfer is ONE kernel that reads
// local memory and stores straight into the peers' exchange buffers (CUDA IPC mappings of PyTorch allocations), and a
// flag protocol in the buffers' headers replaces the collectives' implicit synchronisation:
//
//   push kernel:   payload stores -> __threadfence_system -> device-wide completion counter; the LAST block bumps
//                  the owner's epoch[phase] and writes it to flags[phase][rank] of every peer (st.release.sys)
//   wait kernel:   lane q spins on flags[phase][q] >= epoch[phase] (ld.acquire.sys), bounded by a clock64 budget
//
// A section is rewritten only after every rank passed the NEXT phase's wait, and a rank reaches that wait only after
// the kernels that read the section (same stream) have run -- so no double buffering is needed.
#include "common.cuh"
#include "../../include/ips_b200.h"
#include <cuda.h>
#include <string.h>

namespace {

struct PeerHdr {
    unsigned int epoch[4];        // owner's per-phase epoch (incremented by the owner's push kernels)
    unsigned int done[4];         // block completion counters of the running push kernel
    unsigned int status;          // 0 ok, 1 a wait timed out
    unsigned int pad[7];
    unsigned int flags[4][IPSB_MAX_PEERS];   // written by the peers: their epoch of that phase
};
static_assert(sizeof(PeerHdr) <= IPSB_PEER_HEADER_BYTES, "header too large");

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_peer16(void* p, const int4& v) {
    asm volatile("st.global.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Called by every thread of every block after its payload stores.
__device__ void signal_peers(const ipsb_peer_ctx& c, int phase) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        PeerHdr* me = reinterpret_cast<PeerHdr*>(c.base[c.rank]);
        const unsigned int prev = atomicAdd(&me->done[phase], 1u);
        if (prev == gridDim.x - 1) {                     // every block's stores are ordered before this point
            me->done[phase] = 0;
            const unsigned int e = me->epoch[phase] + 1u;
            me->epoch[phase] = e;
            __threadfence_system();
            for (int q = 0; q < c.world; ++q)
                st_release_sys(&reinterpret_cast<PeerHdr*>(c.base[q])->flags[phase][c.rank], e);
        }
    }
}

__global__ void peer_wait_kernel(ipsb_peer_ctx c, int phase, long long budget_clocks) {
    PeerHdr* me = reinterpret_cast<PeerHdr*>(c.base[c.rank]);
    if ((int)threadIdx.x < c.world) {
        const unsigned int e = me->epoch[phase];
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(&me->flags[phase][threadIdx.x]) - e) < 0) {
            if (clock64() - t0 > budget_clocks) { me->status = 1u; break; }
            __nanosleep(64);
        }
    }
}

struct CandParams {
    const float* z_local; int64_t n_local; const int64_t* cand;
    int B, m, HT; int64_t index_base, L, slot0, cz_off, ci_off;
    const int64_t* rows;      // row of z_local holding candidate (b, j)'s logits; null = cand (table in original order)
};

// one thread per (b, j): HT logits (<= 128 bytes) + one index to every rank
__global__ void __launch_bounds__(256) push_candidates_kernel(ipsb_peer_ctx c, CandParams p) {
    const int64_t total = (int64_t)p.B * p.m;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / p.m), j = (int)(i - (int64_t)b * p.m);
        const int64_t loc = p.cand[i];
        const float* src = p.z_local + ((int64_t)b * p.n_local + (p.rows ? p.rows[i] : loc)) * p.HT;
        const int64_t dst_row = (int64_t)b * p.L + p.slot0 + j;
        float v[32];
#pragma unroll 8
        for (int k = 0; k < p.HT; ++k) v[k] = __ldg(src + k);
        for (int q = 0; q < c.world; ++q) {
            char* base = reinterpret_cast<char*>(c.base[q]);
            float* dz = reinterpret_cast<float*>(base + p.cz_off) + dst_row * p.HT;
            if ((p.HT & 3) == 0) {
                for (int k = 0; k < p.HT; k += 4)
                    *reinterpret_cast<float4*>(dz + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
            } else {
                for (int k = 0; k < p.HT; ++k) dz[k] = v[k];
            }
            reinterpret_cast<int64_t*>(base + p.ci_off)[dst_row] = p.index_base + loc;
        }
    }
    signal_peers(c, 0);
}

struct LogitParams { const float* z_local; int B, HT; int64_t n_local, N, row0, z_off; };

__global__ void __launch_bounds__(256) push_logits_kernel(ipsb_peer_ctx c, LogitParams p) {
    const int64_t per_b = p.n_local * p.HT;                 // floats per image in the slice
    const bool vec = (per_b % 4 == 0) && ((p.row0 * p.HT) % 4 == 0) && ((p.N * p.HT) % 4 == 0);
    if (vec) {
        const int64_t per4 = per_b / 4, total = (int64_t)p.B * per4;
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
            const int64_t b = i / per4, r = i - b * per4;
            const float4 v = __ldg(reinterpret_cast<const float4*>(p.z_local + b * per_b) + r);
            const int64_t dst = (b * p.N + p.row0) * p.HT + 4 * r;
            for (int q = 0; q < c.world; ++q)
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(reinterpret_cast<char*>(c.base[q]) + p.z_off) + dst) = v;
        }
    } else {
        const int64_t total = (int64_t)p.B * per_b;
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
            const int64_t b = i / per_b, r = i - b * per_b;
            const float v = __ldg(p.z_local + i);
            const int64_t dst = (b * p.N + p.row0) * p.HT + r;
            for (int q = 0; q < c.world; ++q)
                reinterpret_cast<float*>(reinterpret_cast<char*>(c.base[q]) + p.z_off)[dst] = v;
        }
    }
    signal_peers(c, 0);
}

struct WinParams {
    const char* local_rows; int64_t n_local, row0; const int64_t* win; const int64_t* ci; int64_t L;
    int B, M; int64_t row_bytes; int slides_per_rank; int64_t out_off; int64_t* idx_out;
};

// one CTA per winner slot (grid-stride): the owner of the patch copies its row to the destination ranks
__global__ void __launch_bounds__(256) push_winners_kernel(ipsb_peer_ctx c, WinParams p) {
    const int64_t slots = (int64_t)p.B * p.M;
    const int64_t row16 = p.row_bytes / 16;
    for (int64_t s = blockIdx.x; s < slots; s += gridDim.x) {
        const int b = (int)(s / p.M), m = (int)(s - (int64_t)b * p.M);
        int64_t g = p.win[s];
        if (p.ci) g = p.ci[(int64_t)b * p.L + g];
        if (threadIdx.x == 0 && p.idx_out) p.idx_out[s] = g;
        if (g < p.row0 || g >= p.row0 + p.n_local) continue;                      // another rank owns this patch (block-uniform)
        const int4* src = reinterpret_cast<const int4*>(p.local_rows + ((int64_t)b * p.n_local + (g - p.row0)) * p.row_bytes);
        int q0 = 0, q1 = c.world;
        int64_t ob = b;
        if (p.slides_per_rank > 0) { q0 = b / p.slides_per_rank; q1 = q0 + 1; ob = b - (int64_t)q0 * p.slides_per_rank; }
        const int64_t dst_off = p.out_off + (ob * p.M + m) * p.row_bytes;
        for (int64_t i = threadIdx.x; i < row16; i += blockDim.x) {
            const int4 v = ipsb::ld_stream16(src + i);
            for (int q = q0; q < q1; ++q)
                st_peer16(reinterpret_cast<char*>(c.base[q]) + dst_off + i * 16, v);
        }
    }
    signal_peers(c, 1);
}

// Small all-gather (synchronised BatchNorm statistics: a few KB per rank).  Two slots alternate with the parity of the
// owner's epoch (identical on all ranks, advanced on the device: correct under CUDA-graph replay): every rank stores its
// `n16` 16-byte words into [slot][rank] of the section of every peer, then signals `phase`.
__global__ void __launch_bounds__(256) allgather_small_kernel(ipsb_peer_ctx c, const int4* __restrict__ src, int n16, int64_t sec_off,
                                                              int64_t slot_stride, int phase) {
    const unsigned int e = reinterpret_cast<const PeerHdr*>(c.base[c.rank])->epoch[phase];   // bumped only after every block signalled
    const int64_t off = sec_off + (int64_t)(e & 1u) * slot_stride + (int64_t)c.rank * n16 * 16;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) {
        const int4 v = src[i];
        for (int q = 0; q < c.world; ++q) st_peer16(reinterpret_cast<char*>(c.base[q]) + off + (int64_t)i * 16, v);
    }
    signal_peers(c, phase);
}

// Waits for `phase`, then copies the (world x n16) words of the slot just filled into dst (a fixed local address).
__global__ void __launch_bounds__(256) wait_gather_kernel(ipsb_peer_ctx c, int phase, long long budget_clocks, int64_t sec_off,
                                                          int64_t slot_stride, int4* __restrict__ dst, int n16_total) {
    PeerHdr* me = reinterpret_cast<PeerHdr*>(c.base[c.rank]);
    const unsigned int e = me->epoch[phase];                       // own push of this exchange already ran: epoch = exchange number
    if ((int)threadIdx.x < c.world) {
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(&me->flags[phase][threadIdx.x]) - e) < 0) {
            if (clock64() - t0 > budget_clocks) { me->status = 1u; break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
    const int4* src = reinterpret_cast<const int4*>(reinterpret_cast<const char*>(c.base[c.rank]) + sec_off + (int64_t)((e - 1u) & 1u) * slot_stride);
    for (int i = threadIdx.x; i < n16_total; i += blockDim.x) dst[i] = __ldcg(src + i);
}

// Synchronised BatchNorm, forward: wait for every rank's [mean | var] (cols each), then in the same kernel combine them
// (equal row counts per rank), write mean / rstd and update the running statistics -- what were a dozen elementwise
// launches per BatchNorm layer.  One block.
__global__ void __launch_bounds__(256) wait_bn_forward_kernel(ipsb_peer_ctx c, int phase, long long budget_clocks, int64_t sec_off,
                                                              int64_t slot_stride, int cols, float momentum, float unbias, float eps,
                                                              float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                              float* __restrict__ running_mean, float* __restrict__ running_var) {
    PeerHdr* me = reinterpret_cast<PeerHdr*>(c.base[c.rank]);
    const unsigned int e = me->epoch[phase];
    if ((int)threadIdx.x < c.world) {
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(&me->flags[phase][threadIdx.x]) - e) < 0) {
            if (clock64() - t0 > budget_clocks) { me->status = 1u; break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
    const float* st = reinterpret_cast<const float*>(reinterpret_cast<const char*>(c.base[c.rank]) + sec_off + (int64_t)((e - 1u) & 1u) * slot_stride);
    const int R = c.world;
    for (int j = threadIdx.x; j < cols; j += blockDim.x) {
        float m = 0.f;
        for (int r = 0; r < R; ++r) m += __ldcg(st + (int64_t)r * 2 * cols + j);
        m /= (float)R;
        float v = 0.f;
        for (int r = 0; r < R; ++r) {
            const float d = __ldcg(st + (int64_t)r * 2 * cols + j) - m;
            v += __ldcg(st + (int64_t)r * 2 * cols + cols + j) + d * d;
        }
        v /= (float)R;
        mean_out[j] = m;
        rstd_out[j] = rsqrtf(v + eps);
        running_mean[j] = running_mean[j] * (1.f - momentum) + m * momentum;
        running_var[j] = running_var[j] * (1.f - momentum) + v * (momentum * unbias);
    }
}

// Wait, then out[i] = sum over the ranks (in rank order: identical on every rank) of the gathered vectors.
__global__ void __launch_bounds__(256) wait_sum_kernel(ipsb_peer_ctx c, int phase, long long budget_clocks, int64_t sec_off,
                                                       int64_t slot_stride, int n, float* __restrict__ out) {
    PeerHdr* me = reinterpret_cast<PeerHdr*>(c.base[c.rank]);
    const unsigned int e = me->epoch[phase];
    if ((int)threadIdx.x < c.world) {
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(&me->flags[phase][threadIdx.x]) - e) < 0) {
            if (clock64() - t0 > budget_clocks) { me->status = 1u; break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
    const float* st = reinterpret_cast<const float*>(reinterpret_cast<const char*>(c.base[c.rank]) + sec_off + (int64_t)((e - 1u) & 1u) * slot_stride);
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        float a = 0.f;
        for (int r = 0; r < c.world; ++r) a += __ldcg(st + (int64_t)r * n + j);
        out[j] = a;
    }
}

__global__ void read_status_kernel(ipsb_peer_ctx c, int* out) { *out = (int)reinterpret_cast<PeerHdr*>(c.base[c.rank])->status; }

int check_ctx(const ipsb_peer_ctx* ctx) {
    IPSB_REQUIRE(ctx != nullptr && ctx->world >= 1 && ctx->world <= IPSB_MAX_PEERS && ctx->rank >= 0 && ctx->rank < ctx->world,
                 "peer: bad context (rank %d of %d)", ctx ? ctx->rank : -1, ctx ? ctx->world : -1);
    for (int q = 0; q < ctx->world; ++q) IPSB_REQUIRE(ctx->base[q] != nullptr, "peer: rank %d's exchange buffer is not mapped", q);
    return 0;
}

int grid_for(int64_t work_items, int per_block) {
    int64_t g = (work_items + per_block - 1) / per_block;
    const int64_t cap = (int64_t)ipsb::sm_count() * 8;
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}

typedef CUresult (*GetRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
GetRangeFn range_fn() {
    static GetRangeFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (GetRangeFn)ptr;
    }
    return fn;
}

}  // namespace

extern "C" {

int ipsb_peer_export(const void* ptr, unsigned char handle_out[64], int64_t* offset_out) {
    IPSB_REQUIRE(ptr && handle_out && offset_out, "peer_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    GetRangeFn fn = range_fn();
    IPSB_REQUIRE(fn != nullptr, "peer_export: cuMemGetAddressRange not available from the driver");
    CUdeviceptr base = 0;
    size_t size = 0;
    const CUresult r = fn(&base, &size, (CUdeviceptr)ptr);
    IPSB_REQUIRE(r == CUDA_SUCCESS, "peer_export: cuMemGetAddressRange failed with %d", (int)r);
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, (void*)base);
    if (e != cudaSuccess)
        return ipsb::fail("peer_export: cudaIpcGetMemHandle -> %s (the exchange buffer must come from a cudaMalloc'ed block: "
                          "PyTorch's allocator without expandable_segments / cudaMallocAsync)", cudaGetErrorString(e));
    memcpy(handle_out, &h, 64);
    *offset_out = (int64_t)((CUdeviceptr)ptr - base);
    return 0;
}

int ipsb_peer_open(const unsigned char handle[64], int64_t offset, void** ptr_out) {
    IPSB_REQUIRE(handle && ptr_out && offset >= 0, "peer_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void* base = nullptr;
    IPSB_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr_out = (char*)base + offset;
    return 0;
}

int ipsb_peer_close(void* ptr, int64_t offset) {
    IPSB_REQUIRE(ptr != nullptr, "peer_close: null pointer");
    IPSB_CUDA(cudaIpcCloseMemHandle((char*)ptr - offset));
    return 0;
}

int ipsb_peer_status(const ipsb_peer_ctx* ctx, int* status_out, void* stream) {
    if (int rc = check_ctx(ctx)) return rc;
    IPSB_REQUIRE(status_out != nullptr, "peer_status: null output");
    int* d = nullptr;
    IPSB_CUDA(cudaMalloc(&d, sizeof(int)));
    read_status_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(*ctx, d);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(status_out, d, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(d);
    if (e != cudaSuccess) return ipsb::fail("peer_status: %s", cudaGetErrorString(e));
    return 0;
}

int ipsb_peer_push_candidates(const ipsb_peer_ctx* ctx, const float* z_local, int64_t n_local, const int64_t* cand,
                              int B, int m, int HT, int64_t index_base, int64_t L, int64_t slot0,
                              int64_t cz_off, int64_t ci_off, void* stream) {
    return ipsb_peer_push_candidates_rows(ctx, z_local, n_local, cand, nullptr, B, m, HT, index_base, L, slot0, cz_off, ci_off, stream);
}

int ipsb_peer_push_candidates_rows(const ipsb_peer_ctx* ctx, const float* z_local, int64_t n_local, const int64_t* cand,
                                   const int64_t* rows, int B, int m, int HT, int64_t index_base, int64_t L, int64_t slot0,
                                   int64_t cz_off, int64_t ci_off, void* stream) {
    if (int rc = check_ctx(ctx)) return rc;
    IPSB_REQUIRE(z_local && cand && B > 0 && m > 0 && HT > 0 && HT <= 32, "peer_push_candidates: bad shape B=%d m=%d HT=%d", B, m, HT);
    IPSB_REQUIRE(slot0 >= 0 && slot0 + m <= L, "peer_push_candidates: slots [%lld, %lld) outside the %lld-entry list",
                 (long long)slot0, (long long)(slot0 + m), (long long)L);
    IPSB_REQUIRE(cz_off >= IPSB_PEER_HEADER_BYTES && ci_off >= IPSB_PEER_HEADER_BYTES && cz_off % 16 == 0 && ci_off % 8 == 0,
                 "peer_push_candidates: sections must lie behind the header, 16-byte aligned");
    CandParams p{z_local, n_local, cand, B, m, HT, index_base, L, slot0, cz_off, ci_off, rows};
    push_candidates_kernel<<<grid_for((int64_t)B * m, 256), 256, 0, (cudaStream_t)stream>>>(*ctx, p);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_peer_push_logits(const ipsb_peer_ctx* ctx, const float* z_local, int B, int64_t n_local, int HT,
                          int64_t N, int64_t row0, int64_t z_off, void* stream) {
    if (int rc = check_ctx(ctx)) return rc;
    IPSB_REQUIRE(z_local && B > 0 && n_local > 0 && HT > 0 && row0 >= 0 && row0 + n_local <= N, "peer_push_logits: bad shape");
    IPSB_REQUIRE(z_off >= IPSB_PEER_HEADER_BYTES && z_off % 16 == 0, "peer_push_logits: section must lie behind the header, 16-byte aligned");
    LogitParams p{z_local, B, HT, n_local, N, row0, z_off};
    push_logits_kernel<<<grid_for((int64_t)B * n_local * HT / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(*ctx, p);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_peer_allgather_small(const ipsb_peer_ctx* ctx, const void* src, int64_t bytes, int64_t sec_off, int64_t slot_stride,
                              int phase, void* dst, void* stream) {
    if (int rc = check_ctx(ctx)) return rc;
    IPSB_REQUIRE(src && dst && bytes > 0 && bytes % 16 == 0 && ((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 16) == 0 && bytes <= (1 << 20),
                 "peer_allgather_small: %lld bytes (16-byte multiples up to 1 MiB, aligned buffers)", (long long)bytes);
    IPSB_REQUIRE(sec_off >= IPSB_PEER_HEADER_BYTES && sec_off % 16 == 0 && slot_stride >= bytes * ctx->world && slot_stride % 16 == 0 &&
                 phase >= 0 && phase < 4, "peer_allgather_small: bad section / phase");
    const int n16 = (int)(bytes / 16);
    int grid = (n16 + 255) / 256;
    if (grid > 8) grid = 8;
    allgather_small_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*ctx, (const int4*)src, n16, sec_off, slot_stride, phase);
    IPSB_LAUNCH_CHECK();
    wait_gather_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(*ctx, phase, 8000000000ll, sec_off, slot_stride, (int4*)dst, n16 * ctx->world);
    IPSB_LAUNCH_CHECK();
    return 0;
}

static int push_small(const ipsb_peer_ctx* ctx, const void* src, int64_t bytes, int64_t sec_off, int64_t slot_stride, int phase,
                      void* stream, const char* who) {
    IPSB_REQUIRE(src && bytes > 0 && bytes % 16 == 0 && ((uintptr_t)src % 16) == 0 && bytes <= (1 << 20),
                 "%s: %lld bytes (16-byte multiples up to 1 MiB, aligned buffer)", who, (long long)bytes);
    IPSB_REQUIRE(sec_off >= IPSB_PEER_HEADER_BYTES && sec_off % 16 == 0 && slot_stride >= bytes * ctx->world && slot_stride % 16 == 0 &&
                 phase >= 0 && phase < 4, "%s: bad section / phase", who);
    const int n16 = (int)(bytes / 16);
    int grid = (n16 + 255) / 256;
    if (grid > 8) grid = 8;
    allgather_small_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*ctx, (const int4*)src, n16, sec_off, slot_stride, phase);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_peer_bn_forward(const ipsb_peer_ctx* ctx, const float* mean_var, int cols, int64_t sec_off, int64_t slot_stride, int phase,
                         float momentum, float unbias, float eps, float* mean_out, float* rstd_out, float* running_mean,
                         float* running_var, void* stream) {
    if (int rc = check_ctx(ctx)) return rc;
    IPSB_REQUIRE(cols > 0 && mean_out && rstd_out && running_mean && running_var, "peer_bn_forward: null argument");
    if (int rc = push_small(ctx, mean_var, (int64_t)cols * 8, sec_off, slot_stride, phase, stream, "peer_bn_forward")) return rc;
    wait_bn_forward_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(*ctx, phase, 8000000000ll, sec_off, slot_stride, cols, momentum, unbias, eps,
                                                               mean_out, rstd_out, running_mean, running_var);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_peer_allgather_sum(const ipsb_peer_ctx* ctx, const float* src, int n, int64_t sec_off, int64_t slot_stride, int phase,
                            float* out, void* stream) {
    if (int rc = check_ctx(ctx)) return rc;
    IPSB_REQUIRE(n > 0 && out, "peer_allgather_sum: null argument");
    if (int rc = push_small(ctx, src, (int64_t)n * 4, sec_off, slot_stride, phase, stream, "peer_allgather_sum")) return rc;
    wait_sum_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(*ctx, phase, 8000000000ll, sec_off, slot_stride, n, out);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_peer_wait(const ipsb_peer_ctx* ctx, int phase, void* stream) {
    if (int rc = check_ctx(ctx)) return rc;
    IPSB_REQUIRE(phase >= 0 && phase < 4, "peer_wait: phase %d", phase);
    peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(*ctx, phase, 8000000000ll);      // ~4 s at 2 GHz, then status = 1
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_peer_push_winners(const ipsb_peer_ctx* ctx, const void* local_rows, int64_t n_local, int64_t row0,
                           const int64_t* win, const int64_t* ci, int64_t L, int B, int M, int64_t row_bytes,
                           int slides_per_rank, int64_t out_off, int64_t* idx_out, void* stream) {
    if (int rc = check_ctx(ctx)) return rc;
    IPSB_REQUIRE(local_rows && win && B > 0 && M > 0 && n_local > 0, "peer_push_winners: bad shape");
    IPSB_REQUIRE(row_bytes > 0 && row_bytes % 16 == 0 && ((uintptr_t)local_rows % 16) == 0, "peer_push_winners: rows must be 16-byte multiples (row_bytes=%lld)",
                 (long long)row_bytes);
    IPSB_REQUIRE(out_off >= IPSB_PEER_HEADER_BYTES && out_off % 16 == 0, "peer_push_winners: section must lie behind the header, 16-byte aligned");
    IPSB_REQUIRE(slides_per_rank >= 0 && (slides_per_rank == 0 || (int64_t)slides_per_rank * ctx->world >= B),
                 "peer_push_winners: %d slides per rank cannot hold %d slides on %d ranks", slides_per_rank, B, ctx->world);
    WinParams p{(const char*)local_rows, n_local, row0, win, ci, L, B, M, row_bytes, slides_per_rank, out_off, idx_out};
    int64_t g = (int64_t)B * M;
    const int64_t cap = (int64_t)ipsb::sm_count() * 8;
    if (g > cap) g = cap;
    push_winners_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(*ctx, p);
    IPSB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
