// Scoring + stable top-M + the sequential selection loop, on a per-patch logit table.
//
// Reference semantics (architecture/ips_net.py:136-155,213-241 and
// architecture/transformer.py:29-34,143-148): every iteration concatenates the M
// memory entries with the next <=I scanned patches, scores the (M+I) buffer with
// softmax over the buffer axis per (head, token), averages heads then tokens,
// and keeps the top M, best first.  Tie-break contract of this library: lowest
// buffer position wins (a stable descending sort), see SURVEY.md section 8c P1.
//
// One CTA owns one image for the whole loop: the memory set lives in shared
// memory across iterations, logits are re-read from the (L2 resident) table.
#include <cooperative_groups.h>
#include <stdlib.h>
#include "common.cuh"
#include "../../include/ips_b200.h"

namespace cg = cooperative_groups;

namespace {

constexpr int kMaxHT = 32;
constexpr int kMaxLpad = 16384;

__device__ __forceinline__ uint32_t order_bits(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float order_bits_inv(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}
// sort key: score descending, then buffer position ascending
__device__ __forceinline__ unsigned long long make_key(float s, uint32_t l) {
    return ((unsigned long long)order_bits(s) << 32) | (unsigned long long)(0xffffffffu - l);
}
__device__ __forceinline__ uint32_t key_pos(unsigned long long k) { return 0xffffffffu - (uint32_t)(k & 0xffffffffull); }

struct Scratch {
    float mx[kMaxHT];
    float sm[kMaxHT];
    float part[32 * kMaxHT];   // per-warp partials
    float slow[1024];          // generic (non power-of-two HT) path
};

// Block-wide reduction of one value per thread into HT classes (class = tid % HT).
// kMax selects max / sum.  Result in out[0..HT).  Deterministic for a fixed blockDim.
template <bool kMax>
__device__ void reduce_classes(float v, int HT, bool pow2, int nt_eff, Scratch* sc, float* out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    if (pow2) {
        for (int o = 16; o >= HT; o >>= 1) {
            float u = __shfl_xor_sync(0xffffffffu, v, o);
            v = kMax ? fmaxf(v, u) : (v + u);
        }
        if (lane < HT) sc->part[warp * kMaxHT + lane] = v;
        __syncthreads();
        if (tid < HT) {
            float r = sc->part[tid];
            for (int w = 1; w < nwarps; ++w) {
                float u = sc->part[w * kMaxHT + tid];
                r = kMax ? fmaxf(r, u) : (r + u);
            }
            out[tid] = r;
        }
    } else {
        sc->slow[tid] = v;
        __syncthreads();
        if (tid < HT) {
            float r = sc->slow[tid];
            for (int i = tid + HT; i < nt_eff; i += HT) {
                float u = sc->slow[i];
                r = kMax ? fmaxf(r, u) : (r + u);
            }
            out[tid] = r;
        }
    }
    __syncthreads();
}

// Candidate l of the current buffer -> row of the logit table.
struct Buffer {
    const float* z;          // (N, HT) table of this image
    const int64_t* perm;     // scan order of this image or nullptr
    const int* mem_src;      // smem, original index of memory entry l < M
    int M, lo;               // new entries l >= M are scan positions lo + (l - M)
    __device__ __forceinline__ int src(int l) const {
        if (l < M) return mem_src[l];
        int p = lo + (l - M);
        return perm ? (int)perm[p] : p;
    }
};

// softmax over the buffer axis per (h,t) -> mean over heads -> mean over tokens.
// Writes sort keys (and optionally raw scores) for l < L.
template <class RowOf>
__device__ void score_buffer(const RowOf& row_of, int L, int H, int T, Scratch* sc,
                             unsigned long long* keys, float* scores_out) {
    const int HT = H * T, tid = threadIdx.x, nthreads = blockDim.x;
    const bool pow2 = (HT & (HT - 1)) == 0;
    const int nt_eff = (nthreads / HT) * HT;
    const int ht = tid % HT, l0 = tid / HT, lstep = nt_eff / HT;
    // pass 1: max per (h,t)
    float v = -INFINITY;
    if (tid < nt_eff)
        for (int l = l0; l < L; l += lstep) v = fmaxf(v, row_of(l)[ht]);
    reduce_classes<true>(v, HT, pow2, nt_eff, sc, sc->mx);
    // pass 2: sum of exp
    v = 0.f;
    if (tid < nt_eff) {
        const float m = sc->mx[ht];
        for (int l = l0; l < L; l += lstep) v += expf(row_of(l)[ht] - m);
    }
    reduce_classes<false>(v, HT, pow2, nt_eff, sc, sc->sm);
    // pass 3: one thread per candidate
    for (int l = tid; l < L; l += nthreads) {
        const float* zr = row_of(l);
        float tok = 0.f;
        for (int t = 0; t < T; ++t) {
            float hs = 0.f;
            for (int h = 0; h < H; ++h) {
                const int c = h * T + t;
                hs += expf(zr[c] - sc->mx[c]) / sc->sm[c];
            }
            tok += hs / (float)H;
        }
        const float s = tok / (float)T;
        if (keys) keys[l] = make_key(s, (uint32_t)l);
        if (scores_out) scores_out[l] = s;
    }
}

// in-place bitonic sort, descending, n = power of two
__device__ void bitonic_desc(unsigned long long* keys, int n) {
    const int tid = threadIdx.x, nthreads = blockDim.x;
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < (n >> 1); i += nthreads) {
                const int lo = 2 * i - (i & (j - 1));
                const int hi = lo + j;
                const bool desc = (lo & k) == 0;
                const unsigned long long a = keys[lo], b = keys[hi];
                if ((a < b) == desc) { keys[lo] = b; keys[hi] = a; }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

int host_next_pow2(int v);
}  // namespace
extern "C" int64_t ipsb_select_loop_workspace_bytes(int B, int N, int HT, int M);
namespace {

struct LoopParams {
    const float* z;
    const int64_t* perm;
    int64_t perm_stride;
    int N, H, T, M, I;
    int64_t* out_pos;
    int64_t* out_src;
    float* out_score;
};

__global__ void __launch_bounds__(1024, 1) select_loop_kernel(LoopParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x, tid = threadIdx.x, nthreads = blockDim.x;
    const int HT = p.H * p.T, M = p.M;
    const int Lmax = M + min(p.I, p.N - M);
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
    int* mem_pos = reinterpret_cast<int*>(keys + next_pow2(Lmax));
    int* mem_src = mem_pos + M;
    Scratch* sc = reinterpret_cast<Scratch*>(mem_src + M);

    const float* z = p.z + (int64_t)b * p.N * HT;
    const int64_t* perm = p.perm ? p.perm + (int64_t)b * p.perm_stride : nullptr;
    for (int r = tid; r < M; r += nthreads) {
        mem_pos[r] = r;
        mem_src[r] = perm ? (int)perm[r] : r;
    }
    __syncthreads();

    const int n_iter = (p.N - M + p.I - 1) / p.I;
    for (int it = 0; it < n_iter; ++it) {
        Buffer buf{z, perm, mem_src, M, M + it * p.I};
        const int hi = min(buf.lo + p.I, p.N);
        const int L = M + (hi - buf.lo);
        const int Lpad = next_pow2(L);
        auto row_of = [&](int l) { return z + (int64_t)buf.src(l) * HT; };
        score_buffer(row_of, L, p.H, p.T, sc, keys, nullptr);
        for (int l = L + tid; l < Lpad; l += nthreads) keys[l] = 0ull;
        __syncthreads();
        bitonic_desc(keys, Lpad);
        // winners: translate buffer position -> (scan position, original index); the
        // slot r is read and rewritten by the same thread, old mem_* stay intact meanwhile
        for (int r = tid; r < M; r += nthreads) {
            const unsigned long long k = keys[r];
            const int l = (int)key_pos(k);
            const int pos = (l < M) ? mem_pos[l] : buf.lo + (l - M);
            const int src = buf.src(l);
            const bool last = (it == n_iter - 1);
            if (last) {
                p.out_pos[(int64_t)b * M + r] = pos;
                p.out_src[(int64_t)b * M + r] = src;
                if (p.out_score) p.out_score[(int64_t)b * M + r] = order_bits_inv((uint32_t)(k >> 32));
            }
            keys[r] = ~(((unsigned long long)(uint32_t)pos << 32) | (uint32_t)src);
        }
        // the memory buffer is kept in SCAN ORDER (ascending position), so "lowest buffer position"
        // always means "scanned first": re-sort the winners by position (descending on the complement)
        const int Mpad = next_pow2(M);
        for (int r = M + tid; r < Mpad; r += nthreads) keys[r] = 0ull;
        __syncthreads();
        bitonic_desc(keys, Mpad);
        for (int r = tid; r < M; r += nthreads) {
            const unsigned long long k = ~keys[r];
            mem_pos[r] = (int)(k >> 32);
            mem_src[r] = (int)(k & 0xffffffffull);
        }
        __syncthreads();
    }
}


// ---- register-resident variant for long buffers -------------------------------------------------
// Each of 1024 threads owns E consecutive buffer positions.  The memory buffer is kept in scan
// order, so an iteration only needs the SET of survivors: a 4-pass radix select finds the score of
// rank M, a stable compaction keeps everything above it plus the first-scanned equals (the
// library's tie-break contract), and the winners are sorted once, after the last iteration.
template <int HT, int E>
__global__ void __launch_bounds__(1024, 1) select_loop_reg_kernel(LoopParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NT = 1024;
    constexpr int CAP = NT * E;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = p.M;
    uint32_t* keyA = reinterpret_cast<uint32_t*>(smem_raw);            // [CAP] order bits of the scores
    int* cand = reinterpret_cast<int*>(keyA + CAP);                    // [CAP] table row of each buffer position
    uint32_t* fin_key = reinterpret_cast<uint32_t*>(cand + CAP);       // [M] score bits of the survivors
    int* mem_pos = reinterpret_cast<int*>(fin_key + M);                // [2][M] scan positions (double buffered)
    int* mem_src = mem_pos + M;                                        // interleaved halves: pos0 src0 pos1 src1
    int* hist = mem_pos + 4 * M;                                       // [256]
    int* sel = hist + 256;                                             // [2]
    int* wsum = sel + 2;                                               // [64]
    float* red = reinterpret_cast<float*>(wsum + 64);                  // [32][HT] partials, [HT] result, [HT] maxima

    const float* z = p.z + (int64_t)b * p.N * HT;
    const int64_t* perm = p.perm ? p.perm + (int64_t)b * p.perm_stride : nullptr;
    for (int r = tid; r < M; r += NT) {
        mem_pos[r] = r;
        mem_src[r] = perm ? (int)perm[r] : r;
    }
    __syncthreads();

    const int n_iter = (p.N - M + p.I - 1) / p.I;
    for (int it = 0; it < n_iter; ++it) {
        const int lo = M + it * p.I;
        const int hi = min(lo + p.I, p.N);
        const int L = M + (hi - lo);
        // ---- candidate -> row of the logit table, cached in shared memory
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int l = tid * E + e;
            int src = -1;
            if (l < L) {
                if (l < M) src = mem_src[l];
                else { const int pp = lo + (l - M); src = perm ? (int)perm[pp] : pp; }
            }
            cand[l] = src;
        }
        // ---- pass 1: max per (h,t); E independent 32-byte row loads in flight per thread
        float acc[HT];
#pragma unroll
        for (int c = 0; c < HT; ++c) acc[c] = -INFINITY;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int src = cand[tid * E + e];
            if (src >= 0) {
                const float4* row = reinterpret_cast<const float4*>(z + (int64_t)src * HT);
#pragma unroll
                for (int c = 0; c < HT / 4; ++c) {
                    const float4 v = __ldg(row + c);
                    acc[4 * c] = fmaxf(acc[4 * c], v.x); acc[4 * c + 1] = fmaxf(acc[4 * c + 1], v.y);
                    acc[4 * c + 2] = fmaxf(acc[4 * c + 2], v.z); acc[4 * c + 3] = fmaxf(acc[4 * c + 3], v.w);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < HT; ++c) acc[c] = ipsb::warp_max(acc[c]);
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < HT; ++c) red[warp * HT + c] = acc[c];
        }
        __syncthreads();
        if (tid < HT) {
            float m = red[tid];
            for (int w = 1; w < 32; ++w) m = fmaxf(m, red[w * HT + tid]);
            red[32 * HT + tid] = m;
        }
        __syncthreads();
        float mx[HT];
#pragma unroll
        for (int c = 0; c < HT; ++c) mx[c] = red[32 * HT + c];
        if (tid < HT) red[33 * HT + tid] = red[32 * HT + tid];          // maxima kept for pass 3
        __syncthreads();
        // ---- pass 2: sums of exp
#pragma unroll
        for (int c = 0; c < HT; ++c) acc[c] = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int src = cand[tid * E + e];
            if (src >= 0) {
                const float4* row = reinterpret_cast<const float4*>(z + (int64_t)src * HT);
#pragma unroll
                for (int c = 0; c < HT / 4; ++c) {
                    const float4 v = __ldg(row + c);
                    acc[4 * c] += expf(v.x - mx[4 * c]); acc[4 * c + 1] += expf(v.y - mx[4 * c + 1]);
                    acc[4 * c + 2] += expf(v.z - mx[4 * c + 2]); acc[4 * c + 3] += expf(v.w - mx[4 * c + 3]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < HT; ++c) acc[c] = ipsb::warp_sum(acc[c]);
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < HT; ++c) red[warp * HT + c] = acc[c];
        }
        __syncthreads();
        if (tid < HT) {
            float sm = red[tid];
            for (int w = 1; w < 32; ++w) sm += red[w * HT + tid];
            red[32 * HT + tid] = sm;
        }
        __syncthreads();
        // ---- pass 3: scores -> ascending keys of the inverted order bits
        const float* mxs = red + 32 * HT;   // sums now; keep maxima in registers
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int l = tid * E + e;
            const int src = cand[l];
            uint32_t key = 0u;
            if (src >= 0) {
                const float* zrow = z + (int64_t)src * HT;
                float tok = 0.f;
                for (int t = 0; t < p.T; ++t) {
                    float hs = 0.f;
                    for (int h = 0; h < p.H; ++h) {
                        const int c = h * p.T + t;
                        hs += expf(__ldg(zrow + c) - red[33 * HT + c]) / mxs[c];
                    }
                    tok += hs / (float)p.H;
                }
                key = order_bits(tok / (float)p.T);
            }
            keyA[l] = key;
        }
        __syncthreads();
        // ---- radix SELECT of the M-th largest key: 4 passes of 8 bits, most significant first
        uint32_t prefix = 0;
        int remaining = M;                      // rank still to locate inside the current prefix class
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            if (tid < 256) hist[tid] = 0;
            __syncthreads();
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int l = tid * E + e;
                if (l < L) {
                    const uint32_t k = keyA[l];
                    if (pass == 0 || (k >> (shift + 8)) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1);
                }
            }
            __syncthreads();
            if (warp == 0) {                    // lane j owns bins [255-8j-7, 255-8j]: scan from the top
                int c[8], tot = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { c[j] = hist[255 - 8 * lane - j]; tot += c[j]; }
                int inc = tot;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
                const int before = inc - tot;   // keys in bins above this lane's range
                if (before < remaining && remaining <= inc) {
                    int run = before;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (run < remaining && remaining <= run + c[j]) { sel[0] = 255 - 8 * lane - j; sel[1] = remaining - run; }
                        run += c[j];
                    }
                }
            }
            __syncthreads();
            prefix = (prefix << 8) | (uint32_t)sel[0];
            remaining = sel[1];
            __syncthreads();
        }
        const uint32_t thr = prefix;            // key of the rank-M candidate; `remaining` of its equals are kept
        // ---- stable compaction in buffer (= scan) order: keep key > thr, and the first `remaining` with key == thr
        int ngt = 0, neq = 0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int l = tid * E + e;
            if (l < L) { const uint32_t k = keyA[l]; ngt += (k > thr); neq += (k == thr); }
        }
        int inc_gt = ngt, inc_eq = neq;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc_gt, o), v = __shfl_up_sync(0xffffffffu, inc_eq, o);
            if (lane >= o) { inc_gt += u; inc_eq += v; }
        }
        if (lane == 31) { wsum[warp] = inc_gt; wsum[32 + warp] = inc_eq; }
        __syncthreads();
        if (warp == 0) {
            int a = wsum[lane], c2 = wsum[32 + lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, a, o), v = __shfl_up_sync(0xffffffffu, c2, o);
                if (lane >= o) { a += u; c2 += v; }
            }
            wsum[lane] = a; wsum[32 + lane] = c2;
        }
        __syncthreads();
        int gt_before = (inc_gt - ngt) + (warp ? wsum[warp - 1] : 0);
        int eq_before = (inc_eq - neq) + (warp ? wsum[32 + warp - 1] : 0);
        int* new_pos = mem_pos + 2 * M;         // second half of the double-buffered memory arrays
        int* new_src = mem_src + 2 * M;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int l = tid * E + e;
            if (l < L) {
                const uint32_t k = keyA[l];
                const bool keep = (k > thr) || (k == thr && eq_before < remaining);
                if (keep) {
                    const int dst = gt_before + (eq_before < remaining ? eq_before : remaining);
                    new_pos[dst] = (l < M) ? mem_pos[l] : lo + (l - M);
                    new_src[dst] = cand[l];
                    fin_key[dst] = k;           // score bits of the survivors (needed for the final ordering)
                }
                gt_before += (k > thr);
                eq_before += (k == thr);
            }
        }
        __syncthreads();
        for (int r = tid; r < M; r += NT) { mem_pos[r] = new_pos[r]; mem_src[r] = new_src[r]; }
        __syncthreads();
    }
    // ---- final ordering: score descending, ties -> scanned first (memory is in scan order)
    unsigned long long* keys64 = reinterpret_cast<unsigned long long*>(smem_raw);     // aliases keyA / cand (no longer needed)
    const int Mpad = next_pow2(M);
    uint32_t fk[(8192 + NT - 1) / NT];
    for (int r = tid, j = 0; r < M; r += NT, ++j) fk[j] = fin_key[r];
    __syncthreads();
    for (int r = tid, j = 0; r < Mpad; r += NT, ++j)
        keys64[r] = (r < M) ? (((unsigned long long)fk[j] << 32) | (unsigned long long)(0xffffffffu - (uint32_t)r)) : 0ull;
    __syncthreads();
    bitonic_desc(keys64, Mpad);
    for (int r = tid; r < M; r += NT) {
        const unsigned long long k = keys64[r];
        const int j = (int)key_pos(k);
        p.out_pos[(int64_t)b * M + r] = mem_pos[j];
        p.out_src[(int64_t)b * M + r] = mem_src[j];
        if (p.out_score) p.out_score[(int64_t)b * M + r] = order_bits_inv((uint32_t)(k >> 32));
    }
}

template <int HT, int E>
int launch_reg(const LoopParams& p, int B, cudaStream_t st) {
    size_t smem = (size_t)1024 * E * 8 + (size_t)p.M * 20 + (256 + 2 + 64) * 4 + 34 * HT * 4 + 64;
    const size_t fin = (size_t)host_next_pow2(p.M) * 8;      // final 64-bit sort aliases the key / cand arrays
    IPSB_REQUIRE(fin <= (size_t)1024 * E * 8, "select_loop: M too large for the final sort");
    IPSB_REQUIRE(smem <= 227 * 1024, "select_loop: %zu bytes of shared memory needed", smem);
    auto kern = select_loop_reg_kernel<HT, E>;
    IPSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<B, 1024, smem, st>>>(p);
    IPSB_LAUNCH_CHECK();
    return 0;
}

// ---- cluster-parallel variant ------------------------------------------------------------------
// A thread-block cluster of NC CTAs owns one image: CTA r holds buffer positions [r*S, (r+1)*S) --
// their logits, scan positions and table rows -- in shared memory.  A pre-pass permutes the logit table
// into scan order, so each iteration streams its chunk with coalesced loads and no dependent gathers.
// Per iteration the CTAs exchange only small partials through distributed shared memory (per-(h,t) max and
// sum, 256-bin radix-select histograms, survivor counts); the M survivors (logits, position, row, score
// bits) move through a double-buffered global scratch that stays in L2.
// Same semantics as the single-CTA kernels: buffer kept in scan order, equal scores -> scanned first.
// histogram increment with warp aggregation: scores of one buffer share their leading bits, so most lanes hit the
// same bin -- one shared-memory atomic per distinct bin per warp instead of one per lane
__device__ __forceinline__ void hist_add(int* hist, uint32_t bin, bool active) {
    const uint32_t act = __ballot_sync(0xffffffffu, active);
    if (!active) return;
    const uint32_t peers = __match_any_sync(act, bin);
    if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&hist[bin], __popc(peers));
}

struct ClusterScratch {
    float part_max[kMaxHT], part_sum[kMaxHT];   // this CTA's partials (read remotely)
    float gmax[kMaxHT], gsum[kMaxHT];           // cluster-wide results
    int hist[4][256];                           // this CTA's radix histograms (read remotely)
    int tot[256];
    int cnt[2];                                 // survivors above / equal to the threshold in this CTA
    int sel[2];
    int wsum[64];
    Scratch red;
};

// ---- scan order drawn on the device (conf.scan_order_rng = 'device') ---------------------------------------------------
// A keyed bijection of [0, N): eight rounds of (xor constant, multiply by an odd constant mod 2^k, xor-shift right) on the
// k = ceil(log2 N) bit domain -- each step a bijection of the k-bit integers --, "cycle walking" (re-apply until the value
// is below N) to restrict it to [0, N).  The 2 x 64-bit key comes from the CUDA generator (a device tensor: no host work,
// capturable in a graph); round constants by splitmix64.  O(1) per element, no sort, no scratch.  The CPU restatement is
// oracle.keyed_scan_order; statistical checks (positions uniform over keys, chunk membership) are in tests/.
__host__ __device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) keyed_scan_order_kernel(const int64_t* __restrict__ key, int rows, int N, int k,
                                                               int64_t* __restrict__ perm) {
    __shared__ uint32_t mul[8], add[8];
    const int row = blockIdx.y;
    if (threadIdx.x == 0) {
        unsigned long long t = (unsigned long long)key[0] ^ splitmix64((unsigned long long)key[1] + (unsigned long long)row);
        for (int r = 0; r < 8; ++r) {
            t = splitmix64(t);
            mul[r] = (uint32_t)(t >> 32) | 1u;
            add[r] = (uint32_t)t;
        }
    }
    __syncthreads();
    const uint32_t mask = (k >= 32) ? 0xffffffffu : ((1u << k) - 1u);
    const int h1 = (k + 1) / 2, h2 = k / 3 + 1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        uint32_t x = (uint32_t)i;
        if (N > 1) {
            do {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    x ^= add[r] & mask;
                    x = (x * mul[r]) & mask;
                    x ^= x >> ((r & 1) ? h2 : h1);
                }
            } while (x >= (uint32_t)N);
        }
        perm[(int64_t)row * N + i] = (int64_t)x;
    }
}

// zs[b, pos, :] = z[b, perm[pos], :], srcs[b, pos] = perm[pos]
__global__ void permute_logits_kernel(const float* __restrict__ z, const int64_t* __restrict__ perm, int64_t perm_stride,
                                      int N, int HT, float* __restrict__ zs, int* __restrict__ srcs, int64_t total) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / HT;
        const int c = (int)(i - row * HT);
        const int64_t b = row / N;
        const int pos = (int)(row - b * N);
        const int src = perm ? (int)perm[b * perm_stride + pos] : pos;
        zs[i] = z[(b * N + src) * HT + c];
        if (c == 0) srcs[row] = src;
    }
}

struct ClusterArgs {
    const float* zs;      // (B, N, HT) logits in scan order
    const int* srcs;      // (B, N) table row of every scan position
    float* m_z;           // (B, 2, M, HT) survivors' logits
    int* m_pos;           // (B, 2, M)
    int* m_src;           // (B, 2, M)
    uint32_t* m_key;      // (B, M)
    unsigned long long* m_runs;   // (B, NC * ceil(M / NC)) sorted runs of the final ordering
    int slice_cap;
    // streamed form (ipsb_select_loop_scan): `zs` is the caller's table, already in scan order and possibly still being
    // written by the projector kernel running next to this one; srcs == null -> table row = perm[pos] (or pos)
    const int64_t* perm;          // (1 | B, N) scan order or null
    int64_t perm_stride;
    const int* tile_flags;        // [ceil(B N / tile_rows)]: != 0 once rows [t tile_rows, (t+1) tile_rows) of zs are written; or null
    int tile_rows;
    int* sync_words;              // [0]: set to 1 when cluster 0 is resident; [1]: != 0 if a wait ran out
};

// rows [g0, g1) of the flat (B N) scan-order table are needed: wait for their tiles (bounded: a producer that never
// comes must not hang the device; the status word tells)
__device__ __forceinline__ void wait_tiles(const ClusterArgs& a, int64_t g0, int64_t g1) {
    if (a.tile_flags == nullptr || g1 <= g0) return;
    const int t0 = (int)(g0 / a.tile_rows), t1 = (int)((g1 - 1) / a.tile_rows);
    for (int t = t0 + (int)threadIdx.x; t <= t1; t += (int)blockDim.x) {
        const int* f = a.tile_flags + t;
        int v = 0;
        for (long long spin = 0; spin < (1ll << 21); ++spin) {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if (v != 0) break;
            __nanosleep(64);
        }
        if (v == 0 && a.sync_words) atomicExch(a.sync_words + 1, 1);
    }
    __syncthreads();
}

template <int NC>
__global__ void __cluster_dims__(NC, 1, 1) __launch_bounds__(512, 1)
select_loop_cluster_kernel(LoopParams p, ClusterArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    constexpr int NT = 512;
    const int cr = (NC == 1) ? 0 : (int)cluster.block_rank();
    const int b = blockIdx.x / NC, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int HT = p.H * p.T, M = p.M, cap = a.slice_cap;
    float* zl = reinterpret_cast<float*>(smem_raw);                     // [cap][HT] logits of this slice
    uint32_t* key = reinterpret_cast<uint32_t*>(zl + (size_t)cap * HT); // [cap]
    int* cand = reinterpret_cast<int*>(key + cap);                      // [cap] table rows
    int* posl = cand + cap;                                             // [cap] scan positions
    size_t front = (size_t)cap * (HT + 3) * 4;
    { const size_t fs = (size_t)next_pow2((M + NC - 1) / NC) * 8; if (front < fs) front = fs; front = (front + 15) / 16 * 16; }
    ClusterScratch* cs = reinterpret_cast<ClusterScratch*>(smem_raw + front);
    uint32_t* allkey = reinterpret_cast<uint32_t*>(cs + 1);             // [M] final ordering

    const float* zs = a.zs + (int64_t)b * p.N * HT;
    const int* srcs = a.srcs ? a.srcs + (int64_t)b * p.N : nullptr;
    const int64_t* perm = a.perm ? a.perm + (int64_t)b * a.perm_stride : nullptr;
    auto src_of = [&](int pos) { return srcs ? srcs[pos] : (perm ? (int)perm[pos] : pos); };
    if (a.sync_words && blockIdx.x == 0 && tid == 0) {       // tells the host-side gate that this kernel holds its SMs
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(a.sync_words), "r"(1) : "memory");
    }
    float* z_buf[2] = {a.m_z + (int64_t)b * 2 * M * HT, a.m_z + (int64_t)b * 2 * M * HT + (int64_t)M * HT};
    int* pos_buf[2] = {a.m_pos + (int64_t)b * 2 * M, a.m_pos + (int64_t)b * 2 * M + M};
    int* src_buf[2] = {a.m_src + (int64_t)b * 2 * M, a.m_src + (int64_t)b * 2 * M + M};
    uint32_t* key_buf = a.m_key + (int64_t)b * M;
    auto csync = [&]() { if (NC == 1) __syncthreads(); else cluster.sync(); };

    // initial memory buffer: the first M scan positions
    wait_tiles(a, (int64_t)b * p.N, (int64_t)b * p.N + M);
    for (int i = cr * NT + tid; i < M * HT; i += NC * NT) z_buf[0][i] = __ldcg(zs + i);
    for (int r = cr * NT + tid; r < M; r += NC * NT) { pos_buf[0][r] = r; src_buf[0][r] = src_of(r); }
    __threadfence();
    csync();

    const bool pow2 = (HT & (HT - 1)) == 0;
    const int nt_eff = (NT / HT) * HT;
    const int n_iter = (p.N - M + p.I - 1) / p.I;
    int cur = 0;
    for (int it = 0; it < n_iter; ++it) {
        const int lo = M + it * p.I;
        const int hi = min(lo + p.I, p.N);
        const int L = M + (hi - lo);
        const int S = (L + NC - 1) / NC;                     // slice length
        const int l0 = cr * S;
        const int n_own = max(0, min(S, L - l0));
        // ---- this slice's logits / positions / rows -> shared memory (coalesced 128-bit streams, 4 loads in flight;
        //      L2-coherent loads: the survivors were written by the other CTAs of the cluster one iteration ago and the
        //      new rows possibly by the projector kernel a moment ago)
        wait_tiles(a, (int64_t)b * p.N + lo, (int64_t)b * p.N + hi);
        {
            const int hv = HT >> 2;                              // float4 granules per row (HT is a power of two >= 4) or 0
            if (hv > 0) {
                const int total4 = n_own * hv;
                float4* zl4 = reinterpret_cast<float4*>(zl);
                for (int e0 = tid; e0 < total4; e0 += 4 * NT) {
                    float4 v4[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int e = e0 + u * NT;
                        if (e < total4) {
                            const int i = e / hv, c4 = e - i * hv, l = l0 + i;
                            const float* src_row = (l < M) ? z_buf[cur] + (int64_t)l * HT : zs + (int64_t)(lo + l - M) * HT;
                            v4[u] = __ldcg(reinterpret_cast<const float4*>(src_row) + c4);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int e = e0 + u * NT;
                        if (e < total4) zl4[e] = v4[u];
                    }
                }
            } else {
                for (int e = tid; e < n_own * HT; e += NT) {
                    const int i = e / HT, c = e - i * HT, l = l0 + i;
                    zl[e] = __ldcg((l < M) ? z_buf[cur] + (int64_t)l * HT + c : zs + (int64_t)(lo + l - M) * HT + c);
                }
            }
        }
        for (int i = tid; i < n_own; i += NT) {
            const int l = l0 + i;
            if (l < M) { posl[i] = __ldcg(pos_buf[cur] + l); cand[i] = __ldcg(src_buf[cur] + l); }
            else { posl[i] = lo + (l - M); cand[i] = src_of(lo + (l - M)); }
        }
        if (tid < 256) { cs->hist[0][tid] = 0; cs->hist[1][tid] = 0; cs->hist[2][tid] = 0; cs->hist[3][tid] = 0; }
        __syncthreads();
        // ---- per-(h,t) max over the whole buffer
        const int ht = tid % HT;
        float v = -INFINITY;
        if (tid < nt_eff)
            for (int e = tid; e < n_own * HT; e += nt_eff) v = fmaxf(v, zl[e]);
        reduce_classes<true>(v, HT, pow2, nt_eff, &cs->red, cs->part_max);
        csync();
        if (tid < HT) {
            float m = cs->part_max[tid];
            if (NC > 1) for (int r = 0; r < NC; ++r) m = fmaxf(m, cluster.map_shared_rank(cs->part_max, r)[tid]);
            cs->gmax[tid] = m;
        }
        __syncthreads();
        // ---- per-(h,t) sum of exp
        v = 0.f;
        if (tid < nt_eff) {
            const float m = cs->gmax[ht];
            for (int e = tid; e < n_own * HT; e += nt_eff) v += expf(zl[e] - m);
        }
        reduce_classes<false>(v, HT, pow2, nt_eff, &cs->red, cs->part_sum);
        csync();
        if (tid < HT) {
            float sm = 0.f;
            if (NC > 1) { for (int r = 0; r < NC; ++r) sm += cluster.map_shared_rank(cs->part_sum, r)[tid]; }
            else sm = cs->part_sum[tid];
            cs->gsum[tid] = sm;
        }
        __syncthreads();
        // ---- scores -> order bits; first radix histogram on the fly
        for (int i0 = 0; i0 < n_own; i0 += NT) {             // warp-uniform trip count (hist_add uses warp votes)
            const int i = i0 + tid;
            uint32_t k = 0;
            if (i < n_own) {
                const float* zr = zl + (size_t)i * HT;
                float tok = 0.f;
                for (int t = 0; t < p.T; ++t) {
                    float hs = 0.f;
                    for (int h = 0; h < p.H; ++h) {
                        const int c = h * p.T + t;
                        hs += expf(zr[c] - cs->gmax[c]) / cs->gsum[c];
                    }
                    tok += hs / (float)p.H;
                }
                k = order_bits(tok / (float)p.T);
                key[i] = k;
            }
            hist_add(cs->hist[0], k >> 24, i < n_own);
        }
        // ---- radix select of the rank-M key: 4 passes of 8 bits, histograms summed over the cluster
        uint32_t prefix = 0;
        int remaining = M;
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            if (pass > 0) {
                for (int i0 = 0; i0 < n_own; i0 += NT) {
                    const int i = i0 + tid;
                    const uint32_t k = (i < n_own) ? key[i] : 0u;
                    hist_add(cs->hist[pass], (k >> shift) & 255u, i < n_own && (k >> (shift + 8)) == prefix);
                }
            }
            csync();
            if (tid < 256) {
                int t = 0;
                if (NC > 1) { for (int r = 0; r < NC; ++r) t += cluster.map_shared_rank(&cs->hist[pass][0], r)[tid]; }
                else t = cs->hist[pass][tid];
                cs->tot[tid] = t;
            }
            __syncthreads();
            if (warp == 0) {
                int c[8], tot = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { c[j] = cs->tot[255 - 8 * lane - j]; tot += c[j]; }
                int inc = tot;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
                const int before = inc - tot;
                if (before < remaining && remaining <= inc) {
                    int run = before;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (run < remaining && remaining <= run + c[j]) { cs->sel[0] = 255 - 8 * lane - j; cs->sel[1] = remaining - run; }
                        run += c[j];
                    }
                }
            }
            __syncthreads();
            prefix = (prefix << 8) | (uint32_t)cs->sel[0];
            remaining = cs->sel[1];
            __syncthreads();
        }
        const uint32_t thr = prefix;
        // ---- stable compaction in scan order: contiguous runs of E slice entries per thread
        const int E = (S + NT - 1) / NT;
        int ngt = 0, neq = 0;
        for (int e = 0; e < E; ++e) {
            const int i = tid * E + e;
            if (i < n_own) { const uint32_t k = key[i]; ngt += (k > thr); neq += (k == thr); }
        }
        int inc_gt = ngt, inc_eq = neq;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc_gt, o), w = __shfl_up_sync(0xffffffffu, inc_eq, o);
            if (lane >= o) { inc_gt += u; inc_eq += w; }
        }
        if (lane == 31) { cs->wsum[warp] = inc_gt; cs->wsum[32 + warp] = inc_eq; }
        __syncthreads();
        if (warp == 0) {
            int g1 = lane < NT / 32 ? cs->wsum[lane] : 0, c2 = lane < NT / 32 ? cs->wsum[32 + lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, g1, o), w = __shfl_up_sync(0xffffffffu, c2, o);
                if (lane >= o) { g1 += u; c2 += w; }
            }
            cs->wsum[lane] = g1; cs->wsum[32 + lane] = c2;
            if (lane == NT / 32 - 1) { cs->cnt[0] = g1; cs->cnt[1] = c2; }
        }
        csync();
        int gt_before = (inc_gt - ngt) + (warp ? cs->wsum[warp - 1] : 0);
        int eq_before = (inc_eq - neq) + (warp ? cs->wsum[32 + warp - 1] : 0);
        if (NC > 1) {
            for (int r = 0; r < cr; ++r) {                   // survivors in the slices before this one
                const int* rc = cluster.map_shared_rank(cs->cnt, r);
                gt_before += rc[0];
                eq_before += rc[1];
            }
        }
        const int nxt = cur ^ 1;
        for (int e = 0; e < E; ++e) {
            const int i = tid * E + e;
            if (i < n_own) {
                const uint32_t k = key[i];
                if ((k > thr) || (k == thr && eq_before < remaining)) {
                    const int dst = gt_before + (eq_before < remaining ? eq_before : remaining);
                    pos_buf[nxt][dst] = posl[i];
                    src_buf[nxt][dst] = cand[i];
                    key_buf[dst] = k;
                    posl[i] = dst;                           // remembered for the cooperative logit copy below
                } else {
                    posl[i] = -1;
                }
                gt_before += (k > thr);
                eq_before += (k == thr);
            }
        }
        __syncthreads();
        for (int e = tid; e < n_own * HT; e += NT) {         // survivors' logits -> next buffer, coalesced per row
            const int i = e / HT, dst = posl[i];
            if (dst >= 0) z_buf[nxt][(int64_t)dst * HT + (e - i * HT)] = zl[e];
        }
        __threadfence();
        csync();                                             // new buffer visible to the whole cluster
        cur = nxt;
    }
    // ---- final ordering (score descending, ties -> scanned first): every CTA sorts its share of the survivors
    //      (bitonic, shared memory), the sorted runs are exchanged through the scratch, and each element's rank is
    //      its place in its own run plus, by binary search, the number of larger elements in the other runs
    {
        const int Sm = (M + NC - 1) / NC;
        const int j0 = cr * Sm;
        const int n_mine = max(0, min(Sm, M - j0));
        const int pad = next_pow2(Sm);
        unsigned long long* run = reinterpret_cast<unsigned long long*>(smem_raw);           // [pad] (aliases zl: free now)
        for (int r = tid; r < pad; r += NT)
            run[r] = (r < n_mine) ? (((unsigned long long)__ldcg(key_buf + j0 + r) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)(j0 + r))) : 0ull;
        __syncthreads();
        bitonic_desc(run, pad);
        unsigned long long* g_runs = a.m_runs + (int64_t)b * NC * Sm;
        for (int r = tid; r < n_mine; r += NT) g_runs[(int64_t)cr * Sm + r] = run[r];
        __threadfence();
        csync();
        unsigned long long* all = reinterpret_cast<unsigned long long*>(allkey);             // [NC][Sm]
        if (NC > 1) {
            for (int r = tid; r < NC * Sm; r += NT) {
                const int rr = r / Sm, k = r - rr * Sm;
                all[r] = (k < max(0, min(Sm, M - rr * Sm))) ? g_runs[r] : 0ull;
            }
            __syncthreads();
        }
        for (int r = tid; r < n_mine; r += NT) {
            const unsigned long long K = run[r];
            int rank = r;
            if (NC > 1) {
                for (int rr = 0; rr < NC; ++rr) {
                    if (rr == cr) continue;
                    const unsigned long long* o = all + rr * Sm;
                    const int n_o = max(0, min(Sm, M - rr * Sm));
                    int lo_ = 0, hi_ = n_o;                      // first index whose element is < K (run is descending)
                    while (lo_ < hi_) {
                        const int mid = (lo_ + hi_) >> 1;
                        if (o[mid] > K) lo_ = mid + 1; else hi_ = mid;
                    }
                    rank += lo_;
                }
            }
            const int j = (int)key_pos(K);
            p.out_pos[(int64_t)b * M + rank] = __ldcg(pos_buf[cur] + j);
            p.out_src[(int64_t)b * M + rank] = __ldcg(src_buf[cur] + j);
            if (p.out_score) p.out_score[(int64_t)b * M + rank] = order_bits_inv((uint32_t)(K >> 32));
        }
    }
}

// ---- tiny buffers (M + I <= 128): one CTA of 128 threads per image, the whole loop in shared memory -------------------
// The traffic configuration (M = 10, I = 32) and the small-M corner of the selection sweep run hundreds to thousands of
// iterations on a buffer of a few dozen entries: a pure latency chain in which the radix select (4 histogram passes with
// shared-memory atomics, ~12 block barriers) and the round trip of the survivors through global memory dominate.  Here
//   * the memory rows (logits, position, table row) stay in shared memory, double buffered across iterations;
//   * the softmax statistics are one online (max, sum) pair per thread, merged through shuffles and one shared-memory step;
//   * the top-M is found by RANK COUNTING: rank_i = #{j : key_j > key_i} + #{j < i : key_j == key_i} (keys read as 128-bit
//     broadcasts) -- no histogram, no atomics, and the rank IS the position in the final best-first order, so the last
//     iteration needs no sort;
//   * survivors are compacted in scan order (the library's tie-break contract) with one block scan.
// 8 block barriers per iteration, no global-memory dependency between iterations.  Measured (B200, H*T = 8): 3.3 us per
// iteration at M = I = 10 (cluster<1> kernel: 5.9), 3.7 at M = 10 / I = 32 (6.0), 4.7 at M = I = 50 (6.3).  The pair
// count grows with L^2: at L = 200 it only ties the radix kernel (7.7 vs 7.2 us) and at L = 1000 it is 4x slower (57 vs
// 14 us), so longer buffers stay on select_loop_cluster_kernel<1>.
__device__ __forceinline__ void merge_ms(float& m, float& s, float m2, float s2) {
    const float M = fmaxf(m, m2);
    const float a = (m == -INFINITY) ? 0.f : s * expf(m - M);
    const float b = (m2 == -INFINITY) ? 0.f : s2 * expf(m2 - M);
    m = M;
    s = a + b;
}

struct SmallScratch {
    float wpart[16][kMaxHT][2];
    float gmax[kMaxHT], gsum[kMaxHT];
    int wsum[32];
};

template <int NT>
__global__ void __launch_bounds__(NT, 1) select_loop_small_kernel(LoopParams p, const float* __restrict__ zs_all,
                                                                  const int* __restrict__ srcs_all, int Lcap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int HT = p.H * p.T, M = p.M;
    float* zl[2];
    zl[0] = reinterpret_cast<float*>(smem_raw);                          // [Lcap][HT]
    zl[1] = zl[0] + (size_t)Lcap * HT;
    int* posb[2]; int* srcb[2];
    posb[0] = reinterpret_cast<int*>(zl[1] + (size_t)Lcap * HT);         // [Lcap] scan positions
    posb[1] = posb[0] + Lcap;
    srcb[0] = posb[1] + Lcap;                                            // [Lcap] table rows
    srcb[1] = srcb[0] + Lcap;
    uint32_t* key = reinterpret_cast<uint32_t*>(srcb[1] + Lcap);         // [Lcap + 4] order bits of the scores (padded to x4)
    int* dstmap = reinterpret_cast<int*>(key + Lcap + 4);                // [Lcap] survivor's row in the next buffer, or -1
    SmallScratch* sc = reinterpret_cast<SmallScratch*>(dstmap + Lcap);

    const float* zs = zs_all + (int64_t)b * p.N * HT;
    const int* srcs = srcs_all + (int64_t)b * p.N;
    // initial memory buffer: the first M scan positions
    for (int e = tid; e < M * HT; e += NT) zl[0][e] = zs[e];
    for (int r = tid; r < M; r += NT) { posb[0][r] = r; srcb[0][r] = srcs[r]; }

    const int n_iter = (p.N - M + p.I - 1) / p.I;
    int cur = 0;
    for (int it = 0; it < n_iter; ++it) {
        const int lo = M + it * p.I;
        const int hi = min(lo + p.I, p.N);
        const int n_new = hi - lo, L = M + n_new;
        const bool last = (it == n_iter - 1);
        float* z = zl[cur];
        // ---- (a) the next chunk: contiguous rows of the scan-ordered table
        for (int e = tid; e < n_new * HT; e += NT) z[M * HT + e] = __ldg(zs + (int64_t)lo * HT + e);
        for (int r = tid; r < n_new; r += NT) { posb[cur][M + r] = lo + r; srcb[cur][M + r] = __ldg(srcs + lo + r); }
        __syncthreads();
        // ---- (b) per-(h,t) max and sum of exp over the buffer: online (max, sum) pairs, fixed merge order
        {
            float m = -INFINITY, sm = 0.f;
            for (int e = tid; e < L * HT; e += NT) m = fmaxf(m, z[e]);   // NT % HT == 0: a thread stays in one class
            for (int e = tid; e < L * HT; e += NT) sm += expf(z[e] - m);
            for (int o = 16; o >= HT; o >>= 1) {
                const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, sm, o);
                merge_ms(m, sm, m2, s2);
            }
            if (lane < HT) { sc->wpart[warp][lane][0] = m; sc->wpart[warp][lane][1] = sm; }
            __syncthreads();
            if (tid < HT) {
                float mm = sc->wpart[0][tid][0], ss = sc->wpart[0][tid][1];
                for (int w = 1; w < NT / 32; ++w) merge_ms(mm, ss, sc->wpart[w][tid][0], sc->wpart[w][tid][1]);
                sc->gmax[tid] = mm; sc->gsum[tid] = ss;
            }
            __syncthreads();
        }
        // ---- (c) scores -> order bits (thread t owns the contiguous candidates [t*E, t*E + E))
        const int E = (L + NT - 1) / NT;
        const int i0 = tid * E;
        for (int e = 0; e < E; ++e) {
            const int i = i0 + e;
            if (i < L) {
                const float* zr = z + (size_t)i * HT;
                float tok = 0.f;
                for (int t = 0; t < p.T; ++t) {
                    float hs = 0.f;
                    for (int h = 0; h < p.H; ++h) {
                        const int c = h * p.T + t;
                        hs += expf(zr[c] - sc->gmax[c]) / sc->gsum[c];
                    }
                    tok += hs / (float)p.H;
                }
                key[i] = order_bits(tok / (float)p.T);
            }
        }
        if (tid < 4) key[L + tid] = 0u;                                  // padding read by the 128-bit rank loop (never larger)
        __syncthreads();
        // ---- (d) rank counting + compaction offsets
        int kept = 0;
        int rank_e[2] = {M, M};                                          // E = 1 (L <= 128 with 128 threads); generic up to 2
        for (int e = 0; e < E; ++e) {
            const int i = i0 + e;
            if (i < L) {
                const uint32_t k = key[i];
                int gt = 0, eqb = 0;
                const uint4* k4 = reinterpret_cast<const uint4*>(key);
                for (int j4 = 0; j4 * 4 < L; ++j4) {
                    const uint4 q = k4[j4];
                    const int j = j4 * 4;
                    gt += (q.x > k) + (q.y > k) + (q.z > k) + (q.w > k);
                    eqb += (q.x == k && j < i) + (q.y == k && j + 1 < i) + (q.z == k && j + 2 < i) + (q.w == k && j + 3 < i);
                }
                rank_e[e] = gt + eqb;
                kept += (rank_e[e] < M);
            }
        }
        int inc = kept;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
        if (lane == 31) sc->wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int v = lane < NT / 32 ? sc->wsum[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
            sc->wsum[lane] = v;
        }
        __syncthreads();
        int dst = (inc - kept) + (warp ? sc->wsum[warp - 1] : 0);        // survivors scanned before this thread's first candidate
        const int nxt = cur ^ 1;
        for (int e = 0; e < E; ++e) {
            const int i = i0 + e;
            if (i >= L) break;
            const int rank = rank_e[e];
            if (rank < M) {
                if (last) {                                              // best first: the rank is the output slot
                    p.out_pos[(int64_t)b * M + rank] = posb[cur][i];
                    p.out_src[(int64_t)b * M + rank] = srcb[cur][i];
                    if (p.out_score) p.out_score[(int64_t)b * M + rank] = order_bits_inv(key[i]);
                }
                posb[nxt][dst] = posb[cur][i];
                srcb[nxt][dst] = srcb[cur][i];
                dstmap[i] = dst++;
            } else {
                dstmap[i] = -1;
            }
        }
        if (last) break;
        __syncthreads();
        // ---- (e) survivors' logits -> the other buffer, rows in scan order (coalesced, conflict free)
        float* zn = zl[nxt];
        for (int e = tid; e < L * HT; e += NT) {
            const int i = e / HT, d = dstmap[i];
            if (d >= 0) zn[d * HT + (e - i * HT)] = z[e];
        }
        __syncthreads();
        cur = nxt;
    }
}

// table already in scan order (ipsb_select_loop_scan): no permuted copy; optional tile flags / residency word
struct ScanArgs {
    const int* tile_flags;
    int tile_rows;
    int* sync_words;
};

template <int NC>
int launch_cluster(const LoopParams& p, int B, int Lmax, void* workspace, int64_t workspace_bytes, cudaStream_t st,
                   const ScanArgs* scan = nullptr) {
    const int HT = p.H * p.T, M = p.M, N = p.N;
    const int cap = (Lmax + NC - 1) / NC;
    const int Sm = (M + NC - 1) / NC;
    size_t front = (size_t)cap * (HT + 3) * 4;
    if (front < (size_t)host_next_pow2(Sm) * 8) front = (size_t)host_next_pow2(Sm) * 8;     // final sort aliases the slice arrays
    front = (front + 15) / 16 * 16;
    const size_t smem = front + sizeof(ClusterScratch) + (size_t)NC * Sm * 8 + 64;
    if (smem > 200 * 1024) return -1;                        // caller falls back
    const int64_t need = ipsb_select_loop_workspace_bytes(B, N, HT, M);
    IPSB_REQUIRE(workspace != nullptr && workspace_bytes >= need, "select_loop: workspace of %lld bytes required", (long long)need);
    char* ws = (char*)workspace;
    ClusterArgs a;
    auto take = [&](size_t bytes) { char* q = ws; ws += (bytes + 255) / 256 * 256; return q; };   // 256-byte aligned pieces
    float* zs = (float*)take((size_t)B * N * HT * 4);
    int* srcs = (int*)take((size_t)B * N * 4);
    a.m_z = (float*)take((size_t)B * 2 * M * HT * 4);
    a.m_pos = (int*)take((size_t)B * 2 * M * 4);
    a.m_src = (int*)take((size_t)B * 2 * M * 4);
    a.m_key = (uint32_t*)take((size_t)B * M * 4);
    a.m_runs = (unsigned long long*)take((size_t)B * (M + 16) * 8);
    a.zs = zs; a.srcs = srcs; a.slice_cap = cap;
    a.perm = nullptr; a.perm_stride = 0; a.tile_flags = nullptr; a.tile_rows = 1; a.sync_words = nullptr;
    if (scan) {
        a.zs = p.z; a.srcs = nullptr; a.perm = p.perm; a.perm_stride = p.perm_stride;
        a.tile_flags = scan->tile_flags; a.tile_rows = scan->tile_rows; a.sync_words = scan->sync_words;
    } else {
        const int64_t total = (int64_t)B * N * HT;
        int64_t g = (total + 255) / 256;
        if (g > (int64_t)ipsb::sm_count() * 8) g = (int64_t)ipsb::sm_count() * 8;
        permute_logits_kernel<<<(unsigned)g, 256, 0, st>>>(p.z, p.perm, p.perm_stride, N, HT, zs, srcs, total);
        IPSB_LAUNCH_CHECK();
    }
    if (NC == 1 && !scan && Lmax <= 128 && getenv("IPSB_SELECT_NO_SMALL") == nullptr) {       // tiny buffers: shared-memory resident loop
        const int Lcap = (Lmax + 3) / 4 * 4;
        const size_t sm_small = (size_t)Lcap * HT * 8 + (size_t)Lcap * 4 * 4 + ((size_t)Lcap + 4) * 4 + (size_t)Lcap * 4 +
                                sizeof(SmallScratch) + 64;
        auto ks = select_loop_small_kernel<128>;
        IPSB_CUDA(cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_small));
        ks<<<B, 128, sm_small, st>>>(p, zs, srcs, Lcap);
        IPSB_LAUNCH_CHECK();
        return 0;
    }
    auto kern = select_loop_cluster_kernel<NC>;
    IPSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (NC > 8) {                                            // 16 CTAs per cluster: beyond the portable size, one cluster per GPC
        IPSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(B * NC)); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = NC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n_clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&n_clusters, kern, &cfg) != cudaSuccess || n_clusters < 1) {
            (void)cudaGetLastError();
            return -1;                                       // caller falls back
        }
    }
    kern<<<B * NC, 512, smem, st>>>(p, a);
    IPSB_LAUNCH_CHECK();
    return 0;
}

// ---- standalone pieces (unit-level parity: P1 / P2) ----------------------------------

__global__ void __launch_bounds__(1024) scores_kernel(const float* z, float* scores, int L, int H, int T) {
    __shared__ Scratch sc;
    const int HT = H * T;
    const float* zb = z + (int64_t)blockIdx.x * L * HT;
    auto row_of = [&](int l) { return zb + (int64_t)l * HT; };
    score_buffer(row_of, L, H, T, &sc, nullptr, scores + (int64_t)blockIdx.x * L);
}

__global__ void __launch_bounds__(1024) topm_kernel(const float* scores, int L, int M, int64_t* idx_out, float* val_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
    const int b = blockIdx.x, tid = threadIdx.x, nthreads = blockDim.x;
    const int Lpad = next_pow2(L);
    for (int l = tid; l < Lpad; l += nthreads)
        keys[l] = (l < L) ? make_key(scores[(int64_t)b * L + l], (uint32_t)l) : 0ull;
    __syncthreads();
    bitonic_desc(keys, Lpad);
    for (int r = tid; r < M; r += nthreads) {
        const unsigned long long k = keys[r];
        idx_out[(int64_t)b * M + r] = key_pos(k);
        if (val_out) val_out[(int64_t)b * M + r] = order_bits_inv((uint32_t)(k >> 32));
    }
}

// Stable top-M of LONG rows (candidate merge of the sequence-sharded schedule: L = ranks * M): radix select of the
// rank-M score straight from global memory (coalesced, L2 resident), compaction of the winners (everything above the
// threshold plus the first-scanned equals), one bitonic sort of the M winners by (score desc, position asc).
__global__ void __launch_bounds__(1024, 1) topm_big_kernel(const float* scores, int L, int M, int64_t* idx_out, float* val_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NT = 1024;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Mpad = next_pow2(M);
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);      // [Mpad]
    int* hist = reinterpret_cast<int*>(keys + Mpad);                                   // [256]
    int* sel = hist + 256;                                                             // [2]
    int* wsum = sel + 2;                                                               // [64]
    const float* s = scores + (int64_t)b * L;
    uint32_t prefix = 0;
    int remaining = M;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        for (int l0 = 0; l0 < L; l0 += NT) {                  // warp-uniform trip count (hist_add votes)
            const int l = l0 + tid;
            const uint32_t k = (l < L) ? order_bits(__ldg(s + l)) : 0u;
            hist_add(hist, (k >> shift) & 255u, l < L && (pass == 0 || (k >> (shift + 8)) == prefix));
        }
        __syncthreads();
        if (warp == 0) {                                      // lane j owns bins [255-8j-7, 255-8j]: scan from the top
            int c[8], tot = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { c[j] = hist[255 - 8 * lane - j]; tot += c[j]; }
            int inc = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
            const int before = inc - tot;
            if (before < remaining && remaining <= inc) {
                int run = before;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (run < remaining && remaining <= run + c[j]) { sel[0] = 255 - 8 * lane - j; sel[1] = remaining - run; }
                    run += c[j];
                }
            }
        }
        __syncthreads();
        prefix = (prefix << 8) | (uint32_t)sel[0];
        remaining = sel[1];
        __syncthreads();
    }
    const uint32_t thr = prefix;                              // key of rank M; the first `remaining` equals are kept
    // each warp owns a contiguous range of the row and walks it 32 entries at a time
    const int per = (((L + 31) / 32) + 31) / 32 * 32;
    const int lo = warp * per, hi = min(L, lo + per);
    int ngt = 0, neq = 0;
    for (int l0 = lo; l0 < hi; l0 += 32) {
        const int l = l0 + lane;
        const uint32_t k = (l < hi) ? order_bits(__ldg(s + l)) : 0u;
        ngt += __popc(__ballot_sync(0xffffffffu, l < hi && k > thr));
        neq += __popc(__ballot_sync(0xffffffffu, l < hi && k == thr));
    }
    if (lane == 0) { wsum[warp] = ngt; wsum[32 + warp] = neq; }
    __syncthreads();
    if (warp == 0) {
        int a = wsum[lane], c2 = wsum[32 + lane];
        const int a0 = a, c0 = c2;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, a, o), v = __shfl_up_sync(0xffffffffu, c2, o);
            if (lane >= o) { a += u; c2 += v; }
        }
        wsum[lane] = a - a0; wsum[32 + lane] = c2 - c0;      // exclusive
    }
    for (int r = M + tid; r < Mpad; r += NT) keys[r] = 0ull;
    __syncthreads();
    int gt_before = wsum[warp], eq_before = wsum[32 + warp];
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (int l0 = lo; l0 < hi; l0 += 32) {
        const int l = l0 + lane;
        const float sv = (l < hi) ? __ldg(s + l) : 0.f;
        const uint32_t k = (l < hi) ? order_bits(sv) : 0u;
        const bool gt = l < hi && k > thr, eq = l < hi && k == thr;
        const uint32_t bg = __ballot_sync(0xffffffffu, gt), be = __ballot_sync(0xffffffffu, eq);
        const int my_gt = gt_before + __popc(bg & lt_mask), my_eq = eq_before + __popc(be & lt_mask);
        if (gt || (eq && my_eq < remaining))
            keys[my_gt + min(my_eq, remaining)] = make_key(sv, (uint32_t)l);
        gt_before += __popc(bg);
        eq_before += __popc(be);
    }
    __syncthreads();
    bitonic_desc(keys, Mpad);
    for (int r = tid; r < M; r += NT) {
        const unsigned long long k = keys[r];
        idx_out[(int64_t)b * M + r] = key_pos(k);
        if (val_out) val_out[(int64_t)b * M + r] = order_bits_inv((uint32_t)(k >> 32));
    }
}

int host_next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace

extern "C" {

int ipsb_scores_from_logits(const float* z, float* scores, int B, int L, int H, int T, void* stream) {
    IPSB_REQUIRE(B > 0 && L > 0 && H > 0 && T > 0, "scores: bad shape B=%d L=%d H=%d T=%d", B, L, H, T);
    IPSB_REQUIRE(H * T <= kMaxHT, "scores: H*T=%d exceeds %d", H * T, kMaxHT);
    const int threads = L * H * T <= 4096 ? 256 : 1024;
    scores_kernel<<<B, threads, 0, (cudaStream_t)stream>>>(z, scores, L, H, T);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_topm_stable(const float* scores, int B, int L, int M, int64_t* idx_out, float* val_out, void* stream) {
    IPSB_REQUIRE(B > 0 && L > 0 && M > 0 && M <= L, "topm: bad shape B=%d L=%d M=%d", B, L, M);
    const int Lpad = host_next_pow2(L);
    if (Lpad > 2048 && M <= 16384) {                          // long rows: radix select from global memory (a single-CTA bitonic
                                                              // sort of 16 384 keys costs > 100 us; this is ~3x faster at L = 10 000)
        const size_t smem_big = (size_t)host_next_pow2(M) * 8 + (256 + 2 + 64) * 4 + 16;
        IPSB_CUDA(cudaFuncSetAttribute(topm_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_big));
        topm_big_kernel<<<B, 1024, smem_big, (cudaStream_t)stream>>>(scores, L, M, idx_out, val_out);
        IPSB_LAUNCH_CHECK();
        return 0;
    }
    const size_t smem = (size_t)Lpad * 8;
    IPSB_CUDA(cudaFuncSetAttribute(topm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    topm_kernel<<<B, Lpad <= 512 ? 256 : 1024, smem, (cudaStream_t)stream>>>(scores, L, M, idx_out, val_out);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int64_t ipsb_select_loop_workspace_bytes(int B, int N, int HT, int M) {
    return (int64_t)B * ((int64_t)N * (HT + 1) + 2ll * M * (HT + 2) + M) * 4 + (int64_t)B * (M + 16) * 8 + 7 * 256;
}

int ipsb_select_loop(const float* z, const int64_t* perm, int64_t perm_batch_stride,
                     int B, int N, int H, int T, int M, int I,
                     int64_t* mem_pos, int64_t* mem_src, float* mem_score,
                     void* workspace, int64_t workspace_bytes, void* stream) {
    IPSB_REQUIRE(B > 0 && N > 0 && M > 0 && I > 0 && H > 0 && T > 0, "select_loop: bad shape");
    IPSB_REQUIRE(M < N, "select_loop: M=%d >= N=%d is the caller's shortcut (ips_net.py:185)", M, N);
    IPSB_REQUIRE(H * T <= kMaxHT, "select_loop: H*T=%d exceeds %d", H * T, kMaxHT);
    const int Lmax = M + (I < N - M ? I : N - M);
    LoopParams p{z, perm, perm_batch_stride, N, H, T, M, I, mem_pos, mem_src, mem_score};
    // shared-memory-resident loop: one CTA per image, or a cluster of 8 CTAs per image for long buffers
    const bool ht_pow2 = ((H * T) & (H * T - 1)) == 0;
    if (ht_pow2 && workspace != nullptr && getenv("IPSB_SELECT_SINGLE_CTA") == nullptr) {
        int rc = (Lmax >= 2048) ? launch_cluster<8>(p, B, Lmax, workspace, workspace_bytes, (cudaStream_t)stream)
                                : launch_cluster<1>(p, B, Lmax, workspace, workspace_bytes, (cudaStream_t)stream);
        if (rc >= 0) return rc;
        if (Lmax >= 2048) {      // slices too long for 8 CTAs' shared memory (candidate merge of 8 ranks: 40 000 entries; H*T = 32)
            rc = launch_cluster<16>(p, B, Lmax, workspace, workspace_bytes, (cudaStream_t)stream);
            if (rc >= 0) return rc;
        }
    }
    // single-CTA long-buffer variant (kept as a cross-check, IPSB_SELECT_SINGLE_CTA=1)
    if (H * T == 8 && Lmax > 2048 && M <= 8192 && N < 65536 * 1024) {
        if (Lmax <= 1024 * 4) return launch_reg<8, 4>(p, B, (cudaStream_t)stream);
        if (Lmax <= 1024 * 10 && (size_t)M * 8 <= 56 * 1024) return launch_reg<8, 10>(p, B, (cudaStream_t)stream);
    }
    const int Lpad = host_next_pow2(Lmax);
    IPSB_REQUIRE(Lpad <= kMaxLpad, "select_loop: M+I=%d exceeds the single-CTA limit %d", Lmax, kMaxLpad);
    const size_t smem = (size_t)Lpad * 8 + (size_t)(2 * M) * 4 + sizeof(Scratch) + 16;
    IPSB_REQUIRE(smem <= 227 * 1024, "select_loop: %zu bytes of shared memory needed (M=%d I=%d)", smem, M, I);
    IPSB_CUDA(cudaFuncSetAttribute(select_loop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    select_loop_kernel<<<B, Lpad <= 512 ? 256 : 1024, smem, (cudaStream_t)stream>>>(p);
    IPSB_LAUNCH_CHECK();
    return 0;
}

// Streamed form of ipsb_select_loop for the feature-bag path: `zs` (B, N, H*T) is ALREADY in scan order (the projector
// kernel gathers its rows through the same `perm`), so no permuted copy is made, and it may still be being written:
// tile_flags[t] != 0 <=> rows [t * tile_rows, (t + 1) * tile_rows) of the flat (B N) table are complete (set by
// ipsb_projector_logits_scan running on another stream); sync_words[0] is set when the loop's cluster is resident (gate
// for the producer's launch: ipsb_wait_word), sync_words[1] != 0 reports a wait that ran out.  tile_flags / sync_words may
// be null (table complete).  Returns -2 when the shape is outside the 8-CTA cluster kernel (caller uses ipsb_select_loop).
int ipsb_select_loop_scan(const float* zs, const int64_t* perm, int64_t perm_batch_stride,
                          int B, int N, int H, int T, int M, int I,
                          int64_t* mem_pos, int64_t* mem_src, float* mem_score,
                          void* workspace, int64_t workspace_bytes,
                          const int* tile_flags, int tile_rows, int* sync_words, void* stream) {
    IPSB_REQUIRE(B > 0 && N > 0 && M > 0 && I > 0 && H > 0 && T > 0 && M < N, "select_loop_scan: bad shape");
    IPSB_REQUIRE(H * T <= kMaxHT, "select_loop_scan: H*T=%d exceeds %d", H * T, kMaxHT);
    IPSB_REQUIRE(tile_flags == nullptr || tile_rows > 0, "select_loop_scan: tile_rows");
    const int Lmax = M + (I < N - M ? I : N - M);
    const bool ht_pow2 = ((H * T) & (H * T - 1)) == 0;
    if (!ht_pow2 || Lmax < 2048 || workspace == nullptr) return -2;
    LoopParams p{zs, perm, perm_batch_stride, N, H, T, M, I, mem_pos, mem_src, mem_score};
    ScanArgs sc{tile_flags, tile_rows, sync_words};
    const int rc = launch_cluster<8>(p, B, Lmax, workspace, workspace_bytes, (cudaStream_t)stream, &sc);
    return rc < 0 ? -2 : rc;
}

// one thread waits (bounded) until *word != 0: orders a launch behind "the other stream's kernel is resident"
__global__ void wait_word_kernel(const int* word) {
    int v = 0;
    for (long long spin = 0; spin < (1ll << 22); ++spin) {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(word) : "memory");
        if (v != 0) break;
        __nanosleep(32);
    }
}
// perm (rows, N) int64: row r = the keyed bijection of [0, N) for key (key[0], key[1], r); key = two int64 on the device
int ipsb_keyed_scan_order(const int64_t* key, int rows, int N, int64_t* perm, void* stream) {
    IPSB_REQUIRE(key && perm && rows > 0 && N > 0, "keyed_scan_order: bad argument");
    int k = 1;
    while (k < 31 && (1ll << k) < (long long)N) ++k;
    int gx = (N + 255) / 256;
    if (gx > 4 * ipsb::sm_count()) gx = 4 * ipsb::sm_count();
    keyed_scan_order_kernel<<<dim3((unsigned)gx, (unsigned)rows), 256, 0, (cudaStream_t)stream>>>(key, rows, N, k, perm);
    IPSB_LAUNCH_CHECK();
    return 0;
}

// CUDA loads kernels lazily, and the first load of a function can wait for running kernels: a loop already spinning on
// the producer's flags would then wait for a producer whose load waits for the loop.  Called once, before the first
// streamed selection, this loads every kernel of the pair while the device is idle.
int ipsb_streamed_preload(void) {
    cudaFuncAttributes fa;
    IPSB_CUDA(cudaFuncGetAttributes(&fa, wait_word_kernel));
    IPSB_CUDA(cudaFuncGetAttributes(&fa, select_loop_cluster_kernel<8>));
    return ipsb_projector_preload();
}
int ipsb_wait_word(const int* word, void* stream) {
    IPSB_REQUIRE(word != nullptr, "wait_word: null");
    wait_word_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(word);
    IPSB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
