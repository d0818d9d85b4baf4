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
#include "common.cuh"
#include "../../include/ips_b200.h"

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
            keys[r] = ((unsigned long long)(uint32_t)pos << 32) | (uint32_t)src;
        }
        __syncthreads();
        for (int r = tid; r < M; r += nthreads) {
            const unsigned long long k = keys[r];
            mem_pos[r] = (int)(k >> 32);
            mem_src[r] = (int)(k & 0xffffffffull);
        }
        __syncthreads();
    }
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
    IPSB_REQUIRE(Lpad <= kMaxLpad, "topm: L=%d exceeds the single-CTA limit %d", L, kMaxLpad);
    const size_t smem = (size_t)Lpad * 8;
    IPSB_CUDA(cudaFuncSetAttribute(topm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    topm_kernel<<<B, Lpad <= 512 ? 256 : 1024, smem, (cudaStream_t)stream>>>(scores, L, M, idx_out, val_out);
    IPSB_LAUNCH_CHECK();
    return 0;
}

int ipsb_select_loop(const float* z, const int64_t* perm, int64_t perm_batch_stride,
                     int B, int N, int H, int T, int M, int I,
                     int64_t* mem_pos, int64_t* mem_src, float* mem_score, void* stream) {
    IPSB_REQUIRE(B > 0 && N > 0 && M > 0 && I > 0 && H > 0 && T > 0, "select_loop: bad shape");
    IPSB_REQUIRE(M < N, "select_loop: M=%d >= N=%d is the caller's shortcut (ips_net.py:185)", M, N);
    IPSB_REQUIRE(H * T <= kMaxHT, "select_loop: H*T=%d exceeds %d", H * T, kMaxHT);
    const int Lmax = M + (I < N - M ? I : N - M);
    const int Lpad = host_next_pow2(Lmax);
    IPSB_REQUIRE(Lpad <= kMaxLpad, "select_loop: M+I=%d exceeds the single-CTA limit %d", Lmax, kMaxLpad);
    const size_t smem = (size_t)Lpad * 8 + (size_t)(2 * M) * 4 + sizeof(Scratch) + 16;
    IPSB_REQUIRE(smem <= 227 * 1024, "select_loop: %zu bytes of shared memory needed (M=%d I=%d)", smem, M, I);
    IPSB_CUDA(cudaFuncSetAttribute(select_loop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LoopParams p{z, perm, perm_batch_stride, N, H, T, M, I, mem_pos, mem_src, mem_score};
    select_loop_kernel<<<B, Lpad <= 512 ? 256 : 1024, smem, (cudaStream_t)stream>>>(p);
    IPSB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
