// Native executor for the eval-mode ResNet patch encoder: issues the whole layer sequence
// of one chunk (stage -> stem -> maxpool -> BasicBlocks -> avgpool -> logits) from C++, so an
// ips() call costs one library call instead of ~25 Python round trips per chunk.
#include <cstdlib>
#include "common.cuh"
#include "pf.cuh"
#include "../../include/ips_b200.h"

#include <vector>

namespace {

// ---- per-launch timing of the executor's kernels (bench.py roofline): CUDA events around every library call on the
//      lane it is issued on, all measured against one base event, so the host can take the union of the intervals of
//      a kernel family across the concurrently running lanes.  Off (zero cost) unless ipsb_profile_begin was called.
struct ProfRec { int kind; cudaEvent_t a, b; };
struct Profiler {
    bool on = false;
    cudaEvent_t base = nullptr;
    std::vector<ProfRec> recs;
};
Profiler& profiler() { static Profiler p; return p; }
template <class F>
int timed_call(int kind, void* stream, const F& f) {
    Profiler& pr = profiler();
    if (!pr.on) return f();
    ProfRec r{kind, nullptr, nullptr};
    IPSB_CUDA(cudaEventCreate(&r.a));
    IPSB_CUDA(cudaEventCreate(&r.b));
    IPSB_CUDA(cudaEventRecord(r.a, (cudaStream_t)stream));
    const int rc = f();
    IPSB_CUDA(cudaEventRecord(r.b, (cudaStream_t)stream));
    pr.recs.push_back(r);
    return rc;
}
#define PROF(kind, stream, call) timed_call(kind, stream, [&]() -> int { return (call); })
enum { K_STAGE = 0, K_STEM = 1, K_CONV = 2, K_POOL = 3, K_LOGITS = 4 };

struct Geo { int H, W, C; };

inline int out_dim(int x, int k, int s, int p) { return (x + 2 * p - k) / s + 1; }
inline size_t esize(int dt) { return dt == IPSB_BF16 ? 2 : 4; }
inline int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }

// largest activation (elements per patch) of each of the three rotating buffers
int64_t max_act_elems(const ipsb_resnet_desc* net, int H, int W, int64_t* staged_elems) {
    *staged_elems = (net->stem.mode >= 3) ? (int64_t)(H + 6) * (W + 6) * 4 : (int64_t)H * W * 4;
    int h = out_dim(H, 7, 2, 3), w = out_dim(W, 7, 2, 3);
    int64_t mx = (int64_t)h * w * net->stem.cout;
    h = out_dim(h, 3, 2, 1); w = out_dim(w, 3, 2, 1);
    for (int b = 0; b < net->n_blocks; ++b) {
        const ipsb_conv_desc& c1 = net->blocks[b].c1;
        h = out_dim(h, c1.kh, c1.stride, c1.pad); w = out_dim(w, c1.kw, c1.stride, c1.pad);
        const int64_t e = (int64_t)h * w * c1.cout;
        if (e > mx) mx = e;
    }
    return mx;
}

int run_conv(const ipsb_resnet_desc* net, const ipsb_conv_desc& c, const void* x, const void* res, void* y,
             int64_t P, int H, int W, int relu, void* stream) {
    if (net->dt == IPSB_BF16)
        return ipsb_conv_bf16_umma(x, c.w, c.scale, c.shift, res, y, P, H, W, c.cin, c.cout, c.kh, c.kw, c.stride,
                                   c.pad, relu, c.mode, stream);
    return ipsb_conv_f32((const float*)x, (const float*)c.w, c.scale, c.shift, (const float*)res, (float*)y, P, H, W,
                         c.cin, c.cout, c.kh, c.kw, c.stride, c.pad, relu, stream);
}

__global__ void iota_mod_kernel(int64_t* out, int64_t first, int64_t n, int64_t mod) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (first + i) % mod;
}

// ---- bf16 path: padded-flat activations after the max-pool (see pf.cuh) -------------------------
struct PfPlan {
    int n_groups;                 // resolution groups = layers of two BasicBlocks
    int H[4], W[4], C[4];         // geometry of each group's activations
    int64_t rows[4];              // PF rows per buffer for `chunk` patches
    int64_t staged_bytes, stem_bytes, group_bytes[4], total;
    // groups whose maps do not fill pixel-box tiles (7x7) and are 256 wide can run DENSE through the im2col-mode kernel:
    // they get their own four (P*H*W, C) buffers (0 bytes when the group is not a candidate), after the PF buffers
    int64_t dense_bytes[4];
};

PfPlan make_pf_plan(const ipsb_resnet_desc* net, int64_t chunk, int H, int W) {
    PfPlan pl;
    pl.staged_bytes = align256(chunk * (int64_t)(H + 6) * (W + 6) * 4 * 2);
    int h = out_dim(H, 7, 2, 3), w = out_dim(W, 7, 2, 3);
    const bool fused = net->stem.mode == 4 && net->stem.cout == 64 && !getenv("IPSB_STEM_UNFUSED");   // stem output stays on chip
    pl.stem_bytes = fused ? 0
                  : (net->stem.mode == 4) ? align256(chunk * (int64_t)(h + 3) * (w + 3) * net->stem.cout * 2)   // wide row order
                                          : align256(chunk * (int64_t)h * w * net->stem.cout * 2);
    h = out_dim(h, 3, 2, 1); w = out_dim(w, 3, 2, 1);
    pl.n_groups = net->n_blocks / 2;
    pl.total = pl.staged_bytes + pl.stem_bytes;
    for (int g = 0; g < pl.n_groups; ++g) {
        const ipsb_conv_desc& c1 = net->blocks[2 * g].c1;
        h = out_dim(h, c1.kh, c1.stride, c1.pad); w = out_dim(w, c1.kw, c1.stride, c1.pad);
        pl.H[g] = h; pl.W[g] = w; pl.C[g] = c1.cout;
        pl.rows[g] = ipsb_pf_rows(chunk, h, w);
        pl.group_bytes[g] = align256(pl.rows[g] * c1.cout * 2);
        pl.total += 4 * pl.group_bytes[g];
        const int hw = h * w;
        const bool candidate = g > 0 && c1.cout % 256 == 0 && (hw & (hw - 1)) != 0 && !getenv("IPSB_NO_IM2COL") && !getenv("IPSB_NO_PAIR");
        pl.dense_bytes[g] = candidate ? align256(chunk * (int64_t)hw * c1.cout * 2) : 0;
    }
    for (int g = 0; g < pl.n_groups; ++g) pl.total += 4 * pl.dense_bytes[g];      // (after every PF buffer: see resnet_logits_pf)
    return pl;
}

// Internal streams for running several chunks concurrently ("lanes").  Every kernel of the bf16 path is a persistent
// grid of one CTA per SM; with two chunks in flight the block scheduler fills the SMs a kernel's last partial wave of
// tiles leaves idle (and the HBM-bound staging kernel of one chunk overlaps the tensor-bound kernels of the other).
constexpr int MAX_LANES = 4;
struct LaneStreams {
    int device = -1;
    cudaStream_t s[MAX_LANES] = {};
    cudaEvent_t done[MAX_LANES] = {};
    cudaEvent_t fork = nullptr;
    // per lane: a side stream for the 1x1 downsample convolution of a stride-2 BasicBlock, which reads the same input as
    // the block's first 3x3 convolution and is needed only by the second one
    cudaStream_t aux[MAX_LANES] = {};
    cudaEvent_t aux_fork[MAX_LANES] = {}, aux_join[MAX_LANES] = {};
};
constexpr int MAX_DEVICES = 16;
int lane_streams(LaneStreams** out) {
    static thread_local LaneStreams per_device[MAX_DEVICES];      // one set per device: nothing leaks when the caller switches
    int dev = 0;
    IPSB_CUDA(cudaGetDevice(&dev));
    IPSB_REQUIRE(dev >= 0 && dev < MAX_DEVICES, "resnet_logits: device index %d", dev);
    LaneStreams& ls = per_device[dev];
    if (ls.device != dev) {
        for (int i = 0; i < MAX_LANES; ++i) {
            IPSB_CUDA(cudaStreamCreateWithFlags(&ls.s[i], cudaStreamNonBlocking));
            IPSB_CUDA(cudaEventCreateWithFlags(&ls.done[i], cudaEventDisableTiming));
            IPSB_CUDA(cudaStreamCreateWithFlags(&ls.aux[i], cudaStreamNonBlocking));
            IPSB_CUDA(cudaEventCreateWithFlags(&ls.aux_fork[i], cudaEventDisableTiming));
            IPSB_CUDA(cudaEventCreateWithFlags(&ls.aux_join[i], cudaEventDisableTiming));
        }
        IPSB_CUDA(cudaEventCreateWithFlags(&ls.fork, cudaEventDisableTiming));
        ls.device = dev;
    }
    *out = &ls;
    return 0;
}

int run_conv_pf(const ipsb_conv_desc& c, const void* x, const void* res, void* y, int64_t P, int H, int W, int relu,
                void* stream, int in_pf = 1, int out_pf = 1) {
    return ipsb_conv_bf16_pf(x, c.w, c.scale, c.shift, res, y, P, H, W, c.cin, c.cout, c.kh, c.kw, c.stride, c.pad, relu,
                             in_pf, out_pf, stream);
}

}  // namespace

extern "C" {

int ipsb_profile_begin(void* stream) {
    Profiler& pr = profiler();
    for (ProfRec& r : pr.recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    pr.recs.clear();
    if (!pr.base) IPSB_CUDA(cudaEventCreate(&pr.base));
    IPSB_CUDA(cudaEventRecord(pr.base, (cudaStream_t)stream));
    pr.on = true;
    return 0;
}

int ipsb_profile_end(int max_records, int* kinds, float* start_ms, float* stop_ms, int* n_out) {
    Profiler& pr = profiler();
    pr.on = false;
    IPSB_REQUIRE(n_out != nullptr, "profile_end: null output");
    int n = 0;
    for (ProfRec& r : pr.recs) {
        if (n < max_records && kinds && start_ms && stop_ms) {
            IPSB_CUDA(cudaEventSynchronize(r.b));
            kinds[n] = r.kind;
            IPSB_CUDA(cudaEventElapsedTime(&start_ms[n], pr.base, r.a));
            IPSB_CUDA(cudaEventElapsedTime(&stop_ms[n], pr.base, r.b));
            ++n;
        }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    pr.recs.clear();
    *n_out = n;
    return 0;
}

int64_t ipsb_resnet_workspace_bytes(const ipsb_resnet_desc* net, int64_t chunk, int C, int H, int W) {
    (void)C;
    const int64_t tail = align256(chunk * net->D * 4) + align256(chunk * 8) + 1024;
    if (net->dt == IPSB_BF16 && net->stem.mode >= 3) return make_pf_plan(net, chunk, H, W).total + tail;
    int64_t staged;
    const int64_t act = max_act_elems(net, H, W, &staged);
    const int64_t es = (int64_t)esize(net->dt);
    return align256(chunk * staged * es) + 4 * align256(chunk * act * es) + tail;
}

static int resnet_logits_pf(const ipsb_resnet_desc* net, const float* patches, int64_t first_row, int64_t n_rows,
                            int C, int H, int W, int64_t n_per_image, int64_t chunk, void* workspace, int64_t lane_bytes, int lanes,
                            int zero_init, float* emb_out, float* z_out, void* caller_stream, const ipsb_image_geo* img_geo = nullptr) {
    const PfPlan pl = make_pf_plan(net, chunk, H, W);
    const int64_t n_chunks = (n_rows + chunk - 1) / chunk;
    if (lanes > n_chunks) lanes = (int)n_chunks;
    LaneStreams* ls = nullptr;
    const bool ds_side = getenv("IPSB_DS_INLINE") == nullptr;     // downsample convolutions on the lane's side stream
    if (lanes > 1 || ds_side) {
        if (int rc = lane_streams(&ls)) return rc;
    }
    size_t pf_bytes = 0;
    for (int g = 0; g < pl.n_groups; ++g) pf_bytes += 4 * (size_t)pl.group_bytes[g];
    // a candidate group runs dense in this call when every chunk has an even number of full 128-pixel tiles
    bool dense[4] = {false, false, false, false};
    for (int g = 0; g < pl.n_groups; ++g) {
        if (pl.dense_bytes[g] == 0) continue;
        const int64_t hw = (int64_t)pl.H[g] * pl.W[g];
        const int64_t last = n_rows % chunk;
        dense[g] = (n_rows >= chunk ? (chunk * hw) % 256 == 0 : true) && (last == 0 || (last * hw) % 256 == 0);
    }
    if (zero_init)   // pad rows of the padded-flat buffers must be zero; kernels keep them zero afterwards
        for (int l = 0; l < lanes; ++l)
            IPSB_CUDA(cudaMemsetAsync((char*)workspace + l * lane_bytes + pl.staged_bytes + pl.stem_bytes, 0, pf_bytes,
                                      (cudaStream_t)caller_stream));
    if (lanes > 1) {
        IPSB_CUDA(cudaEventRecord(ls->fork, (cudaStream_t)caller_stream));
        for (int l = 0; l < lanes; ++l) IPSB_CUDA(cudaStreamWaitEvent(ls->s[l], ls->fork, 0));
    }

    auto run_chunk = [&](int64_t lo, int64_t ci) -> int {
        const int lane = (int)(ci % lanes);
        void* stream = lanes > 1 ? (void*)ls->s[lane] : caller_stream;
        char* ws = (char*)workspace + lane * lane_bytes;
        void* staged = ws;               ws += pl.staged_bytes;
        void* stem_out = ws;             ws += pl.stem_bytes;
        void* gb[4][4];
        for (int g = 0; g < pl.n_groups; ++g)
            for (int i = 0; i < 4; ++i) { gb[g][i] = ws; ws += pl.group_bytes[g]; }
        void* gd[4][4];
        for (int g = 0; g < pl.n_groups; ++g)
            for (int i = 0; i < 4; ++i) { gd[g][i] = ws; ws += pl.dense_bytes[g]; }
        float* emb_ws = (float*)ws;      ws += align256(chunk * net->D * 4);
        int64_t* pos_idx = (int64_t*)ws;
        const int64_t P = (n_rows - lo < chunk) ? n_rows - lo : chunk;
        // stage -> stem (+ max-pool)
        int rc = 0;
        const ipsb_conv_desc& st = net->stem;
        const int hs = out_dim(H, 7, 2, 3), wsz = out_dim(W, 7, 2, 3);
        const int hq = out_dim(hs, 3, 2, 1), wq = out_dim(wsz, 3, 2, 1);
        const int64_t sub = P;                // (sub-chunks that keep the stem output in L2 measured slower: smaller grids)
        const pf::Geo gq = pf::make(P, hq, wq);
        for (int64_t s0 = 0; s0 < P; s0 += sub) {
            const int64_t Ps = (P - s0 < sub) ? P - s0 : sub;
            if (st.mode == 4) {               // space-to-depth frame -> shifted-window stem -> strided max-pool
                rc = PROF(K_STAGE, stream, img_geo ? ipsb_stage_image_s2d(patches, img_geo, first_row + lo + s0, Ps, C, H, W, staged, stream)
                                                   : ipsb_stage_patches_s2d(patches, nullptr, first_row + lo + s0, Ps, C, H, W, staged, stream));
                if (rc) return rc;
                if (st.cout == 64 && !getenv("IPSB_STEM_UNFUSED")) {   // stem + pool fused: the stem output stays on chip
                    rc = PROF(K_STEM, stream, ipsb_stem_pool_s2d(staged, st.w, st.scale, st.shift, (char*)gb[0][0] + (size_t)s0 * gq.Sp * st.cout * 2, Ps, H, W,
                                                                  1, stream));
                    if (rc) return rc;
                    continue;
                }
                rc = PROF(K_STEM, stream, ipsb_conv_bf16_umma(staged, st.w, st.scale, st.shift, nullptr, stem_out, Ps, H, W, 16, st.cout, 7, 7, 2, 3, 1, 4,
                                                               stream));
                if (rc) return rc;
                rc = PROF(K_POOL, stream, ipsb_maxpool3x3s2_pf_strided(stem_out, (char*)gb[0][0] + (size_t)s0 * gq.Sp * st.cout * 2, Ps, hs, wsz, st.cout,
                                                                        wsz + 3, (hs + 3) * (wsz + 3), stream));
                if (rc) return rc;
                continue;
            }
            rc = PROF(K_STAGE, stream, ipsb_stage_patches_padded(patches, nullptr, first_row + lo + s0, Ps, C, H, W, 3, 4, H + 6, W + 6, staged, stream));
            if (rc) return rc;
            rc = PROF(K_STEM, stream, ipsb_conv_bf16_umma(staged, st.w, st.scale, st.shift, nullptr, stem_out, Ps, H + 6, W + 6, 4, st.cout, 7, 7, 2, 3,
                                                           1, 3, stream));
            if (rc) return rc;
            // patch s0 of the chunk starts s0*Sp rows into the padded-flat buffer (same lead-in G0)
            rc = PROF(K_POOL, stream, ipsb_maxpool3x3s2_pf(stem_out, (char*)gb[0][0] + (size_t)s0 * gq.Sp * st.cout * 2, Ps, hs, wsz, st.cout, stream));
            if (rc) return rc;
        }
        int h = out_dim(hs, 3, 2, 1), w = out_dim(wsz, 3, 2, 1);   // geometry of the current activation
        const void* cur = gb[0][0];
        int cur_slot = 0;
        bool cur_dense = false;                  // layout of `cur`: padded-flat, or dense (P*H*W, C) rows
        for (int b = 0; b < net->n_blocks; ++b) {
            const ipsb_block_desc& blk = net->blocks[b];
            const int g = b / 2;
            const bool out_dense = dense[g];
            void** bufs = out_dense ? gd[g] : gb[g];
            int free_ids[3], nf = 0;
            if (b % 2 == 0 && g > 0) { free_ids[0] = 0; free_ids[1] = 1; free_ids[2] = 2; nf = 3; }   // input lives in the previous group
            else for (int i = 0; i < 4; ++i) if (i != cur_slot) free_ids[nf++] = i;
            const void* idt = cur;               // (no downsample: same group, same layout as the output)
            bool join_pending = false;
            if (blk.has_ds && ds_side) {
                // the 1x1 stride-2 downsample is a short, latency-bound launch (1 % of the FLOPs, 22-43 us): it runs next to the
                // block's first convolution on the lane's side stream and is joined before the second one, which adds it in
                cudaStream_t aux = ls->aux[lane];
                IPSB_CUDA(cudaEventRecord(ls->aux_fork[lane], (cudaStream_t)stream));
                IPSB_CUDA(cudaStreamWaitEvent(aux, ls->aux_fork[lane], 0));
                rc = PROF(K_CONV, aux, run_conv_pf(blk.ds, cur, nullptr, bufs[free_ids[0]], P, h, w, 0, aux, !cur_dense, !out_dense));
                IPSB_CUDA(cudaEventRecord(ls->aux_join[lane], aux));
                join_pending = true;
                if (rc) { cudaStreamWaitEvent((cudaStream_t)stream, ls->aux_join[lane], 0); return rc; }
                idt = bufs[free_ids[0]];
            } else if (blk.has_ds) {
                rc = PROF(K_CONV, stream, run_conv_pf(blk.ds, cur, nullptr, bufs[free_ids[0]], P, h, w, 0, stream, !cur_dense, !out_dense));
                if (rc) return rc;
                idt = bufs[free_ids[0]];
            }
            rc = PROF(K_CONV, stream, run_conv_pf(blk.c1, cur, nullptr, bufs[free_ids[1]], P, h, w, 1, stream, !cur_dense, !out_dense));
            if (join_pending) IPSB_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, ls->aux_join[lane], 0));
            if (rc) return rc;
            rc = PROF(K_CONV, stream, run_conv_pf(blk.c2, bufs[free_ids[1]], idt, bufs[free_ids[2]], P, pl.H[g], pl.W[g], 1, stream, !out_dense, !out_dense));
            if (rc) return rc;
            cur = bufs[free_ids[2]];
            cur_slot = free_ids[2];
            cur_dense = out_dense;
            h = pl.H[g]; w = pl.W[g];
        }
        const int gl = pl.n_groups - 1;
        IPSB_REQUIRE(pl.C[gl] == net->D, "resnet_logits: encoder width %d != D %d", pl.C[gl], net->D);
        float* emb = emb_out ? emb_out + lo * net->D : emb_ws;
        rc = PROF(K_POOL, stream, cur_dense ? ipsb_avgpool(cur, emb, P, h * w, pl.C[gl], IPSB_BF16, stream) : ipsb_avgpool_pf(cur, emb, P, h, w, pl.C[gl], stream));
        if (rc) return rc;
        const int64_t* idx = nullptr;
        if (net->add_tab) {
            iota_mod_kernel<<<(unsigned)ipsb::ceil_div(P, 256), 256, 0, (cudaStream_t)stream>>>(pos_idx, first_row + lo, P, n_per_image);
            IPSB_LAUNCH_CHECK();
            idx = pos_idx;
        }
        rc = PROF(K_LOGITS, stream, ipsb_logits(emb, net->U, net->add_tab, idx, z_out + lo * net->HT, P, net->D, net->HT, stream));
        return rc;
    };
    int status = 0;
    int64_t ci = 0;
    for (int64_t lo = 0; lo < n_rows && status == 0; lo += chunk, ++ci) status = run_chunk(lo, ci);
    // every exit joins the lanes back into the caller's stream: after an error the caller's stream must still be ordered
    // behind the kernels already enqueued on the lanes (they write `workspace` and `z_out`)
    if (lanes > 1)
        for (int l = 0; l < lanes; ++l) {
            const cudaError_t e1 = cudaEventRecord(ls->done[l], ls->s[l]);
            const cudaError_t e2 = (e1 == cudaSuccess) ? cudaStreamWaitEvent((cudaStream_t)caller_stream, ls->done[l], 0) : e1;
            if (e2 != cudaSuccess && status == 0) status = ipsb::fail("resnet_logits: lane join -> %s", cudaGetErrorString(e2));
        }
    return status;
}

int ipsb_resnet_logits_image(const ipsb_resnet_desc* net, const float* images, const ipsb_image_geo* geo, int64_t first_row,
                             int64_t n_rows, int C, int H, int W, int64_t chunk,
                             void* workspace, int64_t workspace_bytes, int zero_init, float* emb_out, float* z_out, void* stream) {
    IPSB_REQUIRE(net && images && geo && z_out && workspace, "resnet_logits_image: null argument");
    IPSB_REQUIRE(n_rows > 0 && chunk > 0 && net->n_blocks > 0 && net->n_blocks <= 8 && net->n_blocks % 2 == 0, "resnet_logits_image: bad sizes");
    IPSB_REQUIRE(net->dt == IPSB_BF16 && net->stem.mode == 4, "resnet_logits_image: needs the bf16 path with the space-to-depth stem");
    const int64_t lane_bytes = ipsb_resnet_workspace_bytes(net, chunk, C, H, W);
    IPSB_REQUIRE(workspace_bytes >= lane_bytes, "resnet_logits_image: workspace too small");
    int lanes = (int)(workspace_bytes / lane_bytes);
    if (lanes > MAX_LANES) lanes = MAX_LANES;
    return resnet_logits_pf(net, images, first_row, n_rows, C, H, W, geo->n_per_image, chunk, workspace, lane_bytes, lanes, zero_init,
                            emb_out, z_out, stream, geo);
}

int ipsb_resnet_logits(const ipsb_resnet_desc* net, const float* patches, int64_t first_row, int64_t n_rows,
                       int C, int H, int W, int64_t n_per_image, int64_t chunk,
                       void* workspace, int64_t workspace_bytes, int zero_init, float* emb_out, float* z_out, void* stream) {
    IPSB_REQUIRE(net && patches && z_out && workspace, "resnet_logits: null argument");
    IPSB_REQUIRE(n_rows > 0 && chunk > 0 && net->n_blocks > 0 && net->n_blocks <= 8 && net->n_blocks % 2 == 0, "resnet_logits: bad sizes");
    const int64_t lane_bytes = ipsb_resnet_workspace_bytes(net, chunk, C, H, W);
    IPSB_REQUIRE(workspace_bytes >= lane_bytes, "resnet_logits: workspace too small");
    if (net->dt == IPSB_BF16 && net->stem.mode >= 3) {
        int lanes = (int)(workspace_bytes / lane_bytes);        // a workspace of L lanes runs L chunks concurrently
        if (lanes > MAX_LANES) lanes = MAX_LANES;
        return resnet_logits_pf(net, patches, first_row, n_rows, C, H, W, n_per_image, chunk, workspace, lane_bytes, lanes, zero_init,
                                emb_out, z_out, stream);
    }
    const int dt = net->dt;
    const int64_t es = (int64_t)esize(dt);
    int64_t staged_elems;
    const int64_t act = max_act_elems(net, H, W, &staged_elems);
    char* ws = (char*)workspace;
    void* staged = ws;                       ws += align256(chunk * staged_elems * es);
    void* buf[4];
    for (int i = 0; i < 4; ++i) { buf[i] = ws; ws += align256(chunk * act * es); }
    float* emb_ws = (float*)ws;               ws += align256(chunk * net->D * 4);
    int64_t* pos_idx = (int64_t*)ws;

    for (int64_t lo = 0; lo < n_rows; lo += chunk) {
        const int64_t P = (n_rows - lo < chunk) ? n_rows - lo : chunk;
        int rc;
        int h = out_dim(H, 7, 2, 3), w = out_dim(W, 7, 2, 3);
        if (net->stem.mode == 3) {           // zero-bordered frame for the TMA-fed stem
            rc = ipsb_stage_patches_padded(patches, nullptr, first_row + lo, P, C, H, W, 3, 4, H + 6, W + 6, staged, stream);
            if (rc) return rc;
            rc = run_conv(net, net->stem, staged, nullptr, buf[0], P, H + 6, W + 6, 1, stream);
        } else {
            rc = ipsb_stage_patches(patches, nullptr, first_row + lo, P, C, H, W, 4, dt, staged, stream);
            if (rc) return rc;
            rc = run_conv(net, net->stem, staged, nullptr, buf[0], P, H, W, 1, stream);
        }
        if (rc) return rc;
        rc = ipsb_maxpool3x3s2(buf[0], buf[1], P, h, w, net->stem.cout, dt, stream);
        if (rc) return rc;
        h = out_dim(h, 3, 2, 1); w = out_dim(w, 3, 2, 1);
        int cur = 1;                          // index of the buffer holding the block input
        int c_out = net->stem.cout;
        for (int b = 0; b < net->n_blocks; ++b) {
            const ipsb_block_desc& blk = net->blocks[b];
            int free_ids[3], nf = 0;
            for (int i = 0; i < 4; ++i) if (i != cur) free_ids[nf++] = i;
            const void* idt = buf[cur];
            const int ho = out_dim(h, blk.c1.kh, blk.c1.stride, blk.c1.pad), wo = out_dim(w, blk.c1.kw, blk.c1.stride, blk.c1.pad);
            if (blk.has_ds) {
                rc = run_conv(net, blk.ds, buf[cur], nullptr, buf[free_ids[0]], P, h, w, 0, stream);
                if (rc) return rc;
                idt = buf[free_ids[0]];
            }
            rc = run_conv(net, blk.c1, buf[cur], nullptr, buf[free_ids[1]], P, h, w, 1, stream);
            if (rc) return rc;
            rc = run_conv(net, blk.c2, buf[free_ids[1]], idt, buf[free_ids[2]], P, ho, wo, 1, stream);
            if (rc) return rc;
            cur = free_ids[2];
            h = ho; w = wo; c_out = blk.c2.cout;
        }
        IPSB_REQUIRE(c_out == net->D, "resnet_logits: encoder width %d != D %d", c_out, net->D);
        float* emb = emb_out ? emb_out + lo * net->D : emb_ws;
        rc = ipsb_avgpool(buf[cur], emb, P, h * w, c_out, dt, stream);
        if (rc) return rc;
        const int64_t* idx = nullptr;
        if (net->add_tab) {
            iota_mod_kernel<<<(unsigned)ipsb::ceil_div(P, 256), 256, 0, (cudaStream_t)stream>>>(pos_idx, first_row + lo, P, n_per_image);
            IPSB_LAUNCH_CHECK();
            idx = pos_idx;
        }
        rc = ipsb_logits(emb, net->U, net->add_tab, idx, z_out + lo * net->HT, P, net->D, net->HT, stream);
        if (rc) return rc;
    }
    return 0;
}

}  // extern "C"
