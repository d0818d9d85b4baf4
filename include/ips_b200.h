/*
 * ips_b200 -- C ABI of the B200-native IPS hot path (sm_100a).
 *
 * The reference (benbergner/ips) is pure Python: it has no FFI layer, its
 * boundary is the IPSNet class surface (architecture/ips_net.py:85,169,264).
 * Each entry point below names the reference code it replaces.  The Python
 * drop-in in ips_b200/ binds these with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - buffers are owned by the caller (PyTorch allocates them) and must
 *     outlive the stream work; the library owns only opaque plan handles;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - return value 0 = ok, non-zero = error; ipsb_last_error() gives the
 *     message of the calling thread's last failure; no exceptions cross;
 *   - no CPU fallback exists: without a CUDA device every call fails.
 */
#ifndef IPS_B200_H
#define IPS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IPSB_ABI_VERSION 1
#if defined(__GNUC__)
#define IPSB_API __attribute__((visibility("default")))
#else
#define IPSB_API
#endif

/* activation dtype codes */
#define IPSB_F32 0
#define IPSB_BF16 1

IPSB_API int ipsb_abi_version(void);
IPSB_API const char* ipsb_last_error(void);
/* 0 when a device with compute capability 10.x is current */
IPSB_API int ipsb_device_ok(void);

/* ---------------------------------------------------------------- staging
 * Replaces: patches[:, lo:hi].to(device).reshape(-1, C, ph, pw)
 * (ips_net.py:206,209,223,227) and the shuffle copy (utils/utils.py:39,56).
 * Gathers `n_rows` patches src[row_idx[i]] of a (rows, C, H, W) fp32 NCHW
 * tensor into channels-last (n_rows, H, W, Cpad) of dtype `dt`, zero-filling
 * channels C..Cpad-1.  row_idx == NULL means rows first_row .. first_row+n_rows-1. */
IPSB_API int ipsb_stage_patches(const float* src, const int64_t* row_idx, int64_t first_row, int64_t n_rows,
                       int C, int H, int W, int Cpad, int dt, void* dst, void* stream);
/* Same gather into a zero-bordered bf16 frame (n_rows, Hp, Wp, 4) with the image at (pad_top, pad_left):
 * the input format of the TMA-fed stem (mode 3 of ipsb_conv_bf16_umma: pad_top 3, pad_left 4, Hp=H+6, Wp=W+6). */
IPSB_API int ipsb_stage_patches_padded(const float* src, const int64_t* row_idx, int64_t first_row, int64_t n_rows,
                              int C, int H, int W, int pad_top, int pad_left, int Hp, int Wp, void* dst, void* stream);

/* Space-to-depth staging for the stem (mode 4 of ipsb_conv_bf16_umma): frame pixel (Y',X') of a patch holds 16 bf16,
 * channel (dy*2+dx)*4+c = in(2Y'+dy-4, 2X'+dx-4, c); (H/2+3) x (W/2+3) frame pixels per patch.  The 7x7/2 stem is then a
 * 4x4 stride-1 convolution whose 16 taps are shifted views of one block of frame rows (weights (Cout,256),
 * k = (a*4+b)*16 + (dy*2+dx)*4 + c = w[2a+dy-1, 2b+dx-1, c]).  The stem output keeps the frame's row order
 * (output (oy,ox) of patch p at row p*Sp + oy*(W/2+3) + ox, Sp = (H/2+3)*(W/2+3)); ipsb_maxpool3x3s2_pf_strided reads it. */
IPSB_API int ipsb_stage_patches_s2d(const float* src, const int64_t* row_idx, int64_t first_row, int64_t n_rows, int C, int H, int W,
                           void* dst, void* stream);
/* The same staging through the copy engine (stage_tma.cu): one 4-D tensor-map box per band of 8 frame rows -- the zero
 * padding on all four sides is the TMA unit's out-of-bounds fill --, fp32 -> frame pixels in shared memory, one bulk store
 * per band.  ipsb_stage_patches_s2d dispatches here when ipsb_stage_tma_ok (16-byte aligned source, W % 4 == 0, W <= 248);
 * IPSB_STAGE_NO_TMA=1 keeps the load/store kernel. */
IPSB_API int ipsb_stage_tma_ok(const float* src, int C, int H, int W);
IPSB_API int ipsb_stage_patches_s2d_tma(const float* src, const int64_t* row_idx, int64_t first_row, int64_t n_rows, int C, int H,
                                        int W, void* dst, void* stream);
IPSB_API int ipsb_maxpool3x3s2_pf_strided(const void* x, void* y, int64_t P, int H, int W, int C, int in_Wp, int in_Sp, void* stream);
/* On-device patchify: the reference cuts images into patches on the CPU workers
 * (`img.unfold(1, ph, sh).unfold(2, pw, sw).permute(1, 2, 0, 3, 4).reshape(-1, C, ph, pw)`,
 * data/megapixel_mnist/mnist_dataset.py:47-53, data/traffic/traffic_dataset.py:337-343).  With an ipsb_image_geo the
 * staging kernel and the final gather read the patch rectangles straight out of (B, C, img_h, img_w) fp32 images:
 * patch n of an image is grid row n / n_cols, column n % n_cols, n_cols = (img_w - pw) / stride_w + 1; overlapping
 * strides are allowed.  n_per_image must equal the number of grid cells. */
typedef struct {
    int32_t img_h, img_w;
    int32_t stride_h, stride_w;
    int32_t n_per_image;
} ipsb_image_geo;
IPSB_API int ipsb_stage_image_s2d(const float* img, const ipsb_image_geo* geo, int64_t first_row, int64_t n_rows, int C, int H, int W,
                         void* dst, void* stream);
/* out (B, M, C, H, W) fp32 <- patch idx[b, m] of image b (idx NULL: all patches in order = patchify; idx < 0: zeros) */
IPSB_API int ipsb_gather_patches_image(const float* img, const ipsb_image_geo* geo, const int64_t* idx, int B, int M, int C, int H, int W,
                              float* out, void* stream);

/* im2col rows for the weight-gradient GEMM of the conv encoder's backward pass (A12): x (P,H,W,C) bf16 channels-last ->
 * out (P*Ho*Wo, Kp) bf16, column (r*kw+s)*C + c, zero beyond kh*kw*C; Kp a multiple of 8. */
IPSB_API int ipsb_im2col_bf16(const void* x, void* out, int64_t P, int H, int W, int C, int kh, int kw, int stride, int pad, int Kp,
                              void* stream);

/* conv1 -> bn1 -> relu -> maxpool (architecture/ips_net.py:17-39) in one kernel: s2d frame of P patches of HxW ->
 * padded-flat (ipsb_pf_rows(P, Hq, Wq), 64) bf16 with Hq = (H/2 - 1)/2 + 1; the stem output stays in shared memory.
 * w: (64, 256) bf16 in the mode-4 packing; y's pad rows must already be zero (they are not written). */
IPSB_API int ipsb_stem_pool_s2d(const void* frame, const void* w, const float* scale, const float* shift, void* y, int64_t P, int H, int W,
                       int relu, void* stream);

/* ---------------------------------------------------------------- encoder, fp32 SIMT ("exact" mode)
 * Replaces: conv2d + eval-mode batch_norm [+ residual add] [+ relu] of the
 * truncated torchvision ResNet (ips_net.py:17-52).  x: (P,H,W,Cin) NHWC fp32,
 * w: (kh*kw*Cin, Cout) fp32 with k = (r*kw+s)*Cin+c, y = act(conv*scale+shift+res). */
IPSB_API int ipsb_conv_f32(const float* x, const float* w, const float* scale, const float* shift,
                  const float* res, float* y, int64_t P, int H, int W, int Cin, int Cout,
                  int kh, int kw, int stride, int pad, int relu, void* stream);
/* y[M,N] = act((A[M,K] @ W[N,K]^T) * scale[N] + shift[N]); scale/shift may be NULL.
 * Replaces nn.Linear call sites (ips_net.py:57, transformer.py:76-77). */
IPSB_API int ipsb_linear_f32(const float* a, const float* w, const float* scale, const float* shift,
                    float* y, int64_t M, int N, int K, int relu, void* stream);
/* max_pool2d(3, stride 2, pad 1) on NHWC; dt selects fp32 / bf16 */
IPSB_API int ipsb_maxpool3x3s2(const void* x, void* y, int64_t P, int H, int W, int C, int dt, void* stream);
/* grad-mode step, channels-last fp32: backward of max_pool2d(3, 2, 1) (gather form, torch's first-maximum rule), the
 * BasicBlock tail y = relu(a + b) (ips_net.py:34-50 via torchvision's BasicBlock.forward) and its backward mask */
IPSB_API int ipsb_maxpool3x3s2_bwd_f32(const float* x, const float* dy, float* dx, int64_t P, int H, int W, int C, void* stream);
IPSB_API int ipsb_add_relu_f32(const float* a, const float* b, float* y, int64_t n, void* stream);
IPSB_API int ipsb_relu_bwd_f32(const float* y, const float* dy, float* dx, int64_t n, void* stream);
/* adaptive_avg_pool2d(1): (P,HW,C) dt -> (P,C) fp32 */
IPSB_API int ipsb_avgpool(const void* x, float* y, int64_t P, int HW, int C, int dt, void* stream);
/* LayerNorm without affine (ips_net.py:56): y = (x-mean)/sqrt(var+eps), rows of F floats */
IPSB_API int ipsb_layernorm_rows_f32(const float* x, float* y, int64_t rows, int F, float eps, void* stream);

/* ---------------------------------------------------------------- encoder, bf16 tcgen05 implicit GEMM
 * Same contract as ipsb_conv_f32 with x,res,y bf16 NHWC and w (Cout, K) bf16
 * K-major, K = kh*kw*Cin padded to a multiple of 64.  mode 0: Cin % 64 == 0, TMA-fed
 * (M tile = box of output pixels); mode 2: same contract through a cp.async gather;
 * mode 4: the stem on the space-to-depth frame (H, W = image size, Cin = 16), shifted-window kernel;
 * mode 3: the stem on the zero-bordered frame of ipsb_stage_patches_padded (H, W = frame size), TMA-fed;
 * mode 1: the 7x7/2 stem on 4-channel-padded input (K laid out r*32+(s+1)*4+c, 256). */
IPSB_API int ipsb_conv_bf16_umma(const void* x, const void* w, const float* scale, const float* shift,
                        const void* res, void* y, int64_t P, int H, int W, int Cin, int Cout,
                        int kh, int kw, int stride, int pad, int relu, int mode, void* stream);
/* Padded-flat (PF) activations: P patches of HxW pixels stored as rows of C channels with one zero pixel after
 * every image row and one zero row after every patch, row(p,y,x) = (W+2) + p*(H+1)*(W+1) + y*(W+1) + x
 * (ips_b200/csrc/pf.cuh).  A stride-1 3x3 convolution on PF input/output runs as a shifted-window implicit
 * GEMM that reads its input once; other geometries read/write PF through TMA boxes.  res is PF iff out_pf. */
IPSB_API int64_t ipsb_pf_rows(int64_t P, int H, int W);    /* rows to allocate (zero-initialised once) */
IPSB_API int ipsb_conv_bf16_pf(const void* x, const void* w, const float* scale, const float* shift, const void* res,
                      void* y, int64_t P, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad,
                      int relu, int in_pf, int out_pf, void* stream);
IPSB_API int ipsb_maxpool3x3s2_pf(const void* x, void* y, int64_t P, int H, int W, int C, void* stream);  /* dense bf16 -> PF */
IPSB_API int ipsb_avgpool_pf(const void* x, float* y, int64_t P, int H, int W, int C, void* stream);      /* PF bf16 -> (P,C) fp32 */
/* y[M,N] (fp32) = act((A[M,K] @ W[N,K]^T) * scale + shift), A and W bf16 K-major,
 * K % 64 == 0, N % 64 == 0.  tcgen05 path for the Linear layers. */
IPSB_API int ipsb_linear_bf16_umma(const void* a, const void* w, const float* scale, const float* shift,
                          float* y, int64_t M, int N, int K, int relu, void* stream);
/* General tensor-core GEMM of the grad-mode step (forward / data-gradient / weight-gradient of nn.Linear):
 *   mode 0 (NT): a (M,K), b (N,K)   c = a b^T        mode 1 (NN): a (M,K), b (K,N)   c = a b
 *   mode 2 (TN): a (K,M), b (K,N)   c = a^T b   (contraction over rows; split along K into fp32 partials in
 *   `workspace`, ipsb_gemm_workspace_bytes, then reduced deterministically; needs M % 128 == 0)
 * a, b bf16 row-major; c (M,N) fp32 (c_is_f32) or bf16; c = act(c * scale[N] + shift[N]); N % 64 == 0; row strides 16-byte aligned. */
IPSB_API int64_t ipsb_gemm_workspace_bytes(int mode, int64_t M, int N, int64_t K);
IPSB_API int ipsb_gemm_bf16(int mode, const void* a, const void* b, const float* scale, const float* shift, void* c, int c_is_f32,
                   int64_t M, int N, int64_t K, int relu, void* workspace, int64_t workspace_bytes, void* stream);
/* fp32 CUDA-core GEMM in the same three layouts (any sizes with K % 4 == 0 / N % 4 == 0 as the layout needs): exact mode
 * of the train step and the fallback for shapes the tensor-core kernel does not take (class heads). */
IPSB_API int ipsb_gemm_f32(int mode, const float* a, const float* b, const float* shift, float* c, int64_t M, int N, int64_t K,
                  int relu, void* stream);
IPSB_API int ipsb_colsum_f32(const float* x, const float* y, float* out, float* scratch /* 64*cols floats or NULL */, int64_t rows,
                    int cols, void* stream);   /* out[c] = sum_r x[r,c] * (y ? y[r,c] : 1) */
IPSB_API int ipsb_cast_bf16(const float* x, void* y, int64_t n, void* stream);
/* grad-mode step: both bf16 operand layouts of a conv weight w (Cout, Cin, kh, kw) fp32 in one launch:
 * w_nk (Cout, kh*kw*Cin) K-major (forward, weight gradient) and, if non-NULL, w_t (Cin, kh*kw*Cout) = flipped + transposed
 * (input gradient as a convolution of dy). */
/* weight gradient as the TN GEMM leaves it, (K', Cout) fp32 with k = (r*kw + s)*C + ci, -> nn.Conv2d's (Cout, Cin, kh, kw) */
IPSB_API int ipsb_wgrad_to_oihw(const float* dw_kc, int Cout, int Cin, int C, int kh, int kw, float* out, void* stream);
IPSB_API int ipsb_conv_weight_layouts(const float* w, int Cout, int Cin, int kh, int kw, void* w_nk, void* w_t, void* stream);
/* fp32 (rows,F) -> bf16 with optional no-affine LayerNorm fused (projector prologue) */
IPSB_API int ipsb_rows_to_bf16(const float* x, void* y, int64_t rows, int F, int layernorm, float eps, void* stream);
/* the same for features already stored as bf16 (flat bf16 feature bags halve the bytes of the CAMELYON path);
 * LayerNorm statistics in fp32 on the exactly-upcast values */
IPSB_API int ipsb_rows_bf16_to_bf16(const void* x, void* y, int64_t rows, int F, int layernorm, float eps, void* stream);

/* ---------------------------------------------------------------- scoring
 * The learned-query score is linear in the embedding before the softmax:
 * z[n,h,t] = emb_n . U[:,h,t] with U = k_w_h^T (q_w q_t)_h / sqrt(D_k)
 * (transformer.py:29-31,76-79).  U is (D, H*T), column h*T+t. */
IPSB_API int ipsb_score_basis(const float* q_tok, const float* q_w, const float* k_w, float* U,
                     int D, int H, int Dk, int T, void* stream);
/* z[row, :] = emb[row, :] @ U (+ add_tab[add_idx[row] or row, :]); emb (rows, D) fp32.
 * add_tab (rows_tab, HT) is the positional-encoding contribution pos @ U (ips_net.py:234-238). */
IPSB_API int ipsb_logits(const float* emb, const float* U, const float* add_tab, const int64_t* add_idx,
                float* z, int64_t rows, int D, int HT, void* stream);
/* scores (B,L) from logits (B,L,H*T): softmax over L per (h,t), mean over heads then
 * tokens (transformer.py:31,143-148). */
IPSB_API int ipsb_scores_from_logits(const float* z, float* scores, int B, int L, int H, int T, void* stream);
/* Stable top-M (torch.topk at ips_net.py:148 with lowest-position tie-break):
 * idx_out/val_out (B,M) = first M of a stable descending sort of scores (B,L). */
IPSB_API int ipsb_topm_stable(const float* scores, int B, int L, int M, int64_t* idx_out, float* val_out, void* stream);
/* The whole sequential selection loop of ips_net.py:213-241 on a per-patch logit
 * table z (B,N,H*T) (patch order = ORIGINAL order).  perm (B or 1, N) int64 is the
 * scan order (shuffled position -> original index) or NULL for identity;
 * perm_batch_stride = N for per-instance permutations, 0 for a shared one.
 * Outputs (B,M): mem_pos = winners' shuffled positions (the reference's mem_idx),
 * mem_src = their original indices, mem_score = final-iteration scores; best first.  Equal scores resolve to
 * the patch scanned first.  workspace (ipsb_select_loop_workspace_bytes; may be NULL = slower single-CTA path)
 * holds the scan-ordered logit table and the double-buffered memory set. */
IPSB_API int64_t ipsb_select_loop_workspace_bytes(int B, int N, int HT, int M);
IPSB_API int ipsb_select_loop(const float* z, const int64_t* perm, int64_t perm_batch_stride,
                     int B, int N, int H, int T, int M, int I,
                     int64_t* mem_pos, int64_t* mem_src, float* mem_score,
                     void* workspace, int64_t workspace_bytes, void* stream);
/* Streamed form: zs (B,N,H*T) is already in SCAN order (ipsb_projector_logits_scan) -- no permuted copy -- and may still be
 * being written on another stream: the loop waits (bounded) on tile_flags[(b N + pos) / tile_rows] before it reads a chunk.
 * perm gives mem_src = perm[pos] (NULL: pos).  sync_words (2 ints, zeroed by the caller, or NULL): [0] is set when the
 * loop's kernel is resident (gate the producer's launch on it with ipsb_wait_word so the producer does not take the SMs
 * first), [1] != 0 afterwards means a wait ran out (producer missing).  Returns -2 without launching when the shape is
 * outside the 8-CTA cluster kernel (M + I < 2048, H*T not a power of two, slices beyond shared memory): use
 * ipsb_select_loop.  Same results as ipsb_select_loop on the same scan order, bit for bit. */
IPSB_API int ipsb_select_loop_scan(const float* zs, const int64_t* perm, int64_t perm_batch_stride,
                     int B, int N, int H, int T, int M, int I,
                     int64_t* mem_pos, int64_t* mem_src, float* mem_score,
                     void* workspace, int64_t workspace_bytes,
                     const int* tile_flags, int tile_rows, int* sync_words, void* stream);
IPSB_API int ipsb_wait_word(const int* word, void* stream);
/* Scan order drawn on the device (the reference draws it on the host: randperm / rand().argsort, utils/utils.py:33-58; 0.4 ms
 * of host time for 50 000 patches).  perm (rows, N) int64: row r = a keyed bijection of [0, N) -- eight rounds of xor /
 * odd multiply mod 2^k / xor-shift on k = ceil(log2 N) bits with cycle walking -- for the key (key[0], key[1], r); key = two
 * int64 in DEVICE memory (drawn from the CUDA generator by the caller).  Same distribution family as a shuffle, not the
 * reference's random stream; oracle: ips_oracle.keyed_scan_order. */
IPSB_API int ipsb_keyed_scan_order(const int64_t* key, int rows, int N, int64_t* perm, void* stream);
/* Load the kernels of the streamed pair (loop, gate, projector) now: CUDA's lazy first-use loading of the producer could
 * otherwise wait for the consumer that is already spinning on its flags.  Call once before the first streamed selection. */
IPSB_API int ipsb_streamed_preload(void);
IPSB_API int ipsb_projector_preload(void);

/* ---------------------------------------------------------------- native encoder executor
 * One call = the whole eval-mode patch encoder + logit projection for `n_rows` patches
 * (everything ips_net.py:209/227 + transformer.py:76-79 do per chunk), chunked internally
 * so activations stay L2-sized.  Issues every launch on `stream`; no host synchronisation. */
typedef struct {
    const void* w;        /* bf16 (Cout,Kpad) K-major, or fp32 (K,Cout) in IPSB_F32 mode */
    const float* scale;   /* folded BatchNorm */
    const float* shift;
    int32_t cin, cout, kh, kw, stride, pad, mode, _pad;
} ipsb_conv_desc;

typedef struct {
    ipsb_conv_desc c1, c2, ds;
    int32_t has_ds, _pad;
} ipsb_block_desc;

typedef struct {
    int32_t dt;           /* IPSB_BF16 or IPSB_F32 */
    int32_t n_blocks;     /* BasicBlocks after the stem + maxpool (4 or 8) */
    ipsb_conv_desc stem;  /* 7x7/2 on 4-channel-padded input */
    ipsb_block_desc blocks[8];
    int32_t D, HT;        /* embedding width, H*T */
    const float* U;       /* (D, HT) score basis */
    const float* add_tab; /* (N, HT) positional contribution or NULL */
} ipsb_resnet_desc;

/* bytes of scratch needed for chunks of `chunk` patches of (C,H,W): one "lane".  On the bf16 path a workspace of
 * L x this size (L <= 4) lets ipsb_resnet_logits run L chunks concurrently on internal streams (fork from / join
 * into `stream` with events), which fills the SMs a persistent kernel's last partial wave leaves idle. */
IPSB_API int64_t ipsb_resnet_workspace_bytes(const ipsb_resnet_desc* net, int64_t chunk, int C, int H, int W);
/* patches: (rows,C,H,W) fp32 NCHW on the device; row r reads patch first_row + r; add_tab row = (first_row + r) % n_per_image.
 * emb_out (n_rows, D) fp32 may be NULL; z_out (n_rows, HT) fp32.  zero_init != 0 on the first use of a workspace
 * (the padded-flat activation buffers rely on zero pad rows, which the kernels then preserve). */
IPSB_API int ipsb_resnet_logits(const ipsb_resnet_desc* net, const float* patches, int64_t first_row, int64_t n_rows,
                                int C, int H, int W, int64_t n_per_image, int64_t chunk,
                                void* workspace, int64_t workspace_bytes, int zero_init, float* emb_out, float* z_out,
                                void* stream);
/* same, reading the patches out of whole images (bf16 path with the space-to-depth stem only); H, W = patch size */
IPSB_API int ipsb_resnet_logits_image(const ipsb_resnet_desc* net, const float* images, const ipsb_image_geo* geo, int64_t first_row,
                                      int64_t n_rows, int C, int H, int W, int64_t chunk,
                                      void* workspace, int64_t workspace_bytes, int zero_init, float* emb_out, float* z_out,
                                      void* stream);

/* ---------------------------------------------------------------- aggregator + heads, forward (no-grad)
 * Replaces MultiHeadCrossAttention.forward / MLP.forward / get_preds in inference
 * (architecture/transformer.py:85-132, architecture/ips_net.py:157-166); projections use the GEMM entry points.
 * q_scaled: (T, H*Dk) = q_w(q) / sqrt(Dk); k: (B,M,H*Dk); v: (B,M,H*Dv); out: (B,T,H*Dv), Dv <= 64. */
IPSB_API int ipsb_cross_attention_f32(const float* q_scaled, const float* k, const float* v, float* out,
                                      int B, int M, int H, int Dk, int Dv, int T, void* stream);
/* y = LayerNorm(x + r[row % r_rows]) * gamma + beta over rows of D (r may be NULL) */
IPSB_API int ipsb_residual_layernorm_f32(const float* x, const float* r, int r_rows, const float* gamma, const float* beta,
                                         float* y, int64_t rows, int D, float eps, void* stream);
/* rows of n logits -> softmax (act 0) or sigmoid (act 1) */
IPSB_API int ipsb_head_activation_f32(const float* logits, float* y, int rows, int n, int act, void* stream);
/* Head activation + loss + gradient in one kernel: get_preds (ips_net.py:157-166) followed by compute_loss
 * (training/iterative.py:83-98).  act 0: softmax, loss = mean_b -log(p[class_idx[b]] + eps); act 1: sigmoid, BCELoss against
 * float targets (rows, n).  loss (1), dlogits (rows, n) = d loss / d logits or NULL, probs (rows, n) or NULL. */
IPSB_API int ipsb_head_loss_f32(const float* logits, const int64_t* class_idx, const float* targets, int rows, int n, int act, float eps,
                                float* loss, float* dlogits, float* probs, void* stream);
IPSB_API int ipsb_add_f32(const float* a, const float* b, float* y, int64_t n, void* stream);

/* ---------------------------------------------------------------- grad-mode operators (fp32), forward + backward
 * BatchNorm1d in batch-statistics mode (ips_net.py:58 under net.train()), LayerNorm backward
 * (transformer.py:107,130) and the cross-attention core with a dropout mask (transformer.py:29-41,98). */
IPSB_API int ipsb_bn_stats_f32(const float* x, float* mean, float* var, float* scratch /* 512*cols floats */, int64_t rows, int cols,
                      void* stream);   /* biased variance, two-stage deterministic reduction */
/* batch statistics with x read from HBM once (chunk-wise mean / M2 merged by the parallel-variance formula) and, when
 * rstd != NULL, the tail below in the same two launches; running_* may be NULL */
IPSB_API int ipsb_bn_stats_finalize_f32(const float* x, float* mean, float* var, float* rstd, float* running_mean, float* running_var,
                                        float momentum, float unbias, float eps, float* scratch /* 512*cols floats */, int64_t rows,
                                        int cols, void* stream);
/* rstd = 1/sqrt(var + eps) + nn.BatchNorm's running-statistics update (unbias = rows / (rows - 1)); running_* may be NULL */
IPSB_API int ipsb_bn_finalize_f32(const float* mean, const float* var, int cols, float momentum, float unbias, float eps, float* rstd,
                                  float* running_mean, float* running_var, void* stream);
IPSB_API int ipsb_bn_apply_f32(const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta, float* y,
                      int64_t rows, int cols, int relu, void* stream);
/* sums (2*cols) = [dbeta, dgamma]; dx = gamma*rstd*(g - mean(g) - xhat*mean(g*xhat)), g = dy*(y>0) when relu */
IPSB_API int ipsb_bn_backward_f32(const float* dy, const float* x, const float* y, const float* mean, const float* rstd,
                         const float* gamma, float* sums, float* dx, float* scratch /* 512*cols floats */, int64_t rows, int cols,
                         int relu, void* stream);
/* the two phases of ipsb_bn_backward_f32, for a BatchNorm synchronised over data-parallel ranks: the caller
 * all-reduces `sums` between them and passes the row count of the whole batch */
IPSB_API int ipsb_bn_backward_sums_f32(const float* dy, const float* x, const float* y, const float* mean, const float* rstd,
                                       float* sums, float* scratch /* 512*cols floats */, int64_t rows, int cols, int relu, void* stream);
IPSB_API int ipsb_bn_backward_apply_f32(const float* dy, const float* x, const float* y, const float* mean, const float* rstd,
                                        const float* gamma, const float* sums, float* dx, int64_t rows, int cols, int64_t rows_total,
                                        int relu, void* stream);
/* dx of y = LayerNorm(x)*gamma+beta (gamma may be NULL); xhat (optional) returns the normalised input for dgamma */
IPSB_API int ipsb_layernorm_backward_f32(const float* dy, const float* x, const float* gamma, float* dx, float* xhat, int64_t rows,
                                int D, float eps, void* stream);
/* q_scaled (T,H*Dk), k (B,M,H*Dk), v (B,M,H*Dv), mask (B,H,T,M) of 0/1 or NULL, keep_scale = 1/(1-p);
 * prob (B,H,T,M) softmax weights (saved for the backward), out (B,T,H*Dv) */
IPSB_API int64_t ipsb_attention_chunks(int M);   /* chunks of 256 selected patches the attention kernels split M into */
IPSB_API int ipsb_attention_train_fwd_f32(const float* q_scaled, const float* k, const float* v, const float* mask, float keep_scale,
                                          float* prob, float* out, float* scratch /* B*H*T*chunks*Dv floats */, int B, int M, int H,
                                          int Dk, int Dv, int T, void* stream);
/* dq_part (B,T,H*Dk) is per batch element (sum over B = dq); dk (B,M,H*Dk); dv (B,M,H*Dv) */
IPSB_API int ipsb_attention_train_bwd_f32(const float* q_scaled, const float* k, const float* v, const float* mask, float keep_scale,
                                          const float* prob, const float* out /* forward result */, const float* dout,
                                          float* dq_part /* (B*chunks, T*H*Dk), summed over rows by the caller */, float* dk, float* dv,
                                          int B, int M, int H, int Dk, int Dv, int T, void* stream);

/* ---------------------------------------------------------------- "bf16x3": near-fp32 arithmetic on the bf16 tensor cores
 * Same reference arithmetic as the fp32 SIMT kernels (fp32 conv2d / linear + eval BatchNorm, ips_net.py:17-60).  Every fp32
 * operand is a pair of bf16 values v = hi + lo (hi = bf16(v), lo = bf16(v - hi)); x*w = xh*wh + xh*wl + xl*wh.  Activations
 * are stored "tri": (pixels, 3C) bf16 = [hi | hi | lo], weights as [wh | wl | wh] along the input channels, so ONE launch of
 * ipsb_conv_bf16_f32out with Cin' = 3C accumulates the three products in tensor memory (exact bf16 products, fp32
 * accumulation; y = conv * scale + shift, no activation, dense bf16 in, dense fp32 out; mode 0 = generic, mode 3 = the 7x7/2
 * stem on the zero-bordered 4-channel frame, run three times on hi / lo frames).  ipsb_sum3_split adds up to three partial
 * results (+ a residual given as fp32, as a hi / lo pair, or as a tri tensor when tri_C > 0), applies ReLU and splits the
 * result again (pair, or tri tensor in out_hi when tri_C > 0; and / or stores it as fp32); ipsb_sum3_maxpool_split is the
 * stem's tail (sum, ReLU, max_pool2d(3, 2, 1)); ipsb_stage_patches_padded_split stages fp32 patches as the stem's frames. */
IPSB_API int ipsb_conv_bf16_f32out(const void* x, const void* w, const float* scale, const float* shift, float* y, int64_t P, int H, int W,
                                   int Cin, int Cout, int kh, int kw, int stride, int pad, int mode, void* stream);
IPSB_API int ipsb_sum3_split(const float* a, const float* b, const float* c, const float* res_f32, const void* res_hi, const void* res_lo,
                             int relu, void* out_hi, void* out_lo, float* out_f32, int64_t n, int tri_C, void* stream);
IPSB_API int ipsb_sum3_maxpool_split(const float* a, const float* b, const float* c, int64_t P, int H, int W, int C, int relu,
                                     void* out_hi, void* out_lo, int tri, void* stream);
IPSB_API int ipsb_stage_patches_padded_split(const float* src, int64_t first_row, int64_t n_rows, int C, int H, int W,
                                             void* dst_hi, void* dst_lo, void* stream);

/* ---------------------------------------------------------------- plan folding (one launch per weight update)
 * Replaces what eval-mode `conv -> bn` costs the reference implicitly on every ips() call (ips_net.py:191-193,209,227):
 * here the convolution weights are re-laid out for the kernels and BatchNorm(eval) is folded into per-channel scale /
 * shift ONCE per weight version.  `items_dev` is a DEVICE array of n_items descriptors (all pointers device pointers):
 * w_src (Cout, cin, kh, kw) fp32 -> w_dst in `layout` (element type bf16 if dst_bf16 else fp32), channels zero-padded to
 * cin_pad; scale_dst = bn_weight * rsqrt(bn_var + eps), shift_dst = bn_bias - bn_mean * scale_dst (skipped when
 * scale_dst is NULL; the weight part is skipped when w_dst is NULL). */
#define IPSB_FOLD_KMAJOR 0    /* (Cout, kh*kw*cin_pad), k = (r*kw+s)*cin_pad + c : tcgen05 kernels            */
#define IPSB_FOLD_KN 1        /* (kh*kw*cin_pad, Cout)                            : fp32 SIMT kernels         */
#define IPSB_FOLD_STEM_S2D 2  /* (Cout, 256), k = (a*4+b)*16 + (dy*2+dx)*4 + c = w[2a+dy-1, 2b+dx-1, c]        */
#define IPSB_FOLD_STEM_8X8 3  /* (Cout, 256), k = r*32 + (s+1)*4 + c                                           */
typedef struct ipsb_fold_item {
    const float* w_src;
    void* w_dst;
    const float* bn_weight; const float* bn_bias; const float* bn_mean; const float* bn_var;
    float* scale_dst; float* shift_dst;
    int64_t dst_elems;
    int32_t cout, cin, kh, kw, cin_pad, layout, dst_bf16;
    float eps;
} ipsb_fold_item;
IPSB_API int ipsb_fold_plan(const ipsb_fold_item* items_dev, int n_items, int blocks_per_item, void* stream);

/* ---------------------------------------------------------------- feature projector + score projection, fused
 * Replaces, in the no-grad pass of ips() on feature bags: LayerNorm(no affine) -> Linear -> BatchNorm1d(eval) -> ReLU
 * (ips_net.py:54-60) followed by the key projection + query dot product (transformer.py:71-83, folded into U):
 *     z[r, :] = ReLU(rstd_r * (a_n * (x_r . W_n) - mean_r * b_n) + shift_n) @ U
 * x (rows, K) fp32 or bf16 (x_is_bf16), rounded to bf16 on chip; mean_r / rstd_r are the LayerNorm statistics of the
 * rounded row (fp32, eps); w_bf16 (N, K) bf16 K-major; table (N, 12) fp32 per output column n:
 *     [a_n = BN scale, b_n = a_n * sum_k w_bf16[n, k], shift_n = bias_n * a_n + BN shift, 0, U[n, 0..7]]  (U zero-padded to 8);
 * z (rows, HT) fp32, HT <= 8.  N in {256, 512} (the 128 x N fp32 tile fills tensor memory), K % 128 == 0, K >= 256.
 * Only the features (once) and 4*HT bytes per row of logits touch HBM. */
IPSB_API int ipsb_projector_logits(const void* x, int x_is_bf16, const void* w_bf16, const float* table, float* z,
                                   int64_t rows, int K, int N, int HT, float eps, void* stream);
/* Scan-order form, the producer half of the streamed selection (ips_net.py:200-241: the reference shuffles the bag, then
 * embeds chunk after chunk inside the loop -- here the rows are READ through the shuffle and the loop runs beside the
 * projector).  Row g of z is the logits of feature row (g / rows_per_batch) * rows_per_batch + perm[(g / rows_per_batch) *
 * perm_batch_stride + g % rows_per_batch] (perm int64, NULL = identity; perm_batch_stride = rows_per_batch for
 * per-instance orders, 0 for a shared one).  tile_flags (ceil(rows / 128) ints, zeroed by the caller, or NULL): flag t is
 * set (release, device scope) once rows [128 t, 128 t + 128) of z are written; tiles are handed to the CTAs in rounds, so
 * with all CTAs co-resident they complete in order.  max_ctas > 0 caps the grid (leaves SMs to the loop's cluster). */
IPSB_API int ipsb_projector_logits_scan(const void* x, int x_is_bf16, const void* w_bf16, const float* table, float* z,
                                        int64_t rows, int K, int N, int HT, float eps,
                                        const int64_t* perm, int64_t perm_batch_stride, int64_t rows_per_batch,
                                        int* tile_flags, int max_ctas, void* stream);

/* ---------------------------------------------------------------- per-launch timing of the native executor
 * (measurement only, bench.py's roofline; the reference's counterpart is the track_efficiency bracket,
 * training/iterative.py:129-132,166-171).  Between begin and end every kernel ipsb_resnet_logits* issues is bracketed by
 * CUDA events on the internal stream ("lane") it runs on; ipsb_profile_end returns, per launch, the kernel family
 * (0 stage, 1 stem(+pool), 2 BasicBlock convolution, 3 pooling, 4 logits) and its start / stop time in ms relative to the
 * event recorded on `stream` by ipsb_profile_begin -- lanes overlap, so a family's busy time is the UNION of its
 * intervals.  n_out = records written (at most max_records). */
IPSB_API int ipsb_profile_begin(void* stream);
IPSB_API int ipsb_profile_end(int max_records, int* kinds, float* start_ms, float* stop_ms, int* n_out);

/* ---------------------------------------------------------------- gathers
 * Replaces torch.gather(patches, 1, mem_idx expanded) (ips_net.py:244-247) and the
 * pos-enc gather (:249-250): dst[b,m,:] = src[b*src_batch_stride + idx[b,m], :],
 * rows of row_bytes bytes (multiple of 4; 16-byte vector path when aligned).  A negative index writes a zero
 * row (used by the sequence-sharded gather, where a rank owns only part of the winners). */
IPSB_API int ipsb_gather_rows(const void* src, int64_t src_batch_stride_rows, const int64_t* idx,
                     int B, int M, int64_t row_bytes, void* dst, void* stream);

/* ---------------------------------------------------------------- sequence-sharded selection over NVLink peer memory
 * New work (the reference is single-GPU; SURVEY.md 8e, north_star "each rank scans its slice and keeps a local top-M,
 * the ranks then merge ... M (score, index) candidates plus a global top-M").  One process per GPU; every rank owns an
 * EXCHANGE BUFFER (allocated by PyTorch, zero-initialised) whose CUDA-IPC handle the host side all-gathers once; after
 * ipsb_peer_open every rank holds a device address for every peer's buffer and the kernels below move data with plain
 * stores over NVLink / NVSwitch -- no NCCL call, no host round trip, CUDA-graph capturable.
 *
 * Exchange-buffer layout (identical on all ranks): a 4096-byte header (per-phase epoch counters of the owner, block
 * completion counters, a status word, and flags[phase][rank] written by the peers), then caller-defined sections
 * addressed by byte offsets.  A push kernel writes its payload into the same section of every destination rank, and
 * its last block (device-wide completion counter) publishes epoch+1 in flags[phase][own rank] of every peer with a
 * system-scope release store.  ipsb_peer_wait spins (system-scope acquire loads, bounded: a peer that never arrives
 * sets status != 0 instead of hanging the GPU) until all ranks' flags of that phase reached the owner's epoch.
 * Phases: 0 = candidates / logit slices, 1 = winners. */
#define IPSB_MAX_PEERS 8
#define IPSB_PEER_HEADER_BYTES 4096
typedef struct ipsb_peer_ctx {
    int32_t rank, world;
    void* base[IPSB_MAX_PEERS];   /* address IN THIS PROCESS of every rank's exchange buffer (base[rank] = own) */
} ipsb_peer_ctx;
/* CUDA-IPC export of a device pointer inside a cudaMalloc'ed allocation (PyTorch's caching allocator without
 * expandable segments): handle of the allocation + byte offset of `ptr` in it. */
IPSB_API int ipsb_peer_export(const void* ptr, unsigned char handle_out[64], int64_t* offset_out);
/* Maps a peer's allocation (lazy peer access) and returns the address of its exchange buffer in this process. */
IPSB_API int ipsb_peer_open(const unsigned char handle[64], int64_t offset, void** ptr_out);
IPSB_API int ipsb_peer_close(void* ptr, int64_t offset);
/* status word of the own header (0 = ok, 1 = a wait timed out); synchronises `stream`. */
IPSB_API int ipsb_peer_status(const ipsb_peer_ctx* ctx, int* status_out, void* stream);
/* Candidates of this rank -> slot block [slot0, slot0 + m) of EVERY rank's candidate sections:
 *   cz (B, L, HT) fp32 at byte offset cz_off:  cz[b, slot0 + j, :] = z_local[b, cand[b, j], :]
 *   ci (B, L) int64   at byte offset ci_off:  ci[b, slot0 + j]    = index_base + cand[b, j]
 * z_local (B, n_local, HT) is this rank's slice of the logit table, cand (B, m) local indices (best first).
 * Replaces the two all-gathers (logits, indices) of the candidate merge; signals phase 0. */
IPSB_API int ipsb_peer_push_candidates(const ipsb_peer_ctx* ctx, const float* z_local, int64_t n_local, const int64_t* cand,
                                       int B, int m, int HT, int64_t index_base, int64_t L, int64_t slot0,
                                       int64_t cz_off, int64_t ci_off, void* stream);
/* Same, with the logits of candidate (b, j) read from row rows[b, j] of z_local instead of row cand[b, j] (a table kept in
 * scan order by the streamed selection: rows = the winners' scan positions, cand = their original indices). */
IPSB_API int ipsb_peer_push_candidates_rows(const ipsb_peer_ctx* ctx, const float* z_local, int64_t n_local, const int64_t* cand,
                                            const int64_t* rows, int B, int m, int HT, int64_t index_base, int64_t L,
                                            int64_t slot0, int64_t cz_off, int64_t ci_off, void* stream);
/* This rank's slice (B, n_local, HT) of the logit table -> rows [row0, row0 + n_local) of every rank's full table
 * (B, N, HT) at byte offset z_off ('exact' mode: the loop is replicated); signals phase 0. */
IPSB_API int ipsb_peer_push_logits(const ipsb_peer_ctx* ctx, const float* z_local, int B, int64_t n_local, int HT,
                                   int64_t N, int64_t row0, int64_t z_off, void* stream);
/* Small all-gather over peer memory (synchronised BatchNorm statistics of the data-parallel train step; the
 * reference is single-process): `bytes` (16-byte multiple) of `src` -> every rank; dst (world x bytes, local) receives all
 * ranks' data in rank order.  The section at sec_off holds two slots of slot_stride >= world * bytes that alternate with the
 * device-side epoch of `phase` (safe under CUDA-graph replay).  A tens-of-microseconds NCCL collective becomes two small
 * kernels (push + flag, wait + copy). */
/* Synchronised BatchNorm over peer memory (data-parallel train step, equal rows per rank).  Forward: push this
 * rank's [mean | var] (2 * cols floats), wait for every rank's, and in the waiting kernel combine them (mean of means,
 * mean of var_r + (mean_r - mean)^2), write mean / rstd and update the running statistics.  Backward: push [sum g | sum g*xhat]
 * (n floats), wait, out = the sum over the ranks in rank order.  Two launches per exchange; sections / phases as for
 * ipsb_peer_allgather_small. */
IPSB_API int ipsb_peer_bn_forward(const ipsb_peer_ctx* ctx, const float* mean_var, int cols, int64_t sec_off, int64_t slot_stride,
                                  int phase, float momentum, float unbias, float eps, float* mean_out, float* rstd_out,
                                  float* running_mean, float* running_var, void* stream);
IPSB_API int ipsb_peer_allgather_sum(const ipsb_peer_ctx* ctx, const float* src, int n, int64_t sec_off, int64_t slot_stride, int phase,
                                     float* out, void* stream);
IPSB_API int ipsb_peer_allgather_small(const ipsb_peer_ctx* ctx, const void* src, int64_t bytes, int64_t sec_off, int64_t slot_stride,
                                       int phase, void* dst, void* stream);
/* Waits until every rank has signalled `phase` for the current epoch. */
IPSB_API int ipsb_peer_wait(const ipsb_peer_ctx* ctx, int phase, void* stream);
/* Winners -> output section (B_out, M, row_bytes) at byte offset out_off of the destination ranks.  win (B, M) holds
 * either global patch indices (ci == NULL) or positions into ci (B, L) (the merged candidate list); the global index
 * is also written to idx_out (B, M) on this rank.  A rank copies the rows it owns (row0 <= index < row0 + n_local) out
 * of local_rows (B, n_local, row_bytes), one 128-bit load and up to `world` 128-bit peer stores per 16 bytes.
 * slides_per_rank == 0: every rank receives all B slides (replicated, like IPSNet.ips on every rank);
 * slides_per_rank > 0: slide b goes only to rank b / slides_per_rank, row b % slides_per_rank of its section (the
 * data-parallel train step that follows needs each slide on one rank only).  Signals phase 1.
 * Replaces the all-reduce of zero-padded winners of round 1 (reference gather: ips_net.py:244-247). */
IPSB_API int ipsb_peer_push_winners(const ipsb_peer_ctx* ctx, const void* local_rows, int64_t n_local, int64_t row0,
                                    const int64_t* win, const int64_t* ci, int64_t L, int B, int M, int64_t row_bytes,
                                    int slides_per_rank, int64_t out_off, int64_t* idx_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IPS_B200_H */
