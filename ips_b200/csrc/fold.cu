// Plan folding in ONE launch: after every optimizer step ips() needs the convolution weights in the kernels' operand
// layouts and the eval-mode BatchNorm folded into per-channel scale / shift (architecture/ips_net.py:17-52 runs conv -> bn
// in eval mode inside `ips`, :191-193).  Round 1 rebuilt this with ~160 ATen launches (permute, reshape, cast, rsqrt, mul,
// sub per layer); here a device-resident table of items describes every layer and one kernel folds them all.
#include "common.cuh"
#include "../../include/ips_b200.h"

namespace {

using bf16 = __nv_bfloat16;

__device__ __forceinline__ float w_at(const ipsb_fold_item& it, int co, int c, int r, int s) {
    if (c >= it.cin || r < 0 || s < 0 || r >= it.kh || s >= it.kw) return 0.f;
    return it.w_src[(((int64_t)co * it.cin + c) * it.kh + r) * it.kw + s];
}

// grid = (blocks per item, items)
__global__ void __launch_bounds__(256) fold_plan_kernel(const ipsb_fold_item* __restrict__ items) {
    const ipsb_fold_item it = items[blockIdx.y];
    // ---- BatchNorm (eval): scale = gamma / sqrt(var + eps), shift = beta - mean * scale
    if (blockIdx.x == 0 && it.scale_dst) {
        for (int c = threadIdx.x; c < it.cout; c += blockDim.x) {
            const float sc = it.bn_weight[c] * rsqrtf(it.bn_var[c] + it.eps);
            it.scale_dst[c] = sc;
            it.shift_dst[c] = it.bn_bias[c] - it.bn_mean[c] * sc;
        }
    }
    if (!it.w_dst) return;
    const int64_t n = it.dst_elems;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float v;
        if (it.layout == IPSB_FOLD_KMAJOR) {            // (Cout, kh*kw*cin_pad): k = (r*kw + s)*cin_pad + c
            const int K = it.kh * it.kw * it.cin_pad;
            const int co = (int)(i / K), k = (int)(i - (int64_t)co * K);
            const int c = k % it.cin_pad, rs = k / it.cin_pad;
            v = w_at(it, co, c, rs / it.kw, rs % it.kw);
        } else if (it.layout == IPSB_FOLD_KN) {         // (kh*kw*cin_pad, Cout): row k as above
            const int co = (int)(i % it.cout);
            const int k = (int)(i / it.cout);
            const int c = k % it.cin_pad, rs = k / it.cin_pad;
            v = w_at(it, co, c, rs / it.kw, rs % it.kw);
        } else if (it.layout == IPSB_FOLD_STEM_S2D) {   // (Cout, 256): k = (a*4+b)*16 + (dy*2+dx)*4 + c = w[2a+dy-1, 2b+dx-1, c]
            const int co = (int)(i >> 8), k = (int)(i & 255);
            const int c = k & 3, dx = (k >> 2) & 1, dy = (k >> 3) & 1, b = (k >> 4) & 3, a = k >> 6;
            v = w_at(it, co, c, 2 * a + dy - 1, 2 * b + dx - 1);
        } else {                                        // IPSB_FOLD_STEM_8X8: (Cout, 256): k = r*32 + (s+1)*4 + c
            const int co = (int)(i >> 8), k = (int)(i & 255);
            const int c = k & 3, sp = (k >> 2) & 7, r = k >> 5;
            v = w_at(it, co, c, r, sp - 1);
        }
        if (it.dst_bf16) reinterpret_cast<bf16*>(it.w_dst)[i] = __float2bfloat16(v);
        else             reinterpret_cast<float*>(it.w_dst)[i] = v;
    }
}

}  // namespace

extern "C" int ipsb_fold_plan(const ipsb_fold_item* items_dev, int n_items, int blocks_per_item, void* stream) {
    IPSB_REQUIRE(items_dev != nullptr && n_items > 0 && n_items <= 65535 && blocks_per_item > 0, "fold_plan: bad arguments");
    fold_plan_kernel<<<dim3((unsigned)blocks_per_item, (unsigned)n_items), 256, 0, (cudaStream_t)stream>>>(items_dev);
    IPSB_LAUNCH_CHECK();
    return 0;
}
