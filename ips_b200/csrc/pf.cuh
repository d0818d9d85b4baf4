// "Padded-flat" (PF) channels-last activation layout.
//
// P patches of H x W pixels with C channels are stored as rows of C elements:
//     row(p, y, x) = G0 + p*Sp + y*Wp + x,     Wp = W + 1, Sp = (H + 1) * Wp, G0 = Wp + 1
// i.e. every image row is followed by ONE zero pixel and every patch by ONE zero row (plus a zero
// lead-in of G0 rows).  All eight neighbours of a pixel -- including the ones a 3x3 convolution's zero
// padding would supply -- are then found at the fixed row offsets dy*Wp + dx, so the nine im2col
// operands of a stride-1 3x3 convolution are nine SHIFTED VIEWS of one contiguous block of rows:
// the block is loaded into shared memory once and tcgen05.mma descriptors are pointed at
// base + (r*Wp + s)*128 bytes (the 128B swizzle is a function of the absolute address, so any
// 128-byte-granular start is valid).  Kernels that write PF tensors keep the pad rows zero.
#pragma once
#include <stdint.h>

namespace pf {

struct Geo {
    int H, W, Wp, Sp, G0;
    int64_t rows;          // allocated rows for P patches
};

__host__ __device__ inline Geo make(int64_t P, int H, int W) {
    Geo g;
    g.H = H; g.W = W; g.Wp = W + 1; g.Sp = (H + 1) * (W + 1); g.G0 = W + 2;
    g.rows = (g.G0 + P * g.Sp + g.Wp + 2 + 7) / 8 * 8;
    return g;
}
__host__ __device__ inline int64_t row(const Geo& g, int64_t p, int y, int x) {
    return g.G0 + p * g.Sp + (int64_t)y * g.Wp + x;
}

}  // namespace pf
