// Row-halo 3x3 convolution for RMVPE's 16-channel, W = 128 levels (conv2d_row.cu).
#pragma once
#include "common.cuh"

namespace rvcb {

struct Conv2dRowArgs {
    const __half* x; long ldx; int H, W, cin, cout;     // x: [H, W, ldx] channels last, first `cin` channels used
    const __half* w; int w_rows, w_cols;                // packed [cout_pad, 9*cin] (dh, dw, ci) order, K contiguous (pack_conv2d_3x3, bk = 16)
    const float* bias; bool relu;
    const float* res2; long ldres2;                     // fp32 residual added after the activation (nullable)
    float* out32; long ld32; __half* out16; long ld16;  // either may be null
};
// Launches the kernel and returns true when the shape qualifies (W = 128, cin = cout = 16, aligned operands); false otherwise.
bool conv2d_row_try(const Conv2dRowArgs& a, cudaStream_t stream);

}  // namespace rvcb
