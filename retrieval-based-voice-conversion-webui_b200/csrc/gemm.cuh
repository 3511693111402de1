// Implicit-GEMM engine: C[M,N] = epilogue( sum_seg A_shift(seg)[M,K_seg] * B[N, K_seg]^T )
//
// One kernel family serves every GEMM-shaped op of the RVC hot path:
//   * Linear / 1x1 conv                      1 segment
//   * Conv1d (k taps, dilation d, stride 1)  k segments, A rows shifted by j*d - pad; zero padding =
//                                            TMA out-of-bounds fill (im2col-free)
//   * Conv1d stride 2 (HuBERT extractor)     A viewed as [T/2, 2C]; 2 segments
//   * ConvTranspose1d (k <= 2*stride)        polyphase: N = stride*C_out, 3 row-shift segments
//   * Conv2d 3x3 / ConvTranspose2d (RMVPE)   3-D tensor map (C, W, H), segments carry (dh, dw)
//   * attention QK^T, PV, V^T projection     batched over heads through grid-z strides
//
// A: fp16 activations, channels-last (K contiguous).  B: fp16 packed weights [N_pad, K_total],
// K contiguous, K_total = sum of segment lengths (each padded to a multiple of BLOCK_K).
// Accumulation fp32 (TMEM).  Epilogue (all fp32):
//     t = acc + bias[n] (+ res1[m,n]);  t = act1(t);  v = alpha*t (+ res2[m,n]);
//     out32[m,n] = v;   out16[m,n] = (half) act2(v)
//   gate mode (WaveNet):  columns come in (tanh, sigmoid) pairs -> out[m, n/2] = tanh(t0)*sigmoid(t1)
#pragma once
#include "common.cuh"

namespace rvcb {

constexpr int GEMM_MAX_SEG = 128;

struct GemmSeg {
    int row_off;   // A row shift (1-D mode) or dh (2-D mode)
    int col_off;   // A column (channel) start of this segment
    int dw;        // 2-D mode only: shift along W
    int nk;        // number of BLOCK_K chunks in this segment
};

struct GemmArgs {
    // ---- A operand (fp16) ----
    const __half* A = nullptr;
    long lda = 0;            // elements between consecutive rows (1-D) / pixels (2-D)
    int a_rows = 0;          // 1-D: valid rows (OOB rows read as zero).  2-D: H
    int a_cols = 0;          // valid columns (channels)
    int conv2d_W = 0;        // 0 = 1-D mode; else W of the [H, W, C] image (W in {4..128}, power of 2)
    // ---- B operand (fp16 packed weights / activations) ----
    const __half* B = nullptr;
    long ldb = 0;
    int b_rows = 0, b_cols = 0;   // valid extents (OOB -> zero)
    // ---- problem ----
    int M = 0, N = 0;
    int block_k = 64;        // 64 (SW128), 32 (SW64) or 16 (SW32)
    int nseg = 0;
    GemmSeg seg[GEMM_MAX_SEG];
    // ---- batching over grid z (heads / groups) ----
    int batch = 1;
    long a_row_z = 0, a_col_z = 0, b_row_z = 0, b_col_z = 0, c_z = 0, bias_z = 0;
    int b_col0 = 0;
    // ---- epilogue ----
    const float* bias = nullptr;
    int bias_per_row = 0;
    const float* res1 = nullptr; long ldres1 = 0;
    const float* res2 = nullptr; long ldres2 = 0;
    float alpha = 1.f;
    int act1 = ACT_NONE; float act1_p = 0.f;
    int act2 = ACT_NONE; float act2_p = 0.f;
    int gate = 0;
    float* out32 = nullptr; long ld32 = 0;
    __half* out16 = nullptr; long ld16 = 0;
    // ConvTranspose2d 2x polyphase scatter: N = 4*up2_C ordered (a, b, co); output pixel
    // (2i+a, 2j+b) of a [2H, 2W, ld] image.  0 = off.
    int up2_C = 0;
};

// Runs on the tcgen05/TMA kernel (the product path).
void gemm_tc(const GemmArgs& g, cudaStream_t stream);
// Plain SIMT restatement of the same contract (validation of the tensor-core kernel in tests;
// never used by the model graphs unless RVCB_GEMM=simt is set for debugging).
void gemm_simt(const GemmArgs& g, cudaStream_t stream);
// Dispatch used by the model graphs.
void gemm(const GemmArgs& g, cudaStream_t stream);

void gemm_prof_begin();
void gemm_prof_end(double* ms_total, unsigned long long* launches);
// per-class totals of the last profiled region: [0] streaming kernel, [1] weight-stationary kernel, [2] fused residual block
void gemm_prof_classes(double ms[3], double launches[3], double flops[3], double bytes[3]);

// helpers to fill segments
inline void seg_linear(GemmArgs& g, int K) {
    g.nseg = 1;
    g.seg[0] = {0, 0, 0, ceil_div(K, g.block_k)};
}
// stride-1 conv1d: k taps at rows t + j*dil - pad
inline void seg_conv1d(GemmArgs& g, int C_in, int k, int dil, int pad) {
    RVCB_CHECK(k <= GEMM_MAX_SEG, "too many taps");
    g.nseg = k;
    for (int j = 0; j < k; ++j) g.seg[j] = {j * dil - pad, 0, 0, ceil_div(C_in, g.block_k)};
}
inline void seg_conv2d_3x3(GemmArgs& g, int C_in) {
    g.nseg = 9;
    for (int dh = 0; dh < 3; ++dh)
        for (int dw = 0; dw < 3; ++dw) g.seg[dh * 3 + dw] = {dh - 1, 0, dw - 1, ceil_div(C_in, g.block_k)};
}
inline int gemm_total_k(const GemmArgs& g) {
    int t = 0;
    for (int s = 0; s < g.nseg; ++s) t += g.seg[s].nk * g.block_k;
    return t;
}

}  // namespace rvcb
