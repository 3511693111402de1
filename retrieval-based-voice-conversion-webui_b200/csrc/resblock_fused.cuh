// Fused ResBlock1 kernel of the NSF-HiFi-GAN vocoder (resblock_fused.cu): one launch = x -> x + sum of three (c1, c2) pairs.
#pragma once
#include "common.cuh"
#include "weights.cuh"

namespace rvcb {

struct RBFusedWeights {
    PackedB w;                  // [C, 6*k*C] fp16: convs in execution order (c1_0, c2_0, c1_1, c2_1, c1_2, c2_2), tap-major, K contiguous
    float* bias_tab = nullptr;  // [7][C] fp32: bias applied when the accumulator is read at epilogue step e (see pack_resblock_fused)
    int C = 0, k = 0, dil[3] = {1, 1, 1};
    int halo() const {
        int h = 0;
        for (int i = 0; i < 3; ++i) h += (k - 1) / 2 * (dil[i] + 1);
        return h;
    }
};

RBFusedWeights pack_resblock_fused(DevOwner& own, int C, int k, const int* dil, const float* const* w1, const float* const* b1,
                                   const float* const* w2, const float* const* b2);
bool resblock_fused_supported(int C, int k, const int* dil);
// rows the output buffer must hold (= T; -1 if the configuration has no fused variant)
long resblock_fused_out_rows(int C, int k, const int* dil, int T);
int resblock_fused_tile_rows(int C);
// y[0:T, :] = ResBlock1(x[0:T, :]);  x, y fp32 [rows, C] dense (ld = C), y with T rows
void resblock_fused(const RBFusedWeights& w, const float* x, float* y, int T, cudaStream_t stream);
// out[t, c] = (half) lrelu((y[0] + .. + y[nk-1])[t, c] / nk, slope), out row stride ld
void resblock_mean_lrelu(const float* const* y, int nk, long T, int C, __half* out, long ld, float slope, cudaStream_t stream);

}  // namespace rvcb
