// Fused multi-head self-attention (attn_fused.cu): scores in TMEM, probabilities in shared memory, one launch per layer.
#pragma once
#include "common.cuh"

namespace rvcb {

struct AttnFusedArgs {
    const __half* q = nullptr; long ldq = 0;     // [T, heads*dk] queries; head h at columns h*dk (zero-padded up to dk)
    const __half* k = nullptr; long ldk = 0;     // [T, heads*dk] keys
    const __half* vT = nullptr; long ldv = 0;    // [heads*dv, >= T] values, channel-major (time contiguous); head h at rows h*dv
    int T = 0, heads = 0;
    int dk = 64, dv = 64;                        // (64, 64): HuBERT; (128, 96) + qrel/ev: TextEncoder relative-position attention
    float qscale = 1.f;                          // multiplies q.k (1 when the scaling is folded into the q projection)
    const float* qrel = nullptr;                 // [heads, T, 32] fp32: (q * qscale) E_k^T, columns 0..20 = offsets -10..+10
    const float* ev = nullptr;                   // [21, dv] fp32: E_v
    __half* out = nullptr; long ldo = 0;         // [T, heads*dv] context
};
bool attention_fused_supported(const AttnFusedArgs& a);
void attention_fused(const AttnFusedArgs& a, cudaStream_t stream);

}  // namespace rvcb
