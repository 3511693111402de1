// Fused multi-head self-attention (attn_fused.cu): scores in TMEM, probabilities in shared memory, one launch per layer.
#pragma once
#include "common.cuh"

namespace rvcb {

struct AttnFusedArgs {
    const __half* q = nullptr; long ldq = 0;     // [T, heads*64] queries, softmax scaling already applied; head h at columns h*64
    const __half* k = nullptr; long ldk = 0;     // [T, heads*64] keys
    const __half* vT = nullptr; long ldv = 0;    // [heads*64, >= T] values, channel-major (time contiguous)
    int T = 0, heads = 0, dh = 64;
    __half* out = nullptr; long ldo = 0;         // [T, heads*64] context
};
bool attention_fused_supported(const AttnFusedArgs& a);
void attention_fused(const AttnFusedArgs& a, cudaStream_t stream);

}  // namespace rvcb
