// HuBERT-base content-feature extractor on the implicit-GEMM engine.
// Replaces fairseq HubertModel.extract_features at infer/modules/vc/pipeline.py:102-110,
// infer/lib/rtrvc.py:154-162 (encoder loop documented in rvc/hubert.py:27-91).
//
// Data layout in HBM: activations channels-last [T, C]; fp16 copies feed the tensor cores,
// the residual stream stays fp32.  The stride-2 extractor convs read the previous layer as a
// [T/2, 2C] view so that stride disappears into the K dimension (no im2col).
#include "../../include/rvcb200.h"
#include "api_macros.h"
#include "attn_fused.cuh"
#include "gemm.cuh"
#include "kernels.cuh"
#include "weights.cuh"

using namespace rvcb;

struct rvcb_hubert {
    DevOwner own;
    Arena arena;
    // conv feature extractor
    float* conv0_w = nullptr;                 // [512,10] fp32
    float *gn_g = nullptr, *gn_b = nullptr;
    PackedB convs[6];                         // conv1..6
    float *ln0_g = nullptr, *ln0_b = nullptr;
    PackedB proj; float* proj_b = nullptr;
    PackedB posconv; float* pos_b = nullptr;
    float *eln_g = nullptr, *eln_b = nullptr;
    struct Layer {
        PackedB wqk, wv, wo, w1, w2;
        float *bqk, *bv, *bo, *b1, *b2, *ln1_g, *ln1_b, *ln2_g, *ln2_b;
    };
    std::vector<Layer> layers;
    bool has_final = false;
    PackedB wfinal; float* bfinal = nullptr;
};

static void conv_lens(int n, int* T) {
    static const int K[7] = {10, 3, 3, 3, 3, 2, 2}, S[7] = {5, 2, 2, 2, 2, 2, 2};
    for (int i = 0; i < 7; ++i) {
        n = (n - K[i]) / S[i] + 1;
        T[i] = n;
    }
}

static rvcb_hubert* hubert_build(const rvcb_weights& w) {
    auto* h = new rvcb_hubert();
    try {
        DevOwner& own = h->own;
        h->conv0_w = own.upload(w.get("feature_extractor.conv_layers.0.0.weight").data);
        h->gn_g = own.upload(w.get("feature_extractor.conv_layers.0.2.weight").data);
        h->gn_b = own.upload(w.get("feature_extractor.conv_layers.0.2.bias").data);
        for (int i = 1; i < 7; ++i) {
            const WT& t = w.get("feature_extractor.conv_layers." + std::to_string(i) + ".0.weight");
            h->convs[i - 1] = pack_conv1d(own, t.data.data(), 512, 512, (int)t.dim(2));
        }
        h->ln0_g = own.upload(w.get("layer_norm.weight").data);
        h->ln0_b = own.upload(w.get("layer_norm.bias").data);
        h->proj = pack_linear(own, w.get("post_extract_proj.weight").data.data(), 768, 512);
        h->proj_b = own.upload(w.get("post_extract_proj.bias").data);
        {   // pos_conv: weight_norm(dim=2): w = g[k] * v / ||v[:, :, k]||
            const WT& v = w.get("encoder.pos_conv.0.weight_v");
            const WT& g = w.get("encoder.pos_conv.0.weight_g");
            const int CO = 768, CG = 48, KK = 128, G = 16;
            RVCB_CHECK(v.dim(0) == CO && v.dim(1) == CG && v.dim(2) == KK, "pos_conv shape");
            std::vector<double> nrm(KK, 0.0);
            for (int co = 0; co < CO; ++co)
                for (int ci = 0; ci < CG; ++ci)
                    for (int k = 0; k < KK; ++k) {
                        const double x = v.data[((size_t)co * CG + ci) * KK + k];
                        nrm[k] += x * x;
                    }
            // packed [G*64, KK*64]: row g*64+co_l, col k*64+ci
            std::vector<float> hB((size_t)G * 64 * KK * 64, 0.f);
            for (int co = 0; co < CO; ++co) {
                const int gi = co / CG, col = co % CG;
                for (int ci = 0; ci < CG; ++ci)
                    for (int k = 0; k < KK; ++k)
                        hB[((size_t)(gi * 64 + col) * KK + k) * 64 + ci] =
                            (float)(v.data[((size_t)co * CG + ci) * KK + k] * (g.data[k] / std::sqrt(nrm[k])));
            }
            h->posconv = upload_half(own, hB, G * 64, KK * 64);
            h->pos_b = own.upload(w.get("encoder.pos_conv.0.bias").data);
        }
        h->eln_g = own.upload(w.get("encoder.layer_norm.weight").data);
        h->eln_b = own.upload(w.get("encoder.layer_norm.bias").data);
        for (int i = 0; i < 12; ++i) {
            const std::string p = "encoder.layers." + std::to_string(i) + ".";
            if (!w.has(p + "fc1.weight")) break;
            rvcb_hubert::Layer L{};
            std::vector<float> wqk((size_t)1536 * 768), bqk(1536);
            const WT& wq = w.get(p + "self_attn.q_proj.weight");
            const WT& wk = w.get(p + "self_attn.k_proj.weight");
            const WT& bq = w.get(p + "self_attn.q_proj.bias");
            const WT& bk = w.get(p + "self_attn.k_proj.bias");
            for (size_t j = 0; j < (size_t)768 * 768; ++j) {
                wqk[j] = wq.data[j] * 0.125f;                 // q scaling 64^-0.5 (exact power of two)
                wqk[(size_t)768 * 768 + j] = wk.data[j];
            }
            for (int j = 0; j < 768; ++j) {
                bqk[j] = bq.data[j] * 0.125f;
                bqk[768 + j] = bk.data[j];
            }
            L.wqk = pack_linear(own, wqk.data(), 1536, 768);
            L.bqk = own.upload(bqk);
            L.wv = pack_linear(own, w.get(p + "self_attn.v_proj.weight").data.data(), 768, 768);
            L.bv = own.upload(w.get(p + "self_attn.v_proj.bias").data);
            L.wo = pack_linear(own, w.get(p + "self_attn.out_proj.weight").data.data(), 768, 768);
            L.bo = own.upload(w.get(p + "self_attn.out_proj.bias").data);
            L.w1 = pack_linear(own, w.get(p + "fc1.weight").data.data(), 3072, 768);
            L.b1 = own.upload(w.get(p + "fc1.bias").data);
            L.w2 = pack_linear(own, w.get(p + "fc2.weight").data.data(), 768, 3072);
            L.b2 = own.upload(w.get(p + "fc2.bias").data);
            L.ln1_g = own.upload(w.get(p + "self_attn_layer_norm.weight").data);
            L.ln1_b = own.upload(w.get(p + "self_attn_layer_norm.bias").data);
            L.ln2_g = own.upload(w.get(p + "final_layer_norm.weight").data);
            L.ln2_b = own.upload(w.get(p + "final_layer_norm.bias").data);
            h->layers.push_back(L);
        }
        if (w.has("final_proj.weight")) {
            h->has_final = true;
            h->wfinal = pack_linear(own, w.get("final_proj.weight").data.data(), 256, 768);
            h->bfinal = own.upload(w.get("final_proj.bias").data);
        }
    } catch (...) {
        delete h;
        throw;
    }
    return h;
}

static void hubert_forward(rvcb_hubert* h, const float* d_wav, int n, int output_layer, float* d_out, int* n_frames, cudaStream_t st) {
    int Tl[7];
    conv_lens(n, Tl);
    RVCB_CHECK(Tl[6] >= 1, "hubert: input too short");
    RVCB_CHECK(output_layer >= 1 && output_layer <= (int)h->layers.size(), "hubert: bad output_layer");
    const int T0 = Tl[0], T = Tl[6];
    const int Tp = round_up(T, 8);
    // ---- arena sizing ----
    size_t need = 0;
    auto rnd = [](size_t b) { return (b + 1023) & ~size_t(1023); };
    need += rnd((size_t)T0 * 512 * 4) + rnd(1024 * 8);
    for (int i = 0; i < 7; ++i) need += rnd(((size_t)Tl[i] + 2) * 512 * 2);
    need += rnd((size_t)T * 512 * 4) + rnd((size_t)T * 512 * 2);
    need += 3 * rnd((size_t)T * 768 * 4) + 2 * rnd((size_t)T * 768 * 2);
    need += rnd((size_t)T * 1536 * 2) + rnd((size_t)768 * Tp * 2) + rnd((size_t)12 * T * Tp * 4) + rnd((size_t)12 * T * Tp * 2);
    need += rnd((size_t)T * 3072 * 2);
    need += 1 << 20;
    h->arena.reserve(need);
    h->arena.reset();
    Arena& ar = h->arena;

    // ---- conv feature extractor ----
    float* y0 = ar.alloc<float>((size_t)T0 * 512);
    double* stats = ar.alloc<double>(1024);
    __half* feat[7];
    for (int i = 0; i < 7; ++i) feat[i] = ar.alloc<__half>(((size_t)Tl[i] + 2) * 512);
    hubert_conv0_gn_gelu(d_wav, n, h->conv0_w, h->gn_g, h->gn_b, y0, stats, feat[0], T0, st);
    for (int i = 1; i < 7; ++i) {
        const int k = (i <= 4) ? 3 : 2;
        const int Tin = Tl[i - 1];
        GemmArgs g;
        g.A = feat[i - 1]; g.lda = 1024; g.a_rows = (Tin + 1) / 2; g.a_cols = 1024;
        g.B = h->convs[i - 1].d; g.ldb = h->convs[i - 1].cols; g.b_rows = h->convs[i - 1].rows; g.b_cols = h->convs[i - 1].cols;
        g.M = Tl[i]; g.N = 512;
        if (k == 3) {
            g.nseg = 2;
            g.seg[0] = {0, 0, 0, 16};
            g.seg[1] = {1, 0, 0, 8};
        } else {
            g.nseg = 1;
            g.seg[0] = {0, 0, 0, 16};
        }
        g.act1 = ACT_GELU;
        g.out16 = feat[i]; g.ld16 = 512;
        gemm(g, st);
    }
    // ---- LN(512) -> proj -> pos_conv -> LN(768) ----
    __half* ln16 = ar.alloc<__half>((size_t)T * 512);
    float* featf = ar.alloc<float>((size_t)T * 512);
    half_to_float(feat[6], featf, (long)T * 512, st);
    layernorm_rows(featf, 512, T, 512, h->ln0_g, h->ln0_b, 1e-5f, nullptr, 0, ln16, 512, st);
    float* x32 = ar.alloc<float>((size_t)T * 768);
    float* tmp32 = ar.alloc<float>((size_t)T * 768);
    float* xb32 = ar.alloc<float>((size_t)T * 768);
    __half* x16 = ar.alloc<__half>((size_t)T * 768);
    __half* ctx16 = ar.alloc<__half>((size_t)T * 768);
    {
        GemmArgs g;
        g.A = ln16; g.lda = 512; g.a_rows = T; g.a_cols = 512;
        g.B = h->proj.d; g.ldb = h->proj.cols; g.b_rows = h->proj.rows; g.b_cols = h->proj.cols;
        g.M = T; g.N = 768; seg_linear(g, 512);
        g.bias = h->proj_b;
        g.out32 = xb32; g.ld32 = 768; g.out16 = x16; g.ld16 = 768;
        gemm(g, st);
    }
    {
        GemmArgs g;
        g.A = x16; g.lda = 768; g.a_rows = T; g.a_cols = 768;
        g.B = h->posconv.d; g.ldb = h->posconv.cols; g.b_rows = h->posconv.rows; g.b_cols = h->posconv.cols;
        g.M = T; g.N = 48;
        g.nseg = 128;
        for (int j = 0; j < 128; ++j) g.seg[j] = {j - 64, 0, 0, 1};
        g.batch = 16; g.a_col_z = 48; g.b_row_z = 64; g.c_z = 48; g.bias_z = 48;
        g.bias = h->pos_b; g.act1 = ACT_GELU;
        g.res2 = xb32; g.ldres2 = 768;
        g.out32 = tmp32; g.ld32 = 768;
        gemm(g, st);
    }
    layernorm_rows(tmp32, 768, T, 768, h->eln_g, h->eln_b, 1e-5f, x32, 768, x16, 768, st);

    // ---- transformer layers ----
    __half* qk16 = ar.alloc<__half>((size_t)T * 1536);
    __half* vT16 = ar.alloc<__half>((size_t)768 * Tp);
    float* S32 = ar.alloc<float>((size_t)12 * T * Tp);
    __half* P16 = ar.alloc<__half>((size_t)12 * T * Tp);
    __half* h16 = ar.alloc<__half>((size_t)T * 3072);
    for (int li = 0; li < output_layer; ++li) {
        const rvcb_hubert::Layer& L = h->layers[li];
        const bool last = (li == output_layer - 1);
        {   // q (pre-scaled) | k
            GemmArgs g;
            g.A = x16; g.lda = 768; g.a_rows = T; g.a_cols = 768;
            g.B = L.wqk.d; g.ldb = L.wqk.cols; g.b_rows = L.wqk.rows; g.b_cols = L.wqk.cols;
            g.M = T; g.N = 1536; seg_linear(g, 768);
            g.bias = L.bqk; g.out16 = qk16; g.ld16 = 1536;
            gemm(g, st);
        }
        {   // V^T = Wv x^T + bv  (weights as the A operand -> channel-major V for the PV GEMM)
            GemmArgs g;
            g.A = L.wv.d; g.lda = L.wv.cols; g.a_rows = 768; g.a_cols = 768;
            g.B = x16; g.ldb = 768; g.b_rows = T; g.b_cols = 768;
            g.M = 768; g.N = T; seg_linear(g, 768);
            g.bias = L.bv; g.bias_per_row = 1; g.out16 = vT16; g.ld16 = Tp;
            gemm(g, st);
        }
        static const bool fused_attn = [] { const char* e = getenv("RVCB_ATTN"); return !(e && e[0] == '0'); }();   // RVCB_ATTN=0: the 3-launch path
        AttnFusedArgs at;
        at.q = qk16; at.ldq = 1536; at.k = qk16 + 768; at.ldk = 1536; at.vT = vT16; at.ldv = Tp; at.T = T; at.heads = 12; at.dk = 64; at.dv = 64;
        at.out = ctx16; at.ldo = 768;
        if (fused_attn && attention_fused_supported(at)) {
            // one launch: scores stay in TMEM, probabilities in shared memory (attn_fused.cu)
            attention_fused(at, st);
        } else {
            {   // S[h] = q_h k_h^T
                GemmArgs g;
                g.A = qk16; g.lda = 1536; g.a_rows = T; g.a_cols = 768;
                g.B = qk16; g.ldb = 1536; g.b_rows = T; g.b_cols = 1536;
                g.M = T; g.N = T; seg_linear(g, 64);
                g.batch = 12; g.a_col_z = 64; g.b_col0 = 768; g.b_col_z = 64; g.c_z = (long)T * Tp;
                g.out32 = S32; g.ld32 = Tp;
                gemm(g, st);
            }
            softmax_rows(S32, Tp, 12, T, P16, Tp, nullptr, 0, 0, nullptr, st);
            {   // ctx[:, h] = P_h V_h
                GemmArgs g;
                g.A = P16; g.lda = Tp; g.a_rows = 12 * T; g.a_cols = T;
                g.B = vT16; g.ldb = Tp; g.b_rows = 768; g.b_cols = T;
                g.M = T; g.N = 64; seg_linear(g, T);
                g.batch = 12; g.a_row_z = T; g.b_row_z = 64; g.c_z = 64;
                g.out16 = ctx16; g.ld16 = 768;
                gemm(g, st);
            }
        }
        {   // out proj + residual
            GemmArgs g;
            g.A = ctx16; g.lda = 768; g.a_rows = T; g.a_cols = 768;
            g.B = L.wo.d; g.ldb = L.wo.cols; g.b_rows = L.wo.rows; g.b_cols = L.wo.cols;
            g.M = T; g.N = 768; seg_linear(g, 768);
            g.bias = L.bo; g.res1 = x32; g.ldres1 = 768; g.out32 = tmp32; g.ld32 = 768;
            gemm(g, st);
        }
        layernorm_rows(tmp32, 768, T, 768, L.ln1_g, L.ln1_b, 1e-5f, x32, 768, x16, 768, st);
        {   // fc1 + GELU
            GemmArgs g;
            g.A = x16; g.lda = 768; g.a_rows = T; g.a_cols = 768;
            g.B = L.w1.d; g.ldb = L.w1.cols; g.b_rows = L.w1.rows; g.b_cols = L.w1.cols;
            g.M = T; g.N = 3072; seg_linear(g, 768);
            g.bias = L.b1; g.act1 = ACT_GELU; g.out16 = h16; g.ld16 = 3072;
            gemm(g, st);
        }
        {   // fc2 + residual
            GemmArgs g;
            g.A = h16; g.lda = 3072; g.a_rows = T; g.a_cols = 3072;
            g.B = L.w2.d; g.ldb = L.w2.cols; g.b_rows = L.w2.rows; g.b_cols = L.w2.cols;
            g.M = T; g.N = 768; seg_linear(g, 3072);
            g.bias = L.b2; g.res1 = x32; g.ldres1 = 768; g.out32 = tmp32; g.ld32 = 768;
            gemm(g, st);
        }
        layernorm_rows(tmp32, 768, T, 768, L.ln2_g, L.ln2_b, 1e-5f, last ? d_out : x32, 768, last ? nullptr : x16, 768, st);
    }
    if (n_frames) *n_frames = T;
}

extern "C" {

int rvcb_hubert_create(const rvcb_weights* w, rvcb_hubert** out) {
    RVCB_API_BEGIN
    RVCB_CHECK(w && out, "null argument");
    *out = hubert_build(*w);
    RVCB_API_END
}

int rvcb_hubert_num_frames(int n_samples) {
    int T[7];
    conv_lens(n_samples, T);
    return T[6];
}

int rvcb_hubert_extract_features(rvcb_hubert* h, const float* d_wav, int n_samples, int output_layer, float* d_out, int* n_frames,
                                 void* stream) {
    RVCB_API_BEGIN
    RVCB_CHECK(h && d_wav && d_out, "null argument");
    hubert_forward(h, d_wav, n_samples, output_layer, d_out, n_frames, (cudaStream_t)stream);
    RVCB_API_END
}

int rvcb_hubert_final_proj(rvcb_hubert* h, const float* d_in, int T, float* d_out, void* stream) {
    RVCB_API_BEGIN
    RVCB_CHECK(h && h->has_final, "hubert: no final_proj weights");
    cudaStream_t st = (cudaStream_t)stream;
    h->arena.reserve((size_t)T * 768 * 2 + (1 << 20));
    h->arena.reset();
    __half* x16 = h->arena.alloc<__half>((size_t)T * 768);
    cast_f32_f16(d_in, x16, (long)T * 768, st);
    GemmArgs g;
    g.A = x16; g.lda = 768; g.a_rows = T; g.a_cols = 768;
    g.B = h->wfinal.d; g.ldb = h->wfinal.cols; g.b_rows = h->wfinal.rows; g.b_cols = h->wfinal.cols;
    g.M = T; g.N = 256; seg_linear(g, 768);
    g.bias = h->bfinal; g.out32 = d_out; g.ld32 = 256;
    gemm(g, st);
    RVCB_API_END
}

void rvcb_hubert_destroy(rvcb_hubert* h) { delete h; }

}  // extern "C"
