// SynthesizerTrnMs{256,768}NSFsid.infer on the implicit-GEMM engine
// (rvc/layers/synthesizers.py:159-203): TextEncoder -> prior sample -> flow^-1 -> NSF-HiFi-GAN.
//
// HBM layout: activations channels-last [T, C]; every conv / linear is one gemm() call whose
// taps are TMA-shifted K segments; fp16 copies feed the tensor cores, residual streams and
// resblock accumulators stay fp32 (that is what holds the 1e-3 waveform bound).
#include "../../include/rvcb200.h"
#include "api_macros.h"
#include "attn_fused.cuh"
#include "gemm.cuh"
#include <algorithm>
#include "kernels.cuh"
#include "resblock_fused.cuh"
#include "weights.cuh"

#include <cmath>

using namespace rvcb;

namespace {

struct AttnLayer {
    PackedB wqk, wv, wo, ek, evT, w1, w2;
    float *bqk, *bv, *bo, *b1, *b2, *ln1_g, *ln1_b, *ln2_g, *ln2_b;
    float* ev32 = nullptr;         // emb_rel_v as fp32 [21, kc] (fused attention)
};
struct FlowLayer {
    PackedB pre, in[3], res[2], skip[3], post;
    float *pre_b, *res_b[2], *skip_b[3], *post_b;
};
struct ResBlock {
    PackedB c1[3], c2[3];
    float *b1[3], *b2[3];
    int k, dil[3];
    RBFusedWeights fused;      // the six convolutions packed for the one-launch kernel (resblock_fused.cu); C == 0: not available
};
struct Stage {
    PackedB up; float* up_b; int s, k, cin, cout;
    int noise_k, noise_stride, noise_pad;   // noise conv (nsf.py:176-183), folded into the ups GEMM as extra K columns
    int Mp;                                 // padded count of harmonic-source columns appended to the ups GEMM's A rows (0: no-f0)
    int dm, bk;                             // polyphase reach (taps -dm..dm) and K block of the upsampling GEMM
    ResBlock rb[4];
};

// RVCB_FUSED: comma-separated channel counts that use the fused residual-block kernel ("0" = none); default below
bool fused_mask(int C) {
    static int m32 = -1, m64 = 0, m128 = 0;
    if (m32 < 0) {
        const char* e = getenv("RVCB_FUSED");
        const std::string v = e ? e : "32,64";
        auto has = [&](const char* t) { return ("," + v + ",").find(std::string(",") + t + ",") != std::string::npos; };
        m32 = has("32"); m64 = has("64"); m128 = has("128");
    }
    return C == 32 ? m32 : (C == 64 ? m64 : (C == 128 ? m128 : 0));
}

GemmArgs mk(const __half* A, long lda, int a_rows, int a_cols, const PackedB& B, int M, int N, int bk = 64) {
    GemmArgs g;
    g.A = A; g.lda = lda; g.a_rows = a_rows; g.a_cols = a_cols;
    g.B = B.d; g.ldb = B.cols; g.b_rows = B.rows; g.b_cols = B.cols;
    g.M = M; g.N = N; g.block_k = bk;
    return g;
}

}  // namespace

struct rvcb_synth {
    rvcb_synth_config cfg{};
    DevOwner own;
    Arena arena;
    int upp = 1, kc = 0, HP = 0;
    bool use_f0 = true;      // false: SynthesizerTrnMs*NSFsid_nono (no pitch embedding, plain Generator decoder)
    // enc_p
    PackedB emb_phone; float* emb_phone_b = nullptr; float* emb_pitch = nullptr;
    std::vector<AttnLayer> attn;
    PackedB proj; float* proj_b = nullptr;
    // flow
    std::vector<FlowLayer> flows;
    float* cond_w = nullptr;    // [n_flows*3*2*hidden, gin] rows in interleaved (tanh_i, sigmoid_i) order
    float* cond_b = nullptr;    // same order: cond bias
    float* in_b = nullptr;      // same order: in_layer biases
    float* emb_g = nullptr;
    // dec
    PackedB conv_pre; float *conv_pre_b = nullptr, *dcond_w = nullptr, *dcond_b = nullptr;
    std::vector<Stage> stages;
    float* conv_post_w = nullptr; int conv_post_k = 0;   // [k, C] fp32
    float lin_w = 1.f, lin_b = 0.f;
};

static rvcb_synth* synth_build(const rvcb_synth_config& c, const rvcb_weights& w) {
    auto* h = new rvcb_synth();
    try {
        h->cfg = c;
        DevOwner& own = h->own;
        const int H = c.hidden_channels, F = c.filter_channels, I = c.inter_channels, heads = c.n_heads;
        RVCB_CHECK(H % heads == 0 && H % 8 == 0 && I % 2 == 0, "synth: bad channel config");
        RVCB_CHECK(c.n_upsamples >= 1 && c.n_upsamples <= 8 && c.n_resblock_kernels >= 1 && c.n_resblock_kernels <= 4, "synth: bad decoder config");
        const int kc = H / heads, HP = pad_to(kc, 64);
        h->kc = kc; h->HP = HP;
        h->upp = 1;
        for (int i = 0; i < c.n_upsamples; ++i) h->upp *= c.upsample_rates[i];
        // ---- enc_p ----
        h->emb_phone = pack_linear(own, w.get("enc_p.emb_phone.weight").data.data(), H, c.encoder_dim);
        h->emb_phone_b = own.upload(w.get("enc_p.emb_phone.bias").data);
        h->use_f0 = w.has("enc_p.emb_pitch.weight");
        if (h->use_f0) h->emb_pitch = own.upload(w.get("enc_p.emb_pitch.weight").data);
        for (int l = 0; l < c.n_layers; ++l) {
            const std::string p = "enc_p.encoder.attn_layers." + std::to_string(l) + ".";
            AttnLayer L{};
            // q|k, each head padded to HP columns (zero rows -> zero outputs)
            std::vector<float> wqk((size_t)2 * heads * HP * H, 0.f), bqk((size_t)2 * heads * HP, 0.f);
            const WT &wq = w.get(p + "conv_q.weight"), &wk = w.get(p + "conv_k.weight");
            const WT &bq = w.get(p + "conv_q.bias"), &bk = w.get(p + "conv_k.bias");
            for (int hd = 0; hd < heads; ++hd)
                for (int d = 0; d < kc; ++d) {
                    for (int k = 0; k < H; ++k) {
                        wqk[((size_t)hd * HP + d) * H + k] = wq.data[(size_t)(hd * kc + d) * H + k];
                        wqk[((size_t)(heads + hd) * HP + d) * H + k] = wk.data[(size_t)(hd * kc + d) * H + k];
                    }
                    bqk[(size_t)hd * HP + d] = bq.data[hd * kc + d];
                    bqk[(size_t)(heads + hd) * HP + d] = bk.data[hd * kc + d];
                }
            L.wqk = pack_linear(own, wqk.data(), 2 * heads * HP, H);
            L.bqk = own.upload(bqk);
            L.wv = pack_linear(own, w.get(p + "conv_v.weight").data.data(), H, H);
            L.bv = own.upload(w.get(p + "conv_v.bias").data);
            L.wo = pack_linear(own, w.get(p + "conv_o.weight").data.data(), H, H);
            L.bo = own.upload(w.get(p + "conv_o.bias").data);
            const WT &ek = w.get(p + "emb_rel_k"), &ev = w.get(p + "emb_rel_v");       // [1, 2w+1, kc]
            const int R = (int)ek.dim(1);
            RVCB_CHECK(R == 21, "synth: window_size must be 10");
            {
                std::vector<float> e((size_t)32 * HP, 0.f);
                for (int r = 0; r < R; ++r)
                    for (int d = 0; d < kc; ++d) e[(size_t)r * HP + d] = ek.data[(size_t)r * kc + d];
                L.ek = upload_half(own, e, 32, HP);
                std::vector<float> t((size_t)pad_to(kc, 128) * 64, 0.f);
                for (int r = 0; r < R; ++r)
                    for (int d = 0; d < kc; ++d) t[(size_t)d * 64 + r] = ev.data[(size_t)r * kc + d];
                L.evT = upload_half(own, t, pad_to(kc, 128), 64);
                L.ev32 = own.upload(ev.data);
            }
            L.ln1_g = own.upload(w.get("enc_p.encoder.norm_layers_1." + std::to_string(l) + ".gamma").data);
            L.ln1_b = own.upload(w.get("enc_p.encoder.norm_layers_1." + std::to_string(l) + ".beta").data);
            L.ln2_g = own.upload(w.get("enc_p.encoder.norm_layers_2." + std::to_string(l) + ".gamma").data);
            L.ln2_b = own.upload(w.get("enc_p.encoder.norm_layers_2." + std::to_string(l) + ".beta").data);
            const std::string f = "enc_p.encoder.ffn_layers." + std::to_string(l) + ".";
            L.w1 = pack_conv1d(own, w.get(f + "conv_1.weight").data.data(), F, H, c.kernel_size);
            L.b1 = own.upload(w.get(f + "conv_1.bias").data);
            L.w2 = pack_conv1d(own, w.get(f + "conv_2.weight").data.data(), H, F, c.kernel_size);
            L.b2 = own.upload(w.get(f + "conv_2.bias").data);
            h->attn.push_back(L);
        }
        h->proj = pack_linear(own, w.get("enc_p.proj.weight").data.data(), 2 * I, H);
        h->proj_b = own.upload(w.get("enc_p.proj.bias").data);
        // ---- flow ----
        const int half = I / 2, n_flows = 4;
        std::vector<float> condw((size_t)n_flows * 3 * 2 * H * c.gin_channels), condb((size_t)n_flows * 3 * 2 * H), inb((size_t)n_flows * 3 * 2 * H);
        for (int f = 0; f < n_flows; ++f) {
            const std::string p = "flow.flows." + std::to_string(2 * f) + ".";
            FlowLayer L{};
            L.pre = pack_linear(own, w.get(p + "pre.weight").data.data(), H, half);
            L.pre_b = own.upload(w.get(p + "pre.bias").data);
            const std::vector<float> cw = effective_weight(w, p + "enc.cond_layer");
            const WT& cb = w.get(p + "enc.cond_layer.bias");
            RVCB_CHECK((int)cb.numel() == 6 * H, "synth: WN must have 3 layers");
            for (int l = 0; l < 3; ++l) {
                const std::string li = std::to_string(l);
                const std::vector<float> iw = effective_weight(w, p + "enc.in_layers." + li);        // [2H, H, 5]
                const WT& ib = w.get(p + "enc.in_layers." + li + ".bias");
                const int k5 = (int)(iw.size() / ((size_t)2 * H * H));
                RVCB_CHECK(k5 == 5, "synth: WN kernel must be 5");
                // interleave rows: packed row 2i = tanh channel i, 2i+1 = sigmoid channel H+i
                std::vector<float> perm(iw.size());
                for (int i = 0; i < H; ++i) {
                    memcpy(&perm[(size_t)(2 * i) * H * k5], &iw[(size_t)i * H * k5], sizeof(float) * H * k5);
                    memcpy(&perm[(size_t)(2 * i + 1) * H * k5], &iw[(size_t)(H + i) * H * k5], sizeof(float) * H * k5);
                    const size_t o = ((size_t)f * 3 + l) * 2 * H;
                    for (int s2 = 0; s2 < 2; ++s2) {
                        const int src = s2 == 0 ? i : H + i;
                        inb[o + 2 * i + s2] = ib.data[src];
                        condb[o + 2 * i + s2] = cb.data[(size_t)l * 2 * H + src];
                        memcpy(&condw[(o + 2 * i + s2) * c.gin_channels], &cw[((size_t)l * 2 * H + src) * c.gin_channels],
                               sizeof(float) * c.gin_channels);
                    }
                }
                L.in[l] = pack_conv1d(own, perm.data(), 2 * H, H, k5);
                const std::vector<float> rw = effective_weight(w, p + "enc.res_skip_layers." + li);  // [2H or H, H, 1]
                const WT& rb = w.get(p + "enc.res_skip_layers." + li + ".bias");
                if (l < 2) {
                    L.res[l] = pack_linear(own, rw.data(), H, H);
                    L.res_b[l] = own.upload(std::vector<float>(rb.data.begin(), rb.data.begin() + H));
                    L.skip[l] = pack_linear(own, rw.data() + (size_t)H * H, H, H);
                    L.skip_b[l] = own.upload(std::vector<float>(rb.data.begin() + H, rb.data.end()));
                } else {
                    L.skip[l] = pack_linear(own, rw.data(), H, H);
                    L.skip_b[l] = own.upload(rb.data);
                }
            }
            L.post = pack_linear(own, w.get(p + "post.weight").data.data(), half, H);
            L.post_b = own.upload(w.get(p + "post.bias").data);
            h->flows.push_back(L);
        }
        h->cond_w = own.upload(condw);
        h->cond_b = own.upload(condb);
        h->in_b = own.upload(inb);
        h->emb_g = own.upload(w.get("emb_g.weight").data);
        // ---- dec ----
        if (h->use_f0) {
            h->lin_w = w.get("dec.m_source.l_linear.weight").data[0];
            h->lin_b = w.get("dec.m_source.l_linear.bias").data[0];
        }
        const int C0 = c.upsample_initial_channel;
        h->conv_pre = pack_conv1d(own, w.get("dec.conv_pre.weight").data.data(), C0, I, 7);
        h->conv_pre_b = own.upload(w.get("dec.conv_pre.bias").data);
        h->dcond_w = own.upload(w.get("dec.cond.weight").data);
        h->dcond_b = own.upload(w.get("dec.cond.bias").data);
        int ch = C0;
        for (int i = 0; i < c.n_upsamples; ++i) {
            Stage S{};
            S.s = c.upsample_rates[i]; S.k = c.upsample_kernel_sizes[i]; S.cin = ch; S.cout = ch / 2;
            RVCB_CHECK(S.k >= S.s && (S.k - S.s) % 2 == 0, "synth: upsample kernel must satisfy k >= s with k - s even");
            S.dm = convT1d_reach(S.k, S.s, (S.k - S.s) / 2);       // input-frame reach of the polyphase form: taps d = -dm..dm
            S.bk = S.cin >= 64 ? 64 : 32;                           // K block of the upsampling GEMM (5-stage decoders end at C_in = 32)
            const std::string us = "dec.ups." + std::to_string(i);
            const std::vector<float> uw = effective_weight(w, us);
            std::vector<float> ub((size_t)S.s * S.cout);
            const WT& ubias = w.get(us + ".bias");
            RVCB_CHECK(S.cin % S.bk == 0, "synth: upsample input channels must be multiples of 32");
            if (h->use_f0) {
                const WT& nw = w.get("dec.noise_convs." + std::to_string(i) + ".weight");
                const WT& nb = w.get("dec.noise_convs." + std::to_string(i) + ".bias");
                S.noise_k = (int)nw.dim(2);
                if (i + 1 < c.n_upsamples) {
                    int st = 1;
                    for (int j = i + 1; j < c.n_upsamples; ++j) st *= c.upsample_rates[j];
                    S.noise_stride = st; S.noise_pad = st / 2;
                    RVCB_CHECK(S.noise_k == 2 * st, "synth: noise conv kernel mismatch");
                } else {
                    S.noise_stride = 1; S.noise_pad = 0;
                }
                S.Mp = round_up((S.s - 1) * S.noise_stride + S.noise_k, S.bk);
                S.up = pack_convT1d(own, uw.data(), S.cin, S.cout, S.k, S.s, (S.k - S.s) / 2, S.bk, nw.data.data(), S.noise_k,
                                    S.noise_stride, S.Mp);
                for (int r = 0; r < S.s; ++r)
                    for (int co = 0; co < S.cout; ++co) ub[(size_t)r * S.cout + co] = ubias.data[co] + nb.data[co];
            } else {
                S.Mp = 0;
                S.up = pack_convT1d(own, uw.data(), S.cin, S.cout, S.k, S.s, (S.k - S.s) / 2, S.bk);
                for (int r = 0; r < S.s; ++r)
                    for (int co = 0; co < S.cout; ++co) ub[(size_t)r * S.cout + co] = ubias.data[co];
            }
            S.up_b = own.upload(ub);
            ch = S.cout;
            const int bk = ch >= 64 ? 64 : (ch >= 32 ? 32 : 16);
            RVCB_CHECK(ch % 16 == 0, "synth: decoder channels must be multiples of 16");
            for (int j = 0; j < c.n_resblock_kernels; ++j) {
                ResBlock& R = S.rb[j];
                R.k = c.resblock_kernel_sizes[j];
                const std::string rp = "dec.resblocks." + std::to_string(i * c.n_resblock_kernels + j) + ".";
                std::vector<float> ew1[3], ew2[3];
                const float *pw1[3], *pw2[3], *pb1[3], *pb2[3];
                for (int q = 0; q < 3; ++q) {
                    R.dil[q] = c.resblock_dilations[j][q];
                    ew1[q] = effective_weight(w, rp + "convs1." + std::to_string(q));
                    ew2[q] = effective_weight(w, rp + "convs2." + std::to_string(q));
                    R.c1[q] = pack_conv1d(own, ew1[q].data(), ch, ch, R.k, bk);
                    R.c2[q] = pack_conv1d(own, ew2[q].data(), ch, ch, R.k, bk);
                    const WT& wb1 = w.get(rp + "convs1." + std::to_string(q) + ".bias");
                    const WT& wb2 = w.get(rp + "convs2." + std::to_string(q) + ".bias");
                    R.b1[q] = own.upload(wb1.data);
                    R.b2[q] = own.upload(wb2.data);
                    pw1[q] = ew1[q].data(); pw2[q] = ew2[q].data(); pb1[q] = wb1.data.data(); pb2[q] = wb2.data.data();
                }
                if (resblock_fused_supported(ch, R.k, R.dil)) R.fused = pack_resblock_fused(own, ch, R.k, R.dil, pw1, pb1, pw2, pb2);
            }
            h->stages.push_back(S);
        }
        {   // [1, C, k] -> [k, C]
            const WT& pw = w.get("dec.conv_post.weight");
            const int kp = (int)pw.dim(2);
            std::vector<float> t((size_t)kp * ch);
            for (int ci = 0; ci < ch; ++ci)
                for (int j = 0; j < kp; ++j) t[(size_t)j * ch + ci] = pw.data[(size_t)ci * kp + j];
            h->conv_post_w = own.upload(t);
            h->conv_post_k = kp;
        }
    } catch (...) {
        delete h;
        throw;
    }
    return h;
}

static void synth_forward(rvcb_synth* h, const float* d_phone, int T, int sid, const long long* d_pitch, const float* d_pitchf,
                          const float* d_noise_prior, const float* d_noise_src, int skip_head, int return_length, int return_length2,
                          float* d_wav_out, int* n_out, cudaStream_t st, int keep_head = -1, int keep_length = -1) {
    const rvcb_synth_config& c = h->cfg;
    const int H = c.hidden_channels, F = c.filter_channels, I = c.inter_channels, heads = c.n_heads, kc = h->kc, HP = h->HP;
    const int upp = h->upp;
    RVCB_CHECK(T >= 1 && sid >= 0 && sid < c.spk_embed_dim, "synth: bad T or sid");
    const bool rt = skip_head >= 0 && return_length >= 0;
    // "keep" mode (offline utterances): the caller discards everything outside frames [keep_head, keep_head + keep_length) -- the
    // x_pad context of pipeline.py:241,295.  TextEncoder attention is global and runs over all T frames; the flow (receptive field
    // +-24 frames, the reference's own constant at synthesizers.py:172) and the decoder (+-10 frames: conv_pre 3 + the resblock /
    // transposed-conv halos of the four stages) are local, so they run over the kept frames plus margins only.  Every kept sample
    // sees exactly the operands it sees in the full computation (the NSF sine phase is still accumulated from frame 0, the noise
    // tensors are the full-length ones): the kept output is bit-identical to infer(...)[keep_head*upp : (keep_head+keep_length)*upp].
    const bool keep = !rt && keep_head >= 0 && keep_length >= 1;
    constexpr int kDecMargin = 16, kFlowMargin = 24;
    RVCB_CHECK(!keep || (keep_head + keep_length <= T && return_length2 < 0), "synth: bad keep_head/keep_length");
    const int ds = keep ? std::max(keep_head - kDecMargin, 0) : 0;                         // first decoder frame
    const int de = keep ? std::min(keep_head + keep_length + kDecMargin, T) : T;            // one past the last decoder frame
    const int flow_head = rt ? std::max(skip_head - 24, 0) : (keep ? std::max(ds - kFlowMargin, 0) : 0);
    const int flow_end = keep ? std::min(de + kFlowMargin, T) : T;
    const int dec_head = rt ? skip_head - flow_head : ds - flow_head;
    const int Tf = flow_end - flow_head;                 // frames through the flow
    const int Td = rt ? return_length : de - ds;         // frames into the decoder
    const int Tn = (return_length2 >= 0) ? return_length2 : Td;   // decoder frames after formant resize
    RVCB_CHECK(Tf >= 1 && Td >= 1 && dec_head + Td <= Tf && Tn >= 1, "synth: bad skip_head/return_length");
    const int Tp = round_up(T, 8);
    // ---- arena ----
    auto rnd = [](size_t b) { return (b + 1023) & ~size_t(1023); };
    size_t need = 8u << 20;
    need += rnd((size_t)T * c.encoder_dim * 2) + 4 * rnd((size_t)T * H * 4) + 3 * rnd((size_t)T * H * 2);
    need += rnd((size_t)T * 2 * heads * HP * 2) + rnd((size_t)H * Tp * 2) + rnd((size_t)heads * T * Tp * 4) + rnd((size_t)heads * T * Tp * 2);
    need += rnd((size_t)heads * T * 32 * 4) + rnd((size_t)heads * T * 64 * 2) + rnd((size_t)T * F * 2) + rnd((size_t)T * 2 * I * 4);
    need += 6 * rnd((size_t)Tf * H * 4) + 4 * rnd((size_t)Tf * H * 2);
    need += (keep ? 2 * rnd((size_t)T * upp * 4) + rnd((size_t)T * 4 + 64) : 0) + 2 * rnd((size_t)Tn * upp * 4) + rnd((size_t)Tn * (c.upsample_initial_channel + h->stages[0].Mp) * 2) + 2 * rnd((size_t)Tn * I * 4);
    size_t carry_e = 0;      // halves per carry buffer: widest [T_in, C_in + Mp] input of stages 1.. and the last stage's output
    {
        size_t mx = 0;
        int Tt = Tn, ch = c.upsample_initial_channel;
        for (int i = 0; i < c.n_upsamples; ++i) {
            if (i > 0) carry_e = std::max(carry_e, (size_t)Tt * (ch + h->stages[i].Mp));
            Tt *= c.upsample_rates[i];
            ch /= 2;
            const size_t e = (size_t)Tt * ch;
            mx = std::max(mx, 3 * rnd(e * 4) + 4 * rnd(e * 2));
        }
        carry_e = std::max(carry_e, (size_t)Tt * ch);
        need += mx + 2 * rnd(carry_e * 2) + (16u << 20);     // + whole-tile slack of the fused residual-block outputs
    }
    h->arena.reserve(need);
    h->arena.reset();
    Arena& ar = h->arena;

    // ================= TextEncoder (encoders.py:135-159) =================
    __half* phone16 = ar.alloc<__half>((size_t)T * c.encoder_dim);
    cast_f32_f16(d_phone, phone16, (long)T * c.encoder_dim, st);
    float* lin32 = ar.alloc<float>((size_t)T * H);
    float* x32 = ar.alloc<float>((size_t)T * H);
    float* tmp32 = ar.alloc<float>((size_t)T * H);
    float* ctx32 = ar.alloc<float>((size_t)T * H);
    __half* x16 = ar.alloc<__half>((size_t)T * H);
    __half* ctx16 = ar.alloc<__half>((size_t)T * H);
    {
        GemmArgs g = mk(phone16, c.encoder_dim, T, c.encoder_dim, h->emb_phone, T, H);
        seg_linear(g, c.encoder_dim);
        g.bias = h->emb_phone_b; g.out32 = lin32; g.ld32 = H;
        gemm(g, st);
    }
    textenc_embed(lin32, h->use_f0 ? d_pitch : nullptr, h->emb_pitch, T, H, sqrtf((float)H), x32, x16, st);
    const int NQK = 2 * heads * HP;
    __half* qk16 = ar.alloc<__half>((size_t)T * NQK);
    __half* vT16 = ar.alloc<__half>((size_t)H * Tp);
    float* S32 = ar.alloc<float>((size_t)heads * T * Tp);
    __half* P16 = ar.alloc<__half>((size_t)heads * T * Tp);
    float* qrel32 = ar.alloc<float>((size_t)heads * T * 32);
    __half* prel16 = ar.alloc<__half>((size_t)heads * T * 64);
    __half* ffn16 = ar.alloc<__half>((size_t)T * F);
    const float qscale = 1.f / sqrtf((float)kc);
    const int ks = c.kernel_size, padl = (ks - 1) / 2;
    for (int l = 0; l < c.n_layers; ++l) {
        const AttnLayer& L = h->attn[l];
        {
            GemmArgs g = mk(x16, H, T, H, L.wqk, T, NQK);
            seg_linear(g, H);
            g.bias = L.bqk; g.out16 = qk16; g.ld16 = NQK;
            gemm(g, st);
        }
        {   // V^T (channel-major) = Wv x^T + bv
            GemmArgs g;
            g.A = L.wv.d; g.lda = L.wv.cols; g.a_rows = H; g.a_cols = H;
            g.B = x16; g.ldb = H; g.b_rows = T; g.b_cols = H;
            g.M = H; g.N = T; seg_linear(g, H);
            g.bias = L.bv; g.bias_per_row = 1; g.out16 = vT16; g.ld16 = Tp;
            gemm(g, st);
        }
        {   // relative-key logits (q / sqrt(kc)) E_k^T  -> [heads, T, 32]
            GemmArgs g = mk(qk16, NQK, T, heads * HP, L.ek, T, 32);
            seg_linear(g, HP);
            g.batch = heads; g.a_col_z = HP; g.c_z = (long)T * 32;
            g.alpha = qscale; g.out32 = qrel32; g.ld32 = 32;
            gemm(g, st);
        }
        // Measured (profiles/README.md, r2g): 181 us per layer against ~85 us for the five small launches below -- 2 heads x 13 query
        // blocks are only 26 CTAs, each walking 13 key blocks twice with one softmax warp per SM sub-partition.  Opt-in.
        static const bool fused_attn = [] { const char* e = getenv("RVCB_ATTN_REL"); return e && e[0] == '1'; }();
        AttnFusedArgs at;
        at.q = qk16; at.ldq = NQK; at.k = qk16 + (size_t)heads * HP; at.ldk = NQK; at.vT = vT16; at.ldv = Tp; at.T = T; at.heads = heads;
        at.dk = HP; at.dv = kc; at.qscale = qscale; at.qrel = qrel32; at.ev = L.ev32; at.out = ctx16; at.ldo = H;
        if (fused_attn && attention_fused_supported(at)) {
            // scores in TMEM, probabilities in shared memory, band logits / band values applied in place (attn_fused.cu)
            attention_fused(at, st);
        } else {
            {   // scores = (q k^T) / sqrt(kc)
                GemmArgs g;
                g.A = qk16; g.lda = NQK; g.a_rows = T; g.a_cols = heads * HP;
                g.B = qk16; g.ldb = NQK; g.b_rows = T; g.b_cols = NQK;
                g.M = T; g.N = T; seg_linear(g, HP);
                g.batch = heads; g.a_col_z = HP; g.b_col0 = heads * HP; g.b_col_z = HP; g.c_z = (long)T * Tp;
                g.alpha = qscale; g.out32 = S32; g.ld32 = Tp;
                gemm(g, st);
            }
            softmax_rows(S32, Tp, heads, T, P16, Tp, qrel32, 32, 10, prel16, st);
            {   // P V
                GemmArgs g;
                g.A = P16; g.lda = Tp; g.a_rows = heads * T; g.a_cols = T;
                g.B = vT16; g.ldb = Tp; g.b_rows = H; g.b_cols = T;
                g.M = T; g.N = kc; seg_linear(g, T);
                g.batch = heads; g.a_row_z = T; g.b_row_z = kc; g.c_z = kc;
                g.out32 = ctx32; g.ld32 = H;
                gemm(g, st);
            }
            {   // + relative values: sum_r P[i, i+r-10] E_v[r]
                GemmArgs g = mk(prel16, 64, heads * T, 64, L.evT, T, kc);
                seg_linear(g, 64);
                g.batch = heads; g.a_row_z = T; g.c_z = kc;
                g.res2 = ctx32; g.ldres2 = H; g.out16 = ctx16; g.ld16 = H;
                gemm(g, st);
            }
        }
        {
            GemmArgs g = mk(ctx16, H, T, H, L.wo, T, H);
            seg_linear(g, H);
            g.bias = L.bo; g.res1 = x32; g.ldres1 = H; g.out32 = tmp32; g.ld32 = H;
            gemm(g, st);
        }
        layernorm_rows(tmp32, H, T, H, L.ln1_g, L.ln1_b, 1e-5f, x32, H, x16, H, st);
        {
            GemmArgs g = mk(x16, H, T, H, L.w1, T, F);
            seg_conv1d(g, H, ks, 1, padl);
            g.bias = L.b1; g.act1 = ACT_RELU; g.out16 = ffn16; g.ld16 = F;
            gemm(g, st);
        }
        {
            GemmArgs g = mk(ffn16, F, T, F, L.w2, T, H);
            seg_conv1d(g, F, ks, 1, padl);
            g.bias = L.b2; g.res1 = x32; g.ldres1 = H; g.out32 = tmp32; g.ld32 = H;
            gemm(g, st);
        }
        layernorm_rows(tmp32, H, T, H, L.ln2_g, L.ln2_b, 1e-5f, x32, H, x16, H, st);
    }
    float* stats32 = ar.alloc<float>((size_t)Tf * 2 * I);
    {
        GemmArgs g = mk(x16 + (size_t)flow_head * H, H, Tf, H, h->proj, Tf, 2 * I);
        seg_linear(g, H);
        g.bias = h->proj_b; g.out32 = stats32; g.ld32 = 2 * I;
        gemm(g, st);
    }
    // ================= prior sample + flow^-1 (synthesizers.py:187-189, residuals.py:210-235) =================
    float* zA = ar.alloc<float>((size_t)Tf * I);
    float* zB = ar.alloc<float>((size_t)Tf * I);
    // keep mode: the prior noise is the full-length [inter, T] tensor, read from frame flow_head on
    prior_sample(stats32, keep ? d_noise_prior + flow_head : d_noise_prior, keep ? T : Tf, Tf, I, zA, st);
    const int half = I / 2;
    const float* gvec = h->emb_g + (size_t)sid * c.gin_channels;
    float* cbias = ar.alloc<float>((size_t)4 * 3 * 2 * H);
    matvec(h->cond_w, gvec, h->cond_b, h->in_b, cbias, 4 * 3 * 2 * H, c.gin_channels, st);
    __half* x0_16 = ar.alloc<__half>((size_t)Tf * half);
    float* h32 = ar.alloc<float>((size_t)Tf * H);
    __half* h16 = ar.alloc<__half>((size_t)Tf * H);
    __half* acts16 = ar.alloc<__half>((size_t)Tf * H);
    float* oacc32 = ar.alloc<float>((size_t)Tf * H);
    __half* oacc16 = ar.alloc<__half>((size_t)Tf * H);
    float* zin = zA;
    float* zout = zB;
    for (int f = 3; f >= 0; --f) {
        const FlowLayer& L = h->flows[f];
        flip_channels(zin, zout, x0_16, Tf, I, half, st);
        {
            GemmArgs g = mk(x0_16, half, Tf, half, L.pre, Tf, H);
            seg_linear(g, half);
            g.bias = L.pre_b; g.out32 = h32; g.ld32 = H; g.out16 = h16; g.ld16 = H;
            gemm(g, st);
        }
        for (int l = 0; l < 3; ++l) {
            {
                GemmArgs g = mk(h16, H, Tf, H, L.in[l], Tf, 2 * H);
                seg_conv1d(g, H, 5, 1, 2);
                g.bias = cbias + ((size_t)f * 3 + l) * 2 * H; g.gate = 1; g.out16 = acts16; g.ld16 = H;
                gemm(g, st);
            }
            if (l < 2) {
                GemmArgs g = mk(acts16, H, Tf, H, L.res[l], Tf, H);
                seg_linear(g, H);
                g.bias = L.res_b[l]; g.res1 = h32; g.ldres1 = H; g.out32 = h32; g.ld32 = H; g.out16 = h16; g.ld16 = H;
                gemm(g, st);
            }
            {
                GemmArgs g = mk(acts16, H, Tf, H, L.skip[l], Tf, H);
                seg_linear(g, H);
                g.bias = L.skip_b[l];
                if (l > 0) { g.res1 = oacc32; g.ldres1 = H; }
                g.out32 = oacc32; g.ld32 = H;
                if (l == 2) { g.out16 = oacc16; g.ld16 = H; }
                gemm(g, st);
            }
        }
        {   // x1 = x1 - m   (mean_only coupling, reverse)
            GemmArgs g = mk(oacc16, H, Tf, H, L.post, Tf, half);
            seg_linear(g, H);
            g.bias = L.post_b; g.alpha = -1.f; g.res2 = zout + half; g.ldres2 = I; g.out32 = zout + half; g.ld32 = I;
            gemm(g, st);
        }
        std::swap(zin, zout);
    }
    const float* z = zin + (size_t)dec_head * I;          // [Td, I]
    // ================= NSF-HiFi-GAN (nsf.py:145-191) =================
    float* har = nullptr;
    if (h->use_f0 && keep) {
        // the sine phase is a running sum over ALL frames (generators.py:166-177): build the source for the whole utterance
        // (one elementwise kernel over T * upp samples) and hand the decoder its slice
        float* phase = ar.alloc<float>((size_t)T + 8);
        float* har_full = ar.alloc<float>((size_t)T * upp);
        sine_source(d_pitchf, T, upp, c.sr, d_noise_src, h->lin_w, h->lin_b, phase, har_full, st);
        har = har_full + (size_t)ds * upp;
    } else if (h->use_f0) {
        const float* pf = d_pitchf + (rt ? skip_head : 0);
        float* phase = ar.alloc<float>((size_t)Td + 8);
        har = ar.alloc<float>((size_t)Td * upp);
        sine_source(pf, Td, upp, c.sr, d_noise_src, h->lin_w, h->lin_b, phase, har, st);
    }
    const float* zdec = z;
    if (Tn != Td) {
        if (h->use_f0) {
            float* har2 = ar.alloc<float>((size_t)Tn * upp);
            interp_linear_rows(har, Td * upp, har2, Tn * upp, 1, st);
            har = har2;
        }
        float* z2 = ar.alloc<float>((size_t)Tn * I);
        interp_linear_rows(z, Td, z2, Tn, I, st);
        zdec = z2;
    }
    const long n_har = (long)Tn * upp;
    __half* z16 = ar.alloc<__half>((size_t)Tn * I);
    cast_f32_f16(zdec, z16, (long)Tn * I, st);
    const int C0 = c.upsample_initial_channel;
    float* pre_b = ar.alloc<float>(C0);
    matvec(h->dcond_w, gvec, h->dcond_b, h->conv_pre_b, pre_b, C0, c.gin_channels, st);
    const long ld_in0 = C0 + h->stages[0].Mp;       // stage inputs are [T_in, C_in | Mp harmonic-source columns]
    __half* xin16 = ar.alloc<__half>((size_t)Tn * ld_in0);
    {
        GemmArgs g = mk(z16, I, Tn, I, h->conv_pre, Tn, C0);
        seg_conv1d(g, I, 7, 1, 3);
        g.bias = pre_b; g.act2 = ACT_LRELU; g.act2_p = 0.1f; g.out16 = xin16; g.ld16 = ld_in0;
        gemm(g, st);
    }
    int Tt = Tn;
    __half* carry[2] = {ar.alloc<__half>(carry_e), ar.alloc<__half>(carry_e)};   // lrelu(x) handed from stage to stage
    const size_t stage_mark = ar.off;
    __half* carry16 = nullptr;
    for (int i = 0; i < c.n_upsamples; ++i) {
        const Stage& S = h->stages[i];
        const int Tin = Tt, Tout = Tt * S.s, C = S.cout;
        const int bk = C >= 64 ? 64 : (C >= 32 ? 32 : 16);
        const size_t e = (size_t)Tout * C;
        ar.off = stage_mark;                           // stage scratch is recycled
        __half* prev16 = (i == 0) ? xin16 : carry[(i - 1) & 1];
        __half* next16 = carry[i & 1];
        const int nk = c.n_resblock_kernels;
        // One launch per residual block (resblock_fused.cu) where the kernel has a variant and the sequence is long enough to fill
        // the machine with whole tiles; short sequences (the realtime block) keep the layer-by-layer kernels, whose M tiles are finer.
        bool fused = fused_mask(C) && (size_t)Tout >= 64u * (size_t)resblock_fused_tile_rows(C);
        for (int j = 0; j < nk && fused; ++j) fused = S.rb[j].fused.C == C;
        float* xs32 = ar.alloc<float>(e);
        float *y32 = nullptr, *sum32 = nullptr, *yb[4] = {nullptr, nullptr, nullptr, nullptr};
        __half *xs16 = nullptr, *y16 = nullptr, *t16 = nullptr;
        if (fused) {
            for (int j = 0; j < nk; ++j)
                yb[j] = ar.alloc<float>((size_t)resblock_fused_out_rows(C, S.rb[j].k, S.rb[j].dil, Tout) * C);
        } else {
            y32 = ar.alloc<float>(e);
            sum32 = ar.alloc<float>(e);
            xs16 = ar.alloc<__half>(e);
            y16 = ar.alloc<__half>(e);
            t16 = ar.alloc<__half>(e);
        }
        const long ld_in = S.cin + S.Mp;
        const bool last_stage = (i + 1 == c.n_upsamples);
        const long ld_next = last_stage ? C : (C + h->stages[i + 1].Mp);
        if (S.Mp)   // column m of row t_in: har[t_in * s * stride_n + m - pad] (zero outside the source / past the taps)
            har_columns(har, n_har, prev16 + S.cin, ld_in, Tin, S.Mp, (S.s - 1) * S.noise_stride + S.noise_k, S.s * S.noise_stride,
                        S.noise_pad, st);
        {   // ConvTranspose1d as a polyphase GEMM: [Tin, s*C] == [Tout, C]; + the noise conv as one more K segment.
            // Epilogue emits x (fp32 residual stream) and lrelu(x) (fp16 operand of the first resblock convs).
            GemmArgs g = mk(prev16, ld_in, Tin, (int)ld_in, S.up, Tin, S.s * C, S.bk);
            g.nseg = 2 * S.dm + 1;
            for (int d = 0; d < g.nseg; ++d) g.seg[d] = {d - S.dm, 0, 0, S.cin / S.bk};
            if (S.Mp) g.seg[g.nseg++] = {0, S.cin, 0, S.Mp / S.bk};
            g.bias = S.up_b; g.out32 = xs32; g.ld32 = (long)S.s * C;
            if (!fused) { g.out16 = xs16; g.ld16 = (long)S.s * C; g.act2 = ACT_LRELU; g.act2_p = 0.1f; }
            gemm(g, st);
        }
        if (fused) {
            for (int j = 0; j < nk; ++j) resblock_fused(S.rb[j].fused, xs32, yb[j], Tout, st);
            resblock_mean_lrelu(yb, nk, Tout, C, next16, ld_next, last_stage ? 0.01f : 0.1f, st);
            carry16 = next16;
            Tt = Tout;
            continue;
        }
        // Branch order: widest kernel first.  The mean over branches is order-independent, and this way the branch whose
        // resident weights are largest needs neither the running-sum tile nor the fp16 hand-off tile in shared memory.
        int order[4] = {0, 1, 2, 3};
        for (int a = 1; a < nk && a < 4; ++a)
            for (int b = a; b > 0 && S.rb[order[b]].k > S.rb[order[b - 1]].k; --b) std::swap(order[b], order[b - 1]);
        for (int jo = 0; jo < nk; ++jo) {
            const int j = jo;                       // position in the running sum
            const ResBlock& R = S.rb[order[jo]];
            for (int q = 0; q < 3; ++q) {
                {
                    GemmArgs g = mk(q == 0 ? xs16 : y16, C, Tout, C, R.c1[q], Tout, C, bk);
                    seg_conv1d(g, C, R.k, R.dil[q], (R.k - 1) * R.dil[q] / 2);
                    g.bias = R.b1[q]; g.act2 = ACT_LRELU; g.act2_p = 0.1f; g.out16 = t16; g.ld16 = C;
                    gemm(g, st);
                }
                {
                    GemmArgs g = mk(t16, C, Tout, C, R.c2[q], Tout, C, bk);
                    seg_conv1d(g, C, R.k, 1, (R.k - 1) / 2);
                    g.bias = R.b2[q];
                    g.res1 = (q == 0) ? xs32 : y32; g.ldres1 = C;
                    if (q < 2) {
                        g.out32 = y32; g.ld32 = C; g.out16 = y16; g.ld16 = C; g.act2 = ACT_LRELU; g.act2_p = 0.1f;
                    } else {
                        g.alpha = 1.f / (float)nk;
                        if (j > 0) { g.res2 = sum32; g.ldres2 = C; }
                        g.out32 = sum32; g.ld32 = C;
                        if (j == nk - 1) {
                            g.out16 = next16; g.ld16 = ld_next; g.act2 = ACT_LRELU; g.act2_p = last_stage ? 0.01f : 0.1f;
                        }
                    }
                    gemm(g, st);
                }
            }
        }
        carry16 = next16;
        Tt = Tout;
    }
    if (keep) {
        float* wav_tmp = ar.alloc<float>((size_t)Tt);
        conv_post_tanh(carry16, Tt, h->stages.back().cout, h->conv_post_w, h->conv_post_k, wav_tmp, st);
        CUDA_CHECK(cudaMemcpyAsync(d_wav_out, wav_tmp + (size_t)(keep_head - ds) * upp, sizeof(float) * (size_t)keep_length * upp,
                                   cudaMemcpyDeviceToDevice, st));
        if (n_out) *n_out = keep_length * upp;
        return;
    }
    conv_post_tanh(carry16, Tt, h->stages.back().cout, h->conv_post_w, h->conv_post_k, d_wav_out, st);
    if (n_out) *n_out = Tt;
}

extern "C" {

int rvcb_synth_create(const rvcb_synth_config* cfg, const rvcb_weights* w, rvcb_synth** out) {
    RVCB_API_BEGIN
    RVCB_CHECK(cfg && w && out, "null argument");
    *out = synth_build(*cfg, *w);
    RVCB_API_END
}

int rvcb_synth_infer(rvcb_synth* h, const float* d_phone, int T, int sid, const int64_t* d_pitch, const float* d_pitchf,
                     const float* d_noise_prior, const float* d_noise_src, int skip_head, int return_length, int return_length2,
                     float* d_wav_out, int* n_out, void* stream) {
    RVCB_API_BEGIN
    RVCB_CHECK(h && d_phone && d_noise_prior && d_wav_out, "null argument");
    RVCB_CHECK(!h->use_f0 || (d_pitch && d_pitchf && d_noise_src), "f0 model: pitch, pitchf and the source noise are required");
    synth_forward(h, d_phone, T, sid, (const long long*)d_pitch, d_pitchf, d_noise_prior, d_noise_src, skip_head, return_length,
                  return_length2, d_wav_out, n_out, (cudaStream_t)stream);
    RVCB_API_END
}

int rvcb_synth_infer_keep(rvcb_synth* h, const float* d_phone, int T, int sid, const int64_t* d_pitch, const float* d_pitchf,
                          const float* d_noise_prior, const float* d_noise_src, int keep_head, int keep_length, float* d_wav_out, int* n_out,
                          void* stream) {
    RVCB_API_BEGIN
    RVCB_CHECK(h && d_phone && d_noise_prior && d_wav_out, "null argument");
    RVCB_CHECK(!h->use_f0 || (d_pitch && d_pitchf && d_noise_src), "f0 model: pitch, pitchf and the source noise are required");
    RVCB_CHECK(keep_head >= 0 && keep_length >= 1 && keep_head + keep_length <= T, "bad keep_head / keep_length");
    synth_forward(h, d_phone, T, sid, (const long long*)d_pitch, d_pitchf, d_noise_prior, d_noise_src, -1, -1, -1, d_wav_out, n_out,
                  (cudaStream_t)stream, keep_head, keep_length);
    RVCB_API_END
}

void rvcb_synth_destroy(rvcb_synth* h) { delete h; }

}  // extern "C"
