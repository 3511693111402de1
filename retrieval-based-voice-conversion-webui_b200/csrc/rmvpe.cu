// RMVPE f0 estimator: log-mel -> DeepUnet -> BiGRU -> 360-bin salience -> Hz
// Replaces RMVPE.compute_f0's device part + numpy decode (rvc/f0/rmvpe.py:96-164, mel.py:51-71,
// e2e.py:29-47, deepunet.py).
//
// HBM layout: feature maps channels-last [H = time, W = mel, C]; every 3x3 / 1x1 / transposed conv
// with C_in >= 16 is one implicit-GEMM launch (3-D TMA map, taps = shifted boxes, zero padding =
// OOB fill, BatchNorm folded into weights + bias, ReLU / residual fused in the epilogue, decoder
// concat = channel offset of the output pointer).  The C_in = 1 stem, pooling, the log-mel front
// end (fp32 DFT-as-GEMM), the recurrent BiGRU (thread-block clusters + DSMEM) and the decode
// are SIMT.
#include "../../include/rvcb200.h"
#include "api_macros.h"
#include "conv2d_row.cuh"
#include "gemm.cuh"
#include "kernels.cuh"
#include "weights.cuh"

#include <cooperative_groups.h>

namespace cg = cooperative_groups;
using namespace rvcb;

namespace {

constexpr int N_FFT = 1024, HOP = 160, N_MELS = 128, N_BINS = 513, N_CLASS = 360, DFT_NPAD = 1088;

struct ConvBN {              // conv3x3 (BN folded)
    PackedB w; float* b; int cin, cout, bk;
};
struct Block {               // ConvBlockRes
    ConvBN c1, c2;
    bool has_sc = false;
    PackedB sc_w; float* sc_b = nullptr;
};

int bk_for(int cin) { return cin >= 64 ? 64 : (cin >= 32 ? 32 : 16); }

struct BNFold {
    std::vector<float> s, b;
};
BNFold fold_bn(const rvcb_weights& w, const std::string& p) {
    const WT &g = w.get(p + "weight"), &be = w.get(p + "bias"), &m = w.get(p + "running_mean"), &v = w.get(p + "running_var");
    BNFold f;
    const size_t n = g.data.size();
    f.s.resize(n); f.b.resize(n);
    for (size_t i = 0; i < n; ++i) {
        const float sc = g.data[i] / std::sqrt(v.data[i] + 1e-5f);
        f.s[i] = sc;
        f.b[i] = be.data[i] - m.data[i] * sc;
    }
    return f;
}

}  // namespace

struct rvcb_rmvpe {
    DevOwner own;
    Arena arena;
    float* dft = nullptr;         // [1026, 1024] windowed cos | -sin
    __half* dft16 = nullptr;      // [1088, 3072] the same basis as fp16 (hi | lo | hi / 2048): split-precision tensor-core DFT
    float* melw = nullptr;        // [128, 513]
    // stem (encoder level 0 block 0, C_in = 1): bn0 scalar affine, conv1 1->16 (BN folded), shortcut 1->16
    float bn0_s = 1.f, bn0_b = 0.f;
    float* stem = nullptr;        // [16*9 w1 | 16 b1 | 16 sc_w | 16 sc_b]
    ConvBN stem_c2;
    std::vector<std::vector<Block>> enc, inter, dec;   // [level][block]  (enc[0] has 3 blocks: the stem replaces block 0)
    struct Up { PackedB w; float* b; int cin, cout, bk; };
    std::vector<Up> ups;
    PackedB cnn_w; float* cnn_b = nullptr;
    PackedB wih; float* bih = nullptr;     // [1536, 384] (fwd | bwd), columns permuted to (mel*3 + c)
    float* whh = nullptr;                  // [2][768][256] fp32
    float* bhh = nullptr;                  // [2][768]
    PackedB fc_w; float* fc_b = nullptr;
};

static Block load_block(DevOwner& own, const rvcb_weights& w, const std::string& p, int cin, int cout) {
    Block B;
    BNFold f1 = fold_bn(w, p + "conv.1."), f2 = fold_bn(w, p + "conv.4.");
    B.c1 = {pack_conv2d_3x3(own, w.get(p + "conv.0.weight").data.data(), cout, cin, bk_for(cin), f1.s.data()), own.upload(f1.b), cin, cout,
            bk_for(cin)};
    B.c2 = {pack_conv2d_3x3(own, w.get(p + "conv.3.weight").data.data(), cout, cout, bk_for(cout), f2.s.data()), own.upload(f2.b), cout,
            cout, bk_for(cout)};
    if (cin != cout) {
        B.has_sc = true;
        B.sc_w = pack_linear(own, w.get(p + "shortcut.weight").data.data(), cout, cin, bk_for(cin));
        B.sc_b = own.upload(w.get(p + "shortcut.bias").data);
    }
    return B;
}

static rvcb_rmvpe* rmvpe_build(const rvcb_weights& w) {
    auto* h = new rvcb_rmvpe();
    try {
        DevOwner& own = h->own;
        {   // DFT basis with the periodic Hann window folded in (torch.stft, stft.py:171-180)
            std::vector<float> d((size_t)2 * N_BINS * N_FFT);
            for (int k = 0; k < N_BINS; ++k)
                for (int n = 0; n < N_FFT; ++n) {
                    const double win = 0.5 - 0.5 * std::cos(2.0 * M_PI * n / N_FFT);
                    const double ph = 2.0 * M_PI * (double)((long)k * n % N_FFT) / N_FFT;
                    d[(size_t)k * N_FFT + n] = (float)(win * std::cos(ph));
                    d[(size_t)(N_BINS + k) * N_FFT + n] = (float)(-win * std::sin(ph));
                }
            h->dft = own.upload(d);
            // b = hi + lo with hi = fp16(b), lo = fp16(b - hi): 22 significant bits; the third copy (hi / 2048, exact unless
            // subnormal) pairs with the x2048-scaled low half of the signal (rmvpe_split_kernel)
            std::vector<__half> d16((size_t)DFT_NPAD * 3 * N_FFT, __float2half(0.f));
            for (int r = 0; r < 2 * N_BINS; ++r)
                for (int n = 0; n < N_FFT; ++n) {
                    const float b = d[(size_t)r * N_FFT + n];
                    const __half hi = __float2half_rn(b);
                    const float hf = __half2float(hi);
                    d16[(size_t)r * 3 * N_FFT + n] = hi;
                    d16[(size_t)r * 3 * N_FFT + N_FFT + n] = __float2half_rn(b - hf);
                    d16[(size_t)r * 3 * N_FFT + 2 * N_FFT + n] = __float2half_rn(hf * (1.f / 2048.f));
                }
            h->dft16 = own.upload(d16);
            // librosa.filters.mel(sr=16000, n_fft=1024, n_mels=128, fmin=30, fmax=8000, htk=True), slaney norm (mel.py:27-34)
            std::vector<double> melf(N_MELS + 2);
            const double lo = 2595.0 * std::log10(1.0 + 30.0 / 700.0), hi = 2595.0 * std::log10(1.0 + 8000.0 / 700.0);
            for (int i = 0; i < N_MELS + 2; ++i) melf[i] = 700.0 * (std::pow(10.0, (lo + (hi - lo) * i / (N_MELS + 1)) / 2595.0) - 1.0);
            std::vector<float> mw((size_t)N_MELS * N_BINS);
            for (int i = 0; i < N_MELS; ++i) {
                const double enorm = 2.0 / (melf[i + 2] - melf[i]);
                for (int k = 0; k < N_BINS; ++k) {
                    const double fr = 8000.0 * k / (N_BINS - 1);
                    const double lower = (fr - melf[i]) / (melf[i + 1] - melf[i]);
                    const double upper = (melf[i + 2] - fr) / (melf[i + 2] - melf[i + 1]);
                    mw[(size_t)i * N_BINS + k] = (float)(std::max(0.0, std::min(lower, upper)) * enorm);
                }
            }
            h->melw = own.upload(mw);
        }
        {   // stem
            BNFold b0 = fold_bn(w, "unet.encoder.bn.");
            h->bn0_s = b0.s[0]; h->bn0_b = b0.b[0];
            const std::string p = "unet.encoder.layers.0.conv.0.";
            BNFold f1 = fold_bn(w, p + "conv.1."), f2 = fold_bn(w, p + "conv.4.");
            const WT& w1 = w.get(p + "conv.0.weight");                   // [16,1,3,3]
            RVCB_CHECK(w1.dim(0) == 16 && w1.dim(1) == 1, "rmvpe: unexpected stem shape");
            std::vector<float> st(16 * 9 + 48);
            for (int co = 0; co < 16; ++co) {
                for (int t = 0; t < 9; ++t) st[co * 9 + t] = w1.data[co * 9 + t] * f1.s[co];
                st[144 + co] = f1.b[co];
                st[160 + co] = w.get(p + "shortcut.weight").data[co];
                st[176 + co] = w.get(p + "shortcut.bias").data[co];
            }
            h->stem = own.upload(st);
            h->stem_c2 = {pack_conv2d_3x3(own, w.get(p + "conv.3.weight").data.data(), 16, 16, 16, f2.s.data()), own.upload(f2.b), 16, 16, 16};
        }
        int cin = 16, cout = 16;
        for (int l = 0; l < 5; ++l) {
            std::vector<Block> L;
            for (int b = 0; b < 4; ++b) {
                if (l == 0 && b == 0) continue;
                L.push_back(load_block(own, w, "unet.encoder.layers." + std::to_string(l) + ".conv." + std::to_string(b) + ".",
                                       b == 0 ? cin : cout, cout));
            }
            h->enc.push_back(L);
            cin = cout; cout *= 2;
        }
        // cin = 256, cout = 512 here
        for (int l = 0; l < 4; ++l) {
            std::vector<Block> L;
            for (int b = 0; b < 4; ++b)
                L.push_back(load_block(own, w, "unet.intermediate.layers." + std::to_string(l) + ".conv." + std::to_string(b) + ".",
                                       (l == 0 && b == 0) ? cin : cout, cout));
            h->inter.push_back(L);
        }
        int dc = cout;   // 512
        for (int l = 0; l < 5; ++l) {
            const int oc = dc / 2;
            const std::string p = "unet.decoder.layers." + std::to_string(l) + ".";
            BNFold f = fold_bn(w, p + "conv1.1.");
            rvcb_rmvpe::Up U;
            U.cin = dc; U.cout = oc; U.bk = bk_for(dc);
            U.w = pack_convT2d_up2(own, w.get(p + "conv1.0.weight").data.data(), dc, oc, U.bk, f.s.data());
            std::vector<float> b4((size_t)4 * oc);
            for (int q = 0; q < 4; ++q)
                for (int co = 0; co < oc; ++co) b4[(size_t)q * oc + co] = f.b[co];
            U.b = own.upload(b4);
            h->ups.push_back(U);
            std::vector<Block> L;
            for (int b = 0; b < 4; ++b) L.push_back(load_block(own, w, p + "conv2." + std::to_string(b) + ".", b == 0 ? 2 * oc : oc, oc));
            h->dec.push_back(L);
            dc = oc;
        }
        h->cnn_w = pack_conv2d_3x3(own, w.get("cnn.weight").data.data(), 3, 16, 16);
        h->cnn_b = own.upload(w.get("cnn.bias").data);
        {   // GRU: input projection for both directions; column permutation (c*128+mel) -> (mel*3+c)
            std::vector<float> wi((size_t)1536 * 384), bi(1536), whh((size_t)2 * 768 * 256), bhh(2 * 768);
            for (int dir = 0; dir < 2; ++dir) {
                const std::string sfx = dir ? "_reverse" : "";
                const WT& a = w.get("fc.0.gru.weight_ih_l0" + sfx);
                for (int r = 0; r < 768; ++r)
                    for (int c2 = 0; c2 < 3; ++c2)
                        for (int m = 0; m < 128; ++m) wi[((size_t)dir * 768 + r) * 384 + m * 3 + c2] = a.data[(size_t)r * 384 + c2 * 128 + m];
                memcpy(&bi[dir * 768], w.get("fc.0.gru.bias_ih_l0" + sfx).data.data(), 768 * sizeof(float));
                memcpy(&whh[(size_t)dir * 768 * 256], w.get("fc.0.gru.weight_hh_l0" + sfx).data.data(), (size_t)768 * 256 * sizeof(float));
                memcpy(&bhh[dir * 768], w.get("fc.0.gru.bias_hh_l0" + sfx).data.data(), 768 * sizeof(float));
            }
            h->wih = pack_linear(own, wi.data(), 1536, 384);
            h->bih = own.upload(bi);
            h->whh = own.upload(whh);
            h->bhh = own.upload(bhh);
        }
        h->fc_w = pack_linear(own, w.get("fc.1.weight").data.data(), N_CLASS, 512);
        h->fc_b = own.upload(w.get("fc.1.bias").data);
    } catch (...) {
        delete h;
        throw;
    }
    return h;
}

// ------------------------------------------------------------------------------------------------
// SIMT kernels
// ------------------------------------------------------------------------------------------------
__global__ void reflect_pad_kernel(const float* __restrict__ x, int n, float* __restrict__ y, int pad) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)n + 2 * pad) return;
    long j = i - pad;
    if (j < 0) j = -j;
    if (j >= n) j = 2L * (n - 1) - j;
    y[i] = x[j];
}

// Reflect-padded signal as two fp16 planes of one [2R, HOP] matrix: rows [0, R) hold hi = fp16(x), rows [R, 2R) hold
// fp16((x - hi) * 2048); elements past n + 2 * pad are zero.
__global__ void rmvpe_split_kernel(const float* __restrict__ x, int n, int pad, int R, __half* __restrict__ y) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)R * HOP) return;
    float v = 0.f;
    if (i < (long)n + 2 * pad) {
        long j = i - pad;
        if (j < 0) j = -j;
        if (j >= n) j = 2L * (n - 1) - j;
        v = x[j];
    }
    const __half hi = __float2half_rn(v);
    y[i] = hi;
    y[(long)R * HOP + i] = __float2half_rn((v - __half2float(hi)) * 2048.f);
}

__global__ void magnitude_kernel(const float* __restrict__ spec, float* __restrict__ mag, int nf) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)nf * N_BINS) return;
    const int t = (int)(i / N_BINS), k = (int)(i - (long)t * N_BINS);
    const float re = spec[(long)t * 2 * N_BINS + k], im = spec[(long)t * 2 * N_BINS + N_BINS + k];
    mag[i] = sqrtf(re * re + im * im);
}

// logmel [Tpad, 128] (time-major, zero = "constant 0" padding frames), optional transposed copy [128, nf]
__global__ void logmel_kernel(const float* __restrict__ mel, int nf, int Tpad, float* __restrict__ out, float* __restrict__ out_t) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)Tpad * N_MELS) return;
    const int t = (int)(i / N_MELS), m = (int)(i - (long)t * N_MELS);
    float v = 0.f;
    if (t < nf) {
        v = logf(fmaxf(mel[i], 1e-5f));
        if (out_t) out_t[(long)m * nf + t] = v;
    }
    out[i] = v;
}

// stem: x = bn0(mel) (zero padded AFTER bn); y1 = relu(conv3x3_1->16 (BN folded)); sc = shortcut(x)
__global__ void __launch_bounds__(256) stem_kernel(const float* __restrict__ mel, int T, float s0, float b0, const float* __restrict__ prm,
                                                   __half* __restrict__ y1, float* __restrict__ sc) {
    __shared__ float xs[4][130];
    __shared__ float p[192];
    const int t0 = blockIdx.x * 2;
    for (int i = threadIdx.x; i < 192; i += 256) p[i] = prm[i];
    for (int i = threadIdx.x; i < 4 * 130; i += 256) {
        const int r = i / 130, cidx = i - r * 130;
        const int t = t0 - 1 + r, wv = cidx - 1;
        xs[r][cidx] = (t >= 0 && t < T && wv >= 0 && wv < 128) ? mel[(long)t * 128 + wv] * s0 + b0 : 0.f;
    }
    __syncthreads();
    const int tl = threadIdx.x >> 7, wv = threadIdx.x & 127;
    const int t = t0 + tl;
    if (t >= T) return;
    float xin[9];
#pragma unroll
    for (int dh = 0; dh < 3; ++dh)
#pragma unroll
        for (int dw = 0; dw < 3; ++dw) xin[dh * 3 + dw] = xs[tl + dh][wv + dw];
    const float xc = xin[4];
    const long o = ((long)t * 128 + wv) * 16;
    __align__(16) __half hv[16];
    __align__(16) float sv[16];
#pragma unroll
    for (int co = 0; co < 16; ++co) {
        float acc = p[144 + co];
#pragma unroll
        for (int k = 0; k < 9; ++k) acc = fmaf(p[co * 9 + k], xin[k], acc);
        hv[co] = __float2half_rn(fmaxf(acc, 0.f));
        sv[co] = fmaf(p[160 + co], xc, p[176 + co]);
    }
    *reinterpret_cast<uint4*>(y1 + o) = *reinterpret_cast<const uint4*>(hv);
    *reinterpret_cast<uint4*>(y1 + o + 8) = *reinterpret_cast<const uint4*>(hv + 8);
#pragma unroll
    for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(sc + o + q * 4) = *reinterpret_cast<const float4*>(sv + q * 4);
}

// 2x2 average pool of fp32 [H, W, C] -> fp16 [H/2, W/2, C]
__global__ void avgpool_kernel(const float* __restrict__ x, int H, int W, int C, __half* __restrict__ y) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int Ho = H / 2, Wo = W / 2;
    if (i >= (long)Ho * Wo * C) return;
    const int c = (int)(i % C);
    const long pix = i / C;
    const int wo = (int)(pix % Wo), ho = (int)(pix / Wo);
    const float* b = x + ((long)(2 * ho) * W + 2 * wo) * C + c;
    const float v = (b[0] + b[C] + b[(long)W * C] + b[(long)W * C + C]) * 0.25f;
    y[i] = __float2half_rn(v);
}

// BiGRU recurrence: one 8-CTA cluster per direction.  Each CTA owns 32 hidden units = 96 rows of W_hh, resident in
// REGISTERS (96 fp32 values per thread).  A warp owns 4 units x 3 gates; 8 lanes share one row (float4 k-slices) and
// reduce with 3 shuffles, so a unit's r/z/n sums land in one lane that applies the gates and publishes h_t to all 8 CTAs
// through distributed shared memory.  One split cluster barrier (arrive ... wait) per step; no block barrier.
constexpr int GRU_H = 256, GRU_CL = 8, GRU_U = GRU_H / GRU_CL;     // 32 hidden units per CTA
__global__ void __cluster_dims__(GRU_CL, 1, 1) __launch_bounds__(256)
gru_kernel(const float* __restrict__ gi /*[T,1536]*/, const float* __restrict__ whh /*[2,768,256]*/, const float* __restrict__ bhh /*[2,768]*/,
           int T, float* __restrict__ out32 /*[T,512] or null*/, __half* __restrict__ out16 /*[T,512]*/) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int dir = blockIdx.x / GRU_CL;
    __shared__ __align__(16) float hbuf[2 * GRU_H];      // double-buffered hidden state, written by all 8 CTAs through DSMEM
    const float* Wd = whh + (size_t)dir * 768 * 256;
    for (int i = threadIdx.x; i < 2 * GRU_H; i += blockDim.x) hbuf[i] = 0.f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rsub = lane >> 3, ks = lane & 7;                      // unit within the warp, k-slice
    const int ul = warp * 4 + rsub;                                 // local unit 0..31
    const int unit = rank * GRU_U + ul;                             // hidden unit 0..255
    const bool leader = (ks == 0);
    const float b_r = bhh[dir * 768 + unit], b_z = bhh[dir * 768 + GRU_H + unit], b_n = bhh[dir * 768 + 2 * GRU_H + unit];
    // this lane's slice of W_hh stays in REGISTERS for the whole sequence: 3 gates x 8 float4 (k = 4*(ks + 8i) .. +3)
    float4 wreg[3][8];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int i = 0; i < 8; ++i)
            wreg[g][i] = __ldg(reinterpret_cast<const float4*>(Wd + (size_t)(g * GRU_H + unit) * 256) + ks + 8 * i);
    float* remote[GRU_CL];
#pragma unroll
    for (int cr = 0; cr < GRU_CL; ++cr) remote[cr] = cluster.map_shared_rank(hbuf, cr);
    __syncthreads();
    cluster.sync();
    int cur = 0;
    int t = dir == 0 ? 0 : T - 1;
    const int dt = dir == 0 ? 1 : -1;
    float gir = 0.f, giz = 0.f, gin = 0.f;
    if (leader && T > 0) {
        const float* g0 = gi + (size_t)t * 1536 + dir * 768 + unit;
        gir = __ldg(g0); giz = __ldg(g0 + GRU_H); gin = __ldg(g0 + 2 * GRU_H);
    }
    for (int step = 0; step < T; ++step, t += dt) {
        // prefetch the NEXT step's input projection (independent of the recurrence)
        float nr = 0.f, nz = 0.f, nn = 0.f;
        if (leader && step + 1 < T) {
            const float* g1 = gi + (size_t)(t + dt) * 1536 + dir * 768 + unit;
            nr = __ldg(g1); nz = __ldg(g1 + GRU_H); nn = __ldg(g1 + 2 * GRU_H);
        }
        const float4* hc = reinterpret_cast<const float4*>(hbuf + cur * 256);
        float4 hv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) hv[i] = hc[ks + 8 * i];
        float sums[3];
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 w4 = wreg[g][i];
                a0 = fmaf(w4.x, hv[i].x, a0); a1 = fmaf(w4.y, hv[i].y, a1);
                a2 = fmaf(w4.z, hv[i].z, a2); a3 = fmaf(w4.w, hv[i].w, a3);
            }
            float a = (a0 + a1) + (a2 + a3);
            a += __shfl_xor_sync(0xffffffffu, a, 4);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            sums[g] = a;
        }
        if (leader) {
            const float r = 1.f / (1.f + expf(-(gir + sums[0] + b_r)));
            const float z = 1.f / (1.f + expf(-(giz + sums[1] + b_z)));
            const float n = tanhf(gin + r * (sums[2] + b_n));
            const float hprev = hbuf[cur * 256 + unit];
            const float hn = (1.f - z) * n + z * hprev;
            const int o = (cur ^ 1) * 256 + unit;
#pragma unroll
            for (int cr = 0; cr < GRU_CL; ++cr) remote[cr][o] = hn;
            const size_t og = (size_t)t * 512 + dir * 256 + unit;
            if (out32) out32[og] = hn;
            out16[og] = __float2half_rn(hn);
        }
        cluster.barrier_arrive();          // release: this step's DSMEM writes
        gir = nr; giz = nz; gin = nn;
        cluster.barrier_wait();            // acquire: everyone's h_t is visible
        cur ^= 1;
    }
}

// ---- v2 recurrence: same ownership (8-CTA cluster per direction, W_hh in registers), but h_t is published with
// st.async (remote shared-memory store that completes a transaction count on the DESTINATION CTA's mbarrier), so a step
// costs one DSMEM store latency + one mbarrier wake-up instead of a full release/acquire cluster barrier (which also has
// to drain the step's global stores and flushes L1).  Every lane of a unit group evaluates the gates redundantly; lane
// j < 8 of each warp then sends the warp's 4 consecutive h values as one float4 to CTA j -> one store instruction per
// warp per step.  Two mbarriers (one per h buffer) keep step t+1 and t+2 transactions apart; each expects 1024 B/step.
__device__ __forceinline__ uint32_t gru_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t gru_mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void gru_st_async_v4(uint32_t raddr, float a, float b, float c, float d, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];"
                 :: "r"(raddr), "f"(a), "f"(b), "f"(c), "f"(d), "r"(rbar) : "memory");
}
__device__ __forceinline__ void gru_bar_arm(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gru_bar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "GRU_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra GRU_DONE_%=;\n\t"
        "bra GRU_WAIT_%=;\n\t"
        "GRU_DONE_%=:\n\t}"
        :: "r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ float gru_ldg(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
template <bool FAST, bool TRACE = false>
__global__ void __cluster_dims__(GRU_CL, 1, 1) __launch_bounds__(256)
gru_kernel_v2(const float* __restrict__ gi /*[T,1536]*/, const float* __restrict__ whh /*[2,768,256]*/, const float* __restrict__ bhh /*[2,768]*/,
              int T, float* __restrict__ out32 /*[T,512] or null*/, __half* __restrict__ out16 /*[T,512]*/) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int dir = blockIdx.x / GRU_CL;
    __shared__ __align__(16) float hbuf[2 * GRU_H];      // h_s lives in hbuf[(s & 1) * 256 ...]
    __shared__ __align__(8) uint64_t hbar[2];            // hbar[b] completes when all 256 values of buffer b arrived
    const float* Wd = whh + (size_t)dir * 768 * 256;
    for (int i = threadIdx.x; i < 2 * GRU_H; i += blockDim.x) hbuf[i] = 0.f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rsub = lane >> 3, ks = lane & 7;
    const int ul = warp * 4 + rsub;
    const int unit = rank * GRU_U + ul;
    const float b_r = bhh[dir * 768 + unit], b_z = bhh[dir * 768 + GRU_H + unit], b_n = bhh[dir * 768 + 2 * GRU_H + unit];
    float4 wreg[3][8];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int i = 0; i < 8; ++i)
            wreg[g][i] = __ldg(reinterpret_cast<const float4*>(Wd + (size_t)(g * GRU_H + unit) * 256) + ks + 8 * i);
    const uint32_t bar0 = gru_smem_u32(&hbar[0]);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar0 + 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        gru_bar_arm(bar0 + 8, GRU_H * 4);      // h_1 -> buffer 1
        gru_bar_arm(bar0, GRU_H * 4);          // h_2 -> buffer 0
    }
    // lane j < 8 publishes this warp's float4 (units rank*32 + warp*4 .. +3) to CTA j
    const uint32_t dst_h = gru_mapa(gru_smem_u32(hbuf) + (uint32_t)(rank * GRU_U + warp * 4) * 4u, (uint32_t)(lane & 7));
    const uint32_t dst_b = gru_mapa(bar0, (uint32_t)(lane & 7));
    __syncthreads();
    cluster.sync();
    int t = dir == 0 ? 0 : T - 1;
    const int dt = dir == 0 ? 1 : -1;
    // input projections of step s (gi*), s + 1 (p*) and, issued at the top of step s, s + 2 (q*).  The loads are volatile asm:
    // the compiler otherwise sinks them to their first use at the end of the step and the L2 latency (~300 cycles) lands on the
    // serial chain (measured: the ex2.approx gate variant came out SLOWER than libm for exactly this reason, profiles/r2k)
    float gir = 0.f, giz = 0.f, gin = 0.f, pr = 0.f, pz = 0.f, pn = 0.f;
    {
        const float* g0 = gi + (size_t)t * 1536 + dir * 768 + unit;
        if (T > 0) { gir = gru_ldg(g0); giz = gru_ldg(g0 + GRU_H); gin = gru_ldg(g0 + 2 * GRU_H); }
        if (T > 1) { g0 += dt * 1536; pr = gru_ldg(g0); pz = gru_ldg(g0 + GRU_H); pn = gru_ldg(g0 + 2 * GRU_H); }
    }
    long long tr[5] = {0, 0, 0, 0, 0}, tc0 = 0, tc1 = 0, tc2 = 0, tc3 = 0, tc4 = 0;      // RVCB_GRU_TRACE: where a step's cycles go
    for (int s = 0; s < T; ++s, t += dt) {
        if (TRACE) {
            const long long now = clock64();
            if (s > 64 && s <= 64 + 1024) tr[4] += now - tc4;
            tc0 = now;
        }
        float qr = 0.f, qz = 0.f, qn = 0.f;
        if (s + 2 < T) {
            const float* g2 = gi + (size_t)(t + 2 * dt) * 1536 + dir * 768 + unit;
            qr = gru_ldg(g2); qz = gru_ldg(g2 + GRU_H); qn = gru_ldg(g2 + 2 * GRU_H);
        }
        const int cur = s & 1;
        if (s > 0) {
            gru_bar_wait(bar0 + 8u * cur, (uint32_t)(((s - 1) >> 1) & 1));
            if (threadIdx.x == 0 && s + 2 <= T) gru_bar_arm(bar0 + 8u * cur, GRU_H * 4);      // for h_{s+2}
        }
        if (TRACE) tc1 = clock64();
        const float4* hc = reinterpret_cast<const float4*>(hbuf + cur * 256);
        float4 hv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) hv[i] = hc[ks + 8 * i];
        const float hprev = hbuf[cur * 256 + unit];
        float sums[3];
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 w4 = wreg[g][i];
                a0 = fmaf(w4.x, hv[i].x, a0); a1 = fmaf(w4.y, hv[i].y, a1);
                a2 = fmaf(w4.z, hv[i].z, a2); a3 = fmaf(w4.w, hv[i].w, a3);
            }
            float a = (a0 + a1) + (a2 + a3);
            a += __shfl_xor_sync(0xffffffffu, a, 4);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            sums[g] = a;
        }
        if (TRACE) { tc2 = clock64(); tc2 += (long long)(sums[0] == 12345.f); }
        // the gate math sits on the serial per-step chain (in-kernel clock64 trace, RVCB_GRU_TRACE=1: 350 of ~1400 cycles per step
        // with the libm forms).  FAST (default; RVCB_GRU_FAST=0 = libm): exp through ex2.approx (2 ulp), tanh(x) = 2 sigmoid(2x) - 1
        // (absolute error ~1e-7, far below the fp16 operand rounding of the layers around the GRU): 173 cycles, RMVPE on 16 s
        // 2.98 -> 2.80 ms -- but only with the input-projection loads pinned early (see gru_ldg above)
        float r, z, n;
        if (FAST) {
            r = __fdividef(1.f, 1.f + __expf(-(gir + sums[0] + b_r)));
            z = __fdividef(1.f, 1.f + __expf(-(giz + sums[1] + b_z)));
            n = __fdividef(2.f, 1.f + __expf(-2.f * (gin + r * (sums[2] + b_n)))) - 1.f;
        } else {
            r = 1.f / (1.f + expf(-(gir + sums[0] + b_r)));
            z = 1.f / (1.f + expf(-(giz + sums[1] + b_z)));
            n = tanhf(gin + r * (sums[2] + b_n));
        }
        const float hn = (1.f - z) * n + z * hprev;
        if (TRACE) { tc3 = clock64(); tc3 += (long long)(hn == 12345.f); }
        const float h0 = __shfl_sync(0xffffffffu, hn, 0), h1 = __shfl_sync(0xffffffffu, hn, 8);
        const float h2 = __shfl_sync(0xffffffffu, hn, 16), h3 = __shfl_sync(0xffffffffu, hn, 24);
        if (s + 1 < T && lane < GRU_CL)
            gru_st_async_v4(dst_h + (uint32_t)((cur ^ 1) * 256) * 4u, h0, h1, h2, h3, dst_b + 8u * (cur ^ 1));
        if (ks == 0) {
            const size_t og = (size_t)t * 512 + dir * 256 + unit;
            if (out32) out32[og] = hn;
            out16[og] = __float2half_rn(hn);
        }
        gir = pr; giz = pz; gin = pn;
        pr = qr; pz = qz; pn = qn;
        if (TRACE) {
            tc4 = clock64();
            if (s >= 64 && s < 64 + 1024) { tr[0] += tc1 - tc0; tr[1] += tc2 - tc1; tr[2] += tc3 - tc2; tr[3] += tc4 - tc3; }
        }
    }
    if (TRACE && blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 37))
        printf("[gru trace] thread %d: avg cycles / step: wait %lld, load h + matvec + reduce %lld, gates %lld, publish + store %lld, loop %lld\n",
               (int)threadIdx.x, tr[0] >> 10, tr[1] >> 10, tr[2] >> 10, tr[3] >> 10, tr[4] >> 10);
    cluster.sync();      // nobody exits while a peer may still be storing into its shared memory
}

// salience [T,360] -> f0 [T]  (rmvpe.py:119-137,157-164); one warp per frame
__global__ void decode_kernel(const float* __restrict__ sal, int T, float thred, float* __restrict__ f0) {
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= T) return;
    const float* s = sal + (long)t * N_CLASS;
    float best = -INFINITY;
    int bi = 0;
    for (int b = lane; b < N_CLASS; b += 32) {
        const float v = s[b];
        if (v > best) { best = v; bi = b; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) {
        double ps = 0.0, ws = 0.0;
        for (int b = bi - 4; b <= bi + 4; ++b) {
            if (b < 0 || b >= N_CLASS) continue;
            const double sv = (double)s[b];
            ps += sv * (20.0 * b + 1997.3794084376191);
            ws += sv;
        }
        double cents = ps / ws;
        if (best <= thred) cents = 0.0;
        double f = 10.0 * exp2(cents / 1200.0);
        if (f == 10.0) f = 0.0;
        f0[t] = (float)f;
    }
}

// ------------------------------------------------------------------------------------------------
// graph
// ------------------------------------------------------------------------------------------------
static void conv3x3(const ConvBN& c, const __half* x, long ldx, int Hh, int W, int act, const float* res2, long ldres2, float* out32,
                    long ld32, __half* out16, long ld16, cudaStream_t st) {
    if (act == ACT_RELU || act == ACT_NONE) {      // full-resolution 16-channel levels: one halo load per image row (conv2d_row.cu)
        const Conv2dRowArgs a{x, ldx, Hh, W, c.cin, c.cout, c.w.d, c.w.rows, c.w.cols, c.b, act == ACT_RELU, res2, ldres2, out32, ld32, out16, ld16};
        if (conv2d_row_try(a, st)) return;
    }
    GemmArgs g;
    g.A = x; g.lda = ldx; g.a_rows = Hh; g.a_cols = c.cin; g.conv2d_W = W;
    g.B = c.w.d; g.ldb = c.w.cols; g.b_rows = c.w.rows; g.b_cols = c.w.cols;
    g.M = Hh * W; g.N = c.cout; g.block_k = c.bk;
    seg_conv2d_3x3(g, c.cin);
    g.bias = c.b; g.act1 = act;
    g.res2 = res2; g.ldres2 = ldres2;
    g.out32 = out32; g.ld32 = ld32; g.out16 = out16; g.ld16 = ld16;
    gemm(g, st);
}

// ConvBlockRes: out = relu(bn(conv2(relu(bn(conv1(x)))))) + (shortcut(x) | x)
static void run_block(const Block& B, const __half* x16, long ldx, const float* x32, int Hh, int W, __half* t16, float* sc32, float* out32,
                      __half* out16, long ld16, cudaStream_t st) {
    const float* res = x32;
    if (B.has_sc) {
        GemmArgs g;
        g.A = x16; g.lda = ldx; g.a_rows = Hh; g.a_cols = B.c1.cin; g.conv2d_W = W;
        g.B = B.sc_w.d; g.ldb = B.sc_w.cols; g.b_rows = B.sc_w.rows; g.b_cols = B.sc_w.cols;
        g.M = Hh * W; g.N = B.c1.cout; g.block_k = B.c1.bk;
        g.nseg = 1; g.seg[0] = {0, 0, 0, ceil_div(B.c1.cin, B.c1.bk)};
        g.bias = B.sc_b; g.out32 = sc32; g.ld32 = B.c1.cout;
        gemm(g, st);
        res = sc32;
    }
    conv3x3(B.c1, x16, ldx, Hh, W, ACT_RELU, nullptr, 0, nullptr, 0, t16, B.c1.cout, st);
    conv3x3(B.c2, t16, B.c2.cin, Hh, W, ACT_RELU, res, B.c2.cout, out32, B.c2.cout, out16, ld16, st);
}

static int rmvpe_frames(int n) { return n / HOP + 1; }

static void rmvpe_forward(rvcb_rmvpe* h, const float* d_wav, int n, float thred, float* d_mel, float* d_hidden, float* d_f0, int* n_frames,
                          cudaStream_t st) {
    RVCB_CHECK(n > N_FFT / 2, "rmvpe: input too short for reflect padding");
    const int nf = rmvpe_frames(n);
    const int T = round_up(nf, 32);
    auto rnd = [](size_t b) { return (b + 1023) & ~size_t(1023); };
    size_t need = 64u << 20;
    need += 2 * rnd(((size_t)n + N_FFT + 4 * HOP) * 4) + rnd((size_t)nf * 2 * N_BINS * 4) + rnd((size_t)nf * N_BINS * 4) + 2 * rnd((size_t)T * 128 * 4);
    const size_t px = (size_t)T * 128;
    need += 3 * rnd(px * 16 * 4) + 3 * rnd(px * 16 * 2);        // ping-pong fp32, sc32, fp16 x2, t16 (level-0 size bounds all levels)
    for (int l = 0; l < 5; ++l) need += rnd((px >> (2 * l)) * (32u << l) * 2);   // concat buffers
    need += rnd(px * 16 * 2) + rnd((size_t)T * 384 * 2) + rnd((size_t)T * 1536 * 4) + rnd((size_t)T * 512 * 2) + rnd((size_t)T * 360 * 4);
    h->arena.reserve(need);
    h->arena.reset();
    Arena& ar = h->arena;

    // ---- log-mel ----
    float* spec = ar.alloc<float>((size_t)nf * 2 * N_BINS);
    static const bool mel_fp32 = [] { const char* e = getenv("RVCB_MEL_FP32"); return e && e[0] == '1'; }();
    if (mel_fp32) {
        float* wpad = ar.alloc<float>((size_t)n + N_FFT);
        reflect_pad_kernel<<<(unsigned)ceil_div_l((long)n + N_FFT, 256), 256, 0, st>>>(d_wav, n, wpad, N_FFT / 2);
        KERNEL_CHECK();
        sgemm_nt(wpad, HOP, h->dft, N_FFT, spec, 2 * N_BINS, nf, 2 * N_BINS, N_FFT, st);       // overlapping frames: lda = hop
    } else {
        // Split-precision DFT on the tensor cores: x = hi + lo / 2048 and b = hi + lo in fp16, the three significant products
        // (hi.hi, hi.lo, lo.hi) accumulate in fp32 (the dropped lo.lo term is 2^-22 relative).  A frame is the signal viewed as
        // rows of HOP samples: frame t = rows t .. t+6 (6 x 160 + 64 = 1024), i.e. a 7-tap "convolution" over hop-sized rows,
        // so the operand needs no overlapping-row tensor map and no im2col.
        const int R = (int)ceil_div_l((long)n + N_FFT, HOP) + 1;
        __half* ws = ar.alloc<__half>((size_t)2 * R * HOP + 64);
        rmvpe_split_kernel<<<(unsigned)ceil_div_l((long)R * HOP, 256), 256, 0, st>>>(d_wav, n, N_FFT / 2, R, ws);
        KERNEL_CHECK();
        GemmArgs g;
        g.A = ws; g.lda = HOP; g.a_rows = 2 * R; g.a_cols = HOP;
        g.B = h->dft16; g.ldb = 3 * N_FFT; g.b_rows = 2 * N_BINS; g.b_cols = 3 * N_FFT;
        g.M = nf; g.N = 2 * N_BINS; g.block_k = 32;
        g.nseg = 0;
        for (int part = 0; part < 3; ++part)
            for (int j = 0; j < 7; ++j) g.seg[g.nseg++] = {(part == 2 ? R : 0) + j, 0, 0, j < 6 ? HOP / 32 : (N_FFT - 6 * HOP) / 32};
        g.out32 = spec; g.ld32 = 2 * N_BINS;
        gemm(g, st);
    }
    float* mag = ar.alloc<float>((size_t)nf * N_BINS);
    magnitude_kernel<<<(unsigned)ceil_div_l((long)nf * N_BINS, 256), 256, 0, st>>>(spec, mag, nf);
    KERNEL_CHECK();
    float* melp = ar.alloc<float>((size_t)T * N_MELS);
    sgemm_nt(mag, N_BINS, h->melw, N_BINS, melp, N_MELS, nf, N_MELS, N_BINS, st);
    float* logmel = ar.alloc<float>((size_t)T * N_MELS);
    logmel_kernel<<<(unsigned)ceil_div_l((long)T * N_MELS, 256), 256, 0, st>>>(melp, nf, T, logmel, d_mel);
    KERNEL_CHECK();
    count_launch(3);
    if (!d_hidden && !d_f0) {
        if (n_frames) *n_frames = nf;
        return;
    }
    // ---- U-Net ----
    float* a32 = ar.alloc<float>(px * 16);
    float* b32 = ar.alloc<float>(px * 16);
    float* sc32 = ar.alloc<float>(px * 16);
    __half* a16 = ar.alloc<__half>(px * 16);
    __half* b16 = ar.alloc<__half>(px * 16);
    __half* t16 = ar.alloc<__half>(px * 16);
    __half* cat[5];
    for (int l = 0; l < 5; ++l) cat[l] = ar.alloc<__half>((px >> (2 * l)) * (32u << l));
    // level-0 stem
    stem_kernel<<<ceil_div(T, 2), 256, 0, st>>>(logmel, T, h->bn0_s, h->bn0_b, h->stem, t16, sc32);
    KERNEL_CHECK();
    count_launch();
    int Hh = T, W = 128, C = 16;
    float* cur32 = a32; float* nxt32 = b32;
    __half* cur16 = a16; __half* nxt16 = b16;
    conv3x3(h->stem_c2, t16, 16, Hh, W, ACT_RELU, sc32, 16, cur32, 16, cur16, 16, st);
    for (int l = 0; l < 5; ++l) {
        const std::vector<Block>& L = h->enc[l];
        for (size_t b = 0; b < L.size(); ++b) {
            const bool last = (b + 1 == L.size());
            // the level's last block writes its fp16 output straight into the decoder's concat buffer (channels [C, 2C))
            __half* o16 = last ? cat[l] + C : nxt16;
            const long ld16 = last ? 2 * C : C;
            const int cin = L[b].c1.cin;
            run_block(L[b], cur16, cin, cur32, Hh, W, t16, sc32, nxt32, o16, ld16, st);
            std::swap(cur32, nxt32);
            if (!last) std::swap(cur16, nxt16);
        }
        avgpool_kernel<<<(unsigned)ceil_div_l((long)(Hh / 2) * (W / 2) * C, 256), 256, 0, st>>>(cur32, Hh, W, C, cur16);
        KERNEL_CHECK();
        count_launch();
        Hh /= 2; W /= 2;
        if (l < 4) C *= 2;
    }
    // here: Hh = T/32, W = 4, cur16 holds the pooled [Hh, W, 256]
    for (int l = 0; l < 4; ++l)
        for (int b = 0; b < 4; ++b) {
            const Block& B = h->inter[l][b];
            run_block(B, cur16, B.c1.cin, cur32, Hh, W, t16, sc32, nxt32, nxt16, B.c2.cout, st);
            std::swap(cur32, nxt32);
            std::swap(cur16, nxt16);
        }
    for (int l = 0; l < 5; ++l) {
        const rvcb_rmvpe::Up& U = h->ups[l];
        const int lvl = 4 - l;                       // encoder level whose skip is concatenated
        {
            GemmArgs g;
            g.A = cur16; g.lda = U.cin; g.a_rows = Hh; g.a_cols = U.cin; g.conv2d_W = W;
            g.B = U.w.d; g.ldb = U.w.cols; g.b_rows = U.w.rows; g.b_cols = U.w.cols;
            g.M = Hh * W; g.N = 4 * U.cout; g.block_k = U.bk;
            g.nseg = 4;
            for (int si = 0; si < 4; ++si) g.seg[si] = {si >> 1, 0, si & 1, ceil_div(U.cin, U.bk)};
            g.bias = U.b; g.act1 = ACT_RELU;
            g.out16 = cat[lvl]; g.ld16 = 2 * U.cout; g.up2_C = U.cout;
            gemm(g, st);
        }
        Hh *= 2; W *= 2;
        const std::vector<Block>& L = h->dec[l];
        for (int b = 0; b < 4; ++b) {
            const __half* in16 = (b == 0) ? cat[lvl] : cur16;
            const long ldx = (b == 0) ? 2 * U.cout : U.cout;
            run_block(L[b], in16, ldx, cur32, Hh, W, t16, sc32, nxt32, nxt16, U.cout, st);
            std::swap(cur32, nxt32);
            std::swap(cur16, nxt16);
        }
    }
    // ---- head: cnn 16->3, BiGRU, Linear + sigmoid ----
    __half* gru_in = ar.alloc<__half>((size_t)T * 384);
    {
        const Conv2dRowArgs ra{cur16, 16, T, 128, 16, 3, h->cnn_w.d, h->cnn_w.rows, h->cnn_w.cols, h->cnn_b, false, nullptr, 0, nullptr, 0, gru_in, 3};
        GemmArgs g;
        g.A = cur16; g.lda = 16; g.a_rows = T; g.a_cols = 16; g.conv2d_W = 128;
        g.B = h->cnn_w.d; g.ldb = h->cnn_w.cols; g.b_rows = h->cnn_w.rows; g.b_cols = h->cnn_w.cols;
        g.M = T * 128; g.N = 3; g.block_k = 16;
        seg_conv2d_3x3(g, 16);
        g.bias = h->cnn_b; g.out16 = gru_in; g.ld16 = 3;
        if (!conv2d_row_try(ra, st)) gemm(g, st);
    }
    float* gi = ar.alloc<float>((size_t)T * 1536);
    {
        GemmArgs g;
        g.A = gru_in; g.lda = 384; g.a_rows = T; g.a_cols = 384;
        g.B = h->wih.d; g.ldb = h->wih.cols; g.b_rows = h->wih.rows; g.b_cols = h->wih.cols;
        g.M = T; g.N = 1536; seg_linear(g, 384);
        g.bias = h->bih; g.out32 = gi; g.ld32 = 1536;
        gemm(g, st);
    }
    __half* gru_out = ar.alloc<__half>((size_t)T * 512);
    {
        static const bool v1 = [] { const char* e = getenv("RVCB_GRU"); return e && !strcmp(e, "barrier"); }();
        static const bool fast = [] { const char* e = getenv("RVCB_GRU_FAST"); return !(e && e[0] == '0'); }();
        if (v1) gru_kernel<<<2 * GRU_CL, 256, 0, st>>>(gi, h->whh, h->bhh, T, nullptr, gru_out);
        else if (getenv("RVCB_GRU_TRACE") && fast) gru_kernel_v2<true, true><<<2 * GRU_CL, 256, 0, st>>>(gi, h->whh, h->bhh, T, nullptr, gru_out);
        else if (getenv("RVCB_GRU_TRACE")) gru_kernel_v2<false, true><<<2 * GRU_CL, 256, 0, st>>>(gi, h->whh, h->bhh, T, nullptr, gru_out);
        else if (fast) gru_kernel_v2<true><<<2 * GRU_CL, 256, 0, st>>>(gi, h->whh, h->bhh, T, nullptr, gru_out);
        else gru_kernel_v2<false><<<2 * GRU_CL, 256, 0, st>>>(gi, h->whh, h->bhh, T, nullptr, gru_out);
        KERNEL_CHECK();
        count_launch();
    }
    float* hidden = d_hidden ? d_hidden : ar.alloc<float>((size_t)nf * N_CLASS);
    {
        GemmArgs g;
        g.A = gru_out; g.lda = 512; g.a_rows = T; g.a_cols = 512;
        g.B = h->fc_w.d; g.ldb = h->fc_w.cols; g.b_rows = h->fc_w.rows; g.b_cols = h->fc_w.cols;
        g.M = nf; g.N = N_CLASS; seg_linear(g, 512);
        g.bias = h->fc_b; g.act1 = ACT_SIGMOID; g.out32 = hidden; g.ld32 = N_CLASS;
        gemm(g, st);
    }
    if (d_f0) {
        decode_kernel<<<ceil_div(nf, 8), 256, 0, st>>>(hidden, nf, thred, d_f0);
        KERNEL_CHECK();
        count_launch();
    }
    if (n_frames) *n_frames = nf;
}

extern "C" {

int rvcb_rmvpe_create(const rvcb_weights* w, rvcb_rmvpe** out) {
    RVCB_API_BEGIN
    RVCB_CHECK(w && out, "null argument");
    *out = rmvpe_build(*w);
    RVCB_API_END
}
int rvcb_rmvpe_num_frames(int n_samples) { return rmvpe_frames(n_samples); }
int rvcb_rmvpe_infer(rvcb_rmvpe* h, const float* d_wav, int n_samples, float thred, float* d_mel, float* d_hidden, float* d_f0,
                     int* n_frames, void* stream) {
    RVCB_API_BEGIN
    RVCB_CHECK(h && d_wav, "null argument");
    rmvpe_forward(h, d_wav, n_samples, thred, d_mel, d_hidden, d_f0, n_frames, (cudaStream_t)stream);
    RVCB_API_END
}
void rvcb_rmvpe_destroy(rvcb_rmvpe* h) { delete h; }

}  // extern "C"
