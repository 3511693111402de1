// Spectral gate of the realtime block (reference: infer/modules/gui/torchgate.py:217-280, utils.py:5-40; used at gui.py:974-990 on the
// input and gui.py:1015-1023 on the output) and the windowed-sinc polyphase resampler around rtrvc.RVC.infer (gui.py:851-866:
// torchaudio.transforms.Resample, sinc_interp_hann).  Everything is fp32: the gate is a threshold test on dB values, so the DFT
// runs as an fp32 SIMT GEMM against a window-folded basis (the same scheme as RMVPE's mel front end, rmvpe.cu) rather than on
// the tensor cores.  Sizes are tiny (realtime: 1920-point frames, 21 frames of signal, ~270 frames of noise reference), the
// whole gate is 9 launches and sits inside the captured realtime-block graph.
//
//   pad -> DFT GEMM (frames overlap in place: lda = hop) -> dB / per-bin floor / noise statistics -> mask -> 2-D mask smoothing
//   fused with the multiply -> inverse-DFT GEMM (synthesis window folded in) -> overlap-add / window-envelope normalisation
#include <cmath>
#include <vector>

#include "common.cuh"
#include "kernels.cuh"
#include "gemm.cuh"
#include "../../include/rvcb200.h"
#include "api_macros.h"

namespace rvcb {
namespace {

constexpr float kDbEps = 2.220446049250313e-16f;         // torch.finfo(float64).eps added in float32 (utils.py:7)

__global__ void tg_pad_kernel(const float* __restrict__ x, long n, int half, float* __restrict__ xp) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n + 2L * half) return;
    xp[i] = (i >= half && i < half + n) ? x[i - half] : 0.f;                     // torch.stft(center=True, pad_mode="constant")
}

// Split-precision operands for the tensor-core DFTs (same scheme as RMVPE's mel front end, rmvpe.cu): v = hi + lo / 2048 with hi, lo
// fp16.  Forward: the zero-padded signal as rows of `hop` samples, rows [0, R) = hi, rows [R, 2R) = lo.
__global__ void tg_split_pad_kernel(const float* __restrict__ x, long n, int half, int R, int hop, __half* __restrict__ y) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)R * hop) return;
    const float v = (i >= half && i < half + n) ? x[i - half] : 0.f;
    const __half hi = __float2half_rn(v);
    y[i] = hi;
    y[(long)R * hop + i] = __float2half_rn((v - __half2float(hi)) * 2048.f);
}

// spec: [T, 2F] (re | im).  One block = 32 bins x 8 time lanes.  mode 0: val[f, t] = max(dB, max_t dB - 40) (amp_to_db);
// mode 1: val = |X| (non-stationary branch).  thresh (optional, mode 0): mean_t + n_std * std_t (unbiased) of the floored dB.
__global__ void __launch_bounds__(256) tg_db_kernel(const float* __restrict__ spec, int T, int F, int mode, float top_db, float n_std,
                                                    float* __restrict__ val, float* __restrict__ thresh) {
    __shared__ float s_max[8][33];
    __shared__ double s_a[8][33], s_b[8][33];
    const int fx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int f = blockIdx.x * 32 + fx;
    const bool ok = f < F;
    float mx = -INFINITY;
    if (ok)
        for (int t = ty; t < T; t += 8) {
            const float re = spec[(long)t * 2 * F + f], im = spec[(long)t * 2 * F + F + f];
            const float a = sqrtf(re * re + im * im);
            const float v = mode == 0 ? 20.f * log10f(a + kDbEps) : a;
            if (val) val[(long)f * T + t] = v;
            mx = fmaxf(mx, v);
        }
    if (mode != 0) return;
    s_max[ty][fx] = mx;
    __syncthreads();
    mx = s_max[0][fx];
#pragma unroll
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, s_max[i][fx]);
    const float floor_db = mx - top_db;
    double sum = 0.0;
    if (ok)
        for (int t = ty; t < T; t += 8) {
            float v;
            if (val) {
                v = fmaxf(val[(long)f * T + t], floor_db);
                val[(long)f * T + t] = v;
            } else {
                const float re = spec[(long)t * 2 * F + f], im = spec[(long)t * 2 * F + F + f];
                v = fmaxf(20.f * log10f(sqrtf(re * re + im * im) + kDbEps), floor_db);
            }
            sum += (double)v;
        }
    if (!thresh) return;
    s_a[ty][fx] = sum;
    __syncthreads();
    double tot = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += s_a[i][fx];
    const double mean = tot / T;
    double sq = 0.0;
    if (ok)
        for (int t = ty; t < T; t += 8) {
            float v;
            if (val) v = val[(long)f * T + t];
            else {
                const float re = spec[(long)t * 2 * F + f], im = spec[(long)t * 2 * F + F + f];
                v = fmaxf(20.f * log10f(sqrtf(re * re + im * im) + kDbEps), floor_db);
            }
            const double d = (double)v - mean;
            sq += d * d;
        }
    s_b[ty][fx] = sq;
    __syncthreads();
    if (ty == 0 && ok) {
        double q = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) q += s_b[i][fx];
        const float sd = (float)sqrt(q / (T > 1 ? T - 1 : 1));
        thresh[f] = (float)mean + sd * n_std;                                  // torchgate.py:171-174
    }
}

// mask[f, t] after "prop_decrease * (mask - 1) + 1" (torchgate.py:252).  Stationary: val > thresh[f].  Non-stationary: moving mean of
// |X| over k frames (conv1d padding="same": (k-1)/2 zeros on the left, the rest on the right), slowness ratio, temperature sigmoid.
__global__ void tg_mask_kernel(const float* __restrict__ val, const float* __restrict__ thresh, int T, int F, int nonstat, int k,
                               float n_thresh, float temp, float prop, float* __restrict__ mask) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)F * T) return;
    const int f = (int)(i / T), t = (int)(i % T);
    float m;
    if (!nonstat) m = val[i] > thresh[f] ? 1.f : 0.f;
    else {
        const int left = (k - 1) / 2;
        float s = 0.f;
        for (int j = 0; j < k; ++j) {
            const int tt = t - left + j;
            if (tt >= 0 && tt < T) s += val[(long)f * T + tt];
        }
        const float sm = s / (float)k;
        const float ratio = (val[i] - sm) / (sm + 1e-6f);
        m = 1.f / (1.f + expf(-(ratio - n_thresh) / temp));
    }
    mask[i] = prop * (m - 1.f) + 1.f;
}

// Y[t, f] = X[t, f] * (filt (*) mask)[f, t]: conv2d(padding="same") of the mask with the fr x fc smoothing filter, zero outside.
// y16 (optional, instead of y): the masked spectrum as split fp16 planes [2T, Kp] (rows [0, T) hi, [T, 2T) lo x 2048; columns
// [2F, Kp) stay zero) for the tensor-core inverse DFT.
__global__ void tg_smooth_apply_kernel(const float* __restrict__ spec, const float* __restrict__ mask, const float* __restrict__ filt,
                                       int fr, int fc, int T, int F, float* __restrict__ y, __half* __restrict__ y16, int Kp) {
    extern __shared__ float s_f[];
    for (int i = threadIdx.x; i < fr * fc; i += blockDim.x) s_f[i] = filt[i];
    __syncthreads();
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)F * T) return;
    const int t = (int)(i / F), f = (int)(i % F);
    float s;
    if (fr * fc == 0) s = mask[(long)f * T + t];
    else {
        s = 0.f;
        const int pr = (fr - 1) / 2, pc = (fc - 1) / 2;
        for (int a = 0; a < fr; ++a) {
            const int ff = f - pr + a;
            if (ff < 0 || ff >= F) continue;
            for (int b = 0; b < fc; ++b) {
                const int tt = t - pc + b;
                if (tt >= 0 && tt < T) s = fmaf(s_f[a * fc + b], mask[(long)ff * T + tt], s);
            }
        }
    }
    const float re = spec[(long)t * 2 * F + f] * s, im = spec[(long)t * 2 * F + F + f] * s;
    if (y16) {
        const __half rh = __float2half_rn(re), ih = __float2half_rn(im);
        y16[(long)t * Kp + f] = rh;
        y16[(long)t * Kp + F + f] = ih;
        y16[(long)(T + t) * Kp + f] = __float2half_rn((re - __half2float(rh)) * 2048.f);
        y16[(long)(T + t) * Kp + F + f] = __float2half_rn((im - __half2float(ih)) * 2048.f);
    } else {
        y[(long)t * 2 * F + f] = re;
        y[(long)t * 2 * F + F + f] = im;
    }
}

// torch.istft(center=True): overlap-add of the (already windowed) frames, divided by the overlap-added squared window, trimmed by n_fft/2.
__global__ void tg_ola_kernel(const float* __restrict__ frames, const float* __restrict__ w2, int T, int N, int hop, long n_out,
                              float* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    const long p = i + N / 2;
    long t_hi = p / hop;
    if (t_hi > T - 1) t_hi = T - 1;
    long t_lo = (p - N + hop) / hop;                       // smallest t with p - t*hop < N
    if (p - N + 1 <= 0) t_lo = 0;
    float s = 0.f, e = 0.f;
    for (long t = t_lo; t <= t_hi; ++t) {
        const long k = p - t * hop;
        if (k < 0 || k >= N) continue;
        s += frames[t * N + k];
        e += w2[k];
    }
    out[i] = e > 1e-11f ? s / e : s;
}

// torchaudio.functional.resample's strided convolution: out[i * up + p] = sum_k kern[p, k] * xpad[i * down + k], xpad = x with
// `width` zeros on the left and `width + down` on the right; kern is [up, kw] (computed on the host exactly as torchaudio does).
__global__ void resample_kernel(const float* __restrict__ x, long n, const float* __restrict__ kern, int up, int down, int kw, int width,
                                long n_out, float* __restrict__ out) {
    const long o = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_out) return;
    const long i = o / up;
    const int p = (int)(o % up);
    const float* kp = kern + (long)p * kw;
    const long base = i * down - width;
    float s = 0.f;
    for (int k = 0; k < kw; ++k) {
        const long j = base + k;
        if (j >= 0 && j < n) s = fmaf(kp[k], x[j], s);
    }
    out[o] = s;
}

}  // namespace
}  // namespace rvcb

using namespace rvcb;

struct rvcb_torchgate {
    int sr = 0, n_fft = 0, hop = 0, F = 0, nonstat = 0, n_move = 0, fr = 0, fc = 0;
    float prop = 1.f, n_std = 1.5f, n_thresh = 1.3f, temp = 0.1f;
    float *fwd = nullptr, *inv = nullptr, *w2 = nullptr, *filt = nullptr;
    // tensor-core DFTs (bk != 0): fwd16 [2F, 3N] = (hi | lo | hi / 2048) of the forward basis, inv16 [N, 3 Kp] the same of N x the
    // inverse basis (the 1 / N goes into the GEMM's alpha: the split halves stay in fp16's normal range), Kp = 2F rounded up to 64
    __half *fwd16 = nullptr, *inv16 = nullptr;
    int bk = 0, Kp = 0;
    Arena arena;
    ~rvcb_torchgate() {
        cudaFree(fwd); cudaFree(inv); cudaFree(w2); cudaFree(filt); cudaFree(fwd16); cudaFree(inv16);
    }
};

extern "C" {

int rvcb_torchgate_create(int sr, int n_fft, int hop, int nonstationary, float n_std_thresh_stationary, float n_thresh_nonstationary,
                          float temp_coeff_nonstationary, int n_movemean_nonstationary, float prop_decrease, const float* h_filter,
                          int filter_rows, int filter_cols, rvcb_torchgate** out) {
    RVCB_API_BEGIN
    RVCB_CHECK(out && n_fft >= 16 && (n_fft % 2) == 0 && hop >= 1 && hop <= n_fft, "bad STFT geometry");
    RVCB_CHECK((filter_rows == 0 && filter_cols == 0) || (h_filter && (filter_rows & 1) && (filter_cols & 1)), "smoothing filter must be odd x odd");
    RVCB_CHECK((long)filter_rows * filter_cols * 4 <= 40000, "smoothing filter too large");
    auto* h = new rvcb_torchgate();
    try {
        h->sr = sr; h->n_fft = n_fft; h->hop = hop; h->F = n_fft / 2 + 1; h->nonstat = nonstationary; h->n_move = n_movemean_nonstationary;
        h->prop = prop_decrease; h->n_std = n_std_thresh_stationary; h->n_thresh = n_thresh_nonstationary; h->temp = temp_coeff_nonstationary;
        h->fr = filter_rows; h->fc = filter_cols;
        const int N = n_fft, F = h->F;
        std::vector<float> w(N), w2(N), fwd((size_t)2 * F * N), inv((size_t)N * 2 * F);
        const double tp = 2.0 * M_PI / N;
        for (int n = 0; n < N; ++n) {
            w[n] = (float)(0.5 - 0.5 * std::cos(tp * n));                     // torch.hann_window(N) (periodic), float32
            w2[n] = w[n] * w[n];
        }
        for (int f = 0; f < F; ++f) {
            const double c = (f == 0 || f == N / 2) ? 1.0 : 2.0;             // Hermitian fold of the inverse real DFT
            for (int n = 0; n < N; ++n) {
                const long r = ((long)f * n) % N;                              // exact angle reduction
                const double co = std::cos(tp * r), si = std::sin(tp * r);
                fwd[(size_t)f * N + n] = (float)(co * w[n]);
                fwd[(size_t)(F + f) * N + n] = (float)(-si * w[n]);
                inv[(size_t)n * 2 * F + f] = (float)(c * co * w[n] / N);
                inv[(size_t)n * 2 * F + F + f] = (float)(-c * si * w[n] / N);
            }
        }
        h->fwd = dev_upload(fwd.data(), fwd.size());
        h->inv = dev_upload(inv.data(), inv.size());
        {
            static const bool fp32_only = [] { const char* e = getenv("RVCB_TG_FP32"); return e && e[0] == '1'; }();
            const int rem = N % hop;
            for (int bk : {32, 16})
                if (!h->bk && !fp32_only && hop % bk == 0 && rem % bk == 0 && hop % 8 == 0 && N / hop + (rem ? 1 : 0) <= 32) h->bk = bk;
        }
        if (h->bk) {
            h->Kp = round_up(2 * F, 64);
            std::vector<__half> f16((size_t)2 * F * 3 * N), i16((size_t)N * 3 * h->Kp, __float2half(0.f));
            auto split = [](float b, __half& hi, __half& lo, __half& hs) {
                hi = __float2half_rn(b);
                const float hf = __half2float(hi);
                lo = __float2half_rn(b - hf);
                hs = __float2half_rn(hf * (1.f / 2048.f));
            };
            for (int r = 0; r < 2 * F; ++r)
                for (int n = 0; n < N; ++n) {
                    __half hi, lo, hs;
                    split(fwd[(size_t)r * N + n], hi, lo, hs);
                    f16[(size_t)r * 3 * N + n] = hi; f16[(size_t)r * 3 * N + N + n] = lo; f16[(size_t)r * 3 * N + 2 * N + n] = hs;
                }
            for (int n = 0; n < N; ++n)
                for (int j = 0; j < 2 * F; ++j) {
                    __half hi, lo, hs;
                    split(inv[(size_t)n * 2 * F + j] * (float)N, hi, lo, hs);
                    __half* row = i16.data() + (size_t)n * 3 * h->Kp;
                    row[j] = hi; row[h->Kp + j] = lo; row[2 * h->Kp + j] = hs;
                }
            h->fwd16 = dev_upload(f16.data(), f16.size());
            h->inv16 = dev_upload(i16.data(), i16.size());
        }
        h->w2 = dev_upload(w2.data(), w2.size());
        if (filter_rows) h->filt = dev_upload(h_filter, (size_t)filter_rows * filter_cols);
    } catch (...) {
        delete h;
        throw;
    }
    *out = h;
    RVCB_API_END
}

int64_t rvcb_torchgate_out_len(const rvcb_torchgate* h, int64_t n) { return h ? (int64_t)h->hop * (n / h->hop) : -1; }

int rvcb_torchgate_apply(rvcb_torchgate* h, const float* d_x, int64_t n, const float* d_xn, int64_t n_noise, float* d_y, void* stream) {
    RVCB_API_BEGIN
    RVCB_CHECK(h && d_x && d_y && n >= h->hop, "null / short argument");
    RVCB_CHECK(!d_xn || n_noise >= h->hop, "noise reference shorter than one hop");
    cudaStream_t st = (cudaStream_t)stream;
    const int N = h->n_fft, F = h->F, hop = h->hop;
    const int T = (int)(n / hop) + 1, Tn = d_xn ? (int)(n_noise / hop) + 1 : 0;
    auto rnd = [](size_t b) { return (b + 1023) & ~size_t(1023); };
    size_t need = rnd((size_t)(n + N) * 4) + rnd((size_t)T * 2 * F * 4) * 2 + rnd((size_t)F * T * 4) * 2 + rnd((size_t)F * 4) + rnd((size_t)T * N * 4);
    if (d_xn) need += rnd((size_t)(n_noise + N) * 4) + rnd((size_t)Tn * 2 * F * 4);
    if (h->bk) need += rnd((size_t)(n + N + 3L * hop) * 4 + 128) + rnd((size_t)(n_noise + N + 3L * hop) * 4 + 128) + rnd((size_t)2 * T * h->Kp * 2);
    h->arena.reserve(need + 4096);
    h->arena.reset();
    float* xp = h->arena.alloc<float>(n + N);
    float* spec = h->arena.alloc<float>((size_t)T * 2 * F);
    float* yspec = h->arena.alloc<float>((size_t)T * 2 * F);
    float* val = h->arena.alloc<float>((size_t)F * T);
    float* mask = h->arena.alloc<float>((size_t)F * T);
    float* thresh = h->arena.alloc<float>(F);
    float* frames = h->arena.alloc<float>((size_t)T * N);
    // forward STFT: spec[t, :] = frame t x basis.  Tensor-core form: a frame is N / hop consecutive hop-sized rows of the padded signal
    auto stft = [&](const float* x, long len, int frames, float* pad32, float* out) {
        if (!h->bk) {
            tg_pad_kernel<<<(unsigned)ceil_div_l(len + N, 256), 256, 0, st>>>(x, len, N / 2, pad32);
            KERNEL_CHECK();
            count_launch();
            sgemm_nt(pad32, hop, h->fwd, N, out, 2 * F, frames, 2 * F, N, st);
            return;
        }
        const int R = (int)ceil_div_l(len + N, hop) + 1;
        __half* ws = h->arena.alloc<__half>((size_t)2 * R * hop + 64);
        tg_split_pad_kernel<<<(unsigned)ceil_div_l((long)R * hop, 256), 256, 0, st>>>(x, len, N / 2, R, hop, ws);
        KERNEL_CHECK();
        count_launch();
        GemmArgs g;
        g.A = ws; g.lda = hop; g.a_rows = 2 * R; g.a_cols = hop;
        g.B = h->fwd16; g.ldb = 3 * N; g.b_rows = 2 * F; g.b_cols = 3 * N;
        g.M = frames; g.N = 2 * F; g.block_k = h->bk;
        g.nseg = 0;
        const int q = N / hop, rem = N % hop;
        for (int part = 0; part < 3; ++part) {
            for (int j = 0; j < q; ++j) g.seg[g.nseg++] = {(part == 2 ? R : 0) + j, 0, 0, hop / h->bk};
            if (rem) g.seg[g.nseg++] = {(part == 2 ? R : 0) + q, 0, 0, rem / h->bk};
        }
        g.out32 = out; g.ld32 = 2 * F;
        gemm(g, st);
    };
    stft(d_x, n, T, xp, spec);
    const int fb = ceil_div(F, 32);
    if (h->nonstat) {
        tg_db_kernel<<<fb, 256, 0, st>>>(spec, T, F, 1, 40.f, h->n_std, val, nullptr);
        KERNEL_CHECK();
        count_launch();
    } else if (d_xn) {
        float* np = h->arena.alloc<float>(n_noise + N);
        float* nspec = h->arena.alloc<float>((size_t)Tn * 2 * F);
        stft(d_xn, n_noise, Tn, np, nspec);
        tg_db_kernel<<<fb, 256, 0, st>>>(nspec, Tn, F, 0, 40.f, h->n_std, nullptr, thresh);
        KERNEL_CHECK();
        tg_db_kernel<<<fb, 256, 0, st>>>(spec, T, F, 0, 40.f, h->n_std, val, nullptr);
        KERNEL_CHECK();
        count_launch(2);
    } else {
        tg_db_kernel<<<fb, 256, 0, st>>>(spec, T, F, 0, 40.f, h->n_std, val, thresh);
        KERNEL_CHECK();
        count_launch();
    }
    const long ft = (long)F * T;
    tg_mask_kernel<<<(unsigned)ceil_div_l(ft, 256), 256, 0, st>>>(val, thresh, T, F, h->nonstat, h->n_move, h->n_thresh, h->temp, h->prop, mask);
    KERNEL_CHECK();
    __half* y16 = nullptr;
    if (h->bk) {
        y16 = h->arena.alloc<__half>((size_t)2 * T * h->Kp);
        CUDA_CHECK(cudaMemsetAsync(y16, 0, (size_t)2 * T * h->Kp * sizeof(__half), st));       // the K padding columns
    }
    tg_smooth_apply_kernel<<<(unsigned)ceil_div_l(ft, 256), 256, (size_t)h->fr * h->fc * 4, st>>>(spec, mask, h->filt, h->fr, h->fc, T, F, yspec,
                                                                                                   y16, h->Kp);
    KERNEL_CHECK();
    count_launch(2);
    if (!h->bk) sgemm_nt(yspec, 2 * F, h->inv, 2 * F, frames, N, T, N, 2 * F, st);
    else {
        GemmArgs g;
        g.A = y16; g.lda = h->Kp; g.a_rows = 2 * T; g.a_cols = h->Kp;
        g.B = h->inv16; g.ldb = 3 * h->Kp; g.b_rows = N; g.b_cols = 3 * h->Kp;
        g.M = T; g.N = N; g.block_k = 64;
        g.nseg = 3;
        for (int part = 0; part < 3; ++part) g.seg[part] = {part == 2 ? T : 0, 0, 0, h->Kp / 64};
        g.alpha = 1.f / (float)N;
        g.out32 = frames; g.ld32 = N;
        gemm(g, st);
    }
    const long n_out = (long)hop * (T - 1);
    tg_ola_kernel<<<(unsigned)ceil_div_l(n_out, 256), 256, 0, st>>>(frames, h->w2, T, N, hop, n_out, d_y);
    KERNEL_CHECK();
    count_launch();
    RVCB_API_END
}

void rvcb_torchgate_destroy(rvcb_torchgate* h) { delete h; }

int rvcb_resample_sinc(const float* d_x, int64_t n, const float* d_kernel, int up, int down, int kernel_width, int width, float* d_out,
                       int64_t n_out, void* stream) {
    RVCB_API_BEGIN
    RVCB_CHECK(d_x && d_kernel && d_out && up >= 1 && down >= 1 && kernel_width >= 1 && n_out >= 0, "bad argument");
    if (n_out) {
        resample_kernel<<<(unsigned)ceil_div_l(n_out, 256), 256, 0, (cudaStream_t)stream>>>(d_x, n, d_kernel, up, down, kernel_width, width, n_out, d_out);
        KERNEL_CHECK();
        count_launch();
    }
    RVCB_API_END
}

}  // extern "C"
