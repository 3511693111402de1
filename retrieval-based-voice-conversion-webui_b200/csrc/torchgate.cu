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
__global__ void tg_smooth_apply_kernel(const float* __restrict__ spec, const float* __restrict__ mask, const float* __restrict__ filt,
                                       int fr, int fc, int T, int F, float* __restrict__ y) {
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
    y[(long)t * 2 * F + f] = spec[(long)t * 2 * F + f] * s;
    y[(long)t * 2 * F + F + f] = spec[(long)t * 2 * F + F + f] * s;
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
    Arena arena;
    ~rvcb_torchgate() {
        cudaFree(fwd); cudaFree(inv); cudaFree(w2); cudaFree(filt);
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
    h->arena.reserve(need + 4096);
    h->arena.reset();
    float* xp = h->arena.alloc<float>(n + N);
    float* spec = h->arena.alloc<float>((size_t)T * 2 * F);
    float* yspec = h->arena.alloc<float>((size_t)T * 2 * F);
    float* val = h->arena.alloc<float>((size_t)F * T);
    float* mask = h->arena.alloc<float>((size_t)F * T);
    float* thresh = h->arena.alloc<float>(F);
    float* frames = h->arena.alloc<float>((size_t)T * N);
    tg_pad_kernel<<<(unsigned)ceil_div_l(n + N, 256), 256, 0, st>>>(d_x, n, N / 2, xp);
    KERNEL_CHECK();
    count_launch();
    sgemm_nt(xp, hop, h->fwd, N, spec, 2 * F, T, 2 * F, N, st);
    const int fb = ceil_div(F, 32);
    if (h->nonstat) {
        tg_db_kernel<<<fb, 256, 0, st>>>(spec, T, F, 1, 40.f, h->n_std, val, nullptr);
        KERNEL_CHECK();
        count_launch();
    } else if (d_xn) {
        float* np = h->arena.alloc<float>(n_noise + N);
        float* nspec = h->arena.alloc<float>((size_t)Tn * 2 * F);
        tg_pad_kernel<<<(unsigned)ceil_div_l(n_noise + N, 256), 256, 0, st>>>(d_xn, n_noise, N / 2, np);
        KERNEL_CHECK();
        sgemm_nt(np, hop, h->fwd, N, nspec, 2 * F, Tn, 2 * F, N, st);
        tg_db_kernel<<<fb, 256, 0, st>>>(nspec, Tn, F, 0, 40.f, h->n_std, nullptr, thresh);
        KERNEL_CHECK();
        tg_db_kernel<<<fb, 256, 0, st>>>(spec, T, F, 0, 40.f, h->n_std, val, nullptr);
        KERNEL_CHECK();
        count_launch(3);
    } else {
        tg_db_kernel<<<fb, 256, 0, st>>>(spec, T, F, 0, 40.f, h->n_std, val, thresh);
        KERNEL_CHECK();
        count_launch();
    }
    const long ft = (long)F * T;
    tg_mask_kernel<<<(unsigned)ceil_div_l(ft, 256), 256, 0, st>>>(val, thresh, T, F, h->nonstat, h->n_move, h->n_thresh, h->temp, h->prop, mask);
    KERNEL_CHECK();
    tg_smooth_apply_kernel<<<(unsigned)ceil_div_l(ft, 256), 256, (size_t)h->fr * h->fc * 4, st>>>(spec, mask, h->filt, h->fr, h->fc, T, F, yspec);
    KERNEL_CHECK();
    count_launch(2);
    sgemm_nt(yspec, 2 * F, h->inv, 2 * F, frames, N, T, N, 2 * F, st);
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
