// SIMT kernels around the implicit-GEMM engine (norms, softmax, elementwise, scans).
#pragma once
#include "common.cuh"

namespace rvcb {

// y = LN(x) over the last dim C; x fp32 [rows, ldx]; writes fp32 and/or fp16 copies.
void layernorm_rows(const float* x, long ldx, int rows, int C, const float* gamma, const float* beta, float eps,
                    float* out32, long ld32, __half* out16, long ld16, cudaStream_t s);

// HuBERT conv0 (1 -> 512, k=10, s=5, no bias) + GroupNorm(512 groups) over time + GELU -> fp16 [T0(+pad), 512]
void hubert_conv0_gn_gelu(const float* wav, int n_samples, const float* w /*[512,10]*/, const float* gamma, const float* beta,
                          float* scratch_y /*[T0,512]*/, double* scratch_stats /*[2*512]*/, __half* out16, int T0, cudaStream_t s);

// row softmax of S fp32 [H, T, lds] -> P fp16 [H*T, ldp]; optional relative-position band:
//   S[h,i,j] += qrel[h,i,(j-i+win)] for |j-i|<=win ; prel[h*T+i, r] = P[i, i+r-win] (fp16, ld 64, zero padded)
void softmax_rows(const float* S, long lds, int H, int T, __half* P, long ldp, const float* qrel, long ldq, int win,
                  __half* prel, cudaStream_t s);

void cast_f32_f16(const float* x, __half* y, long n, cudaStream_t s);
void half_to_float(const __half* x, float* y, long n, cudaStream_t s);   // exact widening
// y[r, c] = x[r, c] for a [rows, cols] fp32 matrix with leading dims -> fp16
void cast_f32_f16_2d(const float* x, long ldx, __half* y, long ldy, int rows, int cols, cudaStream_t s);

// y[n] = sum_k W[n,k] x[k] + b[n] (+ add[n])    (tiny conditioning mat-vecs, fp32)
void matvec(const float* W, const float* x, const float* b, const float* add, float* y, int N, int K, cudaStream_t s);

// TextEncoder front: x = lrelu((lin[t,:] + emb_pitch[pitch[t],:]) * scale, 0.1)
void textenc_embed(const float* lin, const long long* pitch, const float* emb_pitch, int T, int C, float scale, float* out32,
                   __half* out16, cudaStream_t s);
// z = (m + exp(logs) * noise * 0.66666): stats fp32 [T, 2*C] (m | logs), noise channel-major [C, ldn]
void prior_sample(const float* stats, const float* noise, long ldn, int T, int C, float* z, cudaStream_t s);
// out[t, c] = in[t, C-1-c]; also emits fp16 of the first `half_c` output channels
void flip_channels(const float* in, float* out, __half* x0_16, int T, int C, int half_c, cudaStream_t s);
// add per-channel vector: y[t,c] = x[t,c] + v[c]  (fp32 in place allowed), optional lrelu->fp16 copy
void add_rowvec(float* x, const float* v, int T, int C, __half* out16, float lrelu_slope, cudaStream_t s);

// NSF source (generators.py:148-194 + nsf.py:57-61): f0 [T] -> har fp32 [T*upp]
void sine_source(const float* f0, int T, int upp, int sr, const float* noise, float lin_w, float lin_b, float* phase_scratch,
                 float* har, cudaStream_t s);
// out[t, m] = (half) har[t*step + m - pad]  (0 for m >= m_valid or outside the source): the harmonic-source columns appended
// to a vocoder stage input, so the noise convolution rides in the ups GEMM as one more K segment
void har_columns(const float* har, long n_har, __half* out, long ld, int T, int Mp, int m_valid, int step, int pad, cudaStream_t s);
// out[t] = tanh(sum_{j,c} x[t + j - k/2, c] * w[j, c])   (conv_post, one output channel; x fp16 [T, C] dense)
void conv_post_tanh(const __half* x, int T, int C, const float* w, int k, float* out, cudaStream_t s);
// linear interpolation along time (F.interpolate mode="linear", align_corners=False) of [T_in, C] -> [T_out, C]
void interp_linear_rows(const float* in, int T_in, float* out, int T_out, int C, cudaStream_t s);

// retrieval epilogue (pipeline.py:140-160)
void upsample_protect(const float* feats, const float* feats0, int T_h, int C, const float* pitchf, int T, float protect,
                      float* out, cudaStream_t s);

// RMS-envelope mix of the converted audio with the input's envelope + peak normalisation to the int16 range, in place on y
// (pipeline.py:26-45, 349-360).  scratch: >= (n1/8000 + n2/(sr2/2) + 8) doubles.
void post_mix(float* y, long n2, int sr2, const float* x16k, long n1, float rate, double* scratch, cudaStream_t s, bool scale = true);

// fp32 SIMT GEMM  C[M,N] = A[M,K] * B[N,K]^T  (A rows may overlap: lda < K is allowed)
void sgemm_nt(const float* A, long lda, const float* B, long ldb, float* C, long ldc, int M, int N, int K, cudaStream_t s);

// forward-backward IIR as a cascade of <= 4 second-order sections [b0 b1 b2 1 a1 a2] (odd extension by `edge` samples,
// edge state zi[sec] * first input sample), float64, block-parallel on one CTA
struct SosCoef { int ns; double sos[4][6]; double zi[4][2]; };
void sosfiltfilt(const SosCoef& c, const float* x, long n, int edge, float* y, double* scratch, cudaStream_t s);
// np.pad(mode="reflect") of a 1-D signal
void reflect_pad(const float* x, long n, long pad, float* out, cudaStream_t s);
void f32_to_i16(const float* x, long n, short* out, cudaStream_t s);
// resize (np.interp with unvoiced -> NaN -> 0) + gap fill + key shift + mel quantisation, float64, numpy operation order
void f0_post(const float* f0, int n_frames, int p_len, double key_factor, double f0_min, double f0_max, long long* pitch, float* pitchf,
             double* scratch, cudaStream_t s);

// Realtime tail of gui.py's audio callback on the device (gui.py:1024-1087): envelope mix (rms_mix_rate < 1, in place on
// `infer`) + SOLA offset search, cross-fade, output block and buffer update.  scratch: >= 2 * (n / zc + 1) + nsearch + 1 floats.
void rt_tail(float* infer, int n, const float* input, int zc, float rms_mix_rate, float* sola_buffer, int block_frame, int nbuf, int nsearch,
             float* out, float* scratch, int* offset, cudaStream_t s, bool use_pv = false);

}  // namespace rvcb
