// SIMT kernels around the implicit-GEMM engine.  All fp32 math; no fast-math.
#include "kernels.cuh"

namespace rvcb {

// let a PDL-launched successor (a tensor-core GEMM) run its prologue while this kernel executes; the successor still waits
// for this grid's completion before touching global memory (griddepcontrol.wait)
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm: one warp per row
// ---------------------------------------------------------------------------------------------
__global__ void layernorm_kernel(const float* __restrict__ x, long ldx, int rows, int C, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, float* out32, long ld32, __half* out16, long ld16) {
    pdl_trigger();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (long)row * ldx;
    float v[32];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int c = lane + i * 32;
        v[i] = c < C ? xr[c] : 0.f;
        s += v[i];
    }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int c = lane + i * 32;
        const float d = c < C ? v[i] - mean : 0.f;
        q += d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int c = lane + i * 32;
        if (c < C) {
            const float y = (v[i] - mean) * rstd * gamma[c] + beta[c];
            if (out32) out32[(long)row * ld32 + c] = y;
            if (out16) out16[(long)row * ld16 + c] = __float2half_rn(y);
        }
    }
}

void layernorm_rows(const float* x, long ldx, int rows, int C, const float* gamma, const float* beta, float eps, float* out32,
                    long ld32, __half* out16, long ld16, cudaStream_t s) {
    RVCB_CHECK(C <= 1024, "layernorm: C too large");
    layernorm_kernel<<<ceil_div(rows, 8), 256, 0, s>>>(x, ldx, rows, C, gamma, beta, eps, out32, ld32, out16, ld16);
    KERNEL_CHECK();
    count_launch();
}

// ---------------------------------------------------------------------------------------------
// HuBERT conv0 + GroupNorm(512 groups of 1 channel, over time) + GELU
// ---------------------------------------------------------------------------------------------
constexpr int C0_TSTEP = 32;
__global__ void hubert_conv0_kernel(const float* __restrict__ wav, int n_samples, const float* __restrict__ w, float* __restrict__ y,
                                    double* __restrict__ stats, int T0) {
    __shared__ float xs[C0_TSTEP * 5 + 8];
    const int c = threadIdx.x;                 // 512 threads = channels
    const int t0 = blockIdx.x * C0_TSTEP;
    for (int i = threadIdx.x; i < C0_TSTEP * 5 + 5; i += blockDim.x) {
        const long g = (long)t0 * 5 + i;
        xs[i] = g < n_samples ? wav[g] : 0.f;
    }
    float wr[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) wr[j] = w[c * 10 + j];
    __syncthreads();
    float s1 = 0.f, s2 = 0.f;
    for (int tt = 0; tt < C0_TSTEP; ++tt) {
        const int t = t0 + tt;
        if (t >= T0) break;
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 10; ++j) acc = fmaf(wr[j], xs[tt * 5 + j], acc);
        y[(long)t * 512 + c] = acc;
        s1 += acc;
        s2 += acc * acc;
    }
    atomicAdd(&stats[c], (double)s1);
    atomicAdd(&stats[512 + c], (double)s2);
}

// thread = channel (512), block = GN_TSTEP time steps: the per-channel mean / rstd (float64 divisions and a square root) are computed
// once per thread instead of once per element (the per-element form ran at 1.2 TB/s: 135 us for 157 MB); same arithmetic per element
constexpr int GN_TSTEP = 32;
__global__ void __launch_bounds__(512) hubert_gn_gelu_kernel(const float* __restrict__ y, const double* __restrict__ stats,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             __half* __restrict__ out, int T0) {
    pdl_trigger();
    const int c = threadIdx.x;
    const double mean = stats[c] / T0;
    const double var = stats[512 + c] / T0 - mean * mean;
    const float rstd = (float)(1.0 / sqrt(var + 1e-5));
    const float m = (float)mean, g = gamma[c], b = beta[c];
    const int t0 = blockIdx.x * GN_TSTEP;
#pragma unroll 4
    for (int tt = 0; tt < GN_TSTEP; ++tt) {
        const int t = t0 + tt;
        if (t >= T0) break;
        const long idx = (long)t * 512 + c;
        const float v = (y[idx] - m) * rstd * g + b;
        out[idx] = __float2half_rn(0.5f * v * (1.f + erff(v * 0.70710678118654752440f)));
    }
}

void hubert_conv0_gn_gelu(const float* wav, int n_samples, const float* w, const float* gamma, const float* beta, float* scratch_y,
                          double* scratch_stats, __half* out16, int T0, cudaStream_t s) {
    CUDA_CHECK(cudaMemsetAsync(scratch_stats, 0, sizeof(double) * 1024, s));
    hubert_conv0_kernel<<<ceil_div(T0, C0_TSTEP), 512, 0, s>>>(wav, n_samples, w, scratch_y, scratch_stats, T0);
    KERNEL_CHECK();
    hubert_gn_gelu_kernel<<<ceil_div(T0, GN_TSTEP), 512, 0, s>>>(scratch_y, scratch_stats, gamma, beta, out16, T0);
    KERNEL_CHECK();
    count_launch(2);
}

// ---------------------------------------------------------------------------------------------
// row softmax (+ relative-position band)
// ---------------------------------------------------------------------------------------------
__global__ void softmax_kernel(const float* __restrict__ S, long lds, int T, __half* __restrict__ P, long ldp,
                               const float* __restrict__ qrel, long ldq, int win, __half* __restrict__ prel) {
    pdl_trigger();
    extern __shared__ float row[];
    __shared__ float red[32];
    const long r = blockIdx.x;                  // h*T + i
    const int i = (int)(r % T);
    const float* sr = S + r * lds;
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < T; j += blockDim.x) {
        float v = sr[j];
        if (qrel) {
            const int d = j - i;
            if (d >= -win && d <= win) v += qrel[r * ldq + d + win];
        }
        row[j] = v;
        mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -INFINITY;
        v = warp_max(v);
        if (threadIdx.x == 0) red[0] = v;
    }
    __syncthreads();
    mx = red[0];
    __syncthreads();
    float sum = 0.f;
    for (int j = threadIdx.x; j < T; j += blockDim.x) {
        const float e = expf(row[j] - mx);
        row[j] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) red[0] = v;
    }
    __syncthreads();
    const float inv = 1.f / red[0];
    __half* pr = P + r * ldp;
    for (int j = threadIdx.x; j < ldp; j += blockDim.x) pr[j] = __float2half_rn(j < T ? row[j] * inv : 0.f);
    if (prel) {
        for (int q = threadIdx.x; q < 64; q += blockDim.x) {
            const int j = i + q - win;
            prel[r * 64 + q] = __float2half_rn((q <= 2 * win && j >= 0 && j < T) ? row[j] * inv : 0.f);
        }
    }
}

void softmax_rows(const float* S, long lds, int H, int T, __half* P, long ldp, const float* qrel, long ldq, int win, __half* prel,
                  cudaStream_t s) {
    RVCB_CHECK((size_t)T * 4 <= 48 * 1024, "softmax: row too long");
    softmax_kernel<<<H * T, 256, T * sizeof(float), s>>>(S, lds, T, P, ldp, qrel, ldq, win, prel);
    KERNEL_CHECK();
    count_launch();
}

// ---------------------------------------------------------------------------------------------
// small elementwise kernels
// ---------------------------------------------------------------------------------------------
__global__ void cast_kernel(const float* __restrict__ x, __half* __restrict__ y, long n) {
    pdl_trigger();
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = __float2half_rn(x[i]);
}
void cast_f32_f16(const float* x, __half* y, long n, cudaStream_t s) {
    cast_kernel<<<(unsigned)ceil_div_l(n, 256), 256, 0, s>>>(x, y, n);
    KERNEL_CHECK();
    count_launch();
}
__global__ void half_to_float_kernel(const __half* __restrict__ x, float* __restrict__ y, long n) {
    pdl_trigger();
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = __half2float(x[i]);
}
void half_to_float(const __half* x, float* y, long n, cudaStream_t s) {
    half_to_float_kernel<<<(unsigned)ceil_div_l(n, 256), 256, 0, s>>>(x, y, n);
    KERNEL_CHECK();
    count_launch();
}
__global__ void cast2d_kernel(const float* __restrict__ x, long ldx, __half* __restrict__ y, long ldy, int rows, int cols) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)rows * cols) return;
    const long r = i / cols;
    const int c = (int)(i - r * cols);
    y[r * ldy + c] = __float2half_rn(x[r * ldx + c]);
}
void cast_f32_f16_2d(const float* x, long ldx, __half* y, long ldy, int rows, int cols, cudaStream_t s) {
    cast2d_kernel<<<(unsigned)ceil_div_l((long)rows * cols, 256), 256, 0, s>>>(x, ldx, y, ldy, rows, cols);
    KERNEL_CHECK();
    count_launch();
}

__global__ void matvec_kernel(const float* __restrict__ W, const float* __restrict__ x, const float* __restrict__ b,
                              const float* __restrict__ add, float* __restrict__ y, int N, int K) {
    pdl_trigger();
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(W[(long)n * K + k], x[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) y[n] = acc + (b ? b[n] : 0.f) + (add ? add[n] : 0.f);
}
void matvec(const float* W, const float* x, const float* b, const float* add, float* y, int N, int K, cudaStream_t s) {
    matvec_kernel<<<ceil_div(N, 8), 256, 0, s>>>(W, x, b, add, y, N, K);
    KERNEL_CHECK();
    count_launch();
}

__global__ void textenc_embed_kernel(const float* __restrict__ lin, const long long* __restrict__ pitch, const float* __restrict__ emb,
                                     int T, int C, float scale, float* __restrict__ o32, __half* __restrict__ o16) {
    pdl_trigger();
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)T * C) return;
    const int t = (int)(i / C), c = (int)(i - (long)t * C);
    float v = lin[i];
    if (pitch) v += emb[(long)pitch[t] * C + c];
    v *= scale;
    v = v > 0.f ? v : v * 0.1f;
    o32[i] = v;
    o16[i] = __float2half_rn(v);
}
void textenc_embed(const float* lin, const long long* pitch, const float* emb_pitch, int T, int C, float scale, float* out32,
                   __half* out16, cudaStream_t s) {
    textenc_embed_kernel<<<(unsigned)ceil_div_l((long)T * C, 256), 256, 0, s>>>(lin, pitch, emb_pitch, T, C, scale, out32, out16);
    KERNEL_CHECK();
    count_launch();
}

__global__ void prior_kernel(const float* __restrict__ stats, const float* __restrict__ noise, long ldn, int T, int C, float* __restrict__ z) {
    pdl_trigger();
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)T * C) return;
    const int t = (int)(i / C), c = (int)(i - (long)t * C);
    const float m = stats[(long)t * 2 * C + c], logs = stats[(long)t * 2 * C + C + c];
    z[i] = m + expf(logs) * noise[(long)c * ldn + t] * 0.66666f;
}
void prior_sample(const float* stats, const float* noise, long ldn, int T, int C, float* z, cudaStream_t s) {
    prior_kernel<<<(unsigned)ceil_div_l((long)T * C, 256), 256, 0, s>>>(stats, noise, ldn, T, C, z);
    KERNEL_CHECK();
    count_launch();
}

__global__ void flip_kernel(const float* __restrict__ in, float* __restrict__ out, __half* __restrict__ x0, int T, int C, int hc) {
    pdl_trigger();
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)T * C) return;
    const int t = (int)(i / C), c = (int)(i - (long)t * C);
    const float v = in[(long)t * C + (C - 1 - c)];
    out[i] = v;
    if (c < hc) x0[(long)t * hc + c] = __float2half_rn(v);
}
void flip_channels(const float* in, float* out, __half* x0_16, int T, int C, int half_c, cudaStream_t s) {
    flip_kernel<<<(unsigned)ceil_div_l((long)T * C, 256), 256, 0, s>>>(in, out, x0_16, T, C, half_c);
    KERNEL_CHECK();
    count_launch();
}

__global__ void add_rowvec_kernel(float* __restrict__ x, const float* __restrict__ v, int T, int C, __half* __restrict__ o16, float slope) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)T * C) return;
    const int c = (int)(i % C);
    const float y = x[i] + v[c];
    x[i] = y;
    if (o16) o16[i] = __float2half_rn(y > 0.f ? y : y * slope);
}
void add_rowvec(float* x, const float* v, int T, int C, __half* out16, float lrelu_slope, cudaStream_t s) {
    add_rowvec_kernel<<<(unsigned)ceil_div_l((long)T * C, 256), 256, 0, s>>>(x, v, T, C, out16, lrelu_slope);
    KERNEL_CHECK();
    count_launch();
}

// ---------------------------------------------------------------------------------------------
// NSF sine source
// ---------------------------------------------------------------------------------------------
// phase[t] = fmod(cumsum_{s<t} r2[s], 1): block-wide scan in double (torch's CPU cumsum accumulates float in double,
// acc_type<float,false>; double partial sums of <= 2^13 floats in (-0.5, 0.5] round to the same float in any order)
__global__ void __launch_bounds__(1024) sine_phase_kernel(const float* __restrict__ f0, int T, int upp, float sr, float* __restrict__ phase) {
    __shared__ double part[1024];
    const int tid = threadIdx.x;
    const int per = (T + 1023) / 1024;
    const int s0 = tid * per;
    const float a_last = (float)upp;
    double loc = 0.0;
    for (int i = 0; i < per; ++i) {
        const int t = s0 + i;
        if (t + 1 < T) {
            const float rad = __fmul_rn(__fdiv_rn(f0[t], sr), a_last);
            loc += (double)(fmodf(__fadd_rn(rad, 0.5f), 1.0f) - 0.5f);
        }
    }
    part[tid] = loc;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {          // inclusive Hillis-Steele scan
        double v = tid >= off ? part[tid - off] : 0.0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    double acc = tid > 0 ? part[tid - 1] : 0.0;
    if (tid == 0) phase[0] = 0.f;
    for (int i = 0; i < per; ++i) {
        const int t = s0 + i;
        if (t + 1 < T) {
            const float rad = __fmul_rn(__fdiv_rn(f0[t], sr), a_last);
            acc += (double)(fmodf(__fadd_rn(rad, 0.5f), 1.0f) - 0.5f);
            phase[t + 1] = fmodf((float)acc, 1.0f);
        }
    }
}
__global__ void sine_wave_kernel(const float* __restrict__ f0, const float* __restrict__ phase, int T, int upp, float sr,
                                 const float* __restrict__ noise, float lw, float lb, float* __restrict__ har) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)T * upp) return;
    const int t = (int)(i / upp), j = (int)(i - (long)t * upp);
    const float f = f0[t];
    float rad = __fmul_rn(__fdiv_rn(f, sr), (float)(j + 1));
    rad = __fadd_rn(rad, phase[t]);
    const float sine = __fmul_rn(sinf(__fmul_rn(6.283185307179586f, rad)), 0.1f);
    const float uv = f > 0.f ? 1.f : 0.f;
    const float namp = __fadd_rn(__fmul_rn(uv, 0.003f), __fdiv_rn(__fmul_rn(1.f - uv, 0.1f), 3.f));
    const float sw = __fadd_rn(__fmul_rn(sine, uv), __fmul_rn(namp, noise[i]));
    har[i] = tanhf(__fadd_rn(__fmul_rn(sw, lw), lb));
}
void sine_source(const float* f0, int T, int upp, int sr, const float* noise, float lin_w, float lin_b, float* phase_scratch,
                 float* har, cudaStream_t s) {
    sine_phase_kernel<<<1, 1024, 0, s>>>(f0, T, upp, (float)sr, phase_scratch);
    KERNEL_CHECK();
    const long n = (long)T * upp;
    sine_wave_kernel<<<(unsigned)ceil_div_l(n, 256), 256, 0, s>>>(f0, phase_scratch, T, upp, (float)sr, noise, lin_w, lin_b, har);
    KERNEL_CHECK();
    count_launch(2);
}

// Harmonic-source columns of a vocoder stage input: out[t, m] = (half) har[t * step + m - pad] for m < m_valid and an
// in-range source index, else 0.  These columns are the A operand of the noise convolution folded into the ups GEMM.
__global__ void har_columns_kernel(const float* __restrict__ har, long n_har, __half* __restrict__ out, long ld, int T, int Mp, int m_valid,
                                   int step, int pad) {
    pdl_trigger();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;      // T * Mp / 2 < 2^31 (pairs of columns)
    const int half_m = Mp >> 1;
    if (i >= T * half_m) return;
    const int t = i / half_m;
    const int m = (i - t * half_m) * 2;
    const long q = (long)t * step + m - pad;
    const float a = (m < m_valid && q >= 0 && q < n_har) ? har[q] : 0.f;
    const float b = (m + 1 < m_valid && q + 1 >= 0 && q + 1 < n_har) ? har[q + 1] : 0.f;
    *reinterpret_cast<__half2*>(out + (long)t * ld + m) = __floats2half2_rn(a, b);
}
void har_columns(const float* har, long n_har, __half* out, long ld, int T, int Mp, int m_valid, int step, int pad, cudaStream_t s) {
    RVCB_CHECK(Mp % 2 == 0 && ld % 2 == 0 && (long)T * (Mp / 2) < (1L << 31), "har_columns: bad shape");
    har_columns_kernel<<<(unsigned)ceil_div_l((long)T * (Mp / 2), 256), 256, 0, s>>>(har, n_har, out, ld, T, Mp, m_valid, step, pad);
    KERNEL_CHECK();
    count_launch();
}

// conv_post (Conv1d(C, 1, k, padding k/2, bias=False) + tanh, nsf.py:187-189): one output channel is a bandwidth-bound
// reduction, not a GEMM.  A block stages (256 + k - 1) input rows in shared memory (row stride C*2 + 16 B: conflict-free
// 16-byte reads at one row per thread) and each thread reduces its k x C window in fp32.
template <int C>
__global__ void __launch_bounds__(256) conv_post_kernel(const __half* __restrict__ x, int T, const float* __restrict__ w /*[k, C]*/, int k,
                                                        float* __restrict__ out) {
    pdl_trigger();
    constexpr int RS = C * 2 + 16;               // bytes per staged row
    constexpr int V = C / 8;                     // uint4 per row
    extern __shared__ __align__(16) unsigned char smem_cp[];
    float* sw = reinterpret_cast<float*>(smem_cp);
    unsigned char* sx = smem_cp + ((k * C * 4 + 15) & ~15);
    const int pad = k / 2;
    const int t0 = blockIdx.x * 256;
    const int rows = 256 + k - 1;
    for (int i = threadIdx.x; i < k * C; i += 256) sw[i] = w[i];
    for (int i = threadIdx.x; i < rows * V; i += 256) {
        const int r = i / V, v = i - r * V;
        const int t = t0 + r - pad;
        uint4 val = make_uint4(0u, 0u, 0u, 0u);
        if (t >= 0 && t < T) val = *reinterpret_cast<const uint4*>(x + (long)t * C + v * 8);
        *reinterpret_cast<uint4*>(sx + r * RS + v * 16) = val;
    }
    __syncthreads();
    const int t = t0 + threadIdx.x;
    if (t >= T) return;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < k; ++j) {
        const unsigned char* row = sx + (threadIdx.x + j) * RS;
        const float* wj = sw + j * C;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const uint4 q = *reinterpret_cast<const uint4*>(row + v * 16);
            const __half2* h2 = reinterpret_cast<const __half2*>(&q);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(h2[e]);
                acc[e] = fmaf(f.x, wj[v * 8 + 2 * e], acc[e]);
                acc[e] = fmaf(f.y, wj[v * 8 + 2 * e + 1], acc[e]);
            }
        }
    }
    out[t] = tanhf((acc[0] + acc[1]) + (acc[2] + acc[3]));
}
void conv_post_tanh(const __half* x, int T, int C, const float* w, int k, float* out, cudaStream_t s) {
    RVCB_CHECK((C == 32 || C == 16 || C == 64) && k >= 1 && k <= 15, "conv_post: unsupported shape");
    const size_t smem = ((size_t)k * C * 4 + 15) / 16 * 16 + (size_t)(256 + k - 1) * (C * 2 + 16);
    const unsigned grid = (unsigned)ceil_div_l(T, 256);
    if (C == 32) conv_post_kernel<32><<<grid, 256, smem, s>>>(x, T, w, k, out);
    else if (C == 16) conv_post_kernel<16><<<grid, 256, smem, s>>>(x, T, w, k, out);
    else conv_post_kernel<64><<<grid, 256, smem, s>>>(x, T, w, k, out);
    KERNEL_CHECK();
    count_launch();
}

__global__ void interp_rows_kernel(const float* __restrict__ in, int T_in, float* __restrict__ out, int T_out, int C) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)T_out * C) return;
    const int t = (int)(i / C), c = (int)(i - (long)t * C);
    const float scale = (float)T_in / (float)T_out;
    float src = ((float)t + 0.5f) * scale - 0.5f;
    if (src < 0.f) src = 0.f;
    int i0 = (int)src;
    if (i0 > T_in - 1) i0 = T_in - 1;
    const int i1 = i0 + 1 < T_in ? i0 + 1 : i0;
    const float l1 = src - (float)i0, l0 = 1.f - l1;
    out[i] = l0 * in[(long)i0 * C + c] + l1 * in[(long)i1 * C + c];
}
void interp_linear_rows(const float* in, int T_in, float* out, int T_out, int C, cudaStream_t s) {
    interp_rows_kernel<<<(unsigned)ceil_div_l((long)T_out * C, 256), 256, 0, s>>>(in, T_in, out, T_out, C);
    KERNEL_CHECK();
    count_launch();
}

__global__ void upsample_protect_kernel(const float* __restrict__ f, const float* __restrict__ f0, int T_h, int C,
                                        const float* __restrict__ pitchf, int T, float protect, float* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)T * C) return;
    const int t = (int)(i / C), c = (int)(i - (long)t * C);
    const int th = t >> 1;
    float v = f[(long)th * C + c];
    if (f0 != nullptr && pitchf != nullptr && protect < 0.5f) {
        const float pf = pitchf[t];
        float pm = pf > 0.f ? 1.f : pf;       // pitchff[pitchf > 0] = 1
        if (pf < 1.f) pm = protect;           // pitchff[pitchf < 1] = protect
        v = v * pm + f0[(long)th * C + c] * (1.f - pm);
    }
    out[i] = v;
}
void upsample_protect(const float* feats, const float* feats0, int T_h, int C, const float* pitchf, int T, float protect, float* out,
                      cudaStream_t s) {
    RVCB_CHECK(T <= 2 * T_h, "upsample_protect: T > 2*T_h");
    upsample_protect_kernel<<<(unsigned)ceil_div_l((long)T * C, 256), 256, 0, s>>>(feats, feats0, T_h, C, pitchf, T, protect, out);
    KERNEL_CHECK();
    count_launch();
}

// ---------------------------------------------------------------------------------------------
// host-DSP epilogue on the device: RMS-envelope mix (change_rms, pipeline.py:26-45) + peak normalisation to the int16
// range (pipeline.py:356-360).  rms frames = librosa.feature.rms(frame = 2*hop, center=True, zero padding).
// ---------------------------------------------------------------------------------------------
__global__ void block_sumsq_kernel(const float* __restrict__ x, long n, int hop, int nblocks, double* __restrict__ sums) {
    // block b covers padded samples [b*hop, (b+1)*hop) of yp = pad(x, hop, hop)  ->  x[(b-1)*hop ... b*hop)
    const int b = blockIdx.x;
    double acc = 0.0;
    const long base = (long)(b - 1) * hop;
    for (int i = threadIdx.x; i < hop; i += blockDim.x) {
        const long j = base + i;
        if (j >= 0 && j < n) {
            const float v = x[j];
            acc += (double)v * (double)v;
        }
    }
    __shared__ double red[256];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[b] = red[0];
}

__device__ __forceinline__ float rms_interp(const double* __restrict__ sums, int nf, int frame_len, long n_out, long i) {
    // F.interpolate(mode="linear", align_corners=False) of the [nf] rms track to n_out samples
    const float scale = (float)nf / (float)n_out;
    float src = scale * ((float)i + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    int i0 = (int)src;
    if (i0 > nf - 1) i0 = nf - 1;
    const int i1 = i0 + 1 < nf ? i0 + 1 : i0;
    const float l1 = src - (float)i0, l0 = 1.f - l1;
    const float r0 = (float)sqrt((sums[i0] + sums[i0 + 1]) / frame_len);
    const float r1 = (float)sqrt((sums[i1] + sums[i1 + 1]) / frame_len);
    return l0 * r0 + l1 * r1;
}

__global__ void rms_mix_kernel(float* __restrict__ y, long n2, const double* __restrict__ s1, int nf1, int fl1, const double* __restrict__ s2,
                               int nf2, int fl2, float rate, int do_mix, unsigned int* __restrict__ amax_bits) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    float v = 0.f;
    if (i < n2) {
        v = y[i];
        if (do_mix) {
            const float r1 = rms_interp(s1, nf1, fl1, n2, i);
            const float r2 = fmaxf(rms_interp(s2, nf2, fl2, n2, i), 1e-6f);
            v *= powf(r1, 1.f - rate) * powf(r2, rate - 1.f);
            y[i] = v;
        }
    }
    float a = fabsf(v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, o));
    if ((threadIdx.x & 31) == 0 && a > 0.f) atomicMax(amax_bits, __float_as_uint(a));
}

__global__ void peak_scale_kernel(float* __restrict__ y, long n, const unsigned int* __restrict__ amax_bits) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float amax = __uint_as_float(*amax_bits) / 0.99f;
    float sc = 32768.f;
    if (amax > 1.f) sc /= amax;
    y[i] *= sc;
}

void post_mix(float* y, long n2, int sr2, const float* x16k, long n1, float rate, double* scratch, cudaStream_t s, bool scale) {
    const int hop1 = 16000 / 2, hop2 = sr2 / 2;
    const int nf1 = 1 + (int)(n1 / hop1), nf2 = 1 + (int)(n2 / hop2);         // 1 + (len + 2*hop - 2*hop)/hop
    double* s1 = scratch;
    double* s2 = scratch + (nf1 + 2);
    unsigned int* amax = reinterpret_cast<unsigned int*>(s2 + (nf2 + 2));
    CUDA_CHECK(cudaMemsetAsync(amax, 0, sizeof(unsigned int), s));
    const int do_mix = rate != 1.f;
    if (do_mix) {
        block_sumsq_kernel<<<nf1 + 1, 256, 0, s>>>(x16k, n1, hop1, nf1 + 1, s1);
        block_sumsq_kernel<<<nf2 + 1, 256, 0, s>>>(y, n2, hop2, nf2 + 1, s2);
        KERNEL_CHECK();
        count_launch(2);
    }
    rms_mix_kernel<<<(unsigned)ceil_div_l(n2, 256), 256, 0, s>>>(y, n2, s1, nf1, 2 * hop1, s2, nf2, 2 * hop2, rate, do_mix, amax);
    if (scale) peak_scale_kernel<<<(unsigned)ceil_div_l(n2, 256), 256, 0, s>>>(y, n2, amax);
    KERNEL_CHECK();
    count_launch(scale ? 2 : 1);
}

// ---------------------------------------------------------------------------------------------
// fp32 SIMT GEMM, 64x64x16 tiles, 4x4 register blocking
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sgemm_nt_kernel(const float* __restrict__ A, long lda, const float* __restrict__ B, long ldb,
                                                       float* __restrict__ C, long ldc, int M, int N, int K) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float acc[4][4] = {};
    const int lk = threadIdx.x & 15, lr = threadIdx.x >> 4;     // loader: k index, row (0..15)
    for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = lr + 16 * i;
            const int k = k0 + lk;
            As[lk][r] = (m0 + r < M && k < K) ? A[(long)(m0 + r) * lda + k] : 0.f;
            Bs[lk][r] = (n0 + r < N && k < K) ? B[(long)(n0 + r) * ldb + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a[i] = As[k][ty * 4 + i];
                b[i] = Bs[k][tx * 4 + i];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
            if (m < M && n < N) C[(long)m * ldc + n] = acc[i][j];
        }
}
void sgemm_nt(const float* A, long lda, const float* B, long ldb, float* C, long ldc, int M, int N, int K, cudaStream_t s) {
    dim3 grid(ceil_div(N, 64), ceil_div(M, 64));
    sgemm_nt_kernel<<<grid, 256, 0, s>>>(A, lda, B, ldb, C, ldc, M, N, K);
    KERNEL_CHECK();
    count_launch();
}

// ------------------------------------------------------------------------------------------------
// input / output plumbing kernels
// ------------------------------------------------------------------------------------------------
// Forward-backward IIR (scipy filtfilt semantics: odd extension by `edge` samples, steady-state edge conditions) as a cascade
// of second-order sections, float64.  The 5th-order direct form of the reference has poles clustered at |z| = 0.98..0.994; its
// 5x5 state matrix is too non-normal to propagate across blocks in float64 (A^L comes out ~1e13 instead of ~1), the 2x2 section
// matrices are benign (|A^L| <= ~30).  The recurrence is cut into IIR_NB blocks of L samples, one thread each, spread over
// many SMs (FP64 issue rate per SM is what bounds a one-CTA version: 1.1 ms measured).  Per direction and section:
//   zero-state end state of every block                         (sos_sweep_kernel, fused with the previous section's re-run)
//   start state of every block: s_k = sum_{j<=k} M^(k-j) u_j     (sos_scan_kernel: log-step prefix over IIR_NB blocks, M = A^L)
//   exact re-run of every block from its start state, in place   (sos_sweep_kernel)
// DF2T section: y = z0 + b0 x;  z0' = z1 + b1 x - a1 y;  z1' = b2 x - a2 y;   A = [[-a1, 1], [-a2, 0]].
constexpr int IIR_NB = 1024, IIR_B = 16, IIR_LOG = 10;
struct SosSec { double q[6]; };
struct SosScanP { double zi[2]; double mp[IIR_LOG][4]; };      // edge state per unit input, and M^(2^d), d = 0..9
__device__ __forceinline__ double sos_step(double& z0, double& z1, const double* c, double x) {
    const double y = z0 + c[0] * x;
    z0 = z1 + c[1] * x - c[4] * y;
    z1 = c[2] * x - c[5] * y;
    return y;
}
// w = odd extension of x (formed in float32 like numpy does on float32 audio, then widened)
__global__ void sos_extend_kernel(const float* __restrict__ x, long n, int edge, double* __restrict__ w) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n + 2L * edge) return;
    float v;
    if (i < edge) v = __fsub_rn(2.0f * x[0], x[edge - i]);
    else if (i < edge + n) v = x[i - edge];
    else v = __fsub_rn(2.0f * x[n - 1], x[n - 2 - (i - edge - n)]);
    w[i] = (double)v;
}
// One thread per recurrence block.  has_run: re-run section `run` exactly from start[] (in place); has_zs: zero-state pass of
// section `zs` over the block's (new) values -> ends[].  Both in one sweep when a section hands over to the next one.
__global__ void __launch_bounds__(32) sos_sweep_kernel(double* __restrict__ w, long m, long L, int dir, int has_run, SosSec run,
                                                       const double* __restrict__ start, int has_zs, SosSec zs,
                                                       double* __restrict__ ends) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= IIR_NB) return;
    const long lo = min((long)t * L, m), hi = min(lo + L, m);
    const long cnt = hi - lo;
    double* p = w + (dir == 0 ? lo : m - 1 - lo);          // element j of the block in processing order is p[j * sg]
    const long sg = dir == 0 ? 1 : -1;
    double r0 = 0.0, r1 = 0.0, z0 = 0.0, z1 = 0.0;
    if (has_run) { r0 = start[2 * t]; r1 = start[2 * t + 1]; }
    long j = 0;
    for (; j + IIR_B <= cnt; j += IIR_B) {
        double v[IIR_B];
#pragma unroll
        for (int k = 0; k < IIR_B; ++k) v[k] = p[(j + k) * sg];
        if (has_run) {
#pragma unroll
            for (int k = 0; k < IIR_B; ++k) {
                v[k] = sos_step(r0, r1, run.q, v[k]);
                p[(j + k) * sg] = v[k];
            }
        }
        if (has_zs) {
#pragma unroll
            for (int k = 0; k < IIR_B; ++k) sos_step(z0, z1, zs.q, v[k]);
        }
    }
    for (; j < cnt; ++j) {
        double v = p[j * sg];
        if (has_run) { v = sos_step(r0, r1, run.q, v); p[j * sg] = v; }
        if (has_zs) sos_step(z0, z1, zs.q, v);
    }
    if (has_zs) {
        const bool full = (cnt == L);                       // only a full block's end state feeds a later block
        ends[2 * t] = full ? z0 : 0.0;
        ends[2 * t + 1] = full ? z1 : 0.0;
    }
}
// start[k] = sum_{j<=k} M^(k-j) u_j with u_0 = zi * x0 (edge condition: the cascade has been fed its first input sample forever;
// x0 = w[x0_idx] read BEFORE the section's sweeps overwrite it -- it is stored in x0_keep by the first scan of a direction)
__global__ void __launch_bounds__(IIR_NB) sos_scan_kernel(const double* __restrict__ ends, double* __restrict__ start, SosScanP sp,
                                                          const double* __restrict__ w, long x0_idx, double* __restrict__ x0_keep,
                                                          int first_of_dir) {
    __shared__ double u[2][IIR_NB][2];
    const int t = threadIdx.x;
    double x0;
    if (first_of_dir) {
        x0 = w[x0_idx];
        if (t == 0) *x0_keep = x0;
    } else {
        x0 = *x0_keep;
    }
    double a0 = t == 0 ? sp.zi[0] * x0 : ends[2 * (t - 1)];
    double a1 = t == 0 ? sp.zi[1] * x0 : ends[2 * (t - 1) + 1];
    int cur = 0;
    u[0][t][0] = a0; u[0][t][1] = a1;
    __syncthreads();
#pragma unroll
    for (int d = 0; d < IIR_LOG; ++d) {
        const int off = 1 << d;
        if (t >= off) {
            const double b0 = u[cur][t - off][0], b1 = u[cur][t - off][1];
            a0 += sp.mp[d][0] * b0 + sp.mp[d][1] * b1;
            a1 += sp.mp[d][2] * b0 + sp.mp[d][3] * b1;
        }
        cur ^= 1;
        u[cur][t][0] = a0; u[cur][t][1] = a1;
        __syncthreads();
    }
    start[2 * t] = a0;
    start[2 * t + 1] = a1;
}
__global__ void sos_finish_kernel(const double* __restrict__ w, long n, int edge, float* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)w[edge + i];
}
void sosfiltfilt(const SosCoef& c, const float* x, long n, int edge, float* y, double* scratch, cudaStream_t s) {
    RVCB_CHECK(c.ns >= 1 && c.ns <= 4 && edge >= 0, "sosfiltfilt: 1..4 sections");
    RVCB_CHECK(n > edge, "The length of the input vector x must be greater than padlen");
    const long m = n + 2L * edge;
    const long L = (m + IIR_NB - 1) / IIR_NB;
    double* w = scratch;
    double* ends = scratch + ((m + 7) & ~7L);
    double* start = ends + 2 * IIR_NB;
    double* x0_keep = start + 2 * IIR_NB;
    SosSec sec[4];
    SosScanP sp[4];
    for (int k = 0; k < c.ns; ++k) {
        for (int i = 0; i < 6; ++i) sec[k].q[i] = c.sos[k][i];
        sp[k].zi[0] = c.zi[k][0]; sp[k].zi[1] = c.zi[k][1];
        // M = A^L by repeated squaring, then its 2^d powers
        double P[4] = {-c.sos[k][4], 1.0, -c.sos[k][5], 0.0}, R[4] = {1.0, 0.0, 0.0, 1.0};
        auto mul = [](const double* a, const double* b, double* o) {
            const double t0 = a[0] * b[0] + a[1] * b[2], t1 = a[0] * b[1] + a[1] * b[3];
            const double t2 = a[2] * b[0] + a[3] * b[2], t3 = a[2] * b[1] + a[3] * b[3];
            o[0] = t0; o[1] = t1; o[2] = t2; o[3] = t3;
        };
        for (long e = L; e > 0; e >>= 1) {
            if (e & 1) mul(R, P, R);
            mul(P, P, P);
        }
        for (int d = 0; d < IIR_LOG; ++d) {
            for (int i = 0; i < 4; ++i) sp[k].mp[d][i] = R[i];
            mul(R, R, R);
        }
    }
    sos_extend_kernel<<<(unsigned)ceil_div_l(m, 256), 256, 0, s>>>(x, n, edge, w);
    KERNEL_CHECK();
    int launches = 1;
    const unsigned gsw = IIR_NB / 32;
    for (int dir = 0; dir < 2; ++dir) {
        // zero-state pass of section 0, then per section: scan -> exact re-run (fused with the next section's zero-state pass)
        sos_sweep_kernel<<<gsw, 32, 0, s>>>(w, m, L, dir, 0, sec[0], start, 1, sec[0], ends);
        KERNEL_CHECK();
        ++launches;
        for (int k = 0; k < c.ns; ++k) {
            sos_scan_kernel<<<1, IIR_NB, 0, s>>>(ends, start, sp[k], w, dir == 0 ? 0 : m - 1, x0_keep, k == 0 ? 1 : 0);
            KERNEL_CHECK();
            const int nxt = (k + 1 < c.ns) ? 1 : 0;
            sos_sweep_kernel<<<gsw, 32, 0, s>>>(w, m, L, dir, 1, sec[k], start, nxt, sec[nxt ? k + 1 : k], ends);
            KERNEL_CHECK();
            launches += 2;
        }
    }
    sos_finish_kernel<<<(unsigned)ceil_div_l(n, 256), 256, 0, s>>>(w, n, edge, y);
    KERNEL_CHECK();
    count_launch(launches + 1);
}

__global__ void reflect_pad_1d_kernel(const float* __restrict__ x, long n, long pad, float* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n + 2 * pad) return;
    // np.pad(mode="reflect") for ANY pad width (a 2 s utterance is shorter than the 3 s pad of the default config): the signal
    // extended with period 2(n-1), edge samples not repeated
    long j = i - pad;
    const long period = 2 * (n - 1);
    if (period == 0) {
        j = 0;
    } else {
        j %= period;
        if (j < 0) j += period;
        if (j >= n) j = period - j;
    }
    out[i] = x[j];
}
void reflect_pad(const float* x, long n, long pad, float* out, cudaStream_t s) {
    RVCB_CHECK(pad >= 0 && n >= 1, "reflect_pad: bad shape");
    reflect_pad_1d_kernel<<<(unsigned)ceil_div_l(n + 2 * pad, 256), 256, 0, s>>>(x, n, pad, out);
    KERNEL_CHECK();
    count_launch();
}
__global__ void f32_to_i16_kernel(const float* __restrict__ x, long n, short* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (short)__float2int_rz(x[i]);
}
void f32_to_i16(const float* x, long n, short* out, cudaStream_t s) {
    f32_to_i16_kernel<<<(unsigned)ceil_div_l(n, 256), 256, 0, s>>>(x, n, out);
    KERNEL_CHECK();
    count_launch();
}

// f0 post-processing, one block.  Phase 1 (parallel): resize by np.interp over the NaN-marked contour (compiled_base.c
// arr_interp: exact-hit shortcut, slope*(x - xp[j]) + fp[j], the two NaN fallbacks), NaN -> 0.  Phase 2 (parallel): the
// gap fill of F0Predictor._interpolate_f0 from nearest-voiced-neighbour indices.  Phase 3 (parallel): key shift and mel quantisation of
// post_process.  All float64 with explicit round-to-nearest mul/add (no FMA contraction), i.e. numpy's arithmetic.
__global__ void __launch_bounds__(256) f0_post_kernel(const float* __restrict__ f0, int L, int n, double key_factor, double mel_min,
                                                      double mel_max, long long* __restrict__ pitch, float* __restrict__ pitchf,
                                                      double* __restrict__ gdata, int use_smem) {
    extern __shared__ double f0_sdata[];
    double* data = use_smem ? f0_sdata : gdata;        // the gap fill walks this array: keep it on chip when it fits
    int* nvb = use_smem ? reinterpret_cast<int*>(f0_sdata + n) : reinterpret_cast<int*>(gdata + n);   // next voiced index
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    auto src = [&](int j) -> double {
        const double v = (double)f0[j];
        return v < 0.001 ? qnan : v;
    };
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double x = (double)((long long)i * L) / (double)n;
        double r;
        if (x > (double)(L - 1)) {
            r = src(L - 1);
        } else {
            const int j = (int)x;                      // xp = arange(L): the bracketing index is floor(x)
            if (j == L - 1 || (double)j == x) {
                r = src(j);
            } else {
                const double y0 = src(j), y1 = src(j + 1);
                const double slope = __ddiv_rn(__dsub_rn(y1, y0), 1.0);
                r = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, (double)j)), y0);
                if (isnan(r)) {
                    r = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, (double)(j + 1))), y1);
                    if (isnan(r) && y0 == y1) r = y0;
                }
            }
        }
        data[i] = isnan(r) ? 0.0 : r;
    }
    __syncthreads();
    {
        // Gap fill of F0Predictor._interpolate_f0, per element: with pv / nv the nearest voiced frames on either side of an
        // unvoiced frame i (run = pv+1 .. nv-1),  nv < n-1 and pv >= 0 -> linear ramp  data[pv] + step * (i - pv),
        // step = (data[nv] - data[pv]) / (nv - pv - 1);  nv < n-1 and no pv -> data[nv];  otherwise (the run reaches the end, or
        // its right neighbour is the very last frame) -> hold data[pv] (0 if none); in that last case the reference's slice
        // assignment data[i:n] also overwrites the (voiced) final frame.  No other voiced frame is ever written, so the fill is
        // done in place; each thread walks one contiguous chunk with carries from a 256-entry summary.
        __shared__ int s_first[256], s_last[256];
        const int t = threadIdx.x;
        const int per = (n + 255) / 256;
        const int c0 = min(t * per, n), c1 = min(c0 + per, n);
        int fv = n, lv = -1;
        for (int i = c0; i < c1; ++i)
            if (data[i] > 0.0) {
                if (fv == n) fv = i;
                lv = i;
            }
        s_first[t] = fv; s_last[t] = lv;
        const bool last_overwritten = n >= 2 && data[n - 1] > 0.0 && !(data[n - 2] > 0.0);   // read before anyone writes
        __syncthreads();
        int pv = -1, nx = n;
        for (int k = t - 1; k >= 0; --k)
            if (s_last[k] >= 0) { pv = s_last[k]; break; }
        for (int k = t + 1; k < 256; ++k)
            if (s_first[k] < n) { nx = s_first[k]; break; }
        for (int i = c1 - 1; i >= c0; --i) {
            if (data[i] > 0.0) nx = i;
            else nvb[i] = nx;
        }
        for (int i = c0; i < c1; ++i) {
            if (data[i] > 0.0) {
                if (i == n - 1 && last_overwritten) data[i] = pv >= 0 ? data[pv] : 0.0;
                pv = i;
                continue;
            }
            const int nv = nvb[i];
            double val;
            if (nv < n - 1) {
                if (pv >= 0) {
                    const double base = data[pv];
                    const double step = __ddiv_rn(__dsub_rn(data[nv], base), (double)(nv - (pv + 1)));
                    val = __dadd_rn(base, __dmul_rn(step, (double)(i - pv)));
                } else {
                    val = data[nv];
                }
            } else {
                val = pv >= 0 ? data[pv] : 0.0;
            }
            data[i] = val;
        }
    }
    __syncthreads();
    const double span = __dsub_rn(mel_max, mel_min);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double f = __dmul_rn(data[i], key_factor);
        double mel = __dmul_rn(1127.0, log(__dadd_rn(1.0, __ddiv_rn(f, 700.0))));
        if (mel > 0.0) mel = __dadd_rn(__ddiv_rn(__dmul_rn(__dsub_rn(mel, mel_min), 254.0), span), 1.0);
        if (mel <= 1.0) mel = 1.0;
        if (mel > 255.0) mel = 255.0;
        pitch[i] = (long long)rint(mel);
        pitchf[i] = (float)f;
    }
}
void f0_post(const float* f0, int n_frames, int p_len, double key_factor, double f0_min, double f0_max, long long* pitch, float* pitchf,
             double* scratch, cudaStream_t s) {
    RVCB_CHECK(n_frames >= 1 && p_len >= 1, "f0_post: empty contour");
    const double mel_min = 1127.0 * std::log(1.0 + f0_min / 700.0), mel_max = 1127.0 * std::log(1.0 + f0_max / 700.0);
    const int use_smem = p_len <= 16000 ? 1 : 0;       // 192 KB of the 227 KB a CTA may opt into
    const size_t smem = use_smem ? (size_t)p_len * 12 : 0;
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(f0_post_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16000 * 12));
        configured = true;
    }
    f0_post_kernel<<<1, 256, smem, s>>>(f0, n_frames, p_len, key_factor, mel_min, mel_max, pitch, pitchf, scratch, use_smem);
    KERNEL_CHECK();
    count_launch();
}

}  // namespace rvcb

// =====================================================================================================================
// Realtime tail of gui.py's audio callback, per block, on the device (gui.py:1024-1087): volume-envelope mix
// (librosa.feature.rms(frame 4*zc, hop zc) of the input and of the converted block, align_corners linear interpolation,
// power-law mix) and SOLA (normalised cross-correlation of the block head with the previous block's tail, arg-max offset,
// sin^2 cross-fade, buffer update).
// =====================================================================================================================
namespace rvcb {

__global__ void rt_rms_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, int frame, int hop, int nf, float* __restrict__ rms) {
    const int sig = blockIdx.x / nf, f = blockIdx.x % nf;
    const float* y = sig ? b : a;
    const int lo = max(f * hop - frame / 2, 0), hi = min(f * hop - frame / 2 + frame, n);     // centred frame, zero padding
    float s = 0.f;
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) s += y[i] * y[i];
    __shared__ float sh[32];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) rms[sig * nf + f] = sqrtf(s / (float)frame);
    }
}

// F.interpolate(rms[None], size = n + 1, mode = "linear", align_corners = True)[0, 0, :-1], then y *= (r1 / max(r2, 1e-3)) ^ (1 - rate)
__global__ void rt_mix_kernel(float* __restrict__ y, int n, const float* __restrict__ rms, int nf, float expo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float scale = (float)(nf - 1) / (float)n;
    const float src = scale * (float)i;
    const int i0 = (int)src, i1 = min(i0 + 1, nf - 1);
    const float l1 = src - (float)i0, l0 = 1.f - l1;
    const float r1 = l0 * rms[i0] + l1 * rms[i1];
    const float r2 = fmaxf(l0 * rms[nf + i0] + l1 * rms[nf + i1], 1e-3f);
    y[i] *= powf(r1 / r2, expo);
}

__global__ void sola_corr_kernel(const float* __restrict__ y, const float* __restrict__ buf, int nbuf, float* __restrict__ score) {
    const int o = blockIdx.x;
    float nom = 0.f, den = 0.f;
    for (int i = threadIdx.x; i < nbuf; i += blockDim.x) {
        const float v = y[o + i];
        nom += v * buf[i];
        den += v * v;
    }
    __shared__ float sh[2][32];
    for (int k = 16; k; k >>= 1) { nom += __shfl_xor_sync(0xffffffffu, nom, k); den += __shfl_xor_sync(0xffffffffu, den, k); }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = nom; sh[1][threadIdx.x >> 5] = den; }
    __syncthreads();
    if (threadIdx.x < 32) {
        nom = threadIdx.x < (blockDim.x >> 5) ? sh[0][threadIdx.x] : 0.f;
        den = threadIdx.x < (blockDim.x >> 5) ? sh[1][threadIdx.x] : 0.f;
        for (int k = 16; k; k >>= 1) { nom += __shfl_xor_sync(0xffffffffu, nom, k); den += __shfl_xor_sync(0xffffffffu, den, k); }
        if (threadIdx.x == 0) score[o] = nom / sqrtf(den + 1e-8f);
    }
}

__device__ __forceinline__ float sola_fade_in(int i, int n) {      // sin(0.5*pi*linspace(0, 1, n))^2, torch's float32 linspace
    if (n <= 1) return 0.f;
    const float step = 1.f / (float)(n - 1);
    const float t = i < n / 2 ? (float)i * step : 1.f - (float)(n - 1 - i) * step;
    const float s = sinf(1.5707963267948966f * t);
    return s * s;
}

// one block: arg-max of the scores (first maximum), cross-fade, output block, new SOLA buffer
__global__ void sola_finish_kernel(const float* __restrict__ y, const float* __restrict__ score, int nscore, float* __restrict__ buf, int nbuf,
                                   int block, float* __restrict__ out, int* __restrict__ offset_out, const float* __restrict__ xf) {
    extern __shared__ float newbuf[];
    __shared__ float bv[32];
    __shared__ int bi[32], off_s;
    float best = -INFINITY;
    int idx = 0x7fffffff;
    for (int i = threadIdx.x; i < nscore; i += blockDim.x) {
        const float v = score[i];
        if (v > best || (v == best && i < idx)) { best = v; idx = i; }
    }
    for (int k = 16; k; k >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, k);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, k);
        if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
    }
    if ((threadIdx.x & 31) == 0) { bv[threadIdx.x >> 5] = best; bi[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x < 32) {
        best = threadIdx.x < (blockDim.x >> 5) ? bv[threadIdx.x] : -INFINITY;
        idx = threadIdx.x < (blockDim.x >> 5) ? bi[threadIdx.x] : 0x7fffffff;
        for (int k = 16; k; k >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, k);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, k);
            if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
        }
        if (threadIdx.x == 0) {
            off_s = idx == 0x7fffffff ? 0 : idx;
            if (offset_out) *offset_out = off_s;
        }
    }
    __syncthreads();
    const int off = off_s;
    auto faded = [&](int i) {          // element i of infer_wav[off:] after the in-place cross-fade of its first nbuf samples
        if (i < nbuf && xf) return xf[i];                 // phase-vocoder cross-fade, computed by rt_pv_* from the same offset
        float v = y[off + i];
        if (i < nbuf) {
            const float fi = sola_fade_in(i, nbuf);
            v = v * fi + buf[i] * (1.f - fi);
        }
        return v;
    };
    for (int i = threadIdx.x; i < nbuf; i += blockDim.x) newbuf[i] = faded(block + i);
    for (int i = threadIdx.x; i < block; i += blockDim.x) out[i] = faded(i);
    __syncthreads();
    for (int i = threadIdx.x; i < nbuf; i += blockDim.x) buf[i] = newbuf[i];
}

// ---- phase-vocoder cross-fade (gui.py:27-48, the use_pv branch of the callback) ----
// offset = arg-max of the SOLA score (first maximum), for the kernels below
__global__ void sola_argmax_kernel(const float* __restrict__ score, int nscore, int* __restrict__ off) {
    __shared__ float bv[32];
    __shared__ int bi[32];
    float best = -INFINITY;
    int idx = 0x7fffffff;
    for (int i = threadIdx.x; i < nscore; i += blockDim.x) {
        const float v = score[i];
        if (v > best || (v == best && i < idx)) { best = v; idx = i; }
    }
    for (int k = 16; k; k >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, k);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, k);
        if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
    }
    if ((threadIdx.x & 31) == 0) { bv[threadIdx.x >> 5] = best; bi[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x < 32) {
        best = threadIdx.x < (blockDim.x >> 5) ? bv[threadIdx.x] : -INFINITY;
        idx = threadIdx.x < (blockDim.x >> 5) ? bi[threadIdx.x] : 0x7fffffff;
        for (int k = 16; k; k >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, k);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, k);
            if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
        }
        if (threadIdx.x == 0) *off = idx == 0x7fffffff ? 0 : idx;
    }
}
// One block per bin f: fa = rfft(a * w), fb = rfft(b * w), w = sqrt(fade_out * fade_in), a = sola_buffer, b = y[off : off + n].
// spec[f] = (|fa| + |fb|) (x2 for the interior bins), spec[F + f] = angle(fa), spec[2F + f] = wrapped angle(fb) - angle(fa).
__global__ void __launch_bounds__(128) rt_pv_dft_kernel(const float* __restrict__ a, const float* __restrict__ y, const int* __restrict__ off,
                                                        int n, float* __restrict__ spec) {
    __shared__ float red[4][4];
    const int f = blockIdx.x, F = n / 2 + 1;
    const float* b = y + *off;
    float are = 0.f, aim = 0.f, bre = 0.f, bim = 0.f;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const float fi = sola_fade_in(k, n);
        const float w = sqrtf((1.f - fi) * fi);
        const int r = (int)(((long)f * k) % n);
        float sn, cs;
        sincospif(2.f * (float)r / (float)n, &sn, &cs);
        const float av = a[k] * w, bv = b[k] * w;
        are = fmaf(av, cs, are); aim = fmaf(-av, sn, aim);
        bre = fmaf(bv, cs, bre); bim = fmaf(-bv, sn, bim);
    }
    float v[4] = {are, aim, bre, bim};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        for (int o = 16; o; o >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][j] = v[j];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t[4];
        for (int j = 0; j < 4; ++j) t[j] = red[0][j] + red[1][j] + red[2][j] + red[3][j];
        float mag = sqrtf(t[0] * t[0] + t[1] * t[1]) + sqrtf(t[2] * t[2] + t[3] * t[3]);
        const bool interior = f >= 1 && ((n % 2 == 0) ? f < F - 1 : true);
        if (interior) mag *= 2.f;
        const float pa = atan2f(t[1], t[0]), pb = atan2f(t[3], t[2]);
        float d = pb - pa;
        d = d - 6.283185307179586f * floorf(d / 2.f / 3.141592653589793f + 0.5f);
        spec[f] = mag; spec[F + f] = pa; spec[2 * F + f] = d;
    }
}
// xf[i] = a[i] fo^2 + b[i] fi^2 + sum_f mag_f cos((2 pi f + d_f) i / n + pa_f) * w_i / n; the 2 pi f i / n part is reduced exactly.
__global__ void rt_pv_synth_kernel(const float* __restrict__ a, const float* __restrict__ y, const int* __restrict__ off, int n,
                                   const float* __restrict__ spec, float* __restrict__ xf) {
    extern __shared__ float s_spec[];
    const int F = n / 2 + 1;
    for (int j = threadIdx.x; j < 3 * F; j += blockDim.x) s_spec[j] = spec[j];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float t = (float)i / (float)n;
    float acc = 0.f;
    for (int f = 0; f < F; ++f) {
        const int r = (int)(((long)f * i) % n);
        const float ph = 6.283185307179586f * ((float)r / (float)n) + s_spec[2 * F + f] * t + s_spec[F + f];
        acc = fmaf(s_spec[f], cosf(ph), acc);
    }
    const float fi = sola_fade_in(i, n), fo = 1.f - fi;
    xf[i] = a[i] * (fo * fo) + y[*off + i] * (fi * fi) + acc * sqrtf(fo * fi) / (float)n;
}

void rt_tail(float* infer, int n, const float* input, int zc, float rms_mix_rate, float* sola_buffer, int block_frame, int nbuf, int nsearch,
             float* out, float* scratch, int* offset, cudaStream_t s, bool use_pv) {
    RVCB_CHECK(n >= block_frame + nbuf + nsearch && nbuf >= 1 && nbuf <= 8192 && zc >= 1, "rt_tail: bad sizes");
    const int nf = 1 + n / zc;
    float* rms = scratch;                       // [2, nf]
    float* score = scratch + 2 * nf;            // [nsearch + 1]
    if (rms_mix_rate < 1.f) {
        RVCB_CHECK(input != nullptr, "rt_tail: the envelope mix needs the input window");
        rt_rms_kernel<<<2 * nf, 256, 0, s>>>(input, infer, n, 4 * zc, zc, nf, rms);
        rt_mix_kernel<<<ceil_div(n, 256), 256, 0, s>>>(infer, n, rms, nf, 1.f - rms_mix_rate);
        count_launch(2);
    }
    sola_corr_kernel<<<nsearch + 1, 256, 0, s>>>(infer, sola_buffer, nbuf, score);
    const float* xf = nullptr;
    if (use_pv) {
        const int F = nbuf / 2 + 1;
        int* off = reinterpret_cast<int*>(score + nsearch + 1);      // [1] (+1 pad)
        float* spec = score + nsearch + 3;                            // [3, F]
        float* x = spec + 3 * F;                                      // [nbuf]
        RVCB_CHECK((size_t)3 * F * sizeof(float) <= 96 * 1024, "rt_tail: phase-vocoder frame too long");
        static bool configured = false;
        if (!configured) {
            CUDA_CHECK(cudaFuncSetAttribute(rt_pv_synth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            configured = true;
        }
        sola_argmax_kernel<<<1, 1024, 0, s>>>(score, nsearch + 1, off);
        rt_pv_dft_kernel<<<F, 128, 0, s>>>(sola_buffer, infer, off, nbuf, spec);
        rt_pv_synth_kernel<<<ceil_div(nbuf, 128), 128, (size_t)3 * F * sizeof(float), s>>>(sola_buffer, infer, off, nbuf, spec, x);
        count_launch(3);
        xf = x;
    }
    sola_finish_kernel<<<1, 1024, nbuf * sizeof(float), s>>>(infer, score, nsearch + 1, sola_buffer, nbuf, block_frame, out, offset, xf);
    KERNEL_CHECK();
    count_launch(2);
}

}  // namespace rvcb
