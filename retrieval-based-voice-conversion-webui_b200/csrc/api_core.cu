// C ABI: runtime + op-level entry points (see include/rvcb200.h).
#include "../../include/rvcb200.h"
#include "common.cuh"
#include "gemm.cuh"

#include <cstring>

namespace rvcb {
unsigned long long g_launch_count = 0;
int g_grid_cap = 0;
static thread_local std::string g_err;
void set_last_error(const std::string& s) { g_err = s; }
}  // namespace rvcb

#include <vector>
#include "api_macros.h"

extern "C" {

const char* rvcb_last_error(void) { return rvcb::g_err.c_str(); }
unsigned long long rvcb_launch_count(void) { return rvcb::g_launch_count; }
int rvcb_set_grid_cap(int max_ctas) {
    const int prev = rvcb::g_grid_cap;
    rvcb::g_grid_cap = max_ctas > 0 ? max_ctas : 0;
    return prev;
}
const char* rvcb_version(void) { return "rvcb200 0.1 (sm_100a)"; }

int rvcb_init(int device) {
    RVCB_API_BEGIN
    int n = 0;
    CUDA_CHECK(cudaGetDeviceCount(&n));
    RVCB_CHECK(n > 0 && device < n, "no CUDA device " + std::to_string(device));
    CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    RVCB_CHECK(prop.major == 10, std::string("librvcb200 is built for sm_100a only; device is sm_") +
                                     std::to_string(prop.major) + std::to_string(prop.minor));
    RVCB_API_END
}

int rvcb_op_gemm(const rvcb_gemm_desc* d, int impl, void* stream) {
    RVCB_API_BEGIN
    rvcb::GemmArgs g;
    g.A = (const __half*)d->A; g.lda = d->lda; g.a_rows = d->a_rows; g.a_cols = d->a_cols; g.conv2d_W = d->conv2d_W;
    g.B = (const __half*)d->B; g.ldb = d->ldb; g.b_rows = d->b_rows; g.b_cols = d->b_cols;
    g.M = d->M; g.N = d->N; g.block_k = d->block_k; g.nseg = d->nseg;
    RVCB_CHECK(d->nseg > 0 && d->nseg <= rvcb::GEMM_MAX_SEG, "bad nseg");
    for (int i = 0; i < d->nseg; ++i) g.seg[i] = {d->seg[i].row_off, d->seg[i].col_off, d->seg[i].dw, d->seg[i].nk};
    g.batch = d->batch; g.a_row_z = d->a_row_z; g.a_col_z = d->a_col_z; g.b_row_z = d->b_row_z; g.b_col_z = d->b_col_z;
    g.c_z = d->c_z; g.bias_z = d->bias_z; g.b_col0 = d->b_col0;
    g.bias = d->bias; g.bias_per_row = d->bias_per_row; g.res1 = d->res1; g.ldres1 = d->ldres1; g.res2 = d->res2; g.ldres2 = d->ldres2;
    g.alpha = d->alpha; g.act1 = d->act1; g.act1_p = d->act1_p; g.act2 = d->act2; g.act2_p = d->act2_p; g.gate = d->gate;
    g.out32 = d->out32; g.ld32 = d->ld32; g.out16 = (__half*)d->out16; g.ld16 = d->ld16; g.up2_C = d->up2_C;
    if (impl == 1) rvcb::gemm_simt(g, (cudaStream_t)stream);
    else rvcb::gemm_tc(g, (cudaStream_t)stream);
    RVCB_API_END
}

}  // extern "C"

// ---- weight container ----------------------------------------------------------------------
#include "weights.cuh"
extern "C" {
int rvcb_weights_create(rvcb_weights** out) {
    RVCB_API_BEGIN
    RVCB_CHECK(out, "null argument");
    *out = new rvcb_weights();
    RVCB_API_END
}
int rvcb_weights_add(rvcb_weights* w, const char* name, const float* host_data, int ndim, const int64_t* shape) {
    RVCB_API_BEGIN
    RVCB_CHECK(w && name && (host_data || ndim == 0), "null argument");
    rvcb_weights::Tensor t;
    int64_t n = 1;
    for (int i = 0; i < ndim; ++i) {
        t.shape.push_back(shape[i]);
        n *= shape[i];
    }
    t.data.assign(host_data, host_data + n);
    w->t[std::string(name)] = std::move(t);
    RVCB_API_END
}
void rvcb_weights_destroy(rvcb_weights* w) { delete w; }
}

#include "kernels.cuh"
extern "C" int rvcb_upsample_protect(const float* d_feats, const float* d_feats0, int T_h, int C, const float* d_pitchf, int T,
                                     float protect, float* d_out, void* stream) {
    RVCB_API_BEGIN
    RVCB_CHECK(d_feats && d_out, "null argument");
    rvcb::upsample_protect(d_feats, d_feats0, T_h, C, d_pitchf, T, protect, d_out, (cudaStream_t)stream);
    RVCB_API_END
}

extern "C" int rvcb_prof_begin(void) {
    RVCB_API_BEGIN
    rvcb::gemm_prof_begin();
    RVCB_API_END
}
extern "C" int rvcb_prof_end(double* gemm_ms, unsigned long long* gemm_launches) {
    RVCB_API_BEGIN
    rvcb::gemm_prof_end(gemm_ms, gemm_launches);
    RVCB_API_END
}

extern "C" int rvcb_post_mix(float* d_wav, int64_t n_out, int tgt_sr, const float* d_audio16k, int64_t n_in, float rms_mix_rate,
                             double* d_scratch, void* stream) {
    RVCB_API_BEGIN
    RVCB_CHECK(d_wav && d_audio16k && d_scratch && n_out > 0 && n_in > 0 && tgt_sr >= 16000, "bad argument");
    rvcb::post_mix(d_wav, n_out, tgt_sr, d_audio16k, n_in, rms_mix_rate, d_scratch, (cudaStream_t)stream);
    RVCB_API_END
}

extern "C" int rvcb_rms_mix(float* d_wav, int64_t n_out, int tgt_sr, const float* d_audio16k, int64_t n_in, float rms_mix_rate,
                            double* d_scratch, void* stream) {
    RVCB_API_BEGIN
    RVCB_CHECK(d_wav && d_audio16k && d_scratch && n_out > 0 && n_in > 0 && tgt_sr >= 16000, "bad argument");
    rvcb::post_mix(d_wav, n_out, tgt_sr, d_audio16k, n_in, rms_mix_rate, d_scratch, (cudaStream_t)stream, false);
    RVCB_API_END
}

// scipy.signal.lfilter's float64 loop (direct form II transposed), one pass; z is the state, updated in place.
// NC known at compile time keeps the state in registers (the recurrence is one add -> mul -> sub dependency chain per sample).
template <int NC>
static void lfilter_df2t_fixed(const double* b, const double* a, const double* x, int64_t n, int64_t stride, double* y, double* zio) {
    double z[NC - 1];
    for (int k = 0; k < NC - 1; ++k) z[k] = zio[k];
    for (int64_t i = 0; i < n; ++i) {
        const double xi = x[i * stride];
        const double yi = z[0] + b[0] * xi;
        for (int k = 1; k < NC - 1; ++k) z[k - 1] = z[k] + xi * b[k] - yi * a[k];      // NC is a constant: fully unrolled
        z[NC - 2] = xi * b[NC - 1] - yi * a[NC - 1];
        y[i * stride] = yi;
    }
    for (int k = 0; k < NC - 1; ++k) zio[k] = z[k];
}
static void lfilter_df2t(const double* b, const double* a, int nc, const double* x, int64_t n, int64_t stride, double* y, double* z) {
    if (nc == 6) return lfilter_df2t_fixed<6>(b, a, x, n, stride, y, z);
    for (int64_t i = 0; i < n; ++i) {
        const double xi = x[i * stride];
        const double yi = z[0] + b[0] * xi;
        for (int k = 1; k < nc - 1; ++k) z[k - 1] = z[k] + xi * b[k] - yi * a[k];
        z[nc - 2] = xi * b[nc - 1] - yi * a[nc - 1];
        y[i * stride] = yi;
    }
}

extern "C" int rvcb_host_filtfilt(const double* b, const double* a, const double* zi, int ncoef, const void* xv, int x_is_f32, int64_t n,
                                  double* y) {
    RVCB_API_BEGIN
    RVCB_CHECK(b && a && zi && xv && y && ncoef >= 2 && ncoef <= 16 && a[0] == 1.0, "bad filter");
    const int64_t edge = 3 * (int64_t)ncoef;
    RVCB_CHECK(n > edge, "The length of the input vector x must be greater than padlen");
    // odd extension: 2*x[0] - x[edge..1], x, 2*x[n-1] - x[n-2..n-edge-1]   (in the input's own precision, as numpy does)
    std::vector<double> ext((size_t)(n + 2 * edge));
    if (x_is_f32) {
        const float* x = static_cast<const float*>(xv);
        for (int64_t i = 0; i < edge; ++i) ext[(size_t)i] = (double)(float)(2.0f * x[0] - x[edge - i]);
        for (int64_t i = 0; i < n; ++i) ext[(size_t)(edge + i)] = (double)x[i];
        for (int64_t i = 0; i < edge; ++i) ext[(size_t)(edge + n + i)] = (double)(float)(2.0f * x[n - 1] - x[n - 2 - i]);
    } else {
        const double* x = static_cast<const double*>(xv);
        for (int64_t i = 0; i < edge; ++i) ext[(size_t)i] = 2.0 * x[0] - x[edge - i];
        for (int64_t i = 0; i < n; ++i) ext[(size_t)(edge + i)] = x[i];
        for (int64_t i = 0; i < edge; ++i) ext[(size_t)(edge + n + i)] = 2.0 * x[n - 1] - x[n - 2 - i];
    }
    const int64_t m = n + 2 * edge;
    double z[16];
    for (int k = 0; k < ncoef - 1; ++k) z[k] = zi[k] * ext[0];
    lfilter_df2t(b, a, ncoef, ext.data(), m, 1, ext.data(), z);                 // forward
    for (int k = 0; k < ncoef - 1; ++k) z[k] = zi[k] * ext[(size_t)(m - 1)];
    lfilter_df2t(b, a, ncoef, ext.data() + (m - 1), m, -1, ext.data() + (m - 1), z);   // backward, in place
    for (int64_t i = 0; i < n; ++i) y[i] = ext[(size_t)(edge + i)];
    RVCB_API_END
}

extern "C" int rvcb_sosfiltfilt(const double* sos, const double* zi, int n_sections, int edge, const float* d_x, int64_t n, float* d_y,
                                double* d_scratch, void* stream) {
    RVCB_API_BEGIN
    RVCB_CHECK(sos && zi && d_x && d_y && d_scratch, "null argument");
    RVCB_CHECK(n_sections >= 1 && n_sections <= 4, "rvcb_sosfiltfilt: 1..4 sections");
    rvcb::SosCoef c{};
    c.ns = n_sections;
    for (int s = 0; s < n_sections; ++s) {
        RVCB_CHECK(sos[s * 6 + 3] == 1.0, "rvcb_sosfiltfilt: sections must be normalised (a0 == 1)");
        for (int k = 0; k < 6; ++k) c.sos[s][k] = sos[s * 6 + k];
        c.zi[s][0] = zi[s * 2];
        c.zi[s][1] = zi[s * 2 + 1];
    }
    rvcb::sosfiltfilt(c, d_x, n, edge, d_y, d_scratch, (cudaStream_t)stream);
    RVCB_API_END
}

extern "C" int rvcb_reflect_pad(const float* d_x, int64_t n, int64_t pad, float* d_out, void* stream) {
    RVCB_API_BEGIN
    RVCB_CHECK(d_x && d_out && n > 0, "bad argument");
    rvcb::reflect_pad(d_x, n, pad, d_out, (cudaStream_t)stream);
    RVCB_API_END
}

extern "C" int rvcb_f32_to_i16(const float* d_x, int64_t n, int16_t* d_out, void* stream) {
    RVCB_API_BEGIN
    RVCB_CHECK(d_x && d_out && n > 0, "bad argument");
    rvcb::f32_to_i16(d_x, n, d_out, (cudaStream_t)stream);
    RVCB_API_END
}

extern "C" int rvcb_f0_post(const float* d_f0, int n_frames, int p_len, double key_factor, double f0_min, double f0_max, int64_t* d_pitch,
                            float* d_pitchf, double* d_scratch, void* stream) {
    RVCB_API_BEGIN
    RVCB_CHECK(d_f0 && d_pitch && d_pitchf && d_scratch, "null argument");
    rvcb::f0_post(d_f0, n_frames, p_len, key_factor, f0_min, f0_max, (long long*)d_pitch, d_pitchf, d_scratch, (cudaStream_t)stream);
    RVCB_API_END
}

#include "resblock_fused.cuh"
extern "C" int64_t rvcb_op_resblock1_out_rows(int C, int k, const int* dil, int T) {
    if (!dil || !rvcb::resblock_fused_supported(C, k, dil)) return -1;
    return rvcb::resblock_fused_out_rows(C, k, dil, T);
}
extern "C" int rvcb_op_resblock1(int C, int k, const int* dil, const float* const* w1, const float* const* b1, const float* const* w2,
                                 const float* const* b2, const float* d_x, int T, float* d_y, void* stream) {
    RVCB_API_BEGIN
    RVCB_CHECK(dil && w1 && b1 && w2 && b2 && d_x && d_y && T > 0, "null argument");
    RVCB_CHECK(rvcb::resblock_fused_supported(C, k, dil), "resblock1: unsupported (C, k, dilations)");
    rvcb::DevOwner own;       // test entry: packs on every call
    const rvcb::RBFusedWeights w = rvcb::pack_resblock_fused(own, C, k, dil, w1, b1, w2, b2);
    rvcb::resblock_fused(w, d_x, d_y, T, (cudaStream_t)stream);
    CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    RVCB_API_END
}

extern "C" int rvcb_rt_tail(float* d_infer, int n, const float* d_input, int zc, float rms_mix_rate, float* d_sola_buffer, int block_frame,
                            int sola_buffer_frame, int sola_search_frame, float* d_out, float* d_scratch, int* d_offset, void* stream) {
    RVCB_API_BEGIN
    RVCB_CHECK(d_infer && d_sola_buffer && d_out && d_scratch, "null argument");
    rvcb::rt_tail(d_infer, n, d_input, zc, rms_mix_rate, d_sola_buffer, block_frame, sola_buffer_frame, sola_search_frame, d_out, d_scratch,
                  d_offset, (cudaStream_t)stream);
    RVCB_API_END
}

extern "C" int rvcb_rt_tail_pv(float* d_infer, int n, const float* d_input, int zc, float rms_mix_rate, float* d_sola_buffer, int block_frame,
                               int sola_buffer_frame, int sola_search_frame, int use_pv, float* d_out, float* d_scratch, int* d_offset,
                               void* stream) {
    RVCB_API_BEGIN
    RVCB_CHECK(d_infer && d_sola_buffer && d_out && d_scratch, "null argument");
    rvcb::rt_tail(d_infer, n, d_input, zc, rms_mix_rate, d_sola_buffer, block_frame, sola_buffer_frame, sola_search_frame, d_out, d_scratch,
                  d_offset, (cudaStream_t)stream, use_pv != 0);
    RVCB_API_END
}

extern "C" int rvcb_prof_classes(double* ms2, double* launches2, double* flops2, double* bytes2) {
    RVCB_API_BEGIN
    RVCB_CHECK(ms2 && launches2 && flops2 && bytes2, "null argument");
    rvcb::gemm_prof_classes(ms2, launches2, flops2, bytes2);
    RVCB_API_END
}
