/* librvcb200 -- C ABI of the B200-native RVC inference hot path.
 *
 * The reference (fumiama/Retrieval-based-Voice-Conversion-WebUI) has no FFI/plugin
 * interface: its seams are duck-typed Python objects.  This header is
 * the boundary a maintainer binds *immediately beneath* those Python seams; each entry point
 * names the reference interface it replaces.  INTEGRATION.md shows the ctypes stubs.
 *
 * Conventions: plain pointers and sizes only; pointers prefixed d_ are DEVICE pointers owned
 * by the caller; `stream` is a cudaStream_t passed as void*; every function returns 0 on
 * success and a negative code on failure with the message available from rvcb_last_error();
 * no C++ exception crosses the ABI; a handle is safe for one caller at a time
 * (the reference's own threading model: SURVEY.md 8b "Threading").
 */
#ifndef RVCB200_H
#define RVCB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rvcb_weights rvcb_weights;
typedef struct rvcb_hubert rvcb_hubert;
typedef struct rvcb_index rvcb_index;
typedef struct rvcb_flat rvcb_flat;
typedef struct rvcb_rmvpe rvcb_rmvpe;
typedef struct rvcb_synth rvcb_synth;

/* ---- runtime ---------------------------------------------------------------------------- */
int rvcb_init(int device);                 /* cudaSetDevice + capability check (sm_100 required) */
const char* rvcb_last_error(void);
unsigned long long rvcb_launch_count(void); /* kernels launched by this library so far */
const char* rvcb_version(void);
/* Cap the CTAs of the persistent GEMM / kNN grids launched from now on (0 = every SM); returns the previous cap.  Process-wide
 * launch-time setting: the front doors use it to run the two independent front branches side by side on disjoint SMs. */
int rvcb_set_grid_cap(int max_ctas);

/* per-launch CUDA-event timing of the implicit-GEMM kernel (bench.py roofline); begin resets, end syncs and sums */
int rvcb_prof_begin(void);
int rvcb_prof_end(double* gemm_ms, unsigned long long* gemm_launches);
/* per-kernel-class totals of the last profiled region; each array has 3 entries: [0] streaming gemm_tc_kernel (+ split-K), [1] weight-
 * stationary gemm_ws*_kernel, [2] fused residual-block kernel (ms, launches, 2*M*N*K flops incl. K padding, algorithmic HBM bytes
 * (classes 1 and 2)) */
int rvcb_prof_classes(double* ms3, double* launches3, double* flops3, double* bytes3);

/* ---- weight container (host fp32 tensors keyed by the reference's state_dict names) ------
 * replaces: torch.load + load_state_dict in rvc/synthesizer.py:10-35, rvc/f0/models.py:9-11,
 * infer/modules/vc/utils.py:24-36 (the Python side reads the .pth/.pt and hands tensors over). */
int rvcb_weights_create(rvcb_weights** out);
int rvcb_weights_add(rvcb_weights* w, const char* name, const float* host_data, int ndim, const int64_t* shape);
void rvcb_weights_destroy(rvcb_weights* w);

/* ---- HuBERT-base content features -----------------------------------------------------------
 * replaces: fairseq HubertModel.extract_features(source, padding_mask, output_layer)
 * called at infer/modules/vc/pipeline.py:102-110, infer/lib/rtrvc.py:154-162 and
 * .final_proj (v1, pipeline.py:110). */
int rvcb_hubert_create(const rvcb_weights* w, rvcb_hubert** out);
int rvcb_hubert_num_frames(int n_samples);
/* d_wav: f32[n_samples]; d_out: f32[T_h, 768]; returns T_h through n_frames */
int rvcb_hubert_extract_features(rvcb_hubert* h, const float* d_wav, int n_samples, int output_layer,
                                 float* d_out, int* n_frames, void* stream);
/* d_in f32[T,768] -> d_out f32[T,256] */
int rvcb_hubert_final_proj(rvcb_hubert* h, const float* d_in, int T, float* d_out, void* stream);
void rvcb_hubert_destroy(rvcb_hubert* h);

/* ---- IVF-Flat retrieval ---------------------------------------------------------------------
 * replaces: faiss.read_index(...).search(npy, k=8) / reconstruct_n and the numpy blend at
 * infer/modules/vc/pipeline.py:113-138, infer/lib/rtrvc.py:169-185. */
int rvcb_index_create(const float* centroids, int nlist, const float* vectors, int64_t ntotal, int d,
                      const int64_t* list_off /*[nlist+1]*/, const int64_t* list_ids /*[ntotal]*/,
                      rvcb_index** out);
int64_t rvcb_index_ntotal(const rvcb_index* ix);
/* nprobe = 1.  d_q f32[nq,d] -> d_D f32[nq,k] (ascending squared L2), d_I i64[nq,k]; missing = (3.4028235e38,-1) */
int rvcb_index_search(rvcb_index* ix, const float* d_q, int nq, int k, float* d_D, int64_t* d_I, void* stream);
/* feats_out = rate * sum_k w_k big_npy[I_k] + (1-rate) * feats_in,  w = (1/D)^2 normalised (pipeline.py:129-138) */
int rvcb_index_blend(rvcb_index* ix, const float* d_feats_in, int nq, int k, const float* d_D, const int64_t* d_I,
                     float index_rate, float* d_feats_out, void* stream);
/* exact brute-force L2 top-1 over d_db f32[n,d] (BASELINE config #5) */
int rvcb_knn_bruteforce_top1(const float* d_db, int64_t n, int d, const float* d_q, int nq, float* d_D,
                             int64_t* d_I, void* stream);
void rvcb_index_destroy(rvcb_index* ix);
/* exact brute-force L2 top-1 for query BATCHES (BASELINE config #5 at nq >= 32): the handle keeps an fp16 mirror of d_db (which
 * must stay alive) and ||v||^2; scores come from the tcgen05 GEMM engine, the 32 best candidates per query are re-ranked with
 * the exact fp32 lane-order distance and a rounding-error certificate sends any doubtful query to the exact scan, so d_D / d_I
 * are bit-identical to rvcb_knn_bruteforce_top1.  RVCB_KNN_TC=0 forces the scan. */
int rvcb_flat_create(const float* d_db, int64_t n, int d, rvcb_flat** out);
int rvcb_flat_search_top1(rvcb_flat* f, const float* d_q, int nq, float* d_D, int64_t* d_I, void* stream);
void rvcb_flat_destroy(rvcb_flat* f);

/* ---- retrieval epilogue: x2 nearest upsample + protect mix (pipeline.py:140-160) ----------- */
/* d_feats f32[T_h,C], d_feats0 (nullable) f32[T_h,C], d_pitchf (nullable) f32[T]; out f32[T,C], T <= 2*T_h */
int rvcb_upsample_protect(const float* d_feats, const float* d_feats0, int T_h, int C, const float* d_pitchf, int T,
                          float protect, float* d_out, void* stream);

/* ---- output epilogue: RMS-envelope mix + peak normalisation to the int16 range, in place (change_rms pipeline.py:26-45,
 * scaling pipeline.py:356-360).  d_wav f32[n_out] at tgt_sr; d_audio16k f32[n_in] = the (filtered) 16 kHz input;
 * d_scratch: >= n_in/8000 + n_out/(tgt_sr/2) + 8 doubles. */
int rvcb_post_mix(float* d_wav, int64_t n_out, int tgt_sr, const float* d_audio16k, int64_t n_in, float rms_mix_rate,
                  double* d_scratch, void* stream);
/* change_rms alone (pipeline.py:349-350), without the peak scaling: for the resample_sr branch (pipeline.py:351-354), where the
 * resampler runs between the mix and the scaling -- rvcb_rms_mix, rvcb_resample_sinc, then rvcb_post_mix(rate = 1) = scaling only. */
int rvcb_rms_mix(float* d_wav, int64_t n_out, int tgt_sr, const float* d_audio16k, int64_t n_in, float rms_mix_rate,
                  double* d_scratch, void* stream);

/* ---- input front end ----------------------------------------------------------------------
 * rvcb_host_filtfilt: HOST function (no GPU): scipy.signal.filtfilt(b, a, x) with its defaults (odd extension,
 * padlen = 3*max(len(a), len(b)), zi scaled by the edge sample), the 48 Hz high-pass of pipeline.py:23,221.  Same
 * direct-form-II-transposed recurrence and operation order as scipy's lfilter, float64.  nb == na == ncoef <= 16,
 * a[0] == 1; zi = scipy.signal.lfilter_zi(b, a) (ncoef-1 values, a constant of the filter).  x is float64[n], or
 * float32[n] when x_is_f32 != 0 (then the odd extension is formed in float32 like numpy does before scipy upcasts). */
int rvcb_host_filtfilt(const double* b, const double* a, const double* zi, int ncoef, const void* x, int x_is_f32, int64_t n, double* y);
/* The same filter on the DEVICE, as a cascade of second-order sections (scipy.signal.sosfiltfilt semantics with padtype "odd",
 * padlen = edge): sos f64[n_sections,6] rows [b0 b1 b2 1 a1 a2], zi f64[n_sections,2] = scipy.signal.sosfilt_zi(sos) (edge state
 * per unit input level).  d_x f32[n] -> d_y f32[n], float64 arithmetic.  The recurrence is cut into 1024 blocks per section
 * (zero-state pass, log-step 2x2 state prefix with A^L, exact re-run), so it agrees with the host filter to ~1e-7, not bit for
 * bit: the direct 5th-order form cannot be block-propagated in float64.  d_scratch: >= n + 2*edge + 4120 doubles. */
int rvcb_sosfiltfilt(const double* sos, const double* zi, int n_sections, int edge, const float* d_x, int64_t n, float* d_y,
                     double* d_scratch, void* stream);
/* np.pad(x, (pad, pad), mode="reflect") on the device (pipeline.py:241): d_out f32[n + 2*pad]; any pad width (periodic reflection) */
int rvcb_reflect_pad(const float* d_x, int64_t n, int64_t pad, float* d_out, void* stream);
/* (int16) x with C truncation, the .astype(np.int16) of modules.py:181: d_out i16[n] */
int rvcb_f32_to_i16(const float* d_x, int64_t n, int16_t* d_out, void* stream);

/* ---- realtime tail of gui.py's audio callback, per block, on the device (gui.py:1024-1087) ---------------------------
 * replaces: the volume-envelope mix (librosa.feature.rms(frame 4*zc, hop zc) of input and output, align_corners interpolation,
 * (rms1 / max(rms2, 1e-3)) ^ (1 - rms_mix_rate), gui.py:1024-1056; skipped when rms_mix_rate >= 1) and the SOLA step (normalised
 * correlation of the block head with the previous tail, arg-max offset, sin^2 cross-fade, buffer update; gui.py:1057-1087, the
 * use_pv = False branch).  d_infer f32[n], n >= block_frame + sola_buffer_frame + sola_search_frame (scaled in place by the mix);
 * d_input f32[>= n]: input window from extra_frame on, at the output rate (nullable when rms_mix_rate >= 1); d_sola_buffer
 * f32[sola_buffer_frame]: state, in/out; d_out f32[block_frame]; d_scratch: >= 2*(n/zc + 1) + sola_search_frame + 1 floats;
 * d_offset (nullable) i32[1]: the chosen offset. */
int rvcb_rt_tail(float* d_infer, int n, const float* d_input, int zc, float rms_mix_rate, float* d_sola_buffer, int block_frame,
                 int sola_buffer_frame, int sola_search_frame, float* d_out, float* d_scratch, int* d_offset, void* stream);

/* The same with the phase-vocoder cross-fade of gui.py:27-48 in place of the sin^2 cross-fade (gui.py:1078-1083, use_pv = True):
 * rfft of the windowed previous tail and new head, magnitudes added, phase advanced linearly from the old phase to the new one.
 * use_pv = 0 is rvcb_rt_tail.  d_scratch: >= 2*(n/zc + 1) + sola_search_frame + 4 + 3*(sola_buffer_frame/2 + 1) + sola_buffer_frame. */
int rvcb_rt_tail_pv(float* d_infer, int n, const float* d_input, int zc, float rms_mix_rate, float* d_sola_buffer, int block_frame,
                    int sola_buffer_frame, int sola_search_frame, int use_pv, float* d_out, float* d_scratch, int* d_offset, void* stream);

/* ---- realtime noise gate and resamplers around RVC.infer, per block, on the device -------------------------------------
 * rvcb_torchgate_*  replaces: infer/modules/gui/torchgate.py TorchGate.__init__ :33-70 / forward :217-280 (the torch.stft /
 * torch.istft branch; stationary mask :128-178 with or without a noise reference, non-stationary mask :180-215, mask smoothing
 * :253-258) as used at gui.py:869-871 (n_fft = 4 * zc, prop_decrease 0.9), :974-990 (input) and :1015-1023 (output).  fp32.
 * h_filter: HOST f32[filter_rows, filter_cols] smoothing filter of _generate_mask_smoothing_filter :72-126 (rows = frequency,
 * odd sizes; 0 x 0 = none).  apply: d_x f32[n], d_xn f32[n_noise] noise reference (nullable: the signal's own statistics),
 * d_y f32[rvcb_torchgate_out_len(h, n)] = hop * (n / hop) samples (torch.istft's length); batch 1 (the GUI's unsqueeze(0)). */
typedef struct rvcb_torchgate rvcb_torchgate;
int rvcb_torchgate_create(int sr, int n_fft, int hop, int nonstationary, float n_std_thresh_stationary, float n_thresh_nonstationary,
                          float temp_coeff_nonstationary, int n_movemean_nonstationary, float prop_decrease, const float* h_filter,
                          int filter_rows, int filter_cols, rvcb_torchgate** out);
int64_t rvcb_torchgate_out_len(const rvcb_torchgate* h, int64_t n);
int rvcb_torchgate_apply(rvcb_torchgate* h, const float* d_x, int64_t n, const float* d_xn, int64_t n_noise, float* d_y, void* stream);
void rvcb_torchgate_destroy(rvcb_torchgate* h);
/* replaces: torchaudio.transforms.Resample.forward (gui.py:851-866 resampler / resampler2, applied at gui.py:991-1000, :1008-1009):
 * out[i * up + p] = sum_k kernel[p, k] * xpad[i * down + k], xpad = x with `width` zeros on the left; d_kernel f32[up, kernel_width]
 * is the windowed-sinc table of torchaudio's _get_sinc_resample_kernel (built by the host mirror with torchaudio's own formula),
 * up / down = new / orig rate over their gcd, n_out = ceil(up * n / down). */
int rvcb_resample_sinc(const float* d_x, int64_t n, const float* d_kernel, int up, int down, int kernel_width, int width, float* d_out,
                       int64_t n_out, void* stream);

/* ---- f0 post-processing on the device (no host round trip between RMVPE and the synthesizer) ----
 * replaces: F0Predictor._resize_f0 + _interpolate_f0 (rvc/f0/f0.py:31-78) and post_process (rvc/f0/gen.py:10-41, without
 * a manual f0 curve), float64 with numpy's operation order.  d_f0 f32[n_frames] (Hz, 0 = unvoiced) -> d_pitch i64[p_len]
 * (coarse mel bins 1..255) and d_pitchf f32[p_len] (Hz, gaps filled, shifted by key_factor = 2^(f0_up_key/12)).
 * d_scratch: >= 2 * p_len doubles. */
int rvcb_f0_post(const float* d_f0, int n_frames, int p_len, double key_factor, double f0_min, double f0_max, int64_t* d_pitch,
                 float* d_pitchf, double* d_scratch, void* stream);

/* ---- RMVPE f0 -------------------------------------------------------------------------------
 * replaces: RMVPE.compute_f0 -> mel_extractor + _mel2hidden + _decode (rvc/f0/rmvpe.py:96-164). */
int rvcb_rmvpe_create(const rvcb_weights* w, rvcb_rmvpe** out);
int rvcb_rmvpe_num_frames(int n_samples);
/* d_wav f32[n]; outputs (each nullable): d_mel f32[128, n_frames] log-mel, d_hidden f32[n_frames,360] salience,
 * d_f0 f32[n_frames] decoded Hz (threshold thred, 0 = unvoiced) */
int rvcb_rmvpe_infer(rvcb_rmvpe* h, const float* d_wav, int n_samples, float thred, float* d_mel, float* d_hidden,
                     float* d_f0, int* n_frames, void* stream);
void rvcb_rmvpe_destroy(rvcb_rmvpe* h);

/* ---- synthesizer ----------------------------------------------------------------------------
 * replaces: SynthesizerTrnMsNSFsid.infer (rvc/layers/synthesizers.py:159-203). */
typedef struct rvcb_synth_config {
    int inter_channels, hidden_channels, filter_channels, n_heads, n_layers, kernel_size;
    int n_resblock_kernels;  int resblock_kernel_sizes[4];  int resblock_dilations[4][3];
    int n_upsamples;         int upsample_rates[8];         int upsample_kernel_sizes[8];      /* 4 (v2, v1/40k) or 5 (v1/32k, v1/48k) stages */
    int upsample_initial_channel, spk_embed_dim, gin_channels, sr, encoder_dim;
} rvcb_synth_config;
int rvcb_synth_create(const rvcb_synth_config* cfg, const rvcb_weights* w, rvcb_synth** out);
/* d_phone f32[T,encoder_dim]; d_pitch i64[T] (1..255); d_pitchf f32[T] Hz;
 * d_noise_prior f32[inter, T - flow_head] (channel-major, like randn_like(m_p)); d_noise_src f32[T_dec*upp];
 * skip_head/return_length/return_length2 < 0 mean "None".  d_wav_out f32[T_out*upp]; n_out receives T_out*upp. */
int rvcb_synth_infer(rvcb_synth* h, const float* d_phone, int T, int sid, const int64_t* d_pitch, const float* d_pitchf,
                     const float* d_noise_prior, const float* d_noise_src, int skip_head, int return_length,
                     int return_length2, float* d_wav_out, int* n_out, void* stream);
/* The offline caller discards the x_pad context of every chunk (pipeline.py:241,295: audio1[t_pad_tgt : -t_pad_tgt]).  This entry returns
 * exactly rvcb_synth_infer(...)[keep_head*upp : (keep_head+keep_length)*upp] -- bit for bit -- but runs the local parts of the model
 * (flow: receptive field +-24 frames; decoder: +-10 frames) only over the kept frames plus margins; the TextEncoder (global attention) and
 * the NSF sine phase (a running sum from frame 0) still cover all T frames.  d_noise_prior f32[inter, T] and d_noise_src f32[T*upp] are
 * the FULL-length tensors of the untrimmed call; d_wav_out f32[keep_length*upp]. */
int rvcb_synth_infer_keep(rvcb_synth* h, const float* d_phone, int T, int sid, const int64_t* d_pitch, const float* d_pitchf,
                          const float* d_noise_prior, const float* d_noise_src, int keep_head, int keep_length, float* d_wav_out,
                          int* n_out, void* stream);
void rvcb_synth_destroy(rvcb_synth* h);

/* ---- op-level entry for the unit tests: the implicit-GEMM engine -------------------------- */
typedef struct rvcb_gemm_seg { int row_off, col_off, dw, nk; } rvcb_gemm_seg;
typedef struct rvcb_gemm_desc {
    const void* A; int64_t lda; int a_rows, a_cols, conv2d_W;
    const void* B; int64_t ldb; int b_rows, b_cols;
    int M, N, block_k, nseg;
    int batch; int64_t a_row_z, a_col_z, b_row_z, b_col_z, c_z, bias_z; int b_col0;
    const float* bias; int bias_per_row;
    const float* res1; int64_t ldres1; const float* res2; int64_t ldres2;
    float alpha; int act1; float act1_p; int act2; float act2_p; int gate;
    float* out32; int64_t ld32; void* out16; int64_t ld16; int up2_C;
    rvcb_gemm_seg seg[128];
} rvcb_gemm_desc;
/* impl: 0 = tcgen05/TMA kernel (product), 1 = SIMT restatement (validation only) */
int rvcb_op_gemm(const rvcb_gemm_desc* d, int impl, void* stream);

/* ---- op-level entry for the unit tests: the fused residual block of the vocoder -------------
 * replaces: ResBlock1.forward (rvc/layers/residuals.py:68-85): for d in dil: x = x + c2(lrelu(c1_d(lrelu(x, 0.1)), 0.1)).
 * Host weights w1[i], w2[i]: f32 [C, C, k] (weight-norm folded), b1[i], b2[i]: f32 [C]; i = 0..2.  d_x f32 [T, C] (channels last);
 * d_y f32 [T, C].  rvcb_op_resblock1_out_rows returns T when (C, k, dil) has a fused variant (C in {32, 64}), else -1. */
int64_t rvcb_op_resblock1_out_rows(int C, int k, const int* dil, int T);
int rvcb_op_resblock1(int C, int k, const int* dil, const float* const* w1, const float* const* b1, const float* const* w2,
                      const float* const* b2, const float* d_x, int T, float* d_y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RVCB200_H */
