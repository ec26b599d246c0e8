// extern "C" surface of libwaveglow_b200.so — thin argument adapters over the kernels' host launchers.
// Declarations (with the reference code each entry point replaces) live in include/waveglow_b200.h.
#include "../../include/waveglow_b200.h"

#include "common.cuh"

namespace wgb {
// wn_tc.cu
int tc_wn_gate(const void*, const void*, const void*, const float*, void*, int, int, int, cudaStream_t);
int tc_wn_res(const void*, const void*, const float*, const void*, void*, int, int, cudaStream_t);
int tc_wn_skip_end(const void*, int, const void*, const float*, const float*, float*, const float*, float*, int, int,
                   int, int, cudaStream_t);
int tc_gemm_plain(const void*, const void*, const float*, void*, int, int, int, int, int, cudaStream_t);
int tc_gemm_split3(const void*, const void*, const void*, const float*, void*, int, int, int, int, long long, long long,
                   cudaStream_t);
int tc_conv1d_taps(const void*, const void*, const float*, void*, int, int, int, int, int, int, int, int, cudaStream_t);
int tc_stft_mag(const void*, const void*, const void*, void*, int, int, int, int, long long, long long, cudaStream_t);
int tc2_stft_mel(const void*, const void*, const void*, const void*, float*, int, int, int, int, int, int, int, float, cudaStream_t);
int tc2_stft_denoise(const void*, const void*, const void*, const float*, float, void*, void*, int, int, int, int, int, int,
                     cudaStream_t);
int tc2_istft_ola(const void*, const void*, const void*, const float*, float*, int, int, int, int, cudaStream_t);
int fft_stft_mel(const float*, const float*, const void*, int, const float*, int, float*, int, int, int, int, float, int*, cudaStream_t);
int fft_denoise(const float*, const float*, const float*, float, const float*, float*, int, int, int, cudaStream_t);
int tc2_gemm_split3(const void*, const void*, const void*, float*, long long, int, int, cudaStream_t);
int tc_stft_mel(const void*, const void*, const void*, const void*, void*, int, int, int, int, long long, long long, int, float,
                cudaStream_t);
int tc_stft_denoise(const void*, const void*, const void*, const float*, float, void*, void*, int, int, int, int, int,
                    long long, long long, cudaStream_t);
// wn_tc2.cu
int tc2_wn_gate(const void*, const void*, const void*, const float*, void*, void*, int, int, int, cudaStream_t);
int tc2_wn_res(const void*, const void*, const float*, const void*, void*, int, int, long long, const void*, float*, int,
               cudaStream_t);
int tc2_wn_gate_mel(const void*, const void*, const void*, const void*, const float*, void*, int, int, int, int,
                    const float*, float*, int, cudaStream_t);
int tc2_wn_gate_mel0(const void*, const void*, const void*, const void*, const float*, void*, int, int, int, const float*,
                     float*, int, cudaStream_t);
// flow.cu
int x_stack(const float*, void*, int, int, long long, int, cudaStream_t);
int end_from_acc(const float*, const float*, float*, const float*, float*, int, int, int, int, const float*, const float*,
                 int, void*, long long, cudaStream_t);
int tc2_wn_skip_end(const void*, int, const void*, const float*, const float*, float*, const float*, float*, int, int,
                    int, int, const float*, const float*, int, void*, long long, cudaStream_t);
// wn_skip16.cu
int tc_wn_skip16_end(const void*, int, const void*, const float*, float*, const float*, float*, int, int, int, int,
                     const float*, const float*, int, void*, long long, const float*, const float*, cudaStream_t);
// ref_f32.cu
int sgemm_nt(const float*, const float*, const float*, void*, int, int, int, int, int, long long, long long, long long,
             long long, long long, int, int, cudaStream_t);
int gate_f32(const float*, float*, long long, int, cudaStream_t);
int act_f32(float*, long long, int, cudaStream_t);
int fused_add_tanh_sigmoid_multiply(const float*, const float*, float*, int, int, int, cudaStream_t);
int res_skip_f32(const float*, float*, float*, long long, int, int, int, cudaStream_t);
// flow.cu
int flow_from_z(const float*, float*, int, int, float, cudaStream_t);
int flow_to_z(const float*, float*, int, int, cudaStream_t);
int flow_mix(float*, const float*, long long, int, cudaStream_t);
int wn_start(const float*, const float*, const float*, void*, int, long long, int, int, int, long long, cudaStream_t);
int end_coupling_f32(const float*, const float*, const float*, float*, const float*, float*, int, int, int, int, int,
                     cudaStream_t);
int upsample_im2col(const float*, void*, int, int, int, int, int, int, cudaStream_t);
int cast_f32_to_bf16(const float*, void*, long long, cudaStream_t);
int audio_to_int16(const float*, void*, long long, float, cudaStream_t);
// training direction: wn_tc2.cu, wn_tc.cu, wn_wgrad.cu, train.cu
int tc2_wn_res_taps(const void*, const void*, const float*, const void*, void*, int, int, long long, int, int, int,
                    cudaStream_t);
int tc2_wn_res_seg(const void*, const void*, int, int, const void*, const float*, const void*, void*, int, int, long long, int,
                   int, int, cudaStream_t);
int tc_gemm_seg(const void*, const void*, int, int, const void*, const float*, const void*, void*, int, int, int, int, int,
                int, int, int, int, cudaStream_t);
int tc_wgrad(const void*, const void*, float*, int, int, int, int, int, int, int, cudaStream_t);
int gate_bwd(const void*, void*, float*, long long, int, cudaStream_t);
int coupling_bwd(float*, const float*, const float*, const float*, const float*, float*, void*, void*, int, int, int, int,
                 cudaStream_t);
int start_bwd(float*, const void*, const float*, long long, int, int, cudaStream_t);
int colsum8_f32(const float*, float*, long long, int, cudaStream_t);
int mix_bwd(float*, const float*, const float*, float*, long long, int, cudaStream_t);
int upsample_wgrad(const float*, const float*, float*, float*, int, int, int, int, int, int, int, int, cudaStream_t);
int adam_step(float*, const float*, float*, float*, long long, float, float, float, float, int, float, cudaStream_t);
int adam_step_dev(float*, const float*, float*, float*, long long, float, float, float, float, int*, float, cudaStream_t);
int logdet(const float*, float*, float*, int, cudaStream_t);
// stft.cu
int stft_reflect_pad(const float*, float*, int, int, int, long long, cudaStream_t);
int stft_reflect_pad_split(const float*, void*, void*, int, int, int, long long, int*, cudaStream_t);
int split_bf16(const float*, void*, void*, long long, cudaStream_t);
int stft_polar(const float*, float*, float*, float*, int, int, int, int, cudaStream_t);
int mel_log(const float*, float*, int, int, int, float, cudaStream_t);
int denoise_scale(float*, const float*, float, long long, int, int, cudaStream_t);
int denoise_scale_split(const float*, const float*, float, void*, void*, long long, int, int, cudaStream_t);
int stft_recombine(const float*, const float*, float*, int, int, int, int, cudaStream_t);
int spec_set_magnitude(float*, const float*, int, int, int, int, cudaStream_t);
int istft_overlap_add(const float*, const double*, float*, int, int, int, int, cudaStream_t);
}  // namespace wgb

using namespace wgb;
static inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

extern "C" {

WGB_API int wgb_abi_version(void) { return WGB_ABI_VERSION; }
WGB_API int wgb_set_tuning(const char* key, int value) { return tuning_set(key, value); }
#ifndef WGB_SOURCE_HASH
#define WGB_SOURCE_HASH "unknown"
#endif
// the marker prefix lets build.py read the hash out of the file without dlopen-ing a possibly stale library
static const char kSourceHash[] = "WGB_SOURCE_HASH=" WGB_SOURCE_HASH;
WGB_API const char* wgb_source_hash(void) { return kSourceHash + 16; }
WGB_API const char* wgb_last_error(void) { return error_buffer(); }

WGB_API int wgb_device_check(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(WGB_ERR_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(WGB_ERR_DEVICE, "device %d not present (%d visible)", device, n);
    int major = 0;
    WGB_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10) return fail(WGB_ERR_DEVICE, "device %d is sm_%d0; this library is sm_100a only", device, major);
    return WGB_OK;
}

WGB_API int wgb_flow_from_z(const float* z, float* x, int batch, int T, float sigma, void* stream) {
    return flow_from_z(z, x, batch, T, sigma, S(stream));
}
WGB_API int wgb_flow_to_z(const float* x, float* z, int batch, int T, void* stream) { return flow_to_z(x, z, batch, T, S(stream)); }
WGB_API int wgb_flow_mix(float* x, const float* w, long long rows, int C, void* stream) { return flow_mix(x, w, rows, C, S(stream)); }
WGB_API int wgb_wn_start(const float* x, const float* w, const float* bias, void* h, int out_bf16, long long rows, int n_ch,
                 int n_half, void* stream) {
    return wn_start(x, w, bias, h, out_bf16, rows, n_ch, n_half, 0, 0, S(stream));
}
WGB_API int wgb_wn_start_padded(const float* x, const float* w, const float* bias, void* h, int out_bf16, int batch, int T,
                                long long h_batch_rows, int n_ch, int n_half, void* stream) {
    return wn_start(x, w, bias, h, out_bf16, static_cast<long long>(batch) * T, n_ch, n_half, T, h_batch_rows, S(stream));
}

WGB_API int wgb_tc_wn_gate(const void* h, const void* cond, const void* w_packed, const float* bias, void* acts, int batch,
                   int T, int dilation, void* stream) {
    return tc_wn_gate(h, cond, w_packed, bias, acts, batch, T, dilation, S(stream));
}
WGB_API int wgb_tc2_wn_gate(const void* h, const void* cond, const void* w_packed, const float* bias, void* acts, int batch,
                            int T, int dilation, void* stream) {
    return tc2_wn_gate(h, cond, w_packed, bias, acts, nullptr, batch, T, dilation, S(stream));
}
WGB_API int wgb_tc2_wn_gate_train(const void* h, const void* cond, const void* w_packed, const float* bias, void* acts,
                                  void* ts, int batch, int T, int dilation, void* stream) {
    WGB_REQUIRE(ts != nullptr, "ts is null");
    return tc2_wn_gate(h, cond, w_packed, bias, acts, ts, batch, T, dilation, S(stream));
}
WGB_API int wgb_tc2_wn_gate_mel(const void* h, const void* mel_stack, const void* w_packed, const void* w_mel,
                                const float* bias, void* acts, int batch, int T, int frames_pad, int dilation,
                                const float* w_comp, float* skip_acc, int skip_first, void* stream) {
    return tc2_wn_gate_mel(h, mel_stack, w_packed, w_mel, bias, acts, batch, T, frames_pad, dilation, w_comp, skip_acc,
                           skip_first, S(stream));
}
WGB_API int wgb_tc2_wn_gate_mel0(const void* x_stack_, const void* mel_stack, const void* w0, const void* w_mel, const float* bias,
                                 void* acts, int batch, int T, int frames_pad, const float* w_comp, float* skip_acc,
                                 int skip_first, void* stream) {
    return tc2_wn_gate_mel0(x_stack_, mel_stack, w0, w_mel, bias, acts, batch, T, frames_pad, w_comp, skip_acc, skip_first,
                            S(stream));
}
WGB_API int wgb_x_stack(const float* x, void* out, int batch, int T, long long out_batch_rows, int n_half, void* stream) {
    return x_stack(x, out, batch, T, out_batch_rows, n_half, S(stream));
}
WGB_API int wgb_end_from_acc(const float* skip_acc, const float* b_end, float* x, const float* w_mix, float* log_s, int batch,
                             int T, int n_half, int direction, const float* next_w_start, const float* next_b_start,
                             int next_n_half, void* h_next, long long h_next_batch_rows, void* stream) {
    return end_from_acc(skip_acc, b_end, x, w_mix, log_s, batch, T, n_half, direction, next_w_start, next_b_start,
                        next_n_half, h_next, h_next_batch_rows, S(stream));
}
WGB_API int wgb_tc2_wn_res(const void* acts, const void* w_res, const float* bias, const void* h_in, void* h_out, int batch,
                           int T, long long h_batch_rows, const void* w16_layer, float* skip_acc, int skip_first,
                           void* stream) {
    return tc2_wn_res(acts, w_res, bias, h_in, h_out, batch, T, h_batch_rows, w16_layer, skip_acc, skip_first, S(stream));
}
WGB_API int wgb_tc2_wn_skip_end(const void* acts_all, int n_layers, const void* w_skip, const float* w_end,
                                const float* b_end, float* x, const float* w_mix, float* log_s, int batch, int T,
                                int n_half, int direction, const float* next_w_start, const float* next_b_start,
                                int next_n_half, void* h_next, long long h_next_batch_rows, void* stream) {
    return tc2_wn_skip_end(acts_all, n_layers, w_skip, w_end, b_end, x, w_mix, log_s, batch, T, n_half, direction,
                           next_w_start, next_b_start, next_n_half, h_next, h_next_batch_rows, S(stream));
}
WGB_API int wgb_tc_wn_skip16_end(const void* acts_all, int n_layers, const void* w16, const float* b_end, float* x,
                                 const float* w_mix, float* log_s, int batch, int T, int n_half, int direction,
                                 const float* next_w_start, const float* next_b_start, int next_n_half, void* h_next,
                                 long long h_next_batch_rows, const float* skip_acc, const float* next_w_mix,
                                 void* stream) {
    return tc_wn_skip16_end(acts_all, n_layers, w16, b_end, x, w_mix, log_s, batch, T, n_half, direction, next_w_start,
                            next_b_start, next_n_half, h_next, h_next_batch_rows, skip_acc, next_w_mix, S(stream));
}
WGB_API int wgb_tc_wn_res(const void* acts, const void* w_res, const float* bias, const void* h_in, void* h_out, int batch,
                  int T, void* stream) {
    return tc_wn_res(acts, w_res, bias, h_in, h_out, batch, T, S(stream));
}
WGB_API int wgb_tc_wn_skip_end(const void* acts_all, int n_layers, const void* w_skip, const float* w_end, const float* b_end,
                       float* x, const float* w_mix, float* log_s, int batch, int T, int n_half, int direction,
                       void* stream) {
    return tc_wn_skip_end(acts_all, n_layers, w_skip, w_end, b_end, x, w_mix, log_s, batch, T, n_half, direction,
                          S(stream));
}

WGB_API int wgb_tc_gemm(const void* a, const void* w, const float* bias, void* c, int out_bf16, int batch, int T, int N,
                        int K, void* stream) {
    return tc_gemm_plain(a, w, bias, c, out_bf16, batch, T, N, K, S(stream));
}

WGB_API int wgb_tc_gemm_split3(const void* a_hi, const void* a_lo, const void* w3, const float* bias, void* c, int batch,
                               int rows, int N, int K, long long row_stride, long long batch_stride, void* stream) {
    return tc_gemm_split3(a_hi, a_lo, w3, bias, c, batch, rows, N, K, row_stride, batch_stride, S(stream));
}

WGB_API int wgb_tc_stft_mag(const void* a_hi, const void* a_lo, const void* w3_paired, void* mag_cl, int batch, int rows,
                            int cp, int K, long long row_stride, long long batch_stride, void* stream) {
    return tc_stft_mag(a_hi, a_lo, w3_paired, mag_cl, batch, rows, cp, K, row_stride, batch_stride, S(stream));
}
WGB_API int wgb_tc_conv1d(const void* a, const void* w, const float* bias, void* c, int out_bf16, int batch, int T, int N,
                          int C, int taps, int dilation, int act, void* stream) {
    return tc_conv1d_taps(a, w, bias, c, out_bf16, batch, T, N, C, taps, dilation, act, S(stream));
}
WGB_API int wgb_tc_stft_mel(const void* a_hi, const void* a_lo, const void* w3_paired, const void* mel_table, float* out,
                            int batch, int rows, int cp, int K, long long row_stride, long long batch_stride, int n_mel,
                            float clip, void* stream) {
    return tc_stft_mel(a_hi, a_lo, w3_paired, mel_table, out, batch, rows, cp, K, row_stride, batch_stride, n_mel, clip,
                       S(stream));
}
WGB_API int wgb_tc_stft_denoise(const void* a_hi, const void* a_lo, const void* w3_paired, const float* bias_spec,
                                float strength, void* hi_out, void* lo_out, int batch, int rows, int cutoff, int cp, int K,
                                long long row_stride, long long batch_stride, void* stream) {
    return tc_stft_denoise(a_hi, a_lo, w3_paired, bias_spec, strength, hi_out, lo_out, batch, rows, cutoff, cp, K,
                           row_stride, batch_stride, S(stream));
}

WGB_API int wgb_tc2_stft_mel(const void* a_hi, const void* a_lo, const void* w3_paired, const void* mel_table, float* out,
                             int batch, int frames, int R, int L, int hop, int n_pass, int n_mel, float clip, void* stream) {
    return tc2_stft_mel(a_hi, a_lo, w3_paired, mel_table, out, batch, frames, R, L, hop, n_pass, n_mel, clip, S(stream));
}
WGB_API int wgb_tc2_stft_denoise(const void* a_hi, const void* a_lo, const void* w3_paired, const float* bias_spec,
                                 float strength, void* hi_out, void* lo_out, int batch, int frames, int R, int L, int hop,
                                 int out_R, void* stream) {
    return tc2_stft_denoise(a_hi, a_lo, w3_paired, bias_spec, strength, hi_out, lo_out, batch, frames, R, L, hop, out_R,
                            S(stream));
}
WGB_API int wgb_tc2_istft_ola(const void* s_hi, const void* s_lo, const void* w_ola, const float* env_tab, float* out,
                              int batch, int frames, int L, int hop, void* stream) {
    return tc2_istft_ola(s_hi, s_lo, w_ola, env_tab, out, batch, frames, L, hop, S(stream));
}
WGB_API int wgb_fft_stft_mel(const float* y, const float* window, const void* mel_slots, int slots_per_lane,
                             const float* mel_w, int bins_used, float* out, int batch, int n, int hop, int n_mel,
                             float clip, int* range_flag, void* stream) {
    return fft_stft_mel(y, window, mel_slots, slots_per_lane, mel_w, bins_used, out, batch, n, hop, n_mel, clip, range_flag,
                        S(stream));
}
WGB_API int wgb_fft_denoise(const float* y, const float* window, const float* bias_spec, float strength,
                            const float* env_tab, float* out, int batch, int n, int hop, void* stream) {
    return fft_denoise(y, window, bias_spec, strength, env_tab, out, batch, n, hop, S(stream));
}
WGB_API int wgb_tc2_gemm_split3(const void* a_hi, const void* a_lo, const void* w3, float* c, long long rows, int N, int K,
                                void* stream) {
    return tc2_gemm_split3(a_hi, a_lo, w3, c, rows, N, K, S(stream));
}

WGB_API int wgb_sgemm_f32(const float* A, const float* W, const float* bias, void* C, int out_bf16, int batch, int M, int N,
                  int K, long long lda, long long a_batch, long long ldw, long long ldc, long long c_batch, int shift,
                  int accumulate, void* stream) {
    return sgemm_nt(A, W, bias, C, out_bf16, batch, M, N, K, lda, a_batch, ldw, ldc, c_batch, shift, accumulate,
                    S(stream));
}
WGB_API int wgb_gate_f32(const float* u, float* acts, long long rows, int n_ch, void* stream) {
    return gate_f32(u, acts, rows, n_ch, S(stream));
}
WGB_API int wgb_fused_add_tanh_sigmoid_multiply(const float* input_a, const float* input_b, float* acts, int batch, int n_ch,
                                                int T, void* stream) {
    return fused_add_tanh_sigmoid_multiply(input_a, input_b, acts, batch, n_ch, T, S(stream));
}
WGB_API int wgb_act_f32(float* x, long long n, int act, void* stream) { return act_f32(x, n, act, S(stream)); }
WGB_API int wgb_res_skip_f32(const float* rs, float* h, float* skip, long long rows, int n_ch, int has_res, int first,
                     void* stream) {
    return res_skip_f32(rs, h, skip, rows, n_ch, has_res, first, S(stream));
}
WGB_API int wgb_end_coupling_f32(const float* skip, const float* w_end, const float* b_end, float* x, const float* w_mix,
                         float* log_s, int batch, int T, int n_ch, int n_half, int direction, void* stream) {
    return end_coupling_f32(skip, w_end, b_end, x, w_mix, log_s, batch, T, n_ch, n_half, direction, S(stream));
}

WGB_API int wgb_upsample_im2col(const float* mel, void* a, int out_bf16, int batch, int n_mel, int F, int taps, int ld_tap,
                        void* stream) {
    return upsample_im2col(mel, a, out_bf16, batch, n_mel, F, taps, ld_tap, S(stream));
}
WGB_API int wgb_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream) {
    return cast_f32_to_bf16(src, dst, n, S(stream));
}

WGB_API int wgb_audio_to_int16(const float* audio, void* pcm, long long n, float scale, void* stream) {
    return audio_to_int16(audio, pcm, n, scale, S(stream));
}

WGB_API int wgb_stft_reflect_pad(const float* y, float* ypad, int batch, int N, int half, long long ld_pad, void* stream) {
    return stft_reflect_pad(y, ypad, batch, N, half, ld_pad, S(stream));
}
WGB_API int wgb_stft_reflect_pad_split(const float* y, void* hi, void* lo, int batch, int N, int half, long long ld_pad,
                                       void* stream) {
    return stft_reflect_pad_split(y, hi, lo, batch, N, half, ld_pad, nullptr, S(stream));
}
WGB_API int wgb_stft_reflect_pad_split_check(const float* y, void* hi, void* lo, int batch, int N, int half, long long ld_pad,
                                             int* range_flag, void* stream) {
    WGB_REQUIRE(range_flag, "null range_flag");
    return stft_reflect_pad_split(y, hi, lo, batch, N, half, ld_pad, range_flag, S(stream));
}
WGB_API int wgb_split_bf16(const float* src, void* hi, void* lo, long long n, void* stream) {
    return split_bf16(src, hi, lo, n, S(stream));
}
WGB_API int wgb_stft_polar(const float* spec, float* mag, float* phase, float* mag_cl, int batch, int F, int cutoff, int cp,
                   void* stream) {
    return stft_polar(spec, mag, phase, mag_cl, batch, F, cutoff, cp, S(stream));
}
WGB_API int wgb_mel_log(const float* raw, float* out, int batch, int F, int n_mel, float clip, void* stream) {
    return mel_log(raw, out, batch, F, n_mel, clip, S(stream));
}
WGB_API int wgb_denoise_scale(float* spec, const float* bias, float strength, long long rows, int cutoff, int cp, void* stream) {
    return denoise_scale(spec, bias, strength, rows, cutoff, cp, S(stream));
}
WGB_API int wgb_spec_set_magnitude(float* spec, const float* target, int batch, int F, int cutoff, int cp, void* stream) {
    return spec_set_magnitude(spec, target, batch, F, cutoff, cp, S(stream));
}
WGB_API int wgb_denoise_scale_split(const float* spec, const float* bias, float strength, void* hi, void* lo, long long rows,
                                    int cutoff, int cp, void* stream) {
    return denoise_scale_split(spec, bias, strength, hi, lo, rows, cutoff, cp, S(stream));
}
WGB_API int wgb_stft_recombine(const float* mag, const float* phase, float* spec, int batch, int F, int cutoff, int cp,
                       void* stream) {
    return stft_recombine(mag, phase, spec, batch, F, cutoff, cp, S(stream));
}
WGB_API int wgb_istft_overlap_add(const float* frames, const double* win_sq, float* out, int batch, int F, int L, int hop,
                          void* stream) {
    return istft_overlap_add(frames, win_sq, out, batch, F, L, hop, S(stream));
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ training direction
WGB_API int wgb_tc_gemm_seg(const void* a0, const void* a1, int n_seg, int seg_mask, const void* w, const float* bias,
                            const void* res, void* c, int out_bf16, int batch, int T, int N, int C, int shift0, int dshift,
                            int act, int stacked, void* stream) {
    return tc_gemm_seg(a0, a1, n_seg, seg_mask, w, bias, res, c, out_bf16, batch, T, N, C, shift0, dshift, act, stacked,
                       S(stream));
}
WGB_API int wgb_tc_wgrad(const void* g, const void* x, float* dw, int batch, int T, int ca, int cb, int taps, int dilation,
                         int accumulate, void* stream) {
    return tc_wgrad(g, x, dw, batch, T, ca, cb, taps, dilation, accumulate, S(stream));
}
WGB_API int wgb_gate_bwd(const void* g_acts, void* ts, float* db, long long rows, int n_ch, void* stream) {
    return gate_bwd(g_acts, ts, db, rows, n_ch, S(stream));
}
WGB_API int wgb_coupling_bwd(float* g_x, const float* x_mix, const float* log_s, const float* g_log_s, const float* w_end_t,
                             float* g_out, void* g_skip, void* stack, int batch, int T, int n_ch, int n_half, void* stream) {
    return coupling_bwd(g_x, x_mix, log_s, g_log_s, w_end_t, g_out, g_skip, stack, batch, T, n_ch, n_half, S(stream));
}
WGB_API int wgb_start_bwd(float* g_x, const void* g_h0, const float* w_start, long long rows, int n_ch, int n_half,
                          void* stream) {
    return start_bwd(g_x, g_h0, w_start, rows, n_ch, n_half, S(stream));
}
WGB_API int wgb_colsum8_f32(const float* a, float* out, long long rows, int accumulate, void* stream) {
    return colsum8_f32(a, out, rows, accumulate, S(stream));
}
WGB_API int wgb_mix_bwd(float* g_x, const float* x_pre, const float* w, float* dw, long long rows, int C, void* stream) {
    return mix_bwd(g_x, x_pre, w, dw, rows, C, S(stream));
}
WGB_API int wgb_upsample_wgrad(const float* mel, const float* g_cond, float* dw, float* db, int batch, int n_mel, int frames,
                               int T, int ld, int ksize, int stride, int n_group, void* stream) {
    return upsample_wgrad(mel, g_cond, dw, db, batch, n_mel, frames, T, ld, ksize, stride, n_group, S(stream));
}
WGB_API int wgb_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                          float eps, int step, float grad_scale, void* stream) {
    return adam_step(p, g, m, v, n, lr, beta1, beta2, eps, step, grad_scale, S(stream));
}
WGB_API int wgb_tc2_wn_res_taps(const void* a, const void* w, const float* bias, const void* h_in, void* h_out, int batch,
                                int T, long long h_batch_rows, int C, int taps, int dilation, void* stream) {
    return tc2_wn_res_taps(a, w, bias, h_in, h_out, batch, T, h_batch_rows, C, taps, dilation, S(stream));
}
WGB_API int wgb_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                              float eps, int* step_dev, float grad_scale, void* stream) {
    return adam_step_dev(p, g, m, v, n, lr, beta1, beta2, eps, step_dev, grad_scale, S(stream));
}
WGB_API int wgb_logdet(const float* w, float* out, float* inv_t, int c, void* stream) {
    return logdet(w, out, inv_t, c, S(stream));
}
WGB_API int wgb_tc2_wn_res_seg(const void* a0, const void* a1, int n_seg, int seg_mask, const void* w, const float* bias,
                               const void* h_in, void* h_out, int batch, int T, long long h_batch_rows, int C, int shift0,
                               int dshift, void* stream) {
    return tc2_wn_res_seg(a0, a1, n_seg, seg_mask, w, bias, h_in, h_out, batch, T, h_batch_rows, C, shift0, dshift, S(stream));
}
