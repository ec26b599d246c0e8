/* waveglow_b200 — C ABI of the B200-native vocoding path (libwaveglow_b200.so).
 *
 * The reference (DonggeunYu/Text2Speech) has no FFI: its hot path sits behind Python nn.Modules
 * (waveglow/glow.py, waveglow/denoiser.py, utils/stft.py, utils/layers.py).  The drop-in Python
 * classes in text2speech_b200/ keep that module API and state_dict layout and call ONLY the
 * functions below (ctypes), so this header is the whole device-side contract.  Each entry point
 * cites the reference code it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; wgb_last_error() then holds a
 *     human-readable message (thread-local).  No CPU fallback exists anywhere: a missing GPU or
 *     a non-sm_100 device is an error.
 *   - all pointers are DEVICE pointers (caller-allocated, caller-owned: torch owns memory);
 *     `stream` is a cudaStream_t passed as void*; calls are asynchronous on that stream.
 *   - activations are channels-last: h/acts [B, T, 512], cond [B, T, 640], flow state x [B, T, 8]
 *     (T = group steps = samples / 8).  "bf16" buffers hold __nv_bfloat16.
 *   - safe to call concurrently from different host threads on different devices/streams.
 */
#ifndef WAVEGLOW_B200_H_
#define WAVEGLOW_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define WGB_ABI_VERSION 1
#if defined(__GNUC__)
#define WGB_API __attribute__((visibility("default")))
#else
#define WGB_API
#endif

WGB_API int wgb_abi_version(void);
/* hex SHA-256 of the sources (csrc + this header) the library was built from: the Python binding compares it with
 * the tree it parses its prototypes from, so a stale library is rebuilt (or refused) instead of being called through
 * newer prototypes. */
WGB_API const char* wgb_source_hash(void);
WGB_API const char* wgb_last_error(void);
/* Process-wide tuning switches, A/B-measured with tools/bench_kernels.py; results never depend on them.
 *   "gate_l2_hint"  gate kernels (wgb_tc2_wn_gate, wgb_tc2_wn_gate_mel*): bit 0 = weight tiles TMA-loaded with L2
 *                   evict_last priority (DEFAULT 1: 5.27 vs 5.58 ms per launch at 64 x 10 s, the weights no longer lose
 *                   their L2 lines to the activations streaming through; profiles/r02k_l2_hint_ab.json), bit 1 = h taps
 *                   with evict_first (slower: 6.0 ms), bit 2 = mel_stack evict_last, bit 3 = acts stores evict_first
 *   "res_l2_hint"   wgb_tc2_wn_res*: bit 0 = weights evict_last, bit 1 = activations evict_first (default 0: 1.20 ms
 *                   either way for bit 0, 1.26 ms with bit 1)
 *   "stft_l2_hint"  wgb_tc2_stft_* / wgb_tc2_istft_ola / wgb_tc2_gemm_split3: bit 0 = basis tiles evict_last (default 0:
 *                   no effect, profiles/r02l_stft_l2_hint_ab.json)
 *   "pdl"           1 (default): the persistent WN kernels (wgb_tc2_wn_*, wgb_tc_wn_skip16_end, wgb_x_stack) are launched
 *                   with programmatic stream serialization: a kernel's CTAs start on an SM as soon as the previous kernel's
 *                   CTA there has exited and run their prologue while its slowest CTAs finish; they wait for the previous
 *                   grid's completion before their first access to activations.  0: plain stream order.
 *   "fft_mel_warps" wgb_fft_stft_mel: 12 (default; 168 registers per thread) or 16 (128 registers, small spills) warps
 *                   per CTA, one CTA per SM */
WGB_API int wgb_set_tuning(const char* key, int value);
/* 0 iff `device` exists and is compute capability 10.x; selects nothing. */
WGB_API int wgb_device_check(int device);

/* ---------------------------------------------------------------- flow state / small convs */

/* x[b,t,c] = sigma * z[b,c,t].  Replaces the noise draws + `sigma*audio` of WaveGlow.infer
 * (glow.py:260-269, :284-289): z [B,8,T] is host-supplied noise in WaveGlow.forward's output layout. */
WGB_API int wgb_flow_from_z(const float* z, float* x, int batch, int T, float sigma, void* stream);
/* z[b,c,t] = x[b,t,c]: the cat(output_audio, 1) of WaveGlow.forward (glow.py:248-249). */
WGB_API int wgb_flow_to_z(const float* x, float* z, int batch, int T, void* stream);
/* x[:, 8-C:] <- W x[:, 8-C:], W as fp32 [8][8] row-major (top-left CxC): Invertible1x1Conv.forward
 * (glow.py:100-101).  log|det W| is computed on the host. */
WGB_API int wgb_flow_mix(float* x, const float* w, long long rows, int C, void* stream);
/* WN.start 1x1 conv (glow.py:122-124,156): h[r,:] = W[512,n_half] x[r, 8-2*n_half : 8-n_half] + b.
 * out_bf16 selects a bf16 (tensor-core path) or fp32 (validation path) h. */
WGB_API int wgb_wn_start(const float* x, const float* w, const float* bias, void* h, int out_bf16, long long rows,
                 int n_ch, int n_half, void* stream);
/* Same, x [B,T,8] dense but h with a row pitch of h_batch_rows >= T per utterance (the padded layout of
 * wgb_tc2_wn_gate_mel: rows T.. of every utterance are guard rows that stay zero). */
WGB_API int wgb_wn_start_padded(const float* x, const float* w, const float* bias, void* h, int out_bf16, int batch, int T,
                                long long h_batch_rows, int n_ch, int n_half, void* stream);

/* ---------------------------------------------------------------- WN layers, BF16 tensor-core path */

/* in_layers[i] (k=3, dilation) + cond_layers[i] + fused_add_tanh_sigmoid_multiply
 * (glow.py:33-40, :159-162).  h bf16 [B,T,512], cond bf16 [B,T,640] -> acts bf16 [B,T,512].
 * w_packed bf16 [1024][2176] and bias fp32 [1024] in the packed row order documented in
 * text2speech_b200/packing.py (pass p: tanh rows 128p.., then sigmoid rows 512+128p..). */
WGB_API int wgb_tc_wn_gate(const void* h, const void* cond, const void* w_packed, const float* bias, void* acts,
                   int batch, int T, int dilation, void* stream);
/* Same contract as wgb_tc_wn_gate, executed by CTA pairs (tcgen05.mma.cta_group::2, UMMA M = 256): the two
 * CTAs of a cluster share each weight tile, halving weight traffic from L2 and shared memory. */
WGB_API int wgb_tc2_wn_gate(const void* h, const void* cond, const void* w_packed, const float* bias, void* acts,
                            int batch, int T, int dilation, void* stream);
/* residual half of res_skip_layers[i] (glow.py:164-166): h_out = h_in + W_res[512][512] acts + b. */
WGB_API int wgb_tc_wn_res(const void* acts, const void* w_res, const float* bias, const void* h_in, void* h_out,
                  int batch, int T, void* stream);
/* skip half of all res_skip_layers (glow.py:167-174) as one K = n_layers*512 GEMM over the stored
 * acts_all bf16 [n_layers][B][T][512], then WN.end (glow.py:175), the affine coupling and — for
 * direction 0 (infer, glow.py:277-282) — the inverse 1x1 conv W^-1 (w_mix fp32 [8][8]); direction 1
 * (forward, glow.py:241-246) writes log_s fp32 [B,n_half,T].  w_skip bf16 [512][n_layers*512],
 * w_end fp32 [512][8] (transposed, zero padded), b_end fp32 [8] with the skip biases folded in. */
WGB_API int wgb_tc_wn_skip_end(const void* acts_all, int n_layers, const void* w_skip, const float* w_end,
                       const float* b_end, float* x, const float* w_mix, float* log_s, int batch, int T,
                       int n_half, int direction, void* stream);

/* The gate layer with cond_layers[i] composed with WaveGlow.upsample (glow.py:183-185,252-258 + :141-143,161):
 * cond_layers[i](regroup(upsample(mel)))[t] is linear in the 4 x 80 mel values reaching group step t, with a
 * weight that depends only on t mod 32, so the kernel tiles rows phase-major (128 frames at one phase) and the
 * conditioning costs K = 320 instead of 640 (K = 1856 per layer instead of 2176); the [B,T,640] cond tensor is
 * not needed.  h bf16 [B,T,512], T = 32 * frames; mel_stack bf16 [B,frames,320] (wgb_upsample_im2col with
 * ld_tap = 80); w_packed as for wgb_tc_wn_gate (its 640 cond columns are not read); w_mel bf16 [32][1024][320] and
 * bias fp32 [1024] from text2speech_b200/packing.py:pack_cond_mel.  Same output as wgb_tc2_wn_gate up to bf16
 * rounding of the composed weight.
 *
 * frames_pad == T/32: h and mel_stack are dense and every utterance is tiled on its own (ceil(frames/128) tiles per
 * phase).  frames_pad > T/32 (by at least dilation/32): PADDED layout: h [B, 32*frames_pad, 512] and mel_stack
 * [B, frames_pad, 320], the guard rows of h (rows >= T of each utterance) are zero and stay zero, and the kernel tiles
 * all utterances as one sequence of B*frames_pad frames (no partial tile per utterance); acts stays dense [B,T,512].
 *
 * Optional skip accumulation (w_comp, skip_acc both non-NULL): the skip halves of res_skip_layers and WN.end are
 * linear with nothing in between (glow.py:167-175), so this layer's contribution to WN.end's output is
 * (W_end W_skip_i) acts_i.  w_comp fp32 [512][8] is that product transposed; each epilogue thread adds (stores, when
 * skip_first) the partial product over its pass's 128 channels, taken from the un-rounded fp32 activations, into its
 * own slot of skip_acc fp32 [4][B*T][8].  wgb_end_from_acc then finishes the flow, and acts need not be kept. */
WGB_API int wgb_tc2_wn_gate_mel(const void* h, const void* mel_stack, const void* w_packed, const void* w_mel,
                                const float* bias, void* acts, int batch, int T, int frames_pad, int dilation,
                                const float* w_comp, float* skip_acc, int skip_first, void* stream);
/* First WN layer with WN.start folded into in_layers[0] (both linear, glow.py:156 + :160, dilation 1):
 *   in_layers[0](start(a0))[t] = sum_tap (W_in0,tap W_start) a0[t+tap-1] + sum_tap (W_in0,tap b_start) [t+tap-1 in range]
 * wgb_x_stack writes, per group step, the K = 64 operand row (bf16 hi / lo split of the three a0 taps of the flow state
 * x [B,T,8] and the in-range indicators; layout in csrc/flow.cu) into out bf16 [B, out_batch_rows, 64];
 * wgb_tc2_wn_gate_mel0 is wgb_tc2_wn_gate_mel with that operand (x_stack, w0 bf16 [1024][64] from
 * packing.py:pack_gate0) in place of the three dilated taps of h: K = 64 + 320 instead of 1536 + 320.  The first
 * layer's in-conv becomes fp32-grade accurate (hi/lo on both sides) instead of reading a bf16-rounded h. */
WGB_API int wgb_x_stack(const float* x, void* out, int batch, int T, long long out_batch_rows, int n_half, void* stream);
WGB_API int wgb_tc2_wn_gate_mel0(const void* x_stack, const void* mel_stack, const void* w0, const void* w_mel,
                                 const float* bias, void* acts, int batch, int T, int frames_pad, const float* w_comp,
                                 float* skip_acc, int skip_first, void* stream);
/* WN.end output = sum of the 4 slots of skip_acc + b_end (skip biases folded in), then the affine coupling and W^-1
 * (direction 0, glow.py:277-282) or forward coupling + log_s (direction 1, :241-246), and optionally WN.start of the
 * next flow (as wgb_tc2_wn_skip_end).  Memory-bound: 160 B in, 32 B (+ 1 KB h_next) out per group step. */
WGB_API int wgb_end_from_acc(const float* skip_acc, const float* b_end, float* x, const float* w_mix, float* log_s,
                             int batch, int T, int n_half, int direction, const float* next_w_start,
                             const float* next_b_start, int next_n_half, void* h_next, long long h_next_batch_rows,
                             void* stream);

/* CTA-pair (cta_group::2) forms of the two entry points above, same contracts: each CTA loads half of every
 * weight tile, the pair issues one M = 256 MMA (half the weight traffic from L2 / shared memory per FLOP).
 * wgb_tc2_wn_skip_end can also run WN.start of the NEXT flow of WaveGlow.infer (glow.py:156 for flow k-1) on the
 * rows it has just updated: h_next bf16 [B,T,512] (row pitch h_next_batch_rows >= T per utterance) = next_w_start
 * fp32 [512][next_n_half] applied to that flow's audio_0 channels + next_b_start fp32 [512]; pass h_next = NULL
 * (and NULL / 0 for the rest) to skip it.  wgb_tc2_wn_res takes the same row pitch for h_in / h_out.
 *
 * wgb_tc2_wn_res can also accumulate the layer's share of WN.end's output while the activations are on chip
 * (w16_layer, skip_acc both non-NULL): a third, N = 16 pass per tile against w16_layer bf16 [16][512] (hi rows 0..7 /
 * lo rows 8..15 of W_end W_skip_i), added (stored when skip_first) into skip_acc fp32 [B*T][8]. */
WGB_API int wgb_tc2_wn_res(const void* acts, const void* w_res, const float* bias, const void* h_in, void* h_out,
                           int batch, int T, long long h_batch_rows, const void* w16_layer, float* skip_acc,
                           int skip_first, void* stream);
WGB_API int wgb_tc2_wn_skip_end(const void* acts_all, int n_layers, const void* w_skip, const float* w_end,
                                const float* b_end, float* x, const float* w_mix, float* log_s, int batch, int T,
                                int n_half, int direction, const float* next_w_start, const float* next_b_start,
                                int next_n_half, void* h_next, long long h_next_batch_rows, void* stream);

/* Skip path and WN.end composed (glow.py:167-175 are linear with nothing in between): one skinny tcgen05 GEMM of
 * acts_all bf16 [n_layers][B][T][512] against w16 bf16 [16][n_layers*512] = the bf16 hi (rows 0..7) and lo (rows 8..15)
 * parts of W_end [W_skip_0 | ... | W_skip_7] (zero rows beyond 2*n_half), b_end fp32 [8] with the skip biases folded
 * in; then the same coupling / W^-1 / log_s / optional next-flow WN.start epilogue as wgb_tc2_wn_skip_end.
 * HBM-bound: 8 KB of activations per group step, 2*4096*16 tensor FLOPs instead of 2*4096*512.
 * skip_acc (optional) fp32 [B*T][8] is added to the product: with the first n-1 layers accumulated by
 * wgb_tc2_wn_res, call this with n_layers = 1 on the last layer's activations and its [16][512] weight slice.
 * next_w_mix (direction 1 only, with a fused next-flow start): fp32 [8][8] W of the NEXT flow's Invertible1x1Conv,
 * applied to the updated row first (WaveGlow.forward runs convinv[k+1] before WN[k+1], glow.py:233). */
WGB_API int wgb_tc_wn_skip16_end(const void* acts_all, int n_layers, const void* w16, const float* b_end, float* x,
                                 const float* w_mix, float* log_s, int batch, int T, int n_half, int direction,
                                 const float* next_w_start, const float* next_b_start, int next_n_half, void* h_next,
                                 long long h_next_batch_rows, const float* skip_acc, const float* next_w_mix,
                                 void* stream);

/* Plain tcgen05 GEMM with the same TMA/TMEM pipeline: C[b,t,n] = sum_k A[b,t,k] W[n,k] + bias[n];
 * A bf16 [B,T,K] (K % 64 == 0), W bf16 [N,K] (N % 256 == 0), C fp32 or bf16 [B,T,N]; bias may be NULL.
 * Serves the dense-basis contractions that are 1x1 convs in the reference (cond/upsample/STFT bases,
 * glow.py:141-143,183-185; stft.py:85-89) and the exact-arithmetic unit tests of the pipeline. */
WGB_API int wgb_tc_gemm(const void* a, const void* w, const float* bias, void* c, int out_bf16, int batch, int T,
                        int N, int K, void* stream);

/* conv1d (stride 1, zero "same" padding) as an implicit GEMM on the same pipeline:
 *   c[b,t,n] = act(bias[n] + sum_tap sum_ch w[n][tap*C + ch] * a[b, t + (tap - (taps-1)/2)*dilation, ch])
 * a bf16 channels-last [B,T,C] (C % 64 == 0), w bf16 [N][taps*C] (N % 256 == 0), c fp32 or bf16 [B,T,N]; act 0 none,
 * 1 tanh, 2 relu.  Every tap is a K segment whose TMA box is shifted in time (out-of-range rows zero-filled).  Serves
 * the Tacotron-2 Postnet (tacotron/modules.py:94-137) with BatchNorm folded into w / bias. */
WGB_API int wgb_tc_conv1d(const void* a, const void* w, const float* bias, void* c, int out_bf16, int batch, int T, int N,
                          int C, int taps, int dilation, int act, void* stream);

/* Split-bf16 GEMM (fp32-grade accuracy on tcgen05): C[b,r,n] = sum_k A[b,r,k] W[n,k] with A = a_hi + a_lo
 * (bf16 parts, row r of batch b at element offset b*batch_stride + r*row_stride, rows may OVERLAP:
 * row_stride = hop < K reads STFT frames straight from the padded signal) and w3 = [W_hi | W_hi | W_lo]
 * bf16 [N][3K]; fp32 C [B,rows,N].  The dense-basis contractions of STFT.transform / STFT.inverse
 * (stft.py:85-89, :105-109).  N % 256 == 0, K % 64 == 0, strides % 8 == 0. */
WGB_API int wgb_tc_gemm_split3(const void* a_hi, const void* a_lo, const void* w3, const float* bias, void* c,
                               int batch, int rows, int N, int K, long long row_stride, long long batch_stride,
                               void* stream);

/* The forward STFT GEMM with its consumer fused into the epilogue.  w3_paired = the split-bf16 forward basis
 * [2cp][3K] with rows in Re/Im-PAIRED order (rows 256p..256p+127 = Re of bins 128p..128p+127, the next 128 rows their
 * Im; cp % 128 == 0), so each accumulator pass holds both parts of its bins:
 *   wgb_tc_stft_mag      |X| fp32 channels-last [B, rows, cp]   (stft.py:85-97; all TacotronSTFT.mel_spectrogram needs)
 *   wgb_tc_stft_denoise  Denoiser.forward's max(|X| - bias*strength, 0) e^{j arg X} (denoiser.py:36-38, stft.py:102-103)
 *                        written as the bf16 hi / lo operands [B*rows][2cp] (Re | Im) of the inverse-basis GEMM. */
/*   wgb_tc_stft_mel      all of TacotronSTFT.mel_spectrogram (layers.py:63-79): |X|, the mel filterbank and
 *                        log(max(., clip)) in the epilogue; out fp32 [B, n_mel, rows].  mel_table: cp x {first filter
 *                        index, weight in it, weight in the next filter, 0} (fp32 x 4): the triangular filters overlap
 *                        pairwise, so a bin feeds at most two adjacent filters (checked by the packer). */
WGB_API int wgb_tc_stft_mel(const void* a_hi, const void* a_lo, const void* w3_paired, const void* mel_table, float* out,
                            int batch, int rows, int cp, int K, long long row_stride, long long batch_stride, int n_mel,
                            float clip, void* stream);
WGB_API int wgb_tc_stft_mag(const void* a_hi, const void* a_lo, const void* w3_paired, void* mag_cl, int batch, int rows,
                            int cp, int K, long long row_stride, long long batch_stride, void* stream);
WGB_API int wgb_tc_stft_denoise(const void* a_hi, const void* a_lo, const void* w3_paired, const float* bias_spec,
                                float strength, void* hi_out, void* lo_out, int batch, int rows, int cutoff, int cp, int K,
                                long long row_stride, long long batch_stride, void* stream);

/* The same three STFT-family GEMMs on CTA pairs (tcgen05 cta_group::2, 5-6 stage ring, half the basis traffic per
 * CTA) with an unpadded spectrum layout -- the defaults of TacotronSTFT.mel_spectrogram / Denoiser.forward when
 * filter_length % 256 == 0 and hop % 8 == 0.  Im of bins 0 and L/2 is exactly zero in the reference's basis
 * (stft.py:46-51), so the L/2 + 1 bins are L real numbers: w3_paired = split-bf16 forward basis [L][3L] whose pass p
 * (256 rows) holds the Re rows of bins 128p..128p+127 followed by their Im rows, with the Re row of bin L/2 in the Im
 * slot of bin 0 (L/256 passes instead of the five of the padded layout above).  a_hi / a_lo = reflect-padded signals
 * as bf16 hi / lo parts [B, R*hop]: frame r of utterance b is flat row b*R + r of ONE frame axis over the batch.
 *   wgb_tc2_stft_mel      layers.py:63-79 in one kernel; only the first n_pass passes run (the ones holding a bin with
 *                         non-zero mel weight: 3 of 4 at 22.05 kHz / fmax 8 kHz); mel_table has L/2 + 1 entries
 *                         (the last one = bin L/2); out fp32 [B, n_mel, frames].
 *   wgb_tc2_stft_denoise  denoiser.py:36-38 + stft.py:102-103; hi_out / lo_out bf16 [B][out_R][L] (out_R >= frames: row
 *                         pitch per utterance; rows >= frames are not written): columns 0..L/2-1
 *                         Re of bins 0..L/2-1, column L/2 Re of bin L/2, columns L/2+1.. Im of bins 1..L/2-1 = the
 *                         K operand of the inverse-basis GEMM (K = L instead of 2 * 640); bias_spec fp32 [L/2 + 1].
 *   wgb_tc2_gemm_split3   C fp32 [rows][N] = (a_hi + a_lo)[rows][K] (W_hi + W_lo)^T, w3 = [W_hi | W_hi | W_lo] bf16
 *                         [N][3K] (stft.py:105-109, the inverse-basis contraction); N % 256 == 0, K % 64 == 0. */
WGB_API int wgb_tc2_stft_mel(const void* a_hi, const void* a_lo, const void* w3_paired, const void* mel_table, float* out,
                             int batch, int frames, int R, int L, int hop, int n_pass, int n_mel, float clip, void* stream);
WGB_API int wgb_tc2_stft_denoise(const void* a_hi, const void* a_lo, const void* w3_paired, const float* bias_spec,
                                 float strength, void* hi_out, void* lo_out, int batch, int frames, int R, int L, int hop,
                                 int out_R, void* stream);
WGB_API int wgb_tc2_gemm_split3(const void* a_hi, const void* a_lo, const void* w3, float* c, long long rows, int N, int K,
                                void* stream);
/* STFT.inverse's conv_transpose1d WITH its overlap-add, the window-sum normalisation, the L/hop scale and the L/2 trim
 * (stft.py:105-128; audio_processing.py:7-48) as one CTA-pair GEMM: output block q (hop samples) = sum_j frame[q-j] W_j
 * with W_j the j-th hop-wide slice of the inverse basis (a tap conv over the frame axis, K = taps * 3L, N = hop), so the
 * [B, frames, L] intermediate never exists.  s_hi / s_lo bf16 [B][frames + taps - 1][L] (column order of
 * wgb_tc2_stft_denoise; the last taps - 1 rows of every utterance ZERO), taps = L / hop; w_ola bf16 [hop][taps * 3L]
 * (tap j: split-bf16 [hi | hi | lo] of inverse-basis samples j*hop .. (j+1)*hop-1); env_tab fp32 [2^taps][hop]: the
 * window-sum envelope for every set of covering frames (bit j: frame q - j exists), accumulated like the reference's
 * host loop, or NULL for window=None; out fp32 [B][hop * (frames - 1)].  hop % 256 == 0, L % hop == 0, L / hop <= 8. */
WGB_API int wgb_tc2_istft_ola(const void* s_hi, const void* s_lo, const void* w_ola, const float* env_tab, float* out,
                              int batch, int frames, int L, int hop, void* stream);

/* ---------------------------------------------------------------- butterfly STFT (stock bases, filter_length 1024)
 * The reference's forward_basis is window * [cos; -sin](2 pi k n / L) and inverse_basis its pseudo-inverse times the window
 * (stft.py:46-60): a real DFT pair.  For those bases the two fused paths below compute the same thing as 1024-point real
 * FFTs in fp32 (one warp per frame, 512-point complex FFT in registers + one shared-memory exchange buffer): 0.05 MFLOP per
 * frame instead of the 2.1 MFLOP dense contraction, the signal read once from HBM and only the result written.  The host
 * (stft.py / layers.py) takes these paths only when the module's basis buffers still equal the constructor's.
 *
 * TacotronSTFT.mel_spectrogram (layers.py:63-79; stft.py:79-97): y fp32 [B][n] -> out fp32 [B][n_mel][n / hop + 1] =
 * log(clamp(mel_basis |STFT(y)|, clip)).  window fp32 [1024] (zero-padded window, ones for window=None).  The filterbank
 * arrives as PIECES of 8 consecutive bins starting at a multiple of 4, dealt to the 32 lanes of a warp: lane l sums the
 * pieces [q][l], q = 0 .. slots_per_lane - 1, in order.  mel_slots int32 [slots_per_lane][32] = first bin / 4 | (filter to
 * emit after this piece + 1) << 8 | (1 if the piece starts a filter) << 16; mel_w fp32 [slots_per_lane][2][32][4] = the
 * piece's 8 weights as two float4 with the lanes innermost (zero outside the filter's span and past bin 512).  All pieces
 * of a filter sit consecutively in ONE lane, so every filter is summed in a fixed order by one thread; unused slots carry
 * zero weights and emit nothing.  bins_used = one past the last bin any piece with weight covers (513 = all; at most 384
 * lets the kernel skip the upper quarter of the spectrum, which the usual 8 kHz filterbank never reads).  range_flag
 * (optional int32) is set to 1 when a sample is outside [-1, 1] or NaN (layers.py:72-73).  n > 512, n_mel <= 128,
 * slots_per_lane <= 16. */
WGB_API int wgb_fft_stft_mel(const float* y, const float* window, const void* mel_slots, int slots_per_lane,
                             const float* mel_w, int bins_used, float* out, int batch, int n, int hop, int n_mel,
                             float clip, int* range_flag, void* stream);
/* Denoiser.forward (denoiser.py:35-40) = STFT.transform, clamp(|X| - bias_spec * strength, 0) with the phase kept,
 * STFT.inverse (stft.py:99-130: overlap-add, window-sum normalisation, L/hop scale, L/2 trim) in ONE kernel: a warp walks a
 * run of consecutive frames and holds the overlap-add in registers.  y fp32 [B][n]; bias_spec fp32 [513]; env_tab fp32
 * [16][hop] as for wgb_tc2_istft_ola (NULL for window=None: no normalisation and no scale); out fp32 [B][hop * (n / hop)].
 * hop must be 256 (= L / 4, the Denoiser's n_overlap = 4). */
WGB_API int wgb_fft_denoise(const float* y, const float* window, const float* bias_spec, float strength,
                            const float* env_tab, float* out, int batch, int n, int hop, void* stream);

/* ---------------------------------------------------------------- FP32 validation path (CUDA cores) */

/* C[b][m][n] (+)= sum_k A[b][m+shift][k] W[n][k] + bias[n]; rows outside [0,M) read as zero, so a
 * dilated conv tap is a GEMM with a row shift (in_layers, glow.py:136-139,160), a 1x1 conv is
 * shift 0 (cond/res_skip/end), the STFT is A = overlapping frames (lda = hop; stft.py:85-89).
 * K, lda, ldw, a_batch must be multiples of 4.  out_bf16: C is bf16 (no accumulate). */
WGB_API int wgb_sgemm_f32(const float* A, const float* W, const float* bias, void* C, int out_bf16, int batch, int M,
                  int N, int K, long long lda, long long a_batch, long long ldw, long long ldc,
                  long long c_batch, int shift, int accumulate, void* stream);
/* acts = tanh(u[:, :C]) * sigmoid(u[:, C:]) with accurate tanhf/expf (glow.py:33-40). */
WGB_API int wgb_gate_f32(const float* u, float* acts, long long rows, int n_ch, void* stream);
/* fused_add_tanh_sigmoid_multiply(input_a, input_b, n_channels) of the reference on ITS layout (glow.py:33-40):
 * input_a, input_b fp32 [B, 2*n_ch, T] channels-first -> acts fp32 [B, n_ch, T]. */
WGB_API int wgb_fused_add_tanh_sigmoid_multiply(const float* input_a, const float* input_b, float* acts, int batch,
                                                int n_ch, int T, void* stream);
/* in-place tanh (act 1) / relu (act 2), accurate libm versions: FP32 validation path of the Postnet. */
WGB_API int wgb_act_f32(float* x, long long n, int act, void* stream);
/* has_res: h += rs[:, :C]; skip (+)= rs[:, C:]   else: skip (+)= rs   (glow.py:165-174). */
WGB_API int wgb_res_skip_f32(const float* rs, float* h, float* skip, long long rows, int n_ch, int has_res,
                     int first, void* stream);
/* WN.end + coupling (+ W^-1 for infer) from an fp32 skip sum [B*T, n_ch]; same math as the
 * epilogue of wgb_tc_wn_skip_end. */
WGB_API int wgb_end_coupling_f32(const float* skip, const float* w_end, const float* b_end, float* x,
                         const float* w_mix, float* log_s, int batch, int T, int n_ch, int n_half,
                         int direction, void* stream);

/* ---------------------------------------------------------------- upsample (ConvTranspose1d as GEMM) */

/* A[b,q,j*ld_tap+c] = mel[b,c,q-j] (zero if q<j or c>=n_mel): the four frames feeding output samples
 * [256q, 256q+256) of WaveGlow.upsample (glow.py:183-185,252); the GEMM against the repacked
 * [20480][taps*ld_tap] weight then writes the regrouped cond [B, 32F, 640] directly (glow.py:254-258). */
WGB_API int wgb_upsample_im2col(const float* mel, void* a, int out_bf16, int batch, int n_mel, int F, int taps,
                        int ld_tap, void* stream);
WGB_API int wgb_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream);

/* pcm[i] = (int16) trunc(audio[i] * scale), saturating: the `audio * MAX_WAV_VALUE` -> astype('int16') of the
 * vocoder CLI (waveglow/inference.py:58-62) done before the device->host copy.  Buffers 16 B aligned. */
WGB_API int wgb_audio_to_int16(const float* audio, void* pcm, long long n, float scale, void* stream);

/* ---------------------------------------------------------------- STFT / mel / denoiser glue */

/* reflect pad by `half` each side into ypad[B, ld_pad] (stft.py:79-83). */
WGB_API int wgb_stft_reflect_pad(const float* y, float* ypad, int batch, int N, int half, long long ld_pad, void* stream);
/* reflect pad + split into bf16 hi/lo parts (operands of wgb_tc_gemm_split3); ld_pad % 8 == 0. */
WGB_API int wgb_stft_reflect_pad_split(const float* y, void* hi, void* lo, int batch, int N, int half,
                                       long long ld_pad, void* stream);
/* the same pass with the input-range asserts of TacotronSTFT.mel_spectrogram (layers.py:72-73: min(y) >= -1, max(y) <= 1)
 * folded in: *range_flag (device int, zeroed by the caller) is set to 1 when a sample is outside [-1, 1] or NaN. */
WGB_API int wgb_stft_reflect_pad_split_check(const float* y, void* hi, void* lo, int batch, int N, int half,
                                             long long ld_pad, int* range_flag, void* stream);
/* hi = bf16(src), lo = bf16(src - hi). */
WGB_API int wgb_split_bf16(const float* src, void* hi, void* lo, long long n, void* stream);
/* spec[B,F,2cp] -> magnitude/phase [B,cutoff,F] (stft.py:91-97), optional channels-last mag_cl[B,F,cp];
 * any of mag / phase / mag_cl may be NULL. */
WGB_API int wgb_stft_polar(const float* spec, float* mag, float* phase, float* mag_cl, int batch, int F, int cutoff,
                   int cp, void* stream);
/* out[b,m,f] = log(max(raw[b,f,m], clip))  (layers.py:77-78; audio_processing.py:70-76). */
WGB_API int wgb_mel_log(const float* raw, float* out, int batch, int F, int n_mel, float clip, void* stream);
/* in-place spectral subtraction on spec rows (denoiser.py:36-38 + the cos/sin recombination of
 * stft.py:102-103, done as a magnitude ratio). */
WGB_API int wgb_denoise_scale(float* spec, const float* bias, float strength, long long rows, int cutoff, int cp,
                      void* stream);
/* The same spectral subtraction, but writing the result as the bf16 hi / lo operands of the inverse-basis
 * wgb_tc_gemm_split3 (spec is read once and not modified). */
WGB_API int wgb_denoise_scale_split(const float* spec, const float* bias, float strength, void* hi, void* lo,
                                    long long rows, int cutoff, int cp, void* stream);
/* Griffin-Lim projection step (audio_processing.py:64-66): keep the phase of spec[B,F,2cp], impose the
 * magnitude target[B,cutoff,F] (in place, no atan2/cos/sin: Re,Im *= target/|X|; |X| = 0 -> phase 0). */
WGB_API int wgb_spec_set_magnitude(float* spec, const float* target, int batch, int F, int cutoff, int cp, void* stream);
/* (magnitude, phase) [B,cutoff,F] -> spec[B,F,2cp]  (stft.py:102-103). */
WGB_API int wgb_stft_recombine(const float* mag, const float* phase, float* spec, int batch, int F, int cutoff, int cp,
                       void* stream);
/* overlap-add + window-sum normalisation + xL/hop + trim (stft.py:105-128; audio_processing.py:7-48):
 * frames [B,F,L] -> out [B, hop*(F-1)]; win_sq = fp64 squared padded window [L], or NULL for the reference's
 * window=None case (no envelope division and no L/hop scale, stft.py:111-125). */
WGB_API int wgb_istft_overlap_add(const float* frames, const double* win_sq, float* out, int batch, int F, int L,
                          int hop, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Training direction: what torch autograd + torch.optim.Adam run for
 * waveglow/train.py:116-124 (`outputs = model((mel, audio)); loss = criterion(outputs); loss.backward();
 * optimizer.step()`), restated as explicit kernels over the tensors the forward kernels already hold.
 * All activations channels-last; "rows" = B*T group steps.
 * --------------------------------------------------------------------------------------------------------------- */
/* wgb_tc2_wn_gate that also stores tanh | sigmoid of the pre-activations as bf16 [B,T,1024] (original channel order):
 * the tensors autograd would save for glow.py:33-40. */
WGB_API int wgb_tc2_wn_gate_train(const void* h, const void* cond, const void* w_packed, const float* bias, void* acts,
                                  void* ts, int batch, int T, int dilation, void* stream);
/* out[b,t,n] = act(bias[n] + sum_s sum_c W[n][s*C+c] A_s[b, t + shift0 + s*dshift, c]) + res[b,t,n] on tcgen05.
 * A_s = a1 if bit s of seg_mask else a0 (bf16 [B,T,C], C % 64 == 0); stacked = 1: a0 is [n_seg,B,T,C] and A_s = plane s.
 * W bf16 [N][n_seg*C] (N % 256 == 0); out / res fp32 (out_bf16 = 0) or bf16; res may be NULL or alias out.  The data gradients of in_layers (dilated taps + residual
 * stream), res_skip_layers (segments [g_h | g_skip]) and cond_layers (accumulating) of glow.py:159-166. */
WGB_API int wgb_tc_gemm_seg(const void* a0, const void* a1, int n_seg, int seg_mask, const void* w, const float* bias,
                            const void* res, void* c, int out_bf16, int batch, int T, int N, int C, int shift0, int dshift,
                            int act, int stacked, void* stream);
/* h_out[b,t,:] = h_in[b,t,:] + bias + sum_tap W[:, tap*C:(tap+1)*C] a[b, t + (tap - (taps-1)/2) dilation, :] with the CTA-pair
 * residual kernel (cta_group::2, h tile updated in shared memory and TMA-stored): the data gradient of in_layers
 * flowing into the residual stream's gradient.  a bf16 [B,T,C], w bf16 [512][taps*C], bias fp32 [512], h bf16
 * [B,h_batch_rows,512] (h_in may alias h_out). */
WGB_API int wgb_tc2_wn_res_taps(const void* a, const void* w, const float* bias, const void* h_in, void* h_out, int batch,
                                int T, long long h_batch_rows, int C, int taps, int dilation, void* stream);
/* The same kernel with a general segmented K: h_out = h_in + bias + sum_s W[:, s*C:(s+1)*C] A_s[b, t + shift0 + s*dshift, :],
 * A_s = a1 if bit s of seg_mask else a0.  The res_skip data gradient g_acts = [g_h | g_skip] W_rs runs through it with
 * h_in = zeros. */
WGB_API int wgb_tc2_wn_res_seg(const void* a0, const void* a1, int n_seg, int seg_mask, const void* w, const float* bias,
                               const void* h_in, void* h_out, int batch, int T, long long h_batch_rows, int C, int shift0,
                               int dshift, void* stream);
/* dw[tap][m][n] (+)= sum_{b,t} g[b,t,m] x[b, t + (tap - (taps-1)/2)*dilation, n] on tcgen05 with both operands
 * MN-major (no transposed copies).  g bf16 [B,T,ca] (ca % 64 == 0), x bf16 [B,T,cb] (cb % 8 == 0), dw fp32
 * [taps][ca][cb]; accumulate = 0 clears dw first.  Weight gradients of in_layers / cond_layers / res_skip_layers. */
WGB_API int wgb_tc_wgrad(const void* g, const void* x, float* dw, int batch, int T, int ca, int cb, int taps, int dilation,
                         int accumulate, void* stream);
/* ts (tanh | sigmoid, bf16 [rows, 2 n_ch]) <- gradient w.r.t. the gate pre-activations given g_acts bf16 [rows, n_ch];
 * db (optional, fp32 [2 n_ch]) <- its column sums = the gradient of the in_layers / cond_layers biases. */
WGB_API int wgb_gate_bwd(const void* g_acts, void* ts, float* db, long long rows, int n_ch, void* stream);
/* Affine coupling + WN.end backward (glow.py:241-246): see csrc/train.cu.  g_x fp32 [rows,8] in/out, x_mix = flow state
 * before the coupling, log_s / g_log_s fp32 [B,n_half,T] (g_log_s may be NULL), w_end_t fp32 [n_ch][8];
 * g_out fp32 [rows,8], g_skip bf16 [rows,n_ch]; stack (optional) bf16 [rows,64] = hi/lo(g_out) | hi/lo(x_mix) | 1 | 0:
 * the operand that lets wgb_tc_wgrad compute the <= 8-channel weight gradients and the bias column sums. */
WGB_API int wgb_coupling_bwd(float* g_x, const float* x_mix, const float* log_s, const float* g_log_s, const float* w_end_t,
                             float* g_out, void* g_skip, void* stack, int batch, int T, int n_ch, int n_half, void* stream);
/* g_x[a0 channels] += g_h0 W_start (glow.py:156); w_start fp32 [n_ch][n_half]. */
WGB_API int wgb_start_bwd(float* g_x, const void* g_h0, const float* w_start, long long rows, int n_ch, int n_half,
                          void* stream);
/* out[j] (+)= sum_r a[r][j]; a fp32 [rows,8]. */
WGB_API int wgb_colsum8_f32(const float* a, float* out, long long rows, int accumulate, void* stream);
/* Invertible 1x1 conv backward (glow.py:97-102): g_x <- W^T g_y on the last C channels (in place),
 * dw[i][j] = sum_r g_y[i] x_pre[j] (fp32 [8][8]; the log-det term is the caller's). */
WGB_API int wgb_mix_bwd(float* g_x, const float* x_pre, const float* w, float* dw, long long rows, int C, void* stream);
/* ConvTranspose1d upsample weight / bias gradient (glow.py:183-185,213-221) from the gradient of the regrouped
 * conditioning tensor g_cond fp32 [B,T,ld]; mel fp32 [B,n_mel,frames]; dw fp32 [n_mel][n_mel][ksize], db [n_mel]. */
WGB_API int wgb_upsample_wgrad(const float* mel, const float* g_cond, float* dw, float* db, int batch, int n_mel, int frames,
                               int T, int ld, int ksize, int stride, int n_group, void* stream);
/* torch.optim.Adam step (train.py:79,124; no weight decay / amsgrad) over one flat fp32 buffer of n (% 4 == 0) values;
 * the gradient is multiplied by grad_scale first (1 / world_size after a sum all-reduce). */
WGB_API int wgb_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                          float eps, int step, float grad_scale, void* stream);
/* The same step with the step counter on the device (incremented by the call): CUDA-graph replayable. */
WGB_API int wgb_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                              float eps, int* step_dev, float grad_scale, void* stream);
/* torch.logdet(W) of a c x c matrix (c <= 8; NaN when det W < 0, as in the reference) and (W^-1)^T = its gradient
 * (glow.py:100): out[0], inv_t fp32 [c][c]. */
WGB_API int wgb_logdet(const float* w, float* out, float* inv_t, int c, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WAVEGLOW_B200_H_ */
