/* seedvc_b200 - C ABI of the B200 (sm_100a) kernels behind Seed-VC's conversion hot path.
 *
 * The reference has no FFI on this path except one JIT-built pybind op
 * (modules/bigvgan/alias_free_activation/cuda/anti_alias_activation.cpp:19-22,
 * `fwd_cuda` in anti_alias_activation_cuda.cu:212-246): raw device pointers in,
 * launch on the caller's stream, dtype dispatch, no allocation of inputs.  This
 * header keeps that shape for every kernel of the path: plain pointers, sizes and
 * a cudaStream_t (passed as void*), no torch types, no allocation, no host sync.
 * Each entry point names the reference site it replaces.  See INTEGRATION.md for
 * the ctypes binding.
 *
 * Return value: SVC_OK, or a negative SVC_ERR_* code with a message available
 * from svc_last_error().  All launches are asynchronous on `stream`.
 *
 * Layout convention: activations are "frames-major" (B, T, C): channel/feature
 * index contiguous.  A `rows`/`bstride`/`rstride` triple describes a strided
 * (B, rows, C) view in ELEMENTS of the tensor's own dtype.
 */
#ifndef SEEDVC_B200_H_
#define SEEDVC_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define SVC_OK 0
#define SVC_ERR_ARG (-1)
#define SVC_ERR_CUDA (-2)
#define SVC_ERR_UNSUPPORTED (-3)

/* operand element types */
#define SVC_BF16 0 /* bf16 operands, tcgen05 tensor cores, fp32 accumulate */
#define SVC_F32 1  /* fp32 operands, fp32 FFMA ("fp32 mode", also small-M GEMMs) */
#define SVC_F16 2  /* fp16 operands (the reference's default GPU precision: fp16 autocast, inference.py:499),
                      same tcgen05 kind::f16 path and speed as bf16, 11-bit mantissa */

/* svc_gemm backends */
#define SVC_BACKEND_AUTO 0 /* bf16 / fp16 -> tcgen05, fp32 -> SIMT */
#define SVC_BACKEND_SIMT 1 /* force the SIMT mainloop (debug cross-check) */

/* epilogue activations */
#define SVC_ACT_NONE 0
#define SVC_ACT_SILU 1
#define SVC_ACT_SWIGLU_PAIR 2  /* out[j] = silu(v[2j]) * v[2j+1]            (FeedForward) */
#define SVC_ACT_TANH_SIG_PAIR 3 /* out[j] = tanh(v[2j]) * sigmoid(v[2j+1])  (WaveNet gate) */
#define SVC_ACT_ROPE 4          /* rotate pairs (2j,2j+1) of columns < rope_cols       */

#define SVC_MAX_SEG 16

const char* svc_last_error(void);
int svc_version(void);

/* ---------------------------------------------------------------------------
 * svc_gemm: out[b,t,:] = epilogue( sum_s A_s[b, t + shift_s, 0:K_s] . W_s[:, 0:K_s]^T )
 *
 * One call covers every dense contraction of the path:
 *   nn.Linear          (diffusion_transformer.py:205-206,266-268,166,437,476-478,...)  1 segment
 *   concat + Linear    (diffusion_transformer.py:185-186, 524-525)  2 segments, no cat
 *   Conv1d, k taps     (bigvgan.py:56-87,285-287,348-350; wavenet.py:125-135)  k segments
 *   ConvTranspose1d    (bigvgan.py:300-316)  3 segments over the polyphase weight
 * Rows of A outside [0, a_rows) read as zero (conv zero padding / TMA OOB fill).
 * W_s is (N, K_s) row-major with row stride w_rstride.
 * Epilogue order: v = acc + bias[n] + rowbias[b][n]; act (pair acts halve N);
 *   v *= gate[b][n]; v += res[b,t,n]; v *= alpha; if (accumulate) v += out_f32[b,t,n];
 *   store out_f32 and/or out_op (operand dtype).
 * ------------------------------------------------------------------------- */
typedef struct svc_gemm_desc {
    int dtype;  /* SVC_BF16 / SVC_F16 / SVC_F32: element type of A, W and out_op */
    int B, T, N; /* output rows per batch, GEMM columns (before pair reduction) */
    int n_seg;
    const void* a_ptr[SVC_MAX_SEG];
    long long a_bstride[SVC_MAX_SEG];
    long long a_rstride[SVC_MAX_SEG];
    int a_rows[SVC_MAX_SEG];  /* valid rows per batch of the A view */
    int a_shift[SVC_MAX_SEG]; /* row offset added to t */
    const void* w_ptr[SVC_MAX_SEG];
    long long w_rstride[SVC_MAX_SEG];
    int K[SVC_MAX_SEG];
    /* epilogue */
    const float* bias;    /* [N] or NULL */
    const float* rowbias; /* [B][N] (stride rowbias_bstride) or NULL */
    long long rowbias_bstride;
    int act;
    const float* rope_tab; /* (pos, 32, 2) cos/sin, used when act == SVC_ACT_ROPE */
    int rope_cols;         /* columns [0, rope_cols) are rotated (q and k) */
    int rope_pos0;         /* position of row t is rope_pos0 + t */
    int q_cols;            /* columns [0, q_cols) are multiplied by q_scale after RoPE */
    float q_scale;
    const float* gate; /* [B][N_out] (stride gate_bstride) or NULL */
    long long gate_bstride;
    const float* res; /* fp32 (B, T, N_out) view or NULL */
    long long res_bstride, res_rstride;
    float alpha;
    int accumulate;
    float* out_f32; /* or NULL */
    long long of_bstride, of_rstride;
    void* out_op; /* operand dtype, or NULL */
    long long oo_bstride, oo_rstride;
    /* optional pair-major copy of rope_tab: (32 pairs, rope_ld positions, 2); lets the tensor-core
       path rotate in the accumulator's row layout with coalesced table reads.  NULL = not given. */
    const float* rope_tab_t;
    int rope_ld;
    /* element type of out_op when it differs from the operands': 0 = same as dtype, else 1 + SVC_BF16 / SVC_F16
       (a bf16 GEMM may write an fp16 operand copy and vice versa; fp32 GEMMs write fp32). */
    int out_op_dtype_p1;
    /* RMS normalisation folded into the GEMMs around it (AdaptiveLayerNorm / RMSNorm, diffusion_transformer.py:30-48):
       y = x * rsqrt(mean(x^2) + eps) * g + a  feeding a Linear W  ==  rsqrt(..) * (x (W * g)^T) + a W^T, so the
       caller folds g into the weight columns and a W^T into `bias`, the GEMM that PRODUCES x reports the row sums of
       squares and the GEMM that CONSUMES x scales its accumulator rows.
       row_ss_out: (B*T, SVC_SS_SLOTS) fp32, contiguous, zeroed once by the caller.  Two-output tensor-core calls
         (out_f32 + out_op) write, per output row, the sum of squares of the final fp32 values of every N tile into
         the leading slots (untouched slots keep their zeros).  N % 32 == 0, at most SVC_SS_SLOTS N tiles.
       row_ss_in: the same array from an earlier call: acc[r,:] *= rsqrt(sum(slots[r]) * rs_inv_dim + rs_eps) before
         the bias.  Tensor-core calls with bias + (RoPE | SwiGLU pair) -> out_op only.
       NULL = off.  Not available on the SIMT / fp32 path (SVC_ERR_UNSUPPORTED). */
    float* row_ss_out;
    const float* row_ss_in;
    float rs_inv_dim, rs_eps;
} svc_gemm_desc;
#define SVC_SS_SLOTS 8

int svc_gemm(const svc_gemm_desc* d, int backend, void* stream);

/* ---------------------------------------------------------------------------
 * svc_attention: non-causal softmax attention with key-length masking.
 * Replaces F.scaled_dot_product_attention + the (B,1,T,T) bool mask
 * (diffusion_transformer.py:255, 518-520).  q/k/v are (B, T, H*64) views with a
 * common bstride/rstride (column slices of the packed wqkv output); RoPE and the
 * 1/sqrt(64) scale are already applied to q,k by the wqkv epilogue.
 * kv_len[b] keys are attended for batch b (device int32 [B]).  head_dim is 64.
 * ------------------------------------------------------------------------- */
int svc_attention(const void* q, const void* k, const void* v, long long qkv_bstride,
                  long long qkv_rstride, void* out, long long out_bstride, long long out_rstride,
                  int B, int T, int H, const int* kv_len, int dtype, int backend, void* stream);

/* ---------------------------------------------------------------------------
 * svc_norm_mod: out = norm(x) * gamma * mul + add, per row.
 *   mode 0: RMSNorm (diffusion_transformer.py:274-285) as used by AdaptiveLayerNorm
 *           (:30-48; v2 dit_model.py:20-54,136-142); gamma = norm.weight.
 *   mode 1: LayerNorm without affine, eps 1e-6 (FinalLayer, :391,401-403).
 * gamma/mul/add are [D] fp32 vectors or NULL.  x is fp32 (B, T, D) strided; out is the
 * operand dtype.
 * ------------------------------------------------------------------------- */
int svc_norm_mod(const float* x, long long x_bstride, long long x_rstride, const float* gamma,
                 const float* mul, const float* add, float eps, int mode, void* out,
                 long long o_bstride, long long o_rstride, int B, int T, int D, int out_dtype,
                 void* stream);
/* Same, and additionally raw_out = x cast to the operand dtype (same strides as out): the U-ViT skip
 * tensor / the operand copy of the hidden state (diffusion_transformer.py:133-141,185-186) falls out of
 * the read the norm does anyway, so the producing GEMM needs only its fp32 output. */
int svc_norm_mod_copy(const float* x, long long x_bstride, long long x_rstride, const float* gamma,
                      const float* mul, const float* add, float eps, int mode, void* out, void* raw_out,
                      long long o_bstride, long long o_rstride, int B, int T, int D, int out_dtype,
                      int raw_dtype /* SVC_* of raw_out: out_dtype, or the other 16-bit format */, void* stream);

/* ---------------------------------------------------------------------------
 * svc_snake_aa: anti-aliased Snake / SnakeBeta = 2x Kaiser-sinc FIR upsample, x + sin^2(a x)/b,
 * FIR low-pass + 2x downsample.  Replaces Activation1d (alias_free_activation/torch/act.py:25-30,
 * resample.py:29-38, filter.py:94-101, activations.py:107-119) and the reference's own CUDA op
 * (alias_free_activation/cuda/anti_alias_activation_cuda.cu:44-179).
 * x: contiguous (B, L, C), fp32 / bf16 / fp16 (x_dtype); out: contiguous (B, L, C) out_dtype.
 * a[c] = exp(alpha_c) (or alpha_c), inv_b[c] = 1 / (beta_c + 1e-9), prepared by the caller.
 * precise = 1 (libm sinf, fp32 FIRs) needs fp32 in and out.  With a 16-bit out_dtype and C % 8 == 0 both
 * FIRs run on the tensor cores (x and the activated 2x signal enter them as IEEE half, taps split hi + lo).
 * ------------------------------------------------------------------------- */
int svc_snake_aa(const void* x, int x_dtype, void* out, int out_dtype, const float* a,
                 const float* inv_b, int B, int L, int C, int precise, void* stream);

/* activation_post + conv_post + clamp/tanh (bigvgan.py:377-384): out (B, L) fp32.
 * w is (ksize, C) fp32 (tap-major), bias NULL when use_bias_at_final is false. */
int svc_snake_conv_post(const float* x, const float* a, const float* inv_b, const float* w,
                        const float* bias, float* out, int B, int L, int C, int ksize,
                        int use_tanh, int precise, void* stream);

/* conv_post + clamp / tanh alone, on an already activated 16-bit (B, L, C) tensor (svc_snake_aa's output): the
 * 16-bit modes run activation_post on the tensor cores and this kernel behind it.  w (ksize, C) fp32, bias or NULL. */
int svc_conv_post(const void* act, int act_dtype, const float* w, const float* bias, float* out, int B, int L,
                  int C, int ksize, int use_tanh, void* stream);

/* ---------------------------------------------------------------------------
 * svc_cfg_euler: x += dt * (c0 v[0] + c1 v[1] + c2 v[2]); rows t < prompt_len and rows
 * t >= x_lens[b] are set to 0; optionally also writes x in the operand dtype.
 * Replaces flow_matching.py:98-110 and v2/cfm.py:86-130.  x: (B, T, C) fp32 contiguous;
 * v: (n_branch*B, T, C) fp32 contiguous, branch-major.
 * ------------------------------------------------------------------------- */
int svc_cfg_euler(float* x, const float* v, int n_branch, float c0, float c1, float c2, float dt,
                  int B, int T, int C, int prompt_len, const int* x_lens, void* x_op, int op_dtype,
                  void* stream);

/* (B, C, T) fp32 -> (B, T, C) in out_dtype; rows t in [zero_from, zero_to) are written as 0.
 * Entry of the sampler/vocoder (flow_matching.py:76-81; diffusion_transformer.py:503-504). */
int svc_bct_to_btc(const float* in, void* out, long long o_bstride, long long o_rstride, int B,
                   int C, int T, int zero_from, int zero_to, int out_dtype, void* stream);
/* (B, T, C) fp32 -> (B, C, T) fp32 */
int svc_btc_to_bct(const float* in, float* out, int B, int T, int C, void* stream);

/* fp32 -> operand dtype, n contiguous elements */
int svc_cast(const float* in, void* out, long long n, int out_dtype, void* stream);

/* Folded RMS norm, weight side: out[s][n][k] = W[n][k] * g[k] * mul[s][k] in out_dtype, s < S (one copy per Euler
 * step).  W fp32 (N, K) with row stride w_rstride; g [K] (RMSNorm weight) and mul [S][mul_stride] (AdaLN weight per
 * step) may be NULL.  The matching bias a_s W^T is an ordinary fp32 svc_gemm.  See svc_gemm_desc.row_ss_in. */
int svc_scale_cols(const float* W, long long w_rstride, const float* g, const float* mul, long long mul_stride,
                   void* out, int out_dtype, int S, int N, int K, void* stream);

/* Reflect padding halo of the WaveNet input (encodec.py:212-228 `pad1d(..., 'reflect')`):
 * buf is (B, T + 2*pad, C) in the operand dtype with the body at rows [pad, pad+len_b);
 * writes rows pad-1-i <- body[i+1] and pad+len_b+i <- body[len_b-2-i], i in [0,pad). */
int svc_reflect_halo(void* buf, long long bstride, long long rstride, int B, int T, int C, int pad,
                     const int* lens, int dtype, void* stream);

/* Sinusoidal timestep features (diffusion_transformer.py:341-359): out (n, 2*half) =
 * [cos(1000 t f_i) | sin(1000 t f_i)], f_i = freqs[i] = exp(-ln(1e4) i / half) (the module's
 * `freqs` buffer, :337-340). */
int svc_timestep_embedding(const float* t, const float* freqs, float* out, int n, int half,
                           void* stream);

/* dst[b, 0:D] = src[b * src_bstride + 0:D] for b < B (token rows, diffusion_transformer.py:512-517) */
int svc_set_rows(const float* src, long long src_bstride, float* dst, long long dst_bstride, int B,
                 int D, void* stream);

/* ---- InterpolateRegulator pieces (modules/length_regulator.py:90-141) ----------------------------
 * svc_interp_rows: out[b,t,:] = src[b, idx[t], :] (+ add_vec[:]) (+ emb[emb_q[b*q_bstride + emb_idx[t]], :])
 *   = F.interpolate(x, size, mode='nearest') as a row gather (:115; idx is the caller's nearest-index
 *   table), `x + f0_mask` (:122) or `x + interpolate(f0_embedding(quantized_f0))` (:124-129).
 *   src fp32 (B, Tin, D); out (B, Tout, D) fp32 or bf16 (dtype). */
int svc_interp_rows(const float* src, long long src_bstride, long long src_rstride, const int* idx,
                    const float* add_vec, const float* emb, const int* emb_q, long long q_bstride,
                    const int* emb_idx, void* out, long long out_bstride, long long out_rstride, int B,
                    int Tout, int D, int dtype, void* stream);
/* svc_groupnorm1_mish: out = Mish(GroupNorm(1 group, C)(x)) for x (B, T, C) fp32 frames-major: mean / var
 *   over all T*C values of a sample (fp64 accumulation), per-channel affine, Mish = y*tanh(softplus(y))
 *   (nn.GroupNorm(groups=1) + nn.Mish, :50-53).  stats_ws: 2*B doubles of scratch. */
int svc_groupnorm1_mish(const float* x, long long bstride, long long rstride, const float* gamma,
                        const float* beta, float eps, double* stats_ws, void* out, long long out_bstride,
                        long long out_rstride, int B, int T, int C, int dtype, int precise, void* stream);
/* svc_mask_rows: x[b, t, :] = 0 for t >= lens[b]  (`out * mask`, :140). */
int svc_mask_rows(float* x, long long bstride, long long rstride, const int* lens, int B, int T, int D,
                  void* stream);

/* ---- mel front-end pieces (modules/audio.py:45-82 `mel_spectrogram`) --------------------------------
 * The STFT is svc_gemm (fp32) over the padded audio viewed as rows of `hop` samples, n_fft/hop segments with
 * row shifts 0..n_fft/hop-1, against the Hann-windowed DFT matrix [cos | -sin].
 * svc_reflect_pad1d: out[b, i] = y[b, reflect(i - pad)], i < L + 2*pad; zeros up to out_len (:58-61). */
int svc_reflect_pad1d(const float* y, long long y_bstride, int B, int L, int pad, float* out,
                      long long out_bstride, long long out_len, void* stream);
/* svc_stft_mag: mag[m,k] = sqrt(re^2 + im^2 + eps), spec row m = [re_0..re_{nb-1} | im_0..im_{nb-1} | pad] (:78) */
int svc_stft_mag(const float* spec, long long spec_rstride, long long rows, int n_bins, float eps, float* mag,
                 long long mag_rstride, void* stream);
/* svc_log_clamp: x = log(max(x, clip))  (dynamic_range_compression_torch, :24-25) */
int svc_log_clamp(float* x, long long n, float clip, void* stream);

/* Stitch n vocoded chunks into one waveform (inference.py:343-350 `crossfade`, :505-527 chunk loop;
 * seed_vc_wrapper.py:190-285 `_stream_wave_chunks`).  waves is (n, wave_stride) fp32, chunk k holds
 * lens[k] samples.  Chunks k < n-1 contribute samples [0, lens[k]-overlap), the last one all of its
 * samples; out offset of chunk k is offs[k] (offs[0] = 0, offs[k+1] = offs[k] + lens[k] - overlap).
 * For k > 0 the first `overlap` samples are w_k[i]*fade_in[i] + w_{k-1}[lens[k-1]-overlap+i]*fade_out[i],
 * evaluated in fp64 and rounded to fp32 once (numpy float32*float64 semantics of the reference);
 * fade_in / fade_out are the reference's cos^2 ramps (fp64, `overlap` entries, device memory). */
int svc_crossfade_stitch(const float* waves, long long wave_stride, const int* lens, const long long* offs,
                         int n_chunks, int overlap, const double* fade_in, const double* fade_out,
                         float* out, long long total, void* stream);

/* Streaming SOLA stitch, one launch for B concurrent streams (real-time-gui.py:1103-1137): offset[b] =
 * argmax_o  sum_i x[o+i]*buf[i] / sqrt(sum_i x[o+i]^2 + 1e-8), o in [0, search]; out[b, i] = x[off+i] for
 * i in [sb, block), x[off+i]*fade_in[i] + buf[i]*fade_out[i] for i < sb; then buf[b, :] = the next sb
 * samples after the block (the state carried to the next tick).  infer (B, infer_len) fp32,
 * sola_buf (B, sb) fp32 in/out, fade_in / fade_out (sb) fp32, out (B, block) fp32, offsets (B) int32. */
int svc_sola_stitch(const float* infer, long long infer_bstride, int infer_len, float* sola_buf,
                    long long buf_bstride, const float* fade_in, const float* fade_out, float* out,
                    long long out_bstride, int* offsets, int B, int sola_buffer_frame, int sola_search_frame,
                    int block_frame, void* stream);

/* ---------------------------------------------------------------------------
 * HiFT vocoder pieces (reference modules/hifigan/generator.py).  The generator's convolutions run on
 * svc_gemm; these are the non-GEMM steps.
 * ------------------------------------------------------------------------- */
/* out = f(x) elementwise on a (B, T, C) fp32 view -> out_dtype view.  kind 0: leaky_relu(slope) (generator.py:404,424),
 * 1: ELU (f0_predictor.py:33-49), 2: Snake x + sin^2(alpha_c x) / (alpha_c + 1e-9) (:79-90), 3: |x| (f0_predictor.py:55) */
int svc_unary(const float* x, long long x_bstride, long long x_rstride, void* out, long long o_bstride,
              long long o_rstride, int B, int T, int C, int kind, float slope, const float* alpha, int out_dtype,
              int precise, void* stream);
/* F0 (B, Tm) Hz -> harmonic source (B, Tm * scale): nn.Upsample(nearest) (:295,367), SineGen.forward (:208-243) with
 * the phase draw `phase` (B, H) and the Gaussian draw `noise` (B, H, Tm * scale) or NULL given by the caller,
 * tanh(Linear_{H -> 1}) (:270-275).  prefix_ws: B * H * Tm doubles of workspace. */
int svc_hift_source(const float* f0, long long f0_bstride, const float* phase, const float* noise, const float* lin_w,
                    float lin_b, double* prefix_ws, float* out, long long out_bstride, int B, int Tm, int H, int scale,
                    float sampling_rate, float sine_amp, float noise_std, float voiced_threshold, void* stream);
/* torch.stft(n_fft 16, hop 4, Hann, center, reflect) of s (B, L) (:372-378) -> (B, rows, Cpad) out_dtype with channels
 * [re_0..re_8, im_0..im_8, 0...]; rows >= L / 4 + 1, rows past that are zero-filled. */
int svc_hift_stft(const float* s, long long s_bstride, void* out, long long o_bstride, long long o_rstride, int B, int L,
                  int rows, int Cpad, int out_dtype, void* stream);
/* conv_post output x (B, TT, >= 18) fp32 -> waveform (B, 4 (TT - 1)): exp / sin heads, magnitude clip, polar,
 * torch.istft(16, 4, Hann, center), clamp to +-audio_limit (:426-435, :380-385). */
int svc_hift_istft(const float* x, long long x_bstride, long long x_rstride, float* wav, long long wav_bstride, int B,
                   int TT, float clip_mag, float audio_limit, int precise, void* stream);

/* ---------------------------------------------------------------------------
 * Graph-level entry point: svc_dit_step = ONE estimator call (all CFG branches) of the Euler loop, i.e.
 * DiT.forward from `cond_x_merge_linear` to the velocity (diffusion_transformer.py:508-537 with Transformer.forward
 * :112-143, TransformerBlock :173-191, the WaveNet / final_mlp heads; v2: modules/v2/dit_wrapper.py:137-152,
 * dit_model.py:109-143), built from the entry points above.  A host in any language prepares the two structs once
 * per model / per solve and then calls svc_dit_step + svc_cfg_euler per Euler step; nothing is allocated inside.
 *
 * Layouts (all contiguous unless a stride is given; R = n_branch * B rows, Tq = T + ntok):
 *   weights   (N, K) row-major in the dtype named; conv weights (k, N, K); w13 / WaveNet in-layers row-interleaved
 *             for the pair activations (SVC_ACT_SWIGLU_PAIR / SVC_ACT_TANH_SIG_PAIR)
 *   ada       per-step row of every AdaLN projection, fp32 (n_steps, n_ada): v1 chunks [w | b] (2D), v2 chunks
 *             [shift, 1+scale, gate, shift, 1+scale, gate] (6D) per layer, final [mul | add], FinalLayer [shift | 1+scale]
 *   h (R, Tq, D) fp32; xn, att (R, Tq, D), qkv (R, Tq, 3D), ff (R, Tq, I) op dtype; h_op, xn_f, skips (R, Tq, D),
 *   x_res / y (R, T, D or Dw), xw_op (R, T + 2 pad, Dw), acts (R, T, wn_layers * Dw), ln (R, T, Dw) stream dtype;
 *   xw, wn_out (R, T, Dw), v (R, T, C) fp32.
 * ------------------------------------------------------------------------- */
#define SVC_MAX_LAYERS 32
#define SVC_MAX_WN_LAYERS 16
#define SVC_MAX_BRANCH 3

typedef struct svc_dit_weights {
    int version;                       /* 1 (modules/diffusion_transformer.py) or 2 (modules/v2) */
    int D, H, L, C, I;
    int time_as_token, style_as_token, uvit, long_skip;
    int head;                          /* 0 = final_mlp, 1 = WaveNet + FinalLayer + conv2 */
    int Dw, wn_layers, wn_kernel;
    int op_dtype, stream_dtype;        /* SVC_* of branch operands / of operands that carry the residual stream */
    const void* wqkv[SVC_MAX_LAYERS];  /* op dtype */
    const void* wo[SVC_MAX_LAYERS];
    const void* w13[SVC_MAX_LAYERS];
    const void* w2[SVC_MAX_LAYERS];
    const float* g_attn[SVC_MAX_LAYERS];
    const float* g_ffn[SVC_MAX_LAYERS];
    const void* skip_w[SVC_MAX_LAYERS]; /* (D, 2D) stream dtype, NULL unless the layer receives a U-ViT skip */
    const float* skip_b[SVC_MAX_LAYERS];
    int ada_attn[SVC_MAX_LAYERS];      /* v1: offset of the attention-norm chunk in the ada row, -1 = plain RMSNorm; v2: offset of the 6D chunk */
    int ada_ffn[SVC_MAX_LAYERS];       /* v1: offset of the ffn-norm chunk, -1 = plain RMSNorm */
    const float* g_final;
    int ada_final;
    const void* merge_wx;              /* x columns of cond_x_merge_linear: (D, C) view, stream dtype */
    long long merge_w_rstride;
    const void* lskip_w;               /* skip_linear (D, D + C), stream dtype, or NULL */
    const float* lskip_b;
    const void* mlp0_w;                /* final_mlp (head 0), stream dtype */
    const float* mlp0_b;
    const void* mlp2_w;
    const float* mlp2_b;
    const void* conv1_w;               /* head 1, stream dtype */
    const float* conv1_b;
    const void* resp_w;
    const void* conv2_w;
    const float* conv2_b;
    const void* fl_w;
    const float* fl_b;
    const void* wn_in_w[SVC_MAX_WN_LAYERS];   /* (k, 2 Dw, Dw) */
    const void* wn_rs_w[SVC_MAX_WN_LAYERS];   /* res half of res_skip: (Dw, Dw) */
    const float* wn_rs_b[SVC_MAX_WN_LAYERS];
    const void* wn_skip_w;             /* (Dw, wn_layers * Dw): all skip halves side by side */
    const float* wn_skip_b;            /* summed skip biases + res_projection bias */
    int ada_fl;
    const float* rope_tab;             /* (positions, 32, 2) */
    const float* rope_tab_t;           /* pair-major copy (32, rope_ld, 2) */
    int rope_ld;
} svc_dit_weights;

typedef struct svc_dit_state {
    int B, T, n_branch, n_steps, n_ada;
    const float* ada;                  /* (n_steps, n_ada) */
    const float* t1;                   /* (n_steps, D): time token rows */
    const float* wn_g;                 /* (n_steps, wn_layers * 2 Dw): WaveNet cond_layer output + in-layer bias */
    int const_kind[SVC_MAX_BRANCH];    /* hoisted columns of cond_x_merge_linear: 0 = (B, T, D) fp32 matrix, 1 = D vector */
    const float* const_ptr[SVC_MAX_BRANCH];
    const float* style_tok;            /* (B, D) */
    const float* style_tok_null;       /* (D) */
    int branch_style[SVC_MAX_BRANCH];  /* 1: the branch sees the style token, 0: the null token */
    const int* kv_len;                 /* (R) keys per row incl. tokens */
    const int* wn_lens;                /* (R) frames per row */
    float* h;
    void* xn;
    void* xn_f;
    void* qkv;
    void* att;
    void* ff;
    void* h_op;
    void* skips[SVC_MAX_LAYERS / 2];
    float* v;
    void* x_res;
    void* y;
    float* xw;
    void* xw_op;
    void* acts;
    float* wn_out;
    void* ln;
} svc_dit_state;

/* velocity of every branch at Euler step s -> st->v (R, T, C) fp32.  x_op: (B, T, C) stream dtype (the copy
 * svc_cfg_euler writes). */
int svc_dit_step(const svc_dit_weights* w, const svc_dit_state* st, int s, const void* x_op, void* stream);

/* ---------------------------------------------------------------------------
 * Graph-level entry point: svc_bigvgan_forward = BigVGAN.forward (modules/bigvgan/bigvgan.py:360-386) with AMPBlock1
 * (:132-141): conv_pre, per stage the polyphase ConvTranspose1d and 3 x 3 x [snake, conv, snake, conv, + x] with the
 * (r0 + r1 + r2) / 3 average fused in the GEMM epilogues, activation_post + conv_post + clamp / tanh.  Weights are the
 * prepared copies (folded weight-norm, (taps, N, K) conv layout, time-to-depth regrouping for the narrow stages,
 * Snake exp(alpha) and 1 / (exp(beta) + 1e-9)).  The caller provides the workspace; nothing is allocated inside.
 * ------------------------------------------------------------------------- */
#define SVC_MAX_STAGES 8
#define SVC_MAX_TAPS 16

typedef struct svc_conv_plan {
    int f;                             /* time-to-depth factor: the (B, L, O) buffer is read as (B, L / f, f * O) */
    int n_taps;
    int shifts[SVC_MAX_TAPS];          /* row shift of every tap (in regrouped rows) */
    const void* w;                     /* (n_taps, f * O, f * I) operand dtype */
    const float* b;                    /* (f * O) */
    int k;                             /* kernel size of the original conv (FLOP accounting only) */
} svc_conv_plan;

typedef struct svc_amp_pair {
    const float *a1, *inv_b1, *a2, *inv_b2;   /* Snake parameters of the two activations, (O) each */
    svc_conv_plan c1, c2;
} svc_amp_pair;

typedef struct svc_bigvgan_stage {
    int u, O;                          /* upsampling factor, output channels */
    int n_delta;
    int deltas[SVC_MAX_TAPS];
    const void* up_w;                  /* (n_delta, u * O, I) polyphase weight */
    const float* up_b;                 /* (u * O) */
    svc_amp_pair pairs[3][3];          /* [resblock][dilation] */
} svc_bigvgan_stage;

typedef struct svc_bigvgan_weights {
    int n_mels, c0, n_stages, n_kernels, n_dil;
    int op_dtype, precise;
    const void* pre_w;                 /* (7, c0, n_mels) */
    const float* pre_b;
    svc_bigvgan_stage stages[SVC_MAX_STAGES];
    const float *post_a, *post_inv_b;
    const float* post_w;               /* (post_k, C_last) fp32 */
    const float* post_b;               /* or NULL */
    int post_k, use_tanh;
} svc_bigvgan_weights;

long long svc_bigvgan_workspace_bytes(const svc_bigvgan_weights* w, int B, int Tm);
/* mel (B, n_mels, Tm) fp32 -> out (B, Tm * prod(u)) fp32 in [-1, 1] */
int svc_bigvgan_forward(const svc_bigvgan_weights* w, const float* mel, void* workspace, float* out, int B, int Tm,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEEDVC_B200_H_ */
