/* zk_b200.h -- C ABI of libzk_b200.so: the sm_100a kernels of the two-stage sliding-window
 * inference path (zenker-audio-detection), one entry point per stage of the path.
 *
 * The reference has NO native / FFI interface of its own: its hot path is
 * Python glue around torchaudio and transformers.  Each entry point below therefore cites the
 * reference (or third-party) call it replaces:
 *
 *   ref:   /root/reference/src/test_long_audio_windows_2stage.py
 *   refc:  /root/reference/src/test_long_audio_windows_2stage_cache.py
 *   TA:    torchaudio 2.11.0  (functional/functional.py, compliance/kaldi.py)
 *   HF:    transformers 5.5.0 (models/audio_spectrogram_transformer/)
 *
 * Conventions
 *   - every pointer named d_* is DEVICE memory owned by the caller; h_* is host memory.
 *   - the library never allocates or frees caller-visible device memory.  The only device
 *     allocations it owns live inside the opaque handles (zk_fbank_plan, zk_model).
 *   - all work is enqueued on `stream` (a cudaStream_t / CUstream); no hidden synchronisation,
 *     except zk_model_create / zk_fbank_plan_create which synchronise once at creation.
 *   - return value: 0 success; <0 a zk_status argument/shape error; >0 a cudaError_t.
 *     zk_last_error_string() returns a thread-local description of the last failure.
 *   - nothing here falls back to a CPU path: on a device that is not sm_100 every compute entry
 *     point returns ZK_ERR_UNSUPPORTED.
 */
#ifndef ZK_B200_H_
#define ZK_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZK_ABI_VERSION 2

typedef void* zk_stream_t; /* cudaStream_t */

enum zk_status {
  ZK_OK = 0,
  ZK_ERR_ARG = -1,         /* null pointer, negative size, misaligned pointer */
  ZK_ERR_SHAPE = -2,       /* a dimension the kernels do not support */
  ZK_ERR_UNSUPPORTED = -3, /* not an sm_100 device / driver entry point missing */
  ZK_ERR_WORKSPACE = -4,   /* workspace too small */
  ZK_ERR_INTERNAL = -5
};

int zk_abi_version(void);
const char* zk_last_error_string(void);
/* 0 when the current device is sm_100 (B200) and the driver exposes cuTensorMapEncodeTiled. */
int zk_device_check(void);

/* ------------------------------------------------------------------------------------------
 * (1) resample -- replaces torchaudio.functional.resample as called by ref:55-58
 *     (TA:functional/functional.py:1405-1432: pad(width, width+orig), conv1d(stride=orig) with
 *     the `new` polyphase filters, truncate to ceil(new*n/orig)), fused with the channel mean
 *     of ref:55-56.
 *     d_in      [channels][ch_pitch] f32 planar (what torchaudio.load returns), n_in samples used
 *     d_taps    [new_][2*width+orig] f32 (TA:...:1341-1402; built by the host in f32)
 *     d_out     [n_out] f32, n_out = ceil(new_*n_in/orig)
 * ------------------------------------------------------------------------------------------ */
int zk_resample_f32(const float* d_in, int64_t n_in, int channels, int64_t ch_pitch, const float* d_taps, int orig,
                    int new_, int width, float* d_out, int64_t n_out, zk_stream_t stream);
/* Same, from interleaved PCM16 frames [n_in][channels] (value/32768 as torchaudio.load
 * normalises), SURVEY.md section 8f #2. */
int zk_resample_pcm16(const int16_t* d_in, int64_t n_in, int channels, const float* d_taps, int orig, int new_,
                      int width, float* d_out, int64_t n_out, zk_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (2)+(3) Kaldi fbank -- replaces torchaudio.compliance.kaldi.fbank as called by
 *     HF:feature_extraction_audio_spectrogram_transformer.py:116-121 (TA:compliance/kaldi.py:514-645):
 *     25 ms / 10 ms frames (400 / 160 samples at 16 kHz), per-frame DC removal, pre-emphasis with
 *     replicate padding, window, zero-pad to 512, |rfft|^2, mel filterbank, log(max(., eps)).
 *     The plan owns the constant tables (window, twiddles, sparse mel bank).
 * ------------------------------------------------------------------------------------------ */
typedef struct zk_fbank_plan zk_fbank_plan;
/* h_window [400] f32, h_mel [num_mel][256] f32 dense bank (TA:compliance/kaldi.py:436-511), both
 * computed by the host with the same torch f32 ops torchaudio uses. num_mel must be 128. */
int zk_fbank_plan_create(const float* h_window, const float* h_mel, int num_mel, float preemph, float log_floor,
                         zk_fbank_plan** out);
void zk_fbank_plan_destroy(zk_fbank_plan* plan);
/* frames of a waveform of n samples: 0 if n < 400 else 1 + (n-400)/160 (TA:compliance/kaldi.py:63-67) */
int64_t zk_fbank_num_frames(int64_t n);
/* Continuous fbank over one recording: d_wave [n] f32 -> d_out [m][128] f32, m = zk_fbank_num_frames(n).
 * window k of the reference equals rows [50k, 50k+98) of this matrix bit-exactly. */
int zk_fbank_f32(const zk_fbank_plan* plan, const float* d_wave, int64_t n, float* d_out, int64_t m,
                 zk_stream_t stream);
/* ASTFeatureExtractor.__call__ contract (HF:...:104-156,215-227): per-window fbank, zero-pad or
 * truncate to max_length rows, then (x-mean)/(2*std) when do_normalize (pad rows become -mean/(2*std)).
 *   d_windows [batch][win_pitch] f32, win_len samples used per window
 *   d_out     [batch][max_length][128] f32 */
int zk_fx_contract_f32(const zk_fbank_plan* plan, const float* d_windows, int batch, int64_t win_len,
                       int64_t win_pitch, int do_normalize, float mean, float std, int max_length, float* d_out,
                       zk_stream_t stream);

/* Dataset normalisation statistics as an EPILOGUE of the feature kernel (utils/compute_ast_normalization_stats.py:
 * 62-80: extractor with do_normalize = False, float64 sum and sum of squares over every element of the zero-padded
 * (B, max_length, 128) features): the un-normalised log-mel values of `batch` waveforms are added to d_acc[0] (sum) and
 * d_acc[1] (sum of squares) in fp64 while they are computed; d_out [batch][max_length][128] receives the features, or
 * is NULL for statistics only (nothing but 16 bytes leaves the SMs).  The caller zeroes d_acc and counts
 * batch * max_length * 128 elements per call. */
int zk_fx_stats_f32(const zk_fbank_plan* plan, const float* d_windows, int batch, int64_t win_len, int64_t win_pitch,
                    int max_length, float* d_out, double* d_acc, zk_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (4) AST forward -- replaces ASTForAudioClassification.forward (HF:modeling_audio_spectrogram_
 *     transformer.py:403-451) for the AST-base geometry (hidden 768, 12 layers, 12 heads, MLP 3072,
 *     patch 16, strides 10/10, 128 mel x max_length frames, LayerNorm eps from config).
 *     Weights are given as fp32 DEVICE pointers in HF layout (nn.Linear weight = [out][in]); the
 *     model handle keeps 16-bit copies (QKV fused) plus fp32 biases / LayerNorm / position table.
 *
 *     Two precisions share one handle:
 *       ZK_PRECISION_FAST     every GEMM / attention operand is ONE 16-bit value (operand_format: fp16 by default,
 *                             11 significant bits; bf16, 8 bits, selectable), fp32 accumulation; the throughput path.
 *       ZK_PRECISION_RECHECK  every operand is TWO fp16 planes (x = hi + lo, 22 bits) and every contraction adds the
 *                             products lo*hi + hi*lo + hi*hi: fp32-class logits (a few 1e-6 from an fp32 CPU forward)
 *                             at ~3.5x the time.  The cascade routes the windows whose fast logits fall within eps of
 *                             a decision threshold through it BEFORE gating (ref:312-320 gates on exact fp32
 *                             probabilities), so thresholded decisions equal the reference's.
 * ------------------------------------------------------------------------------------------ */
#define ZK_AST_LAYERS 12
enum zk_operand_format { ZK_FMT_BF16 = 0, ZK_FMT_F16 = 1 };
enum zk_precision { ZK_PRECISION_FAST = 0, ZK_PRECISION_RECHECK = 1 };
typedef struct zk_ast_layer_weights {
  const float *ln1_w, *ln1_b;     /* layernorm_before */
  const float *q_w, *q_b;         /* attention.attention.query  [768][768], [768] */
  const float *k_w, *k_b;
  const float *v_w, *v_b;
  const float *o_w, *o_b;         /* attention.output.dense */
  const float *ln2_w, *ln2_b;     /* layernorm_after */
  const float *fc1_w, *fc1_b;     /* intermediate.dense [3072][768] */
  const float *fc2_w, *fc2_b;     /* output.dense [768][3072] */
} zk_ast_layer_weights;

typedef struct zk_ast_weights {
  int32_t num_layers;   /* <= ZK_AST_LAYERS (12 for AST-base; fewer only for tests) */
  int32_t max_length;   /* frames per window the position table was built for (1024) */
  int32_t num_labels;   /* 2 */
  float ln_eps;         /* 1e-12 */
  int32_t operand_format; /* zk_operand_format of the FAST path (the re-check planes are always fp16) */
  const float* cls_token;     /* [768] */
  const float* dist_token;    /* [768] */
  const float* pos_emb;       /* [tokens][768], tokens = 2 + 12*((max_length-16)/10+1) */
  const float* patch_w;       /* [768][1][16 freq][16 time] */
  const float* patch_b;       /* [768] */
  zk_ast_layer_weights layer[ZK_AST_LAYERS];
  const float *final_ln_w, *final_ln_b;   /* audio_spectrogram_transformer.layernorm */
  const float *head_ln_w, *head_ln_b;     /* classifier.layernorm */
  const float *head_w, *head_b;           /* classifier.dense [num_labels][768], [num_labels] */
} zk_ast_weights;

typedef struct zk_model zk_model;
int zk_model_create(const zk_ast_weights* w, zk_model** out);
void zk_model_destroy(zk_model* m);
int zk_model_num_tokens(const zk_model* m);
int zk_model_max_length(const zk_model* m);
/* bytes of caller-provided workspace zk_model_forward* needs for `batch` windows at `precision` */
size_t zk_model_workspace_bytes(const zk_model* m, int batch, int precision);

/* Contract path: d_features [rows][max_length][128] f32 (already normalised, what ASTFeatureExtractor returns)
 * -> d_logits [batch][num_labels] f32.  Window i of the batch reads feature row (d_row_index ? d_row_index[i] : i)
 * (the re-check pass re-runs a few rows of a batch it has already seen).  If d_hidden != NULL the final residual
 * stream [batch][tokens][768] f32 (before the last LayerNorm) is copied there (test hook). */
int zk_model_forward(zk_model* m, const float* d_features, const int32_t* d_row_index, int batch, int precision,
                     void* d_workspace, size_t workspace_bytes, float* d_logits, float* d_hidden, zk_stream_t stream);
/* Fused path (ref:322-328 / refc:499-500 without materialising (batch,1024,128)): window i of the
 * batch reads rows [first_frame_i, first_frame_i + valid_frames) of the continuous fbank d_fbank
 * [m][128] (un-normalised), first_frame_i = (d_window_index ? d_window_index[i] : window_base + i)
 * * frames_per_hop; rows >= valid_frames are the pad constant; normalisation (x-mean)/(2*std) is
 * applied while gathering. */
int zk_model_forward_fbank(zk_model* m, const float* d_fbank, int64_t fbank_frames, const int32_t* d_window_index,
                           int window_base, int frames_per_hop, int valid_frames, float mean, float std, int batch,
                           int precision, void* d_workspace, size_t workspace_bytes, float* d_logits, zk_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (5) Stage-1 gate + order-preserving compaction -- replaces ref:111 (softmax) and ref:312-320
 *     (argmax & threshold, np.where), with refc:471-478's optional forward_min_prob.
 *     d_logits [n][2] f32 -> d_probs [n][2] f32 (softmax), d_pred [n] i32 ((argmax==1)&(p1>=thr)),
 *     d_index [<=n] i32 ascending indices of forwarded windows, d_count [1] i32.
 *     min_prob < 0 disables the extra gate.  Integer outputs are bit-exact functions of d_probs.
 * ------------------------------------------------------------------------------------------ */
int zk_gate_compact(const float* d_logits, int n, float threshold, float min_prob, float* d_probs, int32_t* d_pred,
                    int32_t* d_index, int32_t* d_count, zk_stream_t stream);
/* softmax only (Stage 2: ref:111) */
int zk_softmax2(const float* d_logits, int n, float* d_probs, zk_stream_t stream);
/* Decision re-check selection: rows i of d_logits [n][2] whose margin l1 - l0 lies within eps of ANY of the
 * num_margins decision points h_margins[] (HOST array, <= 4 entries; logit(thr) of every threshold applied to the stage:
 * ref:313-317 argmax (0) and p1 >= thr, refc:471-478 forward_min_prob, ref:333 / refc:512-520 p_zenker >= thr2).
 * Order-preserving compaction: d_pos[j] = i (row to overwrite later), d_window[j] = d_src_window ? d_src_window[i] : i
 * (window index to re-run), d_count[0] = how many.  Integer outputs are exact functions of d_logits. */
int zk_band_select(const float* d_logits, int n, const float* h_margins, int num_margins, float eps,
                   const int32_t* d_src_window, int32_t* d_pos, int32_t* d_window, int32_t* d_count, zk_stream_t stream);
/* d_dst[d_pos[j]][0..1] = d_src[j][0..1] for j < count (logits of the re-checked windows back into place) */
int zk_scatter_rows2(const float* d_src, const int32_t* d_pos, int count, float* d_dst, zk_stream_t stream);
/* Dataset normalisation statistics (utils/compute_ast_normalization_stats.py:77-80): d_acc[0] += sum(x), d_acc[1] +=
 * sum(x^2) over n floats, accumulated in float64 (the caller zeroes d_acc before the first batch; zero-padded feature
 * rows simply add 0, so the continuous or the padded layout give the same sums). */
int zk_sum_sumsq_f64(const float* d_x, int64_t n, double* d_acc, zk_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (6) The whole cascade of one recording in ONE call -- replaces the per-recording body of ref:301-348 / refc:433-531
 *     (SURVEY.md section 8b "cascade_run"): continuous fbank, Stage 1 in batches, decision re-check, gate + compaction,
 *     Stage 2 on the compacted windows, re-check, softmax.  A C host drives the path with zk_resample_* + this.
 *       d_audio16k  [n_samples] f32 mono 16 kHz, n_samples >= window_samples (zero-pad a shorter recording, ref:70-73)
 *       outputs     d_probs1 [N][2] f32, d_pred1 [N] i32, d_index [N] i32 (first num_forwarded valid, ascending),
 *                   d_probs2 [N][2] f32 (row j belongs to window d_index[j]);  N = (n_samples - window) / hop + 1
 *       h_counts    HOST: N, forwarded count, how many windows each stage re-ran at ZK_PRECISION_RECHECK
 *     hop_samples must be a multiple of 160 (the fused gather; otherwise use the per-window entry points).
 *     SYNCHRONISATION: the batch counts of later steps are produced on the device, so this entry point -- unlike every
 *     other one -- synchronises `stream` up to four times (band count x 2, gate count) before it returns; all kernels
 *     are still enqueued on `stream`, and the outputs are complete only after the caller synchronises it once more.
 * ------------------------------------------------------------------------------------------ */
typedef struct zk_cascade_params {
  int32_t batch_size;      /* windows per launch of the FAST forward (128) */
  int32_t recheck_batch;   /* windows per launch of the RECHECK forward (62: 4.0 waves of 256-row CTA-pair tiles) */
  int32_t window_samples;  /* int(window_sec * 16000), ref:63 */
  int32_t hop_samples;     /* int(hop_sec * 16000), ref:64 */
  float mean1, std1;       /* Stage-1 extractor statistics (HF:feature_extraction...:155-156) */
  float mean2, std2;       /* Stage-2 */
  float thr1;              /* ref:316 */
  float min_prob;          /* refc:471-478; < 0 = none */
  float thr2;              /* ref:333 */
  int32_t stage2_argmax;   /* refc:512-515 */
  float recheck_eps;       /* half-width (logit units) of the re-check band; 0 disables the re-check */
} zk_cascade_params;
typedef struct zk_cascade_counts {
  int32_t num_windows, num_forwarded, rechecked_s1, rechecked_s2;
} zk_cascade_counts;
/* 0 on bad arguments */
size_t zk_cascade_workspace_bytes(const zk_model* m1, const zk_model* m2, int64_t n_samples, const zk_cascade_params* p);
int zk_cascade_run(const zk_fbank_plan* plan, zk_model* m1, zk_model* m2, const float* d_audio16k, int64_t n_samples,
                   const zk_cascade_params* p, void* d_workspace, size_t workspace_bytes, float* d_probs1, int32_t* d_pred1,
                   int32_t* d_index, float* d_probs2, zk_cascade_counts* h_counts, zk_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Building blocks, exported so the parity tests can pin each kernel separately.
 * ------------------------------------------------------------------------------------------ */
enum zk_gemm_epilogue {
  ZK_EPI_BIAS_BF16 = 0,      /* out 16-bit [M][N] = acc*s + bias (name kept from ABI 1; the format is operand_format) */
  ZK_EPI_BIAS_GELU_BF16 = 1, /* out 16-bit = gelu_erf(acc*s + bias)  (HF activations "gelu") */
  ZK_EPI_BIAS_RESID_F32 = 2, /* out f32 [M][N] += acc*s + bias  (residual stream, in place) */
  ZK_EPI_PATCH_F32 = 3,      /* out f32 row (r/P)*(P+2)+2+r%P = acc*s + bias + pos[2 + r%P] (P = aux_rows) */
  ZK_EPI_BIAS_SPLIT = 4,     /* out fp16 [M][2N]: hi plane at columns [0,N), lo plane at [N,2N), of acc*s + bias */
  ZK_EPI_BIAS_GELU_SPLIT = 5 /* same of gelu_erf(acc*s + bias), erf from libdevice (no polynomial) */
};
/* C = A * W^T, fp32 accumulation in TMEM.  N % 256 == 0, K % 64 == 0.
 *   products == 1: d_a 16-bit [M][K] (row pitch lda elements, 0 = K), d_w 16-bit [N][K] (nn.Linear layout, pitch ldw)
 *   products == 3: split operands, fp16 only: d_a [M][2K] = hi | lo planes, d_w [N][2K] = hi | lo planes;
 *                  C = A_lo W_hi^T + A_hi W_lo^T + A_hi W_hi^T
 *   acc_scale: the accumulator is multiplied by it before the bias (weights pre-scaled by a power of two).
 *   ldo: output row pitch in elements (0 = N, or 2N for the ..._SPLIT epilogues). d_aux: position table for PATCH. */
int zk_gemm16(const void* d_a, int64_t lda, const void* d_w, int64_t ldw, const float* d_bias, void* d_out, int64_t ldo,
              int64_t M, int N, int K, int epilogue, int operand_format, int products, float acc_scale,
              const float* d_aux, int aux_rows, zk_stream_t stream);
/* ABI-1 form: bf16 operands, one product, no scale */
int zk_gemm_bf16(const void* d_a, const void* d_w, const float* d_bias, void* d_out, int64_t M, int N, int K,
                 int epilogue, const float* d_aux, int aux_rows, zk_stream_t stream);
/* rows of 768 f32 -> 16-bit, (x-mean)/sqrt(var+eps)*w+b with biased variance (HF:modeling...:260-261);
 * planes == 2 (fp16 only): d_out [rows][2*768] = hi | lo planes */
int zk_layernorm16(const float* d_x, const float* d_w, const float* d_b, float eps, void* d_out, int64_t rows, int cols,
                   int operand_format, int planes, zk_stream_t stream);
int zk_layernorm_bf16(const float* d_x, const float* d_w, const float* d_b, float eps, void* d_out, int64_t rows,
                      int cols, zk_stream_t stream);
/* d_qkv 16-bit [batch*tokens][2304] (q | k | v, head h at columns h*64) -> d_out 16-bit [batch*tokens][768];
 * softmax(q k^T / 8) v per (window, head), no mask (HF:modeling...:150-181). */
int zk_attention16(const void* d_qkv, void* d_out, int batch, int tokens, int operand_format, zk_stream_t stream);
int zk_attention_bf16(const void* d_qkv, void* d_out, int batch, int tokens, zk_stream_t stream);
/* Re-check precision: d_qkv fp16 [batch*tokens][2*2304] = hi | lo planes of (q | k | v) -> d_out fp16
 * [batch*tokens][2*768] = hi | lo planes; scores and P V as three-product sums, softmax in fp32 with exp2f. */
int zk_attention_split(const void* d_qkv, void* d_out, int batch, int tokens, zk_stream_t stream);
int zk_f32_to_bf16(const float* d_in, void* d_out, int64_t n, zk_stream_t stream);
/* d_in f32 [rows][cols] * scale -> d_out [rows][planes*cols] 16-bit (planes == 2, fp16 only: hi | lo) */
int zk_f32_to_16(const float* d_in, void* d_out, int64_t rows, int cols, int operand_format, int planes, float scale,
                 zk_stream_t stream);
/* zk_attention_bf16 + a device-clock timeline of the first 512 CTAs (128 int64 slots each; tuning aid, see
 * scripts/attn_trace.py for the slot layout). */
int zk_attention_trace(const void* d_qkv, void* d_out, int batch, int tokens, int64_t* d_trace, zk_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Launch accounting (bench.py): every kernel launch of the library is counted per class; with
 * timing enabled each launch is additionally bracketed by CUDA events on ITS stream.
 * zk_prof_collect synchronises the device, adds up the event durations and resets the counters.
 * ------------------------------------------------------------------------------------------ */
enum zk_kernel_class {
  ZK_K_RESAMPLE = 0, ZK_K_FBANK = 1, ZK_K_GATHER = 2, ZK_K_GEMM_PATCH = 3, ZK_K_LAYERNORM = 4,
  ZK_K_GEMM_QKV = 5, ZK_K_ATTENTION = 6, ZK_K_GEMM_OUT = 7, ZK_K_GEMM_FC1 = 8, ZK_K_GEMM_FC2 = 9,
  ZK_K_HEAD = 10, ZK_K_GATE = 11, ZK_K_MISC = 12, ZK_K_TAIL = 13,
  ZK_K_RECHECK = 14, /* every launch of a ZK_PRECISION_RECHECK forward */
  ZK_K_NUM_CLASSES = 15
};
void zk_prof_enable(int time_launches);
/* ms[ZK_K_NUM_CLASSES] (0 where timing was off), launches[ZK_K_NUM_CLASSES]; returns 0 or a cudaError_t */
int zk_prof_collect(float* ms, int64_t* launches);
const char* zk_kernel_class_name(int cls);

#ifdef __cplusplus
}
#endif
#endif /* ZK_B200_H_ */
