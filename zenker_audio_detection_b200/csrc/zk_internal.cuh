// Internal (non-ABI) entry points shared between the translation units of libzk_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace zk {

// C = A W^T on the tcgen05 GEMM (zk_gemm.cu).  `fmt` = FMT_BF16 / FMT_F16 (zk_common.cuh).
//   products 1: a [M][K] (pitch lda), w [N][K] (pitch ldw)
//   products 3: split fp16 operands, a [M][2K] = hi | lo, w [N][2K] = hi | lo
// Pitches of 0 mean "dense".  acc_scale multiplies the accumulator before the bias.  prof_cls: launch-accounting class
// override (-1 = by epilogue).
struct GemmArgs {
  const void* a = nullptr;
  long long lda = 0;
  const void* w = nullptr;
  long long ldw = 0;
  const float* bias = nullptr;
  void* out = nullptr;
  long long ldo = 0;
  long long M = 0;
  int N = 0, K = 0;
  int epilogue = 0;
  int fmt = 0;
  int products = 1;
  float acc_scale = 1.0f;
  const float* aux = nullptr;
  int aux_rows = 0;
  int prof_cls = -1;
  // optional LayerNorm tail (ZK_EPI_BIAS_RESID_F32 with N = 768 on the CTA-pair kernel only, see gemm_ln_tail_ok):
  // ln_out [M][768] 16-bit in `fmt` = LayerNorm(out) with ln_w / ln_b / ln_eps, produced by the same kernel;
  // ln_count = device scratch of gemm_ln_tail_counters(M) ints
  const float* ln_w = nullptr;
  const float* ln_b = nullptr;
  void* ln_out = nullptr;
  int* ln_count = nullptr;
  float ln_eps = 0.f;
};
int gemm16(const GemmArgs& g, cudaStream_t stream);
bool gemm_ln_tail_ok(long long M);                 // true when gemm16 can run the residual GEMM with the LayerNorm tail
inline size_t gemm_ln_tail_counters(long long M) { return (size_t)((M + 255) / 256) * 2; }

int attention16(const void* qkv, void* out, int batch, int tokens, int fmt, cudaStream_t stream);
// re-check precision: qkv fp16 [rows][2*2304] (hi | lo planes of q|k|v) -> out fp16 [rows][2*768] (hi | lo)
int attention_split(const void* qkv, void* out, int batch, int tokens, cudaStream_t stream);
// rows of 768 f32 -> 16-bit GEMM operand; planes == 2 (fp16): out [rows][1536] = hi | lo
int layernorm16(const float* x, const float* w, const float* b, float eps, void* out, long long rows, int cols, int fmt,
                int planes, int prof_cls, cudaStream_t stream);
// in f32 [rows][cols] * scale -> out [rows][planes*cols] 16-bit
int f32_to_16(const float* in, void* out, long long rows, int cols, int fmt, int planes, float scale, cudaStream_t stream);
// max |x| over n floats -> d_out[0] (d_out zeroed by the callee)
int max_abs(const float* in, long long n, float* d_out, cudaStream_t stream);
// last-layer tail (only tokens 0 and 1 of every window reach the classifier, HF:modeling...:378-380):
// rows {0,1} of every window: h (16-bit) -> hq [2*batch][768], x (f32) -> x2 [2*batch][768]
int gather_head_rows(const void* h, const float* x, int batch, int tokens, void* hq, float* x2, cudaStream_t stream);
// attention of the two head queries of every window against all keys (K, V read in place from qkv)
int attention_head_rows(const void* q2, const void* qkv, void* out2, int batch, int tokens, int fmt, cudaStream_t stream);

// patch gather (im2col) for the 16x16 / stride 10 patch embedding; see zk_ops.cu
struct GatherSrc {
  const float* features;      // mode 0: [batch][max_length][128] normalised
  const float* fbank;         // mode 1: continuous fbank [frames][128], un-normalised
  const int32_t* window_index;  // optional: mode 1 window numbers / mode 0 rows of `features` to read
  long long fbank_frames;
  int window_base, frames_per_hop, valid_frames;
  float mean, std2;           // mode 1: (x-mean)/std2
};
// a_out [batch*patches][planes*256] 16-bit (planes == 2: fp16 hi | lo)
int gather_patches(const GatherSrc& src, int batch, int max_length, void* a_out, int fmt, int planes, cudaStream_t stream);
int write_special_tokens(const float* cls, const float* dist, const float* pos, float* x, int batch, int tokens,
                         cudaStream_t stream);
int head_logits(const float* x, int batch, int tokens, const float* fln_w, const float* fln_b, const float* hln_w,
                const float* hln_b, const float* head_w, const float* head_b, int num_labels, float eps, float* logits,
                cudaStream_t stream);

// Every launch made by this thread while one of these is alive is accounted to `cls` (ZK_K_RECHECK for a forward at
// ZK_PRECISION_RECHECK), whatever class the kernel normally belongs to.
struct ProfClassOverride {
  explicit ProfClassOverride(int cls);
  ~ProfClassOverride();
  int prev_;
};

}  // namespace zk
