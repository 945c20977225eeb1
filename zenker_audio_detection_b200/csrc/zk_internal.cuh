// Internal (non-ABI) entry points shared between the translation units of libzk_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace zk {

// out_pitch: elements between output rows (0 = N); prof_cls: launch-accounting class override (-1 = by epilogue)
int gemm_bf16(const void* a, const void* w, const float* bias, void* out, long long M, int N, int K, int epilogue,
              const float* aux, int aux_rows, cudaStream_t stream, long long out_pitch = 0, int prof_cls = -1);
int attention_bf16(const void* qkv, void* out, int batch, int tokens, cudaStream_t stream);
int layernorm_bf16(const float* x, const float* w, const float* b, float eps, void* out, long long rows, int cols,
                   cudaStream_t stream);
int layernorm_bf16_cls(const float* x, const float* w, const float* b, float eps, void* out, long long rows, int cols,
                       int prof_cls, cudaStream_t stream);
int f32_to_bf16(const float* in, void* out, long long n, cudaStream_t stream);
// last-layer tail (only tokens 0 and 1 of every window reach the classifier, HF:modeling...:378-380):
// rows {0,1} of every window: h (bf16) -> hq [2*batch][768], x (f32) -> x2 [2*batch][768]
int gather_head_rows(const void* h, const float* x, int batch, int tokens, void* hq, float* x2, cudaStream_t stream);
// attention of the two head queries of every window against all keys (K, V read in place from qkv)
int attention_head_rows(const void* q2, const void* qkv, void* out2, int batch, int tokens, cudaStream_t stream);

// patch gather (im2col) for the 16x16 / stride 10 patch embedding; see zk_ops.cu
struct GatherSrc {
  const float* features;      // mode 0: [batch][max_length][128] normalised
  const float* fbank;         // mode 1: continuous fbank [frames][128], un-normalised
  const int32_t* window_index;  // mode 1, optional
  long long fbank_frames;
  int window_base, frames_per_hop, valid_frames;
  float mean, std2;           // mode 1: (x-mean)/std2
};
int gather_patches(const GatherSrc& src, int batch, int max_length, void* a_out, cudaStream_t stream);
int write_special_tokens(const float* cls, const float* dist, const float* pos, float* x, int batch, int tokens,
                         cudaStream_t stream);
int head_logits(const float* x, int batch, int tokens, const float* fln_w, const float* fln_b, const float* hln_w,
                const float* hln_b, const float* head_w, const float* head_b, int num_labels, float eps, float* logits,
                cudaStream_t stream);

}  // namespace zk
