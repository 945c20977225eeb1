// Internal (non-ABI) entry points shared between the translation units of libzk_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace zk {

int gemm_bf16(const void* a, const void* w, const float* bias, void* out, long long M, int N, int K, int epilogue,
              const float* aux, int aux_rows, cudaStream_t stream);
int attention_bf16(const void* qkv, void* out, int batch, int tokens, cudaStream_t stream);
int layernorm_bf16(const float* x, const float* w, const float* b, float eps, void* out, long long rows, int cols,
                   cudaStream_t stream);
int f32_to_bf16(const float* in, void* out, long long n, cudaStream_t stream);

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
