// AST-base forward (HF:modeling_audio_spectrogram_transformer.py:403-451) as a fixed sequence of sm_100a kernels.
//
// zk_model owns the packed weights: bf16 [out][in] matrices for the tcgen05 GEMMs (query/key/value fused into one
// [2304][768] matrix), fp32 biases, LayerNorm parameters, cls/dist tokens and the position table.
// Activations live in the caller's workspace:
//   x    fp32 [B*T][768]   residual stream (kept in fp32 end to end)
//   h    bf16 [B*T][768]   LayerNorm output / attention output (GEMM A operands)
//   qkv  bf16 [B*T][2304]  fused projection output, read in place by the attention kernel
//   mlp  bf16 [B*T][3072]  GELU(fc1) output; its head doubles as the patch-gather matrix [B*P][256]
// Per layer: LN -> QKV GEMM(+bias) -> attention -> out-proj GEMM(+bias +residual) -> LN -> fc1 GEMM(+bias, GELU)
//            -> fc2 GEMM(+bias +residual).   LayerNorm, softmax, GELU and all accumulation are fp32.
#include <stdlib.h>
#include <string.h>

#include "zk_b200.h"
#include "zk_common.cuh"
#include "zk_internal.cuh"

namespace {
constexpr int HID = 768, MLP = 3072, QKV = 3 * HID, PATCH_K = 256;

struct LayerDev {
  __nv_bfloat16 *qkv_w, *o_w, *fc1_w, *fc2_w;
  float *qkv_b, *o_b, *fc1_b, *fc2_b, *ln1_w, *ln1_b, *ln2_w, *ln2_b;
};
}  // namespace

struct zk_model {
  int num_layers, max_length, num_labels, tokens, patches;
  float ln_eps;
  void* blob;  // one device allocation holding everything below
  size_t blob_bytes;
  __nv_bfloat16* patch_w;
  float *patch_b, *cls, *dist, *pos;
  LayerDev layer[ZK_AST_LAYERS];
  float *fln_w, *fln_b, *hln_w, *hln_b, *head_w, *head_b;
};

namespace zk {

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Carver {
  uint8_t* base;
  size_t off;
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

static void carve(zk_model* m, uint8_t* base, size_t* total) {
  Carver c{base, 0};
  m->patch_w = c.take<__nv_bfloat16>((size_t)HID * PATCH_K);
  m->patch_b = c.take<float>(HID);
  m->cls = c.take<float>(HID);
  m->dist = c.take<float>(HID);
  m->pos = c.take<float>((size_t)m->tokens * HID);
  for (int l = 0; l < m->num_layers; ++l) {
    LayerDev& L = m->layer[l];
    L.qkv_w = c.take<__nv_bfloat16>((size_t)QKV * HID);
    L.o_w = c.take<__nv_bfloat16>((size_t)HID * HID);
    L.fc1_w = c.take<__nv_bfloat16>((size_t)MLP * HID);
    L.fc2_w = c.take<__nv_bfloat16>((size_t)HID * MLP);
    L.qkv_b = c.take<float>(QKV);
    L.o_b = c.take<float>(HID);
    L.fc1_b = c.take<float>(MLP);
    L.fc2_b = c.take<float>(HID);
    L.ln1_w = c.take<float>(HID);
    L.ln1_b = c.take<float>(HID);
    L.ln2_w = c.take<float>(HID);
    L.ln2_b = c.take<float>(HID);
  }
  m->fln_w = c.take<float>(HID);
  m->fln_b = c.take<float>(HID);
  m->hln_w = c.take<float>(HID);
  m->hln_b = c.take<float>(HID);
  m->head_w = c.take<float>((size_t)m->num_labels * HID);
  m->head_b = c.take<float>(m->num_labels);
  *total = align_up(c.off, 256);
}

struct Workspace {
  float* x;
  __nv_bfloat16 *h, *qkv, *mlp;
};
static size_t carve_ws(const zk_model* m, int batch, uint8_t* base, Workspace* ws) {
  Carver c{base, 0};
  const size_t rows = (size_t)batch * m->tokens;
  Workspace w;
  w.x = c.take<float>(rows * HID);
  w.h = c.take<__nv_bfloat16>(rows * HID);
  w.qkv = c.take<__nv_bfloat16>(rows * QKV);
  w.mlp = c.take<__nv_bfloat16>(rows * MLP);
  if (ws) *ws = w;
  return align_up(c.off, 256);
}

static int forward_impl(zk_model* m, const GatherSrc& src, int batch, void* workspace, size_t workspace_bytes,
                        float* logits, float* hidden, cudaStream_t stream) {
  int rc = device_check();
  if (rc) return rc;
  if (!m || !logits || batch <= 0 || !workspace) {
    set_error("zk_model_forward: bad arguments");
    return ZK_ERR_ARG;
  }
  if (reinterpret_cast<uintptr_t>(workspace) % 256) {
    set_error("zk_model_forward: workspace must be 256-byte aligned");
    return ZK_ERR_ARG;
  }
  Workspace ws;
  const size_t need = carve_ws(m, batch, reinterpret_cast<uint8_t*>(workspace), &ws);
  if (workspace_bytes < need) {
    set_error("zk_model_forward: workspace %zu bytes < %zu needed for batch %d", workspace_bytes, need, batch);
    return ZK_ERR_WORKSPACE;
  }
  const long long rows = (long long)batch * m->tokens;
  const long long prow = (long long)batch * m->patches;
  // embeddings: gather patches -> GEMM (+bias +position) into token rows 2.., cls/dist rows 0,1
  if ((rc = gather_patches(src, batch, m->max_length, ws.mlp, stream))) return rc;
  if ((rc = gemm_bf16(ws.mlp, m->patch_w, m->patch_b, ws.x, prow, HID, PATCH_K, ZK_EPI_PATCH_F32, m->pos, m->patches, stream)))
    return rc;
  if ((rc = write_special_tokens(m->cls, m->dist, m->pos, ws.x, batch, m->tokens, stream))) return rc;
  // The classifier reads tokens 0 and 1 of the last hidden state only, so unless the caller asked for the full hidden
  // state the last layer runs K/V for every token and everything else for those two rows per window (identical
  // logits, 7.2 % fewer flops: 242.2 of 261.0 GFLOP per window are executed).  ZK_FULL_LAST_LAYER=1 disables it.
  static const bool full_last = getenv("ZK_FULL_LAST_LAYER") && atoi(getenv("ZK_FULL_LAST_LAYER")) != 0;
  const bool prune = !hidden && !full_last && m->num_layers >= 1;
  const int full_layers = prune ? m->num_layers - 1 : m->num_layers;
  for (int l = 0; l < full_layers; ++l) {
    const LayerDev& L = m->layer[l];
    if ((rc = layernorm_bf16(ws.x, L.ln1_w, L.ln1_b, m->ln_eps, ws.h, rows, HID, stream))) return rc;
    if ((rc = gemm_bf16(ws.h, L.qkv_w, L.qkv_b, ws.qkv, rows, QKV, HID, ZK_EPI_BIAS_BF16, nullptr, 0, stream))) return rc;
    if ((rc = attention_bf16(ws.qkv, ws.h, batch, m->tokens, stream))) return rc;
    if ((rc = gemm_bf16(ws.h, L.o_w, L.o_b, ws.x, rows, HID, HID, ZK_EPI_BIAS_RESID_F32, nullptr, 0, stream))) return rc;
    if ((rc = layernorm_bf16(ws.x, L.ln2_w, L.ln2_b, m->ln_eps, ws.h, rows, HID, stream))) return rc;
    if ((rc = gemm_bf16(ws.h, L.fc1_w, L.fc1_b, ws.mlp, rows, MLP, HID, ZK_EPI_BIAS_GELU_BF16, nullptr, 0, stream))) return rc;
    if ((rc = gemm_bf16(ws.mlp, L.fc2_w, L.fc2_b, ws.x, rows, HID, MLP, ZK_EPI_BIAS_RESID_F32, nullptr, 0, stream))) return rc;
  }
  if (!prune) {
    if (hidden) ZK_CUDA(cudaMemcpyAsync(hidden, ws.x, (size_t)rows * HID * 4, cudaMemcpyDeviceToDevice, stream));
    return head_logits(ws.x, batch, m->tokens, m->fln_w, m->fln_b, m->hln_w, m->hln_b, m->head_w, m->head_b,
                       m->num_labels, m->ln_eps, logits, stream);
  }
  {
    const LayerDev& L = m->layer[m->num_layers - 1];
    const long long r2 = 2LL * batch;
    // compact buffers of the two head rows per window live in the (otherwise idle) fc1 activation buffer
    Carver c{reinterpret_cast<uint8_t*>(ws.mlp), 0};
    __nv_bfloat16* hq = c.take<__nv_bfloat16>((size_t)r2 * HID);    // LN1 rows, later LN2 rows
    __nv_bfloat16* q2 = c.take<__nv_bfloat16>((size_t)r2 * HID);    // queries, later the attention output
    __nv_bfloat16* att2 = c.take<__nv_bfloat16>((size_t)r2 * HID);
    __nv_bfloat16* mlp2 = c.take<__nv_bfloat16>((size_t)r2 * MLP);
    float* x2 = c.take<float>((size_t)r2 * HID);
    if ((rc = layernorm_bf16(ws.x, L.ln1_w, L.ln1_b, m->ln_eps, ws.h, rows, HID, stream))) return rc;
    // K | V projections of every token, written in place into columns [768, 2304) of the fused QKV buffer
    if ((rc = gemm_bf16(ws.h, L.qkv_w + (size_t)HID * HID, L.qkv_b + HID, ws.qkv + HID, rows, 2 * HID, HID, ZK_EPI_BIAS_BF16,
                        nullptr, 0, stream, QKV, ZK_K_GEMM_QKV)))
      return rc;
    if ((rc = gather_head_rows(ws.h, ws.x, batch, m->tokens, hq, x2, stream))) return rc;
    if ((rc = gemm_bf16(hq, L.qkv_w, L.qkv_b, q2, r2, HID, HID, ZK_EPI_BIAS_BF16, nullptr, 0, stream, 0, ZK_K_TAIL))) return rc;
    if ((rc = attention_head_rows(q2, ws.qkv, att2, batch, m->tokens, stream))) return rc;
    if ((rc = gemm_bf16(att2, L.o_w, L.o_b, x2, r2, HID, HID, ZK_EPI_BIAS_RESID_F32, nullptr, 0, stream, 0, ZK_K_TAIL))) return rc;
    if ((rc = layernorm_bf16_cls(x2, L.ln2_w, L.ln2_b, m->ln_eps, hq, r2, HID, ZK_K_TAIL, stream))) return rc;
    if ((rc = gemm_bf16(hq, L.fc1_w, L.fc1_b, mlp2, r2, MLP, HID, ZK_EPI_BIAS_GELU_BF16, nullptr, 0, stream, 0, ZK_K_TAIL))) return rc;
    if ((rc = gemm_bf16(mlp2, L.fc2_w, L.fc2_b, x2, r2, HID, MLP, ZK_EPI_BIAS_RESID_F32, nullptr, 0, stream, 0, ZK_K_TAIL))) return rc;
    return head_logits(x2, batch, 2, m->fln_w, m->fln_b, m->hln_w, m->hln_b, m->head_w, m->head_b, m->num_labels, m->ln_eps,
                       logits, stream);
  }
}

}  // namespace zk

extern "C" {

int zk_model_create(const zk_ast_weights* w, zk_model** out) {
  using namespace zk;
  int rc = device_check();
  if (rc) return rc;
  if (!w || !out) {
    set_error("zk_model_create: null pointer");
    return ZK_ERR_ARG;
  }
  if (w->num_layers < 1 || w->num_layers > ZK_AST_LAYERS || w->num_labels < 1 || w->num_labels > 64 ||
      w->max_length < 16) {
    set_error("zk_model_create: unsupported geometry (layers %d labels %d max_length %d)", w->num_layers, w->num_labels,
              w->max_length);
    return ZK_ERR_SHAPE;
  }
  zk_model* m = new zk_model();
  memset(m, 0, sizeof(*m));
  m->num_layers = w->num_layers;
  m->max_length = w->max_length;
  m->num_labels = w->num_labels;
  m->ln_eps = w->ln_eps;
  m->patches = 12 * ((w->max_length - 16) / 10 + 1);
  m->tokens = m->patches + 2;
  carve(m, nullptr, &m->blob_bytes);
  cudaError_t e = cudaMalloc(&m->blob, m->blob_bytes);
  if (e != cudaSuccess) {
    delete m;
    return cuda_fail(e, "zk_model_create cudaMalloc");
  }
  size_t total;
  carve(m, reinterpret_cast<uint8_t*>(m->blob), &total);
  cudaStream_t s = 0;
  auto cp32 = [&](float* dst, const float* src, size_t n) {
    if (rc) return;
    if (!src) {
      set_error("zk_model_create: a weight pointer is null");
      rc = ZK_ERR_ARG;
      return;
    }
    cudaError_t ce = cudaMemcpyAsync(dst, src, n * 4, cudaMemcpyDeviceToDevice, s);
    if (ce != cudaSuccess) rc = cuda_fail(ce, "zk_model_create copy");
  };
  auto cv16 = [&](__nv_bfloat16* dst, const float* src, size_t n) {
    if (rc) return;
    if (!src) {
      set_error("zk_model_create: a weight pointer is null");
      rc = ZK_ERR_ARG;
      return;
    }
    rc = f32_to_bf16(src, dst, (long long)n, s);
  };
  cv16(m->patch_w, w->patch_w, (size_t)HID * PATCH_K);
  cp32(m->patch_b, w->patch_b, HID);
  cp32(m->cls, w->cls_token, HID);
  cp32(m->dist, w->dist_token, HID);
  cp32(m->pos, w->pos_emb, (size_t)m->tokens * HID);
  for (int l = 0; l < m->num_layers && !rc; ++l) {
    const zk_ast_layer_weights& W = w->layer[l];
    LayerDev& L = m->layer[l];
    cv16(L.qkv_w, W.q_w, (size_t)HID * HID);
    cv16(L.qkv_w + (size_t)HID * HID, W.k_w, (size_t)HID * HID);
    cv16(L.qkv_w + 2 * (size_t)HID * HID, W.v_w, (size_t)HID * HID);
    cp32(L.qkv_b, W.q_b, HID);
    cp32(L.qkv_b + HID, W.k_b, HID);
    cp32(L.qkv_b + 2 * HID, W.v_b, HID);
    cv16(L.o_w, W.o_w, (size_t)HID * HID);
    cp32(L.o_b, W.o_b, HID);
    cv16(L.fc1_w, W.fc1_w, (size_t)MLP * HID);
    cp32(L.fc1_b, W.fc1_b, MLP);
    cv16(L.fc2_w, W.fc2_w, (size_t)HID * MLP);
    cp32(L.fc2_b, W.fc2_b, HID);
    cp32(L.ln1_w, W.ln1_w, HID);
    cp32(L.ln1_b, W.ln1_b, HID);
    cp32(L.ln2_w, W.ln2_w, HID);
    cp32(L.ln2_b, W.ln2_b, HID);
  }
  cp32(m->fln_w, w->final_ln_w, HID);
  cp32(m->fln_b, w->final_ln_b, HID);
  cp32(m->hln_w, w->head_ln_w, HID);
  cp32(m->hln_b, w->head_ln_b, HID);
  cp32(m->head_w, w->head_w, (size_t)m->num_labels * HID);
  cp32(m->head_b, w->head_b, m->num_labels);
  if (!rc) {
    cudaError_t se = cudaStreamSynchronize(s);
    if (se != cudaSuccess) rc = cuda_fail(se, "zk_model_create sync");
  }
  if (rc) {
    zk_model_destroy(m);
    return rc;
  }
  *out = m;
  return 0;
}

void zk_model_destroy(zk_model* m) {
  if (!m) return;
  cudaFree(m->blob);
  delete m;
}

int zk_model_num_tokens(const zk_model* m) { return m ? m->tokens : 0; }

size_t zk_model_workspace_bytes(const zk_model* m, int batch) {
  if (!m || batch <= 0) return 0;
  return zk::carve_ws(m, batch, nullptr, nullptr);
}

int zk_model_forward(zk_model* m, const float* d_features, int batch, void* d_workspace, size_t workspace_bytes,
                     float* d_logits, float* d_hidden, zk_stream_t stream) {
  if (!d_features) {
    zk::set_error("zk_model_forward: d_features is null");
    return ZK_ERR_ARG;
  }
  zk::GatherSrc src;
  memset(&src, 0, sizeof(src));
  src.features = d_features;
  src.std2 = 1.f;
  return zk::forward_impl(m, src, batch, d_workspace, workspace_bytes, d_logits, d_hidden, (cudaStream_t)stream);
}

int zk_model_forward_fbank(zk_model* m, const float* d_fbank, int64_t fbank_frames, const int32_t* d_window_index,
                           int window_base, int frames_per_hop, int valid_frames, float mean, float std, int batch,
                           void* d_workspace, size_t workspace_bytes, float* d_logits, zk_stream_t stream) {
  if (!d_fbank || fbank_frames <= 0 || frames_per_hop <= 0 || valid_frames < 0) {
    zk::set_error("zk_model_forward_fbank: bad arguments");
    return ZK_ERR_ARG;
  }
  zk::GatherSrc src;
  memset(&src, 0, sizeof(src));
  src.fbank = d_fbank;
  src.fbank_frames = fbank_frames;
  src.window_index = d_window_index;
  src.window_base = window_base;
  src.frames_per_hop = frames_per_hop;
  src.valid_frames = valid_frames;
  src.mean = mean;
  src.std2 = std * 2.0f;
  return zk::forward_impl(m, src, batch, d_workspace, workspace_bytes, d_logits, nullptr, (cudaStream_t)stream);
}

}  // extern "C"
