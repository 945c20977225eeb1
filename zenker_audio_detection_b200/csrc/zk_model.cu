// AST-base forward (HF:modeling_audio_spectrogram_transformer.py:403-451) as a fixed sequence of sm_100a kernels.
//
// zk_model owns the packed weights: 16-bit [out][in] matrices for the tcgen05 GEMMs (query/key/value fused into one
// [2304][768] matrix) stored as fp16 hi | lo planes of the power-of-two-scaled weight, fp32 biases, LayerNorm
// parameters, cls/dist tokens and the position table.
// Activations live in the caller's workspace (16 = fp16 or bf16; P = 1 plane FAST, 2 planes hi | lo at RECHECK):
//   x    fp32 [B*T][768]     residual stream (kept in fp32 end to end)
//   h    16   [B*T][P*768]   LayerNorm output / attention output (GEMM A operands)
//   qkv  16   [B*T][P*2304]  fused projection output, read in place by the attention kernel
//   mlp  16   [B*T][P*3072]  GELU(fc1) output; its head doubles as the patch-gather matrix [B*patches][P*256]
// Per layer: LN -> QKV GEMM(+bias) -> attention -> out-proj GEMM(+bias +residual) -> LN -> fc1 GEMM(+bias, GELU)
//            -> fc2 GEMM(+bias +residual).   LayerNorm, softmax, GELU and all accumulation are fp32.
// ZK_PRECISION_RECHECK runs the same sequence with split operands (three-product contractions, exact erf / exp2),
// see include/zk_b200.h and DESIGN.md section 4b.
#include <stdlib.h>
#include <string.h>

#include "zk_b200.h"
#include "zk_common.cuh"
#include "zk_internal.cuh"

namespace {
constexpr int HID = 768, MLP = 3072, QKV = 3 * HID, PATCH_K = 256;
constexpr int ZK_LN_FUSE_DEFAULT = 0;  // set by measurement, see DESIGN.md section 6

// One nn.Linear weight [N][K] as the kernels read it:
//   planes  fp16 [N][2K] = hi | lo of W * 2^e (e chosen so that max |W| 2^e lies in (2^13, 2^14]: the lo plane stays in
//           fp16's normal range down to |w| = 2^-17 max |W|); inv_scale = 2^-e is applied to the accumulator.
//           The FAST path in fp16 reads the hi plane only (row pitch 2K), the RECHECK path both.
//   bf16    [N][K], unscaled: only allocated when the FAST path was asked to run in bf16.
struct WeightDev {
  __half* planes;
  __nv_bfloat16* bf16;
  float inv_scale;
};
struct LayerDev {
  WeightDev qkv, o, fc1, fc2;
  float *qkv_b, *o_b, *fc1_b, *fc2_b, *ln1_w, *ln1_b, *ln2_w, *ln2_b;
};
}  // namespace

struct zk_model {
  int num_layers, max_length, num_labels, tokens, patches, fmt;
  float ln_eps;
  void* blob;  // one device allocation holding everything below
  size_t blob_bytes;
  WeightDev patch;
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
  const bool bf = m->fmt == FMT_BF16;
  auto weight = [&](WeightDev& w, size_t n, size_t k) {
    w.planes = c.take<__half>(n * 2 * k);
    w.bf16 = bf ? c.take<__nv_bfloat16>(n * k) : nullptr;
  };
  weight(m->patch, HID, PATCH_K);
  m->patch_b = c.take<float>(HID);
  m->cls = c.take<float>(HID);
  m->dist = c.take<float>(HID);
  m->pos = c.take<float>((size_t)m->tokens * HID);
  for (int l = 0; l < m->num_layers; ++l) {
    LayerDev& L = m->layer[l];
    weight(L.qkv, QKV, HID);
    weight(L.o, HID, HID);
    weight(L.fc1, MLP, HID);
    weight(L.fc2, HID, MLP);
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

// Activations: x fp32 residual stream; h / qkv / mlp 16-bit GEMM operands with `planes` planes per row (1 FAST, 2 RECHECK)
struct Workspace {
  float* x;
  uint16_t *h, *qkv, *mlp;
  int* ln_count;  // scratch of the residual GEMMs' LayerNorm tail (zk_gemm.cu)
};
static size_t carve_ws(const zk_model* m, int batch, int planes, uint8_t* base, Workspace* ws) {
  Carver c{base, 0};
  const size_t rows = (size_t)batch * m->tokens;
  Workspace w;
  w.x = c.take<float>(rows * HID);
  w.h = c.take<uint16_t>(rows * HID * planes);
  w.qkv = c.take<uint16_t>(rows * QKV * planes);
  w.mlp = c.take<uint16_t>(rows * MLP * planes);
  w.ln_count = c.take<int>(gemm_ln_tail_counters((long long)rows));
  if (ws) *ws = w;
  return align_up(c.off, 256);
}

static int forward_impl(zk_model* m, const GatherSrc& src, int batch, int precision, void* workspace, size_t workspace_bytes,
                        float* logits, float* hidden, cudaStream_t stream) {
  int rc = device_check();
  if (rc) return rc;
  if (!m || !logits || batch <= 0 || !workspace) {
    set_error("zk_model_forward: bad arguments");
    return ZK_ERR_ARG;
  }
  if (precision != ZK_PRECISION_FAST && precision != ZK_PRECISION_RECHECK) {
    set_error("zk_model_forward: unknown precision %d", precision);
    return ZK_ERR_ARG;
  }
  if (reinterpret_cast<uintptr_t>(workspace) % 256) {
    set_error("zk_model_forward: workspace must be 256-byte aligned");
    return ZK_ERR_ARG;
  }
  const bool hp = precision == ZK_PRECISION_RECHECK;
  const int planes = hp ? 2 : 1;
  const int fmt = hp ? FMT_F16 : m->fmt;  // format of the activation operands
  Workspace ws;
  const size_t need = carve_ws(m, batch, planes, reinterpret_cast<uint8_t*>(workspace), &ws);
  if (workspace_bytes < need) {
    set_error("zk_model_forward: workspace %zu bytes < %zu needed for batch %d at precision %d", workspace_bytes, need, batch,
              precision);
    return ZK_ERR_WORKSPACE;
  }
  ProfClassOverride prof_override(hp ? ZK_K_RECHECK : -1);
  const long long rows = (long long)batch * m->tokens;
  const long long prow = (long long)batch * m->patches;
  // out = epilogue(A W^T): A has `planes` planes of K columns; W is read as the precision asks
  // ln_w != nullptr: the residual GEMM also produces h = LayerNorm(out) (ln_w, ln_b) for the next GEMM (LayerNorm tail)
  auto gemm = [&](const void* a, const WeightDev& w, size_t w_row0, const float* bias, void* out, long long ldo, long long M,
                  int N, int K, int epilogue, int prof_cls, const float* ln_w = nullptr, const float* ln_b = nullptr) {
    GemmArgs g;
    if (ln_w) {
      g.ln_w = ln_w, g.ln_b = ln_b, g.ln_out = ws.h, g.ln_count = ws.ln_count, g.ln_eps = m->ln_eps;
    }
    g.a = a;
    g.lda = (long long)planes * K;
    if (!hp && m->fmt == FMT_BF16) {
      g.w = w.bf16 + w_row0 * K;
      g.ldw = K;
      g.acc_scale = 1.0f;
    } else {
      g.w = w.planes + w_row0 * 2 * K;
      g.ldw = 2LL * K;
      g.acc_scale = w.inv_scale;
    }
    g.bias = bias;
    g.out = out;
    g.ldo = ldo;
    g.M = M, g.N = N, g.K = K;
    g.epilogue = epilogue;
    g.fmt = fmt;
    g.products = hp ? 3 : 1;
    g.aux = m->pos;
    g.aux_rows = m->patches;
    g.prof_cls = prof_cls;
    return gemm16(g, stream);
  };
  const int EPI_LIN = hp ? ZK_EPI_BIAS_SPLIT : ZK_EPI_BIAS_BF16;
  const int EPI_GELU = hp ? ZK_EPI_BIAS_GELU_SPLIT : ZK_EPI_BIAS_GELU_BF16;
  // embeddings: gather patches -> GEMM (+bias +position) into token rows 2.., cls/dist rows 0,1
  if ((rc = gather_patches(src, batch, m->max_length, ws.mlp, fmt, planes, stream))) return rc;
  if ((rc = gemm(ws.mlp, m->patch, 0, m->patch_b, ws.x, 0, prow, HID, PATCH_K, ZK_EPI_PATCH_F32, -1))) return rc;
  if ((rc = write_special_tokens(m->cls, m->dist, m->pos, ws.x, batch, m->tokens, stream))) return rc;
  // The classifier reads tokens 0 and 1 of the last hidden state only, so unless the caller asked for the full hidden
  // state the last layer runs K/V for every token and everything else for those two rows per window (identical
  // logits, 7.2 % fewer flops: 242.2 of 261.0 GFLOP per window are executed).  ZK_FULL_LAST_LAYER=1 disables it; the
  // re-check precision always runs the full layer (it sees a few windows per recording).
  static const bool full_last = getenv("ZK_FULL_LAST_LAYER") && atoi(getenv("ZK_FULL_LAST_LAYER")) != 0;
  const bool prune = !hp && !hidden && !full_last && m->num_layers >= 1;
  const int full_layers = prune ? m->num_layers - 1 : m->num_layers;
  // LayerNorm fused into the residual GEMM that produces its input (north_star (4), zk_gemm.cu "LayerNorm tail"):
  // ZK_LN_FUSE bit 0 = fc2 -> layernorm_before of the next layer, bit 1 = out-projection -> layernorm_after.
  // FAST precision only; bit-identical to the separate kernel (same row arithmetic).
  static const int ln_fuse_env = getenv("ZK_LN_FUSE") ? atoi(getenv("ZK_LN_FUSE")) : ZK_LN_FUSE_DEFAULT;
  const int ln_fuse = (!hp && gemm_ln_tail_ok(rows)) ? ln_fuse_env : 0;
  bool h_is_ln1 = false;  // ws.h already holds layernorm_before of the coming layer
  for (int l = 0; l < full_layers; ++l) {
    const LayerDev& L = m->layer[l];
    if (!h_is_ln1 && (rc = layernorm16(ws.x, L.ln1_w, L.ln1_b, m->ln_eps, ws.h, rows, HID, fmt, planes, -1, stream))) return rc;
    h_is_ln1 = false;
    if ((rc = gemm(ws.h, L.qkv, 0, L.qkv_b, ws.qkv, (long long)planes * QKV, rows, QKV, HID, EPI_LIN, -1))) return rc;
    if (hp) {
      if ((rc = attention_split(ws.qkv, ws.h, batch, m->tokens, stream))) return rc;
    } else {
      if ((rc = attention16(ws.qkv, ws.h, batch, m->tokens, fmt, stream))) return rc;
    }
    if (ln_fuse & 2) {
      // the tail overwrites ws.h (this GEMM's A operand) row block by row block, each only after every column tile of
      // the block has run its whole K loop
      if ((rc = gemm(ws.h, L.o, 0, L.o_b, ws.x, 0, rows, HID, HID, ZK_EPI_BIAS_RESID_F32, -1, L.ln2_w, L.ln2_b))) return rc;
    } else {
      if ((rc = gemm(ws.h, L.o, 0, L.o_b, ws.x, 0, rows, HID, HID, ZK_EPI_BIAS_RESID_F32, -1))) return rc;
      if ((rc = layernorm16(ws.x, L.ln2_w, L.ln2_b, m->ln_eps, ws.h, rows, HID, fmt, planes, -1, stream))) return rc;
    }
    if ((rc = gemm(ws.h, L.fc1, 0, L.fc1_b, ws.mlp, (long long)planes * MLP, rows, MLP, HID, EPI_GELU, -1))) return rc;
    const LayerDev* next = l + 1 < m->num_layers ? &m->layer[l + 1] : nullptr;  // (the pruned last layer starts with its LN1 too)
    if ((ln_fuse & 1) && next) {
      if ((rc = gemm(ws.mlp, L.fc2, 0, L.fc2_b, ws.x, 0, rows, HID, MLP, ZK_EPI_BIAS_RESID_F32, -1, next->ln1_w, next->ln1_b)))
        return rc;
      h_is_ln1 = true;
    } else {
      if ((rc = gemm(ws.mlp, L.fc2, 0, L.fc2_b, ws.x, 0, rows, HID, MLP, ZK_EPI_BIAS_RESID_F32, -1))) return rc;
    }
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
    uint16_t* hq = c.take<uint16_t>((size_t)r2 * HID);    // LN1 rows, later LN2 rows
    uint16_t* q2 = c.take<uint16_t>((size_t)r2 * HID);    // queries, later the attention output
    uint16_t* att2 = c.take<uint16_t>((size_t)r2 * HID);
    uint16_t* mlp2 = c.take<uint16_t>((size_t)r2 * MLP);
    float* x2 = c.take<float>((size_t)r2 * HID);
    if (!h_is_ln1 && (rc = layernorm16(ws.x, L.ln1_w, L.ln1_b, m->ln_eps, ws.h, rows, HID, fmt, 1, -1, stream))) return rc;
    // K | V projections of every token, written in place into columns [768, 2304) of the fused QKV buffer
    if ((rc = gemm(ws.h, L.qkv, HID, L.qkv_b + HID, ws.qkv + HID, QKV, rows, 2 * HID, HID, ZK_EPI_BIAS_BF16, ZK_K_GEMM_QKV)))
      return rc;
    if ((rc = gather_head_rows(ws.h, ws.x, batch, m->tokens, hq, x2, stream))) return rc;
    if ((rc = gemm(hq, L.qkv, 0, L.qkv_b, q2, 0, r2, HID, HID, ZK_EPI_BIAS_BF16, ZK_K_TAIL))) return rc;
    if ((rc = attention_head_rows(q2, ws.qkv, att2, batch, m->tokens, fmt, stream))) return rc;
    if ((rc = gemm(att2, L.o, 0, L.o_b, x2, 0, r2, HID, HID, ZK_EPI_BIAS_RESID_F32, ZK_K_TAIL))) return rc;
    if ((rc = layernorm16(x2, L.ln2_w, L.ln2_b, m->ln_eps, hq, r2, HID, fmt, 1, ZK_K_TAIL, stream))) return rc;
    if ((rc = gemm(hq, L.fc1, 0, L.fc1_b, mlp2, 0, r2, MLP, HID, ZK_EPI_BIAS_GELU_BF16, ZK_K_TAIL))) return rc;
    if ((rc = gemm(mlp2, L.fc2, 0, L.fc2_b, x2, 0, r2, HID, MLP, ZK_EPI_BIAS_RESID_F32, ZK_K_TAIL))) return rc;
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
  if (w->operand_format != ZK_FMT_BF16 && w->operand_format != ZK_FMT_F16) {
    set_error("zk_model_create: unknown operand format %d", w->operand_format);
    return ZK_ERR_ARG;
  }
  zk_model* m = new zk_model();
  memset(m, 0, sizeof(*m));
  m->num_layers = w->num_layers;
  m->max_length = w->max_length;
  m->num_labels = w->num_labels;
  m->ln_eps = w->ln_eps;
  m->fmt = w->operand_format;
  m->patches = 12 * ((w->max_length - 16) / 10 + 1);
  m->tokens = m->patches + 2;
  carve(m, nullptr, &m->blob_bytes);
  cudaError_t e = cudaMalloc(&m->blob, m->blob_bytes + 256);  // + a float for the max-|w| reductions
  if (e != cudaSuccess) {
    delete m;
    return cuda_fail(e, "zk_model_create cudaMalloc");
  }
  size_t total;
  carve(m, reinterpret_cast<uint8_t*>(m->blob), &total);
  float* d_max = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(m->blob) + m->blob_bytes);
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
  // One GEMM weight made of `parts` row blocks of [n_each][k] (query | key | value for the fused projection): a common
  // power-of-two scale from the largest magnitude, then hi | lo fp16 planes (and the plain bf16 copy if asked for).
  auto weight = [&](WeightDev& dst, const float* const* srcs, int parts, size_t n_each, size_t k) {
    if (rc) return;
    float mx = 0.f;
    for (int i = 0; i < parts && !rc; ++i) {
      if (!srcs[i]) {
        set_error("zk_model_create: a weight pointer is null");
        rc = ZK_ERR_ARG;
        return;
      }
      float v = 0.f;
      if ((rc = max_abs(srcs[i], (long long)(n_each * k), d_max, s))) return;
      cudaError_t ce = cudaMemcpyAsync(&v, d_max, sizeof(float), cudaMemcpyDeviceToHost, s);
      if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
      if (ce != cudaSuccess) {
        rc = cuda_fail(ce, "zk_model_create max |w|");
        return;
      }
      if (!(v == v) || v > 3.0e38f) {
        set_error("zk_model_create: a weight tensor holds NaN or Inf");
        rc = ZK_ERR_ARG;
        return;
      }
      mx = v > mx ? v : mx;
    }
    int ex = 0;
    if (mx > 0.f) {
      frexpf(mx, &ex);        // mx = f * 2^ex, f in [0.5, 1)
      ex = 14 - ex;           // mx * 2^ex in [2^13, 2^14)
      if (ex > 60) ex = 60;
      if (ex < -60) ex = -60;
    }
    const float scale = ldexpf(1.0f, ex);
    dst.inv_scale = ldexpf(1.0f, -ex);
    for (int i = 0; i < parts && !rc; ++i) {
      rc = f32_to_16(srcs[i], dst.planes + (size_t)i * n_each * 2 * k, (long long)n_each, (int)k, FMT_F16, 2, scale, s);
      if (!rc && dst.bf16) rc = f32_to_16(srcs[i], dst.bf16 + (size_t)i * n_each * k, (long long)n_each, (int)k, FMT_BF16, 1, 1.0f, s);
    }
  };
  auto weight1 = [&](WeightDev& dst, const float* src, size_t n, size_t k) {
    const float* one[1] = {src};
    weight(dst, one, 1, n, k);
  };
  weight1(m->patch, w->patch_w, HID, PATCH_K);
  cp32(m->patch_b, w->patch_b, HID);
  cp32(m->cls, w->cls_token, HID);
  cp32(m->dist, w->dist_token, HID);
  cp32(m->pos, w->pos_emb, (size_t)m->tokens * HID);
  for (int l = 0; l < m->num_layers && !rc; ++l) {
    const zk_ast_layer_weights& W = w->layer[l];
    LayerDev& L = m->layer[l];
    const float* qkv[3] = {W.q_w, W.k_w, W.v_w};
    weight(L.qkv, qkv, 3, HID, HID);
    cp32(L.qkv_b, W.q_b, HID);
    cp32(L.qkv_b + HID, W.k_b, HID);
    cp32(L.qkv_b + 2 * HID, W.v_b, HID);
    weight1(L.o, W.o_w, HID, HID);
    cp32(L.o_b, W.o_b, HID);
    weight1(L.fc1, W.fc1_w, MLP, HID);
    cp32(L.fc1_b, W.fc1_b, MLP);
    weight1(L.fc2, W.fc2_w, HID, MLP);
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
int zk_model_max_length(const zk_model* m) { return m ? m->max_length : 0; }

size_t zk_model_workspace_bytes(const zk_model* m, int batch, int precision) {
  if (!m || batch <= 0) return 0;
  return zk::carve_ws(m, batch, precision == ZK_PRECISION_RECHECK ? 2 : 1, nullptr, nullptr);
}

int zk_model_forward(zk_model* m, const float* d_features, const int32_t* d_row_index, int batch, int precision,
                     void* d_workspace, size_t workspace_bytes, float* d_logits, float* d_hidden, zk_stream_t stream) {
  if (!d_features) {
    zk::set_error("zk_model_forward: d_features is null");
    return ZK_ERR_ARG;
  }
  zk::GatherSrc src;
  memset(&src, 0, sizeof(src));
  src.features = d_features;
  src.window_index = d_row_index;
  src.std2 = 1.f;
  return zk::forward_impl(m, src, batch, precision, d_workspace, workspace_bytes, d_logits, d_hidden, (cudaStream_t)stream);
}

int zk_model_forward_fbank(zk_model* m, const float* d_fbank, int64_t fbank_frames, const int32_t* d_window_index,
                           int window_base, int frames_per_hop, int valid_frames, float mean, float std, int batch,
                           int precision, void* d_workspace, size_t workspace_bytes, float* d_logits, zk_stream_t stream) {
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
  return zk::forward_impl(m, src, batch, precision, d_workspace, workspace_bytes, d_logits, nullptr, (cudaStream_t)stream);
}

}  // extern "C"
