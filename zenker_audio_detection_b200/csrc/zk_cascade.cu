// zk_cascade_run: the whole two-stage cascade of one recording behind ONE C entry point (replaces the per-recording
// body of ref:301-348 / refc:433-531 from "mono 16 kHz waveform in device memory" to "per-window scores").
//
//   continuous fbank -> Stage-1 FAST forward in batches (windows gathered from the fbank) -> decision re-check of the
//   windows within eps of a threshold (band select, RECHECK forward, scatter) -> softmax + gate + compaction ->
//   Stage-2 FAST forward on the compacted index list -> re-check -> softmax
//
// It is a plain host function over the other entry points of this library (nothing here is a kernel): a C or C++ host
// drives the path with zk_resample_* + this call.  The batch counts of the later steps depend on counts produced on the
// device, so the function synchronises `stream` at most four times (the two band counts and the gate count; see
// include/zk_b200.h) -- the one place in the ABI that does.
#include <math.h>
#include <string.h>

#include "zk_b200.h"
#include "zk_common.cuh"

namespace {

struct Carve {
  uint8_t* base;
  size_t off;
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) / 256 * 256;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

struct Layout {
  float *fbank, *logits1, *logits2, *hi;
  int32_t *pos, *win, *counts;
  void* model_ws;
  size_t model_ws_bytes, total;
};

int64_t num_windows(int64_t n_samples, int win, int hop) {  // ref:62-75 for n_samples >= win
  return (n_samples - win) / hop + 1;
}

Layout carve(const zk_model* m1, const zk_model* m2, int64_t n_samples, const zk_cascade_params* p, uint8_t* base) {
  Layout L;
  Carve c{base, 0};
  const int64_t n = num_windows(n_samples, p->window_samples, p->hop_samples);
  const int64_t frames = zk_fbank_num_frames(n_samples);
  L.fbank = c.take<float>((size_t)frames * 128);
  L.logits1 = c.take<float>((size_t)n * 2);
  L.logits2 = c.take<float>((size_t)n * 2);
  L.hi = c.take<float>((size_t)n * 2);
  L.pos = c.take<int32_t>((size_t)n);
  L.win = c.take<int32_t>((size_t)n);
  L.counts = c.take<int32_t>(4);
  size_t ws = 0;
  const zk_model* ms[2] = {m1, m2};
  for (const zk_model* m : ms) {
    const size_t a = zk_model_workspace_bytes(m, p->batch_size, ZK_PRECISION_FAST);
    const size_t b = p->recheck_eps > 0.f ? zk_model_workspace_bytes(m, p->recheck_batch, ZK_PRECISION_RECHECK) : 0;
    ws = ws > a ? ws : a;
    ws = ws > b ? ws : b;
  }
  L.model_ws_bytes = ws;
  L.model_ws = c.take<uint8_t>(ws);
  L.total = (c.off + 255) / 256 * 256;
  return L;
}

// logit(p): the margin l1 - l0 at which softmax(...)[1] == p; thresholds at or beyond 0 / 1 have no band
int add_margin(float* m, int n, float prob) {
  if (!(prob > 0.f) || !(prob < 1.f)) return n;
  const float v = (float)log((double)prob / (1.0 - (double)prob));
  for (int i = 0; i < n; ++i)
    if (fabsf(m[i] - v) <= 1e-12f) return n;
  m[n] = v;
  return n + 1;
}

int check_params(const zk_cascade_params* p) {
  if (!p || p->batch_size <= 0 || p->recheck_batch <= 0 || p->window_samples <= 0 || p->hop_samples <= 0 ||
      !(p->recheck_eps >= 0.f)) {
    zk::set_error("zk_cascade: bad parameters");
    return ZK_ERR_ARG;
  }
  if (p->hop_samples % 160) {
    zk::set_error("zk_cascade: hop of %d samples is not a multiple of the 10 ms frame shift (use the per-window entry points)",
                  p->hop_samples);
    return ZK_ERR_SHAPE;
  }
  return 0;
}

}  // namespace

extern "C" {

size_t zk_cascade_workspace_bytes(const zk_model* m1, const zk_model* m2, int64_t n_samples, const zk_cascade_params* p) {
  if (!m1 || !m2 || check_params(p) || n_samples < p->window_samples) return 0;
  return carve(m1, m2, n_samples, p, nullptr).total;
}

int zk_cascade_run(const zk_fbank_plan* plan, zk_model* m1, zk_model* m2, const float* d_audio16k, int64_t n_samples,
                   const zk_cascade_params* p, void* d_workspace, size_t workspace_bytes, float* d_probs1, int32_t* d_pred1,
                   int32_t* d_index, float* d_probs2, zk_cascade_counts* h_counts, zk_stream_t stream) {
  int rc = zk::device_check();
  if (rc) return rc;
  if ((rc = check_params(p))) return rc;
  if (!plan || !m1 || !m2 || !d_audio16k || !d_workspace || !d_probs1 || !d_pred1 || !d_index || !d_probs2 || !h_counts) {
    zk::set_error("zk_cascade_run: null pointer");
    return ZK_ERR_ARG;
  }
  if (n_samples < p->window_samples) {
    zk::set_error("zk_cascade_run: %lld samples are less than one window (%d); zero-pad first (ref:70-73)", (long long)n_samples,
                  p->window_samples);
    return ZK_ERR_SHAPE;
  }
  if (reinterpret_cast<uintptr_t>(d_workspace) % 256) {
    zk::set_error("zk_cascade_run: workspace must be 256-byte aligned");
    return ZK_ERR_ARG;
  }
  const Layout L = carve(m1, m2, n_samples, p, reinterpret_cast<uint8_t*>(d_workspace));
  if (workspace_bytes < L.total) {
    zk::set_error("zk_cascade_run: workspace %zu bytes < %zu needed", workspace_bytes, L.total);
    return ZK_ERR_WORKSPACE;
  }
  const int64_t n64 = num_windows(n_samples, p->window_samples, p->hop_samples);
  if (n64 > 0x7fffffffLL / 2) {
    zk::set_error("zk_cascade_run: too many windows (%lld)", (long long)n64);
    return ZK_ERR_SHAPE;
  }
  const int n = (int)n64;
  const int64_t frames = zk_fbank_num_frames(n_samples);
  const int frames_per_hop = p->hop_samples / 160;
  const int64_t wf = zk_fbank_num_frames(p->window_samples);
  const int max_len = zk_model_max_length(m1);
  if (zk_model_max_length(m2) != max_len) {
    zk::set_error("zk_cascade_run: the two models were built for different max_length (%d, %d)", max_len, zk_model_max_length(m2));
    return ZK_ERR_SHAPE;
  }
  const int valid_frames = (int)(wf < max_len ? wf : max_len);
  cudaStream_t s = (cudaStream_t)stream;
  memset(h_counts, 0, sizeof(*h_counts));
  h_counts->num_windows = n;

  if ((rc = zk_fbank_f32(plan, d_audio16k, n_samples, L.fbank, frames, stream))) return rc;

  // logits of `count` windows (window numbers index[i], or i when index is null) in batches at `precision`
  auto forward = [&](zk_model* m, float mean, float std, const int32_t* index, int count, int precision, float* logits) -> int {
    const int B = precision == ZK_PRECISION_RECHECK ? p->recheck_batch : p->batch_size;
    for (int base = 0; base < count; base += B) {
      const int b = count - base < B ? count - base : B;
      int r = zk_model_forward_fbank(m, L.fbank, frames, index ? index + base : nullptr, base, frames_per_hop, valid_frames,
                                     mean, std, b, precision, L.model_ws, L.model_ws_bytes, logits + 2 * (size_t)base, stream);
      if (r) return r;
    }
    return 0;
  };
  auto read_count = [&](const int32_t* d, int32_t* h) -> int {
    ZK_CUDA(cudaMemcpyAsync(h, d, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    ZK_CUDA(cudaStreamSynchronize(s));
    return 0;
  };
  // re-run the rows of `logits` (windows src[i] or i) that sit within eps of a decision point, overwrite them in place
  auto recheck = [&](zk_model* m, float mean, float std, float* logits, int count, const int32_t* src, const float* margins,
                     int nm, int32_t* h_n) -> int {
    *h_n = 0;
    if (!(p->recheck_eps > 0.f) || nm == 0 || count == 0) return 0;
    int r = zk_band_select(logits, count, margins, nm, p->recheck_eps, src, L.pos, L.win, L.counts, stream);
    if (r) return r;
    if ((r = read_count(L.counts, h_n))) return r;
    if (*h_n == 0) return 0;
    if ((r = forward(m, mean, std, L.win, *h_n, ZK_PRECISION_RECHECK, L.hi))) return r;
    return zk_scatter_rows2(L.hi, L.pos, *h_n, logits, stream);
  };

  float mg1[4], mg2[4];
  int n1 = add_margin(mg1, 0, 0.5f);  // argmax (ref:313, and the bare argmax of the summary, ref:156)
  n1 = add_margin(mg1, n1, p->thr1);
  if (p->min_prob >= 0.f) n1 = add_margin(mg1, n1, p->min_prob);
  const int n2 = add_margin(mg2, 0, p->stage2_argmax ? 0.5f : p->thr2);

  if ((rc = forward(m1, p->mean1, p->std1, nullptr, n, ZK_PRECISION_FAST, L.logits1))) return rc;
  if ((rc = recheck(m1, p->mean1, p->std1, L.logits1, n, nullptr, mg1, n1, &h_counts->rechecked_s1))) return rc;
  if ((rc = zk_gate_compact(L.logits1, n, p->thr1, p->min_prob, d_probs1, d_pred1, d_index, L.counts + 1, stream))) return rc;
  if ((rc = read_count(L.counts + 1, &h_counts->num_forwarded))) return rc;
  const int k = h_counts->num_forwarded;
  if (k > 0) {
    if ((rc = forward(m2, p->mean2, p->std2, d_index, k, ZK_PRECISION_FAST, L.logits2))) return rc;
    if ((rc = recheck(m2, p->mean2, p->std2, L.logits2, k, d_index, mg2, n2, &h_counts->rechecked_s2))) return rc;
    if ((rc = zk_softmax2(L.logits2, k, d_probs2, stream))) return rc;
  }
  return 0;
}

}  // extern "C"
