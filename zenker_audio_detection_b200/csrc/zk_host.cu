// Host-side plumbing of libzk_b200: thread-local error string, device check, TMA tensor maps.
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "zk_b200.h"
#include "zk_common.cuh"
#include "zk_internal.cuh"

namespace zk {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return (int)e;
}

static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static int g_check = 1;  // 1 = not yet run
static std::once_flag g_once;

static void do_device_check() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    g_check = cuda_fail(e, "cudaGetDevice");
    return;
  }
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) {
    g_check = cuda_fail(e, "cudaGetDeviceProperties");
    return;
  }
  if (p.major != 10) {
    set_error("device %s is sm_%d%d; libzk_b200 contains sm_100a code only (no fallback path)", p.name, p.major, p.minor);
    g_check = ZK_ERR_UNSUPPORTED;
    return;
  }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || fn == nullptr) {
    set_error("driver entry point cuTensorMapEncodeTiled not available");
    g_check = ZK_ERR_UNSUPPORTED;
    return;
  }
  g_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  g_check = 0;
}

int device_check() {
  std::call_once(g_once, do_device_check);
  if (g_check != 0 && g_err[0] == 0) set_error("device check failed earlier in this process (code %d)", g_check);
  return g_check;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static int make_tmap_2d(CUtensorMap* out, CUtensorMapDataType dt, uint32_t esize, const void* base, uint64_t rows,
                        uint64_t cols, uint64_t pitch_elems, uint32_t box_rows, uint32_t box_cols) {
  int rc = device_check();
  if (rc) return rc;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (pitch_elems * esize) % 16 || box_cols * esize != 128 || box_rows > 256) {
    set_error("make_tmap_2d: bad alignment/box (base %p pitch %llu box %ux%u esize %u)", base,
              (unsigned long long)pitch_elems, box_rows, box_cols, esize);
    return ZK_ERR_ARG;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {pitch_elems * esize};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(out, dt, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows %llu cols %llu pitch %llu box %ux%u", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)pitch_elems, box_rows, box_cols);
    return ZK_ERR_INTERNAL;
  }
  return 0;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                      uint32_t box_rows, uint32_t box_cols) {
  return make_tmap_2d(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, rows, cols, pitch_elems, box_rows, box_cols);
}
int make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t slabs, uint64_t rows, uint64_t cols,
                      uint64_t pitch_elems, uint64_t slab_pitch_elems, uint32_t box_rows, uint32_t box_cols) {
  int rc = device_check();
  if (rc) return rc;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (pitch_elems * 2) % 16 || (slab_pitch_elems * 2) % 16 || box_cols != 64 ||
      box_rows > 256) {
    set_error("make_tmap_bf16_3d: bad alignment/box");
    return ZK_ERR_ARG;
  }
  cuuint64_t gdim[3] = {cols, rows, slabs};
  cuuint64_t gstr[2] = {pitch_elems * 2, slab_pitch_elems * 2};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (3d) failed (%d) slabs %llu rows %llu cols %llu", (int)r, (unsigned long long)slabs,
              (unsigned long long)rows, (unsigned long long)cols);
    return ZK_ERR_INTERNAL;
  }
  return 0;
}
int make_tmap_f32_3d(CUtensorMap* out, const void* base, uint64_t slabs, uint64_t rows, uint64_t cols,
                     uint64_t pitch_elems, uint64_t slab_pitch_elems, uint32_t box_rows, uint32_t box_cols) {
  int rc = device_check();
  if (rc) return rc;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (pitch_elems * 4) % 16 || (slab_pitch_elems * 4) % 16 || box_cols != 32 ||
      box_rows > 256) {
    set_error("make_tmap_f32_3d: bad alignment/box");
    return ZK_ERR_ARG;
  }
  cuuint64_t gdim[3] = {cols, rows, slabs};
  cuuint64_t gstr[2] = {pitch_elems * 4, slab_pitch_elems * 4};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (f32 3d) failed (%d)", (int)r);
    return ZK_ERR_INTERNAL;
  }
  return 0;
}
int make_tmap_f32_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                     uint32_t box_rows, uint32_t box_cols) {
  return make_tmap_2d(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, rows, cols, pitch_elems, box_rows, box_cols);
}

int ensure_dynamic_smem(const void* kernel, int bytes, unsigned long long* done) {
  int dev = 0;
  ZK_CUDA(cudaGetDevice(&dev));
  const unsigned long long bit = 1ull << (dev & 63);
  if (__atomic_load_n(done, __ATOMIC_ACQUIRE) & bit) return 0;
  ZK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  __atomic_fetch_or(done, bit, __ATOMIC_RELEASE);
  return 0;
}

// ---------------------------------------------------------------------------------------------- profiler
struct ProfRec {
  int cls;
  cudaEvent_t a, b;
};
static std::mutex g_prof_mu;
static bool g_prof_time = false;
static long long g_prof_count[ZK_K_NUM_CLASSES] = {0};
static std::vector<ProfRec> g_prof_recs;
static std::vector<cudaEvent_t> g_prof_pool;

static cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) {
    cudaEvent_t e = g_prof_pool.back();
    g_prof_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

static thread_local int g_prof_override = -1;
ProfClassOverride::ProfClassOverride(int cls) : prev_(g_prof_override) { g_prof_override = cls; }
ProfClassOverride::~ProfClassOverride() { g_prof_override = prev_; }

ProfScope::ProfScope(int cls, cudaStream_t stream) : cls_(cls), stream_(stream), stop_(nullptr) {
  if (g_prof_override >= 0) cls = cls_ = g_prof_override;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_count[cls]++;
  if (!g_prof_time) return;
  ProfRec r{cls, prof_event(), prof_event()};
  if (!r.a || !r.b) return;
  cudaEventRecord(r.a, stream);
  stop_ = r.b;
  g_prof_recs.push_back(r);
}
ProfScope::~ProfScope() {
  if (stop_) cudaEventRecord((cudaEvent_t)stop_, stream_);
}

}  // namespace zk

extern "C" {
void zk_prof_enable(int time_launches) {
  std::lock_guard<std::mutex> lk(zk::g_prof_mu);
  zk::g_prof_time = time_launches != 0;
}

int zk_prof_collect(float* ms, int64_t* launches) {
  cudaError_t e = cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(zk::g_prof_mu);
  for (int i = 0; i < ZK_K_NUM_CLASSES; ++i) {
    if (ms) ms[i] = 0.f;
    if (launches) launches[i] = zk::g_prof_count[i];
    zk::g_prof_count[i] = 0;
  }
  for (auto& r : zk::g_prof_recs) {
    float t = 0.f;
    if (e == cudaSuccess && cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess && ms) ms[r.cls] += t;
    zk::g_prof_pool.push_back(r.a);
    zk::g_prof_pool.push_back(r.b);
  }
  zk::g_prof_recs.clear();
  if (e != cudaSuccess) return zk::cuda_fail(e, "zk_prof_collect");
  return 0;
}

const char* zk_kernel_class_name(int cls) {
  static const char* names[ZK_K_NUM_CLASSES] = {"resample", "fbank", "gather_patches", "gemm_patch", "layernorm",
                                               "gemm_qkv", "attention", "gemm_out", "gemm_fc1", "gemm_fc2",
                                               "head", "gate", "misc", "last_layer_tail", "recheck"};
  return (cls >= 0 && cls < ZK_K_NUM_CLASSES) ? names[cls] : "?";
}

int zk_abi_version(void) { return ZK_ABI_VERSION; }
const char* zk_last_error_string(void) { return zk::g_err; }
int zk_device_check(void) { return zk::device_check(); }
}
