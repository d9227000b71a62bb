// Host-side plumbing shared by every entry point: last-error string, TMA tensor-map encoding through the
// driver entry point, device queries.
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace f5b {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// Encoding a tensor map costs a few microseconds of host time; the hot loops re-use a small set of (pointer, shape) pairs
// (same workspace, same weights every ODE step), so encoded maps are cached.  Key = every input of the encode call.
struct TmapKey {
  uint64_t v[10];
  bool operator==(const TmapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t x : k.v) { h ^= x; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmaps;
static std::mutex g_tmap_mu;
static bool tmap_lookup(const TmapKey& k, CUtensorMap* out) {
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  auto it = g_tmaps.find(k);
  if (it == g_tmaps.end()) return false;
  *out = it->second;
  return true;
}
static void tmap_store(const TmapKey& k, const CUtensorMap& m) {
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  if (g_tmaps.size() > 65536) g_tmaps.clear();
  g_tmaps.emplace(k, m);
}

static CUtensorMapDataType dtype_of(int elt_bytes) {
  return elt_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
}

int make_tmap_2d(CUtensorMap* out, const void* base, int elt_bytes, uint64_t inner, uint64_t outer,
                 uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer, bool swizzle128) {
  EncodeTiledFn enc = get_encode();
  F5B_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
  F5B_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base %p not 16B aligned", base);
  F5B_CHECK(row_stride_bytes % 16 == 0, "TMA row stride %llu not a multiple of 16", (unsigned long long)row_stride_bytes);
  const TmapKey key = {{2, (uint64_t)(uintptr_t)base, (uint64_t)elt_bytes, inner, outer, 0, row_stride_bytes, 0,
                        ((uint64_t)box_inner << 32) | box_outer, (uint64_t)swizzle128}};
  if (tmap_lookup(key, out)) return 0;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, dtype_of(elt_bytes), 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  F5B_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d) failed: %d (dims %llu x %llu, stride %llu, box %u x %u)", (int)r,
            (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes, box_inner, box_outer);
  tmap_store(key, *out);
  return 0;
}

int make_tmap_3d(CUtensorMap* out, const void* base, int elt_bytes, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2,
                 bool swizzle128) {
  EncodeTiledFn enc = get_encode();
  F5B_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
  F5B_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base %p not 16B aligned", base);
  F5B_CHECK(stride1_bytes % 16 == 0 && stride2_bytes % 16 == 0, "TMA strides must be multiples of 16");
  const TmapKey key = {{3, (uint64_t)(uintptr_t)base, (uint64_t)elt_bytes, d0, d1, d2, stride1_bytes, stride2_bytes,
                        ((uint64_t)b0 << 40) | ((uint64_t)b1 << 20) | b2, (uint64_t)swizzle128}};
  if (tmap_lookup(key, out)) return 0;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, dtype_of(elt_bytes), 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  F5B_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed: %d", (int)r);
  tmap_store(key, *out);
  return 0;
}

// ---------------------------------------------------------------- launch accounting / profiling
struct ProfState {
  bool on = false;
  std::vector<cudaEvent_t> ev;          // pairs (start, stop)
  std::vector<int> ev_kind;
  size_t used = 0;
  unsigned long long launches[K_NUM] = {0};
  double flops[K_NUM] = {0}, bytes[K_NUM] = {0};
};
static ProfState g_prof;
static std::mutex g_prof_mu;

static thread_local int g_kind_override = -1;
KindOverride::KindOverride(int kind) : prev(g_kind_override) { g_kind_override = kind; }
KindOverride::~KindOverride() { g_kind_override = prev; }

LaunchScope::LaunchScope(int kind_, cudaStream_t s_, double fl, double by, int n) : kind(kind_), s(s_), slot(-1) {
  if (g_kind_override >= 0 && kind != K_SPECTRAL) kind = g_kind_override;  // (the iSTFT stays in "spectral")
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.launches[kind] += n;
  g_prof.flops[kind] += fl;
  g_prof.bytes[kind] += by;
  if (!g_prof.on) return;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;  // no events inside a capture
  if (g_prof.used + 2 > g_prof.ev.size()) {
    const size_t add = 4096;
    for (size_t i = 0; i < add; ++i) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return;
      g_prof.ev.push_back(e);
    }
    g_prof.ev_kind.resize(g_prof.ev.size() / 2);
  }
  slot = (int)(g_prof.used / 2);
  g_prof.ev_kind[slot] = kind;
  cudaEventRecord(g_prof.ev[g_prof.used], s);
  g_prof.used += 2;
}
LaunchScope::~LaunchScope() {
  if (slot >= 0) cudaEventRecord(g_prof.ev[2 * slot + 1], s);
}

int g_pdl = 0;  // f5b_set_dependent_launch(): the Python host turns it on for the launch-bound regimes (see include/f5b200.h)

int sm_count() {
  static int cache[64] = {0};  // per device: a process may drive several GPUs
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  int n = cache[dev];
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    cache[dev] = n;
  }
  return n;
}

}  // namespace f5b

extern "C" {

// Profiling / accounting interface used by bench.py (not on the product path).
void f5b_prof_reset(int enable) {
  std::lock_guard<std::mutex> lk(f5b::g_prof_mu);
  f5b::g_prof.on = enable != 0;
  f5b::g_prof.used = 0;
  for (int k = 0; k < f5b::K_NUM; ++k) {
    f5b::g_prof.launches[k] = 0;
    f5b::g_prof.flops[k] = 0;
    f5b::g_prof.bytes[k] = 0;
  }
}

// add (launches, -, flops, bytes) per kind: a replayed CUDA graph re-issues launches the LaunchScope only saw at capture time
void f5b_prof_add(const double* delta, int n_kinds) {
  if (delta == nullptr || n_kinds != f5b::K_NUM) return;
  std::lock_guard<std::mutex> lk(f5b::g_prof_mu);
  for (int k = 0; k < f5b::K_NUM; ++k) {
    f5b::g_prof.launches[k] += (unsigned long long)delta[k * 4 + 0];
    f5b::g_prof.flops[k] += delta[k * 4 + 2];
    f5b::g_prof.bytes[k] += delta[k * 4 + 3];
  }
}
int f5b_prof_enabled(void) { return f5b::g_prof.on ? 1 : 0; }

// out[k*4 + {0,1,2,3}] = launches, device milliseconds (sum of event-bracketed launches; 0 if profiling was off),
// algorithmic FLOPs, algorithmic bytes of kernel class k (f5b::KernelKind order).  Synchronises the device.
int f5b_prof_read(double* out, int n_kinds) {
  using namespace f5b;
  F5B_CHECK(out != nullptr && n_kinds == K_NUM, "f5b_prof_read: expected %d kinds", (int)K_NUM);
  F5B_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double ms[K_NUM] = {0};
  for (size_t i = 0; i + 1 < g_prof.used; i += 2) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, g_prof.ev[i], g_prof.ev[i + 1]) == cudaSuccess) ms[g_prof.ev_kind[i / 2]] += t;
  }
  for (int k = 0; k < K_NUM; ++k) {
    out[k * 4 + 0] = (double)g_prof.launches[k];
    out[k * 4 + 1] = ms[k];
    out[k * 4 + 2] = g_prof.flops[k];
    out[k * 4 + 3] = g_prof.bytes[k];
  }
  return 0;
}

}  // extern "C"

extern "C" int f5b_set_dependent_launch(int on) {
  f5b::g_pdl = on ? 1 : 0;
  return 0;
}
