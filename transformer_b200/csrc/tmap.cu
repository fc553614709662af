// Host utilities: thread-local error string, TMA descriptor encoding through the driver entry point
// (no link-time dependency on libcuda, so the library loads on a CPU-only box for the symbol check).
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace bx {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

static std::atomic<long long> g_launches{0};
void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

// ---- optional per-launch CUDA-event timing of the tensor-core kernels (bench.py's live roofline numbers) ----
struct ProfRec { cudaEvent_t e0, e1; int cat; double work; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_event_pool;
static std::mutex g_prof_mu;
static int g_gemm_cat = 0;
void prof_set_gemm_cat(int cat) { g_gemm_cat = cat; }
int prof_gemm_cat() { return g_gemm_cat; }
bool prof_on() { return g_prof_on; }
void prof_enable(int on) { g_prof_on = on != 0; }
static cudaEvent_t prof_event() {
  if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
void* prof_begin(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEvent_t e = prof_event();
  cudaEventRecord(e, st);
  return e;
}
void prof_end(void* begin, int cat, double work, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEvent_t e1 = prof_event();
  cudaEventRecord(e1, st);
  g_prof.push_back({reinterpret_cast<cudaEvent_t>(begin), e1, cat, work});
}
// out[cat*3 + {0,1,2}] = {total ms, total work (flops), launches}; call after the stream is synchronised.
int prof_collect(double* out, int ncat) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < ncat * 3; ++i) out[i] = 0.0;
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess && r.cat >= 0 && r.cat < ncat) {
      out[r.cat * 3 + 0] += ms; out[r.cat * 3 + 1] += r.work; out[r.cat * 3 + 2] += 1.0;
    }
    g_event_pool.push_back(r.e0); g_event_pool.push_back(r.e1);
  }
  g_prof.clear();
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static CUtensorMapSwizzle swz(int bytes) {
  switch (bytes) {
    case 32: return CU_TENSOR_MAP_SWIZZLE_32B;
    case 64: return CU_TENSOR_MAP_SWIZZLE_64B;
    case 128: return CU_TENSOR_MAP_SWIZZLE_128B;
    default: return CU_TENSOR_MAP_SWIZZLE_NONE;
  }
}

int encode_tmap_2d(CUtensorMap* tm, int dtype, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                   uint32_t box_inner, uint32_t box_outer, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode();
  BX_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  const uint64_t es = dtype == 0 ? 2 : 4;
  BX_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base %p not 16-byte aligned", base);
  BX_REQUIRE((ld_elems * es) % 16 == 0, "TMA row pitch %llu B not a multiple of 16", (unsigned long long)(ld_elems * es));
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * es};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz(swizzle_bytes),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BX_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d) failed: %d (inner %llu outer %llu ld %llu box %u x %u)",
             (int)r, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld_elems, box_inner,
             box_outer);
  return 0;
}

int encode_tmap_3d(CUtensorMap* tm, int dtype, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                   uint64_t stride1_elems, uint64_t stride2_elems, uint32_t b0, uint32_t b1, uint32_t b2,
                   int swizzle_bytes) {
  EncodeTiledFn fn = get_encode();
  BX_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  const uint64_t es = dtype == 0 ? 2 : 4;
  BX_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base %p not 16-byte aligned", base);
  BX_REQUIRE((stride1_elems * es) % 16 == 0 && (stride2_elems * es) % 16 == 0, "TMA strides not multiples of 16 B");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_elems * es, stride2_elems * es};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz(swizzle_bytes),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BX_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed: %d", (int)r);
  return 0;
}

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  return dev < kMaxDevices ? dev : kMaxDevices - 1;
}

int sm_count() {
  static std::atomic<int> n[kMaxDevices];
  const int dev = current_device();
  int v = n[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    n[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

}  // namespace bx
