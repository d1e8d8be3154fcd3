// C-ABI plumbing shared by all kernels: thread-local error string, launch check,
// driver entry point for cuTensorMapEncodeTiled (resolved at run time so the
// library loads -- and exports its symbols -- on a machine without libcuda).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "geo4d_b200.h"

namespace g4 {

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    (void)cudaGetLastError();
    return G4_ERR_CUDA;
  }
  ++g_launches;
  return G4_OK;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GEO4D_PDL");
    v = (e && e[0] == '1') ? 1 : 0;   // off by default: no gain under CUDA-graph replay (profiles/)
  }
  return v == 1;
}

int device_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    set_last_error("cudaGetDevice failed");
    return -1;
  }
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
      set_last_error("cannot query SM count");
      return -1;
    }
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(p);
    else (void)cudaGetLastError();
  }
  return fn;
}

bool once_per_device(unsigned long long* mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return true; }
  const unsigned long long bit = 1ull << (dev & 63);
  const unsigned long long old = __atomic_fetch_or(mask, bit, __ATOMIC_ACQ_REL);
  return (old & bit) == 0;
}
void unlatch_device(unsigned long long* mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return; }
  __atomic_fetch_and(mask, ~(1ull << (dev & 63)), __ATOMIC_ACQ_REL);
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_last_error("cuTensorMapEncodeTiled driver entry point unavailable (no CUDA driver?)");
    return G4_ERR_DRIVER;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr,
                   bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u] "
                   "stride0 %llu",
                   (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                   (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                   box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0,
                   (unsigned long long)(rank > 1 ? strides_bytes[0] : 0));
    return G4_ERR_DRIVER;
  }
  return G4_OK;
}

}  // namespace g4

extern "C" int geo4d_abi_version(void) { return GEO4D_ABI_VERSION; }
extern "C" uint64_t geo4d_launch_count(void) { return g4::g_launches; }
extern "C" const char* geo4d_last_error(void) { return g4::g_err; }
extern "C" int geo4d_device_supported(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return major == 10 ? 1 : 0;
}
