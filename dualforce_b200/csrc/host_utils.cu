#include "host_utils.h"

#include <string.h>

#include "../../include/mova_b200.h"

namespace mv {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// Pinned, device-mapped diagnostics record (see common.cuh: mbar_timeout_trap).
static uint32_t* g_dbg_host = nullptr;
uint32_t* debug_host_record() {
  if (g_dbg_host == nullptr) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, 64, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) return nullptr;
    memset(p, 0, 64);
    g_dbg_host = static_cast<uint32_t*>(p);
  }
  return g_dbg_host;
}
int debug_device_pointer(uint32_t** dptr) {
  uint32_t* h = debug_host_record();
  if (h == nullptr) return -1;
  void* d = nullptr;
  if (cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) return -1;
  *dptr = static_cast<uint32_t*>(d);
  return 0;
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
      set_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
      return nullptr;
    }
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

static int encode_nd(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                     const uint32_t* box) {
  auto fn = get_encode_fn();
  if (fn == nullptr) return -3;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_error("tensor map: base pointer %p is not 16-byte aligned", base);
    return -1;
  }
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i >= 1) {
      gstrides[i - 1] = strides_elems[i - 1] * 2;  // bytes
      if (gstrides[i - 1] % 16 != 0) {
        set_error("tensor map: stride %llu bytes of dim %d is not a multiple of 16",
                  (unsigned long long)gstrides[i - 1], i);
        return -1;
      }
    }
  }
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdims, gstrides, gbox, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu box %u,%u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return -3;
  }
  return 0;
}

int encode_tmap_2d(CUtensorMap* m, const void* base, uint64_t dim0, uint64_t dim1, uint64_t stride1, uint32_t box0,
                   uint32_t box1) {
  uint64_t dims[2] = {dim0, dim1};
  uint64_t strides[1] = {stride1};
  uint32_t box[2] = {box0, box1};
  return encode_nd(m, base, 2, dims, strides, box);
}

int encode_tmap_3d(CUtensorMap* m, const void* base, uint64_t dim0, uint64_t dim1, uint64_t dim2, uint64_t stride1,
                   uint64_t stride2, uint32_t box0, uint32_t box1, uint32_t box2) {
  uint64_t dims[3] = {dim0, dim1, dim2};
  uint64_t strides[2] = {stride1, stride2};
  uint32_t box[3] = {box0, box1, box2};
  return encode_nd(m, base, 3, dims, strides, box);
}

}  // namespace mv

extern "C" {

int mova_b200_abi_version(void) { return MOVA_B200_ABI_VERSION; }

const char* mova_b200_last_error(void) { return mv::get_error(); }

const uint32_t* mova_b200_debug_record(void) { return mv::debug_host_record(); }

int mova_b200_device_check(int device) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    mv::set_error("cudaGetDeviceProperties(%d) failed: %s", device, cudaGetErrorString(e));
    return -2;
  }
  if (prop.major != 10) {
    mv::set_error("device %d is sm_%d%d; mova_b200 kernels are built for sm_100a only (no fallback)", device,
                  prop.major, prop.minor);
    return -1;
  }
  return 0;
}

}  // extern "C"
