// Host-side helpers shared by the C-ABI entry points: error reporting, TMA tensor-map encoding
// (driver entry point fetched through the runtime, so the library does not link libcuda).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

namespace mv {

void set_error(const char* fmt, ...);
const char* get_error();

#define MV_CHECK_CUDA(expr)                                                                    \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      mv::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return -2;                                                                               \
    }                                                                                          \
  } while (0)

#define MV_REQUIRE(cond, ...)    \
  do {                           \
    if (!(cond)) {               \
      mv::set_error(__VA_ARGS__); \
      return -1;                 \
    }                            \
  } while (0)

// 16-word pinned record a trapping kernel fills in (word 0 == 0x4d564442 when valid)
uint32_t* debug_host_record();
int debug_device_pointer(uint32_t** dptr);

// number of SMs of the current device (cached per device)
int sm_count();

// bf16 tensor maps, 128B swizzle.  Dimensions innermost-first; strides in ELEMENTS for dims >= 1.
// Returns 0 or a negative error (message in get_error()).
int encode_tmap_2d(CUtensorMap* m, const void* base, uint64_t dim0, uint64_t dim1, uint64_t stride1,
                   uint32_t box0, uint32_t box1);
int encode_tmap_3d(CUtensorMap* m, const void* base, uint64_t dim0, uint64_t dim1, uint64_t dim2,
                   uint64_t stride1, uint64_t stride2, uint32_t box0, uint32_t box1, uint32_t box2);

}  // namespace mv
