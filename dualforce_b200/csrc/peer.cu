// Ulysses head <-> sequence exchange over NVSwitch peer memory -- the data path of the context-parallel all-to-all
// without NCCL kernels.
//
// Replaces the two all-to-alls yunchang's LongContextAttention issues inside USPAttention.forward
// (mova/diffusion/models/wan_video_dit.py:192-208) around every video self-attention.
//
// Why not NCCL here.  The exchange of head group g+1 is meant to travel while head group g is in the tensor cores.
// An NCCL all-to-all is a kernel: it needs SMs, and the attention kernel holds every SM with one 213 KB CTA, so the
// exchange only advances when attention CTAs retire (measured at cp = 8, profiles/r02_timeline_cp8.json: 0.10-0.16 ms
// per head group alone, 0.75-2.1 ms beside attention -- the exchange, not the attention, sets the layer time).
// Here every rank owns one receive window (cudaMalloc, exported with cudaIpcGetMemHandle, mapped by its peers); the
// per-(group, peer) chunks the GEMM already wrote destination-rank-major are contiguous, so the exchange is
//     n x cudaMemcpyAsync(peer window + offset, chunk)      -- copy engines over NVLink, no SM, no kernel
//     the epoch into a flag word in each destination's window
// and the consumer's stream waits for its own flag words before the attention / o-projection launch.
//
// Two implementations of the flag side, selected per call:
//   * stream memory operations (default): cuStreamWriteValue64 puts the epoch into a word of this rank's window, n
//     8-byte copy-engine copies carry it to the peers' flag words (stream order keeps them behind the data), and the
//     consumer waits with cuStreamBatchMemOp(WAIT_VALUE_64 >=).  No kernel anywhere: measured on two B200s with the
//     kernel version below, the 32-thread signal kernel queued behind the attention kernel (which owns every SM's
//     registers and shared memory) and reached the peer 7 ms late (profiles/r02_timeline_cp2_peer_kernel_flags.json);
//   * kernels (fallback / diagnostics): one 32-thread kernel stores the flags (st.release.sys), one polling kernel waits
//     (ld.acquire.sys) and traps after `timeout_ms`, so a lost peer is an error on this rank instead of a silent wait.
#include <stdint.h>
#include <string.h>

#include "../../include/mova_b200.h"
#include "host_utils.h"

namespace mv {

constexpr int PEER_MAX_FLAGS = 32;

struct PeerFlagList {
  unsigned long long* p[PEER_MAX_FLAGS];
};

__global__ void peer_signal_kernel(PeerFlagList flags, int n, unsigned long long epoch) {
  const int i = threadIdx.x;
  if (i < n) {
    // everything this stream queued before the kernel (the copies into the same peer) is complete; the fence orders it
    // before the flag for any observer in the system
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flags.p[i]), "l"(epoch) : "memory");
  }
}

__global__ void peer_wait_kernel(const unsigned long long* flags, int n, unsigned long long epoch,
                                 unsigned long long timeout_ns, uint32_t* dbg) {
  const int i = threadIdx.x;
  if (i < n) {
    unsigned long long t0 = 0, now = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    unsigned long long seen = 0;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flags + i) : "memory");
      if (seen >= epoch) break;
      __nanosleep(200);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (now - t0 > timeout_ns) {
        if (dbg != nullptr) {
          dbg[1] = static_cast<uint32_t>(i);
          dbg[2] = static_cast<uint32_t>(epoch);
          dbg[3] = static_cast<uint32_t>(seen);
          dbg[0] = 0x4d565057u;  // "MVPW": peer wait timed out
          __threadfence_system();
        }
        __trap();
      }
    }
  }
  __syncthreads();
  __threadfence_system();
}

struct MemOps {
  PFN_cuStreamWriteValue64_v11070 write = nullptr;
  PFN_cuStreamBatchMemOp_v11070 batch = nullptr;
  bool tried = false;
};

static MemOps* memops() {
  static MemOps m;
  if (!m.tried) {
    m.tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue64", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      m.write = reinterpret_cast<PFN_cuStreamWriteValue64_v11070>(p);
    p = nullptr;
    if (cudaGetDriverEntryPoint("cuStreamBatchMemOp", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      m.batch = reinterpret_cast<PFN_cuStreamBatchMemOp_v11070>(p);
  }
  return (m.write != nullptr && m.batch != nullptr) ? &m : nullptr;
}

static bool memops_supported() {
  if (memops() == nullptr) return false;
  int dev = 0, can64 = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  // cudaDevAttrReserved122 == CU_DEVICE_ATTRIBUTE_CAN_USE_64_BIT_STREAM_MEM_OPS
  if (cudaDeviceGetAttribute(&can64, static_cast<cudaDeviceAttr>(122), dev) != cudaSuccess) return false;
  return can64 != 0;
}

}  // namespace mv

extern "C" {

int mova_b200_peer_memops_supported(void) { return mv::memops_supported() ? 1 : 0; }

int mova_b200_peer_alloc(int64_t nbytes, void** ptr, void* handle64) {
  MV_REQUIRE(nbytes > 0 && ptr != nullptr && handle64 != nullptr, "peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
  void* p = nullptr;
  MV_CHECK_CUDA(cudaMalloc(&p, static_cast<size_t>(nbytes)));
  cudaError_t e = cudaMemset(p, 0, static_cast<size_t>(nbytes));
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    mv::set_error("peer_alloc: %s", cudaGetErrorString(e));
    cudaFree(p);
    return -2;
  }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return 0;
}

int mova_b200_peer_open(const void* handle64, void** ptr) {
  MV_REQUIRE(ptr != nullptr && handle64 != nullptr, "peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  MV_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr = p;
  return 0;
}

int mova_b200_peer_close(void* ptr) {
  MV_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}

int mova_b200_peer_free(void* ptr) {
  MV_CHECK_CUDA(cudaFree(ptr));
  return 0;
}

int mova_b200_peer_push(int n_copies, void* const* dst, const void* const* src, const int64_t* nbytes, int n_flags,
                        void* const* flags, int64_t local_mask, int64_t epoch, void* epoch_src, void* stream) {
  MV_REQUIRE(n_copies >= 0 && n_flags >= 0 && n_flags <= mv::PEER_MAX_FLAGS, "peer_push: %d copies / %d flags (max %d)",
             n_copies, n_flags, mv::PEER_MAX_FLAGS);
  MV_REQUIRE(epoch > 0, "peer_push: epoch must be positive");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (int i = 0; i < n_copies; ++i) {
    if (nbytes[i] <= 0) continue;
    MV_CHECK_CUDA(cudaMemcpyAsync(dst[i], src[i], static_cast<size_t>(nbytes[i]), cudaMemcpyDefault, s));
  }
  if (n_flags == 0) return 0;
  for (int i = 0; i < n_flags; ++i)
    MV_REQUIRE((reinterpret_cast<uintptr_t>(flags[i]) & 7) == 0, "peer_push: flag %d is not 8-byte aligned", i);
  if (epoch_src != nullptr) {
    // no kernel: the epoch goes into a word of this rank's window, copy-engine copies carry it to the flag words
    mv::MemOps* m = mv::memops();
    MV_REQUIRE(m != nullptr, "peer_push: stream memory operations are not available in this driver");
    MV_REQUIRE((reinterpret_cast<uintptr_t>(epoch_src) & 7) == 0, "peer_push: epoch_src is not 8-byte aligned");
    CUresult r = m->write(reinterpret_cast<CUstream>(s), reinterpret_cast<CUdeviceptr>(epoch_src),
                          static_cast<cuuint64_t>(epoch), CU_STREAM_WRITE_VALUE_DEFAULT);
    MV_REQUIRE(r == CUDA_SUCCESS, "peer_push: cuStreamWriteValue64 failed (CUresult %d)", static_cast<int>(r));
    for (int i = 0; i < n_flags; ++i) {
      if ((local_mask >> i) & 1) {
        // a word of this device's own memory: written by the stream front end (a same-device cudaMemcpyAsync runs on
        // the SMs and would queue behind a kernel that owns them -- profiles/r02_ce_overlap_probe.json)
        r = m->write(reinterpret_cast<CUstream>(s), reinterpret_cast<CUdeviceptr>(flags[i]),
                     static_cast<cuuint64_t>(epoch), CU_STREAM_WRITE_VALUE_DEFAULT);
        MV_REQUIRE(r == CUDA_SUCCESS, "peer_push: cuStreamWriteValue64 failed (CUresult %d)", static_cast<int>(r));
      } else {
        MV_CHECK_CUDA(cudaMemcpyAsync(flags[i], epoch_src, 8, cudaMemcpyDefault, s));
      }
    }
    return 0;
  }
  mv::PeerFlagList fl;
  for (int i = 0; i < n_flags; ++i) fl.p[i] = static_cast<unsigned long long*>(flags[i]);
  mv::peer_signal_kernel<<<1, 32, 0, s>>>(fl, n_flags, static_cast<unsigned long long>(epoch));
  MV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mova_b200_peer_wait(const void* flags, int n_flags, int64_t epoch, int timeout_ms, int use_memops, void* stream) {
  MV_REQUIRE(flags != nullptr && n_flags > 0 && n_flags <= 1024, "peer_wait: bad flag range (%d)", n_flags);
  MV_REQUIRE((reinterpret_cast<uintptr_t>(flags) & 7) == 0, "peer_wait: flags are not 8-byte aligned");
  MV_REQUIRE(epoch > 0 && timeout_ms > 0, "peer_wait: epoch and timeout must be positive");
  if (use_memops) {
    mv::MemOps* m = mv::memops();
    MV_REQUIRE(m != nullptr, "peer_wait: stream memory operations are not available in this driver");
    CUstreamBatchMemOpParams ops[64];
    int done = 0;
    while (done < n_flags) {
      const int n = (n_flags - done) < 64 ? (n_flags - done) : 64;
      for (int i = 0; i < n; ++i) {
        memset(&ops[i], 0, sizeof(ops[i]));
        ops[i].waitValue.operation = CU_STREAM_MEM_OP_WAIT_VALUE_64;
        ops[i].waitValue.address =
            reinterpret_cast<CUdeviceptr>(static_cast<const unsigned long long*>(flags) + done + i);
        ops[i].waitValue.value64 = static_cast<cuuint64_t>(epoch);
        ops[i].waitValue.flags = CU_STREAM_WAIT_VALUE_GEQ;
      }
      CUresult r = m->batch(reinterpret_cast<CUstream>(stream), static_cast<unsigned>(n), ops, 0);
      MV_REQUIRE(r == CUDA_SUCCESS, "peer_wait: cuStreamBatchMemOp failed (CUresult %d)", static_cast<int>(r));
      done += n;
    }
    return 0;
  }
  uint32_t* dbg = nullptr;
  mv::debug_device_pointer(&dbg);
  const int threads = ((n_flags + 31) / 32) * 32;
  mv::peer_wait_kernel<<<1, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const unsigned long long*>(flags), n_flags, static_cast<unsigned long long>(epoch),
      static_cast<unsigned long long>(timeout_ms) * 1000000ull, dbg);
  MV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
