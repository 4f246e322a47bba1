// nn.Linear with fused epilogue on tcgen05:  C[M,N] = epi(A[M,K] . W[N,K]^T + bias).
//
// Replaces cuBLAS + ATen elementwise at the reference call sites
//   mova/diffusion/models/wan_video_dit.py:171-174 (q,k,v,o), :218-221 (cross-attn), :270-271 (ffn + GELU-tanh),
//   :254-255,:287-290 (gated residuals), mova/diffusion/models/interactionv2.py:218-221,:251,:535.
//
// Design (B200):
//   * persistent kernel, one CTA (CG=1) or one CTA pair (CG=2, `cta_group::2`) per SM / SM pair,
//   * output tile 128*CG x 256, K step 64 (one 128B swizzle atom), 4 (CG=1) / 6 (CG=2) TMA stages,
//   * both operands K-major straight from the row-major activations [M,K] and nn.Linear weights [N,K],
//   * fp32 accumulators in TMEM, double buffered (2 x 256 columns) so the epilogue of tile i overlaps the
//     main loop of tile i+1,
//   * warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..7 = epilogue
//     (TMEM -> registers -> bias / GELU / gated residual -> bf16 -> swizzled smem -> TMA store),
//   * tiles are walked in N-panels sized so that the weight panel (<= 48 MB) stays L2 resident while the activations
//     stream past it once per panel; TMA loads of W carry the L2 evict_last hint, the C stores evict_first (ncu of the
//     round-1 kernel: 8.3-9.3 GB of DRAM traffic per QKV GEMM against 1.9 GB algorithmic -- DRAM joules are clock on a
//     power-capped part).
#include <atomic>
#include "common.cuh"
#include "host_utils.h"
#include "../../include/mova_b200.h"

namespace mv {

constexpr int GEMM_BM = 128;  // rows per CTA
constexpr int GEMM_BN = 256;  // columns per tile
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_STAGING_BYTES = 128 * 64 * 2;

template <int CG>
struct GemmCfg {
  static constexpr int B_ROWS = GEMM_BN / CG;
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = B_ROWS * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (CG == 1) ? 4 : 6;
  static constexpr int OFF_STAGING = STAGES * STAGE_BYTES;
  static constexpr int OFF_BIAS = OFF_STAGING + 2 * GEMM_STAGING_BYTES;
  static constexpr int OFF_GATE = OFF_BIAS + GEMM_BN * 4;
  static constexpr int OFF_BARS = OFF_GATE + GEMM_BN * 4;
  static constexpr int NUM_BARS = 2 * STAGES + 4;
  static constexpr int OFF_TMEM_PTR = OFF_BARS + NUM_BARS * 8;
  static constexpr int SMEM_BYTES = OFF_TMEM_PTR + 16;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");
};

struct GemmParams {
  int M, N, K;
  int seg_k;  // A is split along K into segments of seg_k columns living a_seg_stride apart (3rd TMA dimension)
  int seg_n;  // C is split along N into segments of seg_n columns living c_seg_stride apart (3rd TMA dimension)
  const __nv_bfloat16* bias;      // [N] or null
  const __nv_bfloat16* residual;  // [M, ldr] or null
  long long ldr;
  const float* gate;  // [N] or null
  float scale;
  int panel;  // n-tiles per L2 panel
};

__device__ __forceinline__ void gemm_decode_tile(int tile, int m_tiles, int n_tiles, int panel_w, int& m_blk,
                                                 int& n_blk) {
  const int per_panel = m_tiles * panel_w;
  const int panel = tile / per_panel;
  const int within = tile - panel * per_panel;
  const int n_begin = panel * panel_w;
  const int pw = min(panel_w, n_tiles - n_begin);
  m_blk = within / pw;
  n_blk = n_begin + (within - m_blk * pw);
}

__device__ __forceinline__ float gelu_tanh_f(float x) {
  // 0.5 x (1 + tanh(u)),  u = sqrt(2/pi) (x + 0.044715 x^3)   ==   x * sigmoid(2u)
  const float u = 0.7978845608028654f * x * (1.0f + 0.044715f * x * x);
  const float e = fast_exp2(-2.0f * 1.4426950408889634f * u);
  return x * fast_rcp(1.0f + e);
}

template <int CG, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  using Cfg = GemmCfg<CG>;
  extern __shared__ __align__(1024) uint8_t smem[];

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool is_leader = (rank == 0);

  const uint32_t smem_base = smem_u32(smem);
  float* sBias = reinterpret_cast<float*>(smem + Cfg::OFF_BIAS);
  float* sGate = reinterpret_cast<float*>(smem + Cfg::OFF_GATE);
  const uint32_t bars = smem_base + Cfg::OFF_BARS;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * Cfg::STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * Cfg::STAGES + 2 + a); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + Cfg::OFF_TMEM_PTR);

  if (threadIdx.x == 0) {
    if ((smem_base & 1023u) != 0) __trap();  // swizzle-128B tiles need a 1024-byte aligned window
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
  }
  if (threadIdx.x == 32) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      // one arrival: the leader's expect_tx covering BOTH CTAs' bytes.  The peer only issues its TMA loads against the
      // leader's barrier; their complete_tx may land before the expect_tx (the tx-count is signed within a phase) but
      // never in an earlier phase: the peer waits for its multicast `empty` first, i.e. after the leader's MMAs consumed
      // the stage, so the leader's barrier is already in the next phase with its own arrival still pending.
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);  // one tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4 * CG);  // one arrive per epilogue warp of every CTA in the pair
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) tmem_alloc<CG>(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), 512);

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int m_tiles = (p.M + GEMM_BM * CG - 1) / (GEMM_BM * CG);
  const int n_tiles = (p.N + GEMM_BN - 1) / GEMM_BN;
  const int total_tiles = m_tiles * n_tiles;
  const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int cluster_id = blockIdx.x / CG;
  const int num_clusters = gridDim.x / CG;

  if (warp_idx == 0) {
    // ============================== TMA producer ==============================
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        int m_blk, n_blk;
        gemm_decode_tile(tile, m_tiles, n_tiles, p.panel, m_blk, n_blk);
        const int row0 = m_blk * GEMM_BM * CG + static_cast<int>(rank) * GEMM_BM;
        const int col0 = n_blk * GEMM_BN + static_cast<int>(rank) * Cfg::B_ROWS;
        int k_in_seg = 0, seg = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sA = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sB = sA + Cfg::A_BYTES;
          if (k_in_seg >= p.seg_k) { k_in_seg = 0; ++seg; }
          if constexpr (CG == 1) {
            mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
            tma_load_3d(sA, &tmA, full_bar(stage), k_in_seg, row0, seg);
            tma_load_2d_hint(sB, &tmB, full_bar(stage), kb * GEMM_BK, col0, L2_EVICT_LAST);
          } else {
            const uint32_t leader_full = mapa_shared(full_bar(stage), 0);
            if (is_leader) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
            tma_load_3d_pair(sA, &tmA, leader_full, k_in_seg, row0, seg);
            tma_load_2d_pair_hint(sB, &tmB, leader_full, kb * GEMM_BK, col0, L2_EVICT_LAST);
          }
          k_in_seg += GEMM_BK;
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ============================== MMA issuer (leader CTA only) ==============================
    if (is_leader && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM * CG, GEMM_BN, 0, 0);
      uint32_t stage = 0, phase = 0;
      uint32_t tile_iter = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++tile_iter) {
        const uint32_t acc = tile_iter & 1;
        const uint32_t acc_phase = (tile_iter >> 1) & 1;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * GEMM_BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sA = smem_base + stage * Cfg::STAGE_BYTES;
          const uint64_t a_desc = umma_desc_k_sw128(sA);
          const uint64_t b_desc = umma_desc_k_sw128(sA + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            umma_ss<CG>(d_tmem, umma_desc_advance(a_desc, k * 32), umma_desc_advance(b_desc, k * 32), idesc,
                        (kb > 0 || k > 0) ? 1u : 0u);
          }
          if constexpr (CG == 1) umma_commit(empty_bar(stage)); else umma_commit_pair(empty_bar(stage), 0x3);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        if constexpr (CG == 1) umma_commit(tfull_bar(acc)); else umma_commit_pair(tfull_bar(acc), 0x3);
      }
    }
  } else if (warp_idx >= 4) {
    // ============================== epilogue ==============================
    const int ew = warp_idx - 4;          // == warp_idx % 4 : TMEM lane quarter
    const int et = threadIdx.x - 128;     // 0..127
    const int row_in_tile = ew * 32 + lane;
    const uint32_t staging = smem_base + Cfg::OFF_STAGING;
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
    uint32_t tile_iter = 0;
    uint32_t chunk_counter = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++tile_iter) {
      int m_blk, n_blk;
      gemm_decode_tile(tile, m_tiles, n_tiles, p.panel, m_blk, n_blk);
      const int row0 = m_blk * GEMM_BM * CG + static_cast<int>(rank) * GEMM_BM;
      const int n0 = n_blk * GEMM_BN;
      const uint32_t acc = tile_iter & 1;
      const uint32_t acc_phase = (tile_iter >> 1) & 1;

      for (int i = et; i < GEMM_BN; i += 128) {
        const int n = n0 + i;
        sBias[i] = (p.bias != nullptr && n < p.N) ? __bfloat162float(p.bias[n]) : 0.0f;
        if constexpr (EPI == MOVA_EPI_RESIDUAL)
          sGate[i] = ((p.gate != nullptr && n < p.N) ? p.gate[n] : 1.0f) * p.scale;
      }

      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();

      const int grow = row0 + row_in_tile;
#pragma unroll 1
      for (int c = 0; c < GEMM_BN / 64; ++c, ++chunk_counter) {
        const uint32_t buf = staging + (chunk_counter & 1) * GEMM_STAGING_BYTES;
        const int ncol0 = n0 + c * 64;
        if (ncol0 >= p.N) continue;  // uniform across the CTA: nothing to store for this chunk

        // the TMA store that last read this staging buffer (2 chunks ago) must have drained
        if (et == 0) tma_store_wait_read<1>();
        named_bar_sync(1, 128);

#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint4 res[4];
          if constexpr (EPI == MOVA_EPI_RESIDUAL) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int ncol = ncol0 + h * 32 + j * 8;
              if (grow < p.M && ncol < p.N)
                res[j] = *reinterpret_cast<const uint4*>(p.residual + static_cast<long long>(grow) * p.ldr + ncol);
              else
                res[j] = make_uint4(0, 0, 0, 0);
            }
          }
          uint32_t v[32];
          tmem_ld_x32(lane_taddr + acc * GEMM_BN + c * 64 + h * 32, v);
          tmem_wait_ld();
          const float* bias_s = sBias + c * 64 + h * 32;
          const float* gate_s = sGate + c * 64 + h * 32;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t packed[4];
            const uint32_t rr[4] = {res[j].x, res[j].y, res[j].z, res[j].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int idx = j * 8 + e * 2;
              float x0 = __uint_as_float(v[idx]) + bias_s[idx];
              float x1 = __uint_as_float(v[idx + 1]) + bias_s[idx + 1];
              if constexpr (EPI == MOVA_EPI_GELU_TANH) {
                x0 = gelu_tanh_f(x0);
                x1 = gelu_tanh_f(x1);
              } else if constexpr (EPI == MOVA_EPI_RESIDUAL) {
                x0 = fmaf(gate_s[idx], x0, bf16lo(rr[e]));
                x1 = fmaf(gate_s[idx + 1], x1, bf16hi(rr[e]));
              }
              packed[e] = pack_bf16x2(x0, x1);
            }
            const int chunk16 = h * 4 + j;  // 16-byte chunk index inside the 128-byte staging row
            const uint32_t dst = buf + row_in_tile * 128 + ((chunk16 ^ (row_in_tile & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(packed[0]), "r"(packed[1]),
                         "r"(packed[2]), "r"(packed[3])
                         : "memory");
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (et == 0) {
          const int cseg = ncol0 / p.seg_n;
          tma_store_3d_hint(&tmC, buf, ncol0 - cseg * p.seg_n, row0, cseg, L2_EVICT_FIRST);
          tma_store_commit();
        }
      }
      // accumulator stage drained: hand it back to the MMA issuer (lives in the leader CTA)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 1) mbar_arrive(tempty_bar(acc));
        else mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));
      }
    }
    if (et == 0) tma_store_wait<0>();
  }

  // ============================== teardown ==============================
  __syncwarp();
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp_idx == 2) tmem_dealloc<CG>(tmem_base, 512);
}

template <int CG, int EPI>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmParams& p,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<CG>;
  auto kernel = gemm_bf16_kernel<CG, EPI>;
  debug_attach();
  static std::atomic<bool> configured[64];  // zero-initialised; per-device "attribute set" latch, safe across host threads
  int dev = 0;
  MV_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
    MV_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
  }
  const int m_tiles = (p.M + GEMM_BM * CG - 1) / (GEMM_BM * CG);
  const int n_tiles = (p.N + GEMM_BN - 1) / GEMM_BN;
  const int total = m_tiles * n_tiles;
  int clusters = sm_count() / CG;
  if (clusters > total) clusters = total;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CG, 1, 1);
  cfg.blockDim = dim3(GEMM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = CG;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  MV_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, tmC, p));
  return 0;
}

}  // namespace mv

extern "C" int mova_b200_linear(const void* A, int64_t lda, const void* W, int64_t ldw, const void* bias, void* C,
                                int64_t ldc, int M, int N, int K, int epilogue, const void* residual, int64_t ldr,
                                const float* gate, float scale, int cta_group, void* stream) {
  return mova_b200_linear_ex(A, lda, K, 0, W, ldw, bias, C, ldc, N, 0, M, N, K, epilogue, residual, ldr, gate, scale,
                             cta_group, stream);
}

extern "C" int mova_b200_linear_ex(const void* A, int64_t lda, int seg_k, int64_t seg_stride, const void* W,
                                   int64_t ldw, const void* bias, void* C, int64_t ldc, int seg_n,
                                   int64_t c_seg_stride, int M, int N, int K, int epilogue, const void* residual,
                                   int64_t ldr, const float* gate, float scale, int cta_group, void* stream) {
  using namespace mv;
  MV_REQUIRE(A && W && C, "mova_b200_linear: null operand pointer");
  MV_REQUIRE(M >= 0 && N > 0 && K > 0, "mova_b200_linear: bad shape M=%d N=%d K=%d", M, N, K);
  MV_REQUIRE(N % 8 == 0 && K % 8 == 0, "mova_b200_linear: N (%d) and K (%d) must be multiples of 8", N, K);
  MV_REQUIRE(seg_k > 0 && seg_k <= K && K % seg_k == 0, "mova_b200_linear: seg_k (%d) must divide K (%d)", seg_k, K);
  MV_REQUIRE(seg_k == K || seg_k % GEMM_BK == 0, "mova_b200_linear: a segmented A needs seg_k %% %d == 0", GEMM_BK);
  MV_REQUIRE(seg_n > 0 && seg_n <= N && N % seg_n == 0, "mova_b200_linear: seg_n (%d) must divide N (%d)", seg_n, N);
  MV_REQUIRE(seg_n == N || seg_n % 64 == 0, "mova_b200_linear: a segmented C needs seg_n %% 64 == 0");
  MV_REQUIRE(seg_n == N || epilogue != MOVA_EPI_RESIDUAL, "mova_b200_linear: residual epilogue with a segmented C");
  MV_REQUIRE(lda >= seg_k && ldw >= K && ldc >= seg_n, "mova_b200_linear: leading dimension smaller than row length");
  const int ncseg = N / seg_n;
  if (ncseg == 1) c_seg_stride = ldc * static_cast<int64_t>(M > 0 ? M : 1);
  MV_REQUIRE(c_seg_stride % 8 == 0 && c_seg_stride > 0, "mova_b200_linear: c_seg_stride must be a positive multiple of 8");
  const int nseg = K / seg_k;
  if (nseg == 1) seg_stride = lda * static_cast<int64_t>(M > 0 ? M : 1);
  MV_REQUIRE(seg_stride % 8 == 0 && seg_stride > 0, "mova_b200_linear: seg_stride must be a positive multiple of 8");
  MV_REQUIRE(epilogue >= MOVA_EPI_BIAS && epilogue <= MOVA_EPI_RESIDUAL, "mova_b200_linear: unknown epilogue %d",
             epilogue);
  if (epilogue == MOVA_EPI_RESIDUAL) {
    MV_REQUIRE(residual != nullptr && ldr >= N && ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0,
               "mova_b200_linear: residual epilogue needs a 16B-aligned residual with ldr >= N, ldr %% 8 == 0");
  }
  if (M == 0) return 0;
  if (cta_group == 0) {
    // Chosen per shape.  Measured on B200 (profiles/r02_gemm_vs_cublas.jsonl): the 256 x 256 CTA-pair tile runs at
    // 1.42-1.51 PFLOP/s against 1.31-1.32 for the 128 x 256 single-CTA tile on the 360p video shapes (each SM reads half
    // the B operand from shared memory and the pair loads 2/3 of the bytes per FLOP from L2 -- on a power-capped part
    // that is clock), and is >= cuBLASLt on 8 of the 9 shapes timed.  Problems that do not fill the GPU with
    // single-CTA tiles (the 403-token audio tower, the text keys) keep the smaller tile for parallelism.
    const long long tiles1 = static_cast<long long>((M + GEMM_BM - 1) / GEMM_BM) * ((N + GEMM_BN - 1) / GEMM_BN);
    cta_group = (tiles1 >= sm_count()) ? 2 : 1;
  }
  MV_REQUIRE(cta_group == 1 || cta_group == 2, "mova_b200_linear: cta_group must be 0, 1 or 2");

  CUtensorMap tmA, tmB, tmC;
  int rc;
  if ((rc = encode_tmap_3d(&tmA, A, seg_k, M, nseg, lda, seg_stride, GEMM_BK, GEMM_BM, 1)) != 0) return rc;
  if ((rc = encode_tmap_2d(&tmB, W, K, N, ldw, GEMM_BK, GEMM_BN / cta_group)) != 0) return rc;
  if ((rc = encode_tmap_3d(&tmC, C, seg_n, M, ncseg, ldc, c_seg_stride, 64, GEMM_BM, 1)) != 0) return rc;

  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.seg_k = seg_k;
  p.seg_n = seg_n;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.residual = static_cast<const __nv_bfloat16*>(residual);
  p.ldr = ldr;
  p.gate = gate;
  p.scale = scale;
  // n-tiles per panel: the panel's rows of W (panel x 256 x K bf16) should sit in L2 (126 MB, shared with the streaming
  // operands) while all M-tiles pass: <= 48 MB -> 16 tiles at K = 5120, 6 at K = 13824
  {
    const long long per_tile = 2LL * GEMM_BN * K;
    long long pw = (48LL << 20) / per_tile;
    p.panel = static_cast<int>(pw < 2 ? 2 : (pw > 16 ? 16 : pw));
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);

#define MV_GEMM_DISPATCH(CGV)                                                                           \
  switch (epilogue) {                                                                                   \
    case MOVA_EPI_BIAS: return launch_gemm<CGV, MOVA_EPI_BIAS>(tmA, tmB, tmC, p, s);                    \
    case MOVA_EPI_GELU_TANH: return launch_gemm<CGV, MOVA_EPI_GELU_TANH>(tmA, tmB, tmC, p, s);          \
    default: return launch_gemm<CGV, MOVA_EPI_RESIDUAL>(tmA, tmB, tmC, p, s);                           \
  }
  if (cta_group == 1) { MV_GEMM_DISPATCH(1) } else { MV_GEMM_DISPATCH(2) }
#undef MV_GEMM_DISPATCH
}
