// Round-1 schedule of the attention forward (two 128-row query tiles per CTA, S/P aliased per tile), kept as variant
// 3 of mova_b200_attn_fwd_variant for A/B measurements against the round-2 kernel in attn_pair.cu, which is what
// mova_b200_attn_fwd runs.  Also home of the C-ABI entry points of both.
//
// Non-causal softmax attention forward on tcgen05, head_dim 128:  O = softmax(Q K^T * scale) V.
//
// Replaces flash_attention() (mova/diffusion/models/wan_video_dit.py:58-91) at its three call sites:
// video / audio self-attention (:188), text cross-attention (:241) and the a2v / v2a bridge
// (mova/diffusion/models/interactionv2.py:250).  q/k/v/o stay in the reference's flat [B, S, H*D] layout
// (head h = columns [128h, 128h+128)), addressed through 3-D TMA tensor maps -- no rearrange copies.
//
// Design (B200), one CTA = 256 query rows (two 128-row tiles) x one head:
//   * warps 0-3 / 4-7: softmax warpgroup of tile 0 / tile 1 (thread == one query row, TMEM lane == row),
//     warp 8: TMA producer, warp 9: tcgen05 issuer + TMEM owner, warps 10-11 idle (register donors).
//   * TMEM (512 columns): S0 | S1 | O0 | O1, fp32.  P (bf16) is written back over the first 64 columns of
//     its S tile and consumed straight from TMEM as the A operand of the P.V MMA.
//   * K and V tiles (128 keys x 128 dims, 32 KB) stream through one 5-slot ring in the order
//     K0 V0 K1 V1 ...; V is used as an MN-major B operand, so no transpose is ever made.
//   * the issuer warp walks the schedule warp-uniformly (descriptors in uniform registers, one elected lane
//     issues):  PV0(j) QK0(j+1) PV1(j) QK1(j+1), so one tile's softmax overlaps the other tile's MMAs; tcgen05
//     executes in issue order, which is what makes the S/P aliasing safe.
//   * P is released to the tensor core in two halves (keys 0-63, 64-127; one mbarrier each): the first half of
//     P.V runs while the second half of the row is still in the exponential unit.
//   * online softmax in the exp2 domain, packed f32x2 math, lazy rescaling: the running maximum baked into O and l
//     is only advanced when the new block maximum exceeds it by more than 2^8, so O is touched rarely; a
//     configurable share of the exponentials runs as a polynomial on the FMA pipe (MUFU relief).
//   * epilogue: O/l -> bf16 -> swizzled smem (the dead Q tile) -> TMA store (rows >= Sq are clipped).
//
// Measured alternatives that LOST on B200 (kept in git history): 64-key sub-blocks with double-buffered
// score tiles (1045 TFLOP/s: the P-over-S aliasing still chains S(i+1) to P(i-1), and the per-sub-block fixed
// costs double), the same with a software-pipelined softmax (1042), turn-taking between the two softmax
// warpgroups on the MUFU phase (938), two threads per query row / 16 softmax warps (1178: the phase lengths did not
// move, i.e. they are set by shared units -- MUFU 16 lanes/clk, TMEM read port -- not by per-thread issue).
// This kernel: 1271-1279 TFLOP/s at S = 43120, 40 heads.
#include <atomic>
#include <stdlib.h>

#include "attn_common.cuh"

namespace mv {

constexpr int AT_THREADS = 384;
constexpr int AT_NS = 5;  // K/V ring slots
constexpr int AT_TILE_BYTES = 128 * 128 * 2;
constexpr int AT_HALF_BYTES = AT_TILE_BYTES / 2;  // one 64-column (128-byte-wide) swizzle panel
constexpr int AT_OFF_Q = 0;
constexpr int AT_OFF_KV = 2 * AT_TILE_BYTES;
constexpr int AT_OFF_BARS = AT_OFF_KV + AT_NS * AT_TILE_BYTES;
constexpr int AT_BAR_QFULL = 0;                        // [2]
constexpr int AT_BAR_KVFULL = 2;                       // [NS]
constexpr int AT_BAR_KVEMPTY = AT_BAR_KVFULL + AT_NS;  // [NS]
constexpr int AT_BAR_SFULL = AT_BAR_KVEMPTY + AT_NS;   // [2]
constexpr int AT_BAR_PREADY = AT_BAR_SFULL + 2;        // [tile][half of the key block] = [4]
constexpr int AT_BAR_ODONE = AT_BAR_PREADY + 4;        // [2]
constexpr int AT_NUM_BARS = AT_BAR_ODONE + 2;
constexpr int AT_OFF_TMEM_PTR = AT_OFF_BARS + AT_NUM_BARS * 8;
constexpr int AT_SMEM_BYTES = AT_OFF_TMEM_PTR + 16;
static_assert(AT_SMEM_BYTES <= 232448, "attention shared memory budget exceeded");

constexpr uint32_t AT_TMEM_S = 0;    // + 128 * tile
constexpr uint32_t AT_TMEM_O = 256;  // + 128 * tile
constexpr float AT_RESCALE_THRESHOLD = 8.0f;  // log2 units
constexpr int AT_DEFAULT_EMU = 4;
// EMU: how many of every 16 score pairs take the polynomial path (0 = all MUFU, 8 = half and half)
template <int EMU, bool TRACE>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int row_base = blockIdx.x * 256;
  const int nt = (row_base + 128 < p.Sq) ? 2 : 1;  // live query tiles of this CTA
  const int n_kv = (p.Skv + 127) >> 7;
  // event trace of CTA (0,0,0): softmax thread 0 of each tile and the tcgen05 issuer
  int trace_n = 0;
  const bool tracing = TRACE && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && p.trace != nullptr;
  auto ev = [&](int region, int id) {
    if (TRACE && tracing && trace_n < 4096)
      p.trace[region * 4096 + trace_n++] = (static_cast<unsigned long long>(clock64()) << 8) | static_cast<unsigned>(id);
  };

  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bars = smem_base + AT_OFF_BARS;
  auto bar = [&](int idx) { return bars + 8u * idx; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + AT_OFF_TMEM_PTR);

  if (threadIdx.x == 0) {
    if ((smem_base & 1023u) != 0) __trap();  // swizzle-128B tiles need a 1024-byte aligned window
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < 2; ++i) mbar_init(bar(AT_BAR_QFULL + i), 1);
    for (int i = 0; i < AT_NS; ++i) {
      mbar_init(bar(AT_BAR_KVFULL + i), 1);
      mbar_init(bar(AT_BAR_KVEMPTY + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(AT_BAR_SFULL + i), 1);
      mbar_init(bar(AT_BAR_PREADY + 2 * i), 4);  // one arrive per softmax warp
      mbar_init(bar(AT_BAR_PREADY + 2 * i + 1), 4);
      mbar_init(bar(AT_BAR_ODONE + i), 1);
    }
    fence_barrier_init();
  }
  if (warp_idx == 9) tmem_alloc<1>(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx < 8) {
    // =========================== softmax warpgroups ===========================
    setmaxnreg_inc_208();
    const int tile = warp_idx >> 2;
    if (tile < nt) {
      const int r = (warp_idx & 3) * 32 + lane;  // row inside the tile == TMEM lane
      const uint32_t lane_sel = static_cast<uint32_t>((warp_idx & 3) * 32) << 16;
      const uint32_t t_s = tmem_base + lane_sel + AT_TMEM_S + tile * 128;
      const uint32_t t_o = tmem_base + lane_sel + AT_TMEM_O + tile * 128;
      const float c = p.scale_log2;
      const int tail = p.Skv - (n_kv - 1) * 128;  // valid keys of the last block, 1..128
      float m_used = -INFINITY;  // maximum (raw score units) the running O and l are expressed against
      float l = 0.f;

#pragma unroll 1
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(bar(AT_BAR_SFULL + tile), j & 1);
        tc_fence_after();
        if ((threadIdx.x & 127) == 0) ev(tile, 1);
        uint32_t s[128];
#pragma unroll
        for (int q = 0; q < 4; ++q) tmem_ld_x32(t_s + q * 32, reinterpret_cast<uint32_t(&)[32]>(s[q * 32]));
        tmem_wait_ld();
        if ((threadIdx.x & 127) == 0) ev(tile, 2);
        if (j == n_kv - 1 && tail < 128) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i >= tail) s[i] = 0xff800000u;  // -inf
        }
        {
        // 8 independent chains: the 3-input FMNMX has a long dependent-issue latency (4 chains cost ~420 cycles)
        float mx[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) mx[k] = fmaxf(__uint_as_float(s[k]), __uint_as_float(s[k + 8]));
#pragma unroll
        for (int i = 16; i < 128; i += 16) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            mx[k] = fmaxf(mx[k], fmaxf(__uint_as_float(s[i + k]), __uint_as_float(s[i + k + 8])));
        }
        const float m_new = fmaxf(m_used, fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])),
                                                fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7]))));
        if (j == 0) {
          m_used = m_new;  // nothing accumulated yet
        } else {
          const bool need = (m_new - m_used) * c > AT_RESCALE_THRESHOLD;
          if (__any_sync(0xffffffffu, need)) {
            // PV(j-1) of this tile completed before S(j) was committed, so O is quiescent here
            const float f = fast_exp2((m_used - m_new) * c);
            m_used = m_new;
            l *= f;
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {
              uint32_t o[32];
              tmem_ld_x32(t_o + q * 32, o);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
              tmem_st_x32(t_o + q * 32, o);
            }
            tmem_wait_st();
          }
        }
        }
        if ((threadIdx.x & 127) == 0) ev(tile, 3);
        const float neg = -m_used * c;
        const float2 c2 = make_float2(c, c);
        const float2 neg2 = make_float2(neg, neg);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q == 2) {
            // first half of P (keys 0..63) is complete: let the issuer start P.V on it while the second half is
            // still in the exponential unit
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(AT_BAR_PREADY + 2 * tile));
          }
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float2 x = __ffma2_rn(
                make_float2(__uint_as_float(s[q * 32 + 2 * e]), __uint_as_float(s[q * 32 + 2 * e + 1])), c2, neg2);
            float2 pv;
            if (((e + 1) * EMU) / 16 > (e * EMU) / 16) {  // evenly spread, resolved at compile time
              pv = exp2_poly2(x);
            } else {
              pv.x = fast_exp2(x.x);
              pv.y = fast_exp2(x.y);
            }
            s[q * 32 + 2 * e] = __float_as_uint(pv.x);
            s[q * 32 + 2 * e + 1] = __float_as_uint(pv.y);
            pk[e] = pack_bf16x2(pv.x, pv.y);
          }
          tmem_st_x16(t_s + q * 16, pk);
        }
        if ((threadIdx.x & 127) == 0) ev(tile, 4);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(AT_BAR_PREADY + 2 * tile + 1));
        if ((threadIdx.x & 127) == 0) ev(tile, 5);
        // row sum, off the critical path
        float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
#pragma unroll
        for (int i = 0; i < 128; i += 8) {
          a0 = __fadd2_rn(a0, make_float2(__uint_as_float(s[i]), __uint_as_float(s[i + 1])));
          a1 = __fadd2_rn(a1, make_float2(__uint_as_float(s[i + 2]), __uint_as_float(s[i + 3])));
          a2 = __fadd2_rn(a2, make_float2(__uint_as_float(s[i + 4]), __uint_as_float(s[i + 5])));
          a3 = __fadd2_rn(a3, make_float2(__uint_as_float(s[i + 6]), __uint_as_float(s[i + 7])));
        }
        a0 = __fadd2_rn(__fadd2_rn(a0, a1), __fadd2_rn(a2, a3));
        l += a0.x + a0.y;
      }

      // ---- epilogue: O / l -> bf16 -> swizzled smem (dead Q tile) -> TMA store ----
      mbar_wait(bar(AT_BAR_ODONE + tile), 0);
      tc_fence_after();
      const float inv = 1.0f / l;
      const uint32_t stage = smem_base + AT_OFF_Q + tile * AT_TILE_BYTES;
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        uint32_t o[32];
        tmem_ld_x32(t_o + q * 32, o);
        tmem_wait_ld();
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            w[e] = pack_bf16x2(__uint_as_float(o[v * 8 + 2 * e]) * inv, __uint_as_float(o[v * 8 + 2 * e + 1]) * inv);
          const int chunk16 = (q & 1) * 4 + v;
          const uint32_t dst = stage + (q >> 1) * AT_HALF_BYTES + r * 128 + ((chunk16 ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                       "r"(w[3])
                       : "memory");
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + tile, 128);
      const int row0 = row_base + tile * 128;
      if ((threadIdx.x & 127) == 0) {
        tma_store_3d(&tmO, stage, h * 128, row0, b);
        tma_store_3d(&tmO, stage + AT_HALF_BYTES, h * 128 + 64, row0, b);
        tma_store_commit();
        tma_store_wait<0>();
      }
      if (p.lse != nullptr && row0 + r < p.Sq)
        p.lse[(static_cast<long long>(b) * p.H + h) * p.Sq + row0 + r] = m_used * p.scale + __logf(l);
    }
  } else {
    setmaxnreg_dec_88();
    if (warp_idx == 8) {
      // =========================== TMA producer ===========================
      if (elect_one()) {
        const int c0 = h * 128;
        auto load_tile = [&](const CUtensorMap* m, uint32_t dst, uint32_t full, int row) {
          mbar_arrive_expect_tx(full, AT_TILE_BYTES);
          tma_load_3d(dst, m, full, c0, row, b);
          tma_load_3d(dst + AT_HALF_BYTES, m, full, c0 + 64, row, b);
        };
        load_tile(&tmQ, smem_base + AT_OFF_Q, bar(AT_BAR_QFULL + 0), row_base);
        uint32_t slot = 0, phase = 0;
        for (int t = 0; t < 2 * n_kv; ++t) {
          mbar_wait(bar(AT_BAR_KVEMPTY + slot), phase ^ 1);
          load_tile((t & 1) ? &tmV : &tmK, smem_base + AT_OFF_KV + slot * AT_TILE_BYTES, bar(AT_BAR_KVFULL + slot),
                    (t >> 1) * 128);
          if (t == 0 && nt == 2)
            load_tile(&tmQ, smem_base + AT_OFF_Q + AT_TILE_BYTES, bar(AT_BAR_QFULL + 1), row_base + 128);
          if (++slot == AT_NS) { slot = 0; phase ^= 1; }
        }
      }
    } else if (warp_idx == 9) {
      // =========================== tcgen05 issuer ===========================
      // The whole warp walks the schedule (so addresses and descriptors live in uniform registers); one elected
      // lane issues the MMAs and commits.
      constexpr uint32_t IDESC_QK = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t IDESC_PV = umma_idesc_bf16(128, 128, 0, 1);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      auto issue_qk = [&](int tile, uint32_t kbase) {
        const uint64_t qd = umma_desc_k_sw128(smem_base + AT_OFF_Q + tile * AT_TILE_BYTES);
        const uint64_t kd = umma_desc_k_sw128(kbase);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t off16 = ((ks >> 2) * AT_HALF_BYTES + (ks & 3) * 32) >> 4;
            umma_ss<1>(tmem_u + AT_TMEM_S + tile * 128, qd + off16, kd + off16, IDESC_QK, ks > 0 ? 1u : 0u);
          }
          umma_commit(bar(AT_BAR_SFULL + tile));
        }
        __syncwarp();
      };
      // P.V over keys [64*half, 64*half + 64) of the block
      auto issue_pv_half = [&](int tile, uint32_t vbase, int half, bool acc) {
        const uint64_t vd = umma_desc_mn_sw128(vbase + half * 8192, AT_HALF_BYTES, 1024);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            umma_ts(tmem_u + AT_TMEM_O + tile * 128, tmem_u + AT_TMEM_S + tile * 128 + half * 32 + ks * 8,
                    vd + ((ks * 2048) >> 4), IDESC_PV, (acc || half > 0 || ks > 0) ? 1u : 0u);
          }
        }
        __syncwarp();
      };
      auto commit = [&](uint32_t b_) {
        if (elect_one()) umma_commit(b_);
        __syncwarp();
      };
      uint32_t slot = 0, phase = 0;
      auto next_slot = [&]() { if (++slot == AT_NS) { slot = 0; phase ^= 1; } };
      auto slot_addr = [&](uint32_t s_) { return smem_base + AT_OFF_KV + s_ * AT_TILE_BYTES; };

      // S(0) of both tiles
      mbar_wait(bar(AT_BAR_QFULL + 0), 0);
      mbar_wait(bar(AT_BAR_KVFULL + slot), phase);
      tc_fence_after();
      issue_qk(0, slot_addr(slot));
      if (nt == 2) {
        mbar_wait(bar(AT_BAR_QFULL + 1), 0);
        tc_fence_after();
        issue_qk(1, slot_addr(slot));
      }
      commit(bar(AT_BAR_KVEMPTY + slot));
      next_slot();

      for (int j = 0; j < n_kv; ++j) {
        const bool last = (j == n_kv - 1);
        const uint32_t vslot = slot;
        mbar_wait(bar(AT_BAR_KVFULL + vslot), phase);
        next_slot();
        const uint32_t kslot = slot;
        const uint32_t kphase = phase;
        for (int tile = 0; tile < nt; ++tile) {
          // one barrier per half of P: a single two-phase barrier would let the softmax run two phases ahead of
          // this warp, which a parity wait cannot tell apart from "not there yet"
          mbar_wait(bar(AT_BAR_PREADY + 2 * tile), j & 1);
          tc_fence_after();
          if (tile == 0) ev(2, 10); else ev(2, 11);
          issue_pv_half(tile, slot_addr(vslot), 0, j > 0);
          mbar_wait(bar(AT_BAR_PREADY + 2 * tile + 1), j & 1);
          tc_fence_after();
          issue_pv_half(tile, slot_addr(vslot), 1, true);
          if (tile == 0) ev(2, 12); else ev(2, 13);
          if (last) commit(bar(AT_BAR_ODONE + tile));
          if (!last) {
            if (tile == 0) {
              mbar_wait(bar(AT_BAR_KVFULL + kslot), kphase);
              tc_fence_after();
            }
            issue_qk(tile, slot_addr(kslot));
            if (tile == 0) ev(2, 14); else ev(2, 15);
          }
        }
        commit(bar(AT_BAR_KVEMPTY + vslot));
        if (!last) {
          commit(bar(AT_BAR_KVEMPTY + kslot));
          next_slot();
        }
      }
    }
  }

  // =========================== teardown ===========================
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 9) tmem_dealloc<1>(tmem_base, 512);
}

template <int EMU, bool TRACE>
static int launch_attn(dim3 grid, cudaStream_t stream, const CUtensorMap& tmQ, const CUtensorMap& tmK,
                       const CUtensorMap& tmV, const CUtensorMap& tmO, const AttnParams& p) {
  auto kernel = attn_fwd_kernel<EMU, TRACE>;
  static std::atomic<bool> configured[64];  // zero-initialised; per-device "attribute set" latch, safe across host threads
  int dev = 0;
  MV_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
    MV_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES));
    if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
  }
  kernel<<<grid, AT_THREADS, AT_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, tmO, p);
  MV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mv

// The schedule mova_b200_attn_fwd runs (chosen from the A/B measurements in profiles/): 92 = CTA-pair kernel
#ifndef MOVA_ATTN_DEFAULT_VARIANT
#define MOVA_ATTN_DEFAULT_VARIANT 92
#endif
#ifndef MOVA_ATTN_SHORT_KV_BLOCKS
#define MOVA_ATTN_SHORT_KV_BLOCKS 8
#endif
#ifndef MOVA_ATTN_DEFAULT_EMU
#define MOVA_ATTN_DEFAULT_EMU 4
#endif

extern "C" int mova_b200_attn_fwd(const void* q, int64_t q_bs, int64_t q_ss, const void* k, int64_t k_bs, int64_t k_ss,
                                  const void* v, int64_t v_bs, int64_t v_ss, void* o, int64_t o_bs, int64_t o_ss,
                                  float* lse, int B, int Sq, int Skv, int H, int D, float softmax_scale,
                                  void* stream) {
  // Schedule by shape (measured on B200, profiles/r02_attn_schedules.log):
  //  * long key sequences (video self-attention, v2a bridge): the round-2 CTA-pair kernel -- 1325 vs 1273 TFLOP/s at
  //    43120 x 43120 x 40 heads, and twice the CTAs for the 403-query v2a shape (445 vs 211 TFLOP/s);
  //  * a handful of key blocks against many queries (text cross-attention: 512 keys, a2v bridge: 403 keys): the per-CTA
  //    prologue / epilogue dominates, and the round-1 kernel amortises it over 256 query rows per CTA (920 vs 613).
  const int n_kv = (Skv + 127) / 128;
  const int variant = (n_kv <= MOVA_ATTN_SHORT_KV_BLOCKS && Sq > 256) ? 3 : MOVA_ATTN_DEFAULT_VARIANT;
  return mova_b200_attn_fwd_variant(q, q_bs, q_ss, k, k_bs, k_ss, v, v_bs, v_ss, o, o_bs, o_ss, lse, B, Sq, Skv, H, D,
                                    softmax_scale, variant, MOVA_ATTN_DEFAULT_EMU, nullptr, stream);
}

extern "C" int mova_b200_attn_fwd_variant(const void* q, int64_t q_bs, int64_t q_ss, const void* k, int64_t k_bs,
                                          int64_t k_ss, const void* v, int64_t v_bs, int64_t v_ss, void* o,
                                          int64_t o_bs, int64_t o_ss, float* lse, int B, int Sq, int Skv, int H, int D,
                                          float softmax_scale, int variant, int emu, unsigned long long* trace,
                                          void* stream) {
  using namespace mv;
  MV_REQUIRE(q && k && v && o, "mova_b200_attn_fwd: null pointer");
  MV_REQUIRE(D == 128, "mova_b200_attn_fwd: head_dim %d unsupported (the MOVA towers and bridge use 128)", D);
  MV_REQUIRE(B >= 1 && H >= 1 && Sq >= 0 && Skv >= 1, "mova_b200_attn_fwd: bad shape B=%d Sq=%d Skv=%d H=%d", B, Sq,
             Skv, H);
  MV_REQUIRE(B <= 65535 && H <= 65535, "mova_b200_attn_fwd: B and H must fit a grid dimension");
  MV_REQUIRE(softmax_scale > 0.f, "mova_b200_attn_fwd: softmax_scale must be positive");
  MV_REQUIRE(variant == 3 || variant == 92, "mova_b200_attn_fwd_variant: variant must be 92 (round-2 schedule, CTA "
             "pair) or 3 (round-1 schedule); got %d", variant);
  MV_REQUIRE(emu == 4 || (variant == 3 && (emu == 0 || emu == 8)),
             "mova_b200_attn_fwd_variant: exp2 emulation share must be 4 (variant 3 also has 0 and 8); got %d", emu);
  const int64_t row = static_cast<int64_t>(H) * D;
  MV_REQUIRE(q_ss >= row && k_ss >= row && v_ss >= row && o_ss >= row,
             "mova_b200_attn_fwd: sequence stride smaller than H*D");
  if (Sq == 0) return 0;
  // a batch stride is irrelevant (and may be anything) when B == 1
  if (B == 1) {
    q_bs = q_ss * Sq;
    o_bs = o_ss * Sq;
    k_bs = k_ss * Skv;
    v_bs = v_ss * Skv;
  }

  // K / V boxes: one key block per load (K: this CTA's half of the block in the CTA-pair kernels)
  const int bn = 128;
  const int cg = (variant == 92) ? 2 : 1;
  const uint32_t k_box_rows = static_cast<uint32_t>(bn / cg);
  // O box: the round-2 kernels store one 64-column panel per warpgroup
  CUtensorMap tmQ, tmK, tmV, tmO;
  int rc;
  if ((rc = encode_tmap_3d(&tmQ, q, row, Sq, B, q_ss, q_bs, 64, 128, 1)) != 0) return rc;
  if ((rc = encode_tmap_3d(&tmK, k, row, Skv, B, k_ss, k_bs, 64, k_box_rows, 1)) != 0) return rc;
  if ((rc = encode_tmap_3d(&tmV, v, row, Skv, B, v_ss, v_bs, 64, static_cast<uint32_t>(bn), 1)) != 0) return rc;
  if ((rc = encode_tmap_3d(&tmO, o, row, Sq, B, o_ss, o_bs, 64, 128, 1)) != 0) return rc;

  AttnParams p;
  p.Sq = Sq;
  p.Skv = Skv;
  p.H = H;
  p.scale = softmax_scale;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.lse = lse;
  p.trace = trace;

  debug_attach();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (variant != 3) return launch_attn_pair(cg, bn, emu, trace != nullptr, B, Sq, H, st, tmQ, tmK, tmV, tmO, p);
  dim3 grid((Sq + 255) / 256, H, B);
  if (trace != nullptr) return launch_attn<4, true>(grid, st, tmQ, tmK, tmV, tmO, p);
  switch (emu) {
    case 0: return launch_attn<0, false>(grid, st, tmQ, tmK, tmV, tmO, p);
    case 8: return launch_attn<8, false>(grid, st, tmQ, tmK, tmV, tmO, p);
    default: return launch_attn<4, false>(grid, st, tmQ, tmK, tmV, tmO, p);
  }
}
