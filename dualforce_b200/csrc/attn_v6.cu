// EXPERIMENTAL attention schedule (selected with MOVA_ATTN_VARIANT=v6; attn.cu's v3 schedule is the default).
// Status (round 1, B200): bit-for-bit the same results as v3 and the whole -m gpu suite passes with it, but it is
// SLOWER: 477 TFLOP/s with a try_wait-based poll (try_wait parks the thread), 835 with test_wait, vs 1267 for v3 at
// S = 43120 x 40 heads.  The dependency graph below is sound; the cost is in the polling issuer and in paying the
// per-sub-block fixed costs (TMEM load, max chain, store + arrive) twice per 128 keys.  Kept for round 2.
//
// Same contract, tiles, K/V ring, softmax arithmetic and epilogue as attn.cu; what changes is the dependency graph
// between the softmax warps and the tensor core.  The event timelines of v1/v3 (profiles/README.md) show each query
// tile alternating between ~2100 busy softmax cycles and a ~1200-cycle wait for the next S tile, because P is written
// over S: QK(j+1) cannot be issued before PV(j) has consumed P(j).  Here
//   * keys are consumed in 64-key sub-blocks; TMEM holds ONE 64-column score tile per query tile (S0 | S1, columns
//     0..127), a DOUBLE-BUFFERED 32-column bf16 P tile per query tile in its own columns (128..255) and the two
//     128-column output accumulators (256..511);
//   * as soon as the softmax warps have pulled S(i) into registers they release the score columns (SFREE), so the
//     tensor core computes S(i+1) WHILE the exponentials of sub-block i run; P(i) lands in buffer i&1 and P.V(i) is
//     issued whenever it is ready -- the softmax only ever waits for the tensor core if the tensor core is the
//     bottleneck;
//   * the issuer warp is event driven: it polls (SFREE -> issue QK) and (PREADY -> issue PV) of both tiles instead of
//     walking a fixed order, so a slow tile never blocks the other.
// Barrier-phase safety (a parity wait cannot tell "two phases ahead" from "not yet"): every producer is gated on its
// consumer -- QK(i+1) needs SFREE(i), SFREE(i) needs S(i); P(i+2) needs PVDONE(i), PV(i) needs P(i) -- so no barrier
// can run more than one phase ahead of its waiter.
#include <stdlib.h>

#include "attn_common.cuh"

namespace mv {

constexpr int V6_THREADS = 384;
constexpr int V6_NS = 5;
constexpr int V6_TILE_BYTES = 128 * 128 * 2;
constexpr int V6_HALF_BYTES = V6_TILE_BYTES / 2;
constexpr int V6_OFF_Q = 0;
constexpr int V6_OFF_KV = 2 * V6_TILE_BYTES;
constexpr int V6_OFF_BARS = V6_OFF_KV + V6_NS * V6_TILE_BYTES;
constexpr int V6_BAR_QFULL = 0;                        // [2]
constexpr int V6_BAR_KVFULL = 2;                       // [NS]
constexpr int V6_BAR_KVEMPTY = V6_BAR_KVFULL + V6_NS;  // [NS]
constexpr int V6_BAR_SFULL = V6_BAR_KVEMPTY + V6_NS;   // [tile]          issuer  -> softmax, phase per sub-block
constexpr int V6_BAR_SFREE = V6_BAR_SFULL + 2;         // [tile]          softmax -> issuer,  phase per sub-block
constexpr int V6_BAR_PREADY = V6_BAR_SFREE + 2;        // [tile][buffer]  softmax -> issuer,  phase per 2 sub-blocks
constexpr int V6_BAR_PVDONE = V6_BAR_PREADY + 4;       // [tile][buffer]  issuer  -> softmax, phase per 2 sub-blocks
constexpr int V6_BAR_ODONE = V6_BAR_PVDONE + 4;        // [tile]
constexpr int V6_NUM_BARS = V6_BAR_ODONE + 2;
constexpr int V6_OFF_TMEM_PTR = V6_OFF_BARS + V6_NUM_BARS * 8;
constexpr int V6_SMEM_BYTES = V6_OFF_TMEM_PTR + 16;
static_assert(V6_SMEM_BYTES <= 232448, "attention shared memory budget exceeded");

constexpr uint32_t V6_TMEM_S = 0;    // + 64 * tile
constexpr uint32_t V6_TMEM_P = 128;  // + 64 * tile + 32 * buffer
constexpr uint32_t V6_TMEM_O = 256;  // + 128 * tile
constexpr float V6_RESCALE_THRESHOLD = 8.0f;

template <int EMU>
__global__ void __launch_bounds__(V6_THREADS, 1)
attn_fwd_v6_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                   const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int row_base = blockIdx.x * 256;
  const int nt = (row_base + 128 < p.Sq) ? 2 : 1;  // live query tiles of this CTA
  const int n_kv = (p.Skv + 127) >> 7;             // 128-key K/V tiles
  const int nsub = (p.Skv + 63) >> 6;              // 64-key sub-blocks

  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bars = smem_base + V6_OFF_BARS;
  auto bar = [&](int idx) { return bars + 8u * idx; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + V6_OFF_TMEM_PTR);

  if (threadIdx.x == 0) {
    if ((smem_base & 1023u) != 0) __trap();  // swizzle-128B tiles need a 1024-byte aligned window
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < 2; ++i) mbar_init(bar(V6_BAR_QFULL + i), 1);
    for (int i = 0; i < V6_NS; ++i) {
      mbar_init(bar(V6_BAR_KVFULL + i), 1);
      mbar_init(bar(V6_BAR_KVEMPTY + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(V6_BAR_SFULL + i), 1);
      mbar_init(bar(V6_BAR_SFREE + i), 4);  // one arrive per softmax warp
      mbar_init(bar(V6_BAR_ODONE + i), 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar(V6_BAR_PREADY + i), 4);
      mbar_init(bar(V6_BAR_PVDONE + i), 1);
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
      const uint32_t t_s = tmem_base + lane_sel + V6_TMEM_S + tile * 64;
      const uint32_t t_p = tmem_base + lane_sel + V6_TMEM_P + tile * 64;
      const uint32_t t_o = tmem_base + lane_sel + V6_TMEM_O + tile * 128;
      const float c = p.scale_log2;
      const int tail = p.Skv - (nsub - 1) * 64;  // valid keys of the last sub-block, 1..64
      float m_used = -INFINITY;  // maximum (raw score units) the running O and l are expressed against
      float l = 0.f;

#pragma unroll 1
      for (int i = 0; i < nsub; ++i) {
        const int buf = i & 1;
        mbar_wait(bar(V6_BAR_SFULL + tile), i & 1);
        tc_fence_after();
        uint32_t s[64];
        tmem_ld_x32(t_s, reinterpret_cast<uint32_t(&)[32]>(s[0]));
        tmem_ld_x32(t_s + 32, reinterpret_cast<uint32_t(&)[32]>(s[32]));
        tmem_wait_ld();
        // the scores are in registers: hand the S columns back so Q.K^T of sub-block i+1 overlaps this sub-block
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(V6_BAR_SFREE + tile));
        if (i == nsub - 1 && tail < 64) {
#pragma unroll
          for (int k = 0; k < 64; ++k)
            if (k >= tail) s[k] = 0xff800000u;  // -inf
        }
        float mx[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) mx[k] = fmaxf(__uint_as_float(s[k]), __uint_as_float(s[k + 8]));
#pragma unroll
        for (int o = 16; o < 64; o += 16) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            mx[k] = fmaxf(mx[k], fmaxf(__uint_as_float(s[o + k]), __uint_as_float(s[o + k + 8])));
        }
        const float m_new = fmaxf(
            m_used, fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7]))));
        if (i == 0) {
          m_used = m_new;  // nothing accumulated yet
        } else {
          const bool need = (m_new - m_used) * c > V6_RESCALE_THRESHOLD;
          if (__any_sync(0xffffffffu, need)) {
            // O must be quiescent: P.V(i-1) has to be complete.  Sub-block i-1 waited for P.V(i-3) on the same
            // barrier before it wrote its P, so that barrier is at most one phase behind: the parity is unambiguous.
            mbar_wait(bar(V6_BAR_PVDONE + tile * 2 + (buf ^ 1)), ((i - 1) >> 1) & 1);
            tc_fence_after();
            const float f = fast_exp2((m_used - m_new) * c);
            m_used = m_new;
            l *= f;
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {
              uint32_t o[32];
              tmem_ld_x32(t_o + q * 32, o);
              tmem_wait_ld();
#pragma unroll
              for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * f);
              tmem_st_x32(t_o + q * 32, o);
            }
            tmem_wait_st();
          }
        }
        // P buffer `buf` was last read by P.V(i-2)
        if (i >= 2) {
          mbar_wait(bar(V6_BAR_PVDONE + tile * 2 + buf), ((i - 2) >> 1) & 1);
          tc_fence_after();
        }
        const float neg = -m_used * c;
        const float2 c2 = make_float2(c, c);
        const float2 neg2 = make_float2(neg, neg);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
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
          tmem_st_x16(t_p + buf * 32 + q * 16, pk);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(V6_BAR_PREADY + tile * 2 + buf));
        float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
#pragma unroll
        for (int k = 0; k < 64; k += 8) {
          a0 = __fadd2_rn(a0, make_float2(__uint_as_float(s[k]), __uint_as_float(s[k + 1])));
          a1 = __fadd2_rn(a1, make_float2(__uint_as_float(s[k + 2]), __uint_as_float(s[k + 3])));
          a2 = __fadd2_rn(a2, make_float2(__uint_as_float(s[k + 4]), __uint_as_float(s[k + 5])));
          a3 = __fadd2_rn(a3, make_float2(__uint_as_float(s[k + 6]), __uint_as_float(s[k + 7])));
        }
        a0 = __fadd2_rn(__fadd2_rn(a0, a1), __fadd2_rn(a2, a3));
        l += a0.x + a0.y;
      }

      // ---- epilogue: O / l -> bf16 -> swizzled smem (dead Q tile) -> TMA store ----
      mbar_wait(bar(V6_BAR_ODONE + tile), 0);
      tc_fence_after();
      const float inv = 1.0f / l;
      const uint32_t stage = smem_base + V6_OFF_Q + tile * V6_TILE_BYTES;
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
          const uint32_t dst = stage + (q >> 1) * V6_HALF_BYTES + r * 128 + ((chunk16 ^ (r & 7)) << 4);
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
        tma_store_3d(&tmO, stage + V6_HALF_BYTES, h * 128 + 64, row0, b);
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
          mbar_arrive_expect_tx(full, V6_TILE_BYTES);
          tma_load_3d(dst, m, full, c0, row, b);
          tma_load_3d(dst + V6_HALF_BYTES, m, full, c0 + 64, row, b);
        };
        load_tile(&tmQ, smem_base + V6_OFF_Q, bar(V6_BAR_QFULL + 0), row_base);
        uint32_t slot = 0, phase = 0;
        for (int t = 0; t < 2 * n_kv; ++t) {
          mbar_wait(bar(V6_BAR_KVEMPTY + slot), phase ^ 1);
          load_tile((t & 1) ? &tmV : &tmK, smem_base + V6_OFF_KV + slot * V6_TILE_BYTES, bar(V6_BAR_KVFULL + slot),
                    (t >> 1) * 128);
          if (t == 0 && nt == 2)
            load_tile(&tmQ, smem_base + V6_OFF_Q + V6_TILE_BYTES, bar(V6_BAR_QFULL + 1), row_base + 128);
          if (++slot == V6_NS) { slot = 0; phase ^= 1; }
        }
      }
    } else if (warp_idx == 9) {
      // =========================== tcgen05 issuer (event driven, warp uniform) ===========================
      constexpr uint32_t IDESC_QK = umma_idesc_bf16(128, 64, 0, 0);   // S_sub[128 x 64] = Q[128 x 128] . K_sub^T
      constexpr uint32_t IDESC_PV = umma_idesc_bf16(128, 128, 0, 1);  // O[128 x 128] += P_sub[128 x 64] . V_sub
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      auto slot_addr = [&](int idx) { return smem_base + V6_OFF_KV + (idx % V6_NS) * V6_TILE_BYTES; };
      // a test every lane evaluates; "any lane saw the phase complete" is a warp-uniform fact
      auto ready = [&](uint32_t b_, uint32_t parity) { return __any_sync(0xffffffffu, mbar_test_wait(b_, parity)) != 0; };
      // K / V tiles whose arrival this warp has already observed (ring indices 2j / 2j+1).  Never block here: the
      // producer may be waiting for a slot that only a P.V this very loop has yet to issue can release.
      int k_seen = 0, v_seen = 0;
      auto have_k = [&](int j) {
        while (k_seen <= j) {
          if (!ready(bar(V6_BAR_KVFULL + (2 * k_seen) % V6_NS), ((2 * k_seen) / V6_NS) & 1)) return false;
          ++k_seen;
        }
        return true;
      };
      auto have_v = [&](int j) {
        while (v_seen <= j) {
          if (!ready(bar(V6_BAR_KVFULL + (2 * v_seen + 1) % V6_NS), ((2 * v_seen + 1) / V6_NS) & 1)) return false;
          ++v_seen;
        }
        return true;
      };
      auto issue_qk = [&](int tile, int sub) {
        const uint64_t qd = umma_desc_k_sw128(smem_base + V6_OFF_Q + tile * V6_TILE_BYTES);
        const uint64_t kd = umma_desc_k_sw128(slot_addr(2 * (sub >> 1)) + (sub & 1) * 8192);  // keys [64*(sub&1), +64)
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t off16 = ((ks >> 2) * V6_HALF_BYTES + (ks & 3) * 32) >> 4;
            umma_ss<1>(tmem_u + V6_TMEM_S + tile * 64, qd + off16, kd + off16, IDESC_QK, ks > 0 ? 1u : 0u);
          }
          umma_commit(bar(V6_BAR_SFULL + tile));
        }
        __syncwarp();
      };
      auto issue_pv = [&](int tile, int sub) {
        const uint64_t vd = umma_desc_mn_sw128(slot_addr(2 * (sub >> 1) + 1) + (sub & 1) * 8192, V6_HALF_BYTES, 1024);
        const uint32_t a_tmem = tmem_u + V6_TMEM_P + tile * 64 + (sub & 1) * 32;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_ts(tmem_u + V6_TMEM_O + tile * 128, a_tmem + ks * 8, vd + ((ks * 2048) >> 4), IDESC_PV,
                    (sub > 0 || ks > 0) ? 1u : 0u);
          umma_commit(bar(V6_BAR_PVDONE + tile * 2 + (sub & 1)));
          if (sub == nsub - 1) umma_commit(bar(V6_BAR_ODONE + tile));
        }
        __syncwarp();
      };
      auto release = [&](int ring_idx) {
        if (elect_one()) umma_commit(bar(V6_BAR_KVEMPTY + ring_idx % V6_NS));
        __syncwarp();
      };

      int qk_next[2] = {0, 0};  // next sub-block whose Q.K^T is to be issued, per tile
      int pv_next[2] = {0, 0};  // next sub-block whose P.V is to be issued, per tile
      int k_rel = 0, v_rel = 0;  // K / V tiles already handed back to the producer
      if (nt == 1) { qk_next[1] = nsub; pv_next[1] = nsub; }
      mbar_wait(bar(V6_BAR_QFULL + 0), 0);
      if (nt == 2) mbar_wait(bar(V6_BAR_QFULL + 1), 0);
      tc_fence_after();
      while (pv_next[0] < nsub || pv_next[1] < nsub) {
#pragma unroll
        for (int tile = 0; tile < 2; ++tile) {
          // Q.K^T of the next sub-block as soon as the softmax has pulled the previous scores out of TMEM
          const int iq = qk_next[tile];
          if (iq < nsub && (iq == 0 || ready(bar(V6_BAR_SFREE + tile), (iq - 1) & 1)) && have_k(iq >> 1)) {
            tc_fence_after();
            issue_qk(tile, iq);
            qk_next[tile] = iq + 1;
          }
          // P.V of the oldest sub-block whose P has landed
          const int ip = pv_next[tile];
          if (ip < nsub && ready(bar(V6_BAR_PREADY + tile * 2 + (ip & 1)), (ip >> 1) & 1) && have_v(ip >> 1)) {
            tc_fence_after();
            issue_pv(tile, ip);
            pv_next[tile] = ip + 1;
          }
        }
        // hand K / V tiles back once both query tiles have issued every MMA that reads them
        const int q_min = min(qk_next[0], qk_next[1]);
        const int p_min = min(pv_next[0], pv_next[1]);
        while (k_rel < n_kv && (q_min >= 2 * k_rel + 2 || q_min >= nsub)) { release(2 * k_rel); ++k_rel; }
        while (v_rel < n_kv && (p_min >= 2 * v_rel + 2 || p_min >= nsub)) { release(2 * v_rel + 1); ++v_rel; }
      }
    }
  }

  // =========================== teardown ===========================
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 9) tmem_dealloc<1>(tmem_base, 512);
}

template <int EMU>
static int launch_v6(dim3 grid, cudaStream_t stream, const CUtensorMap& tmQ, const CUtensorMap& tmK,
                     const CUtensorMap& tmV, const CUtensorMap& tmO, const AttnParams& p) {
  auto kernel = attn_fwd_v6_kernel<EMU>;
  static bool configured[64] = {false};
  int dev = 0;
  MV_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    MV_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, V6_SMEM_BYTES));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  kernel<<<grid, V6_THREADS, V6_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, tmO, p);
  MV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_attn_v6(dim3 grid, cudaStream_t stream, const CUtensorMap& tmQ, const CUtensorMap& tmK,
                   const CUtensorMap& tmV, const CUtensorMap& tmO, const AttnParams& p, int emu) {
  debug_attach();
  switch (emu) {
    case 0: return launch_v6<0>(grid, stream, tmQ, tmK, tmV, tmO, p);
    case 4: return launch_v6<4>(grid, stream, tmQ, tmK, tmV, tmO, p);
    case 8: return launch_v6<8>(grid, stream, tmQ, tmK, tmV, tmO, p);
    default: MV_REQUIRE(false, "attention variant v6 is built for MOVA_ATTN_EMU 0, 4 and 8 only (got %d)", emu);
  }
  return 0;
}

}  // namespace mv
