// Non-causal softmax attention forward on tcgen05, head_dim 128:  O = softmax(Q K^T * scale) V  -- the round-2 kernel.
//
// Replaces flash_attention() (mova/diffusion/models/wan_video_dit.py:58-91) at its three call sites: video / audio
// self-attention (:188), text cross-attention (:241) and the a2v / v2a bridge (mova/diffusion/models/interactionv2.py:250).
// q/k/v/o stay in the reference's flat [B, S, H*D] layout (head h = columns [128h, 128h+128)), addressed through 3-D TMA
// tensor maps -- no rearrange copies.
//
// Why a new schedule.  The round-1 kernel (two 128-row query tiles per CTA, one score tile S and one accumulator O per
// query tile, P written over S) is bound by a dependent chain per query tile
//      softmax(j) -> P.V(j) -> Q.K^T(j+1) -> softmax(j+1)
// because Q.K^T(j+1) overwrites the tensor-memory columns P(j) lives in: ncu shows the tensor pipe 74 % active and
// cuDNN's fused attention 8-20 % ahead on the same box (profiles/r02_kernels_vs_libs.jsonl).  All 512 TMEM columns
// were in use (2 x S + 2 x O), so the chain could not be cut by double buffering.  This kernel re-budgets TMEM:
//
//   * one CTA = ONE 128-row query tile x one head;  TMEM = S0 | S1 | S2 | O  (3 score buffers, one accumulator);
//   * Q.K^T of block j+3 is issued as soon as P.V(j) has been issued, i.e. up to two blocks ahead of the softmax, so
//     the score tile of the next block is ready before a softmax warp asks for it;
//   * two softmax warpgroups work on the SAME query rows and take alternate key blocks (A: even, B: odd), thread ==
//     query row == TMEM lane in both.  They share one accumulator O and one reference maximum m per row:
//       - the owner of block j reads m(j-1) from shared memory (published by the other warpgroup; one mbarrier token per
//         warp pair), decides m(j) with the lazy-rescale rule (only advance m when the block maximum exceeds it by more
//         than 2^8) and publishes it -- a ~100-cycle serial section per block, everything else runs concurrently;
//       - if m advances (rare), the owner waits for P.V(j-1) to complete (mbarrier committed by the issuer), rescales
//         O in place, and only then releases P(j); the other warpgroup rebases its private row sum when it next sees
//         the new m.  Row sums are private (l_A, l_B) and added at the end;
//   * CG = 2: two CTAs on one TPC form a pair (`cta_group::2`): the MMAs are M = 256 (each CTA its own 128 query rows),
//     the 128-key K tile and the V tile are split between the two CTAs' shared memories (K by keys, V by head-dim
//     columns), so each SM reads half the B-operand bytes per FLOP from shared memory and loads half the K/V bytes
//     from L2 -- on a power-capped part (sw_power_cap is active in every run) joules per FLOP are throughput;
//   * K and V stream through one ring in consumption order  K0 K1 K2 V0 K3 V1 K4 ...;  V is consumed MN-major (no
//     transpose);  P is released to the tensor core in two halves (keys 0-63, 64-127);
//   * online softmax in the exp2 domain, packed f32x2 math, a configurable share of the exponentials evaluated as a
//     polynomial on the FMA pipe (MUFU relief);
//   * epilogue: the two warpgroups take 64 output columns each: O / (l_A + l_B) -> bf16 -> swizzled smem (the dead Q
//     tile) -> TMA store (rows >= Sq are clipped).
#include <atomic>
#include <stdlib.h>

#include "attn_common.cuh"

namespace mv {

constexpr int AP_THREADS = 384;
constexpr int AP_TILE_BYTES = 128 * 128 * 2;   // one 128 x 128 bf16 tile
constexpr int AP_PANEL_BYTES = 128 * 64 * 2;   // 128 rows of one 64-column (128-byte) swizzle panel
constexpr uint32_t AP_TMEM_O = 384;            // S buffers at columns 0 / 128 / 256
constexpr float AP_RESCALE_THRESHOLD = 8.0f;   // log2 units

// BN: keys per block.  128 -> 3 score buffers, 96 -> 4 (the accumulator takes 128 of the 512 TMEM columns): with the
// same tensor memory a fourth buffer lets Q.K^T run three blocks ahead of the softmax instead of two.
template <int CG, int BN>
struct APCfg {
  static_assert(BN == 128 || BN == 96, "key block must be 128 or 96");
  static constexpr int NBUF = 384 / BN;                        // score buffers (columns [BN * i, BN * i + BN))
  static constexpr int SLOT_BYTES = BN * 128 * 2 / CG;         // this CTA's part of the K (or V) tile of one key block
  static constexpr int NS = 2 * NBUF * CG;                     // ring slots
  static constexpr int K_PANEL_BYTES = (BN / CG) * 128;        // one 64-column panel of this CTA's K part
  static constexpr int V_PANEL_BYTES = BN * 128;               // one 64-column panel of the V tile (all BN keys)
  static constexpr int PV_KSTEPS = BN / 16;                    // tcgen05 K steps of P.V; the first 4 (64 keys) are part 0
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_KV = AP_TILE_BYTES;
  static constexpr int OFF_MREF = OFF_KV + NS * SLOT_BYTES;  // float[128]: reference maximum per query row
  static constexpr int OFF_LSUM = OFF_MREF + 512;            // float[2][128]: private row sums, exchanged at the end
  static constexpr int OFF_BARS = OFF_LSUM + 1024;
  static constexpr int BAR_QFULL = 0;
  static constexpr int BAR_KVFULL = 1;                  // [NS]
  static constexpr int BAR_KVEMPTY = BAR_KVFULL + NS;   // [NS]
  static constexpr int BAR_SFULL = BAR_KVEMPTY + NS;    // [NBUF]
  static constexpr int BAR_PREADY = BAR_SFULL + NBUF;   // [NBUF score buffers][2 parts of the key block]
  static constexpr int BAR_PVDONE = BAR_PREADY + 2 * NBUF;  // [2] (block parity)
  static constexpr int BAR_TOKEN = BAR_PVDONE + 2;      // [2 warpgroups][4 warps]: "m of my block is published"
  static constexpr int NUM_BARS = BAR_TOKEN + 8;
  static constexpr int OFF_TMEM_PTR = OFF_BARS + NUM_BARS * 8;
  static constexpr int SMEM_BYTES = OFF_TMEM_PTR + 16;
  static_assert(SMEM_BYTES <= 232448, "attention shared memory budget exceeded");
};

// D[tmem] (+)= A[tmem] * B[smem], single CTA or CTA pair
template <int CG>
__device__ __forceinline__ void umma_ts_cg(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  if constexpr (CG == 1) {
    umma_ts(d_tmem, a_tmem, b_desc, idesc, accumulate);
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// EMU: how many of every 16 score pairs take the polynomial exp2 (0 = all MUFU, 8 = half and half)
template <int CG, int BN, int EMU, bool TRACE>
__global__ void __launch_bounds__(AP_THREADS, 1)
attn_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                 const AttnParams p) {
  using Cfg = APCfg<CG, BN>;
  constexpr int NBUF = Cfg::NBUF;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool is_leader = (rank == 0);
  const int row0 = blockIdx.x * 128;  // this CTA's query tile (may lie entirely beyond Sq in the last pair)
  const int n_kv = (p.Skv + BN - 1) / BN;

  int trace_n = 0;
  const bool tracing = TRACE && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && p.trace != nullptr;
  auto ev = [&](int region, int id) {
    if (TRACE && tracing && trace_n < 4096)
      p.trace[region * 4096 + trace_n++] = (static_cast<unsigned long long>(clock64()) << 8) | static_cast<unsigned>(id);
  };

  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bars = smem_base + Cfg::OFF_BARS;
  auto bar = [&](int idx) { return bars + 8u * idx; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + Cfg::OFF_TMEM_PTR);
  volatile float* mref = reinterpret_cast<volatile float*>(smem + Cfg::OFF_MREF);
  volatile float* lsum = reinterpret_cast<volatile float*>(smem + Cfg::OFF_LSUM);

  if (threadIdx.x == 0) {
    if ((smem_base & 1023u) != 0) __trap();  // swizzle-128B tiles need a 1024-byte aligned window
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    mbar_init(bar(Cfg::BAR_QFULL), 1);
    for (int i = 0; i < Cfg::NS; ++i) {
      mbar_init(bar(Cfg::BAR_KVFULL + i), 1);   // the leader's expect_tx (covers both CTAs' bytes when CG == 2)
      mbar_init(bar(Cfg::BAR_KVEMPTY + i), 1);  // one tcgen05.commit
    }
    for (int i = 0; i < NBUF; ++i) {
      mbar_init(bar(Cfg::BAR_SFULL + i), 1);
      mbar_init(bar(Cfg::BAR_PREADY + 2 * i), 4 * CG);  // one arrive per softmax warp of the owning warpgroup(s)
      mbar_init(bar(Cfg::BAR_PREADY + 2 * i + 1), 4 * CG);
    }
    for (int i = 0; i < 2; ++i) mbar_init(bar(Cfg::BAR_PVDONE + i), 1);
    for (int i = 0; i < 8; ++i) mbar_init(bar(Cfg::BAR_TOKEN + i), 1);
    fence_barrier_init();
  }
  if (warp_idx == 9) tmem_alloc<CG>(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), 512);
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx < 8) {
    // =========================== softmax warpgroups (A: even key blocks, B: odd) ===========================
    setmaxnreg_inc_208();
    const int wg = warp_idx >> 2;
    const int w = warp_idx & 3;
    const int r = w * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>(w * 32) << 16;
    const uint32_t t_o = tmem_base + lane_sel + AP_TMEM_O;
    const float c = p.scale_log2;
    const int tail = p.Skv - (n_kv - 1) * BN;  // valid keys of the last block, 1..BN
    float m_mine = -INFINITY;  // the reference maximum (raw score units) my row sum l is expressed against
    float l = 0.f;
    // barrier addresses the softmax warps arrive on (the issuer lives in the leader CTA)
    auto pready_bar = [&](int buf, int half) -> uint32_t {
      const uint32_t a = bar(Cfg::BAR_PREADY + 2 * buf + half);
      return (CG == 2) ? mapa_shared(a, 0) : a;
    };
    auto arrive_issuer = [&](uint32_t a) {
      if constexpr (CG == 2) mbar_arrive_cluster(a); else mbar_arrive(a);
    };

    int buf = wg;       // j % NBUF
    int sphase = 0;     // (j / NBUF) & 1
#pragma unroll 1
    for (int j = wg; j < n_kv; j += 2) {
      const uint32_t t_s = tmem_base + lane_sel + buf * BN;
      mbar_wait(bar(Cfg::BAR_SFULL + buf), sphase);
      tc_fence_after();
      if ((threadIdx.x & 127) == 0) ev(wg, 1);
      uint32_t s[BN];
#pragma unroll
      for (int q = 0; q < BN / 32; ++q) tmem_ld_x32(t_s + q * 32, reinterpret_cast<uint32_t(&)[32]>(s[q * 32]));
      // while the score tile is on its way from tensor memory: take the reference maximum published by the owner of
      // block j-1 (the other warpgroup, which passed this point most of a block ago)
      float m_prev = -INFINITY;
      if (j > 0) {
        mbar_wait(bar(Cfg::BAR_TOKEN + (wg ^ 1) * 4 + w), ((j - 1) >> 1) & 1);
        m_prev = mref[r];
        if (m_prev != m_mine) {  // the other warpgroup advanced m: rebase my private row sum
          if (l != 0.f) l *= fast_exp2((m_mine - m_prev) * c);
          m_mine = m_prev;
        }
      }
      tmem_wait_ld();
      if ((threadIdx.x & 127) == 0) ev(wg, 2);
      if (j == n_kv - 1 && tail < BN) {
#pragma unroll
        for (int i = 0; i < BN; ++i)
          if (i >= tail) s[i] = 0xff800000u;  // -inf
      }
      // block maximum: 8 independent chains (the 3-input FMNMX has a long dependent-issue latency)
      float mx[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) mx[k] = fmaxf(__uint_as_float(s[k]), __uint_as_float(s[k + 8]));
#pragma unroll
      for (int i = 16; i < BN; i += 16) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          mx[k] = fmaxf(mx[k], fmaxf(__uint_as_float(s[i + k]), __uint_as_float(s[i + k + 8])));
      }
      const float mblk = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])),
                               fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])));
      // ---- serial section: decide the reference maximum of this block and publish it ----
      if (j == 0) {
        m_mine = mblk;  // nothing accumulated yet
      } else {
        const float m_new = fmaxf(m_prev, mblk);
        const bool need = (m_new - m_prev) * c > AP_RESCALE_THRESHOLD;
        if (__any_sync(0xffffffffu, need)) {
          // O must be quiescent: P.V(j-1) has completed, and no later P.V can be issued before I release P(j)
          mbar_wait(bar(Cfg::BAR_PVDONE + ((j - 1) & 1)), ((j - 1) >> 1) & 1);
          tc_fence_after();
          const float f = fast_exp2((m_prev - m_new) * c);
          m_mine = m_new;
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
      mref[r] = m_mine;
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(Cfg::BAR_TOKEN + wg * 4 + w));
      if ((threadIdx.x & 127) == 0) ev(wg, 3);

      // ---- exponentials: P = 2^(s*c - m*c) -> bf16 over the first 64 columns of the score buffer; the private row
      //      sum rides along on the FMA pipe (the phase is bound by the 16-lane/clk exponential unit) ----
      const float neg = -m_mine * c;
      const float2 c2 = make_float2(c, c);
      const float2 neg2 = make_float2(neg, neg);
      float2 acc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
      for (int q = 0; q < BN / 32; ++q) {
        if (q == 2) {
          // first part of P (keys 0..63) is complete: the issuer may start P.V on it
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_issuer(pready_bar(buf, 0));
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
          acc[e & 3] = __fadd2_rn(acc[e & 3], pv);
          pk[e] = pack_bf16x2(pv.x, pv.y);
        }
        tmem_st_x16(t_s + q * 16, pk);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_issuer(pready_bar(buf, 1));
      if ((threadIdx.x & 127) == 0) ev(wg, 4);
      acc[0] = __fadd2_rn(__fadd2_rn(acc[0], acc[1]), __fadd2_rn(acc[2], acc[3]));
      l += acc[0].x + acc[0].y;
      // next block of this warpgroup: j + 2
      buf += 2;
      if (buf >= NBUF) { buf -= NBUF; sphase ^= 1; }
    }

    // ---- both warpgroups: rebase to the final reference maximum, add the row sums ----
    const int last = n_kv - 1;
    if (wg != (last & 1)) {
      mbar_wait(bar(Cfg::BAR_TOKEN + (last & 1) * 4 + w), (last >> 1) & 1);
      const float m_fin = mref[r];
      if (m_fin != m_mine) {
        if (l != 0.f) l *= fast_exp2((m_mine - m_fin) * c);
        m_mine = m_fin;
      }
    }
    lsum[wg * 128 + r] = l;
    named_bar_sync(1, 256);
    const float l_tot = lsum[r] + lsum[128 + r];

    // ---- epilogue: warpgroup wg takes output columns [64 wg, 64 wg + 64): O / l -> bf16 -> swizzled smem -> TMA ----
    mbar_wait(bar(Cfg::BAR_PVDONE + (last & 1)), (last >> 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / l_tot;
    const uint32_t stage = smem_base + Cfg::OFF_Q + wg * AP_PANEL_BYTES;  // the Q tile is dead: all Q.K^T completed
#pragma unroll 1
    for (int q = 0; q < 2; ++q) {
      uint32_t o[32];
      tmem_ld_x32(t_o + wg * 64 + q * 32, o);
      tmem_wait_ld();
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        uint32_t wv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          wv[e] = pack_bf16x2(__uint_as_float(o[v * 8 + 2 * e]) * inv, __uint_as_float(o[v * 8 + 2 * e + 1]) * inv);
        const int chunk16 = q * 4 + v;
        const uint32_t dst = stage + r * 128 + ((chunk16 ^ (r & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(wv[0]), "r"(wv[1]), "r"(wv[2]),
                     "r"(wv[3])
                     : "memory");
      }
    }
    fence_proxy_async_smem();
    named_bar_sync(2 + wg, 128);
    if ((threadIdx.x & 127) == 0) {
      tma_store_3d(&tmO, stage, h * 128 + wg * 64, row0, b);
      tma_store_commit();
      tma_store_wait<0>();
    }
    if (wg == 0 && p.lse != nullptr && row0 + r < p.Sq)
      p.lse[(static_cast<long long>(b) * p.H + h) * p.Sq + row0 + r] = m_mine * p.scale + __logf(l_tot);
  } else {
    setmaxnreg_dec_88();
    if (warp_idx == 8) {
      // =========================== TMA producer (every CTA loads its own parts) ===========================
      if (elect_one()) {
        const int c0 = h * 128;
        const uint32_t q_full = (CG == 2) ? mapa_shared(bar(Cfg::BAR_QFULL), 0) : bar(Cfg::BAR_QFULL);
        if (is_leader) mbar_arrive_expect_tx(bar(Cfg::BAR_QFULL), AP_TILE_BYTES * CG);
        if constexpr (CG == 1) {
          tma_load_3d(smem_base + Cfg::OFF_Q, &tmQ, q_full, c0, row0, b);
          tma_load_3d(smem_base + Cfg::OFF_Q + AP_PANEL_BYTES, &tmQ, q_full, c0 + 64, row0, b);
        } else {
          tma_load_3d_pair(smem_base + Cfg::OFF_Q, &tmQ, q_full, c0, row0, b);
          tma_load_3d_pair(smem_base + Cfg::OFF_Q + AP_PANEL_BYTES, &tmQ, q_full, c0 + 64, row0, b);
        }
        uint32_t slot = 0, phase = 0;
        auto load_part = [&](bool is_v, int blk) {
          mbar_wait(bar(Cfg::BAR_KVEMPTY + slot), phase ^ 1);
          const uint32_t dst = smem_base + Cfg::OFF_KV + slot * Cfg::SLOT_BYTES;
          const uint32_t full_local = bar(Cfg::BAR_KVFULL + slot);
          if constexpr (CG == 1) {
            mbar_arrive_expect_tx(full_local, Cfg::SLOT_BYTES);
            const CUtensorMap* m = is_v ? &tmV : &tmK;
            tma_load_3d(dst, m, full_local, c0, blk * BN, b);
            tma_load_3d(dst + Cfg::V_PANEL_BYTES, m, full_local, c0 + 64, blk * BN, b);
          } else {
            const uint32_t full = mapa_shared(full_local, 0);
            if (is_leader) mbar_arrive_expect_tx(full_local, 2 * Cfg::SLOT_BYTES);  // both CTAs' parts
            if (is_v) {
              // V part: all BN keys, head-dim columns [64 rank, 64 rank + 64)  (B operand split along N = d)
              tma_load_3d_pair(dst, &tmV, full, c0 + 64 * static_cast<int>(rank), blk * BN, b);
            } else {
              // K part: keys [BN/2 rank, BN/2 rank + BN/2) of the block, all 128 head-dim columns (B split along N = keys)
              const int krow = blk * BN + (BN / 2) * static_cast<int>(rank);
              tma_load_3d_pair(dst, &tmK, full, c0, krow, b);
              tma_load_3d_pair(dst + Cfg::K_PANEL_BYTES, &tmK, full, c0 + 64, krow, b);
            }
          }
          if (++slot == Cfg::NS) { slot = 0; phase ^= 1; }
        };
        const int n_pro = n_kv < NBUF ? n_kv : NBUF;
        for (int i = 0; i < n_pro; ++i) load_part(false, i);
        for (int j = 0; j < n_kv; ++j) {
          load_part(true, j);
          if (j + NBUF < n_kv) load_part(false, j + NBUF);
        }
      }
    } else if (warp_idx == 9) {
      // =========================== tcgen05 issuer (leader CTA of the pair) ===========================
      // The whole warp walks the schedule (addresses and descriptors live in uniform registers); one elected lane
      // issues the MMAs and commits.
      if (is_leader) {
        constexpr uint32_t IDESC_QK = umma_idesc_bf16(128 * CG, BN, 0, 0);
        constexpr uint32_t IDESC_PV = umma_idesc_bf16(128 * CG, 128, 0, 1);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        auto commit = [&](uint32_t b_) {
          if (elect_one()) {
            if constexpr (CG == 1) umma_commit(b_); else umma_commit_pair(b_, 0x3);
          }
          __syncwarp();
        };
        auto issue_qk = [&](int buf, uint32_t kbase) {
          const uint64_t qd = umma_desc_k_sw128(smem_base + Cfg::OFF_Q);
          const uint64_t kd = umma_desc_k_sw128(kbase);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint32_t offq = ((ks >> 2) * AP_PANEL_BYTES + (ks & 3) * 32) >> 4;
              const uint32_t offk = ((ks >> 2) * Cfg::K_PANEL_BYTES + (ks & 3) * 32) >> 4;
              umma_ss<CG>(tmem_u + buf * BN, qd + offq, kd + offk, IDESC_QK, ks > 0 ? 1u : 0u);
            }
          }
          __syncwarp();
        };
        // P.V over part 0 (keys 0..63: K steps 0-3) or part 1 (keys 64..BN-1) of the block
        auto issue_pv_half = [&](int buf, uint32_t vbase, int half, bool acc) {
          const uint64_t vd = umma_desc_mn_sw128(vbase + half * 8192, Cfg::V_PANEL_BYTES, 1024);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              if (half * 4 + ks < Cfg::PV_KSTEPS)
                umma_ts_cg<CG>(tmem_u + AP_TMEM_O, tmem_u + buf * BN + half * 32 + ks * 8, vd + ((ks * 2048) >> 4),
                               IDESC_PV, (acc || half > 0 || ks > 0) ? 1u : 0u);
            }
          }
          __syncwarp();
        };
        uint32_t slot = 0, phase = 0;
        auto slot_addr = [&](uint32_t s_) { return smem_base + Cfg::OFF_KV + s_ * Cfg::SLOT_BYTES; };
        auto next_slot = [&]() { if (++slot == Cfg::NS) { slot = 0; phase ^= 1; } };

        mbar_wait(bar(Cfg::BAR_QFULL), 0);
        const int n_pro = n_kv < NBUF ? n_kv : NBUF;
        for (int i = 0; i < n_pro; ++i) {
          mbar_wait(bar(Cfg::BAR_KVFULL + slot), phase);
          tc_fence_after();
          issue_qk(i, slot_addr(slot));
          commit(bar(Cfg::BAR_SFULL + i));
          commit(bar(Cfg::BAR_KVEMPTY + slot));
          next_slot();
        }
        int buf = 0, pphase = 0;  // j % NBUF, (j / NBUF) & 1
        for (int j = 0; j < n_kv; ++j) {
          const uint32_t vslot = slot;
          mbar_wait(bar(Cfg::BAR_KVFULL + vslot), phase);
          next_slot();
          mbar_wait(bar(Cfg::BAR_PREADY + 2 * buf), pphase);
          tc_fence_after();
          ev(2, 10 + (j & 1));
          issue_pv_half(buf, slot_addr(vslot), 0, j > 0);
          mbar_wait(bar(Cfg::BAR_PREADY + 2 * buf + 1), pphase);
          tc_fence_after();
          issue_pv_half(buf, slot_addr(vslot), 1, true);
          commit(bar(Cfg::BAR_KVEMPTY + vslot));
          commit(bar(Cfg::BAR_PVDONE + (j & 1)));
          ev(2, 12 + (j & 1));
          if (j + NBUF < n_kv) {
            // the score buffer P(j) lived in is free once P.V(j) is queued (the tensor pipe runs in issue order)
            mbar_wait(bar(Cfg::BAR_KVFULL + slot), phase);
            tc_fence_after();
            issue_qk(buf, slot_addr(slot));
            commit(bar(Cfg::BAR_SFULL + buf));
            commit(bar(Cfg::BAR_KVEMPTY + slot));
            next_slot();
            ev(2, 14 + (j & 1));
          }
          if (++buf == NBUF) { buf = 0; pphase ^= 1; }
        }
      }
    }
  }

  // =========================== teardown ===========================
  __syncwarp();
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp_idx == 9) tmem_dealloc<CG>(tmem_base, 512);
}

template <int CG, int BN, int EMU, bool TRACE>
static int launch_attn_pair_t(int B, int Sq, int H, cudaStream_t stream, const CUtensorMap& tmQ, const CUtensorMap& tmK,
                              const CUtensorMap& tmV, const CUtensorMap& tmO, const AttnParams& p) {
  using Cfg = APCfg<CG, BN>;
  auto kernel = attn_pair_kernel<CG, BN, EMU, TRACE>;
  static std::atomic<bool> configured[64];  // zero-initialised; per-device "attribute set" latch, safe across host threads
  int dev = 0;
  MV_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
    MV_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
  }
  const int tiles = (Sq + 127) / 128;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(((tiles + CG - 1) / CG) * CG, H, B);
  cfg.blockDim = dim3(AP_THREADS, 1, 1);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = CG;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  MV_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, tmQ, tmK, tmV, tmO, p));
  return 0;
}

// cg: 1 = single CTA per query tile, 2 = CTA pair;  bn: keys per block (128: 3 score buffers, 96: 4);
// emu: polynomial-exp2 share in 16ths of the score pairs
int launch_attn_pair(int cg, int bn, int emu, bool trace, int B, int Sq, int H, cudaStream_t stream,
                     const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const CUtensorMap& tmO,
                     const AttnParams& p) {
#define MV_AP(CGV, BNV, EMUV)                                                                              \
  if (cg == CGV && bn == BNV && emu == EMUV)                                                               \
    return trace ? launch_attn_pair_t<CGV, BNV, 4, true>(B, Sq, H, stream, tmQ, tmK, tmV, tmO, p)          \
                 : launch_attn_pair_t<CGV, BNV, EMUV, false>(B, Sq, H, stream, tmQ, tmK, tmV, tmO, p)
  // Instantiated: what ships.  Measured and dropped (profiles/r02_attn_schedules_call*.log, r02_attn_bn96_vs_bn128.log;
  // 43120 x 43120 x 40 heads, sustained): CG = 1 1251-1273 TFLOP/s, BN = 96 / 4 score buffers 1259 (CG = 2) and 1152
  // (CG = 1), exp2 emulation share 0 / 6 / 8 of 16: 1314 / 1224 / 1136, against 1325-1340 for CG = 2, BN = 128, share 4.
  MV_AP(2, 128, 4);
#undef MV_AP
  set_error("launch_attn_pair: unsupported cta_group %d / key block %d / exp2 emulation share %d", cg, bn, emu);
  return -1;
}

}  // namespace mv
