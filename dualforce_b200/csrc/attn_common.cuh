// Shared between the attention kernels (attn_pair.cu: the round-2 schedule, attn.cu: the round-1 schedule + C ABI).
#pragma once

#include "common.cuh"
#include "host_utils.h"
#include "../../include/mova_b200.h"

namespace mv {

struct AttnParams {
  int Sq, Skv, H;
  float scale;       // softmax scale
  float scale_log2;  // scale * log2(e)
  float* lse;        // [B, H, Sq] or null
  unsigned long long* trace;  // diagnostics: 3 regions of 4096 (clock << 8 | event) records of CTA (0,0,0), or null
};

__device__ __forceinline__ void setmaxnreg_inc_208() { asm volatile("setmaxnreg.inc.sync.aligned.u32 208;"); }
__device__ __forceinline__ void setmaxnreg_dec_88() { asm volatile("setmaxnreg.dec.sync.aligned.u32 88;"); }

// 2^x for a pair of scores on the FMA/ALU pipes instead of the 16-lane/clk MUFU unit (which at head_dim 128 is as
// busy as the tensor cores): round x to the nearest integer n with the 1.5*2^23 trick, evaluate a degree-3 minimax
// polynomial of 2^r on r = x - n in [-0.5, 0.5] (max relative error 7.5e-5, far below the bf16 rounding of P) with
// packed f32x2 instructions, then add n to the exponent field.
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.0f);
  x.y = fmaxf(x.y, -126.0f);
  const float2 magic = make_float2(12582912.0f, 12582912.0f);
  const float2 t = __fadd2_rn(x, magic);
  const float2 n = __fadd2_rn(t, make_float2(-12582912.0f, -12582912.0f));
  const float2 r = __ffma2_rn(n, make_float2(-1.0f, -1.0f), x);
  float2 p = __ffma2_rn(r, make_float2(0.0551716648f, 0.0551716648f), make_float2(0.2426111251f, 0.2426111251f));
  p = __ffma2_rn(p, r, make_float2(0.6932609677f, 0.6932609677f));
  p = __ffma2_rn(p, r, make_float2(0.9999280572f, 0.9999280572f));
  p.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  p.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
  return p;
}


int launch_attn_pair(int cg, int bn, int emu, bool trace, int B, int Sq, int H, cudaStream_t stream,
                     const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const CUtensorMap& tmO,
                     const AttnParams& p);

}  // namespace mv
