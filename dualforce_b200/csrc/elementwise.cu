// Memory-bound kernels of the DiT block: LayerNorm (+affine, +AdaLN modulate), full-width RMSNorm + RoPE,
// modulation vector add, LSE merge.  All are one-pass (each activation row is read once into registers,
// written once), 16-byte vectorised and warp-shuffle reduced; bf16 I/O, fp32 math.
//
// Reference call sites: mova/diffusion/models/wan_video_dit.py:94-96 (modulate), :131-137 (interleaved RoPE),
// :175-176,:181-187 (RMSNorm(dim) on q,k), :267-269,:286,:289 (LayerNorms), :279-280 (modulation + t_mod);
// mova/diffusion/models/interactionv2.py:40-72 (rotate-half RoPE), :222-223,:229-249, :322,:349 (y_norm).
#include "common.cuh"
#include "host_utils.h"
#include "../../include/mova_b200.h"

namespace mv {

constexpr int EW_THREADS = 128;
constexpr int EW_MAXV = 8;  // 16-byte vectors cached per thread -> rows up to 128*8*8 = 8192 channels

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sum over the whole block (EW_THREADS = 4 warps); every thread gets the result
__device__ __forceinline__ float block_sum(float v, float* red /* [4] */) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5;
  __syncthreads();  // protect `red` against the previous reduction's readers
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  return red[0] + red[1] + red[2] + red[3];
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
  f[4] = bf16lo(u.z); f[5] = bf16hi(u.z); f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

// ------------------------------------------------------------------------------------------------
// y = LN(x) [*w + b] [*(1+scale) + shift]
// ------------------------------------------------------------------------------------------------
template <bool AFFINE, bool MODULATE>
__global__ void __launch_bounds__(EW_THREADS)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ y, long long ldy,
                 int d, float eps, const __nv_bfloat16* __restrict__ ln_w, const __nv_bfloat16* __restrict__ ln_b,
                 const float* __restrict__ shift, const float* __restrict__ scale) {
  __shared__ float red[4];
  const long long row = blockIdx.x;
  const int nvec = d >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * ldx);
  uint4 cache[EW_MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < EW_MAXV; ++i) {
    const int v = threadIdx.x + i * EW_THREADS;
    if (v < nvec) cache[i] = xr[v];
  }
#pragma unroll
  for (int i = 0; i < EW_MAXV; ++i) {
    const int v = threadIdx.x + i * EW_THREADS;
    if (v < nvec) {
      float f[8];
      unpack8(cache[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) s += f[e];
    }
  }
  const float mean = block_sum(s, red) / static_cast<float>(d);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < EW_MAXV; ++i) {
    const int v = threadIdx.x + i * EW_THREADS;
    if (v < nvec) {
      float f[8];
      unpack8(cache[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float c = f[e] - mean;
        ss += c * c;
      }
    }
  }
  const float rstd = rsqrtf(block_sum(ss, red) / static_cast<float>(d) + eps);

  uint4* yr = reinterpret_cast<uint4*>(y + row * ldy);
#pragma unroll
  for (int i = 0; i < EW_MAXV; ++i) {
    const int v = threadIdx.x + i * EW_THREADS;
    if (v < nvec) {
      float f[8];
      unpack8(cache[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = (f[e] - mean) * rstd;
      if constexpr (AFFINE) {
        float w[8], b[8];
        unpack8(reinterpret_cast<const uint4*>(ln_w)[v], w);
        unpack8(reinterpret_cast<const uint4*>(ln_b)[v], b);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = fmaf(f[e], w[e], b[e]);
      }
      if constexpr (MODULATE) {
        const float4 sc0 = reinterpret_cast<const float4*>(scale)[2 * v];
        const float4 sc1 = reinterpret_cast<const float4*>(scale)[2 * v + 1];
        const float4 sh0 = reinterpret_cast<const float4*>(shift)[2 * v];
        const float4 sh1 = reinterpret_cast<const float4*>(shift)[2 * v + 1];
        const float sc[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
        const float sh[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = fmaf(f[e], 1.0f + sc[e], sh[e]);
      }
      yr[v] = pack8(f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// in-place x = RoPE(RMSNorm_full(x) * w)
// Threads own PAIRS of vectors (hv, hv + 8) of one 128-wide head so both RoPE conventions are thread-local.
// ------------------------------------------------------------------------------------------------
constexpr int RR_MAXP = 4;  // vector pairs per thread -> d up to 128 * 4 * 16 = 8192

template <int ROPE>
__global__ void __launch_bounds__(EW_THREADS)
rmsnorm_rope_kernel(__nv_bfloat16* __restrict__ x, long long ldx, int d, int seg_vecs, long long seg_stride,
                    const __nv_bfloat16* __restrict__ w, float eps, const float* __restrict__ cos_tab,
                    const float* __restrict__ sin_tab) {
  __shared__ float red[4];
  const long long row = blockIdx.x;
  const int npair = d >> 4;
  // logical 16-byte vector v of the row lives in segment v / seg_vecs (segments are seg_stride elements apart)
  __nv_bfloat16* xrow = x + row * ldx;
  auto vec_ptr = [&](int v) -> uint4* {
    const int seg = v / seg_vecs;
    return reinterpret_cast<uint4*>(xrow + seg * seg_stride) + (v - seg * seg_vecs);
  };
  uint4 lo[RR_MAXP], hi[RR_MAXP];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < RR_MAXP; ++i) {
    const int pi = threadIdx.x + i * EW_THREADS;
    if (pi < npair) {
      const int v = (pi >> 3) * 16 + (pi & 7);
      lo[i] = *vec_ptr(v);
      hi[i] = *vec_ptr(v + 8);
    }
  }
#pragma unroll
  for (int i = 0; i < RR_MAXP; ++i) {
    const int pi = threadIdx.x + i * EW_THREADS;
    if (pi < npair) {
      float a[8], b[8];
      unpack8(lo[i], a);
      unpack8(hi[i], b);
#pragma unroll
      for (int e = 0; e < 8; ++e) ss += a[e] * a[e] + b[e] * b[e];
    }
  }
  const float inv = rsqrtf(block_sum(ss, red) / static_cast<float>(d) + eps);

#pragma unroll
  for (int i = 0; i < RR_MAXP; ++i) {
    const int pi = threadIdx.x + i * EW_THREADS;
    if (pi < npair) {
      const int hv = pi & 7;
      const int v = (pi >> 3) * 16 + hv;
      float a[8], b[8], wa[8], wb[8];
      unpack8(lo[i], a);
      unpack8(hi[i], b);
      unpack8(reinterpret_cast<const uint4*>(w)[v], wa);
      unpack8(reinterpret_cast<const uint4*>(w)[v + 8], wb);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        a[e] = a[e] * inv * wa[e];
        b[e] = b[e] * inv * wb[e];
      }
      if constexpr (ROPE == MOVA_ROPE_INTERLEAVED) {
        // table row: 64 (cos|sin) values, one per complex pair (2i, 2i+1) of the head
        const float4 ca = reinterpret_cast<const float4*>(cos_tab + row * 64)[hv];
        const float4 sa = reinterpret_cast<const float4*>(sin_tab + row * 64)[hv];
        const float4 cb = reinterpret_cast<const float4*>(cos_tab + row * 64)[hv + 8];
        const float4 sb = reinterpret_cast<const float4*>(sin_tab + row * 64)[hv + 8];
        const float cA[4] = {ca.x, ca.y, ca.z, ca.w}, sA[4] = {sa.x, sa.y, sa.z, sa.w};
        const float cB[4] = {cb.x, cb.y, cb.z, cb.w}, sB[4] = {sb.x, sb.y, sb.z, sb.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float ar = a[2 * e], ai = a[2 * e + 1];
          a[2 * e] = ar * cA[e] - ai * sA[e];
          a[2 * e + 1] = ar * sA[e] + ai * cA[e];
          const float br = b[2 * e], bi = b[2 * e + 1];
          b[2 * e] = br * cB[e] - bi * sB[e];
          b[2 * e + 1] = br * sB[e] + bi * cB[e];
        }
      } else if constexpr (ROPE == MOVA_ROPE_HALF) {
        // table row: 128 values; out = x * cos + rotate_half(x) * sin, rotate_half(x) = cat(-x2, x1)
        const float* cr = cos_tab + row * 128;
        const float* sr = sin_tab + row * 128;
        float c1[8], s1[8], c2[8], s2[8];
        *reinterpret_cast<float4*>(c1) = reinterpret_cast<const float4*>(cr)[2 * hv];
        *reinterpret_cast<float4*>(c1 + 4) = reinterpret_cast<const float4*>(cr)[2 * hv + 1];
        *reinterpret_cast<float4*>(s1) = reinterpret_cast<const float4*>(sr)[2 * hv];
        *reinterpret_cast<float4*>(s1 + 4) = reinterpret_cast<const float4*>(sr)[2 * hv + 1];
        *reinterpret_cast<float4*>(c2) = reinterpret_cast<const float4*>(cr)[2 * (hv + 8)];
        *reinterpret_cast<float4*>(c2 + 4) = reinterpret_cast<const float4*>(cr)[2 * (hv + 8) + 1];
        *reinterpret_cast<float4*>(s2) = reinterpret_cast<const float4*>(sr)[2 * (hv + 8)];
        *reinterpret_cast<float4*>(s2 + 4) = reinterpret_cast<const float4*>(sr)[2 * (hv + 8) + 1];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float x1 = a[e], x2 = b[e];
          a[e] = x1 * c1[e] - x2 * s1[e];
          b[e] = x2 * c2[e] + x1 * s2[e];
        }
      }
      *vec_ptr(v) = pack8(a);
      *vec_ptr(v + 8) = pack8(b);
    }
  }
}

__global__ void add_to_f32_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                  float* __restrict__ out, long long n) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) out[i] = __bfloat162float(a[i]) + (b != nullptr ? __bfloat162float(b[i]) : 0.0f);
}

// one thread per (row, head, 8-channel vector)
__global__ void lse_merge_kernel(const __nv_bfloat16* __restrict__ o_parts, const float* __restrict__ lse_parts,
                                 int n_parts, __nv_bfloat16* __restrict__ out, long long ldo,
                                 float* __restrict__ lse_out, int rows, int H, int D) {
  const int vec_per_head = D >> 3;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(rows) * H * vec_per_head;
  if (idx >= total) return;
  const int vh = static_cast<int>(idx % vec_per_head);
  const int h = static_cast<int>((idx / vec_per_head) % H);
  const int r = static_cast<int>(idx / (static_cast<long long>(vec_per_head) * H));
  float m = -INFINITY;
  for (int p = 0; p < n_parts; ++p) m = fmaxf(m, lse_parts[(static_cast<long long>(p) * H + h) * rows + r]);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float denom = 0.f;
  for (int p = 0; p < n_parts; ++p) {
    const float l = lse_parts[(static_cast<long long>(p) * H + h) * rows + r];
    const float wgt = (m == -INFINITY) ? 0.f : __expf(l - m);
    denom += wgt;
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(o_parts + ((static_cast<long long>(p) * rows + r) * H + h) * D + vh * 8), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = fmaf(wgt, f[e], acc[e]);
  }
  const float invd = denom > 0.f ? 1.0f / denom : 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] *= invd;
  *reinterpret_cast<uint4*>(out + static_cast<long long>(r) * ldo + h * D + vh * 8) = pack8(acc);
  if (lse_out != nullptr && vh == 0) lse_out[static_cast<long long>(h) * rows + r] = m + __logf(denom);
}

}  // namespace mv

extern "C" {

int mova_b200_layernorm(const void* x, int64_t ldx, void* y, int64_t ldy, int L, int d, float eps, const void* ln_w,
                        const void* ln_b, const float* shift, const float* scale, void* stream) {
  using namespace mv;
  MV_REQUIRE(x && y, "mova_b200_layernorm: null pointer");
  MV_REQUIRE(d > 0 && d % 8 == 0 && d <= EW_THREADS * EW_MAXV * 8, "mova_b200_layernorm: d=%d unsupported", d);
  MV_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0 && ldx >= d && ldy >= d, "mova_b200_layernorm: bad leading dimension");
  MV_REQUIRE((ln_w == nullptr) == (ln_b == nullptr), "mova_b200_layernorm: ln_w and ln_b must come together");
  MV_REQUIRE((shift == nullptr) == (scale == nullptr), "mova_b200_layernorm: shift and scale must come together");
  if (L <= 0) return 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const auto* xp = static_cast<const __nv_bfloat16*>(x);
  auto* yp = static_cast<__nv_bfloat16*>(y);
  const auto* wp = static_cast<const __nv_bfloat16*>(ln_w);
  const auto* bp = static_cast<const __nv_bfloat16*>(ln_b);
  const bool aff = ln_w != nullptr, mod = shift != nullptr;
  if (aff && mod) layernorm_kernel<true, true><<<L, EW_THREADS, 0, s>>>(xp, ldx, yp, ldy, d, eps, wp, bp, shift, scale);
  else if (aff) layernorm_kernel<true, false><<<L, EW_THREADS, 0, s>>>(xp, ldx, yp, ldy, d, eps, wp, bp, shift, scale);
  else if (mod) layernorm_kernel<false, true><<<L, EW_THREADS, 0, s>>>(xp, ldx, yp, ldy, d, eps, wp, bp, shift, scale);
  else layernorm_kernel<false, false><<<L, EW_THREADS, 0, s>>>(xp, ldx, yp, ldy, d, eps, wp, bp, shift, scale);
  MV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mova_b200_rmsnorm_rope(void* x, int64_t ldx, int L, int d, int head_dim, const void* w, float eps,
                           const float* cos_tab, const float* sin_tab, int rope_mode, void* stream) {
  return mova_b200_rmsnorm_rope_seg(x, ldx, d, 0, L, d, head_dim, w, eps, cos_tab, sin_tab, rope_mode, stream);
}

int mova_b200_rmsnorm_rope_seg(void* x, int64_t ldx, int seg_len, int64_t seg_stride, int L, int d, int head_dim,
                               const void* w, float eps, const float* cos_tab, const float* sin_tab, int rope_mode,
                               void* stream) {
  using namespace mv;
  MV_REQUIRE(x && w, "mova_b200_rmsnorm_rope: null pointer");
  MV_REQUIRE(d > 0 && d % 128 == 0 && d <= EW_THREADS * RR_MAXP * 16, "mova_b200_rmsnorm_rope: d=%d unsupported", d);
  MV_REQUIRE(seg_len > 0 && seg_len % 128 == 0 && d % seg_len == 0,
             "mova_b200_rmsnorm_rope: seg_len (%d) must be a multiple of 128 dividing d (%d)", seg_len, d);
  MV_REQUIRE(ldx % 8 == 0 && ldx >= seg_len && seg_stride % 8 == 0, "mova_b200_rmsnorm_rope: bad leading dimension");
  const int seg_vecs = seg_len / 8;
  MV_REQUIRE(rope_mode >= MOVA_ROPE_NONE && rope_mode <= MOVA_ROPE_HALF, "mova_b200_rmsnorm_rope: bad rope_mode %d",
             rope_mode);
  if (rope_mode != MOVA_ROPE_NONE) {
    MV_REQUIRE(head_dim == 128, "mova_b200_rmsnorm_rope: RoPE needs head_dim 128 (got %d)", head_dim);
    MV_REQUIRE(cos_tab && sin_tab, "mova_b200_rmsnorm_rope: RoPE tables missing");
  }
  if (L <= 0) return 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto* xp = static_cast<__nv_bfloat16*>(x);
  const auto* wp = static_cast<const __nv_bfloat16*>(w);
  if (rope_mode == MOVA_ROPE_NONE)
    rmsnorm_rope_kernel<MOVA_ROPE_NONE><<<L, EW_THREADS, 0, s>>>(xp, ldx, d, seg_vecs, seg_stride, wp, eps, cos_tab, sin_tab);
  else if (rope_mode == MOVA_ROPE_INTERLEAVED)
    rmsnorm_rope_kernel<MOVA_ROPE_INTERLEAVED><<<L, EW_THREADS, 0, s>>>(xp, ldx, d, seg_vecs, seg_stride, wp, eps, cos_tab, sin_tab);
  else
    rmsnorm_rope_kernel<MOVA_ROPE_HALF><<<L, EW_THREADS, 0, s>>>(xp, ldx, d, seg_vecs, seg_stride, wp, eps, cos_tab, sin_tab);
  MV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mova_b200_add_to_f32(const void* a, const void* b, float* out, int64_t n, void* stream) {
  using namespace mv;
  MV_REQUIRE(a && out, "mova_b200_add_to_f32: null pointer");
  if (n <= 0) return 0;
  const int threads = 256;
  const long long blocks = (n + threads - 1) / threads;
  add_to_f32_kernel<<<static_cast<unsigned>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b), out, n);
  MV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mova_b200_lse_merge(const void* o_parts, const float* lse_parts, int n_parts, void* out, int64_t ldo,
                        float* lse_out, int rows, int H, int D, void* stream) {
  using namespace mv;
  MV_REQUIRE(o_parts && lse_parts && out, "mova_b200_lse_merge: null pointer");
  MV_REQUIRE(n_parts >= 1 && D % 8 == 0 && ldo % 8 == 0 && ldo >= static_cast<int64_t>(H) * D,
             "mova_b200_lse_merge: bad arguments");
  if (rows <= 0) return 0;
  const long long total = static_cast<long long>(rows) * H * (D / 8);
  const int threads = 256;
  lse_merge_kernel<<<static_cast<unsigned>((total + threads - 1) / threads), threads, 0,
                     static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(o_parts), lse_parts,
                                                          n_parts, static_cast<__nv_bfloat16*>(out), ldo, lse_out,
                                                          rows, H, D);
  MV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
