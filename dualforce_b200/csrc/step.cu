// The small kernels either side of the dual-tower forward inside one denoising step
// (MOVA.inference_single_step, mova/diffusion/pipelines/pipeline_mova.py:500-609):
//
//   * patchify:    im2col of the stride == kernel Conv3d / Conv1d patch embedding
//                  (wan_video_dit.py:367-368,399-409; wan_audio_dit.py:143-145,180-189), so the convolution
//                  itself runs as one tcgen05 GEMM  [L, C*pt*ph*pw] x [C*pt*ph*pw, dim],
//   * unpatchify:  'b (f h w) (x y z c) -> b c (f x) (h y) (w z)' (wan_video_dit.py:411-416;
//                  wan_audio_dit.py:191-195 is the same map with y = z = 1),
//   * sinusoidal:  sinusoidal_embedding_1d in fp64 (wan_video_dit.py:99-103), timestep read from device memory
//                  (no host synchronisation),
//   * gemv_f32:    the M = 1 time_embedding / time_projection MLPs (wan_video_dit.py:374-380), which the
//                  reference runs under autocast(float32) (pipeline_mova.py:544-549): fp32 activations,
//                  bf16-valued weights, fp32 accumulation, SiLU fused on either side.
//
//   * cfg_euler:   classifier-free-guidance combine + flow-match Euler update of the latents in one pass
//                  (pipeline_mova.py:456-460; schedulers/flow_match_pair.py:213-227).
//
// All of them are HBM-bound byte movers: coalesced 16-byte accesses on the large side, one pass.
#include "common.cuh"
#include "host_utils.h"
#include "../../include/mova_b200.h"

namespace mv {

// ------------------------------------------------------------------------------------------------
// patchify: out[l, k] = bf16(x[c, f*pt+dt, h*ph+dh, w*pw+dw]),  l = (f*Hp + h)*Wp + w,
//           k = ((c*pt + dt)*ph + dh)*pw + dw   (== Conv3d weight.view(dim, -1) column order)
// One thread per output element: consecutive threads write consecutive k (coalesced), and read runs of
// pw consecutive input elements; the whole input (12-25 MB at 360p) is L2 resident after the first touch.
// ------------------------------------------------------------------------------------------------
template <typename TIn>
__global__ void __launch_bounds__(256)
patchify_kernel(const TIn* __restrict__ x, __nv_bfloat16* __restrict__ out, long long ldo, int C, int F, int H,
                int W, int pt, int ph, int pw, long long total) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int K = C * pt * ph * pw;
  const long long l = idx / K;
  int k = static_cast<int>(idx - l * K);
  const int Hp = H / ph, Wp = W / pw;
  const int w = static_cast<int>(l % Wp);
  const int h = static_cast<int>((l / Wp) % Hp);
  const int f = static_cast<int>(l / (static_cast<long long>(Wp) * Hp));
  const int dw = k % pw; k /= pw;
  const int dh = k % ph; k /= ph;
  const int dt = k % pt;
  const int c = k / pt;
  const long long src = ((static_cast<long long>(c) * F + (f * pt + dt)) * H + (h * ph + dh)) * W + (w * pw + dw);
  out[l * ldo + (idx - l * K)] = __float2bfloat16(static_cast<float>(x[src]));
}

// ------------------------------------------------------------------------------------------------
// unpatchify: out[c, f*pt+x, h*ph+y, w*pw+z] = in[l, ((x*ph + y)*pw + z)*Cout + c]
// One thread per output element (coalesced writes along the innermost output axis).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
unpatchify_kernel(const __nv_bfloat16* __restrict__ in, long long ldi, __nv_bfloat16* __restrict__ out, int Cout,
                  int Fp, int Hp, int Wp, int pt, int ph, int pw, long long total) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int Wo = Wp * pw, Ho = Hp * ph, Fo = Fp * pt;
  const int wo = static_cast<int>(idx % Wo);
  const int ho = static_cast<int>((idx / Wo) % Ho);
  const int fo = static_cast<int>((idx / (static_cast<long long>(Wo) * Ho)) % Fo);
  const int c = static_cast<int>(idx / (static_cast<long long>(Wo) * Ho * Fo));
  const int w = wo / pw, z = wo - w * pw;
  const int h = ho / ph, y = ho - h * ph;
  const int f = fo / pt, xx = fo - f * pt;
  const long long l = (static_cast<long long>(f) * Hp + h) * Wp + w;
  const int col = ((xx * ph + y) * pw + z) * Cout + c;
  out[idx] = in[l * ldi + col];
}

// ------------------------------------------------------------------------------------------------
// sinusoidal_embedding_1d: out[i] = cos(t * 10000^(-i/half)), out[half + i] = sin(...), fp64 math -> fp32
// ------------------------------------------------------------------------------------------------
__global__ void sinusoidal_kernel(const float* __restrict__ t, float* __restrict__ out, int half) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= half) return;
  const double pos = static_cast<double>(t[0]);
  const double ang = pos * pow(10000.0, -static_cast<double>(i) / static_cast<double>(half));
  out[i] = static_cast<float>(cos(ang));
  out[half + i] = static_cast<float>(sin(ang));
}

// ------------------------------------------------------------------------------------------------
// y[n] = post(sum_k pre(x[k]) * W[n, k] + b[n]);  x, y fp32; W, b bf16.  One warp per output row, the activation
// vector staged once per block in shared memory (pre-activation applied there), weight rows streamed with
// 16-byte loads -- the kernel is bound by the single pass over W.
// ------------------------------------------------------------------------------------------------
constexpr int GEMV_THREADS = 256;
constexpr int GEMV_ROWS_PER_WARP = 4;

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + __expf(-v)); }

__global__ void __launch_bounds__(GEMV_THREADS)
gemv_f32_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ W, long long ldw,
                const __nv_bfloat16* __restrict__ bias, float* __restrict__ y, __nv_bfloat16* __restrict__ y_bf16,
                int N, int K, int pre_act, int post_act) {
  extern __shared__ float xs[];
  for (int k = threadIdx.x; k < K; k += GEMV_THREADS) {
    const float v = x[k];
    xs[k] = pre_act ? silu_f(v) : v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = (blockIdx.x * (GEMV_THREADS / 32) + warp) * GEMV_ROWS_PER_WARP;
  const int nvec = K >> 3;
#pragma unroll 1
  for (int r = 0; r < GEMV_ROWS_PER_WARP; ++r) {
    const int n = row0 + r;
    if (n >= N) break;  // warp-uniform
    const uint4* wr = reinterpret_cast<const uint4*>(W + static_cast<long long>(n) * ldw);
    float acc = 0.f;
    for (int v = lane; v < nvec; v += 32) {
      const uint4 u = wr[v];
      const float* xv = xs + v * 8;
      acc = fmaf(bf16lo(u.x), xv[0], acc); acc = fmaf(bf16hi(u.x), xv[1], acc);
      acc = fmaf(bf16lo(u.y), xv[2], acc); acc = fmaf(bf16hi(u.y), xv[3], acc);
      acc = fmaf(bf16lo(u.z), xv[4], acc); acc = fmaf(bf16hi(u.z), xv[5], acc);
      acc = fmaf(bf16lo(u.w), xv[6], acc); acc = fmaf(bf16hi(u.w), xv[7], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      if (bias != nullptr) acc += __bfloat162float(bias[n]);
      if (post_act) acc = silu_f(acc);
      y[n] = acc;
      if (y_bf16 != nullptr) y_bf16[n] = __float2bfloat16(acc);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// classifier-free guidance + flow-match Euler update, one pass (pipeline_mova.py:456-460 + flow_match_pair.py:213-227):
//   out = sample + (nega + s * (posi - nega)) * dsigma          (nega == nullptr: out = sample + posi * dsigma)
// predictions bf16 (converted with .float() by the reference), latents fp32; 8 elements per thread.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cfg_euler_kernel(const __nv_bfloat16* __restrict__ posi, const __nv_bfloat16* __restrict__ nega,
                 const float* __restrict__ sample, float* __restrict__ out, long long n, float cfg_scale,
                 float dsigma) {
  const long long i0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i0 >= n) return;
  if (i0 + 8 <= n) {
    const uint4 up = *reinterpret_cast<const uint4*>(posi + i0);
    const uint32_t pw[4] = {up.x, up.y, up.z, up.w};
    uint32_t nw[4] = {0, 0, 0, 0};
    if (nega != nullptr) {
      const uint4 un = *reinterpret_cast<const uint4*>(nega + i0);
      nw[0] = un.x; nw[1] = un.y; nw[2] = un.z; nw[3] = un.w;
    }
    const float4 s0 = *reinterpret_cast<const float4*>(sample + i0);
    const float4 s1 = *reinterpret_cast<const float4*>(sample + i0 + 4);
    const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    float r[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float p0 = bf16lo(pw[e]), p1 = bf16hi(pw[e]);
      float g0 = p0, g1 = p1;
      if (nega != nullptr) {
        const float n0 = bf16lo(nw[e]), n1 = bf16hi(nw[e]);
        g0 = n0 + cfg_scale * (p0 - n0);
        g1 = n1 + cfg_scale * (p1 - n1);
      }
      r[2 * e] = sv[2 * e] + g0 * dsigma;
      r[2 * e + 1] = sv[2 * e + 1] + g1 * dsigma;
    }
    *reinterpret_cast<float4*>(out + i0) = make_float4(r[0], r[1], r[2], r[3]);
    *reinterpret_cast<float4*>(out + i0 + 4) = make_float4(r[4], r[5], r[6], r[7]);
  } else {
    for (long long i = i0; i < n; ++i) {
      const float pv = __bfloat162float(posi[i]);
      float g = pv;
      if (nega != nullptr) {
        const float nv = __bfloat162float(nega[i]);
        g = nv + cfg_scale * (pv - nv);
      }
      out[i] = sample[i] + g * dsigma;
    }
  }
}

}  // namespace mv

extern "C" {

int mova_b200_cfg_euler(const void* posi, const void* nega, const float* sample, float* out, int64_t n,
                        float cfg_scale, float dsigma, void* stream) {
  using namespace mv;
  MV_REQUIRE(posi && sample && out, "mova_b200_cfg_euler: null pointer");
  MV_REQUIRE(((reinterpret_cast<uintptr_t>(posi) | reinterpret_cast<uintptr_t>(nega) |
               reinterpret_cast<uintptr_t>(sample) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
             "mova_b200_cfg_euler: operands must be 16-byte aligned");
  if (n <= 0) return 0;
  const int threads = 256;
  const long long groups = (n + 7) / 8;
  const unsigned blocks = static_cast<unsigned>((groups + threads - 1) / threads);
  cfg_euler_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(posi), static_cast<const __nv_bfloat16*>(nega), sample, out, n, cfg_scale,
      dsigma);
  MV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mova_b200_patchify(const void* x, int x_is_f32, int C, int F, int H, int W, int pt, int ph, int pw, void* out,
                       int64_t ldo, void* stream) {
  using namespace mv;
  MV_REQUIRE(x && out, "mova_b200_patchify: null pointer");
  MV_REQUIRE(C > 0 && F > 0 && H > 0 && W > 0 && pt > 0 && ph > 0 && pw > 0, "mova_b200_patchify: bad shape");
  MV_REQUIRE(F % pt == 0 && H % ph == 0 && W % pw == 0,
             "mova_b200_patchify: latent %dx%dx%d is not a multiple of the patch %dx%dx%d", F, H, W, pt, ph, pw);
  const int K = C * pt * ph * pw;
  MV_REQUIRE(ldo >= K, "mova_b200_patchify: ldo (%lld) < K (%d)", static_cast<long long>(ldo), K);
  const long long L = static_cast<long long>(F / pt) * (H / ph) * (W / pw);
  const long long total = L * K;
  if (total == 0) return 0;
  const int threads = 256;
  const unsigned blocks = static_cast<unsigned>((total + threads - 1) / threads);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto* op = static_cast<__nv_bfloat16*>(out);
  if (x_is_f32)
    patchify_kernel<float><<<blocks, threads, 0, s>>>(static_cast<const float*>(x), op, ldo, C, F, H, W, pt, ph, pw, total);
  else
    patchify_kernel<__nv_bfloat16><<<blocks, threads, 0, s>>>(static_cast<const __nv_bfloat16*>(x), op, ldo, C, F, H,
                                                               W, pt, ph, pw, total);
  MV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mova_b200_unpatchify(const void* in, int64_t ldi, void* out, int Cout, int Fp, int Hp, int Wp, int pt, int ph,
                         int pw, void* stream) {
  using namespace mv;
  MV_REQUIRE(in && out, "mova_b200_unpatchify: null pointer");
  MV_REQUIRE(Cout > 0 && Fp > 0 && Hp > 0 && Wp > 0 && pt > 0 && ph > 0 && pw > 0, "mova_b200_unpatchify: bad shape");
  MV_REQUIRE(ldi >= static_cast<int64_t>(Cout) * pt * ph * pw, "mova_b200_unpatchify: ldi smaller than the row");
  const long long total = static_cast<long long>(Cout) * Fp * pt * Hp * ph * Wp * pw;
  const int threads = 256;
  const unsigned blocks = static_cast<unsigned>((total + threads - 1) / threads);
  unpatchify_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in), ldi, static_cast<__nv_bfloat16*>(out), Cout, Fp, Hp, Wp, pt, ph, pw, total);
  MV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mova_b200_sinusoidal(const float* t, float* out, int dim, void* stream) {
  using namespace mv;
  MV_REQUIRE(t && out, "mova_b200_sinusoidal: null pointer");
  MV_REQUIRE(dim > 0 && dim % 2 == 0, "mova_b200_sinusoidal: dim (%d) must be even", dim);
  const int half = dim / 2;
  sinusoidal_kernel<<<(half + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(t, out, half);
  MV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mova_b200_gemv_f32(const float* x, const void* W, int64_t ldw, const void* bias, float* y, void* y_bf16, int N,
                       int K, int pre_act, int post_act, void* stream) {
  using namespace mv;
  MV_REQUIRE(x && W && y, "mova_b200_gemv_f32: null pointer");
  MV_REQUIRE(N > 0 && K > 0 && K % 8 == 0 && ldw >= K && ldw % 8 == 0, "mova_b200_gemv_f32: bad shape N=%d K=%d ldw=%lld",
             N, K, static_cast<long long>(ldw));
  MV_REQUIRE((reinterpret_cast<uintptr_t>(W) & 15) == 0, "mova_b200_gemv_f32: W must be 16-byte aligned");
  MV_REQUIRE(K * 4 <= 48 * 1024, "mova_b200_gemv_f32: K=%d exceeds the shared-memory staging buffer", K);
  const int rows_per_block = (GEMV_THREADS / 32) * GEMV_ROWS_PER_WARP;
  const unsigned blocks = static_cast<unsigned>((N + rows_per_block - 1) / rows_per_block);
  gemv_f32_kernel<<<blocks, GEMV_THREADS, K * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<const __nv_bfloat16*>(W), ldw, static_cast<const __nv_bfloat16*>(bias), y,
      static_cast<__nv_bfloat16*>(y_bf16), N, K, pre_act, post_act);
  MV_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
