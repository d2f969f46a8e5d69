// The remaining CUDA-core kernels of the thin ends of the network: K3/K4 forward (profile + counts
// heads) and the input-convolution data gradient that only attribution needs.  (The input
// convolution forward, the heads' data gradient and every parameter gradient run on tensor cores:
// rowconv_tc.cu, wgrad_tc.cu.)
//   reference: models.py:85-90 (initial conv), models.py:20-50 (heads).
#include "common.cuh"
#include "ptx.cuh"

namespace bp {

__device__ __forceinline__ float bf16_at(const __nv_bfloat16* p) { return __bfloat162float(*p); }

__device__ __forceinline__ void store_planes(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t idx,
                                             float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[idx] = h;
  if (lo) lo[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// ------------------------------------------------------------------------------------------
// K1 data gradient (attribution only): dx[b,s,c] = sum_j sum_f gz[b,s-j,f] w[j][c][f]
// block = 64 positions; warp per position strip, lanes over filters, shuffle reduction.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
input_conv_dgrad_kernel(int L_in, int L_out, int W, int Fw, int F,
                        const __nv_bfloat16* __restrict__ gz_hi,
                        const __nv_bfloat16* __restrict__ gz_lo, const float* __restrict__ w,
                        float* __restrict__ dx) {
  extern __shared__ float sm[];
  float* sw = sm;  // [W][4][Fw]
  for (int i = threadIdx.x; i < W * 4 * Fw; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  for (int s = blockIdx.x * 64 + warp; s < min(L_in, (int)(blockIdx.x * 64 + 64)); s += nwarps) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int j = 0; j < W; ++j) {
      const int t = s - j;
      if (t < 0 || t >= L_out) continue;
      const size_t base = (static_cast<size_t>(b) * L_out + t) * F;
      for (int f = lane; f < Fw; f += 32) {
        float g = bf16_at(gz_hi + base + f);
        if (gz_lo) g += bf16_at(gz_lo + base + f);
        a0 = fmaf(g, sw[(j * 4 + 0) * Fw + f], a0);
        a1 = fmaf(g, sw[(j * 4 + 1) * Fw + f], a1);
        a2 = fmaf(g, sw[(j * 4 + 2) * Fw + f], a2);
        a3 = fmaf(g, sw[(j * 4 + 3) * Fw + f], a3);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o);
      a3 += __shfl_xor_sync(0xffffffffu, a3, o);
    }
    if (lane == 0)
      *reinterpret_cast<float4*>(dx + (static_cast<size_t>(b) * L_in + s) * 4) =
          make_float4(a0, a1, a2, a3);
  }
}

// ------------------------------------------------------------------------------------------
// K3/K4 forward: profile logits for all heads + global average pool.
// block = 128 positions of one sequence (128 threads, one position each, all T tasks).
// ------------------------------------------------------------------------------------------
template <int TP>
__global__ void __launch_bounds__(128)
heads_fwd_kernel(int L_in, int L_out, int W, int Fw, int F, int T,
                 const __nv_bfloat16* __restrict__ act_hi, const __nv_bfloat16* __restrict__ act_lo,
                 const float* __restrict__ pw, const float* __restrict__ pb,
                 float* __restrict__ logits, float* __restrict__ gap) {
  extern __shared__ float sm[];
  const int rows = 128 + W - 1;
  constexpr int CH = 64;        // channels staged per pass (any number of filters fits)
  constexpr int stride = CH + 1;
  float* sa = sm;  // [rows][CH+1] fp32 (hi+lo summed)
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * 128;
  const int t = t0 + threadIdx.x;
  float acc[TP];
#pragma unroll
  for (int k = 0; k < TP; ++k) acc[k] = (k < T) ? pb[k] : 0.f;
  for (int c0 = 0; c0 < Fw; c0 += CH) {
    const int nc = min(CH, Fw - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < rows * nc; i += blockDim.x) {
      const int r = i / nc, c = i % nc;
      const int tt = t0 + r;
      float v = 0.f;
      if (tt < L_in) {
        const size_t idx = (static_cast<size_t>(b) * L_in + tt) * F + c0 + c;
        v = bf16_at(act_hi + idx);
        if (act_lo) v += bf16_at(act_lo + idx);
      }
      sa[r * stride + c] = v;
    }
    __syncthreads();
    // global average pool partial sums over this block's own 128 rows
    if (gap) {
      for (int c = threadIdx.x; c < nc; c += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < 128; ++r) s += sa[r * stride + c];  // rows >= L_in are zero
        atomicAdd(gap + static_cast<size_t>(b) * F + c0 + c, s / static_cast<float>(L_in));
      }
    }
    if (t < L_out) {
      for (int j = 0; j < W; ++j) {
        const float* ar = sa + (threadIdx.x + j) * stride;
        const float* wr = pw + (static_cast<size_t>(j) * Fw + c0) * T;
        for (int c = 0; c < nc; ++c) {
          const float a = ar[c];
#pragma unroll
          for (int k = 0; k < TP; ++k)
            if (k < T) acc[k] = fmaf(a, __ldg(wr + c * T + k), acc[k]);
        }
      }
    }
  }
  if (t < L_out) {
    float* lo = logits + (static_cast<size_t>(b) * L_out + t) * T;
#pragma unroll
    for (int k = 0; k < TP; ++k)
      if (k < T) lo[k] = acc[k];
  }
}

__global__ void counts_dense_kernel(int B, int Fw, int F, int H, const float* __restrict__ gap,
                                    const float* __restrict__ cw, const float* __restrict__ cb,
                                    float* __restrict__ logcounts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  const int b = i / H, h = i % H;
  float s = cb[h];
  for (int c = 0; c < Fw; ++c) s = fmaf(gap[static_cast<size_t>(b) * F + c], cw[c * H + h], s);
  logcounts[i] = s;
}

static int pad_tasks(int T) { return T <= 2 ? 2 : (T <= 4 ? 4 : (T <= 8 ? 8 : 16)); }

}  // namespace bp

using namespace bp;

extern "C" int bpb_input_conv_dgrad(int B, int L_in, int L_out, int W, int Fw, int F,
                                    const bpb_bf16* gz_hi, const bpb_bf16* gz_lo, const float* w,
                                    float* dx, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(gz_hi && w && dx, "bpb_input_conv_dgrad: NULL argument");
  const size_t smem = static_cast<size_t>(W) * 4 * Fw * 4;
  static bool configured = false;
  if (!configured) {
    BP_CHECK_CUDA(cudaFuncSetAttribute(input_conv_dgrad_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  BP_REQUIRE(smem <= 200 * 1024, "bpb_input_conv_dgrad: filter too large for shared memory");
  dim3 grid((L_in + 63) / 64, B);
  input_conv_dgrad_kernel<<<grid, 256, smem, stream>>>(
      L_in, L_out, W, Fw, F, reinterpret_cast<const __nv_bfloat16*>(gz_hi),
      reinterpret_cast<const __nv_bfloat16*>(gz_lo), w, dx);
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int bpb_heads_fwd(int B, int L_in, int L_out, int W, int Fw, int F, int T, int H,
                             const bpb_bf16* act_hi, const bpb_bf16* act_lo, const float* pw,
                             const float* pb, const float* cw, const float* cb, float* logits,
                             float* gap, float* logcounts, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(act_hi && pw && pb && cw && cb && logits && gap && logcounts,
             "bpb_heads_fwd: NULL argument");
  BP_REQUIRE(T >= 1 && T <= 16, "bpb_heads_fwd: total tasks T=%d must be in 1..16", T);
  BP_REQUIRE(L_out == L_in - W + 1, "bpb_heads_fwd: L_out must be L_in - W + 1");
  const size_t smem = static_cast<size_t>(128 + W - 1) * 65 * 4;  // 64-channel staging passes
  BP_REQUIRE(smem <= 200 * 1024, "bpb_heads_fwd: filter width W=%d too large", W);
  BP_CHECK_CUDA(cudaMemsetAsync(gap, 0, static_cast<size_t>(B) * F * 4, stream));
  dim3 grid((L_in + 127) / 128, B);
  const __nv_bfloat16* ah = reinterpret_cast<const __nv_bfloat16*>(act_hi);
  const __nv_bfloat16* al = reinterpret_cast<const __nv_bfloat16*>(act_lo);
#define LAUNCH_HF(TP)                                                                          \
  do {                                                                                         \
    static bool configured = false;                                                            \
    if (!configured) {                                                                         \
      BP_CHECK_CUDA(cudaFuncSetAttribute(heads_fwd_kernel<TP>,                                 \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                         200 * 1024));                                         \
      configured = true;                                                                       \
    }                                                                                          \
    heads_fwd_kernel<TP><<<grid, 128, smem, stream>>>(L_in, L_out, W, Fw, F, T, ah, al, pw, pb, \
                                                      logits, gap);                            \
  } while (0)
  switch (pad_tasks(T)) {
    case 2: LAUNCH_HF(2); break;
    case 4: LAUNCH_HF(4); break;
    case 8: LAUNCH_HF(8); break;
    default: LAUNCH_HF(16); break;
  }
#undef LAUNCH_HF
  BP_CHECK_CUDA(cudaGetLastError());
  counts_dense_kernel<<<(B * H + 127) / 128, 128, 0, stream>>>(B, Fw, F, H, gap, cw, cb, logcounts);
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

