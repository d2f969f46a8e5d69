// K1 (input convolution on one-hot DNA), K3/K4 (profile + counts heads) and their gradients.
// These are the thin ends of the network (7 % of the FLOPs, SURVEY.md 8a rows 1,3,4); this first
// version runs them on CUDA cores with shared-memory tiling and fp32 weights.
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
// K1 weight gradient.  dw[j][c][f] = sum x[b,t+j,c] gz[b,t,f].  A block walks a strip of
// positions; thread (j, f) keeps one accumulator per base and adds gz[t'-j][f] into the
// accumulator of the base found at t' (uniform branch).  Partials go to a workspace.
// ------------------------------------------------------------------------------------------
constexpr int kIwStrip = 256;

__global__ void __launch_bounds__(1024)
input_conv_wgrad_kernel(int B, int L_in, int L_out, int W, int Fw, int F,
                        const uint8_t* __restrict__ x, const __nv_bfloat16* __restrict__ gz_hi,
                        const __nv_bfloat16* __restrict__ gz_lo, float* __restrict__ ws,
                        int strips_per_seq) {
  extern __shared__ float sm[];
  float* sg = sm;                                                   // [kIwStrip][Fw]
  float4* sx = reinterpret_cast<float4*>(sg + kIwStrip * Fw);       // [kIwStrip + W - 1] one-hot rows as fp32
  const int n_out = W * 4 * Fw + Fw;
  const int n_pairs = W * Fw;
  // per-thread accumulators for up to 2 (j,f) pairs
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  float accb = 0.f;
  const int total_strips = B * strips_per_seq;
  for (int strip = blockIdx.x; strip < total_strips; strip += gridDim.x) {
    const int b = strip / strips_per_seq;
    const int t0 = (strip % strips_per_seq) * kIwStrip;
    __syncthreads();
    for (int i = threadIdx.x; i < kIwStrip * Fw; i += blockDim.x) {
      const int r = i / Fw, f = i % Fw;
      const int t = t0 + r;
      float g = 0.f;
      if (t < L_out) {
        const size_t idx = (static_cast<size_t>(b) * L_out + t) * F + f;
        g = bf16_at(gz_hi + idx);
        if (gz_lo) g += bf16_at(gz_lo + idx);
      }
      sg[i] = g;
    }
    const int nrows = kIwStrip + W - 1;
    for (int i = threadIdx.x; i < nrows; i += blockDim.x) {
      const int r = t0 + i;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < L_in) {
        const uchar4 u = *reinterpret_cast<const uchar4*>(x + (static_cast<size_t>(b) * L_in + r) * 4);
        v = make_float4(u.x, u.y, u.z, u.w);
      }
      sx[i] = v;
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int pair = blockIdx.y * 2048 + threadIdx.x + s * blockDim.x;
      if (pair < n_pairs) {
        const int j = pair / Fw, f = pair % Fw;
        for (int r = 0; r < kIwStrip; ++r) {
          const float g = sg[r * Fw + f];
          const float4 xv = sx[r + j];
          acc[s][0] = fmaf(xv.x, g, acc[s][0]);
          acc[s][1] = fmaf(xv.y, g, acc[s][1]);
          acc[s][2] = fmaf(xv.z, g, acc[s][2]);
          acc[s][3] = fmaf(xv.w, g, acc[s][3]);
        }
      }
    }
    if (blockIdx.y == 0 && threadIdx.x < Fw) {
      for (int r = 0; r < kIwStrip; ++r) accb += sg[r * Fw + threadIdx.x];
    }
  }
  float* wsp = ws + static_cast<size_t>(blockIdx.x) * n_out;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int pair = blockIdx.y * 2048 + threadIdx.x + s * blockDim.x;
    if (pair < n_pairs) {
      const int j = pair / Fw, f = pair % Fw;
#pragma unroll
      for (int c = 0; c < 4; ++c) wsp[(j * 4 + c) * Fw + f] = acc[s][c];
    }
  }
  if (blockIdx.y == 0 && threadIdx.x < Fw) wsp[W * 4 * Fw + threadIdx.x] = accb;
}

__global__ void partial_reduce_kernel(const float* __restrict__ ws, int nparts, long long n,
                                      long long n_first, float* __restrict__ out_first,
                                      float* __restrict__ out_second, int accumulate) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float s0 = 0.f, s1 = 0.f;
    int s = 0;
    for (; s + 2 <= nparts; s += 2) {
      s0 += ws[s * n + i];
      s1 += ws[(s + 1) * n + i];
    }
    if (s < nparts) s0 += ws[s * n + i];
    const float tot = s0 + s1;
    float* dst = (i < n_first) ? (out_first + i) : (out_second ? out_second + (i - n_first) : nullptr);
    if (dst) *dst = accumulate ? (*dst + tot) : tot;
  }
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
  const int stride = Fw + 1;
  float* sa = sm;  // [rows][Fw+1] fp32 (hi+lo summed)
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * 128;
  for (int i = threadIdx.x; i < rows * Fw; i += blockDim.x) {
    const int r = i / Fw, c = i % Fw;
    const int t = t0 + r;
    float v = 0.f;
    if (t < L_in) {
      const size_t idx = (static_cast<size_t>(b) * L_in + t) * F + c;
      v = bf16_at(act_hi + idx);
      if (act_lo) v += bf16_at(act_lo + idx);
    }
    sa[r * stride + c] = v;
  }
  __syncthreads();
  // global average pool partial sums over this block's own 128 rows
  if (gap) {
    for (int c = threadIdx.x; c < Fw; c += blockDim.x) {
      float s = 0.f;
      for (int r = 0; r < 128; ++r) s += sa[r * stride + c];  // rows >= L_in are zero
      atomicAdd(gap + static_cast<size_t>(b) * F + c, s / static_cast<float>(L_in));
    }
  }
  const int t = t0 + threadIdx.x;
  if (t < L_out) {
    float acc[TP];
#pragma unroll
    for (int k = 0; k < TP; ++k) acc[k] = (k < T) ? pb[k] : 0.f;
    for (int j = 0; j < W; ++j) {
      const float* ar = sa + (threadIdx.x + j) * stride;
      const float* wr = pw + static_cast<size_t>(j) * Fw * T;
      for (int c = 0; c < Fw; ++c) {
        const float a = ar[c];
#pragma unroll
        for (int k = 0; k < TP; ++k)
          if (k < T) acc[k] = fmaf(a, __ldg(wr + c * T + k), acc[k]);
      }
    }
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

// ------------------------------------------------------------------------------------------
// K3/K4 backward, data gradient:  g[b,s,c] = sum_j sum_k dlogits[b,s-j,k] pw[j][c][k]
//                                          + sum_h dlc[b,h] cw[c][h] / L_in
// block = 64 positions, warp = position strip, lane = channel pair.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
heads_dgrad_kernel(int L_in, int L_out, int W, int Fw, int F, int T, int H,
                   const float* __restrict__ pw, const float* __restrict__ cw,
                   const float* __restrict__ dlogits, const float* __restrict__ dlc,
                   __nv_bfloat16* __restrict__ g_hi, __nv_bfloat16* __restrict__ g_lo,
                   __nv_bfloat16* __restrict__ gz_hi, __nv_bfloat16* __restrict__ gz_lo,
                   const unsigned long long* __restrict__ mask_in,
                   const __nv_bfloat16* __restrict__ mult_hi,
                   const __nv_bfloat16* __restrict__ mult_lo) {
  extern __shared__ float sm[];
  const int rows = 64 + W - 1;
  float* sd = sm;  // [rows][T]: dlogits rows s0-(W-1) .. s0+63
  const int b = blockIdx.y;
  const int s0 = blockIdx.x * 64;
  for (int i = threadIdx.x; i < rows * T; i += blockDim.x) {
    const int r = i / T, k = i % T;
    const int t = s0 - (W - 1) + r;
    sd[i] = (t >= 0 && t < L_out) ? dlogits[(static_cast<size_t>(b) * L_out + t) * T + k] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  for (int c = lane; c < F; c += 32) {
    float cterm = 0.f;
    if (c < Fw)
      for (int h = 0; h < H; ++h) cterm = fmaf(dlc[b * H + h], cw[c * H + h], cterm);
    cterm /= static_cast<float>(L_in);
    for (int sl = warp; sl < 64; sl += nwarps) {
      const int s = s0 + sl;
      if (s >= L_in) break;
      float acc = cterm;
      if (c < Fw) {
        for (int j = 0; j < W; ++j) {
          const float* dr = sd + (sl + (W - 1) - j) * T;
          const float* wr = pw + (static_cast<size_t>(j) * Fw + c) * T;
          for (int k = 0; k < T; ++k) acc = fmaf(dr[k], __ldg(wr + k), acc);
        }
      } else {
        acc = 0.f;
      }
      const size_t row = static_cast<size_t>(b) * L_in + s;
      const size_t idx = row * F + c;
      if (g_hi) store_planes(g_hi, g_lo, idx, acc);
      if (gz_hi) {
        float m = 1.f;
        if (mult_hi) {
          m = bf16_at(mult_hi + idx);
          if (mult_lo) m += bf16_at(mult_lo + idx);
        } else if (mask_in) {
          m = ((mask_in[row * (F / 64) + (c >> 6)] >> (c & 63)) & 1ull) ? 1.f : 0.f;
        }
        store_planes(gz_hi, gz_lo, idx, acc * m);
      }
    }
  }
}

// K3/K4 backward, parameter gradients.  thread = (tap j, channel c), T accumulators; blocks walk
// strips of 128 positions and write partials to a workspace.
template <int TP>
__global__ void __launch_bounds__(512)
heads_wgrad_kernel(int B, int L_in, int L_out, int W, int Fw, int F, int T,
                   const __nv_bfloat16* __restrict__ act_hi,
                   const __nv_bfloat16* __restrict__ act_lo, const float* __restrict__ dlogits,
                   float* __restrict__ ws, int strips_per_seq) {
  extern __shared__ float sm[];
  const int rows = 128 + W - 1;
  const int stride = Fw + 1;
  float* sa = sm;                    // [rows][Fw+1]
  float* sd = sa + rows * stride;    // [128][T]
  const int n_pairs = W * Fw;
  const int n_out = n_pairs * T + T;
  float* wsp = ws + static_cast<size_t>(blockIdx.x) * n_out;
  for (int i = threadIdx.x; i < n_out; i += blockDim.x) wsp[i] = 0.f;
  const int total = B * strips_per_seq;
  for (int strip = blockIdx.x; strip < total; strip += gridDim.x) {
    const int b = strip / strips_per_seq;
    const int t0 = (strip % strips_per_seq) * 128;
    __syncthreads();
    for (int i = threadIdx.x; i < rows * Fw; i += blockDim.x) {
      const int r = i / Fw, c = i % Fw;
      const int t = t0 + r;
      float v = 0.f;
      if (t < L_in) {
        const size_t idx = (static_cast<size_t>(b) * L_in + t) * F + c;
        v = bf16_at(act_hi + idx);
        if (act_lo) v += bf16_at(act_lo + idx);
      }
      sa[r * stride + c] = v;
    }
    for (int i = threadIdx.x; i < 128 * T; i += blockDim.x) {
      const int t = t0 + i / T;
      sd[i] = (t < L_out) ? dlogits[(static_cast<size_t>(b) * L_out + t) * T + (i % T)] : 0.f;
    }
    __syncthreads();
    for (int pair = threadIdx.x; pair < n_pairs; pair += blockDim.x) {
      const int j = pair / Fw, c = pair % Fw;
      float acc[TP];
#pragma unroll
      for (int k = 0; k < TP; ++k) acc[k] = 0.f;
      for (int r = 0; r < 128; ++r) {
        const float a = sa[(r + j) * stride + c];
#pragma unroll
        for (int k = 0; k < TP; ++k)
          if (k < T) acc[k] = fmaf(a, sd[r * T + k], acc[k]);
      }
#pragma unroll
      for (int k = 0; k < TP; ++k)
        if (k < T) wsp[pair * T + k] += acc[k];
    }
    if (threadIdx.x < T) {
      float s = 0.f;
      for (int r = 0; r < 128; ++r) s += sd[r * T + threadIdx.x];
      wsp[n_pairs * T + threadIdx.x] += s;
    }
  }
}

// dcw[c][h] = sum_b gap[b,c] dlc[b,h];  dcb[h] = sum_b dlc[b,h]
__global__ void counts_wgrad_kernel(int B, int Fw, int F, int H, const float* __restrict__ gap,
                                    const float* __restrict__ dlc, float* __restrict__ dcw,
                                    float* __restrict__ dcb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Fw * H) {
    const int c = i / H, h = i % H;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s = fmaf(gap[static_cast<size_t>(b) * F + c], dlc[b * H + h], s);
    dcw[i] = s;
  } else if (i < Fw * H + H) {
    const int h = i - Fw * H;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dlc[b * H + h];
    dcb[h] = s;
  }
}

static int pad_tasks(int T) { return T <= 2 ? 2 : (T <= 4 ? 4 : (T <= 8 ? 8 : 16)); }

}  // namespace bp

using namespace bp;

static int iw_blocks(int B, int L_out) {
  int strips = B * ((L_out + kIwStrip - 1) / kIwStrip);
  int g = sm_count();
  if (g <= 0) g = 1;
  return strips < g ? strips : g;
}

extern "C" long long bpb_input_conv_wgrad_ws_bytes(int B, int L_out, int W, int Fw) {
  if (sm_count() <= 0) {
    set_error("bpb_input_conv_wgrad_ws_bytes: no CUDA device");
    return -1;
  }
  return static_cast<long long>(iw_blocks(B, L_out)) * (W * 4 * Fw + Fw) * 4;
}

extern "C" int bpb_input_conv_wgrad(int B, int L_in, int L_out, int W, int Fw, int F,
                                    const uint8_t* x, const bpb_bf16* gz_hi, const bpb_bf16* gz_lo,
                                    float* dw, float* db, float* ws, long long ws_bytes,
                                    void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(x && gz_hi && dw && db && ws, "bpb_input_conv_wgrad: NULL argument");
  BP_REQUIRE(Fw <= 1024, "bpb_input_conv_wgrad: Fw=%d exceeds 1024", Fw);
  const int blocks = iw_blocks(B, L_out);
  const long long n_out = static_cast<long long>(W) * 4 * Fw + Fw;
  BP_REQUIRE(ws_bytes >= blocks * n_out * 4, "bpb_input_conv_wgrad: workspace too small");
  const size_t smem = static_cast<size_t>(kIwStrip) * Fw * 4 + (kIwStrip + W - 1) * 16 + 16;
  static bool configured = false;
  if (!configured) {
    BP_CHECK_CUDA(cudaFuncSetAttribute(input_conv_wgrad_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  BP_REQUIRE(smem <= 200 * 1024, "bpb_input_conv_wgrad: too many filters for shared memory");
  const int strips_per_seq = (L_out + kIwStrip - 1) / kIwStrip;
  dim3 wgrid(blocks, (W * Fw + 2047) / 2048);
  input_conv_wgrad_kernel<<<wgrid, 1024, smem, stream>>>(
      B, L_in, L_out, W, Fw, F, x, reinterpret_cast<const __nv_bfloat16*>(gz_hi),
      reinterpret_cast<const __nv_bfloat16*>(gz_lo), ws, strips_per_seq);
  BP_CHECK_CUDA(cudaGetLastError());
  const int rb = static_cast<int>((n_out + 255) / 256);
  partial_reduce_kernel<<<rb, 256, 0, stream>>>(ws, blocks, n_out,
                                                static_cast<long long>(W) * 4 * Fw, dw, db, 0);
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

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
  const size_t smem = static_cast<size_t>(128 + W - 1) * (Fw + 1) * 4;
  BP_REQUIRE(smem <= 200 * 1024, "bpb_heads_fwd: tile does not fit shared memory (W=%d Fw=%d)", W,
             Fw);
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

static int hw_blocks(int B, int L_in) {
  int strips = B * ((L_in + 127) / 128);
  int g = sm_count();
  if (g <= 0) g = 1;
  return strips < g ? strips : g;
}

extern "C" long long bpb_heads_bwd_ws_bytes(int B, int L_in, int W, int Fw, int T) {
  if (sm_count() <= 0) {
    set_error("bpb_heads_bwd_ws_bytes: no CUDA device");
    return -1;
  }
  return static_cast<long long>(hw_blocks(B, L_in)) * (static_cast<long long>(W) * Fw * T + T) * 4;
}

extern "C" int bpb_heads_bwd(int B, int L_in, int L_out, int W, int Fw, int F, int T, int H,
                             const bpb_bf16* act_hi, const bpb_bf16* act_lo, const float* gap,
                             const float* pw, const float* cw, const float* dlogits,
                             const float* dlogcounts, float* dpw, float* dpb, float* dcw,
                             float* dcb, bpb_bf16* g_hi, bpb_bf16* g_lo, bpb_bf16* gz_hi,
                             bpb_bf16* gz_lo, const uint64_t* mask_in, const bpb_bf16* mult_hi,
                             const bpb_bf16* mult_lo, float* ws, long long ws_bytes,
                             void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(pw && cw && dlogits && dlogcounts, "bpb_heads_bwd: NULL argument");
  BP_REQUIRE(T >= 1 && T <= 16, "bpb_heads_bwd: total tasks T=%d must be in 1..16", T);
  BP_REQUIRE(g_hi || gz_hi || dpw, "bpb_heads_bwd: nothing requested");
  if (g_hi || gz_hi) {
    const size_t smem = static_cast<size_t>(64 + W - 1) * T * 4;
    dim3 grid((L_in + 63) / 64, B);
    heads_dgrad_kernel<<<grid, 256, smem, stream>>>(
        L_in, L_out, W, Fw, F, T, H, pw, cw, dlogits, dlogcounts,
        reinterpret_cast<__nv_bfloat16*>(g_hi), reinterpret_cast<__nv_bfloat16*>(g_lo),
        reinterpret_cast<__nv_bfloat16*>(gz_hi), reinterpret_cast<__nv_bfloat16*>(gz_lo),
        reinterpret_cast<const unsigned long long*>(mask_in),
        reinterpret_cast<const __nv_bfloat16*>(mult_hi),
        reinterpret_cast<const __nv_bfloat16*>(mult_lo));
    BP_CHECK_CUDA(cudaGetLastError());
  }
  if (dpw) {
    BP_REQUIRE(act_hi && gap && dpb && dcw && dcb && ws, "bpb_heads_bwd: NULL gradient argument");
    const int blocks = hw_blocks(B, L_in);
    const long long n_out = static_cast<long long>(W) * Fw * T + T;
    BP_REQUIRE(ws_bytes >= blocks * n_out * 4, "bpb_heads_bwd: workspace too small");
    const size_t smem = (static_cast<size_t>(128 + W - 1) * (Fw + 1) + 128 * T) * 4;
    BP_REQUIRE(smem <= 200 * 1024, "bpb_heads_bwd: tile does not fit shared memory");
    const int strips_per_seq = (L_in + 127) / 128;
    const __nv_bfloat16* ah = reinterpret_cast<const __nv_bfloat16*>(act_hi);
    const __nv_bfloat16* al = reinterpret_cast<const __nv_bfloat16*>(act_lo);
#define LAUNCH_HW(TP)                                                                          \
  do {                                                                                         \
    static bool configured = false;                                                            \
    if (!configured) {                                                                         \
      BP_CHECK_CUDA(cudaFuncSetAttribute(heads_wgrad_kernel<TP>,                               \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                         200 * 1024));                                         \
      configured = true;                                                                       \
    }                                                                                          \
    heads_wgrad_kernel<TP><<<blocks, 512, smem, stream>>>(B, L_in, L_out, W, Fw, F, T, ah, al, \
                                                          dlogits, ws, strips_per_seq);        \
  } while (0)
    switch (pad_tasks(T)) {
      case 2: LAUNCH_HW(2); break;
      case 4: LAUNCH_HW(4); break;
      case 8: LAUNCH_HW(8); break;
      default: LAUNCH_HW(16); break;
    }
#undef LAUNCH_HW
    BP_CHECK_CUDA(cudaGetLastError());
    const int rb = static_cast<int>((n_out + 255) / 256);
    partial_reduce_kernel<<<rb, 256, 0, stream>>>(ws, blocks, n_out,
                                                  static_cast<long long>(W) * Fw * T, dpw, dpb, 0);
    BP_CHECK_CUDA(cudaGetLastError());
    counts_wgrad_kernel<<<(Fw * H + H + 127) / 128, 128, 0, stream>>>(B, Fw, F, H, gap,
                                                                      dlogcounts, dcw, dcb);
    BP_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}
