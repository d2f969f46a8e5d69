// K5 (fused log-softmax / multinomial NLL / log-counts MSE forward+backward), K9 (Keras Adam),
// weight re-layout for the tcgen05 kernels, the epoch jitter gather that replaces libslide, and
// the scalar regressions of the transformation / combined models.
#include "common.cuh"
#include "ptx.cuh"

namespace bp {

template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, Op op, T* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = scratch[0];
  for (int i = 1; i < nw; ++i) v = op(v, scratch[i]);
  return v;
}

struct SumF { __device__ float operator()(float a, float b) const { return a + b; } };
struct MaxF { __device__ float operator()(float a, float b) const { return fmaxf(a, b); } };
struct SumD { __device__ double operator()(double a, double b) const { return a + b; } };

// ------------------------------------------------------------------------------------------
// K5.  losses.py:15-39 (tfp Multinomial.log_prob over the flattened L*T_h block of one head),
// losses.py:42-63 + Keras sum_over_batch_size.  One block per (sequence, head).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
loss_kernel(int B, int L, int T, int H, const int* __restrict__ task_off,
            const float* __restrict__ logits, const float* __restrict__ logcounts,
            const float* __restrict__ counts, const float* __restrict__ logcounts_true,
            const float* __restrict__ profile_weight, const float* __restrict__ lambda,
            float* __restrict__ losses, float* __restrict__ dlogits,
            float* __restrict__ dlogcounts) {
  __shared__ float scratch_f[32];
  __shared__ double scratch_d[32];
  const int b = blockIdx.x, h = blockIdx.y;
  const int k0 = task_off[h], k1 = task_off[h + 1];
  const int Th = k1 - k0;
  const int n = L * Th;
  const float* lg = logits + static_cast<size_t>(b) * L * T;
  const float* ct = counts + static_cast<size_t>(b) * L * T;
  float mx = -INFINITY;
  double sN = 0.0, sLg = 0.0, sKl = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int idx = (i / Th) * T + k0 + (i % Th);
    const float l = lg[idx], k = ct[idx];
    mx = fmaxf(mx, l);
    sN += k;
    sLg += lgammaf(k + 1.f);
    sKl += static_cast<double>(k) * l;
  }
  mx = block_reduce(mx, MaxF(), scratch_f);
  sN = block_reduce(sN, SumD(), scratch_d);
  sLg = block_reduce(sLg, SumD(), scratch_d);
  sKl = block_reduce(sKl, SumD(), scratch_d);
  float se = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int idx = (i / Th) * T + k0 + (i % Th);
    se += __expf(lg[idx] - mx);
  }
  se = block_reduce(se, SumF(), scratch_f);
  const float lse = mx + logf(se);
  const float N = static_cast<float>(sN);
  const float invB = 1.f / static_cast<float>(B);
  if (threadIdx.x == 0) {
    const double logp = lgamma(sN + 1.0) - sLg + sKl - sN * static_cast<double>(lse);
    atomicAdd(&losses[h], static_cast<float>(-logp) * invB);
    const float y = logcounts_true ? logcounts_true[b * H + h] : logf(N);
    const float d = logcounts[b * H + h] - y;
    atomicAdd(&losses[H + h], lambda[h] * d * d * invB);
    if (dlogcounts) dlogcounts[b * H + h] = 2.f * lambda[h] * d * invB;
  }
  if (dlogits) {
    float* dl = dlogits + static_cast<size_t>(b) * L * T;
    const float wh = profile_weight[h] * invB;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int idx = (i / Th) * T + k0 + (i % Th);
      dl[idx] = wh * (N * __expf(lg[idx] - lse) - ct[idx]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// K9.  Keras-3 Adam on a flat bucket.
// ------------------------------------------------------------------------------------------
__global__ void adam_kernel(long long n, float* __restrict__ p, const float* __restrict__ g,
                            float* __restrict__ m, float* __restrict__ v,
                            const float* __restrict__ step_and_lr, float b1, float b2, float eps,
                            float gscale) {
  const double step = static_cast<double>(step_and_lr[0]);
  const float lr = step_and_lr[1];
  const float alpha = static_cast<float>(static_cast<double>(lr) * sqrt(1.0 - pow((double)b2, step)) /
                                         (1.0 - pow((double)b1, step)));
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = m[i] + (gi - m[i]) * (1.f - b1);
    const float vi = v[i] + (gi * gi - v[i]) * (1.f - b2);
    m[i] = mi;
    v[i] = vi;
    p[i] -= alpha * mi / (sqrtf(vi) + eps);
  }
}

// fp32 Keras kernels [layers][3][Fw][Fw] (tap, cin, cout) -> bf16 operand layouts
// [layers][planes][3][F][F]
__global__ void prep_conv3_kernel(int n_layers, int Fw, int F, const float* __restrict__ w,
                                  __nv_bfloat16* __restrict__ fwd, __nv_bfloat16* __restrict__ bwd,
                                  int precise) {
  const long long per = 3LL * F * F;
  const long long total = per * n_layers;
  const int planes = precise ? 2 : 1;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int l = static_cast<int>(i / per);
    const int r = static_cast<int>(i % per);
    const int j = r / (F * F), n = (r / F) % F, c = r % F;
    const float* wl = w + static_cast<size_t>(l) * 3 * Fw * Fw;
    // fwd[j][n][c] = w[j][c][n] ; bwd[j][n][c] = w[j][n][c]
    const float vf = (n < Fw && c < Fw) ? wl[(j * Fw + c) * Fw + n] : 0.f;
    const float vb = (n < Fw && c < Fw) ? wl[(j * Fw + n) * Fw + c] : 0.f;
    const __nv_bfloat16 hf = __float2bfloat16_rn(vf), hb = __float2bfloat16_rn(vb);
    const size_t o = static_cast<size_t>(l) * planes * per + r;
    fwd[o] = hf;
    bwd[o] = hb;
    if (precise) {
      fwd[o + per] = __float2bfloat16_rn(vf - __bfloat162float(hf));
      bwd[o + per] = __float2bfloat16_rn(vb - __bfloat162float(hb));
    }
  }
}

// ------------------------------------------------------------------------------------------
// generators.py:129-138 / libslide.c:44-91
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void slide_kernel(const T* __restrict__ src, T* __restrict__ dst, int src_cols,
                             int dst_cols, int depth, const int* __restrict__ row_idx,
                             const int* __restrict__ col_idx) {
  const int r = blockIdx.x;
  const size_t n = static_cast<size_t>(dst_cols) * depth;
  const T* s = src + (static_cast<size_t>(r) * src_cols + col_idx[r]) * depth;
  T* d = dst + static_cast<size_t>(row_idx[r]) * n;
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
}

__global__ void __launch_bounds__(256)
log_row_sums_kernel(const float* __restrict__ vals, float* __restrict__ out, int n_per_row) {
  __shared__ float scratch[32];
  const float* v = vals + static_cast<size_t>(blockIdx.x) * n_per_row;
  float s = 0.f;
  for (int i = threadIdx.x; i < n_per_row; i += blockDim.x) s += v[i];
  s = block_reduce(s, SumF(), scratch);
  if (threadIdx.x == 0) out[blockIdx.x] = logf(s);
}

// ------------------------------------------------------------------------------------------
// models.py:118-175 (simple transformation) and layers.py:52-75 (CountsLogSumExp)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_fn(int a, float x) {
  return a == 1 ? 1.f / (1.f + __expf(-x)) : (a == 2 ? fmaxf(x, 0.f) : x);
}

__global__ void transform_fwd_kernel(long long n, const float* __restrict__ u,
                                     float* __restrict__ y, int n_terms,
                                     const int* __restrict__ act,
                                     const float* __restrict__ coeffs) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float x = u[i];
    float s = 0.f;
    for (int t = 0; t < n_terms; ++t) {
      const float m1 = coeffs[4 * t], b1 = coeffs[4 * t + 1];
      const float m2 = coeffs[4 * t + 2], b2 = coeffs[4 * t + 3];
      if (act[t] == 0) s += m1 * x + b1;
      else s += m1 * act_fn(act[t], m2 * x + b2) + b1;
    }
    y[i] = s;
  }
}

// dcoeffs[t] += sum_i dy[i] * d y_i / d {m1,b1,m2,b2}
__global__ void __launch_bounds__(256)
transform_bwd_kernel(long long n, const float* __restrict__ u, const float* __restrict__ dy,
                     int n_terms, const int* __restrict__ act, const float* __restrict__ coeffs,
                     float* __restrict__ dcoeffs) {
  __shared__ float scratch[32];
  for (int t = 0; t < n_terms; ++t) {
    const float m1 = coeffs[4 * t], m2 = coeffs[4 * t + 2], b2 = coeffs[4 * t + 3];
    const int a = act[t];
    float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      const float x = u[i], d = dy[i];
      if (a == 0) {
        g0 += d * x;
        g1 += d;
      } else {
        const float zin = m2 * x + b2;
        const float f = act_fn(a, zin);
        const float df = (a == 1) ? f * (1.f - f) : (zin > 0.f ? 1.f : 0.f);
        g0 += d * f;
        g1 += d;
        g2 += d * m1 * df * x;
        g3 += d * m1 * df;
      }
    }
    g0 = block_reduce(g0, SumF(), scratch);
    g1 = block_reduce(g1, SumF(), scratch);
    g2 = block_reduce(g2, SumF(), scratch);
    g3 = block_reduce(g3, SumF(), scratch);
    if (threadIdx.x == 0) {
      atomicAdd(&dcoeffs[4 * t], g0);
      atomicAdd(&dcoeffs[4 * t + 1], g1);
      atomicAdd(&dcoeffs[4 * t + 2], g2);
      atomicAdd(&dcoeffs[4 * t + 3], g3);
    }
  }
}

// out = a + b
__global__ void add_kernel(long long n, const float* __restrict__ a, const float* __restrict__ b,
                           float* __restrict__ out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = a[i] + b[i];
}

// layers.py:66-75: out = log(exp(a) + exp(b)) in fp64 (no max subtraction, like the reference);
// where use[h] == 0 the residual value passes through (models.py:376-380).
__global__ void counts_lse_kernel(int B, int H, const int* __restrict__ use,
                                  const float* __restrict__ bias_lc,
                                  const float* __restrict__ resid_lc, float* __restrict__ out,
                                  float* __restrict__ dresid_scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  const int h = i % H;
  if (use[h]) {
    const double ea = exp(static_cast<double>(bias_lc[i])), eb = exp(static_cast<double>(resid_lc[i]));
    out[i] = static_cast<float>(log(ea + eb));
    if (dresid_scale) dresid_scale[i] = static_cast<float>(eb / (ea + eb));
  } else {
    out[i] = resid_lc[i];
    if (dresid_scale) dresid_scale[i] = 1.f;
  }
}

__global__ void mul_kernel(long long n, const float* __restrict__ a, const float* __restrict__ b,
                           float* __restrict__ out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = a[i] * b[i];
}

static int ew_blocks(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = static_cast<long long>(sm_count() > 0 ? sm_count() : 1) * 8;
  return static_cast<int>(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace bp

using namespace bp;

extern "C" int bpb_loss_fwd_bwd(int B, int L, int T, int H, const int* task_off,
                                const float* logits, const float* logcounts, const float* counts,
                                const float* logcounts_true, const float* profile_weight,
                                const float* lambda, float* losses, float* dlogits,
                                float* dlogcounts, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(task_off && logits && logcounts && counts && profile_weight && lambda && losses,
             "bpb_loss_fwd_bwd: NULL argument");
  BP_REQUIRE(B > 0 && L > 0 && T > 0 && H > 0, "bpb_loss_fwd_bwd: bad sizes");
  BP_CHECK_CUDA(cudaMemsetAsync(losses, 0, static_cast<size_t>(2) * H * sizeof(float), stream));
  dim3 grid(B, H);
  loss_kernel<<<grid, 256, 0, stream>>>(B, L, T, H, task_off, logits, logcounts, counts,
                                        logcounts_true, profile_weight, lambda, losses, dlogits,
                                        dlogcounts);
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int bpb_adam_step(long long n, float* param, const float* grad, float* m, float* v,
                             const float* step_and_lr, float beta1, float beta2, float eps,
                             float grad_scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(param && grad && m && v && step_and_lr, "bpb_adam_step: NULL argument");
  if (n <= 0) return 0;
  adam_kernel<<<ew_blocks(n), 256, 0, stream>>>(n, param, grad, m, v, step_and_lr, beta1, beta2,
                                                eps, grad_scale);
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int bpb_prep_conv3_weights(int n_layers, int Fw, int F, const float* w, bpb_bf16* fwd,
                                      bpb_bf16* bwd, int precise, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(w && fwd && bwd, "bpb_prep_conv3_weights: NULL argument");
  BP_REQUIRE(F % 64 == 0 && Fw > 0 && Fw <= F && n_layers > 0,
             "bpb_prep_conv3_weights: bad F=%d Fw=%d layers=%d", F, Fw, n_layers);
  prep_conv3_kernel<<<ew_blocks(3LL * F * F * n_layers), 256, 0, stream>>>(
      n_layers, Fw, F, w, reinterpret_cast<__nv_bfloat16*>(fwd),
      reinterpret_cast<__nv_bfloat16*>(bwd), precise);
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int bpb_slide_u8(const uint8_t* src, uint8_t* dst, int rows, int src_cols, int dst_cols,
                            int depth, const int* row_idx, const int* col_idx, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(src && dst && row_idx && col_idx, "bpb_slide_u8: NULL argument");
  if (rows <= 0) return 0;
  slide_kernel<uint8_t><<<rows, 256, 0, stream>>>(src, dst, src_cols, dst_cols, depth, row_idx,
                                                  col_idx);
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int bpb_slide_f32(const float* src, float* dst, int rows, int src_cols, int dst_cols,
                             int depth, const int* row_idx, const int* col_idx, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(src && dst && row_idx && col_idx, "bpb_slide_f32: NULL argument");
  if (rows <= 0) return 0;
  slide_kernel<float><<<rows, 256, 0, stream>>>(src, dst, src_cols, dst_cols, depth, row_idx,
                                                col_idx);
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int bpb_log_row_sums(const float* vals, float* out, int rows, int n_per_row,
                                void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(vals && out, "bpb_log_row_sums: NULL argument");
  if (rows <= 0) return 0;
  log_row_sums_kernel<<<rows, 256, 0, stream>>>(vals, out, n_per_row);
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int bpb_transform_fwd(long long n, const float* u, float* y, int n_terms,
                                 const int* act, const float* coeffs, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(u && y && act && coeffs && n_terms > 0, "bpb_transform_fwd: bad argument");
  if (n <= 0) return 0;
  transform_fwd_kernel<<<ew_blocks(n), 256, 0, stream>>>(n, u, y, n_terms, act, coeffs);
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int bpb_transform_bwd(long long n, const float* u, const float* dy, int n_terms,
                                 const int* act, const float* coeffs, float* dcoeffs,
                                 void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(u && dy && act && coeffs && dcoeffs && n_terms > 0, "bpb_transform_bwd: bad argument");
  if (n <= 0) return 0;
  int blocks = ew_blocks(n);
  if (blocks > 64) blocks = 64;
  transform_bwd_kernel<<<blocks, 256, 0, stream>>>(n, u, dy, n_terms, act, coeffs, dcoeffs);
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int bpb_add_f32(long long n, const float* a, const float* b, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(a && b && out, "bpb_add_f32: NULL argument");
  if (n <= 0) return 0;
  add_kernel<<<ew_blocks(n), 256, 0, stream>>>(n, a, b, out);
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int bpb_mul_f32(long long n, const float* a, const float* b, float* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(a && b && out, "bpb_mul_f32: NULL argument");
  if (n <= 0) return 0;
  mul_kernel<<<ew_blocks(n), 256, 0, stream>>>(n, a, b, out);
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int bpb_counts_lse(int B, int H, const int* use_bias, const float* bias_lc,
                              const float* resid_lc, float* out, float* dresid_scale,
                              void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(use_bias && bias_lc && resid_lc && out, "bpb_counts_lse: NULL argument");
  counts_lse_kernel<<<(B * H + 127) / 128, 128, 0, stream>>>(B, H, use_bias, bias_lc, resid_lc,
                                                             out, dresid_scale);
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// DeepSHAP combiners.  standard (shap.py:26-30):  phi[q,t,c] = mean_r mult[q,r,t,c] (x - ref)
// hypothetical (interpretUtils.py:1295-1312):     phi[q,t,i] = mean_r sum_c (1[c==i] - ref_c) mult_c
// ------------------------------------------------------------------------------------------
namespace bp {
__global__ void shap_combine_kernel(int Q, int n, int L, const uint8_t* __restrict__ x,
                                    const uint8_t* __restrict__ refs,
                                    const float* __restrict__ mult, int hypothetical,
                                    float* __restrict__ phi) {
  const long long total = static_cast<long long>(Q) * L;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int q = static_cast<int>(i / L), t = static_cast<int>(i % L);
    const uchar4 xv = *reinterpret_cast<const uchar4*>(x + i * 4);
    const float xf[4] = {float(xv.x), float(xv.y), float(xv.z), float(xv.w)};
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < n; ++r) {
      const size_t row = (static_cast<size_t>(q) * n + r) * L + t;
      const uchar4 rv = *reinterpret_cast<const uchar4*>(refs + row * 4);
      const float4 m = *reinterpret_cast<const float4*>(mult + row * 4);
      const float rf[4] = {float(rv.x), float(rv.y), float(rv.z), float(rv.w)};
      const float mf[4] = {m.x, m.y, m.z, m.w};
      if (!hypothetical) {
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[c] += mf[c] * (xf[c] - rf[c]);
      } else {
        float base = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) base -= rf[c] * mf[c];
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[c] += base + mf[c];
      }
    }
    const float inv = 1.f / static_cast<float>(n);
    *reinterpret_cast<float4*>(phi + i * 4) =
        make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
  }
}
}  // namespace bp

extern "C" int bpb_shap_combine(int Q, int n, int L, const uint8_t* x, const uint8_t* refs,
                                const float* mult, int hypothetical, float* phi, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(x && refs && mult && phi && n > 0, "bpb_shap_combine: bad argument");
  if (Q <= 0) return 0;
  bp::shap_combine_kernel<<<bp::ew_blocks(static_cast<long long>(Q) * L), 256, 0, stream>>>(
      Q, n, L, x, refs, mult, hypothetical, phi);
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// refs[q*n + r, t, :] = x[q, perm[r, t], :]   (interpretUtils.py:1216-1222, kmer size 1)
namespace bp {
__global__ void shuffle_rows_kernel(int Q, int n, int L, const uint8_t* __restrict__ x,
                                    const int* __restrict__ perm, uint8_t* __restrict__ refs) {
  const long long total = static_cast<long long>(Q) * n * L;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i % L);
    const int r = static_cast<int>((i / L) % n);
    const int q = static_cast<int>(i / (static_cast<long long>(L) * n));
    const int src = perm[r * L + t];
    reinterpret_cast<uint32_t*>(refs)[i] =
        reinterpret_cast<const uint32_t*>(x)[static_cast<size_t>(q) * L + src];
  }
}
}  // namespace bp

extern "C" int bpb_shuffle_rows(int Q, int n, int L, const uint8_t* x, const int* perm,
                                uint8_t* refs, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(x && perm && refs && n > 0, "bpb_shuffle_rows: bad argument");
  if (Q <= 0) return 0;
  bp::shuffle_rows_kernel<<<bp::ew_blocks(static_cast<long long>(Q) * n * L), 256, 0, stream>>>(
      Q, n, L, x, perm, refs);
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// interpretFlat profile metric (interpretFlat.py:187-237) under DeepSHAP (shap.py:642-678).
//   l~ = logits[:, :, selected tasks] - mean over the (L x tasks) block
//   metric v = sum_j softmax(stop_gradient(l~))_j * l~_j
// The product of two varying tensors takes the 2-input rule: the multiplier of the query-half l~
// is (p_x + p_r) / 2, zero where |l~_x - l~_r| < 1e-7; the stop-gradient branch contributes
// nothing; the mean-normalisation is linear, so dlogits_j = m_j - mean(m) on the selected
// columns.  One block per (query, reference) pair.
// ------------------------------------------------------------------------------------------
namespace bp {
__global__ void __launch_bounds__(256)
flat_profile_seed_kernel(int n, int L, int T, const int* __restrict__ sel,
                         const float* __restrict__ lx_all, const float* __restrict__ lr_all,
                         float* __restrict__ dlogits, float* __restrict__ vx,
                         float* __restrict__ vr) {
  __shared__ float scratch[32];
  const int r = blockIdx.x;
  const int q = r / n;
  const float* lx = lx_all + static_cast<size_t>(q) * L * T;
  const float* lr = lr_all + static_cast<size_t>(r) * L * T;
  float* dl = dlogits + static_cast<size_t>(r) * L * T;
  const int N = L * T;
  float sx = 0.f, sr = 0.f, mxx = -INFINITY, mxr = -INFINITY;
  int cnt = 0;
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    if (sel[i % T]) {
      sx += lx[i]; sr += lr[i];
      mxx = fmaxf(mxx, lx[i]); mxr = fmaxf(mxr, lr[i]);
      ++cnt;
    }
  sx = block_reduce(sx, SumF(), scratch);
  sr = block_reduce(sr, SumF(), scratch);
  mxx = block_reduce(mxx, MaxF(), scratch);
  mxr = block_reduce(mxr, MaxF(), scratch);
  const float J = block_reduce(static_cast<float>(cnt), SumF(), scratch);
  const float mean_x = sx / J, mean_r = sr / J;
  float ex = 0.f, er = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    if (sel[i % T]) {
      ex += __expf(lx[i] - mxx);
      er += __expf(lr[i] - mxr);
    }
  ex = block_reduce(ex, SumF(), scratch);
  er = block_reduce(er, SumF(), scratch);
  float sm = 0.f, ax = 0.f, ar = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    if (sel[i % T]) {
      const float px = __expf(lx[i] - mxx) / ex, pr = __expf(lr[i] - mxr) / er;
      const float tx = lx[i] - mean_x, tr = lr[i] - mean_r;
      sm += (fabsf(tx - tr) < 1e-7f) ? 0.f : 0.5f * (px + pr);
      ax += px * tx;
      ar += pr * tr;
    }
  sm = block_reduce(sm, SumF(), scratch);
  ax = block_reduce(ax, SumF(), scratch);
  ar = block_reduce(ar, SumF(), scratch);
  const float mbar = sm / J;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    float g = 0.f;
    if (sel[i % T]) {
      const float px = __expf(lx[i] - mxx) / ex, pr = __expf(lr[i] - mxr) / er;
      const float tx = lx[i] - mean_x, tr = lr[i] - mean_r;
      g = ((fabsf(tx - tr) < 1e-7f) ? 0.f : 0.5f * (px + pr)) - mbar;
    }
    dl[i] = g;
  }
  if (threadIdx.x == 0) {
    if (vr) vr[r] = ar;
    if (vx && r % n == 0) vx[q] = ax;
  }
}
}  // namespace bp

extern "C" int bpb_flat_profile_seed(int Q, int n, int L, int T, const int* sel,
                                     const float* logits_x, const float* logits_r,
                                     float* dlogits, float* vx, float* vr, void* stream_) {
  BP_REQUIRE(sel && logits_x && logits_r && dlogits && Q > 0 && n > 0,
             "bpb_flat_profile_seed: bad argument");
  bp::flat_profile_seed_kernel<<<Q * n, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      n, L, T, sel, logits_x, logits_r, dlogits, vx, vr);
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}
