// K5 (fused log-softmax / multinomial NLL / log-counts MSE forward+backward), K9 (Keras Adam),
// weight re-layout for the tcgen05 kernels, the epoch jitter gather that replaces libslide, and
// the scalar regressions of the transformation / combined models.
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "ptx.cuh"

namespace bp {

// Block-wide reduction in a fixed order (deterministic): butterfly inside each warp, then the
// first warp combines the (up to 32) warp results with a second butterfly (lanes beyond the warp
// count contribute the identity `id`).
// Every thread returns the result.  (A loop over the warp results in EVERY thread costs a
// thousand shared-memory reads per call with 32 warps -- measured 1.7k cycles per reduction.)
template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, Op op, T* scratch, T id) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = (blockDim.x + 31) >> 5;
  __syncthreads();  // scratch may still be read by the previous call
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    T w = lane < nw ? scratch[lane] : id;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w = op(w, __shfl_xor_sync(0xffffffffu, w, o));
    if (lane == 0) scratch[0] = w;
  }
  __syncthreads();
  return scratch[0];
}

struct SumF { __device__ float operator()(float a, float b) const { return a + b; } };
struct MaxF { __device__ float operator()(float a, float b) const { return fmaxf(a, b); } };
struct SumD { __device__ double operator()(double a, double b) const { return a + b; } };

// ------------------------------------------------------------------------------------------
// K5.  losses.py:15-39 (tfp Multinomial.log_prob over the flattened L*T_h block of one head),
// losses.py:42-63 + Keras sum_over_batch_size.  One block per (sequence, head).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
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
  mx = block_reduce(mx, MaxF(), scratch_f, -INFINITY);
  sN = block_reduce(sN, SumD(), scratch_d, 0.0);
  sLg = block_reduce(sLg, SumD(), scratch_d, 0.0);
  sKl = block_reduce(sKl, SumD(), scratch_d, 0.0);
  float se = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int idx = (i / Th) * T + k0 + (i % Th);
    se += __expf(lg[idx] - mx);
  }
  se = block_reduce(se, SumF(), scratch_f, 0.f);
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
// K5 + packing, one launch: the losses of K5, and the loss gradient written straight into the
// packed bf16 tensor d8 that the heads' tensor-core kernels read (bpb_heads_pack_grad's layout),
// the heads' bias gradients and the loss sums reduced over the batch in a FIXED order by the last
// block to finish (no float atomics, no memset node).  One block per (sequence, head), one
// thread per profile row.
// ------------------------------------------------------------------------------------------
// (hi, lo) += (eh, el) with the rounding error of hi + eh kept in lo (Knuth two-sum; no fast-math
// reassociation is enabled for this library, so the compiler keeps the sequence as written)
__device__ __forceinline__ void two_sum_acc(float& hi, float& lo, float eh, float el) {
  const float t = __fadd_rn(hi, eh);
  const float bp = __fadd_rn(t, -hi);
  const float err = __fadd_rn(__fadd_rn(hi, -__fadd_rn(t, -bp)), __fadd_rn(eh, -bp));
  hi = t;
  lo = __fadd_rn(lo, __fadd_rn(err, el));
}

// Block-wide reduction of N per-thread float partials through shared memory, accumulated in
// double by ONE warp per quantity (lane j takes elements j, j+32, ... in four chains, a
// butterfly joins the lanes): fixed order, and no double-precision work in the 1024-thread part
// of the kernel.  Quantity 0 is a maximum when MAX0.  `buf` holds N * blockDim.x floats; the
// results land in res (double) and res_f (float), valid after the call for every thread.
template <int N, bool MAX0>
__device__ __forceinline__ void block_reduce_smem(const float (&v)[N], int n_used, float* buf,
                                                  double* res, float* res_f) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = blockDim.x;
  __syncthreads();  // buf / res may still be read by the previous call
#pragma unroll
  for (int i = 0; i < N; ++i)
    if (i < n_used) buf[i * nt + threadIdx.x] = v[i];
  __syncthreads();
  if (warp < n_used) {
    // The sums are carried as unevaluated float pairs (hi + lo, error-free two-sum): ~48
    // mantissa bits without touching the FP64 pipe, whose latency dominated this kernel when
    // the chains were double adds.
    const float* src = buf + warp * nt;
    if (MAX0 && warp == 0) {
      float m = -INFINITY;
      for (int j = lane; j < nt; j += 32) m = fmaxf(m, src[j]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (lane == 0) {
        res[0] = static_cast<double>(m);
        res_f[0] = m;
      }
    } else {
      float hi[4] = {0.f, 0.f, 0.f, 0.f}, lo[4] = {0.f, 0.f, 0.f, 0.f};
      for (int j = lane; j < nt; j += 128) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float e = (j + 32 * q < nt) ? src[j + 32 * q] : 0.f;
          two_sum_acc(hi[q], lo[q], e, 0.f);
        }
      }
      two_sum_acc(hi[0], lo[0], hi[1], lo[1]);
      two_sum_acc(hi[2], lo[2], hi[3], lo[3]);
      two_sum_acc(hi[0], lo[0], hi[2], lo[2]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float eh = __shfl_xor_sync(0xffffffffu, hi[0], o);
        const float el = __shfl_xor_sync(0xffffffffu, lo[0], o);
        two_sum_acc(hi[0], lo[0], eh, el);
      }
      if (lane == 0) {
        res[warp] = static_cast<double>(hi[0]) + static_cast<double>(lo[0]);
        res_f[warp] = hi[0] + lo[0];
      }
    }
  }
  __syncthreads();
}

constexpr int kLossMaxTh = 8;  // tasks per head handled by the fused kernel

// lgamma for the total count of a region: Stirling's series from x = 16 up (next term < 1e-16),
// the library routine below (it is several thousand cycles of one thread on the critical path)
__device__ __forceinline__ double lgamma_total(double x) {
  if (x < 16.0) return lgamma(x);
  const double xi = 1.0 / x, x2 = xi * xi;
  const double corr = xi * (1.0 / 12.0 - x2 * (1.0 / 360.0 - x2 * (1.0 / 1260.0 - x2 * (1.0 / 1680.0 -
                      x2 * (1.0 / 1188.0)))));
  return (x - 0.5) * log(x) - x + 0.91893853320467274178 + corr;
}

struct LossPack {
  int L_in, W, Cp, planes;
  long long plane_stride;  // elements between the bf16 remainder planes of d8
  __nv_bfloat16* d8;
  float *dpb, *dcb;
  float* ws;  // [B][2H] loss terms | [B][T] profile-bias terms | [B][H] dlogcounts | counter
  // optional: the heads' forward left per-tile partial sums of the counts columns (bpb_heads_fwd_tc
  // with logcounts == NULL); this kernel then finishes log-counts itself (and stores them)
  const float* cpart;  // [B][tiles][H]
  const float* cb;     // [H]
  int tiles;
  float* logcounts_out;
  unsigned* dbg;
};

__global__ void __launch_bounds__(1024)
loss_pack_kernel(int B, int L, int T, int H, const int* __restrict__ task_off,
                 const float* __restrict__ logits, const float* __restrict__ logcounts,
                 const float* __restrict__ counts, const float* __restrict__ logcounts_true,
                 const float* __restrict__ profile_weight, const float* __restrict__ lambda,
                 float* __restrict__ losses, float* __restrict__ dlogits,
                 float* __restrict__ dlogcounts, const LossPack pk) {
  __shared__ float red_buf[kLossMaxTh * 1024];
  __shared__ double red_res[kLossMaxTh];
  __shared__ float red_res_f[kLossMaxTh];
  __shared__ float s_dlc;
  __shared__ int s_last;
  const int b = blockIdx.x, h = blockIdx.y;
  pdl_launch_dependents();
  const int k0 = task_off[h];
  const int Th = task_off[h + 1] - k0;
  // scalars of this (sequence, head): fetched now so that their latency hides behind pass 1
  const float w_prof = profile_weight[h], lam_h = lambda[h];
  pdl_wait();  // logits and per-tile count sums are the predecessor's output
  float lc_pred;
  if (pk.cpart != nullptr) {
    // the same sum, in the same order, as heads_counts_finish_kernel
    float cs = 0.f;
    for (int tIdx = 0; tIdx < pk.tiles; ++tIdx)
      cs += pk.cpart[(static_cast<size_t>(b) * pk.tiles + tIdx) * H + h];
    lc_pred = pk.cb[h] + cs / static_cast<float>(pk.L_in);
    if (threadIdx.x == 0) pk.logcounts_out[b * H + h] = lc_pred;
  } else {
    lc_pred = logcounts[b * H + h];
  }
  const float lc_true = logcounts_true ? logcounts_true[b * H + h] : 0.f;
  const float* lg = logits + static_cast<size_t>(b) * L * T + k0;
  const float* ct = counts + static_cast<size_t>(b) * L * T + k0;
  float* part_loss = pk.ws;
  float* part_pb = part_loss + static_cast<size_t>(B) * 2 * H;
  float* part_dlc = part_pb + static_cast<size_t>(B) * T;
  unsigned* counter = reinterpret_cast<unsigned*>(part_dlc + static_cast<size_t>(B) * H);
#ifdef BPB_ROLE_PROFILE
  const long long t_start = clock64();
  __shared__ unsigned lp_st[16];
#define LP_STAMP(slot) \
  if (threadIdx.x == 0) lp_st[slot] = static_cast<unsigned>(clock64() - t_start)
#else
#define LP_STAMP(slot)
#endif

  // the thread's first row stays in registers for the three passes (L <= 1024: its only row)
  float l0[kLossMaxTh], c0[kLossMaxTh];
  const int tid = threadIdx.x;
#pragma unroll
  for (int c = 0; c < kLossMaxTh; ++c) {
    const bool on = c < Th && tid < L;
    l0[c] = on ? lg[tid * T + c] : 0.f;
    c0[c] = on ? ct[tid * T + c] : 0.f;
  }
  float mx = -INFINITY;
  // per-thread partials in float (a thread owns L / 1024 rows), batch sums in double
  float sN = 0.f, sLg = 0.f, sKl = 0.f;
  for (int t = tid; t < L; t += blockDim.x) {
#pragma unroll
    for (int c = 0; c < kLossMaxTh; ++c)
      if (c < Th) {
        const float l = (t == tid) ? l0[c] : lg[t * T + c];
        const float k = (t == tid) ? c0[c] : ct[t * T + c];
        mx = fmaxf(mx, l);
        sN += k;
        sLg += lgammaf(k + 1.f);
        sKl = fmaf(k, l, sKl);
      }
  }
  LP_STAMP(1);
  double dN, dLg, dKl;
  {
    const float v4[4] = {mx, sN, sLg, sKl};
    block_reduce_smem<4, true>(v4, 4, red_buf, red_res, red_res_f);
    mx = red_res_f[0];
    dN = red_res[1]; dLg = red_res[2]; dKl = red_res[3];  // (used by thread 0 only)
  }
  const float N = red_res_f[1];
  LP_STAMP(2);
  float se = 0.f;
  for (int t = tid; t < L; t += blockDim.x) {
#pragma unroll
    for (int c = 0; c < kLossMaxTh; ++c)
      if (c < Th) se += __expf(((t == tid) ? l0[c] : lg[t * T + c]) - mx);
  }
  {
    const float v1[1] = {se};
    block_reduce_smem<1, false>(v1, 1, red_buf, red_res, red_res_f);
    se = red_res_f[0];
  }
  LP_STAMP(3);
  const float lse = mx + logf(se);
  const float invB = 1.f / static_cast<float>(B);
  if (threadIdx.x == 0) {
    const double logp = lgamma_total(dN + 1.0) - dLg + dKl - dN * static_cast<double>(lse);
    const float y = logcounts_true ? lc_true : logf(N);
    const float d = lc_pred - y;
    const float dlc = 2.f * lam_h * d * invB;
    part_loss[(static_cast<size_t>(b) * 2) * H + h] = static_cast<float>(-logp) * invB;
    part_loss[(static_cast<size_t>(b) * 2 + 1) * H + h] = lam_h * d * d * invB;
    part_dlc[b * H + h] = dlc;
    if (dlogcounts) dlogcounts[b * H + h] = dlc;
    s_dlc = dlc;
  }
  __syncthreads();
  LP_STAMP(4);

  // gradient wrt the logits: fp32 copy (optional), packed bf16 planes, per-task sums
  const int Lp = pk.L_in + pk.W - 1;
  const float wh = w_prof * invB;
  float csum[kLossMaxTh];
#pragma unroll
  for (int c = 0; c < kLossMaxTh; ++c) csum[c] = 0.f;
  for (int t = threadIdx.x; t < L; t += blockDim.x) {
    const long long row = (static_cast<long long>(b) * Lp + t + pk.W - 1) * pk.Cp + k0;
#pragma unroll
    for (int c = 0; c < kLossMaxTh; ++c)
      if (c < Th) {
        const float l = (t == tid) ? l0[c] : lg[t * T + c];
        const float k = (t == tid) ? c0[c] : ct[t * T + c];
        const float dl = wh * (N * __expf(l - lse) - k);
        if (dlogits) dlogits[(static_cast<size_t>(b) * L + t) * T + k0 + c] = dl;
        csum[c] += dl;
        float r = dl;
        for (int pl = 0; pl < pk.planes; ++pl) {
          const __nv_bfloat16 hv = __float2bfloat16_rn(r);
          pk.d8[pl * pk.plane_stride + row + c] = hv;
          r -= __bfloat162float(hv);
        }
      }
  }
  LP_STAMP(5);
  // the counts gradient rides on every row of its own channel (tap 0 of the heads' data gradient)
  {
    const float v = s_dlc / static_cast<float>(pk.L_in);
    for (int r0 = threadIdx.x; r0 < Lp; r0 += blockDim.x) {
      const long long idx = (static_cast<long long>(b) * Lp + r0) * pk.Cp + T + h;
      float r = v;
      for (int pl = 0; pl < pk.planes; ++pl) {
        const __nv_bfloat16 hv = __float2bfloat16_rn(r);
        pk.d8[pl * pk.plane_stride + idx] = hv;
        r -= __bfloat162float(hv);
      }
    }
  }
  LP_STAMP(6);
  {
    block_reduce_smem<kLossMaxTh, false>(csum, Th, red_buf, red_res, red_res_f);
    if (threadIdx.x < Th) part_pb[static_cast<size_t>(b) * T + k0 + threadIdx.x] = red_res_f[threadIdx.x];
  }

  // last block to finish adds the per-sequence terms in sequence order
  LP_STAMP(7);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned ticket = atomicAdd(counter, 1u);
    s_last = (ticket == gridDim.x * gridDim.y - 1u);
  }
  __syncthreads();
  LP_STAMP(8);
  if (s_last) {
    __threadfence();
    // one warp per output (round-robin): lane j adds sequences j, j+32, ... then a butterfly --
    // a fixed order, with the B loads of an output in flight together
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int o = warp; o < 2 * H + T + H; o += nw) {
      const float* src;
      int stride, col;
      if (o < 2 * H) { src = part_loss; stride = 2 * H; col = o; }
      else if (o < 2 * H + T) { src = part_pb; stride = T; col = o - 2 * H; }
      else { src = part_dlc; stride = H; col = o - 2 * H - T; }
      float tot = 0.f;
      for (int bb = lane; bb < B; bb += 32) tot += __ldcg(src + static_cast<size_t>(bb) * stride + col);
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, sh);
      if (lane == 0) {
        if (o < 2 * H) losses[o] = tot;
        else if (o < 2 * H + T) { if (pk.dpb) pk.dpb[col] = tot; }
        else if (pk.dcb) pk.dcb[col] = tot;
      }
    }
    if (threadIdx.x == 0) *counter = 0u;  // ready for the next launch
    LP_STAMP(9);
  }
#ifdef BPB_ROLE_PROFILE
  if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && pk.dbg)
    for (int i = 1; i < 9; ++i) atomicAdd(pk.dbg + 1024 + i, lp_st[i]);
  if (threadIdx.x == 0 && s_last && pk.dbg) atomicAdd(pk.dbg + 1024 + 9, lp_st[9] - lp_st[8]);
#endif
}

// ------------------------------------------------------------------------------------------
// K9.  Keras-3 Adam on a flat bucket.
// ------------------------------------------------------------------------------------------
__global__ void adam_kernel(long long n, float* __restrict__ p, const float* __restrict__ g,
                            float* __restrict__ m, float* __restrict__ v,
                            const float* __restrict__ step_and_lr, float b1, float b2, float eps,
                            float gscale) {
  // bias-corrected step size: double-precision pow is long and slow, one thread per block does it
  __shared__ float s_alpha;
  if (threadIdx.x == 0) {
    const double step = static_cast<double>(step_and_lr[0]);
    s_alpha = static_cast<float>(static_cast<double>(step_and_lr[1]) *
                                 sqrt(1.0 - pow((double)b2, step)) / (1.0 - pow((double)b1, step)));
  }
  __syncthreads();
  const float alpha = s_alpha;
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


// ------------------------------------------------------------------------------------------
// K9 + operand refresh, one launch: the Adam update of every parameter, and each updated weight
// scattered straight into the bf16 operand layouts the tensor-core kernels read (what
// bpb_prep_conv3_weights / bpb_prep_input_conv_weights / bpb_prep_heads_dgrad_weights /
// bpb_prep_heads_fwd_weights produce: same formulas, bit-equal results).  The padding entries of
// those layouts are never written: they stay zero from allocation.  The step counter lives in
// state[0] and is advanced by the last block to finish (state[2] is its ticket counter).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void scatter_planes(__nv_bfloat16* dst, long long plane_stride,
                                               int planes, long long idx, float v) {
  float r = v;
  for (int pl = 0; pl < planes; ++pl) {
    const __nv_bfloat16 h = __float2bfloat16_rn(r);
    dst[pl * plane_stride + idx] = h;
    r -= __bfloat162float(h);
  }
}

// Data-parallel variant (PEER): the gradient all-reduce is part of this kernel.  Every rank's
// gradient bucket and a small flag block live in memory its peers map over NVLink (CUDA IPC);
// the kernel announces "my gradients are complete" in every peer's flag block, waits until all
// peers have announced theirs, then each thread adds the world's values of its element straight
// out of peer memory in rank order (a one-shot all-reduce: identical sums, hence identical
// parameters, on every rank) and applies the update; the last block to finish tells every peer
// that this rank no longer reads its bucket.  Flag block of a rank: words 0..7 ready[r],
// 8..15 done[r], 16 sequence number of the last completed exchange, 17 wait bound in seconds
// (0 = 30 s, set by the host), 18 error word.  Waits are bounded in WALL time (%globaltimer, so
// the bound does not move with the SM clock): a peer that has not arrived when the bound expires
// -- ranks legitimately diverge by seconds around checkpoints, validation or a first-step graph
// capture -- makes the waiting thread record 1 + the missing rank in the error word and go on
// (with that peer's stale bucket), instead of trapping and leaving a sticky fault in the CUDA
// context; the host reads the error word (parallel.PeerGradExchange.check) and raises.
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void peer_spin_until(const unsigned* flag, unsigned want, unsigned* mine,
                                                unsigned who) {
  if (static_cast<int>(ld_acquire_sys(flag) - want) >= 0) return;
  const unsigned long long t0 = globaltimer_ns();
  const unsigned secs = mine[17] ? mine[17] : 30u;
  const unsigned long long bound = static_cast<unsigned long long>(secs) * 1000000000ULL;
  unsigned spins = 0;
  while (static_cast<int>(ld_acquire_sys(flag) - want) < 0) {
    __nanosleep(40);
    if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > bound) {
      atomicMax(mine + 18, 1u + who);  // a peer is missing: tell the host, do not trap
      return;
    }
  }
}

__global__ void peer_wait_kernel(const bpb_peer_args pa) {
  // the previous exchange must be over on every peer before this rank's gradients change again
  unsigned* mine = pa.flags[pa.rank];
  if (threadIdx.x < static_cast<unsigned>(pa.world))
    peer_spin_until(mine + 8 + threadIdx.x, mine[16], mine, threadIdx.x);
}

template <bool PEER>
__global__ void __launch_bounds__(256)
adam_prep_kernel(long long n, float* __restrict__ p, const float* __restrict__ g,
                 float* __restrict__ m, float* __restrict__ v, float* __restrict__ state, float b1,
                 float b2, float eps, float gscale, const bpb_prep_table tb,
                 const bpb_peer_args pa) {
  unsigned seq = 0;
  if (PEER) {
    unsigned* mine = pa.flags[pa.rank];
    seq = mine[16] + 1u;
    if (blockIdx.x == 0 && threadIdx.x < static_cast<unsigned>(pa.world)) {
      __threadfence_system();
      st_release_sys(pa.flags[threadIdx.x] + pa.rank, seq);  // ready[rank] on peer threadIdx.x
    }
    if (threadIdx.x < static_cast<unsigned>(pa.world))
      peer_spin_until(mine + threadIdx.x, seq, mine, threadIdx.x);
    __syncthreads();
  }
  // bias-corrected step size: double-precision pow is long and slow, one thread per block does it
  // -- and before the dependency wait: `state` is written only by this kernel's own previous
  // launch (long complete), so the two pows overlap the tail of the weight-gradient reduction
  __shared__ float s_alpha;
  const float step_f = state[0] + 1.f;
  if (threadIdx.x == 0) {
    const double step = static_cast<double>(step_f);
    s_alpha = static_cast<float>(static_cast<double>(state[1]) * sqrt(1.0 - pow((double)b2, step)) /
                                 (1.0 - pow((double)b1, step)));
  }
  pdl_wait();  // gradients of the predecessor
  __syncthreads();
  const float alpha = s_alpha;
  const long long F = tb.F;
  const long long per = 3 * F * F;
  const long long n_dil = per * tb.n_layers;
  const long long n_c1 = static_cast<long long>(tb.W_in) * 4 * tb.Fw;
  const long long n_pw = static_cast<long long>(tb.W_out) * tb.Fw * tb.T;
  const long long n_cw = static_cast<long long>(tb.Fw) * tb.H;
  const long long K_in = (tb.W_in * 8 + 63) / 64 * 64;
  const long long K_hd = (tb.W_out * tb.Cp + 63) / 64 * 64;
  const int KC = tb.F >> 6;
  __nv_bfloat16* w_fwd = reinterpret_cast<__nv_bfloat16*>(tb.w_fwd);
  __nv_bfloat16* w_bwd = reinterpret_cast<__nv_bfloat16*>(tb.w_bwd);
  __nv_bfloat16* w_in_op = reinterpret_cast<__nv_bfloat16*>(tb.w_in_op);
  __nv_bfloat16* w_hd_op = reinterpret_cast<__nv_bfloat16*>(tb.w_hd_op);
  __nv_bfloat16* w_hf_op = reinterpret_cast<__nv_bfloat16*>(tb.w_hf_op);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float gsum;
    if (PEER) {
      gsum = 0.f;
      for (int rk = 0; rk < pa.world; ++rk) gsum += __ldcv(pa.grads[rk] + i);  // uncached peer reads
    } else {
      gsum = g[i];
    }
    const float gi = gsum * gscale;
    const float mi = m[i] + (gi - m[i]) * (1.f - b1);
    const float vi = v[i] + (gi * gi - v[i]) * (1.f - b2);
    m[i] = mi;
    v[i] = vi;
    const float w = p[i] - alpha * mi / (sqrtf(vi) + eps);
    p[i] = w;
    long long r = i - tb.off_dil_w;
    if (r >= 0 && r < n_dil) {
      // 32-bit index arithmetic (the host checks n < 2^31): 64-bit divisions were most of this
      // kernel's instructions
      const unsigned r32 = static_cast<unsigned>(r), per32 = static_cast<unsigned>(per),
                     F32 = static_cast<unsigned>(F);
      const unsigned l = r32 / per32, rr = r32 - l * per32;
      const unsigned jc = rr / F32, cout = rr - jc * F32;  // jc = j * F + cin
      const unsigned j = jc / F32, cin = jc - j * F32;
      const long long base = static_cast<long long>(l) * tb.planes * per;
      scatter_planes(w_fwd, per, tb.planes, base + (j * F + cout) * F + cin, w);
      scatter_planes(w_bwd, per, tb.planes, base + (j * F + cin) * F + cout, w);
      continue;
    }
    r = i - tb.off_conv1_w;
    if (r >= 0 && r < n_c1) {
      const unsigned r32 = static_cast<unsigned>(r), Fw32 = static_cast<unsigned>(tb.Fw);
      const unsigned jc = r32 / Fw32, nn = r32 - jc * Fw32;  // jc = j * 4 + c4
      const unsigned j = jc >> 2, c4 = jc & 3u;
      scatter_planes(w_in_op, F * K_in, tb.in_planes, nn * K_in + j * 8 + c4, w);
      continue;
    }
    r = i - tb.off_prof_w;
    int j = -1, c = 0, col = 0;
    if (r >= 0 && r < n_pw) {
      const unsigned r32 = static_cast<unsigned>(r), T32 = static_cast<unsigned>(tb.T),
                     Fw32 = static_cast<unsigned>(tb.Fw);
      const unsigned jc = r32 / T32;  // j * Fw + c
      col = static_cast<int>(r32 - jc * T32);
      j = static_cast<int>(jc / Fw32);
      c = static_cast<int>(jc - static_cast<unsigned>(j) * Fw32);
    } else {
      r = i - tb.off_cnt_w;
      if (r >= 0 && r < n_cw) {
        const unsigned r32 = static_cast<unsigned>(r), H32 = static_cast<unsigned>(tb.H);
        j = 0;
        c = static_cast<int>(r32 / H32);
        col = tb.T + static_cast<int>(r32 - static_cast<unsigned>(c) * H32);
      }
    }
    if (j >= 0) {
      // heads' data-gradient operand [F][K_hd], k = (W_out-1-j)*Cp + col (counts ride on tap 0,
      // which is j' = 0 there as well)
      const long long jp = (col < tb.T) ? (tb.W_out - 1 - j) : 0;
      scatter_planes(w_hd_op, F * K_hd, tb.planes, c * K_hd + jp * tb.Cp + col, w);
      if (w_hf_op != nullptr) {
        const int cc = c & 63;
        const long long in_tile = static_cast<long long>(col) * 64 + (((cc >> 3) ^ (col & 7)) << 3) + (cc & 7);
        float rem = w;
        for (int pl = 0; pl < tb.planes; ++pl) {
          const __nv_bfloat16 hb = __float2bfloat16_rn(rem);
          // grouped image of bpb_prep_heads_fwd_weights: (group, plane, tap, chunk in group)
          const int KCg = tb.hf_kcg > 0 ? tb.hf_kcg : KC;
          const long long tile =
              ((static_cast<long long>((c >> 6) / KCg) * tb.planes + pl) * tb.W_out + j) * KCg +
              ((c >> 6) % KCg);
          w_hf_op[tile * (static_cast<long long>(tb.N_heads) * 64) + in_tile] = hb;
          rem -= __bfloat162float(hb);
        }
      }
    }
  }
  // advance the step counter once every block has read it
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned* ticket = reinterpret_cast<unsigned*>(state + 2);
    if (atomicAdd(ticket, 1u) == gridDim.x - 1u) {
      *ticket = 0u;
      state[0] = step_f;
      if (PEER) {
        // every block of this rank has read the peers' buckets: release them
        unsigned* mine = pa.flags[pa.rank];
        mine[16] = seq;
        __threadfence_system();
        for (int rk = 0; rk < pa.world; ++rk) st_release_sys(pa.flags[rk] + 8 + pa.rank, seq);
      }
    }
  }
}


// ------------------------------------------------------------------------------------------
// Prediction-side helpers.
// utils.py:420-486 oneHotEncode: ASCII bases -> [n][4] uint8, columns A C G T, upper or lower
// case, anything else all-zero (allowN) -- the caller learns how many bytes were not ACGT.
// utils.py:523-572 logitsToProfile: softmax over the L x T_h block of one head times
// exp(logcounts), one block per (sequence, head).
// ------------------------------------------------------------------------------------------
__global__ void onehot_encode_kernel(long long n, const uint8_t* __restrict__ seq,
                                     uint32_t* __restrict__ out, unsigned long long* n_other) {
  unsigned other = 0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint8_t c = seq[i] & 0xDF;  // fold case
    uint32_t w = 0;
    if (c == 'A') w = 0x00000001u;
    else if (c == 'C') w = 0x00000100u;
    else if (c == 'G') w = 0x00010000u;
    else if (c == 'T') w = 0x01000000u;
    else ++other;
    out[i] = w;  // four uint8 {A, C, G, T}, little endian
  }
  if (n_other != nullptr && other) atomicAdd(n_other, static_cast<unsigned long long>(other));
}

__global__ void __launch_bounds__(256)
logits_to_profile_kernel(int L, int T, int H, const int* __restrict__ task_off,
                         const float* __restrict__ logits, const float* __restrict__ logcounts,
                         float* __restrict__ profile) {
  __shared__ float scratch[32];
  const int b = blockIdx.x, h = blockIdx.y;
  const int k0 = task_off[h], Th = task_off[h + 1] - k0;
  const float* lg = logits + static_cast<size_t>(b) * L * T + k0;
  float* pr = profile + static_cast<size_t>(b) * L * T + k0;
  const int n = L * Th;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, lg[(i / Th) * T + (i % Th)]);
  mx = block_reduce(mx, MaxF(), scratch, -INFINITY);
  float se = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) se += expf(lg[(i / Th) * T + (i % Th)] - mx);
  se = block_reduce(se, SumF(), scratch, 0.f);
  const float scale = expf(logcounts[b * H + h]) / se;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int idx = (i / Th) * T + (i % Th);
    pr[idx] = expf(lg[idx] - mx) * scale;
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
  s = block_reduce(s, SumF(), scratch, 0.f);
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
    g0 = block_reduce(g0, SumF(), scratch, 0.f);
    g1 = block_reduce(g1, SumF(), scratch, 0.f);
    g2 = block_reduce(g2, SumF(), scratch, 0.f);
    g3 = block_reduce(g3, SumF(), scratch, 0.f);
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

// models.py:349-380 in one launch: the frozen bias model's raw outputs go through the per-head
// transformation (models.py:118-175; coefficient table [H][n_p + n_c][4], profile terms first),
// profile logits are added to the residual's, log-counts are combined by CountsLogSumExp (fp64)
// where use[h], else the residual's pass through.  Thread i < B*L*T handles one logit, the next
// B*H threads one log-count.
__global__ void combine_bias_kernel(int B, int L, int T, int H, const int* __restrict__ task_off,
                                    const float* __restrict__ bias_logits,
                                    const float* __restrict__ bias_lc, int n_p,
                                    const int* __restrict__ p_act, int n_c,
                                    const int* __restrict__ c_act,
                                    const float* __restrict__ coeffs, const int* __restrict__ use,
                                    const float* __restrict__ resid_logits,
                                    const float* __restrict__ resid_lc,
                                    float* __restrict__ out_logits, float* __restrict__ out_lc,
                                    float* __restrict__ dresid_scale,
                                    float* __restrict__ bias_logits_t,
                                    float* __restrict__ bias_lc_t) {
  const long long nl = static_cast<long long>(B) * L * T;
  const long long total = nl + static_cast<long long>(B) * H;
  const int per = n_p + n_c;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    if (i < nl) {
      const int col = static_cast<int>(i % T);
      int h = 0;
      while (h + 1 < H && task_off[h + 1] <= col) ++h;
      const float x = bias_logits[i];
      float s = x;
      if (n_p > 0) {
        s = 0.f;
        const float* c = coeffs + static_cast<long long>(h) * per * 4;
        for (int t = 0; t < n_p; ++t) {
          const float m1 = c[4 * t], b1 = c[4 * t + 1], m2 = c[4 * t + 2], b2 = c[4 * t + 3];
          s += (p_act[t] == 0) ? m1 * x + b1 : m1 * act_fn(p_act[t], m2 * x + b2) + b1;
        }
      }
      if (bias_logits_t) bias_logits_t[i] = s;
      out_logits[i] = s + resid_logits[i];
    } else {
      const long long j = i - nl;
      const int h = static_cast<int>(j % H);
      const float x = bias_lc[j];
      float s = x;
      if (n_c > 0) {
        s = 0.f;
        const float* c = coeffs + (static_cast<long long>(h) * per + n_p) * 4;
        for (int t = 0; t < n_c; ++t) {
          const float m1 = c[4 * t], b1 = c[4 * t + 1], m2 = c[4 * t + 2], b2 = c[4 * t + 3];
          s += (c_act[t] == 0) ? m1 * x + b1 : m1 * act_fn(c_act[t], m2 * x + b2) + b1;
        }
      }
      if (bias_lc_t) bias_lc_t[j] = s;
      if (use[h]) {
        const double ea = exp(static_cast<double>(s)), eb = exp(static_cast<double>(resid_lc[j]));
        out_lc[j] = static_cast<float>(log(ea + eb));
        if (dresid_scale) dresid_scale[j] = static_cast<float>(eb / (ea + eb));
      } else {
        out_lc[j] = resid_lc[j];
        if (dresid_scale) dresid_scale[j] = 1.f;
      }
    }
  }
}

// DeepSHAP through the output stage of a transformation / combined model (shap.py:612-639 applied
// to models.py:118-175 and 349-380): the rescale rule replaces the local derivative of every
// Sigmoid / Relu / Exp / Log by (out_x - out_r) / (in_x - in_r) of the (query, reference) pair --
// the ordinary derivative at the query where |in_x - in_r| < 1e-6 (shap.py:634-638) -- while
// linear regressions, adds and casts keep their ordinary gradients (shap.py:737-779).  Given the
// gradient of the metric wrt the COMBINED outputs of every pair, this writes the gradients wrt
// the residual network's outputs and wrt the bias network's RAW outputs, i.e. the seeds of the two
// rescale-rule backward passes.  Pair i uses query i / n.
__device__ __forceinline__ double rescale_term(int a, double m1, double m2, double b2, double ux,
                                               double ur) {
  if (a == 0) return m1;  // linear regression: ordinary gradient
  const double zx = m2 * ux + b2, zr = m2 * ur + b2;
  double fx, fr, dfx;
  if (a == 1) {
    fx = 1.0 / (1.0 + exp(-zx));
    fr = 1.0 / (1.0 + exp(-zr));
    dfx = fx * (1.0 - fx);
  } else {
    fx = zx > 0.0 ? zx : 0.0;
    fr = zr > 0.0 ? zr : 0.0;
    dfx = zx > 0.0 ? 1.0 : 0.0;
  }
  const double din = zx - zr;
  return m1 * (fabs(din) < 1e-6 ? dfx : (fx - fr) / din) * m2;
}

__device__ __forceinline__ double transform_value(int n_t, const int* act, const float* c, double u) {
  if (n_t == 0) return u;
  double s = 0.0;
  for (int t = 0; t < n_t; ++t) {
    const double m1 = c[4 * t], b1 = c[4 * t + 1], m2 = c[4 * t + 2], b2 = c[4 * t + 3];
    const double z = m2 * u + b2;
    s += b1 + m1 * (act[t] == 0 ? u : (act[t] == 1 ? 1.0 / (1.0 + exp(-z)) : (z > 0.0 ? z : 0.0)));
  }
  return s;
}

__global__ void combined_shap_seed_kernel(
    int R, int n, int L, int T, int H, const int* __restrict__ task_off,
    const float* __restrict__ bl_x, const float* __restrict__ bl_r,   // bias raw logits
    const float* __restrict__ bc_x, const float* __restrict__ bc_r,   // bias raw log-counts
    const float* __restrict__ rc_x, const float* __restrict__ rc_r,   // residual log-counts
    int n_p, const int* __restrict__ p_act, int n_c, const int* __restrict__ c_act,
    const float* __restrict__ coeffs, const int* __restrict__ use,
    const float* __restrict__ dl_c, const float* __restrict__ dc_c,   // wrt combined outputs
    float* __restrict__ dl_r, float* __restrict__ dc_r, float* __restrict__ dl_b,
    float* __restrict__ dc_b) {
  const long long nl = static_cast<long long>(R) * L * T;
  const long long total = nl + static_cast<long long>(R) * H;
  const int per = n_p + n_c;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    if (i < nl) {
      const long long pair = i / (static_cast<long long>(L) * T);
      const long long within = i - pair * L * T;
      const int col = static_cast<int>(within % T);
      int h = 0;
      while (h + 1 < H && task_off[h + 1] <= col) ++h;
      const float g = dl_c[i];
      if (dl_r) dl_r[i] = g;  // Add: ordinary gradient
      if (dl_b) {
        double m = 1.0;
        if (n_p > 0) {
          m = 0.0;
          const float* c = coeffs + static_cast<long long>(h) * per * 4;
          const double ux = bl_x[(pair / n) * L * T + within], ur = bl_r[i];
          for (int t = 0; t < n_p; ++t)
            m += rescale_term(p_act[t], c[4 * t], c[4 * t + 2], c[4 * t + 3], ux, ur);
        }
        dl_b[i] = static_cast<float>(g * m);
      }
    } else {
      const long long j = i - nl;
      const long long pair = j / H;
      const int h = static_cast<int>(j % H);
      const long long jx = (pair / n) * H + h;
      const double g = dc_c[j];
      const float* c = coeffs + (static_cast<long long>(h) * per + n_p) * 4;
      double mt = 1.0;  // transformation multiplier of the bias log-counts
      if (n_c > 0) {
        mt = 0.0;
        for (int t = 0; t < n_c; ++t)
          mt += rescale_term(c_act[t], c[4 * t], c[4 * t + 2], c[4 * t + 3], bc_x[jx], bc_r[j]);
      }
      if (rc_x == nullptr) {  // transformation model alone: no residual branch
        if (dc_b) dc_b[j] = static_cast<float>(g * mt);
        continue;
      }
      if (!use[h]) {  // Identity on the residual's log-counts (models.py:376-380)
        if (dc_r) dc_r[j] = static_cast<float>(g);
        if (dc_b) dc_b[j] = 0.f;
        continue;
      }
      // CountsLogSumExp (layers.py:66-75, fp64): Exp on each branch, Add, Log
      const double ax = transform_value(n_c, c_act, c, bc_x[jx]);
      const double ar = transform_value(n_c, c_act, c, bc_r[j]);
      const double bx = rc_x[jx], br = rc_r[j];
      const double eax = exp(ax), ear = exp(ar), ebx = exp(bx), ebr = exp(br);
      const double sx = eax + ebx, sr = ear + ebr;
      const double m_log = fabs(sx - sr) < 1e-6 ? 1.0 / sx : (log(sx) - log(sr)) / (sx - sr);
      const double m_ea = fabs(ax - ar) < 1e-6 ? eax : (eax - ear) / (ax - ar);
      const double m_eb = fabs(bx - br) < 1e-6 ? ebx : (ebx - ebr) / (bx - br);
      if (dc_r) dc_r[j] = static_cast<float>(g * m_log * m_eb);
      if (dc_b) dc_b[j] = static_cast<float>(g * m_log * m_ea * mt);
    }
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
  // one block per (sequence, head) is all the parallelism the reductions allow, so the block is
  // as wide as it can be: the kernel is bound by the latency of its serial passes
  loss_kernel<<<grid, 1024, 0, stream>>>(B, L, T, H, task_off, logits, logcounts, counts,
                                        logcounts_true, profile_weight, lambda, losses, dlogits,
                                        dlogcounts);
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}


extern "C" long long bpb_loss_pack_ws_elems(int B, int T, int H) {
  return static_cast<long long>(B) * (2 * H + T + H) + 4;
}

extern "C" int bpb_loss_fwd_bwd_pack(int B, int L, int T, int H, const int* task_off,
                                     const float* logits, const float* logcounts,
                                     const float* counts, const float* logcounts_true,
                                     const float* profile_weight, const float* lambda,
                                     float* losses, float* dlogits, float* dlogcounts, int L_in,
                                     int W, int Cp, int planes, bpb_bf16* d8, float* dpb,
                                     float* dcb, float* ws, const int* host_task_off,
                                     const float* counts_partials, int tiles, const float* cb,
                                     void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(task_off && logits && logcounts && counts && profile_weight && lambda && losses &&
                 d8 && ws && host_task_off,
             "bpb_loss_fwd_bwd_pack: NULL argument");
  BP_REQUIRE(counts_partials == nullptr || (cb != nullptr && tiles == (L_in + 127) / 128),
             "bpb_loss_fwd_bwd_pack: counts_partials needs cb and tiles = ceil(L_in / 128)");
  BP_REQUIRE(B > 0 && L > 0 && T > 0 && H > 0, "bpb_loss_fwd_bwd_pack: bad sizes");
  BP_REQUIRE(Cp % 8 == 0 && Cp >= T + H && L == L_in - W + 1 && (planes == 1 || planes == 2),
             "bpb_loss_fwd_bwd_pack: bad geometry (Cp=%d must be a multiple of 8 >= T+H=%d)", Cp,
             T + H);
  BP_REQUIRE(2 * H + T + H <= 1024, "bpb_loss_fwd_bwd_pack: too many heads / tasks");
  for (int h = 0; h < H; ++h)
    BP_REQUIRE(host_task_off[h + 1] - host_task_off[h] >= 1 &&
                   host_task_off[h + 1] - host_task_off[h] <= kLossMaxTh,
               "bpb_loss_fwd_bwd_pack: head %d has %d tasks (1..%d supported; use bpb_loss_fwd_bwd + "
               "bpb_heads_pack_grad)", h, host_task_off[h + 1] - host_task_off[h], kLossMaxTh);
  LossPack pk{};
  pk.L_in = L_in;
  pk.W = W;
  pk.Cp = Cp;
  pk.planes = planes;
  // same buffer layout as bpb_heads_pack_grad / bpb_heads_d8_elems: each plane has a spare tail
  pk.plane_stride = static_cast<long long>(B) * (L_in + W - 1) * Cp + 8 * 512;
  pk.d8 = reinterpret_cast<__nv_bfloat16*>(d8);
  pk.dpb = dpb;
  pk.dcb = dcb;
  pk.ws = ws;
  pk.cpart = counts_partials;
  pk.cb = cb;
  pk.tiles = tiles;
  pk.logcounts_out = const_cast<float*>(logcounts);
  pk.dbg = debug_buffer();
  dim3 grid(B, H);
  static const int n_threads = [] {
    const char* e = getenv("BPB_LOSS_THREADS");
    const int v = e ? atoi(e) : 1024;
    return (v == 256 || v == 512 || v == 1024) ? v : 1024;
  }();
  // programmatic serialization: the successor (the heads' data gradient) may run its prologue
  // under this kernel's tail, and this kernel's launch latency hides under the heads' forward
  BP_CHECK_CUDA(launch_pdl(loss_pack_kernel, grid, dim3(n_threads), 0, stream, B, L, T, H, task_off,
                           logits, logcounts, counts, logcounts_true, profile_weight, lambda, losses,
                           dlogits, dlogcounts, pk));
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


extern "C" int bpb_adam_prep_step(long long n, float* param, const float* grad, float* m, float* v,
                                  float* state, float beta1, float beta2, float eps,
                                  float grad_scale, const bpb_prep_table* table, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(param && grad && m && v && state && table, "bpb_adam_prep_step: NULL argument");
  const bpb_prep_table& t = *table;
  BP_REQUIRE(t.F > 0 && t.F % 64 == 0 && t.Fw > 0 && t.Fw <= t.F && t.n_layers >= 0 &&
                 (t.planes == 1 || t.planes == 2) && t.in_planes >= 1 && t.in_planes <= 3 &&
                 t.Cp % 8 == 0 && t.Cp >= t.T + t.H,
             "bpb_adam_prep_step: bad table");
  BP_REQUIRE(t.w_in_op && t.w_hd_op && (t.n_layers == 0 || (t.w_fwd && t.w_bwd)) &&
                 (t.w_hf_op == nullptr || t.N_heads >= t.T + t.H),
             "bpb_adam_prep_step: missing operand buffers");
  BP_REQUIRE(n < (1ll << 31), "bpb_adam_prep_step: more than 2^31 parameters");
  if (n <= 0) return 0;
  BP_CHECK_CUDA(launch_pdl(adam_prep_kernel<false>, dim3(ew_blocks(n)), dim3(256), 0, stream, n, param,
                           static_cast<const float*>(grad), m, v, state, beta1, beta2, eps,
                           grad_scale, t, bpb_peer_args{}));
  return 0;
}

static int check_peer_args(const char* who, const bpb_peer_args* pa) {
  BP_REQUIRE(pa != nullptr && pa->world >= 1 && pa->world <= 8 && pa->rank >= 0 && pa->rank < pa->world,
             "%s: bad peer arguments (world 1..8)", who);
  for (int r = 0; r < pa->world; ++r)
    BP_REQUIRE(pa->grads[r] != nullptr && pa->flags[r] != nullptr, "%s: peer %d is not mapped", who, r);
  return 0;
}

extern "C" int bpb_adam_prep_step_peer(long long n, float* param, float* m, float* v, float* state,
                                       float beta1, float beta2, float eps,
                                       const bpb_prep_table* table, const bpb_peer_args* peers,
                                       void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(param && m && v && state && table, "bpb_adam_prep_step_peer: NULL argument");
  if (int rc = check_peer_args("bpb_adam_prep_step_peer", peers)) return rc;
  BP_REQUIRE(n < (1ll << 31), "bpb_adam_prep_step_peer: more than 2^31 parameters");
  if (n <= 0) return 0;
  // every block waits for the peers at its start: the whole grid must be resident
  int blocks = ew_blocks(n);
  adam_prep_kernel<true><<<blocks, 256, 0, stream>>>(n, param, nullptr, m, v, state, beta1, beta2, eps,
                                                     1.f / static_cast<float>(peers->world), *table,
                                                     *peers);
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int bpb_peer_wait(const bpb_peer_args* peers, void* stream_) {
  if (int rc = check_peer_args("bpb_peer_wait", peers)) return rc;
  peer_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream_)>>>(*peers);
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Memory a peer process can map (CUDA IPC): the gradient bucket and flag block of the exchange.
extern "C" int bpb_peer_alloc(long long bytes, void** ptr, unsigned char* handle64) {
  BP_REQUIRE(ptr && handle64 && bytes > 0, "bpb_peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  BP_CHECK_CUDA(cudaMalloc(ptr, static_cast<size_t>(bytes)));
  BP_CHECK_CUDA(cudaMemset(*ptr, 0, static_cast<size_t>(bytes)));
  cudaIpcMemHandle_t h;
  BP_CHECK_CUDA(cudaIpcGetMemHandle(&h, *ptr));
  memcpy(handle64, &h, 64);
  BP_CHECK_CUDA(cudaDeviceSynchronize());
  return 0;
}

extern "C" int bpb_peer_open(const unsigned char* handle64, void** ptr) {
  BP_REQUIRE(ptr && handle64, "bpb_peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  BP_CHECK_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" int bpb_peer_close(void* ptr) {
  if (ptr) BP_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}

extern "C" int bpb_peer_free(void* ptr) {
  if (ptr) BP_CHECK_CUDA(cudaFree(ptr));
  return 0;
}


extern "C" int bpb_onehot_encode(long long n, const unsigned char* seq, unsigned char* onehot,
                                 unsigned long long* n_other, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(seq && onehot, "bpb_onehot_encode: NULL argument");
  BP_REQUIRE((reinterpret_cast<uintptr_t>(onehot) & 3) == 0, "bpb_onehot_encode: onehot must be 4-byte aligned");
  if (n <= 0) return 0;
  if (n_other) BP_CHECK_CUDA(cudaMemsetAsync(n_other, 0, sizeof(unsigned long long), stream));
  onehot_encode_kernel<<<ew_blocks(n), 256, 0, stream>>>(n, seq, reinterpret_cast<uint32_t*>(onehot),
                                                         n_other);
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int bpb_logits_to_profile(int B, int L, int T, int H, const int* task_off,
                                     const float* logits, const float* logcounts, float* profile,
                                     void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(task_off && logits && logcounts && profile, "bpb_logits_to_profile: NULL argument");
  BP_REQUIRE(B > 0 && L > 0 && T > 0 && H > 0, "bpb_logits_to_profile: bad sizes");
  dim3 grid(B, H);
  logits_to_profile_kernel<<<grid, 256, 0, stream>>>(L, T, H, task_off, logits, logcounts, profile);
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

extern "C" int bpb_combine_bias(int B, int L, int T, int H, const int* task_off,
                                const float* bias_logits, const float* bias_lc, int n_pterms,
                                const int* p_act, int n_cterms, const int* c_act,
                                const float* coeffs, const int* use_bias,
                                const float* resid_logits, const float* resid_lc,
                                float* out_logits, float* out_lc, float* dresid_scale,
                                float* bias_logits_t, float* bias_lc_t, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(task_off && bias_logits && bias_lc && use_bias && resid_logits && resid_lc &&
                 out_logits && out_lc, "bpb_combine_bias: NULL argument");
  BP_REQUIRE(n_pterms >= 0 && n_cterms >= 0 && (n_pterms + n_cterms == 0 || coeffs) &&
                 (n_pterms == 0 || p_act) && (n_cterms == 0 || c_act),
             "bpb_combine_bias: transformation terms without their tables");
  const long long total = static_cast<long long>(B) * L * T + static_cast<long long>(B) * H;
  if (total <= 0) return 0;
  combine_bias_kernel<<<ew_blocks(total), 256, 0, stream>>>(
      B, L, T, H, task_off, bias_logits, bias_lc, n_pterms, p_act, n_cterms, c_act, coeffs,
      use_bias, resid_logits, resid_lc, out_logits, out_lc, dresid_scale, bias_logits_t, bias_lc_t);
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int bpb_combined_shap_seed(int R, int n, int L, int T, int H, const int* task_off,
                                      const float* bias_logits_x, const float* bias_logits_r,
                                      const float* bias_lc_x, const float* bias_lc_r,
                                      const float* resid_lc_x, const float* resid_lc_r,
                                      int n_pterms, const int* p_act, int n_cterms,
                                      const int* c_act, const float* coeffs, const int* use_bias,
                                      const float* dlogits_c, const float* dlc_c,
                                      float* dlogits_resid, float* dlc_resid, float* dlogits_bias,
                                      float* dlc_bias, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(task_off && bias_logits_x && bias_logits_r && bias_lc_x && bias_lc_r && dlogits_c &&
                 dlc_c && n > 0 && R % n == 0,
             "bpb_combined_shap_seed: bad argument");
  BP_REQUIRE((resid_lc_x == nullptr) == (resid_lc_r == nullptr) && (resid_lc_x == nullptr || use_bias),
             "bpb_combined_shap_seed: residual log-counts need both halves and use_bias");
  BP_REQUIRE((n_pterms + n_cterms == 0 || coeffs) && (n_pterms == 0 || p_act) &&
                 (n_cterms == 0 || c_act),
             "bpb_combined_shap_seed: transformation terms without their tables");
  const long long total = static_cast<long long>(R) * L * T + static_cast<long long>(R) * H;
  if (total <= 0) return 0;
  combined_shap_seed_kernel<<<ew_blocks(total), 256, 0, stream>>>(
      R, n, L, T, H, task_off, bias_logits_x, bias_logits_r, bias_lc_x, bias_lc_r, resid_lc_x,
      resid_lc_r, n_pterms, p_act, n_cterms, c_act, coeffs, use_bias, dlogits_c, dlc_c,
      dlogits_resid, dlc_resid, dlogits_bias, dlc_bias);
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
  sx = block_reduce(sx, SumF(), scratch, 0.f);
  sr = block_reduce(sr, SumF(), scratch, 0.f);
  mxx = block_reduce(mxx, MaxF(), scratch, -INFINITY);
  mxr = block_reduce(mxr, MaxF(), scratch, -INFINITY);
  const float J = block_reduce(static_cast<float>(cnt), SumF(), scratch, 0.f);
  const float mean_x = sx / J, mean_r = sr / J;
  float ex = 0.f, er = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    if (sel[i % T]) {
      ex += __expf(lx[i] - mxx);
      er += __expf(lr[i] - mxr);
    }
  ex = block_reduce(ex, SumF(), scratch, 0.f);
  er = block_reduce(er, SumF(), scratch, 0.f);
  float sm = 0.f, ax = 0.f, ar = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    if (sel[i % T]) {
      const float px = __expf(lx[i] - mxx) / ex, pr = __expf(lr[i] - mxr) / er;
      const float tx = lx[i] - mean_x, tr = lr[i] - mean_r;
      sm += (fabsf(tx - tr) < 1e-7f) ? 0.f : 0.5f * (px + pr);
      ax += px * tx;
      ar += pr * tr;
    }
  sm = block_reduce(sm, SumF(), scratch, 0.f);
  ax = block_reduce(ax, SumF(), scratch, 0.f);
  ar = block_reduce(ar, SumF(), scratch, 0.f);
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
