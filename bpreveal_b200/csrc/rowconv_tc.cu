// K1 forward (one-hot input convolution, models.py:85-90) and the data gradient of the profile /
// counts heads (backward of models.py:39-49) as position-major tcgen05 GEMMs on the engine of
// conv_core.cuh.
//
// Both are "window" convolutions over a narrow channels-last tensor: the one-hot sequence padded
// to 8 channels (x8), or the loss gradient d8 = [dlogits | dlogcounts / L | 0] padded to Cp
// channels.  Row t of the implicit im2col matrix is the contiguous run of W*C elements starting at
// element t*C, so a TMA tensor map whose row stride (C*2 bytes) is smaller than its row extent
// delivers im2col tiles straight into swizzled shared memory: no gather kernel, no K1 multiplies
// spent on anything but the 8 x W window, K padded only to the next 16.
#include "conv_core.cuh"

namespace bp {

// STRIP staging reads whole strips of 127 + 2*nk16 positions, also for the last tile of the last
// sequence: window tensors (x8, d8 planes) carry this many spare elements at their end.
constexpr long long kStripPadElems = 8 * 512;

__global__ void __launch_bounds__(256)
onehot_expand_kernel(long long n_pos, const uint32_t* __restrict__ x, uint4* __restrict__ x8) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n_pos) return;
  const uint32_t v = x[i];  // 4 one-hot bytes, little endian: A C G T
  uint4 o;
  // byte (0|1) -> bf16 (0.0|1.0 = 0x3F80), two per word
  o.x = ((v & 0x1u) * 0x3F80u) | (((v >> 8) & 0x1u) * 0x3F800000u);
  o.y = (((v >> 16) & 0x1u) * 0x3F80u) | (((v >> 24) & 0x1u) * 0x3F800000u);
  o.z = 0u;
  o.w = 0u;
  x8[i] = o;
}

__device__ __forceinline__ void store_split(__nv_bfloat16* dst, long long plane_stride, int planes,
                                            long long idx, float v) {
  // successive bf16 remainders: 1 plane = bf16, 2 planes ~ 16 mantissa bits, 3 planes = fp32
  float r = v;
  for (int pl = 0; pl < planes; ++pl) {
    const __nv_bfloat16 h = __float2bfloat16_rn(r);
    dst[pl * plane_stride + idx] = h;
    r -= __bfloat162float(h);
  }
}

// w [W][4][Fw] fp32 -> op [planes][F][KC*64] bf16, k = j*8 + c
__global__ void __launch_bounds__(256)
prep_input_conv_kernel(int W, int Fw, int F, int K, int planes, const float* __restrict__ w,
                       __nv_bfloat16* __restrict__ op) {
  const long long n_el = static_cast<long long>(F) * K;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n_el) return;
  const int n = static_cast<int>(i / K), k = static_cast<int>(i % K);
  const int j = k >> 3, c = k & 7;
  const float v = (j < W && c < 4 && n < Fw) ? w[(j * 4 + c) * Fw + n] : 0.f;
  store_split(op, n_el, planes, i, v);
}

// d8 [planes][B][L_in + W - 1][Cp]: row r <-> logits position r - (W - 1)
__global__ void __launch_bounds__(256)
heads_pack_grad_kernel(int B, int L_in, int L_out, int W, int T, int H, int Cp, int planes,
                       const float* __restrict__ dlogits, const float* __restrict__ dlc,
                       __nv_bfloat16* __restrict__ d8) {
  const int Lp = L_in + W - 1;
  const long long n_el = static_cast<long long>(B) * Lp * Cp;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n_el) return;
  const long long plane_stride = n_el + kStripPadElems;  // see bpb_heads_d8_elems
  const int ch = static_cast<int>(i % Cp);
  const int r = static_cast<int>((i / Cp) % Lp);
  const int b = static_cast<int>(i / (static_cast<long long>(Cp) * Lp));
  const int t = r - (W - 1);
  float v = 0.f;
  if (ch < T) {
    if (t >= 0 && t < L_out) v = dlogits[(static_cast<long long>(b) * L_out + t) * T + ch];
  } else if (ch < T + H) {
    v = dlc[b * H + (ch - T)] / static_cast<float>(L_in);
  }
  store_split(d8, plane_stride, planes, i, v);
}

// Bias gradients of the heads in fp32 (they are sums of terms that cancel almost exactly -- the
// multinomial gradient of a head sums to zero -- so they are not taken from the bf16 d8 tensor):
//   dpb[k] = sum_{b,t} dlogits[b,t,k]    dcb[h] = sum_b dlogcounts[b,h]
// One block per output, fixed summation order (deterministic).
__global__ void __launch_bounds__(1024)
heads_bias_grad_kernel(long long n_rows, int T, int B, int H, const float* __restrict__ dlogits,
                       const float* __restrict__ dlc, float* __restrict__ dpb,
                       float* __restrict__ dcb) {
  __shared__ float red[1024];
  const int o = blockIdx.x;
  float s = 0.f;
  if (o < T) {
    if ((T == 1 || T == 2 || T == 4) && (reinterpret_cast<uintptr_t>(dlogits) & 15) == 0) {
      // coalesced 16-byte reads of the whole [rows][T] array, many in flight: element e belongs
      // to channel e % T, and T divides 4
      const long long n = n_rows * T, n4 = n >> 2;
      const float4* p4 = reinterpret_cast<const float4*>(dlogits);
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 8
      for (long long e = threadIdx.x; e < n4; e += blockDim.x) {
        const float4 v = __ldg(p4 + e);
        s0 += v.x; s1 += v.y; s2 += v.z; s3 += v.w;
      }
      if (T == 1) s = (s0 + s1) + (s2 + s3);
      else if (T == 2) s = (o == 0) ? s0 + s2 : s1 + s3;
      else s = (o == 0) ? s0 : (o == 1) ? s1 : (o == 2) ? s2 : s3;
      if (threadIdx.x == 0)
        for (long long e = n4 << 2; e < n; ++e)
          if (static_cast<int>(e % T) == o) s += dlogits[e];
    } else {
      for (long long r = threadIdx.x; r < n_rows; r += blockDim.x) s += dlogits[r * T + o];
    }
  } else {
    for (int b = threadIdx.x; b < B; b += blockDim.x) s += dlc[b * H + (o - T)];
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 512; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (o < T) dpb[o] = red[0];
    else dcb[o - T] = red[0];
  }
}

// pw [W][Fw][T], cw [Fw][H] fp32 -> op [planes][F][KC*64] bf16, k = j'*Cp + ch with j' = W-1-j;
// the counts columns ride on tap j' = 0 only.
__global__ void __launch_bounds__(256)
prep_heads_dgrad_kernel(int W, int Fw, int F, int T, int H, int Cp, int K, int planes,
                        const float* __restrict__ pw, const float* __restrict__ cw,
                        __nv_bfloat16* __restrict__ op) {
  const long long n_el = static_cast<long long>(F) * K;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n_el) return;
  const int c = static_cast<int>(i / K), k = static_cast<int>(i % K);
  const int jp = k / Cp, ch = k % Cp;
  float v = 0.f;
  if (c < Fw && jp < W) {
    if (ch < T) v = pw[(static_cast<long long>(W - 1 - jp) * Fw + c) * T + ch];
    else if (ch < T + H && jp == 0) v = cw[c * H + (ch - T)];
  }
  store_split(op, n_el, planes, i, v);
}

static int window_kc(int W, int C) { return (W * C + 63) / 64; }
static int window_last_nk(int W, int C) {
  const int rem = W * C - (window_kc(W, C) - 1) * 64;
  return (rem + 15) / 16;
}

}  // namespace bp

extern "C" int bpb_onehot_expand(int B, int L, const uint8_t* x, bpb_bf16* x8, void* stream_) {
  using namespace bp;
  BP_REQUIRE(x && x8 && B > 0 && L > 0, "bpb_onehot_expand: bad arguments");
  const long long n = static_cast<long long>(B) * L;
  onehot_expand_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0,
                         static_cast<cudaStream_t>(stream_)>>>(
      n, reinterpret_cast<const uint32_t*>(x), reinterpret_cast<uint4*>(x8));
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" long long bpb_x8_elems(int B, int L) {
  return static_cast<long long>(B) * L * 8 + bp::kStripPadElems;
}

extern "C" long long bpb_input_conv_wop_elems(int W, int F, int planes) {
  return static_cast<long long>(planes) * F * bp::window_kc(W, 8) * 64;
}

extern "C" int bpb_prep_input_conv_weights(int W, int Fw, int F, const float* w, bpb_bf16* w_op,
                                           int planes, void* stream_) {
  using namespace bp;
  BP_REQUIRE(w && w_op && W > 0 && Fw > 0 && F >= Fw && F % 64 == 0 && planes >= 1 && planes <= 3,
             "bpb_prep_input_conv_weights: bad arguments");
  const int K = window_kc(W, 8) * 64;
  const long long n = static_cast<long long>(F) * K;
  prep_input_conv_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0,
                           static_cast<cudaStream_t>(stream_)>>>(
      W, Fw, F, K, planes, w, reinterpret_cast<__nv_bfloat16*>(w_op));
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int bpb_input_conv_fwd(const bpb_input_conv_args* a, void* stream_) {
  using namespace bp;
  BP_REQUIRE(a != nullptr, "bpb_input_conv_fwd: null args");
  BP_REQUIRE(a->x8 && a->w_op && a->out_hi, "bpb_input_conv_fwd: x8 / w_op / out_hi must not be NULL");
  BP_REQUIRE(a->F > 0 && a->F % 64 == 0 && a->Fw <= a->F, "bpb_input_conv_fwd: bad F=%d Fw=%d", a->F,
             a->Fw);
  BP_REQUIRE(a->L_out == a->L_in - a->W + 1 && a->L_out > 0, "bpb_input_conv_fwd: bad lengths");
  BP_REQUIRE(a->w_planes >= 1 && a->w_planes <= 3, "bpb_input_conv_fwd: w_planes must be 1..3");
  const int planes = a->out_lo ? 2 : 1;
  int epi = EPI_FWD;
  if (a->mult_hi) {
    BP_REQUIRE(a->zx_in && a->zx_div > 0 && !a->z_out && !a->mask_out,
               "bpb_input_conv_fwd: mult needs zx_in and excludes z_out / mask_out");
    BP_REQUIRE(planes == 1 || a->mult_lo, "bpb_input_conv_fwd: precise path needs mult_lo");
    epi = EPI_FWD_R;
  } else if (a->z_out) {
    epi = EPI_FWD_Z;
  }
  ConvDesc d{};
  d.who = "bpb_input_conv_fwd";
  d.B = a->B;
  d.L_out = a->L_out;
  d.F = a->F;
  d.epi = epi;
  d.planes = planes;
  d.a_ptr[0] = a->x8;
  d.n_aplanes = 1;
  d.a_inner = static_cast<unsigned long long>(a->W) * 8;
  d.a_rows = a->L_out;
  d.a_row_stride_bytes = 16;
  d.a_batch_stride_bytes = static_cast<unsigned long long>(a->L_in) * 16;
  d.n_taps = 1;
  d.KC = window_kc(a->W, 8);
  d.last_nk = window_last_nk(a->W, 8);
  d.tap_off[0] = 0;
  d.w_ptr = a->w_op;
  d.n_wplanes = a->w_planes;
  d.npass = a->w_planes;  // the one-hot operand is exact in bf16: x*w_hi + x*w_mid (+ x*w_lo)
  d.pa_bits = 0;
  d.pw_bits = 0b100100;
  d.bias = a->bias;
  d.bias_n = a->Fw;
  d.out[0] = a->out_hi;
  d.out[1] = a->out_lo;
  d.out2[0] = a->mult_hi;
  d.out2[1] = a->mult_lo;
  d.mask_out = a->mask_out;
  d.zx_in = a->zx_in;
  d.zx_div = a->zx_div;
  d.z_out = a->z_out;
  d.allow_halo = false;
  d.allow_strip = true;  // x8 must be padded by bpb_strip_pad_elems() elements at its end
  return conv_run<false>(d, static_cast<cudaStream_t>(stream_));
}

static long long d8_plane_elems(int B, int L_in, int W, int Cp) {
  return static_cast<long long>(B) * (L_in + W - 1) * Cp + bp::kStripPadElems;
}

extern "C" long long bpb_heads_d8_elems(int B, int L_in, int W, int Cp, int planes) {
  return planes * d8_plane_elems(B, L_in, W, Cp);
}

extern "C" long long bpb_heads_dgrad_wop_elems(int W, int F, int Cp, int planes) {
  return static_cast<long long>(planes) * F * bp::window_kc(W, Cp) * 64;
}

extern "C" int bpb_heads_pack_grad(int B, int L_in, int L_out, int W, int T, int H, int Cp,
                                   int planes, const float* dlogits, const float* dlogcounts,
                                   bpb_bf16* d8, float* dpb, float* dcb, void* stream_) {
  using namespace bp;
  BP_REQUIRE(dlogits && dlogcounts && d8, "bpb_heads_pack_grad: NULL argument");
  BP_REQUIRE(Cp % 8 == 0 && Cp >= T + H && L_out == L_in - W + 1 && (planes == 1 || planes == 2),
             "bpb_heads_pack_grad: bad geometry (Cp=%d must be a multiple of 8 >= T+H=%d)", Cp, T + H);
  const long long n = static_cast<long long>(B) * (L_in + W - 1) * Cp;
  heads_pack_grad_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0,
                           static_cast<cudaStream_t>(stream_)>>>(
      B, L_in, L_out, W, T, H, Cp, planes, dlogits, dlogcounts,
      reinterpret_cast<__nv_bfloat16*>(d8));
  BP_CHECK_CUDA(cudaGetLastError());
  if (dpb != nullptr) {
    BP_REQUIRE(dcb != nullptr, "bpb_heads_pack_grad: dpb and dcb go together");
    heads_bias_grad_kernel<<<T + H, 1024, 0, static_cast<cudaStream_t>(stream_)>>>(
        static_cast<long long>(B) * L_out, T, B, H, dlogits, dlogcounts, dpb, dcb);
    BP_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

extern "C" int bpb_prep_heads_dgrad_weights(int W, int Fw, int F, int T, int H, int Cp,
                                            const float* pw, const float* cw, bpb_bf16* w_op,
                                            int planes, void* stream_) {
  using namespace bp;
  BP_REQUIRE(pw && cw && w_op && Cp % 8 == 0 && Cp >= T + H && F % 64 == 0 && Fw <= F &&
                 (planes == 1 || planes == 2),
             "bpb_prep_heads_dgrad_weights: bad arguments");
  const int K = window_kc(W, Cp) * 64;
  const long long n = static_cast<long long>(F) * K;
  prep_heads_dgrad_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0,
                            static_cast<cudaStream_t>(stream_)>>>(
      W, Fw, F, T, H, Cp, K, planes, pw, cw, reinterpret_cast<__nv_bfloat16*>(w_op));
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int bpb_heads_dgrad_tc(const bpb_heads_dgrad_args* a, void* stream_) {
  using namespace bp;
  BP_REQUIRE(a != nullptr, "bpb_heads_dgrad_tc: null args");
  BP_REQUIRE(a->d8_hi && a->w_op && (a->g_hi || a->gz_hi), "bpb_heads_dgrad_tc: NULL argument");
  BP_REQUIRE(a->F > 0 && a->F % 64 == 0 && a->Cp % 8 == 0 && a->L_out == a->L_in - a->W + 1,
             "bpb_heads_dgrad_tc: bad geometry");
  BP_REQUIRE(!(a->mask_in && a->mult_hi), "bpb_heads_dgrad_tc: give mask_in or mult, not both");
  const int planes = a->d8_lo ? 2 : 1;
  if (planes == 2)
    BP_REQUIRE((!a->g_hi || a->g_lo) && (!a->gz_hi || a->gz_lo) && (!a->mult_hi || a->mult_lo),
               "bpb_heads_dgrad_tc: precise path needs the lo planes");
  const int Lp = a->L_in + a->W - 1;
  ConvDesc d{};
  d.who = "bpb_heads_dgrad_tc";
  d.B = a->B;
  d.L_out = a->L_in;  // rows of the result = activations of the last tower layer
  d.F = a->F;
  d.epi = a->mult_hi ? EPI_BWD_MULT : EPI_BWD_MASK;
  d.planes = planes;
  d.a_ptr[0] = a->d8_hi;
  d.a_ptr[1] = a->d8_lo;
  d.n_aplanes = planes;
  d.a_inner = static_cast<unsigned long long>(a->W) * a->Cp;
  d.a_rows = a->L_in;
  d.a_row_stride_bytes = static_cast<unsigned long long>(a->Cp) * 2;
  d.a_batch_stride_bytes = static_cast<unsigned long long>(Lp) * a->Cp * 2;
  d.n_taps = 1;
  d.KC = window_kc(a->W, a->Cp);
  d.last_nk = window_last_nk(a->W, a->Cp);
  d.tap_off[0] = 0;
  d.w_ptr = a->w_op;
  d.n_wplanes = planes;
  d.npass = planes == 2 ? 3 : 1;
  d.pa_bits = 0b010;
  d.pw_bits = 0b010000;
  d.out[0] = a->g_hi;
  d.out[1] = a->g_lo;
  d.out2[0] = a->gz_hi;
  d.out2[1] = a->gz_lo;
  d.mask_in = a->mask_in;
  d.mult[0] = a->mult_hi;
  d.mult[1] = a->mult_lo;
  d.allow_halo = false;
  d.allow_strip = true;  // taken when Cp == 8 (16 bytes per position)
  return conv_run<false>(d, static_cast<cudaStream_t>(stream_));
}
