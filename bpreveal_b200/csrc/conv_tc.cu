// K2 / K6 / K8: width-3 dilated convolution of the BPNet tower as a tcgen05 implicit GEMM.
//
// Replaces, for models.py:92-104 of the reference, the chain SpaceToBatchND -> Conv2D ->
// BatchToSpaceND -> BiasAdd -> Relu -> StridedSlice (Cropping1D) -> AddV2 that TensorFlow runs per
// dilated layer, its data gradient, and the DeepSHAP-rescale variant of that
// gradient (shap.py:612-639).
//
// GEMM view: M = 128 consecutive positions of one sequence, N = NT output channels, K = 3 taps x
// F input channels.  The A operand of tap j is the activation tile at row offset tap_off[j]: one
// TMA box [128 rows x 64 ch] (128-byte swizzle, K-major) per (tap, 64-channel chunk, plane); rows
// outside the sequence are zero-filled by TMA, which gives the 'valid' forward conv and the 'full'
// backward conv from the same code.  Accumulators live in TMEM (4 or 2 stages), one elected thread
// issues tcgen05.mma, four epilogue warps read TMEM, fuse bias / ReLU / crop+residual (forward) or
// residual-gradient + ReLU-mask / rescale-multiplier (backward) and TMA-store bf16 planes.
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-5
// epilogue (TMEM quadrant = warp % 4).
#include "common.cuh"
#include "ptx.cuh"

namespace bp {

constexpr int kTileM = 128;
constexpr int kATileBytes = kTileM * 128;  // [128 x 64] bf16
constexpr int kMaxStages = 12;

struct ConvKParams {
  int B, L_out, tiles_per_seq, n_ntiles, NC, n_planes, n_passes, F;
  int tap_off[3];
  int mode, stages;
  int L_res, resid_off;
  int zx_div;
  int has_out, has_out2;
  const float* bias;
  const __nv_bfloat16 *resid_hi, *resid_lo;
  const __nv_bfloat16 *mult_hi, *mult_lo;
  const unsigned long long* mask_in;
  unsigned long long* mask_out;
  const float* zx_in;
  float* z_out;
};

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]);
  u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
  return u;
}
// lo plane = bf16(value - float(bf16(value)))
__device__ __forceinline__ void split8(const float* f, uint4& hi, uint4& lo) {
  hi = pack8(f);
  float h[8], r[8];
  unpack8(hi, h);
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = f[i] - h[i];
  lo = pack8(r);
}

template <int NT, bool RESIDENT>
__global__ void __launch_bounds__(192, 1)
conv3_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO0,
                const __grid_constant__ CUtensorMap tmO1, const __grid_constant__ CUtensorMap tmP0,
                const __grid_constant__ CUtensorMap tmP1, const ConvKParams p) {
  constexpr int ACC = (NT <= 128) ? 4 : 2;
  constexpr uint32_t TMEM_COLS = ACC * NT;
  constexpr int B_TILE_BYTES = NT * 128;
  constexpr int STAGE_BYTES = kATileBytes + (RESIDENT ? 0 : B_TILE_BYTES);
  constexpr uint32_t IDESC = umma_idesc_bf16(128, NT, 0, 0);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));

  const int w_tiles = RESIDENT ? p.n_planes * 3 * p.NC : 0;
  uint8_t* w_smem = smem;
  uint8_t* ring = w_smem + static_cast<size_t>(w_tiles) * B_TILE_BYTES;
  uint8_t* stg = ring + static_cast<size_t>(p.stages) * STAGE_BYTES;
  const int n_stg = (p.has_out + p.has_out2) * p.n_planes;
  float* s_bias = reinterpret_cast<float*>(stg + static_cast<size_t>(n_stg) * kATileBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + p.F);
  uint64_t* full = bars;
  uint64_t* empty = bars + kMaxStages;
  uint64_t* acc_full = bars + 2 * kMaxStages;
  uint64_t* acc_empty = acc_full + 4;
  uint64_t* w_full = acc_empty + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.B * p.tiles_per_seq * p.n_ntiles;
  const int nsteps = 3 * p.NC * p.n_passes;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmW);
    if (p.n_planes == 2) tma_prefetch_desc(&tmA1);
    if (p.has_out) tma_prefetch_desc(&tmO0);
    if (p.has_out2) tma_prefetch_desc(&tmP0);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < ACC; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  for (int i = threadIdx.x; i < p.F; i += blockDim.x) s_bias[i] = p.bias ? p.bias[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    if (elect_one()) {
      if (RESIDENT) {
        mbar_expect_tx(w_full, static_cast<uint32_t>(w_tiles) * B_TILE_BYTES);
        for (int wp = 0; wp < p.n_planes; ++wp)
          for (int j = 0; j < 3; ++j)
            for (int c = 0; c < p.NC; ++c)
              tma_load_2d(&tmW, w_smem + static_cast<size_t>((wp * 3 + j) * p.NC + c) * B_TILE_BYTES,
                          w_full, c * 64, (wp * 3 + j) * p.F);
      }
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_ntiles;
        const int mt = (tile / p.n_ntiles) % p.tiles_per_seq;
        const int b = tile / (p.n_ntiles * p.tiles_per_seq);
        const int t0 = mt * kTileM;
        for (int step = 0; step < nsteps; ++step) {
          const int pass = step % p.n_passes;
          const int c = (step / p.n_passes) % p.NC;
          const int j = step / (p.n_passes * p.NC);
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], STAGE_BYTES);
          uint8_t* dst = ring + static_cast<size_t>(stage) * STAGE_BYTES;
          tma_load_3d(pass == 1 ? &tmA1 : &tmA0, dst, &full[stage], c * 64, t0 + p.tap_off[j], b);
          if (!RESIDENT) {
            const int wp = (pass == 2) ? 1 : 0;
            tma_load_2d(&tmW, dst + kATileBytes, &full[stage], c * 64,
                        (wp * 3 + j) * p.F + nt * NT);
          }
          if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer
    if (elect_one()) {
      if (RESIDENT) mbar_wait(w_full, 0);
      uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&acc_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * NT;
        for (int step = 0; step < nsteps; ++step) {
          const int pass = step % p.n_passes;
          const int c = (step / p.n_passes) % p.NC;
          const int j = step / (p.n_passes * p.NC);
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(ring + static_cast<size_t>(stage) * STAGE_BYTES);
          uint32_t b_addr;
          if (RESIDENT) {
            const int wp = (pass == 2) ? 1 : 0;
            b_addr = smem_u32(w_smem + static_cast<size_t>((wp * 3 + j) * p.NC + c) * B_TILE_BYTES);
          } else {
            b_addr = a_addr + kATileBytes;
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16(d_tmem, umma_smem_desc(a_addr + k * 32, 0, 1024),
                      umma_smem_desc(b_addr + k * 32, 0, 1024), IDESC, (step | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[as]);
        if (++as == ACC) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------- epilogue warps
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const int chunks_per_row = p.F >> 6;
    uint32_t as = 0, aphase = 0;
    uint8_t* stg_out = stg;
    uint8_t* stg_out2 = stg + static_cast<size_t>(p.has_out * p.n_planes) * kATileBytes;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_ntiles;
      const int mt = (tile / p.n_ntiles) % p.tiles_per_seq;
      const int b = tile / (p.n_ntiles * p.tiles_per_seq);
      const int t0 = mt * kTileM;
      const int t = t0 + row;
      const bool valid = t < p.L_out;
      const int tr = t + p.resid_off;
      const bool rvalid = valid && p.resid_hi != nullptr && tr >= 0 && tr < p.L_res;
      const size_t orow = static_cast<size_t>(b) * p.L_out + t;

      for (int ch = 0; ch < NT / 64; ++ch) {
        const int n0 = nt * NT + ch * 64;
        // Issue the row-wise global reads first so that they overlap the accumulator wait.
        uint4 rr[8], rl[8];
        if (rvalid) {
          const uint4* rp = reinterpret_cast<const uint4*>(
              p.resid_hi + (static_cast<size_t>(b) * p.L_res + tr) * p.F + n0);
#pragma unroll
          for (int g = 0; g < 8; ++g) rr[g] = __ldg(rp + g);
          if (p.resid_lo) {
            const uint4* rq = reinterpret_cast<const uint4*>(
                p.resid_lo + (static_cast<size_t>(b) * p.L_res + tr) * p.F + n0);
#pragma unroll
            for (int g = 0; g < 8; ++g) rl[g] = __ldg(rq + g);
          }
        }
        unsigned long long mbits = ~0ull;
        if (p.mode == 1 && p.mask_in != nullptr && valid)
          mbits = __ldg(p.mask_in + orow * chunks_per_row + (n0 >> 6));

        if (ch == 0) {
          mbar_wait(&acc_full[as], aphase);
          tc_fence_after();
        }
        uint32_t acc[64];
        const uint32_t taddr = tmem_base + as * NT + ch * 64 + (static_cast<uint32_t>(q * 32) << 16);
        tmem_ld32(taddr, acc);
        tmem_ld32(taddr + 32, acc + 32);
        tmem_ld_wait();
        if (ch == NT / 64 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[as]);
        }

        // staging buffers must be free (previous TMA stores have read them)
        if (et == 0) tma_store_wait_read<0>();
        named_bar_sync(1, 128);

        unsigned long long obits = 0ull;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float v[8], r[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(acc[g * 8 + i]);
          if (rvalid) {
            unpack8(rr[g], r);
            if (p.resid_lo) {
              float r2[8];
              unpack8(rl[g], r2);
#pragma unroll
              for (int i = 0; i < 8; ++i) r[i] += r2[i];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = 0.f;
          }
          float o[8], o2[8];
          if (p.mode == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += s_bias[n0 + g * 8 + i];
            if (p.zx_in != nullptr) {
              // DeepSHAP rescale multiplier of this reference against the query's pre-activation
              // (shap.py:623-638); only the query-half gradient is used downstream (shap.py:374).
              float zx[8];
              if (valid) {
                const float4* zp = reinterpret_cast<const float4*>(
                    p.zx_in + (static_cast<size_t>(b / p.zx_div) * p.L_out + t) * p.F + n0 + g * 8);
                const float4 z0 = __ldg(zp), z1 = __ldg(zp + 1);
                zx[0] = z0.x; zx[1] = z0.y; zx[2] = z0.z; zx[3] = z0.w;
                zx[4] = z1.x; zx[5] = z1.y; zx[6] = z1.z; zx[7] = z1.w;
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) zx[i] = 0.f;
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float din = zx[i] - v[i];
                const float dout = fmaxf(zx[i], 0.f) - fmaxf(v[i], 0.f);
                o2[i] = (fabsf(din) < 1e-6f) ? (zx[i] > 0.f ? 1.f : 0.f) : dout / din;
              }
            }
            if (p.z_out != nullptr && valid) {
              float4* zo = reinterpret_cast<float4*>(p.z_out + orow * p.F + n0 + g * 8);
              zo[0] = make_float4(v[0], v[1], v[2], v[3]);
              zo[1] = make_float4(v[4], v[5], v[6], v[7]);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (v[i] > 0.f) obits |= 1ull << (g * 8 + i);
              o[i] = fmaxf(v[i], 0.f) + r[i];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = v[i] + r[i];
            if (p.mult_hi != nullptr) {
              float m[8];
              if (valid) {
                const uint4 mu = __ldg(reinterpret_cast<const uint4*>(p.mult_hi + orow * p.F + n0) + g);
                unpack8(mu, m);
                if (p.mult_lo) {
                  float m2[8];
                  const uint4 ml = __ldg(reinterpret_cast<const uint4*>(p.mult_lo + orow * p.F + n0) + g);
                  unpack8(ml, m2);
#pragma unroll
                  for (int i = 0; i < 8; ++i) m[i] += m2[i];
                }
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) m[i] = 0.f;
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) o2[i] = o[i] * m[i];
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) o2[i] = ((mbits >> (g * 8 + i)) & 1ull) ? o[i] : 0.f;
            }
          }
          const uint32_t off = sw128(row, g);
          if (p.has_out) {
            if (p.n_planes == 2) {
              uint4 h, l;
              split8(o, h, l);
              *reinterpret_cast<uint4*>(stg_out + off) = h;
              *reinterpret_cast<uint4*>(stg_out + kATileBytes + off) = l;
            } else {
              *reinterpret_cast<uint4*>(stg_out + off) = pack8(o);
            }
          }
          if (p.has_out2) {
            if (p.n_planes == 2) {
              uint4 h, l;
              split8(o2, h, l);
              *reinterpret_cast<uint4*>(stg_out2 + off) = h;
              *reinterpret_cast<uint4*>(stg_out2 + kATileBytes + off) = l;
            } else {
              *reinterpret_cast<uint4*>(stg_out2 + off) = pack8(o2);
            }
          }
        }
        if (p.mode == 0 && p.mask_out != nullptr && valid)
          p.mask_out[orow * chunks_per_row + (n0 >> 6)] = obits;

        fence_proxy_async();
        named_bar_sync(1, 128);
        if (et == 0) {
          if (p.has_out) {
            tma_store_3d(&tmO0, stg_out, n0, t0, b);
            if (p.n_planes == 2) tma_store_3d(&tmO1, stg_out + kATileBytes, n0, t0, b);
          }
          if (p.has_out2) {
            tma_store_3d(&tmP0, stg_out2, n0, t0, b);
            if (p.n_planes == 2) tma_store_3d(&tmP1, stg_out2 + kATileBytes, n0, t0, b);
          }
          tma_store_commit();
        }
      }
      if (++as == ACC) { as = 0; aphase ^= 1; }
    }
    if (et == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

template <int NT, bool RESIDENT>
static int launch_conv3(const CUtensorMap* maps, const ConvKParams& kp, size_t smem_bytes, int grid,
                        cudaStream_t stream) {
  auto kern = conv3_tc_kernel<NT, RESIDENT>;
  static bool configured = false;
  if (!configured) {
    BP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       max_smem_optin()));
    configured = true;
  }
  kern<<<grid, 192, smem_bytes, stream>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5],
                                          maps[6], kp);
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace bp

extern "C" int bpb_conv3_tc(const bpb_conv3_args* a, void* stream_) {
  using namespace bp;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(a != nullptr, "bpb_conv3_tc: null args");
  BP_REQUIRE(a->F > 0 && a->F % 64 == 0, "bpb_conv3_tc: F=%d must be a positive multiple of 64",
             a->F);
  BP_REQUIRE(a->B > 0 && a->L_in > 0 && a->L_out > 0, "bpb_conv3_tc: bad geometry B=%d L_in=%d L_out=%d",
             a->B, a->L_in, a->L_out);
  BP_REQUIRE(a->in_hi && a->w_hi, "bpb_conv3_tc: in_hi / w_hi must not be NULL");
  BP_REQUIRE(a->mode == 0 || a->mode == 1, "bpb_conv3_tc: mode must be 0 or 1");
  const int n_planes = a->in_lo ? 2 : 1;
  if (n_planes == 2)
    BP_REQUIRE(a->w_lo != nullptr, "bpb_conv3_tc: precise path needs w_lo");
  const int has_out = a->out_hi ? 1 : 0, has_out2 = a->out2_hi ? 1 : 0;
  BP_REQUIRE(has_out || has_out2 || a->z_out || a->mask_out, "bpb_conv3_tc: no output requested");
  if (n_planes == 2) {
    BP_REQUIRE(!has_out || a->out_lo, "bpb_conv3_tc: precise path needs out_lo");
    BP_REQUIRE(!has_out2 || a->out2_lo, "bpb_conv3_tc: precise path needs out2_lo");
  }
  if (a->mode == 0 && has_out2)
    BP_REQUIRE(a->zx_in != nullptr && a->zx_div > 0, "bpb_conv3_tc: forward out2 needs zx_in");

  const int F = a->F;
  const int NT = (F % 256 == 0) ? 256 : (F % 128 == 0 ? 128 : 64);
  ConvKParams kp{};
  kp.B = a->B;
  kp.L_out = a->L_out;
  kp.tiles_per_seq = (a->L_out + kTileM - 1) / kTileM;
  kp.n_ntiles = F / NT;
  kp.NC = F / 64;
  kp.n_planes = n_planes;
  kp.n_passes = n_planes == 2 ? 3 : 1;
  kp.F = F;
  for (int j = 0; j < 3; ++j) kp.tap_off[j] = a->tap_off[j];
  kp.mode = a->mode;
  kp.L_res = a->L_res;
  kp.resid_off = a->resid_off;
  kp.zx_div = a->zx_div > 0 ? a->zx_div : 1;
  kp.has_out = has_out;
  kp.has_out2 = has_out2;
  kp.bias = a->bias;
  kp.resid_hi = reinterpret_cast<const __nv_bfloat16*>(a->resid_hi);
  kp.resid_lo = reinterpret_cast<const __nv_bfloat16*>(a->resid_lo);
  kp.mult_hi = reinterpret_cast<const __nv_bfloat16*>(a->mult_hi);
  kp.mult_lo = reinterpret_cast<const __nv_bfloat16*>(a->mult_lo);
  kp.mask_in = reinterpret_cast<const unsigned long long*>(a->mask_in);
  kp.mask_out = reinterpret_cast<unsigned long long*>(a->mask_out);
  kp.zx_in = a->zx_in;
  kp.z_out = a->z_out;

  const int max_smem = max_smem_optin();
  BP_REQUIRE(max_smem > 0, "bpb_conv3_tc: no CUDA device");
  const size_t fixed = 1024 /*align slack*/ + static_cast<size_t>(has_out + has_out2) * n_planes * kATileBytes +
                       static_cast<size_t>(F) * 4 + (2 * kMaxStages + 9) * 8 + 16;
  const size_t b_tile = static_cast<size_t>(NT) * 128;
  const size_t w_res = static_cast<size_t>(n_planes) * 3 * kp.NC * b_tile;
  bool resident = (kp.n_ntiles == 1) && (NT <= 128) &&
                  (fixed + w_res + 4 * static_cast<size_t>(kATileBytes) <= static_cast<size_t>(max_smem));
  const size_t stage_bytes = kATileBytes + (resident ? 0 : b_tile);
  const size_t avail = static_cast<size_t>(max_smem) - fixed - (resident ? w_res : 0);
  int stages = static_cast<int>(avail / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  BP_REQUIRE(stages >= 2, "bpb_conv3_tc: not enough shared memory for F=%d planes=%d", F, n_planes);
  kp.stages = stages;
  const size_t smem_bytes = fixed + (resident ? w_res : 0) + stages * stage_bytes;

  CUtensorMap maps[7];
  memset(maps, 0, sizeof(maps));
  const uint64_t wrows = static_cast<uint64_t>(3) * F;
  if (!make_tmap_3d(&maps[0], a->in_hi, F, a->L_in, a->B, 64, kTileM)) return 3;
  if (n_planes == 2 && !make_tmap_3d(&maps[1], a->in_lo, F, a->L_in, a->B, 64, kTileM)) return 3;
  // Weights: planes are separate allocations; the kernel addresses plane wp at row offset
  // wp*3*F, so the hi/lo planes must be contiguous when both are used.
  if (n_planes == 2)
    BP_REQUIRE(a->w_lo == a->w_hi + static_cast<size_t>(3) * F * F,
               "bpb_conv3_tc: w_lo must directly follow w_hi in memory");
  if (!make_tmap_2d(&maps[2], a->w_hi, F, wrows * n_planes, 64, NT)) return 3;
  if (has_out) {
    if (!make_tmap_3d(&maps[3], a->out_hi, F, a->L_out, a->B, 64, kTileM)) return 3;
    if (n_planes == 2 && !make_tmap_3d(&maps[4], a->out_lo, F, a->L_out, a->B, 64, kTileM)) return 3;
  }
  if (has_out2) {
    if (!make_tmap_3d(&maps[5], a->out2_hi, F, a->L_out, a->B, 64, kTileM)) return 3;
    if (n_planes == 2 && !make_tmap_3d(&maps[6], a->out2_lo, F, a->L_out, a->B, 64, kTileM)) return 3;
  }

  const int total_tiles = kp.B * kp.tiles_per_seq * kp.n_ntiles;
  int grid = sm_count();
  if (grid > total_tiles) grid = total_tiles;

  if (NT == 64) {
    return resident ? launch_conv3<64, true>(maps, kp, smem_bytes, grid, stream)
                    : launch_conv3<64, false>(maps, kp, smem_bytes, grid, stream);
  } else if (NT == 128) {
    return resident ? launch_conv3<128, true>(maps, kp, smem_bytes, grid, stream)
                    : launch_conv3<128, false>(maps, kp, smem_bytes, grid, stream);
  }
  return launch_conv3<256, false>(maps, kp, smem_bytes, grid, stream);
}
