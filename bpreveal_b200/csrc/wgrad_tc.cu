// K7: weight and bias gradient of the width-3 dilated convolution (models.py:92-96 of the
// reference; TensorFlow computes it as Conv2DBackpropFilter + BiasAddGrad).
//
//   dw[j][c][n] = sum_{b,t} act[b, t + j*dil, c] * gz[b,t,n]     db[n] = sum_{b,t} gz[b,t,n]
//
// GEMM view: M = 64 input channels of one (tap, channel-chunk) "unit", N = NT output channels,
// K = positions.  Both operands are the same [128 positions x 64 channels] TMA tiles the forward
// kernel uses, read by tcgen05.mma as MN-major (channels contiguous, positions along K).  Each CTA
// owns an (unit group, N tile) and a slice of the position tiles (split-K), keeps its partial
// sums in TMEM for the whole launch, writes them to a workspace, and a second tiny kernel adds the
// partials in a fixed order (deterministic; no atomics).  db comes from one extra unit whose A
// operand is a constant tile of ones.
#include "common.cuh"
#include "ptx.cuh"

namespace bp {

constexpr int kWgTile = 128 * 128;  // bytes of one [128 x 64] bf16 tile
constexpr int kWgMaxStages = 8;

struct WgradKParams {
  int B, L_out, tiles_per_seq, F, NC, dil;
  int n_planes, n_passes;
  int n_units, UG, n_groups, n_ntiles, n_items, ksplit;
  int stages;
  float* ws;  // [ksplit][3*F*F + F]
};

template <int NT>
__global__ void __launch_bounds__(192, 1)
wgrad3_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmG0, const __grid_constant__ CUtensorMap tmG1,
                 const WgradKParams p) {
  constexpr int NSUB = NT / 64;
  constexpr uint32_t IDESC = umma_idesc_bf16(64, NT, 1, 1);
  constexpr uint32_t TMEM_COLS = 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int stage_tiles = NSUB + p.UG;
  const size_t stage_bytes = static_cast<size_t>(stage_tiles) * kWgTile;
  uint8_t* ones = smem;
  uint8_t* ring = smem + kWgTile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + p.stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWgMaxStages;
  uint64_t* acc_full = bars + 2 * kWgMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int item = blockIdx.x % p.n_items;
  const int split = blockIdx.x / p.n_items;
  const int grp = item / p.n_ntiles;
  const int nt = item % p.n_ntiles;
  const int u0 = grp * p.UG;
  const int nu = min(p.UG, p.n_units - u0);
  const bool do_db = (grp == 0);
  const int kblocks = p.B * p.tiles_per_seq;
  const int my_kb = (kblocks - split + p.ksplit - 1) / p.ksplit;  // split < ksplit <= kblocks
  const int visits = my_kb * p.n_passes;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmG0);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  {
    // constant tile of bf16 ones (layout-invariant), read by the async proxy
    uint4 v;
    v.x = v.y = v.z = v.w = 0x3F803F80u;
    for (int i = threadIdx.x; i < kWgTile / 16; i += blockDim.x)
      reinterpret_cast<uint4*>(ones)[i] = v;
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int v = 0; v < visits; ++v) {
        const int kb = split + (v / p.n_passes) * p.ksplit;
        const int pass = v % p.n_passes;
        const int b = kb / p.tiles_per_seq;
        const int t0 = (kb % p.tiles_per_seq) * 128;
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], static_cast<uint32_t>((NSUB + nu) * kWgTile));
        uint8_t* dst = ring + stage * stage_bytes;
        const CUtensorMap* mg = (pass == 2) ? &tmG1 : &tmG0;
        const CUtensorMap* ma = (pass == 1) ? &tmA1 : &tmA0;
        for (int s = 0; s < NSUB; ++s)
          tma_load_3d(mg, dst + s * kWgTile, &full[stage], nt * NT + s * 64, t0, b);
        for (int i = 0; i < nu; ++i) {
          const int u = u0 + i;
          const int j = u / p.NC, c = u % p.NC;
          tma_load_3d(ma, dst + (NSUB + i) * kWgTile, &full[stage], c * 64, t0 + j * p.dil, b);
        }
        if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      uint32_t db_started = 0;
      for (int v = 0; v < visits; ++v) {
        const int pass = v % p.n_passes;
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t g_addr = smem_u32(ring + stage * stage_bytes);
        for (int i = 0; i < nu; ++i) {
          const uint32_t a_addr = g_addr + (NSUB + i) * kWgTile;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16(tmem_base + i * NT, umma_smem_desc(a_addr + k * 2048, kWgTile, 1024),
                      umma_smem_desc(g_addr + k * 2048, kWgTile, 1024), IDESC,
                      (v | k) != 0 ? 1u : 0u);
        }
        if (do_db && pass != 1) {
          const uint32_t o_addr = smem_u32(ones);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16(tmem_base + p.UG * NT, umma_smem_desc(o_addr + k * 2048, kWgTile, 1024),
                      umma_smem_desc(g_addr + k * 2048, kWgTile, 1024), IDESC,
                      (db_started | k) != 0 ? 1u : 0u);
          db_started = 1;
        }
        umma_commit(&empty[stage]);
        if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1; }
      }
      umma_commit(acc_full);
    }
  } else {
    // epilogue: TMEM partial sums -> workspace slice of this split
    const int q = warp & 3;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float* wsp = p.ws + static_cast<size_t>(split) * (static_cast<size_t>(3) * p.F * p.F + p.F);
    for (int i = 0; i < nu; ++i) {
      const int u = u0 + i;
      const int j = u / p.NC, c = u % p.NC;
      // M=64 accumulators occupy lanes 0-15 of each TMEM quadrant: row = 16*q + lane.
      const int m = 16 * q + lane;
      for (int col = 0; col < NT; col += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + i * NT + col + (static_cast<uint32_t>(q * 32) << 16), r);
        tmem_ld_wait();
        if (lane < 16) {
          float4* dst = reinterpret_cast<float4*>(
              wsp + (static_cast<size_t>(j) * p.F + c * 64 + m) * p.F + nt * NT + col);
#pragma unroll
          for (int x = 0; x < 8; ++x)
            dst[x] = make_float4(__uint_as_float(r[4 * x]), __uint_as_float(r[4 * x + 1]),
                                 __uint_as_float(r[4 * x + 2]), __uint_as_float(r[4 * x + 3]));
        }
      }
    }
    if (do_db && q == 0) {
      for (int col = 0; col < NT; col += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + p.UG * NT + col, r);
        tmem_ld_wait();
        if (lane == 0) {
          float* dst = wsp + static_cast<size_t>(3) * p.F * p.F + nt * NT + col;
#pragma unroll
          for (int x = 0; x < 32; ++x) dst[x] = __uint_as_float(r[x]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// out[i] = sum_s ws[s][i] in a fixed order (deterministic).  One block owns 32 consecutive outputs;
// warp w adds the splits s = w, w + 8, ... (each a coalesced 128-byte read, four independent
// chains), then the eight warp partials are added in warp order.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ ws, int ksplit, long long n_dw, long long n_db,
                    float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float part[8][32];
  const long long n = n_dw + n_db;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long i = static_cast<long long>(blockIdx.x) * 32 + lane;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (i < n) {
    int s = warp;
    for (; s + 24 < ksplit; s += 32) {
      s0 += ws[(s + 0) * n + i];
      s1 += ws[(s + 8) * n + i];
      s2 += ws[(s + 16) * n + i];
      s3 += ws[(s + 24) * n + i];
    }
    for (; s < ksplit; s += 8) s0 += ws[s * n + i];
  }
  part[warp][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (warp == 0 && i < n) {
    float tot = part[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) tot += part[w][lane];
    if (i < n_dw) dw[i] = tot;
    else if (db) db[i - n_dw] = tot;
  }
}

struct WgradPlan {
  int NT, UG, n_units, n_groups, n_ntiles, n_items, ksplit, stages;
  size_t smem_bytes, ws_bytes;
};

static bool plan_wgrad(int B, int L_out, int F, int n_planes, WgradPlan* pl) {
  const int NT = (F % 256 == 0) ? 256 : (F % 128 == 0 ? 128 : 64);
  pl->NT = NT;
  pl->n_units = 3 * (F / 64);
  int ug = 512 / NT - 1;  // one accumulator slot is reserved for the bias-gradient unit
  if (ug > pl->n_units) ug = pl->n_units;
  if (ug > 3) ug = 3;
  pl->UG = ug;
  pl->n_groups = (pl->n_units + ug - 1) / ug;
  pl->n_ntiles = F / NT;
  pl->n_items = pl->n_groups * pl->n_ntiles;
  const int kblocks = B * ((L_out + 127) / 128);
  int ks = sm_count() / pl->n_items;
  if (ks < 1) ks = 1;
  if (ks > kblocks) ks = kblocks;
  pl->ksplit = ks;
  const size_t stage_bytes = static_cast<size_t>(NT / 64 + ug) * kWgTile;
  const size_t fixed = 1024 + kWgTile + (2 * kWgMaxStages + 2) * 8 + 16;
  const int max_smem = max_smem_optin();
  int stages = static_cast<int>((static_cast<size_t>(max_smem) - fixed) / stage_bytes);
  if (stages > kWgMaxStages) stages = kWgMaxStages;
  if (stages < 1) return false;
  pl->stages = stages;
  pl->smem_bytes = fixed + stages * stage_bytes;
  pl->ws_bytes = static_cast<size_t>(ks) * (static_cast<size_t>(3) * F * F + F) * sizeof(float);
  (void)n_planes;
  return true;
}

template <int NT>
static int launch_wgrad(const CUtensorMap* maps, const WgradKParams& kp, size_t smem, int grid,
                        cudaStream_t stream) {
  auto kern = wgrad3_tc_kernel<NT>;
  static bool configured = false;
  if (!configured) {
    BP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       max_smem_optin()));
    configured = true;
  }
  kern<<<grid, 192, smem, stream>>>(maps[0], maps[1], maps[2], maps[3], kp);
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace bp

extern "C" long long bpb_wgrad3_ws_bytes(int B, int L_out, int F, int precise) {
  bp::WgradPlan pl;
  if (bp::max_smem_optin() <= 0) {
    bp::set_error("bpb_wgrad3_ws_bytes: no CUDA device");
    return -1;
  }
  if (F <= 0 || F % 64 != 0 || !bp::plan_wgrad(B, L_out, F, precise ? 2 : 1, &pl)) {
    bp::set_error("bpb_wgrad3_ws_bytes: unsupported F=%d", F);
    return -1;
  }
  return static_cast<long long>(pl.ws_bytes);
}

extern "C" int bpb_wgrad3_tc(const bpb_wgrad3_args* a, void* stream_) {
  using namespace bp;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(a != nullptr, "bpb_wgrad3_tc: null args");
  BP_REQUIRE(a->F > 0 && a->F % 64 == 0, "bpb_wgrad3_tc: F=%d must be a positive multiple of 64",
             a->F);
  BP_REQUIRE(a->act_hi && a->gz_hi && a->dw && a->ws, "bpb_wgrad3_tc: NULL tensor");
  BP_REQUIRE((a->act_lo != nullptr) == (a->gz_lo != nullptr),
             "bpb_wgrad3_tc: act_lo and gz_lo must both be given or both be NULL");
  BP_REQUIRE(max_smem_optin() > 0, "bpb_wgrad3_tc: no CUDA device");
  const int n_planes = a->act_lo ? 2 : 1;
  WgradPlan pl;
  BP_REQUIRE(plan_wgrad(a->B, a->L_out, a->F, n_planes, &pl), "bpb_wgrad3_tc: cannot plan F=%d",
             a->F);
  BP_REQUIRE(static_cast<size_t>(a->ws_bytes) >= pl.ws_bytes,
             "bpb_wgrad3_tc: workspace too small (%lld < %zu)", a->ws_bytes, pl.ws_bytes);

  WgradKParams kp{};
  kp.B = a->B;
  kp.L_out = a->L_out;
  kp.tiles_per_seq = (a->L_out + 127) / 128;
  kp.F = a->F;
  kp.NC = a->F / 64;
  kp.dil = a->dil;
  kp.n_planes = n_planes;
  kp.n_passes = n_planes == 2 ? 3 : 1;
  kp.n_units = pl.n_units;
  kp.UG = pl.UG;
  kp.n_groups = pl.n_groups;
  kp.n_ntiles = pl.n_ntiles;
  kp.n_items = pl.n_items;
  kp.ksplit = pl.ksplit;
  kp.stages = pl.stages;
  kp.ws = a->ws;

  CUtensorMap maps[4];
  memset(maps, 0, sizeof(maps));
  if (!make_tmap_3d(&maps[0], a->act_hi, a->F, a->L_in, a->B, 64, 128)) return 3;
  if (n_planes == 2 && !make_tmap_3d(&maps[1], a->act_lo, a->F, a->L_in, a->B, 64, 128)) return 3;
  if (!make_tmap_3d(&maps[2], a->gz_hi, a->F, a->L_out, a->B, 64, 128)) return 3;
  if (n_planes == 2 && !make_tmap_3d(&maps[3], a->gz_lo, a->F, a->L_out, a->B, 64, 128)) return 3;

  const int grid = pl.n_items * pl.ksplit;
  int rc;
  if (pl.NT == 64) rc = launch_wgrad<64>(maps, kp, pl.smem_bytes, grid, stream);
  else if (pl.NT == 128) rc = launch_wgrad<128>(maps, kp, pl.smem_bytes, grid, stream);
  else rc = launch_wgrad<256>(maps, kp, pl.smem_bytes, grid, stream);
  if (rc) return rc;

  const long long n_dw = static_cast<long long>(3) * a->F * a->F;
  const long long n_db = a->F;
  const int blocks = static_cast<int>((n_dw + n_db + 31) / 32);
  wgrad_reduce_kernel<<<blocks, 256, 0, stream>>>(a->ws, pl.ksplit, n_dw, n_db, a->dw, a->db);
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}
