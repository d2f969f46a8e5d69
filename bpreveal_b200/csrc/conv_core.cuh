// The tcgen05 implicit-GEMM convolution engine shared by every position-major contraction of the
// network: the width-3 dilated tower layers (K2 / K6 / K8), the one-hot input convolution (K1) and
// the data gradient of the profile heads.
//
// Replaces, for models.py:92-104 of the reference, the chain SpaceToBatchND -> Conv2D ->
// BatchToSpaceND -> BiasAdd -> Relu -> StridedSlice (Cropping1D) -> AddV2 that TensorFlow runs per
// dilated layer, its data gradient, and the DeepSHAP-rescale variant of that
// gradient (shap.py:612-639).
//
// GEMM view: M = 128 consecutive positions of one sequence, N = NT output channels, K = 3 taps x
// F input channels.  Rows outside the sequence are zero-filled by TMA, which gives the 'valid'
// forward conv and the 'full' backward conv from the same code.
//
// Two operand-staging schemes:
//   HALO   (small dilations, weights resident): ONE TMA box of 128 + span rows per tile and
//          channel chunk; the three taps are UMMA descriptors that start at row offsets
//          0 / d / 2d of that box (SWIZZLE_128B is a function of the absolute smem address, so a
//          descriptor may start at any row).  Activations are read from L2 once instead of three
//          times, and the forward residual a[t+d] is read back from the same box.
//   SLAB   (any dilation / width): one [128 x 64] box per (tap, chunk, pass), as a classic
//          multi-stage K pipeline; weights resident in smem when they fit, else streamed.
//
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, then EG groups of eight
// epilogue warps (two groups on the bf16 path: 576 threads); streamed-weight launches add a second
// producer warp for the weight tiles, the masked-operand backward (XF) six masker warps.
// Epilogue warp w reads TMEM lane quadrant w % 4 (rows) and one 32-column half of each
// 64-channel chunk, so two warps per scheduler hide each other's latency.  The epilogue fuses
// bias / ReLU / ReLU-mask bits / crop + residual (forward) or residual gradient + mask / rescale
// multiplier (backward) and writes bf16 planes through per-warp swizzled staging tiles and TMA
// stores.  The residual comes from the HALO box (forward), from its own TMA ring (F = 64
// backward) or straight from global memory (SLAB-staged forward, every streamed-weight launch).
// F >= 256 training launches run as CTA pairs (cta_group::2, template flag CG2).
#pragma once
#include "common.cuh"
#include "ptx.cuh"

#include <cstdlib>
#include <cstring>

namespace bp {

constexpr int kTileM = 128;
constexpr int kTileBytes = kTileM * 128;  // [128 x 64] bf16
constexpr int kMaxStages = 12;
constexpr int kMaxRStages = 6;
constexpr int kMaxAcc = 6;
constexpr int kBiasTileBytes = 64 * 128;
constexpr int kBiasMmaBytes = kBiasTileBytes + 128 * 128;  // bias tile + tile of ones
constexpr int kEpiThreads = 256;

enum { EPI_FWD = 0, EPI_FWD_Z = 1, EPI_FWD_R = 2, EPI_BWD_MASK = 3, EPI_BWD_MULT = 4 };

struct ConvKParams {
  int B, L_out, tiles_per_seq, n_ntiles, F;
  // K structure: n_taps taps x KC chunks of 64 input channels; the last chunk of every tap issues
  // last_nk (1..4) K=16 MMAs (narrow inputs such as the 8 x 25 one-hot window).
  int n_taps, KC, last_nk;
  // operand planes: pass i multiplies A plane (pa_bits >> i) & 1 with W plane (pw_bits >> 2i) & 3
  int npass, pa_bits, pw_bits, n_aplanes, n_wplanes;
  int step_b, step_mt, step_nt;  // gridDim.x tiles expressed as (sequences, M tiles, N tiles)
  // ceil(2^32 / n_ntiles) and ceil(2^32 / tiles_per_seq) (0 when the divisor is 1): the first
  // tile of a CTA is found with two multiply-highs instead of three integer divisions, which
  // sat on the single-thread start-up chain of every launch
  unsigned inv_ntiles, inv_tiles_per_seq;
  // STRIP staging (window convolutions over a 16-byte-per-position tensor, resident weights): the
  // A operand of a tile is the raw strip of 127 + 2*nk16 positions; a NON-swizzled K-major UMMA
  // descriptor (LBO 16 B, SBO 128 B) reads the implicit im2col matrix straight out of it.
  int strip;         // 1: STRIP staging
  int strip_bytes;   // bytes copied per (tile, A plane)
  int strip_stride;  // the same padded to 1024
  int nk16;          // K = 16 * nk16
  long long a_batch_bytes;      // STRIP: byte stride between sequences
  const uint8_t* a_base[2];     // STRIP: the A planes
  int tap_row[3];  // SLAB: row offset of tap j relative to t0; HALO: row of tap j inside the box
  int box_off;     // HALO: first box row = t0 + box_off
  int sub_bytes;   // HALO: bytes of one (chunk, plane) box, padded to 1024
  int box_bytes;   // HALO: bytes TMA writes per box (rows * 128)
  int stages, rstages, nstg;
  int resid_off, resid_row;  // residual tile = rows t0 + resid_off (resid_row: inside the box)
  int zx_div;
  int has_out, has_out2;
  int epi_groups, ring_resid;
  // 1: the bias enters the accumulator through one extra K = 16 MMA per tile (a tile of ones
  // times a tile holding the bias as three bf16 remainders in its first K columns) instead of 32
  // FADDs and 8 shared-memory reads per epilogue thread
  int bias_mma;
  const float* bias;
  int bias_n;  // valid entries of bias (the rest of the F channels get 0)
  const __nv_bfloat16 *mult_hi, *mult_lo;
  const uint32_t* mask_in;
  uint32_t* mask_out;
  const float* zx_in;
  float* z_out;
  // XF (backward, bf16 path, F = 64): the A operand is `in (.) mask_a`.  The tiles TMA delivers
  // are ANDed with the 1-bit ReLU mask of their rows IN SHARED MEMORY by six extra warps before
  // the MMA reads them, so the caller passes the unmasked gradient g of the layer above and that
  // layer's mask instead of a materialised gz = g (.) mask plane (one write and one read of a
  // whole activation plane per layer).  mask_a: [B][L_a] uint64 (F = 64: one word per row); the
  // words of a tile's rows arrive by a 1-D bulk copy on the same mbarrier as the tile itself.
  const uint8_t* mask_a;
  int L_a;                     // rows per sequence of the A operand
  long long mask_rows_total;   // B * L_a (even)
  int xf_rows;                 // rows of one staged A tile (HALO: the box, SLAB: 128)
  int mstage;                  // bytes of one stage's mask words (multiple of 16)
  // Streamed-weight launches (F >= 256) read the residual straight from global memory in the
  // epilogue instead of through a TMA ring: [B][L_res][F] bf16 planes
  const uint8_t* resid_ptr[2];
  int L_res;
  int cg2;                     // 1: CTA pairs (cta_group::2), tiles_per_seq counts 256-row pair tiles
  int dbg_flags;               // PROFILE builds only (BPB_DBG_FLAGS): 1 no stores, 2 no residual
                               // loads, 4 no masking work -- ablations for the role profile
  unsigned* dbg;
};

constexpr int kXfThreads = 192;  // the six masker warps of an XF launch

struct ConvMaps {
  CUtensorMap A[2], W, O[2], P[2], R[2];
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
__device__ __forceinline__ uint4 lds128(const uint8_t* p) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(smem_u32(p)));
  return v;
}
__device__ __forceinline__ void sts128(uint8_t* p, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p)), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

__device__ __forceinline__ float4 lds_f4(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(smem_u32(p)));
  return v;
}
// 1.0f where v > 0 else 0.0f (one FSET.BF on the ALU pipe).  The ReLU, the residual add and the
// mask-bit accumulation then run on the FMA pipes: relu(v) + r == fma(v, flag, r) for finite v, and
// the 32 mask bits are summed as flag * 2^k in two fp32 accumulators (16 bits each, exact).
__device__ __forceinline__ float gt0_flag(float v) {
  float f;
  asm("set.gt.f32.f32 %0, %1, 0f00000000;" : "=f"(f) : "f"(v));
  return f;
}

// Persistent tile walk without per-tile integer divisions: tile = (b, mt, nt), nt fastest,
// advanced by gridDim.x tiles per iteration.
struct TileWalk {
  int b, mt, nt;
  __device__ __forceinline__ void init(const ConvKParams& p, unsigned first = blockIdx.x) {
    const unsigned tile = first;  // < 2^16, divisors < 2^16: the multiply-high is exact
    const unsigned t1 = p.inv_ntiles ? __umulhi(tile, p.inv_ntiles) : tile;
    nt = static_cast<int>(tile - t1 * static_cast<unsigned>(p.n_ntiles));
    const unsigned bb = p.inv_tiles_per_seq ? __umulhi(t1, p.inv_tiles_per_seq) : t1;
    mt = static_cast<int>(t1 - bb * static_cast<unsigned>(p.tiles_per_seq));
    b = static_cast<int>(bb);
  }
  __device__ __forceinline__ void next(const ConvKParams& p) {
    nt += p.step_nt;
    if (nt >= p.n_ntiles) { nt -= p.n_ntiles; ++mt; }
    mt += p.step_mt;
    if (mt >= p.tiles_per_seq) { mt -= p.tiles_per_seq; ++b; }
    b += p.step_b;
  }
};

// ring index += n (mod size), flipping the phase bit at every wrap
__device__ __forceinline__ void ring_advance(uint32_t& idx, uint32_t& phase, uint32_t n,
                                             uint32_t size) {
  idx += n;
  while (idx >= size) { idx -= size; phase ^= 1; }
}

// XF: the mask words of the A rows [row0, row0 + xf_rows) of sequence b as one 16-byte aligned
// 1-D bulk copy.  Row g of the flattened [B * L_a] mask lands at byte (g - (g0 & ~1)) * 8 of the
// stage's buffer, so shared and global addresses agree modulo 16; rows outside the tensor are
// not copied (their data rows are zero-filled by TMA, whatever bits the buffer holds).
struct XfCopy {
  long long src_row;
  uint32_t dst_off, bytes;
};
__device__ __forceinline__ XfCopy xf_copy(const ConvKParams& p, int b, int row0) {
  const long long g0 = static_cast<long long>(b) * p.L_a + row0;
  const long long g0e = g0 & ~1LL;
  const long long cs = (g0 > 0 ? g0 : 0) & ~1LL;
  long long ce = g0 + p.xf_rows;
  if (ce > p.mask_rows_total) ce = p.mask_rows_total;
  ce = (ce + 1) & ~1LL;
  XfCopy c;
  c.src_row = cs;
  c.dst_off = static_cast<uint32_t>((cs - g0e) * 8);
  c.bytes = ce > cs ? static_cast<uint32_t>((ce - cs) * 8) : 0u;
  return c;
}

template <int EPI, int NT, bool RESIDENT, int PLANES, bool HALO, int EG, bool RESID, bool XF = false,
          bool CG2 = false>
__global__ void __launch_bounds__(64 + kEpiThreads * EG + (XF ? kXfThreads : 0) +
                                  (RESIDENT ? 0 : 32), 1)
conv_tc_kernel(const __grid_constant__ ConvMaps tm, const ConvKParams p) {
  // CG2: CTA pairs (cluster of two) issue M = 256 MMAs on streamed weights: each CTA stages its
  // own 128 rows of A and HALF of the weight tile, so the shared-memory pipe moves 8 KB instead
  // of 12 KB per MMA (the bound of the F = 512 launches).  The even CTA (the
  // leader) owns the operand-full and accumulator-empty barriers and issues; both CTAs run
  // producers and epilogues on their own half of the pair's 256 rows.
  static_assert(!CG2 || (!RESIDENT && NT == 256 && PLANES == 1 && !HALO && !XF && EG == 2 && RESID),
                "CG2: streamed-weight bf16 tower layers");
  const uint32_t crank = CG2 ? cluster_ctarank() : 0u;
  // CG2: the pair walks pair-tiles of 256 rows (p.tiles_per_seq counts those); this CTA's half
  const unsigned first_tile = CG2 ? (blockIdx.x >> 1) : blockIdx.x;
  auto tile_row0 = [crank](int mt) -> int {
    return (CG2 ? 2 * mt + static_cast<int>(crank) : mt) * kTileM;
  };
  static_assert(!XF || (EPI == EPI_BWD_MASK && NT == 64 && RESIDENT && PLANES == 1 && RESID),
                "XF: bf16 backward of an F = 64 layer with resident weights");
  // accumulator stages: a multiple of the group count, so a stage (and its mbarriers) always
  // belongs to the same epilogue group
  constexpr int ACC = (EG == 3) ? ((NT <= 64) ? 6 : 3) : ((NT <= 128) ? 4 : 2);
  static_assert(ACC <= kMaxAcc && ACC * NT <= 512, "TMEM accumulator stages");
  constexpr int NCH = NT / 64;
  constexpr uint32_t TMEM_COLS = (ACC * NT <= 128) ? 128 : (ACC * NT <= 256 ? 256 : 512);
  constexpr int B_TILE_BYTES = (CG2 ? NT / 2 : NT) * 128;  // CG2: this CTA's half of the tile
  constexpr int SLAB_BYTES = kTileBytes + (RESIDENT ? 0 : B_TILE_BYTES);
  constexpr uint32_t IDESC = umma_idesc_bf16(CG2 ? 256 : 128, NT, 0, 0);
  constexpr bool IS_FWD = (EPI <= EPI_FWD_R);
  // the forward residual a[t + d] already sits in the HALO box
  // Streamed weights: NO residual ring.  A tile of NT = 256 columns needs four residual chunks;
  // a two-slot sub-ring per epilogue group made the producer wait for the epilogue of tile i
  // (which starts when the MMAs of tile i are done) before it could issue the first operand
  // loads of tile i + 1: a bubble of two chunk visits + one load latency per tile -- the F = 512
  // data gradient sat at 61 % of the tensor rate because of it.  The epilogue has cycles to
  // spare there (a tile's MMAs take ~14 k cycles, a chunk visit ~3 k), so it reads its residual
  // rows with plain global loads issued at the top of the visit; the 64 KB of ring become
  // operand stages.
  // The SLAB-staged forward with resident weights (F = 64, d >= 128) does the same: 100.2 ->
  // 98.3 us over the nine forward launches of C1.  On the F = 64 data gradient it measured
  // neutral with more spills (-DBPB_DIRECT_ALL), so that keeps its ring; on the HALO-staged
  // forward, whose residual is already in the box, it was slower (98.5 -> 112.6 us).
#if defined(BPB_DIRECT_ALL)
  constexpr bool DIRECT_RESID = RESID && !(HALO && IS_FWD);
#else
  constexpr bool DIRECT_RESID = RESID && (!RESIDENT || (IS_FWD && !HALO));
#endif
  constexpr bool RING_RESID = RESID && !(HALO && IS_FWD) && !DIRECT_RESID;
  static_assert(!HALO || RESIDENT, "HALO staging needs resident weights");
  static_assert(!HALO || RESID || !IS_FWD, "HALO forward is the residual layer");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));

  const int w_tiles = RESIDENT ? p.n_wplanes * p.n_taps * p.KC : 0;
  const int stage_bytes = HALO ? p.KC * p.n_aplanes * p.sub_bytes
                               : (p.strip ? p.n_aplanes * p.strip_stride : SLAB_BYTES);
  const int n_out = p.has_out + p.has_out2;
  const int stg_bytes = n_out * PLANES * kTileBytes;
  const int rstage_bytes = PLANES * kTileBytes;
  constexpr bool CAN_BIAS_MMA = IS_FWD && RESIDENT && NT == 64;
  const bool bias_mma = CAN_BIAS_MMA && p.bias_mma != 0;
  uint8_t* w_smem = smem;
  uint8_t* bias_tile = w_smem + static_cast<size_t>(w_tiles) * B_TILE_BYTES;  // [64 n][64 k]
  uint8_t* ones_tile = bias_tile + kBiasTileBytes;                            // [128 m][64 k]
  uint8_t* ring = bias_tile + (bias_mma ? kBiasMmaBytes : 0);
  uint8_t* stg = ring + static_cast<size_t>(p.stages) * stage_bytes;
  uint8_t* rring = stg + static_cast<size_t>(p.nstg) * stg_bytes;
  float* s_bias = reinterpret_cast<float*>(
      rring + (RING_RESID ? static_cast<size_t>(p.rstages) * rstage_bytes : 0));
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + p.F);
  uint64_t* full = bars;
  uint64_t* empty = full + kMaxStages;
  uint64_t* acc_full = empty + kMaxStages;
  uint64_t* acc_empty = acc_full + kMaxAcc;
  uint64_t* rfull = acc_empty + kMaxAcc;
  uint64_t* rempty = rfull + kMaxRStages;
  uint64_t* w_full = rempty + kMaxRStages;
  uint64_t* b_full = w_full + 1;
  uint64_t* xfull = b_full + 1;  // XF: "stage masked, the MMA may read it"
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xfull + kMaxStages);
  // XF: per-stage mask words, 16-byte aligned (bulk-copy destination)
  uint8_t* mring = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~static_cast<uintptr_t>(15));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // let the next kernel of the stream start its prologue as soon as this grid's CTAs retire
  pdl_launch_dependents();
#ifdef BPB_ROLE_PROFILE
  const long long t_kernel_start = clock64();
  if (threadIdx.x < 32) prof_smem()[threadIdx.x] = 0;  // published by the __syncthreads below
#endif

  if (warp == 0) {
    // one lane per barrier (pair): the prologue is on the critical path of every launch
    static_assert(kMaxStages <= 16 && kMaxAcc <= 8 && kMaxRStages <= 6,
                  "lane map of the barrier initialisation");
    // branch-free lane map: lanes 0..15 ring stages, 16..19 accumulators, 24..27 residual ring
    {
      const bool is_ring = lane < p.stages;
      const bool is_acc = lane >= 16 && lane < 16 + ACC;
      const bool is_res = lane >= 24 && lane < 24 + kMaxRStages;
      uint64_t* b0 = is_ring ? &full[lane] : is_acc ? &acc_full[lane - 16] : &rfull[lane - 24];
      uint64_t* b1 = is_ring ? &empty[lane] : is_acc ? &acc_empty[lane - 16] : &rempty[lane - 24];
      // HALO forward: the box is released by the MMA commit AND by the 8 epilogue warps
      // (CG2: the leader's accumulator-empty barriers collect the epilogue warps of both CTAs)
      const uint32_t c1 = is_ring ? ((HALO && IS_FWD && RESID && !DIRECT_RESID) ? 9u : 1u)
                                  : ((CG2 && is_acc) ? 16u : 8u);
      if (is_ring || is_acc || is_res) {
        mbar_init(b0, 1);
        mbar_init(b1, c1);
      }
      if (XF && is_ring) mbar_init(&xfull[lane], kXfThreads / 32);
    }
    if (lane == 30) mbar_init(b_full, 1);
    if (lane == 31) {
      tma_prefetch_desc(&tm.A[0]);
      tma_prefetch_desc(&tm.W);
      if (p.n_aplanes == 2) tma_prefetch_desc(&tm.A[1]);
      if (p.has_out) tma_prefetch_desc(&tm.O[0]);
      if (p.has_out2) tma_prefetch_desc(&tm.P[0]);
      if (RING_RESID) tma_prefetch_desc(&tm.R[0]);
    }
    fence_barrier_init();
    // Split barrier: the producer needs nothing from the other warps (not even the TMEM base),
    // so it only ARRIVES and issues its first loads while TMEM is still being allocated; the
    // memory latency of the first tile is the longest link of the start-up chain.
    __syncwarp();
    named_bar_arrive(8, blockDim.x);
  } else {
    if (warp == 1) {
      // The resident weights are fetched by the MMA warp (their only reader), off the
      // producer's chain: they are parameters, never written by an in-flight predecessor.
      if (elect_one()) {
        mbar_init(w_full, 1);
        fence_barrier_init();
        if (RESIDENT) {
          mbar_expect_tx(w_full, static_cast<uint32_t>(w_tiles) * B_TILE_BYTES);
          for (int wp = 0; wp < p.n_wplanes; ++wp)
            for (int j = 0; j < p.n_taps; ++j)
              for (int c = 0; c < p.KC; ++c)
                tma_load_2d(&tm.W,
                            w_smem + static_cast<size_t>((wp * p.n_taps + j) * p.KC + c) * B_TILE_BYTES,
                            w_full, c * 64, (wp * p.n_taps + j) * p.F);
        }
      }
      __syncwarp();
      if (CG2) tmem_alloc_cg2<TMEM_COLS>(tmem_slot);
      else tmem_alloc<TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    named_bar_sync(8, blockDim.x);
    tc_fence_after();
  }
  // CG2: no remote arrive / remotely signalled load before both CTAs initialised their barriers
  if (CG2) cluster_sync_all();
  const uint32_t tmem_base = (warp == 0) ? 0u : *tmem_slot;
#ifdef BPB_ROLE_PROFILE
#define BPB_STAMP(slot) prof_smem()[slot] = static_cast<unsigned>(clock64() - t_kernel_start)
  if (threadIdx.x == 0) BPB_STAMP(20);
#else
#define BPB_STAMP(slot)
#endif

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    if (elect_one()) {
      // the activations are written by the predecessor grid: wait for it before the first load
      BPB_STAMP(25);
      pdl_wait();
      BPB_STAMP(21);
      // The residual ring is split into EG sub-rings, one per epilogue group (tile i belongs to
      // group i % EG): a slot is then always consumed by the same warps, which observe every
      // phase of its mbarrier.  (A consumer that skipped phases cannot tell "the phase I skipped
      // is still in flight" from "my phase completed" -- seen as a rare deadlock.)
      const uint32_t RG = static_cast<uint32_t>(p.rstages) / EG;
      uint32_t stage = 0, phase = 0, ntile = 0;
      uint32_t rsg[3] = {0, 0, 0}, rphg[3] = {0, 0, 0};
      TileWalk tw;
      for (tw.init(p, first_tile); tw.b < p.B; tw.next(p), ++ntile) {
        const int nt = tw.nt, b = tw.b;
        const int t0 = tile_row0(tw.mt);
        if (HALO) {
          if (ntile == 0) BPB_STAMP(29);
          mbar_wait_wd(&empty[stage], phase ^ 1, p.dbg, 0x100 + stage, ntile);
          XfCopy xc{0, 0u, 0u};
          if (XF) xc = xf_copy(p, b, t0 + p.box_off);
          mbar_expect_tx(&full[stage],
                         static_cast<uint32_t>(p.KC * p.n_aplanes * p.box_bytes) + xc.bytes);
          if (ntile == 0) BPB_STAMP(30);
          uint8_t* dst = ring + static_cast<size_t>(stage) * stage_bytes;
          for (int c = 0; c < p.KC; ++c)
            for (int pl = 0; pl < p.n_aplanes; ++pl)
              tma_load_3d(&tm.A[pl], dst + static_cast<size_t>(c * p.n_aplanes + pl) * p.sub_bytes,
                          &full[stage], c * 64, t0 + p.box_off, b);
          if (XF && xc.bytes)
            bulk_load_1d(mring + static_cast<size_t>(stage) * p.mstage + xc.dst_off,
                         p.mask_a + xc.src_row * 8, xc.bytes, &full[stage]);
          if (ntile == 0) BPB_STAMP(26);
          if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1; }
        } else if (RESIDENT && p.strip) {
          mbar_wait_wd(&empty[stage], phase ^ 1, p.dbg, 0x100 + stage, ntile);
          mbar_expect_tx(&full[stage], static_cast<uint32_t>(p.n_aplanes * p.strip_bytes));
          uint8_t* dst = ring + static_cast<size_t>(stage) * stage_bytes;
          for (int pl = 0; pl < p.n_aplanes; ++pl)
            bulk_load_1d(dst + pl * p.strip_stride,
                         p.a_base[pl] + static_cast<size_t>(b) * p.a_batch_bytes +
                             static_cast<size_t>(t0) * 16,
                         static_cast<uint32_t>(p.strip_bytes), &full[stage]);
          if (ntile == 0) BPB_STAMP(26);
          if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1; }
        } else {
          // Resident weights: one A tile per (tap, chunk, A plane), shared by every pass that
          // multiplies that plane; streamed weights: one (A, W) tile pair per pass.
          const int inner = RESIDENT ? p.n_aplanes : p.npass;
          const int nsteps = p.n_taps * p.KC * inner;
          for (int step = 0; step < nsteps; ++step) {
            const int pass = step % inner;  // RESIDENT: the A plane
            const int c = (step / inner) % p.KC;
            const int j = step / (inner * p.KC);
            mbar_wait_wd(&empty[stage], phase ^ 1, p.dbg, 0x100 + stage, ntile);
            XfCopy xc{0, 0u, 0u};
            if (XF) xc = xf_copy(p, b, t0 + p.tap_row[j]);
            uint8_t* dst = ring + static_cast<size_t>(stage) * SLAB_BYTES;
            if (CG2) {
              // the leader arms its barrier with the bytes of both CTAs; a completion of the
              // peer that lands first only drives the pending count negative for a moment
              if (crank == 0) mbar_expect_tx(&full[stage], 2 * SLAB_BYTES);
              tma_load_3d_cg2(&tm.A[(p.pa_bits >> pass) & 1], dst, &full[stage], c * 64,
                              t0 + p.tap_row[j], b);
            } else {
              mbar_expect_tx(&full[stage], SLAB_BYTES + xc.bytes);
              tma_load_3d(&tm.A[RESIDENT ? pass : ((p.pa_bits >> pass) & 1)], dst, &full[stage],
                          c * 64, t0 + p.tap_row[j], b);
            }
            if (XF && xc.bytes)
              bulk_load_1d(mring + static_cast<size_t>(stage) * p.mstage + xc.dst_off,
                           p.mask_a + xc.src_row * 8, xc.bytes, &full[stage]);
            // (!RESIDENT: the weight tile of this stage is fetched by the second producer warp)
            if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1; }
          }
        }
#ifdef BPB_ROLE_PROFILE
        if (RING_RESID && !(p.dbg_flags & 2)) {
#else
        if (RING_RESID) {
#endif
          const uint32_t g = (EG > 1) ? (ntile % EG) : 0u;
          for (int ch = 0; ch < NCH; ++ch) {
            const uint32_t rs = g * RG + rsg[g];
            const uint32_t rphase = rphg[g];
            mbar_wait_wd(&rempty[rs], rphase ^ 1, p.dbg, 0x200 + rs, ntile);
            mbar_expect_tx(&rfull[rs], static_cast<uint32_t>(rstage_bytes));
            uint8_t* dst = rring + static_cast<size_t>(rs) * rstage_bytes;
            for (int pl = 0; pl < PLANES; ++pl)
              tma_load_3d(&tm.R[pl], dst + pl * kTileBytes, &rfull[rs], nt * NT + ch * 64,
                          t0 + p.resid_off, b);
            if (++rsg[g] == RG) { rsg[g] = 0; rphg[g] ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer (CG2: the leader's)
    if ((!CG2 || crank == 0) && elect_one()) {
      if (RESIDENT) mbar_wait_wd(w_full, 0, p.dbg, 0x300);
      if (bias_mma) mbar_wait_wd(b_full, 0, p.dbg, 0x310);
      BPB_STAMP(28);
      uint32_t stage = 0, phase = 0, as = 0, aphase = 0, ntile = 0;
#ifdef BPB_ROLE_PROFILE  // ablation 16: the MMA does not wait for the masker (timing only)
      uint64_t* const a_ready = (XF && !(p.dbg_flags & 16)) ? xfull : full;
#else
      uint64_t* const a_ready = XF ? xfull : full;  // XF: operands are usable once masked
#endif
      // Fast path (the bf16 tower layer: 3 taps, one 64-channel chunk, one pass, resident
      // weights): every descriptor word that does not depend on the stage is built once here, so
      // the issue loop is 12 x (one add + one tcgen05.mma) per tile.  The single issuing thread
      // was measured to be the bottleneck of the pipeline when it rebuilt descriptors per MMA.
      const bool fast = RESIDENT && p.n_taps == 3 && p.KC == 1 && p.npass == 1;
      constexpr uint32_t DHI = umma_desc_hi(1024);
      uint32_t b_lo[3][4], a_row[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        a_row[j] = HALO ? static_cast<uint32_t>(p.tap_row[j]) * 8u : 0u;  // rows * 128 B >> 4
#pragma unroll
        for (int k = 0; k < 4; ++k)
          b_lo[j][k] = umma_desc_lo(smem_u32(w_smem) + j * B_TILE_BYTES + k * 32, 0);
      }
      const uint32_t ring_lo = smem_u32(ring) >> 4;
      const uint32_t stage_lo = static_cast<uint32_t>(stage_bytes) >> 4;
      TileWalk tw;
      for (tw.init(p, first_tile); tw.b < p.B; tw.next(p), ++ntile) {
        if (CG2) {  // arrivals come from both CTAs: cluster-scope acquire
          uint32_t spins = 0;
          while (!mbar_try_wait_cluster(&acc_empty[as], aphase ^ 1)) {
            if (++spins == (1u << 24)) mbar_timeout(p.dbg, 0x400 + as, aphase ^ 1, ntile);
            if (spins == (1u << 24) + (1u << 22)) __trap();
          }
        } else {
          mbar_wait_wd(&acc_empty[as], aphase ^ 1, p.dbg, 0x400 + as, ntile);
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * NT;
        // bias: D = ones[128 x 16] * bias_tile[16 x 64] opens the accumulation of the tile
        uint32_t accum0 = 0;
        if (bias_mma) {
          tc_fence_after();
          umma_bf16_lohi(d_tmem, umma_desc_lo(smem_u32(ones_tile), 0), umma_desc_hi(1024),
                         umma_desc_lo(smem_u32(bias_tile), 0), umma_desc_hi(1024), IDESC, 0u);
          accum0 = 1;
        }
        if (fast && HALO) {
          mbar_wait_wd(&a_ready[stage], phase, p.dbg, 0x500 + stage, ntile);
#ifdef BPB_ROLE_PROFILE
          if (ntile == 0) BPB_STAMP(27);
#endif
          tc_fence_after();
          const uint32_t a0 = ring_lo + stage * stage_lo;
#pragma unroll
          for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_lohi(d_tmem, (a0 + a_row[j] + 2u * k) & 0x3FFFu, DHI, b_lo[j][k], DHI, IDESC,
                             (j | k) != 0 ? 1u : accum0);
          umma_commit(&empty[stage]);
          if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1; }
        } else if (fast) {
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            mbar_wait_wd(&a_ready[stage], phase, p.dbg, 0x500 + stage, ntile);
            tc_fence_after();
            const uint32_t a0 = ring_lo + stage * stage_lo;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_lohi(d_tmem, (a0 + 2u * k) & 0x3FFFu, DHI, b_lo[j][k], DHI, IDESC,
                             (j | k) != 0 ? 1u : accum0);
            umma_commit(&empty[stage]);
            if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1; }
          }
        } else if (RESIDENT && p.strip) {
          mbar_wait_wd(&full[stage], phase, p.dbg, 0x500 + stage, ntile);
          tc_fence_after();
          const uint32_t a0 = (ring_lo + stage * stage_lo) & 0x3FFFu;
          // no-swizzle K-major: LBO = 16 B (next core matrix along K), SBO = 128 B (along M)
          constexpr uint32_t A_HI = (128u >> 4) | (1u << 14);
          constexpr uint32_t A_LBO = (16u >> 4) << 16;
          uint32_t accum = accum0;
          for (int ps = 0; ps < p.npass; ++ps) {
            const uint32_t ap = (p.pa_bits >> ps) & 1;
            const uint32_t wp = (p.pw_bits >> (2 * ps)) & 3;
            const uint32_t a_pl = a0 + ap * (static_cast<uint32_t>(p.strip_stride) >> 4);
            // one add per MMA on the issuing thread: the descriptors advance by constants (A: 32 B
            // per K step; W: 32 B inside a 64-wide chunk tile, then on to the next tile)
            uint32_t a_lo = a_pl | A_LBO;
            uint32_t b_lo = umma_desc_lo(smem_u32(w_smem) + wp * p.KC * B_TILE_BYTES, 0);
            const int nk4 = p.nk16 >> 2, nkr = p.nk16 & 3;
            for (int k4 = 0; k4 < nk4; ++k4) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_bf16_lohi(d_tmem, a_lo + 2u * k, A_HI, b_lo + 2u * k, DHI, IDESC, accum);
                accum = 1;
              }
              a_lo += 8u;
              b_lo += static_cast<uint32_t>(B_TILE_BYTES) >> 4;
            }
            for (int k = 0; k < nkr; ++k) {
              umma_bf16_lohi(d_tmem, a_lo + 2u * k, A_HI, b_lo + 2u * k, DHI, IDESC, accum);
              accum = 1;
            }
          }
          umma_commit(&empty[stage]);
          if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1; }
        } else if (HALO) {
          mbar_wait_wd(&full[stage], phase, p.dbg, 0x500 + stage, ntile);
          tc_fence_after();
          const uint32_t base = smem_u32(ring + static_cast<size_t>(stage) * stage_bytes);
          uint32_t accum = accum0;
          for (int pass = 0; pass < p.npass; ++pass)
            for (int c = 0; c < p.KC; ++c)
              for (int j = 0; j < p.n_taps; ++j) {
                const uint32_t a_addr =
                    base + (c * p.n_aplanes + ((p.pa_bits >> pass) & 1)) * p.sub_bytes +
                    p.tap_row[j] * 128;
                const uint32_t b_addr =
                    smem_u32(w_smem) +
                    ((((p.pw_bits >> (2 * pass)) & 3) * p.n_taps + j) * p.KC + c) * B_TILE_BYTES;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma_bf16(d_tmem, umma_smem_desc(a_addr + k * 32, 0, 1024),
                            umma_smem_desc(b_addr + k * 32, 0, 1024), IDESC, accum);
                  accum = 1;
                }
              }
          umma_commit(&empty[stage]);
          if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1; }
        } else {
          const int inner = RESIDENT ? p.n_aplanes : p.npass;
          const int nsteps = p.n_taps * p.KC * inner;
          uint32_t accum = accum0;
          for (int step = 0; step < nsteps; ++step) {
            const int pass = step % inner;  // RESIDENT: the A plane held by this stage
            const int c = (step / inner) % p.KC;
            const int j = step / (inner * p.KC);
            const int nk = (c == p.KC - 1) ? p.last_nk : 4;
            mbar_wait_wd(&full[stage], phase, p.dbg, 0x500 + stage, ntile);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(ring + static_cast<size_t>(stage) * SLAB_BYTES);
            if (RESIDENT) {
              for (int ps = 0; ps < p.npass; ++ps) {
                if (((p.pa_bits >> ps) & 1) != pass) continue;
                const int wp = (p.pw_bits >> (2 * ps)) & 3;
                const uint32_t b_addr =
                    smem_u32(w_smem) + ((wp * p.n_taps + j) * p.KC + c) * B_TILE_BYTES;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (k < nk) {
                    umma_bf16(d_tmem, umma_smem_desc(a_addr + k * 32, 0, 1024),
                              umma_smem_desc(b_addr + k * 32, 0, 1024), IDESC, accum);
                    accum = 1;
                  }
              }
            } else {
              const uint32_t b_addr = a_addr + kTileBytes;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (k < nk) {
                  if (CG2)
                    umma_bf16_cg2(d_tmem, umma_smem_desc(a_addr + k * 32, 0, 1024),
                                  umma_smem_desc(b_addr + k * 32, 0, 1024), IDESC, accum);
                  else
                    umma_bf16(d_tmem, umma_smem_desc(a_addr + k * 32, 0, 1024),
                              umma_smem_desc(b_addr + k * 32, 0, 1024), IDESC, accum);
                  accum = 1;
                }
            }
            if (CG2) umma_commit_cg2(&empty[stage]);  // frees the stage in both CTAs
            else umma_commit(&empty[stage]);
            if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1; }
          }
        }
        if (CG2) umma_commit_cg2(&acc_full[as]);
        else umma_commit(&acc_full[as]);
        if (++as == ACC) { as = 0; aphase ^= 1; }
      }
    }
  } else if (!RESIDENT && warp == 2 + 8 * EG) {
    // ------------------------------------------------------------- second TMA producer
    // Streamed weights (F >= 256): every K step needs an activation tile AND a weight tile, and
    // a tensor load costs ~150-280 cycles of issue on the thread that makes it -- two loads per
    // step on ONE thread took as long as the step's four MMAs (the F = 512 launches sat at 50 %
    // tensor-pipe activity however many stages the ring had).  This warp issues the weight
    // tiles; the first producer arms the stage's mbarrier with the bytes of both (a completion
    // that lands before the expect only drives the pending count negative for a moment: the
    // phase cannot complete before the first producer's arrive).  Weights are parameters: no
    // dependency wait.
    if (elect_one()) {
      uint32_t stage = 0, phase = 0, ntile = 0;
      TileWalk tw;
      for (tw.init(p, first_tile); tw.b < p.B; tw.next(p), ++ntile) {
        const int nsteps = p.n_taps * p.KC * p.npass;
        for (int step = 0; step < nsteps; ++step) {
          const int pass = step % p.npass;
          const int c = (step / p.npass) % p.KC;
          const int j = step / (p.npass * p.KC);
          mbar_wait_wd(&empty[stage], phase ^ 1, p.dbg, 0xF00 + stage, ntile);
          const int wp = (p.pw_bits >> (2 * pass)) & 3;
          if (CG2)  // this CTA's half of the N = 256 rows, signalled on the leader's barrier
            tma_load_2d_cg2(&tm.W, ring + static_cast<size_t>(stage) * SLAB_BYTES + kTileBytes,
                            &full[stage], c * 64,
                            (wp * p.n_taps + j) * p.F + tw.nt * NT + static_cast<int>(crank) * (NT / 2));
          else
            tma_load_2d(&tm.W, ring + static_cast<size_t>(stage) * SLAB_BYTES + kTileBytes,
                        &full[stage], c * 64, (wp * p.n_taps + j) * p.F + tw.nt * NT);
          if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (XF && warp >= 2 + 8 * EG) {
    // ------------------------------------------------------------- XF masker warps
    // Every staged A tile is ANDed with the ReLU bits of its rows before the MMA reads it.  A row
    // of a [rows x 64 ch] SWIZZLE_128B tile is eight 16-byte chunks of eight channels; byte c of
    // the row's mask word covers chunk c, so chunk i of the tile (row-major) has mask byte i.
    // A thread takes four consecutive chunks (half a row) per step: one 32-bit load of mask
    // bytes, four 128-bit loads, the lane masks from PRMT sign replication, four stores.  Lanes
    // 0-15 of a warp take the first halves of 16 consecutive rows, lanes 16-31 the second halves:
    // every quarter-warp then touches eight different rows at the same logical chunk, the
    // conflict-free pattern of a 128-byte-swizzled tile (two lanes per row -- one row's two
    // halves in neighbouring lanes -- measured 2-way bank conflicts on every access).  The six
    // warps of a step cover 96 rows.  (The host only selects XF on the fast path: 3 taps, one
    // chunk, one pass.)
    const int mw_ = warp - (2 + 8 * EG);    // masker warp 0..5
    const int xrow = mw_ * 16 + (lane & 15);
    const uint32_t xhalf = static_cast<uint32_t>(lane) >> 4;
    const int xrows = p.xf_rows;            // rows per staged tile
    const int nsub = HALO ? 1 : 3;          // staged tiles per output tile
    uint32_t xoff[4];                       // byte offsets of this thread's step-0 chunks
#pragma unroll
    for (int k = 0; k < 4; ++k) xoff[k] = sw128(static_cast<uint32_t>(xrow), xhalf * 4u + k);
    uint32_t stage = 0, phase = 0, ntile = 0;
    TileWalk tw;
    for (tw.init(p, first_tile); tw.b < p.B; tw.next(p), ++ntile) {
      const int t0 = tile_row0(tw.mt);
      for (int sub = 0; sub < nsub; ++sub) {
        const int row0 = t0 + (HALO ? p.box_off : p.tap_row[sub]);
        const uint32_t par =
            static_cast<uint32_t>((static_cast<long long>(tw.b) * p.L_a + row0) & 1);
        uint8_t* tile = ring + static_cast<size_t>(stage) * stage_bytes;
        const uint8_t* mb = mring + static_cast<size_t>(stage) * p.mstage + 8u * par;
        mbar_wait_wd(&full[stage], phase, p.dbg, 0xE00 + stage, ntile);
#ifdef BPB_ROLE_PROFILE
        long long tm0 = clock64();
        const int xre = (p.dbg_flags & 4) ? 0 : xrows;
#else
        const int xre = xrows;
#endif
        // rows <= 256 = 2.67 x 96: at most three steps per thread.  Step `it` handles row
        // xrow + 96 it (same swizzle phase, 96 = 0 mod 8), so its byte offsets are the step-0
        // offsets + 96 * 128 * it.  Inside a step the lane masks are
        // computed in independent waves (multiplies, PRMTs, ANDs) over the four chunks: written
        // as one chain per chunk the compiler serialised everything through two temporaries and
        // a step cost ~600 cycles of pure ALU latency.
        const uint32_t tile_a = smem_u32(tile), mb_a = smem_u32(mb) + 8u * xrow + 4u * xhalf;
#pragma unroll 1
        for (int it = 0; it < 3; ++it) {
          if (xrow + 96 * it >= xre) break;
          const uint32_t toff = tile_a + 96u * 128u * it;
          uint32_t m4;
          uint4 v[4];
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(m4) : "r"(mb_a + 8u * 96u * it));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(v[k].x), "=r"(v[k].y), "=r"(v[k].z), "=r"(v[k].w)
                         : "r"(toff + xoff[k]));
          uint32_t ea[4], eb[4], w[4][4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // bit i of a nibble -> most significant bit of byte i (no two partial products
            // meet, so no carries); PRMT then replicates those sign bits over the 16-bit lanes
            ea[k] = ((m4 >> (8 * k)) & 0xFu) * 0x10204080u;
            eb[k] = ((m4 >> (8 * k + 4)) & 0xFu) * 0x10204080u;
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            w[k][0] = prmt(ea[k], 0u, 0x9988u);
            w[k][1] = prmt(ea[k], 0u, 0xBBAAu);
            w[k][2] = prmt(eb[k], 0u, 0x9988u);
            w[k][3] = prmt(eb[k], 0u, 0xBBAAu);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            v[k].x &= w[k][0];
            v[k].y &= w[k][1];
            v[k].z &= w[k][2];
            v[k].w &= w[k][3];
          }
#pragma unroll
          for (int k = 0; k < 4; ++k)
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(toff + xoff[k]),
                         "r"(v[k].x), "r"(v[k].y), "r"(v[k].z), "r"(v[k].w)
                         : "memory");
        }
        // generic-proxy writes -> visible to the tensor core (async proxy), then hand over
#ifdef BPB_ROLE_PROFILE
        {
          const long long tnow = clock64();
          if (lane == 0) atomicAdd(prof_smem() + 7, static_cast<unsigned>(tnow - tm0));
          tm0 = tnow;
        }
#endif
#ifdef BPB_ROLE_PROFILE
        if (!(p.dbg_flags & 32))  // ablation 32: no proxy fence (timing only)
#endif
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&xfull[stage]);
#ifdef BPB_ROLE_PROFILE
        if (lane == 0) atomicAdd(prof_smem() + 0, static_cast<unsigned>(clock64() - tm0));
#endif
        if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------- epilogue warps (2..)
    // EG groups of 8 warps; group g owns this CTA's tiles g, g + EG, ... (its own staging buffer
    // and named barrier), so one group's stores / barrier waits overlap the other's math.
    // The bias is a parameter: written only by the optimiser, never by an in-flight predecessor,
    // so it is fetched before the dependency wait -- and by the warps that use it, off the
    // producer's and the MMA issuer's start-up path.
    for (int i = static_cast<int>(threadIdx.x) - 64; i < p.F; i += EG * kEpiThreads)
      s_bias[i] = (p.bias && i < p.bias_n) ? p.bias[i] : 0.f;
    if (bias_mma) {
      const int te = static_cast<int>(threadIdx.x) - 64;
      uint4 one4;
      one4.x = one4.y = one4.z = one4.w = 0x3F803F80u;
      for (int i = te; i < 128 * 128 / 16; i += EG * kEpiThreads)
        reinterpret_cast<uint4*>(ones_tile)[i] = one4;
      if (te < 64) {
        // row n of the K-major bias tile: k = 0..2 hold bias[n] as successive bf16 remainders
        // (exact to fp32), every other k is zero; 16-byte chunks XOR-ed with the row (SW128)
        const float bv = (p.bias && te < p.bias_n) ? p.bias[te] : 0.f;
        const __nv_bfloat16 h0 = __float2bfloat16_rn(bv);
        const float r1 = bv - __bfloat162float(h0);
        const __nv_bfloat16 h1 = __float2bfloat16_rn(r1);
        const __nv_bfloat16 h2 = __float2bfloat16_rn(r1 - __bfloat162float(h1));
        uint4 c0;
        c0.x = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) |
               (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
        c0.y = static_cast<uint32_t>(__bfloat16_as_ushort(h2));
        c0.z = c0.w = 0u;
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(bias_tile + te * 128 + ((c ^ (te & 7)) << 4)) = (c == 0) ? c0 : z4;
      }
      fence_proxy_async();
    }
    named_bar_sync(7, EG * kEpiThreads);
    if (bias_mma && threadIdx.x == 64) mbar_arrive(b_full);
    pdl_wait();  // global reads (masks, multipliers) and TMA stores below
    const int grp = (warp - 2) >> 3;
    const int q = warp & 3;
    const int half = ((warp - 2) >> 2) & 1;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64 - grp * kEpiThreads;
    const int bar_id = 1 + grp;
    const int chunks_per_row = p.F >> 6;
    const uint32_t RG = static_cast<uint32_t>(p.rstages) / EG;  // this group's residual sub-ring
    const uint32_t rbase = static_cast<uint32_t>(grp) * RG;
    uint32_t as = 0, aphase = 0, rs = 0, rphase = 0, hs = 0, hphase = 0, kc = 0;
    TileWalk tw;
    tw.init(p, first_tile);
    for (int skip = 0; skip < grp && tw.b < p.B; ++skip) {  // this group's first tile
      tw.next(p);
      ring_advance(as, aphase, 1, ACC);
      ring_advance(hs, hphase, 1, p.stages);
    }
    uint32_t ntile = grp;
    for (; tw.b < p.B; tw.next(p), ntile += EG) {
      const int nt = tw.nt, b = tw.b;
      const int t0 = tile_row0(tw.mt);
      const int bz = (EPI == EPI_FWD_R) ? b / p.zx_div : 0;
      const int t = t0 + row;
      const bool valid = t < p.L_out;
      const size_t orow = static_cast<size_t>(b) * p.L_out + t;

#pragma unroll 1
      for (int ch = 0; ch < NCH; ++ch, ++kc) {
        const int n0 = nt * NT + ch * 64;
        const int cq = half * 4;  // first 16-byte chunk of this thread's 32 columns
        // (Prefetching these bits one chunk visit ahead was measured: -15 % on an L2-hot d = 2
        // launch, +2 % on the real step -- the extra tile-walk arithmetic costs more than the
        // load it hides once the masks come out of L2 anyway.)
        uint32_t mbits = ~0u;
        if (EPI == EPI_BWD_MASK && !XF && p.mask_in != nullptr && valid)
          mbits = __ldg(p.mask_in + (orow * chunks_per_row + (n0 >> 6)) * 2 + half);

        // ---- residual (already in shared memory: HALO box or residual ring)
        // The slot is handed back to the TMA producer only AFTER the values read from it have
        // been consumed (``release`` below, after the staging writes).  Arriving right behind
        // the ld.shared instructions is not enough: they are asynchronous, the arrive does not
        // depend on their destination registers, and a refill that lands while some of them are
        // still outstanding was observed -- about once per hundred launches of C1's d = 64
        // layer a warp added the residual of the tile four steps ahead
        // (tests/test_gpu_baseline_cfgs.py::test_every_c1_tower_shape...).  In-order issue puts
        // every instruction behind the stores that use the loaded registers behind the loads.
        uint4 rr[4], rl[4];
        uint64_t* release = nullptr;
        if (RESID) {
          if (DIRECT_RESID) {
            const int tr = t + p.resid_off;
            if (tr >= 0 && tr < p.L_res) {  // rows outside the tensor count as zero
              const size_t roff = ((static_cast<size_t>(b) * p.L_res + tr) * p.F + n0 + half * 32) * 2;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                rr[g] = __ldg(reinterpret_cast<const uint4*>(p.resid_ptr[0] + roff) + g);
                if (PLANES == 2) rl[g] = __ldg(reinterpret_cast<const uint4*>(p.resid_ptr[1] + roff) + g);
              }
            } else {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                rr[g] = make_uint4(0u, 0u, 0u, 0u);
                if (PLANES == 2) rl[g] = make_uint4(0u, 0u, 0u, 0u);
              }
            }
          } else if (RING_RESID) {
#ifdef BPB_ROLE_PROFILE
            if (!(p.dbg_flags & 2))
#endif
            mbar_wait_wd(&rfull[rbase + rs], rphase, p.dbg, 0x600 + rbase + rs, ntile);
            const uint8_t* src = rring + static_cast<size_t>(rbase + rs) * rstage_bytes;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              rr[g] = lds128(src + sw128(row, cq + g));
              if (PLANES == 2) rl[g] = lds128(src + kTileBytes + sw128(row, cq + g));
            }
            release = &rempty[rbase + rs];
#ifdef BPB_ROLE_PROFILE
            if (p.dbg_flags & 2) release = nullptr;
#endif
            ring_advance(rs, rphase, 1, RG);
          } else {
            if (ch == 0) mbar_wait_wd(&full[hs], hphase, p.dbg, 0x700 + hs, ntile);
            // n_ntiles == 1 in HALO mode, so the chunk index is the channel-chunk index
            const uint8_t* src = ring + static_cast<size_t>(hs) * stage_bytes +
                                 static_cast<size_t>(ch * p.n_aplanes) * p.sub_bytes;
            const uint32_t brow = row + p.resid_row;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              rr[g] = lds128(src + sw128(brow, cq + g));
              if (PLANES == 2) rl[g] = lds128(src + p.sub_bytes + sw128(brow, cq + g));
            }
            if (ch == NCH - 1) {
              release = &empty[hs];
              ring_advance(hs, hphase, EG, p.stages);
            }
          }
        }

#ifdef BPB_ROLE_PROFILE
        long long tp0 = clock64();
#define BPB_PHASE(slot)                                                            \
  do {                                                                              \
    const long long tnow = clock64();                                               \
    if (lane == 0) atomicAdd(prof_smem() + (slot), static_cast<unsigned>(tnow - tp0)); \
    tp0 = tnow;                                                                     \
  } while (0)
#else
#define BPB_PHASE(slot)
#endif
        // ---- accumulator
        if (ch == 0) {
          mbar_wait_wd(&acc_full[as], aphase, p.dbg, 0x800 + as, ntile);
          tc_fence_after();
#ifdef BPB_ROLE_PROFILE
          if (et == 0 && ntile == 0) BPB_STAMP(22);
#endif
        }
        uint32_t acc[32];
        tmem_ld32(tmem_base + as * NT + ch * 64 + half * 32 + (static_cast<uint32_t>(q * 32) << 16),
                  acc);
        tmem_ld_wait();
        if (ch == NCH - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG2) mbar_arrive_leader(&acc_empty[as]);
            else mbar_arrive(&acc_empty[as]);
          }
        }
        BPB_PHASE(9);   // acc wait + TMEM load

        // Every warp owns the [32 rows x 32 channels] corner of the tile it computes: its own
        // 2 KB staging tiles (64-byte swizzle) and its own TMA stores.  No CTA-wide barrier in
        // the epilogue: `barrier` was the top stall of the HALO-staged launches when one thread
        // stored the whole [128 x 64] tile for eight warps (two named barriers per tile).
        const int wepi = warp - 2;  // 0 .. 8 EG - 1
        uint8_t* sb = stg + static_cast<size_t>(wepi) * (n_out * PLANES * 2048);
        uint8_t* sb2 = sb + static_cast<size_t>(p.has_out * PLANES) * 2048;
        if (lane == 0) tma_store_wait_read<0>();  // this warp's previous stores have read sb
        __syncwarp();

        BPB_PHASE(10);  // staging buffer free (wait_read + barrier)
        float obits_lo = 0.f, obits_hi = 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float v[8], r[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(acc[g * 8 + i]);
          if (RESID) {
            unpack8(rr[g], r);
            if (PLANES == 2) {
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
          if (IS_FWD) {
            if (!bias_mma) {
              const float4 b0 = lds_f4(s_bias + n0 + half * 32 + g * 8);
              const float4 b1 = lds_f4(s_bias + n0 + half * 32 + g * 8 + 4);
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
              v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            }
            if (EPI == EPI_FWD_R) {
              // DeepSHAP rescale multiplier of this reference against the query's pre-activation
              // (shap.py:623-638); only the query-half gradient is used downstream (shap.py:374).
              float zx[8];
              if (valid) {
                const float4* zp = reinterpret_cast<const float4*>(
                    p.zx_in + (static_cast<size_t>(bz) * p.L_out + t) * p.F + n0 +
                    half * 32 + g * 8);
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
                // approximate division (2 ulp): the multiplier is stored in bf16 planes, and a
                // full-precision divide was a third of this epilogue's instructions
                o2[i] = (fabsf(din) < 1e-6f) ? (zx[i] > 0.f ? 1.f : 0.f) : __fdividef(dout, din);
              }
            }
            if (EPI == EPI_FWD_Z) {
              if (valid) {
                float4* zo = reinterpret_cast<float4*>(p.z_out + orow * p.F + n0 + half * 32 + g * 8);
                zo[0] = make_float4(v[0], v[1], v[2], v[3]);
                zo[1] = make_float4(v[4], v[5], v[6], v[7]);
              }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (EPI != EPI_FWD_R) {
                const float flag = gt0_flag(v[i]);
                if (g < 2) obits_lo = fmaf(flag, static_cast<float>(1u << (g * 8 + i)), obits_lo);
                else obits_hi = fmaf(flag, static_cast<float>(1u << (g * 8 + i - 16)), obits_hi);
                o[i] = fmaf(v[i], flag, r[i]);
              } else {
                // the reference pass of DeepSHAP keeps multipliers, not ReLU masks (the ABI rejects
                // mask_out together with out2): no mask bits, and max(v, 0) is shared with the
                // rescale rule above
                o[i] = fmaxf(v[i], 0.f) + r[i];
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = v[i] + r[i];
            if (EPI == EPI_BWD_MULT) {
              float m[8];
              if (valid) {
                const size_t moff = orow * p.F + n0 + half * 32 + g * 8;
                unpack8(__ldg(reinterpret_cast<const uint4*>(p.mult_hi + moff)), m);
                if (PLANES == 2) {
                  float m2[8];
                  unpack8(__ldg(reinterpret_cast<const uint4*>(p.mult_lo + moff)), m2);
#pragma unroll
                  for (int i = 0; i < 8; ++i) m[i] += m2[i];
                }
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) m[i] = 0.f;
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) o2[i] = o[i] * m[i];
            } else if (!XF) {
#pragma unroll
              for (int i = 0; i < 8; ++i) o2[i] = ((mbits >> (g * 8 + i)) & 1u) ? o[i] : 0.f;
            }
          }
          const uint32_t off = sw64(static_cast<uint32_t>(lane), static_cast<uint32_t>(g));
          if (p.has_out) {
            if (PLANES == 2) {
              uint4 h, l;
              split8(o, h, l);
              sts128(sb + off, h);
              sts128(sb + 2048 + off, l);
            } else {
              sts128(sb + off, pack8(o));
            }
          }
          if ((EPI == EPI_FWD_R || !IS_FWD) && !XF && p.has_out2) {
            if (PLANES == 2) {
              uint4 h, l;
              split8(o2, h, l);
              sts128(sb2 + off, h);
              sts128(sb2 + 2048 + off, l);
            } else {
              sts128(sb2 + off, pack8(o2));
            }
          }
        }

        BPB_PHASE(11);  // math + staging writes
        if (RESID && release != nullptr) {
          // every lane's residual values went through the stores above: the slot may be refilled
          __syncwarp();
          if (lane == 0) mbar_arrive(release);
        }
        // ---- publish: generic-proxy writes -> async proxy, then this warp's own stores
        fence_proxy_async();
        BPB_PHASE(12);  // proxy fence
        __syncwarp();
#ifdef BPB_ROLE_PROFILE
        if (lane == 0 && !(p.dbg_flags & 1)) {
#else
        if (lane == 0) {
#endif
          const int c0 = n0 + half * 32, r0 = t0 + q * 32;
          if (p.has_out) {
            tma_store_3d(&tm.O[0], sb, c0, r0, b);
            if (PLANES == 2) tma_store_3d(&tm.O[1], sb + 2048, c0, r0, b);
#ifdef BPB_ROLE_PROFILE
            if (p.dbg_flags & 8) tma_store_3d(&tm.O[0], sb, c0, r0, b);  // ablation: store twice
#endif
          }
          if (!XF && p.has_out2) {
            tma_store_3d(&tm.P[0], sb2, c0, r0, b);
            if (PLANES == 2) tma_store_3d(&tm.P[1], sb2 + 2048, c0, r0, b);
          }
          tma_store_commit();
        }
        BPB_PHASE(13);  // barrier + TMA store issue
        // The mask goes out AFTER the proxy fence above: fence.proxy.async compiles to a
        // MEMBAR that drains this warp's outstanding stores, and a global store issued before
        // it would put a round trip to L2 on the critical path of every chunk.
        if (EPI != EPI_FWD_R && p.mask_out != nullptr && valid)
          p.mask_out[(orow * chunks_per_row + (n0 >> 6)) * 2 + half] =
              static_cast<uint32_t>(obits_lo) | (static_cast<uint32_t>(obits_hi) << 16);
      }
      ring_advance(as, aphase, EG, ACC);
      // skip the other groups' tiles
      for (int skip = 1; skip < EG && tw.b < p.B; ++skip) tw.next(p);
    }
#ifdef BPB_ROLE_PROFILE
    if (et == 0 && grp == 0) BPB_STAMP(23);
#endif
    if (lane == 0) tma_store_wait_all<0>();
#ifdef BPB_ROLE_PROFILE
    if (et == 0 && grp == 0) BPB_STAMP(24);
#endif
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (CG2) {
    // neither CTA may free its TMEM, or exit with barriers the other still signals, before both
    // are through
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_cg2<TMEM_COLS>(tmem_base);
  } else if (warp == 1) {
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
#ifdef BPB_ROLE_PROFILE
  if (threadIdx.x < 32 && p.dbg != nullptr && blockIdx.x < 8) {
    unsigned v = prof_smem()[threadIdx.x];
    if (threadIdx.x == 15) v = static_cast<unsigned>(clock64() - t_kernel_start);
    atomicAdd(p.dbg + 1024 + 32 * blockIdx.x + threadIdx.x, v);
  }
#endif
}

// ------------------------------------------------------------------------------------------------
// Host side: description of one launch, shared-memory plan, template dispatch.
// ------------------------------------------------------------------------------------------------
struct ConvDesc {
  const char* who;  // ABI entry point, for error messages
  int B, L_out, F;  // sequences, output rows per sequence, output channels (multiple of 64)
  int epi;          // EPI_*
  int planes;       // planes (1 | 2) of out / out2 / resid
  // A operand: bf16 tensor(s) [B][a_rows][a_inner] with arbitrary (16-byte multiple) strides
  const void* a_ptr[2];
  int n_aplanes;
  unsigned long long a_inner, a_rows, a_row_stride_bytes, a_batch_stride_bytes;
  // K structure
  int n_taps, KC, last_nk;
  int tap_off[3];
  // W operand: bf16 [n_wplanes][n_taps][F][KC * 64], contiguous
  const void* w_ptr;
  int n_wplanes;
  int npass, pa_bits, pw_bits;
  // epilogue
  const float* bias;
  int bias_n;
  const void* resid[2];
  int L_res, resid_off;
  void* out[2];
  void* out2[2];
  const void* mask_in;
  void* mask_out;
  const void* mask_a;  // XF: 1-bit mask of the A operand's rows [B][a_rows] (F = 64), or NULL
  const void* mult[2];
  const float* zx_in;
  int zx_div;
  float* z_out;
  bool allow_halo;
  bool allow_strip;  // window convolution over a 16-byte-per-position tensor (n_taps == 1)
};

// BPB_STRIP=0 disables STRIP staging (falls back to overlapping-row TMA boxes).
static const bool g_strip_enabled = [] {
  const char* e = getenv("BPB_STRIP");
  return !(e && e[0] == '0');
}();

// Tuning knob (BPB_EPI_GROUPS=1|2|3, default 2): most 8-warp epilogue groups used on the bf16 path.
// Three groups (-DBPB_ENABLE_EG3, F = 64 with resident weights) were measured SLOWER than two on
// C1 (forward 123 vs 105 us per step, data gradient 156 vs 136): 72 instead of 96 registers per
// thread, three staging buffers leave a shorter operand ring, and the math phases of the groups
// already compete for issue slots.  The kernel stays general; the instantiation is off.
static int g_epi_groups = [] {
  const char* e = getenv("BPB_EPI_GROUPS");
#ifdef BPB_ENABLE_EG3
  if (e && e[0] == '3') return 3;
#endif
  return (e && e[0] == '1') ? 1 : 2;
}();

// Tuning knob (BPB_BIAS_MMA=0|1, default 1): forward bias through the tensor core.
static bool g_bias_mma = [] {
  const char* e = getenv("BPB_BIAS_MMA");
  return !(e && e[0] == '0');
}();

template <int EPI, int NT, bool RESIDENT, int PLANES, bool HALO, int EG, bool RESID, bool XF = false,
          bool CG2 = false>
static int launch_conv_eg(const ConvMaps& maps, const ConvKParams& kp, size_t smem_bytes, int grid,
                          cudaStream_t stream) {
  auto kern = conv_tc_kernel<EPI, NT, RESIDENT, PLANES, HALO, EG, RESID, XF, CG2>;
  static bool configured = false;
  if (!configured) {
    BP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       max_smem_optin()));
    configured = true;
  }
  const dim3 block(64 + kEpiThreads * EG + (XF ? kXfThreads : 0) + (RESIDENT ? 0 : 32));
  if (CG2) BP_CHECK_CUDA(launch_pdl_pairs(kern, dim3(grid), block, smem_bytes, stream, maps, kp));
  else BP_CHECK_CUDA(launch_pdl(kern, dim3(grid), block, smem_bytes, stream, maps, kp));
  return 0;
}

// Two epilogue groups whenever the plan has two staging buffers and group-stable rings.
template <int EPI, int NT, bool RESIDENT, int PLANES, bool HALO, bool RESID>
static int launch_conv(const ConvMaps& maps, const ConvKParams& kp, size_t smem_bytes, int grid,
                       cudaStream_t stream) {
  if constexpr ((EPI == EPI_FWD || EPI == EPI_BWD_MASK) && NT == 256 && !RESIDENT && PLANES == 1 &&
                !HALO && RESID)
    if (kp.cg2)
      return launch_conv_eg<EPI, NT, false, 1, false, 2, RESID, false, true>(maps, kp, smem_bytes,
                                                                             grid, stream);
  if constexpr (EPI == EPI_BWD_MASK && NT == 64 && RESIDENT && PLANES == 1 && RESID)
    if (kp.mask_a != nullptr)  // conv_run planned two staging buffers / group-stable rings for it
      return launch_conv_eg<EPI, NT, RESIDENT, 1, HALO, 2, RESID, true>(maps, kp, smem_bytes, grid,
                                                                        stream);
#ifdef BPB_ENABLE_EG3
  if constexpr (NT == 64 && RESIDENT)
    if (PLANES == 1 && kp.nstg == 3 && kp.epi_groups >= 3 &&
        (kp.rstages % 3 == 0 || !kp.ring_resid))
      return launch_conv_eg<EPI, NT, RESIDENT, 1, HALO, 3, RESID>(maps, kp, smem_bytes, grid, stream);
#endif
  if (PLANES == 1 && kp.nstg == 2 && kp.epi_groups >= 2 && (kp.rstages % 2 == 0 || !kp.ring_resid))
    return launch_conv_eg<EPI, NT, RESIDENT, 1, HALO, 2, RESID>(maps, kp, smem_bytes, grid, stream);
  return launch_conv_eg<EPI, NT, RESIDENT, PLANES, HALO, 1, RESID>(maps, kp, smem_bytes, grid, stream);
}

template <int EPI, int PLANES, bool RESID>
static int dispatch_shape(int NT, bool resident, bool halo, const ConvMaps& maps,
                          const ConvKParams& kp, size_t smem, int grid, cudaStream_t s) {
  if (NT == 64) {
    if constexpr (RESID)
      if (halo) return launch_conv<EPI, 64, true, PLANES, true, RESID>(maps, kp, smem, grid, s);
    if (resident) return launch_conv<EPI, 64, true, PLANES, false, RESID>(maps, kp, smem, grid, s);
    return launch_conv<EPI, 64, false, PLANES, false, RESID>(maps, kp, smem, grid, s);
  }
  if (NT == 128) {
    if constexpr (RESID)
      if (halo) return launch_conv<EPI, 128, true, PLANES, true, RESID>(maps, kp, smem, grid, s);
    if (resident) return launch_conv<EPI, 128, true, PLANES, false, RESID>(maps, kp, smem, grid, s);
    return launch_conv<EPI, 128, false, PLANES, false, RESID>(maps, kp, smem, grid, s);
  }
  return launch_conv<EPI, 256, false, PLANES, false, RESID>(maps, kp, smem, grid, s);
}

template <int EPI, bool RESID>
static int dispatch_planes(int planes, int NT, bool resident, bool halo, const ConvMaps& maps,
                           const ConvKParams& kp, size_t smem, int grid, cudaStream_t s) {
  if (planes == 2) return dispatch_shape<EPI, 2, RESID>(NT, resident, halo, maps, kp, smem, grid, s);
  return dispatch_shape<EPI, 1, RESID>(NT, resident, halo, maps, kp, smem, grid, s);
}

// CTA pairs that can be resident at once (a pair must sit inside one GPC; with an odd number of
// usable SMs in a GPC one SM stays empty): asked of the runtime once, for the forward pair kernel at
// the largest dynamic shared-memory size; sm_count() / 2 if the query fails.
template <bool RESID>
static int cg2_max_pairs() {
  static const int pairs = [] {
    int n = sm_count() / 2;
    if constexpr (RESID) {
      auto kern = conv_tc_kernel<EPI_FWD, 256, false, 1, false, 2, true, false, true>;
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               max_smem_optin()) == cudaSuccess) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * n);
        cfg.blockDim = dim3(64 + kEpiThreads * 2 + 32);
        cfg.dynamicSmemBytes = max_smem_optin();
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int q = 0;
        if (cudaOccupancyMaxActiveClusters(&q, kern, &cfg) == cudaSuccess && q > 0 && q < n) n = q;
      }
      (void)cudaGetLastError();
    }
    return n > 0 ? n : 1;
  }();
  return pairs;
}

// RESID selects the family a translation unit instantiates: true = residual tower layers (HALO
// allowed), false = plain position-major GEMMs (input convolution, head data gradient).
template <bool RESID>
static int conv_run(const ConvDesc& d, cudaStream_t stream) {
  const int F = d.F;
  const int NT = (F % 256 == 0) ? 256 : (F % 128 == 0 ? 128 : 64);
  const int n_planes = d.planes;
  const int has_out = d.out[0] ? 1 : 0, has_out2 = d.out2[0] ? 1 : 0;
  const bool is_fwd = d.epi <= EPI_FWD_R;
  ConvKParams kp{};
  kp.B = d.B;
  kp.L_out = d.L_out;
  kp.tiles_per_seq = (d.L_out + kTileM - 1) / kTileM;
  kp.n_ntiles = F / NT;
  kp.F = F;
  kp.n_taps = d.n_taps;
  kp.KC = d.KC;
  kp.last_nk = d.last_nk;
  kp.npass = d.npass;
  kp.pa_bits = d.pa_bits;
  kp.pw_bits = d.pw_bits;
  kp.n_aplanes = d.n_aplanes;
  kp.n_wplanes = d.n_wplanes;
  kp.resid_off = d.resid_off;
  kp.zx_div = d.zx_div > 0 ? d.zx_div : 1;
  kp.has_out = has_out;
  kp.has_out2 = has_out2;
  kp.epi_groups = g_epi_groups;
  kp.bias = d.bias;
  kp.bias_n = d.bias_n;
  kp.mult_hi = reinterpret_cast<const __nv_bfloat16*>(d.mult[0]);
  kp.mult_lo = reinterpret_cast<const __nv_bfloat16*>(d.mult[1]);
  kp.mask_in = reinterpret_cast<const uint32_t*>(d.mask_in);
  kp.mask_out = reinterpret_cast<uint32_t*>(d.mask_out);
  kp.zx_in = d.zx_in;
  kp.z_out = d.z_out;
  kp.dbg = debug_buffer();
  {
    const char* e = getenv("BPB_DBG_FLAGS");
    kp.dbg_flags = e ? atoi(e) : 0;
  }
  const bool xf = d.mask_a != nullptr;
  if (xf) {
    BP_REQUIRE(RESID && d.epi == EPI_BWD_MASK && n_planes == 1 && F == 64 && d.KC == 1 &&
                   d.n_taps == 3 && d.npass == 1 && d.last_nk == 4 && g_epi_groups >= 2,
               "%s: mask_a (operand masking in shared memory) needs the bf16 backward of an "
               "F = 64 layer", d.who);
    BP_REQUIRE(has_out && !has_out2 && d.mask_in == nullptr,
               "%s: with mask_a the only output is out (the consumers mask it themselves)", d.who);
    BP_REQUIRE(((static_cast<unsigned long long>(d.B) * d.a_rows) & 1ull) == 0,
               "%s: mask_a needs an even B * L_in (16-byte aligned bulk copies of mask words)",
               d.who);
    kp.mask_a = static_cast<const uint8_t*>(d.mask_a);
    kp.L_a = static_cast<int>(d.a_rows);
    kp.mask_rows_total = static_cast<long long>(d.B) * static_cast<long long>(d.a_rows);
  }

  // ---- shared-memory plan
  const int max_smem = max_smem_optin();
  BP_REQUIRE(max_smem > 0, "%s: no CUDA device", d.who);
  int min_off = d.tap_off[0], max_off = d.tap_off[0];
  for (int j = 1; j < d.n_taps; ++j) {
    if (d.tap_off[j] < min_off) min_off = d.tap_off[j];
    if (d.tap_off[j] > max_off) max_off = d.tap_off[j];
  }
  const int span = max_off - min_off;
  const int n_out = has_out + has_out2;
  // CTA pairs (cta_group::2) for the streamed-weight bf16 training launches: BPB_CG2=0 turns it off
  const bool g_cg2 = [] {  // read per call: the parity tests run both shapes in one process
    const char* e = getenv("BPB_CG2");
    return !(e && e[0] == '0');
  }();
  bool cg2 = RESID && NT == 256 && n_planes == 1 && g_cg2 && g_epi_groups >= 2 &&
             (d.epi == EPI_FWD || d.epi == EPI_BWD_MASK) && !xf && sm_count() >= 2;
  size_t b_tile = static_cast<size_t>(cg2 ? NT / 2 : NT) * 128;
  const size_t w_res = static_cast<size_t>(d.n_wplanes) * d.n_taps * d.KC * b_tile;
  const size_t fixed = 1024 /*align slack*/ + static_cast<size_t>(F) * 4 +
                       (3 * kMaxStages + 2 * kMaxAcc + 2 * kMaxRStages + 2) * 8 + 48;
  const bool can_reside = (kp.n_ntiles == 1) && (NT <= 128);
  // With two epilogue groups (bf16 path, two staging buffers) the residual ring must have an EVEN
  // number of slots: a slot is then always consumed by the same group.  With an odd ring a group
  // would skip every other phase of a slot's mbarrier, and a parity wait cannot tell "the phase I
  // skipped is still in flight" from "my phase is complete" (observed as a rare deadlock).
  const bool two_groups = (n_planes == 1) && g_epi_groups >= 2;
  const bool three_groups = (n_planes == 1) && g_epi_groups >= 3 && NT == 64;
  int r_hi = two_groups ? 4 : 3;
  if (const char* e = getenv("BPB_RSTAGES")) {  // tuning knob: residual ring slots (even: 2..4)
    const int v = atoi(e);
    if (v >= 2 && v <= kMaxRStages && (!two_groups || v % 2 == 0)) r_hi = v;
  }

  // STRIP staging: resident weights, one tap, rows one 16-byte position apart
  const int nk16 = (d.KC - 1) * 4 + d.last_nk;
  const int strip_bytes = (127 + 2 * nk16) * 16;
  // multiple of 1024 so that everything carved after the ring keeps the swizzle-atom alignment
  const int strip_stride = (strip_bytes + 1023) / 1024 * 1024;
  const bool strip_ok = !RESID && d.allow_strip && can_reside && d.n_taps == 1 &&
                        d.a_row_stride_bytes == 16 && g_strip_enabled;
  bool halo = false, resident = false;
  int stages = 0, rstages = 0, nstg = 0;
  size_t stage_bytes = 0, mstage = 0;
  int box_rows = 0;
  auto try_plan = [&](bool want_halo, bool want_res, int want_nstg, int want_rst, int min_stages) {
#if defined(BPB_DIRECT_ALL)
    const bool ring_resid = false;
#else
    // (forward: from the HALO box or by direct loads; streamed weights: direct loads)
    const bool ring_resid = RESID && !is_fwd && want_res;
#endif
    // (not for STRIP staging: that kernel is bound by its MMA issue, the extra MMA costs more
    // than the adds it saves -- measured)
    const bool want_bias_mma = want_res && is_fwd && NT == 64 && d.bias != nullptr && g_bias_mma &&
                               !(strip_ok && !want_halo);
    size_t used = fixed + (want_res ? w_res : 0) + (want_bias_mma ? kBiasMmaBytes : 0) +
                  static_cast<size_t>(want_nstg) * n_out * n_planes * kTileBytes +
                  (ring_resid ? static_cast<size_t>(want_rst) * n_planes * kTileBytes : 0);
    if (used >= static_cast<size_t>(max_smem)) return false;
    size_t sb;
    int rows = 0;
    if (want_halo) {
      rows = kTileM + span;
      sb = static_cast<size_t>(d.KC) * d.n_aplanes *
           ((static_cast<size_t>(rows) * 128 + 1023) / 1024 * 1024);
    } else if (want_res && strip_ok) {
      sb = static_cast<size_t>(d.n_aplanes) * strip_stride;
    } else {
      sb = kTileBytes + (want_res ? 0 : b_tile);
    }
    // XF: every stage also holds the mask words of its rows (+ 2: 16-byte alignment of the copy)
    const size_t ms = xf ? static_cast<size_t>((want_halo ? rows : kTileM) + 2) * 8 : 0;
    int st = static_cast<int>((static_cast<size_t>(max_smem) - used) / (sb + ms));
    if (st > kMaxStages) st = kMaxStages;
    if (want_halo && st > 6) st = 6;
    // same-group ownership of every HALO box slot as well
    if (want_halo && two_groups && is_fwd && want_nstg == 2 && st > 2 && (st & 1)) st -= 1;
    if (want_halo && is_fwd && want_nstg == 3) st -= st % 3;
    if (st < min_stages) return false;
    halo = want_halo; resident = want_res; nstg = want_nstg; rstages = ring_resid ? want_rst : 0;
    stages = st; stage_bytes = sb; box_rows = rows; mstage = ms;
    return true;
  };
  const bool halo_ok = RESID && d.allow_halo && can_reside && span <= 128 && span > 0 &&
                       d.last_nk == 4;
  bool planned = false;
  if (cg2) {  // needs the two-group plan (two staging buffers, even residual ring)
    planned = try_plan(false, false, 2, r_hi, 3) || try_plan(false, false, 2, 2, 2);
    if (!planned) {
      cg2 = false;
      b_tile = static_cast<size_t>(NT) * 128;
    }
  }
  // Three epilogue groups (3 staging buffers, residual sub-rings of one slot) when shared memory
  // still leaves a useful operand ring; the weights must be resident (F = 64 towers).
  if (three_groups && halo_ok) planned = try_plan(true, true, 3, 3, is_fwd ? 3 : 2);
  if (!planned && three_groups && can_reside && !strip_ok) planned = try_plan(false, true, 3, 3, 6);
  if (!planned && three_groups && can_reside && strip_ok) planned = try_plan(false, true, 3, 3, 3);
  if (!planned && halo_ok) planned = try_plan(true, true, 2, r_hi, 3) || try_plan(true, true, 2, 2, 2);
  if (!planned && can_reside)
    planned = try_plan(false, true, 2, r_hi, 4) || try_plan(false, true, 2, 2, 3) ||
              try_plan(false, true, 1, 2, 3);
  if (!planned)
    planned = try_plan(false, false, 2, r_hi, 3) || try_plan(false, false, 2, 2, 2) ||
              try_plan(false, false, 1, 2, 2);
  BP_REQUIRE(planned, "%s: not enough shared memory for F=%d planes=%d", d.who, F, n_planes);
  kp.stages = stages;
  kp.bias_mma = (resident && is_fwd && NT == 64 && d.bias != nullptr && g_bias_mma &&
                 !(strip_ok && !halo)) ? 1 : 0;
  if (resident && !halo && strip_ok) {
    kp.strip = 1;
    kp.strip_bytes = strip_bytes;
    kp.strip_stride = strip_stride;
    kp.nk16 = nk16;
    kp.a_batch_bytes = static_cast<long long>(d.a_batch_stride_bytes);
    kp.a_base[0] = static_cast<const uint8_t*>(d.a_ptr[0]);
    kp.a_base[1] = static_cast<const uint8_t*>(d.a_ptr[1]);
  }
  kp.resid_ptr[0] = static_cast<const uint8_t*>(d.resid[0]);
  kp.resid_ptr[1] = static_cast<const uint8_t*>(d.resid[1]);
  kp.L_res = d.L_res;
  kp.rstages = rstages > 0 ? rstages : 1;
  kp.ring_resid = rstages > 0 ? 1 : 0;
  kp.nstg = nstg;
  if (xf) {
    BP_REQUIRE(resident && nstg == 2 && (rstages % 2 == 0), "%s: mask_a: unexpected plan", d.who);
    kp.xf_rows = halo ? box_rows : kTileM;
    kp.mstage = static_cast<int>(mstage);
  }
  if (halo) {
    kp.box_off = min_off;
    kp.box_bytes = box_rows * 128;
    kp.sub_bytes = (kp.box_bytes + 1023) / 1024 * 1024;
    for (int j = 0; j < d.n_taps; ++j) kp.tap_row[j] = d.tap_off[j] - min_off;
    kp.resid_row = d.resid_off - min_off;
    if (is_fwd) {
      BP_REQUIRE(d.resid[0] == d.a_ptr[0] && d.resid[1] == d.a_ptr[1] && kp.resid_row >= 0 &&
                     kp.resid_row <= span,
                 "%s: forward residual must be the input tensor at an offset inside the taps",
                 d.who);
    }
  } else {
    for (int j = 0; j < d.n_taps; ++j) kp.tap_row[j] = d.tap_off[j];
  }
  const size_t smem_bytes = fixed + (resident ? w_res : 0) + (kp.bias_mma ? kBiasMmaBytes : 0) +
                            stages * (stage_bytes + mstage) +
                            static_cast<size_t>(nstg) * n_out * n_planes * kTileBytes +
                            static_cast<size_t>(rstages) * n_planes * kTileBytes;

  ConvMaps maps;
  memset(&maps, 0, sizeof(maps));
  const uint32_t a_box = halo ? static_cast<uint32_t>(box_rows) : kTileM;
  for (int pl = 0; pl < d.n_aplanes; ++pl)
    if (!make_tmap_3d_strided(&maps.A[pl], d.a_ptr[pl], d.a_inner, d.a_rows, d.B,
                              d.a_row_stride_bytes, d.a_batch_stride_bytes, 64, a_box))
      return 3;
  const uint64_t wrows = static_cast<uint64_t>(d.n_wplanes) * d.n_taps * F;
  if (!make_tmap_2d(&maps.W, d.w_ptr, static_cast<uint64_t>(d.KC) * 64, wrows, 64,
                    cg2 ? NT / 2 : NT))
    return 3;
  for (int pl = 0; pl < n_planes; ++pl) {
    if (has_out && !make_tmap_3d_sw64(&maps.O[pl], d.out[pl], F, d.L_out, d.B)) return 3;
    if (has_out2 && !make_tmap_3d_sw64(&maps.P[pl], d.out2[pl], F, d.L_out, d.B)) return 3;
    if (rstages > 0 && !make_tmap_3d(&maps.R[pl], d.resid[pl], F, d.L_res, d.B, 64, kTileM))
      return 3;
  }

  kp.cg2 = cg2 ? 1 : 0;
  if (cg2) kp.tiles_per_seq = (kp.tiles_per_seq + 1) / 2;  // pair tiles of 256 rows
  const int total_tiles = kp.B * kp.tiles_per_seq * kp.n_ntiles;
  int walkers = cg2 ? cg2_max_pairs<RESID>() : sm_count();  // CTAs, or resident CTA pairs
  if (walkers > total_tiles) walkers = total_tiles;
  const int grid = cg2 ? 2 * walkers : walkers;
  kp.step_nt = walkers % kp.n_ntiles;
  kp.step_mt = (walkers / kp.n_ntiles) % kp.tiles_per_seq;
  kp.step_b = walkers / (kp.n_ntiles * kp.tiles_per_seq);
  auto inv32 = [](int d) -> unsigned {
    return d <= 1 ? 0u : static_cast<unsigned>(((1ull << 32) + static_cast<unsigned>(d) - 1) / static_cast<unsigned>(d));
  };
  BP_REQUIRE(grid < 65536 && kp.n_ntiles < 65536 && kp.tiles_per_seq < 65536,
             "%s: tile counts out of range", d.who);
  kp.inv_ntiles = inv32(kp.n_ntiles);
  kp.inv_tiles_per_seq = inv32(kp.tiles_per_seq);

  switch (d.epi) {
    case EPI_FWD:
      return dispatch_planes<EPI_FWD, RESID>(n_planes, NT, resident, halo, maps, kp, smem_bytes, grid, stream);
    case EPI_FWD_Z:
      return dispatch_planes<EPI_FWD_Z, RESID>(n_planes, NT, resident, halo, maps, kp, smem_bytes, grid, stream);
    case EPI_FWD_R:
      return dispatch_planes<EPI_FWD_R, RESID>(n_planes, NT, resident, halo, maps, kp, smem_bytes, grid, stream);
    case EPI_BWD_MASK:
      return dispatch_planes<EPI_BWD_MASK, RESID>(n_planes, NT, resident, halo, maps, kp, smem_bytes, grid, stream);
    default:
      return dispatch_planes<EPI_BWD_MULT, RESID>(n_planes, NT, resident, halo, maps, kp, smem_bytes, grid, stream);
  }
}

}  // namespace bp
