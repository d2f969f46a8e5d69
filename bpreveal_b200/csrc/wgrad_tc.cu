// K7: every parameter gradient of the network as one split-K tcgen05 kernel: the width-3 dilated
// convolutions (models.py:92-96 of the reference; TensorFlow computes them as
// Conv2DBackpropFilter + BiasAddGrad), the one-hot input convolution (models.py:85-90) and the
// profile / counts heads (models.py:39-49).
//
//   dw[j][c][n] = sum_{b,t} act[b, t + j*dil, c] * gz[b,t,n]     db[n] = sum_{b,t} gz[b,t,n]
//
// For the input convolution `act` is the implicit im2col view of the 8-channel one-hot tensor,
// for the heads `gz` is the implicit im2col view of the packed loss gradient d8 (overlapping-row
// TMA maps, see rowconv_tc.cu); the reduction kernel scatters the sums into the Keras layouts.
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

#include <cstdlib>

namespace bp {

constexpr int kWgTile = 128 * 128;  // bytes of one [128 x 64] bf16 tile
constexpr int kWgMaxStages = 8;

constexpr int kWgMaxJobs = 12;
constexpr int kWgMaskStage = (128 + 2) * 8;  // XF: mask words of one G tile (+ alignment slack)

// One launch serves up to kWgMaxJobs independent contractions of the same shape class (all tower
// layers of a step): job j owns the CTAs [cta_begin[j], cta_begin[j+1]) -- a share of the SMs
// proportional to its position tiles -- so that split-K partials, prologue and tail are paid once
// per step instead of once per layer.
struct WgradKParams {
  int B, NC;
  int n_passes, pa_bits, pg_bits, db_mask;  // pass i: A plane (pa_bits >> i) & 1, G plane likewise
  int n_units, UG, n_groups, n_ntiles, n_items;  // n_units = the largest job's
  int n_units_j[kWgMaxJobs];  // units (tap x chunk) of job j; jobs may differ when n_groups == 1
  int stages, stage_bytes;
  int Ncols;   // columns of the result (NT * n_ntiles)
  int n_jobs;
  int cta_begin[kWgMaxJobs + 1];
  int tiles_per_seq[kWgMaxJobs];
  int dil[kWgMaxJobs];
  int halo[kWgMaxJobs];        // 1: ONE TMA box of 128 + (n_taps-1)*dil rows serves every tap
  int halo_bytes[kWgMaxJobs];  // bytes TMA writes for that box
  // STRIP staging of the A operand (window gradients over a 16-byte-per-position tensor): the raw
  // strip of 128 + 8*n_units positions is copied once per position tile and read through
  // NON-swizzled MN-major descriptors (LBO 128 B along K, SBO 16 B along MN), unit u at +128*u B.
  int pair_units;  // 1: two units per M = 128 MMA
  int strip[kWgMaxJobs], strip_bytes[kWgMaxJobs];
  long long a_batch_bytes[kWgMaxJobs];
  const uint8_t* a_base[kWgMaxJobs];
  long long ws_off[kWgMaxJobs];  // float offset of the job's [ksplit][(n_units*64+1)*Ncols] partials
  // XF (bf16 path, 64 columns): job j's G operand is g (.) gmask[j] -- the tile TMA delivers is
  // ANDed with the 1-bit ReLU mask of its rows in shared memory by the (otherwise idle) epilogue
  // warps before the MMA reads it, so the backward never materialises gz = g (.) mask.
  // gmask[j]: [B][g_rows[j]] uint64, or NULL; the words of a tile's rows arrive by a 1-D bulk
  // copy on the tile's own mbarrier (see XfCopy in conv_core.cuh for the alignment scheme).
  const uint8_t* gmask[kWgMaxJobs];
  int g_rows[kWgMaxJobs];
  float* ws;
  unsigned* dbg;
};

struct WgradMaps {
  CUtensorMap A[2][kWgMaxJobs], G[2][kWgMaxJobs];
};

template <int NT, bool XF = false>
__global__ void __launch_bounds__(192, 1)
wgrad3_tc_kernel(const __grid_constant__ WgradMaps tm, const WgradKParams p) {
  constexpr int NSUB = NT / 64;
  constexpr uint32_t IDESC = umma_idesc_bf16(64, NT, 1, 1);
  // Two units per MMA (M = 128): the second 64-row block of the MN-major A operand lies one
  // "leading byte offset" after the first -- the next tile (SLAB), dil rows further into the box
  // (HALO), contiguous (STRIP) or, for the last unit of an odd count, the constant tile of ones
  // whose product is the bias gradient.  Halves the tensor-pipe time of this kernel, which the
  // M = 64 form kept 69 % busy re-reading operands (profiles/other_kernels_full_r1_final_summary).
  constexpr uint32_t IDESC2 = umma_idesc_bf16(128, NT, 1, 1);
  constexpr uint32_t TMEM_COLS = 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const size_t stage_bytes = static_cast<size_t>(p.stage_bytes);
  // the ones tile sits ABOVE the ring: a pair (unit, ones) reaches it with a positive offset
  uint8_t* ring = smem;
  uint8_t* ones = ring + p.stages * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ones + kWgTile);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWgMaxStages;
  uint64_t* acc_full = bars + 2 * kWgMaxStages;
  uint64_t* xfull = acc_full + 1;  // XF: "G tile masked, the MMA may read it"
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xfull + kWgMaxStages);
  uint8_t* mring = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~static_cast<uintptr_t>(15));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#ifdef BPB_ROLE_PROFILE
  const long long t_kernel_start = clock64();
  if (threadIdx.x < 32) prof_smem()[threadIdx.x] = 0;
#define WG_STAMP(slot) prof_smem()[slot] = static_cast<unsigned>(clock64() - t_kernel_start)
#else
#define WG_STAMP(slot)
#endif
  int job = 0;
  while (job + 1 < p.n_jobs && static_cast<int>(blockIdx.x) >= p.cta_begin[job + 1]) ++job;
  const int local = blockIdx.x - p.cta_begin[job];
  const int ksplit = (p.cta_begin[job + 1] - p.cta_begin[job]) / p.n_items;
  const int tiles_per_seq = p.tiles_per_seq[job];
  const int dil = p.dil[job];
  const bool halo = p.halo[job] != 0;
  const CUtensorMap* tmA0 = &tm.A[0][job];
  const CUtensorMap* tmA1 = &tm.A[1][job];
  const CUtensorMap* tmG0 = &tm.G[0][job];
  const CUtensorMap* tmG1 = &tm.G[1][job];
  const int item = local % p.n_items;
  const int split = local / p.n_items;
  const int grp = item / p.n_ntiles;
  const int nt = item % p.n_ntiles;
  const int n_units = p.n_units_j[job];
  const bool strip = p.strip[job] != 0;
  const uint8_t* const gmask = XF ? p.gmask[job] : nullptr;
  const bool xf = XF && gmask != nullptr;  // (compile-time off: the default launch pays nothing)
  const int u0 = grp * p.UG;
  const int nu = min(p.UG, n_units - u0);
  const bool do_db = (grp == 0);
  // accumulator slots of NT columns: real pairs, then the odd unit (paired with the ones tile on
  // the single-pass path), then -- if it is not paired -- the bias-gradient unit
  const int n_pairs = p.pair_units ? (nu >> 1) : 0;
  const int n_single = nu - 2 * n_pairs;  // M = 64 units when pairing is off, else 0 | 1
  // (not for STRIP jobs: in the non-swizzled MN-major form the rows of an operand are contiguous,
  // there is no second-block offset to point at the ones tile)
  const bool pair_ones = p.pair_units && n_single == 1 && do_db && p.n_passes == 1 && !strip;
  const int db_slot = n_pairs + n_single;  // used when !pair_ones
  const int kblocks = p.B * tiles_per_seq;
  const int my_kb = (kblocks - split + ksplit - 1) / ksplit;  // split < ksplit <= kblocks
  const int visits = my_kb * p.n_passes;

  pdl_launch_dependents();
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(tmA0);
    tma_prefetch_desc(tmG0);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
      mbar_init(&xfull[i], 4);
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
  if (threadIdx.x == 0) WG_STAMP(20);

  // everything below reads or writes global memory of earlier kernels (operands, workspace)
  pdl_wait();
  if (threadIdx.x == 0) WG_STAMP(21);
  if (warp == 0) {
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int v = 0; v < visits; ++v) {
        const int kb = split + (v / p.n_passes) * ksplit;
        const int pass = v % p.n_passes;
        const int b = kb / tiles_per_seq;
        const int t0 = (kb % tiles_per_seq) * 128;
        mbar_wait_wd(&empty[stage], phase ^ 1, p.dbg, 0x1100 + stage, v);
        // XF: mask words of rows [t0, t0 + 128) of sequence b (16-byte aligned source / target)
        long long m_src = 0;
        uint32_t m_dst = 0, m_bytes = 0;
        if (xf) {
          const long long total = static_cast<long long>(p.B) * p.g_rows[job];
          const long long g0 = static_cast<long long>(b) * p.g_rows[job] + t0;
          m_src = g0 & ~1LL;
          long long ce = g0 + 128;
          if (ce > total) ce = total;
          ce = (ce + 1) & ~1LL;
          m_bytes = static_cast<uint32_t>((ce - m_src) * 8);
        }
        mbar_expect_tx(&full[stage],
                       (strip ? static_cast<uint32_t>(NSUB * kWgTile + p.strip_bytes[job])
                        : halo  ? static_cast<uint32_t>(NSUB * kWgTile + p.halo_bytes[job])
                                : static_cast<uint32_t>((NSUB + nu) * kWgTile)) + m_bytes);
        uint8_t* dst = ring + stage * stage_bytes;
        if (xf) bulk_load_1d(mring + stage * kWgMaskStage + m_dst, gmask + m_src * 8, m_bytes,
                             &full[stage]);
        const CUtensorMap* mg = ((p.pg_bits >> pass) & 1) ? tmG1 : tmG0;
        const CUtensorMap* ma = ((p.pa_bits >> pass) & 1) ? tmA1 : tmA0;
        for (int s = 0; s < NSUB; ++s)
          tma_load_3d(mg, dst + s * kWgTile, &full[stage], nt * NT + s * 64, t0, b);
        if (strip) {
          bulk_load_1d(dst + NSUB * kWgTile,
                       p.a_base[job] + static_cast<size_t>(b) * p.a_batch_bytes[job] +
                           static_cast<size_t>(t0) * 16,
                       static_cast<uint32_t>(p.strip_bytes[job]), &full[stage]);
        } else if (halo) {
          // rows t0 .. t0 + 127 + (n_taps-1)*dil of the single channel chunk: every tap at once
          tma_load_3d(ma, dst + NSUB * kWgTile, &full[stage], 0, t0, b);
        } else {
          for (int i = 0; i < nu; ++i) {
            const int u = u0 + i;
            const int j = u / p.NC, c = u % p.NC;
            tma_load_3d(ma, dst + (NSUB + i) * kWgTile, &full[stage], c * 64, t0 + j * dil, b);
          }
        }
        if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      uint32_t db_started = 0;
      // Descriptor words that do not depend on the stage are built once: the issue loop is one
      // integer add per MMA (the single issuing thread is otherwise the bottleneck).
      constexpr uint32_t DHI = umma_desc_hi(1024);
      const uint32_t lbo_bits = ((static_cast<uint32_t>(kWgTile) >> 4) & 0x3FFFu) << 16;
      const uint32_t ring_lo = smem_u32(ring) >> 4;
      const uint32_t stage_lo = static_cast<uint32_t>(stage_bytes) >> 4;
      const uint32_t ones_lo = (smem_u32(ones) >> 4) | lbo_bits;
      // offset (in 16-byte units) of unit i's A tile inside a stage
      uint32_t a_off[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        a_off[i] = strip ? (NSUB * kWgTile + (u0 + i) * 128) >> 4
                  : halo  ? (NSUB * kWgTile + (u0 + i) * dil * 128) >> 4
                          : ((NSUB + i) * kWgTile) >> 4;
      // STRIP: no swizzle, LBO = 128 B (8 positions along K), SBO = 16 B (8 elements along MN),
      // K advances by 16 positions = 256 B per MMA
      const uint32_t a_lbo = strip ? ((128u >> 4) << 16) : lbo_bits;
      const uint32_t a_hi = strip ? ((16u >> 4) | (1u << 14)) : DHI;
      const uint32_t a_kstep = strip ? (256u >> 4) : 128u;
      // distance between the two units of a pair, as the descriptor's leading byte offset
      const uint32_t pair_lbo = strip ? a_lbo
                                : halo ? ((static_cast<uint32_t>(dil) * 128u) >> 4) << 16
                                       : lbo_bits;
      for (int v = 0; v < visits; ++v) {
        const int pass = v % p.n_passes;
        mbar_wait_wd(xf ? &xfull[stage] : &full[stage], phase, p.dbg, 0x1500 + stage, v);
#ifdef BPB_ROLE_PROFILE
        if (v == 0) WG_STAMP(22);
#endif
        tc_fence_after();
        const uint32_t g_lo = ((ring_lo + stage * stage_lo) & 0x3FFFu) | lbo_bits;
        const uint32_t st_lo = (ring_lo + stage * stage_lo) & 0x3FFFu;
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
          if (pr < n_pairs) {
            // HALO: tap i starts i*dil rows into the box (the 128-byte swizzle is a function of
            // the absolute shared-memory address, so an MN-major operand may start on any row)
            const uint32_t a_lo = (st_lo + a_off[2 * pr]) | pair_lbo;
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_bf16_lohi(tmem_base + pr * NT, a_lo + k * a_kstep, a_hi, g_lo + k * 128u, DHI,
                             IDESC2, (v | k) != 0 ? 1u : 0u);
          }
        }
        if (pair_ones) {
          // (last unit, ones): the second block starts at the ones tile, wherever this stage is
          const uint32_t a_addr = st_lo + a_off[nu - 1];
          const uint32_t lbo = ((smem_u32(ones) >> 4) - a_addr) & 0x3FFFu;
          const uint32_t a_lo = a_addr | (lbo << 16);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16_lohi(tmem_base + n_pairs * NT, a_lo + k * a_kstep, a_hi, g_lo + k * 128u, DHI,
                           IDESC2, (v | k) != 0 ? 1u : 0u);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (i < n_single) {
              const uint32_t a_lo = (st_lo + a_off[2 * n_pairs + i]) | a_lbo;
#pragma unroll
              for (int k = 0; k < 8; ++k)
                umma_bf16_lohi(tmem_base + (n_pairs + i) * NT, a_lo + k * a_kstep, a_hi,
                               g_lo + k * 128u, DHI, IDESC, (v | k) != 0 ? 1u : 0u);
            }
          }
          if (do_db && ((p.db_mask >> pass) & 1)) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_bf16_lohi(tmem_base + db_slot * NT, ones_lo + k * 128u, DHI, g_lo + k * 128u, DHI,
                             IDESC, (db_started | k) != 0 ? 1u : 0u);
            db_started = 1;
          }
        }
        umma_commit(&empty[stage]);
        if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1; }
      }
      umma_commit(acc_full);
      WG_STAMP(23);
    }
  } else {
    if (xf) {
      // XF masker: until the accumulators are final these four warps have nothing else to do.
      // Byte c of a row's mask word covers its 16-byte chunk c.  Lanes 0-15 of a warp take the
      // first halves (four chunks) of 16 consecutive rows, lanes 16-31 the second halves: every
      // quarter-warp touches eight different rows at the same logical chunk, which is
      // conflict-free under the 128-byte swizzle.  Four warps cover 64 rows per step.
      const int xrow = (warp - 2) * 16 + (lane & 15);
      const uint32_t xhalf = static_cast<uint32_t>(lane) >> 4;
      uint32_t xoff[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) xoff[k] = sw128(static_cast<uint32_t>(xrow), xhalf * 4u + k);
      uint32_t stage = 0, phase = 0;
      for (int v = 0; v < visits; ++v) {
        const int kb = split + (v / p.n_passes) * ksplit;
        const int b = kb / tiles_per_seq;
        const int t0 = (kb % tiles_per_seq) * 128;
        const uint32_t par =
            static_cast<uint32_t>((static_cast<long long>(b) * p.g_rows[job] + t0) & 1);
        const uint32_t tile_a = smem_u32(ring + stage * stage_bytes);
        const uint32_t mb_a = smem_u32(mring + stage * kWgMaskStage) + 8u * par + 8u * xrow +
                              4u * xhalf;
        mbar_wait_wd(&full[stage], phase, p.dbg, 0x1900 + stage, v);
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const uint32_t toff = tile_a + 64u * 128u * it;
          uint32_t m4;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(m4) : "r"(mb_a + 8u * 64u * it));
          uint4 x[4];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(x[k].x), "=r"(x[k].y), "=r"(x[k].z), "=r"(x[k].w)
                         : "r"(toff + xoff[k]));
          uint32_t ea[4], eb[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            ea[k] = ((m4 >> (8 * k)) & 0xFu) * 0x10204080u;
            eb[k] = ((m4 >> (8 * k + 4)) & 0xFu) * 0x10204080u;
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            x[k].x &= prmt(ea[k], 0u, 0x9988u);
            x[k].y &= prmt(ea[k], 0u, 0xBBAAu);
            x[k].z &= prmt(eb[k], 0u, 0x9988u);
            x[k].w &= prmt(eb[k], 0u, 0xBBAAu);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k)
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(toff + xoff[k]),
                         "r"(x[k].x), "r"(x[k].y), "r"(x[k].z), "r"(x[k].w)
                         : "memory");
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&xfull[stage]);
        if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1; }
      }
    }
    // epilogue: TMEM partial sums -> workspace slice of this split
    const int q = warp & 3;
    mbar_wait_wd(acc_full, 0, p.dbg, 0x1800, 0);
    tc_fence_after();
    if (threadIdx.x == 64) WG_STAMP(24);
    const size_t rows = static_cast<size_t>(n_units) * 64;
    float* wsp = p.ws + p.ws_off[job] + static_cast<size_t>(split) * (rows + 1) * p.Ncols;
    // M = 128 slots: accumulator row = TMEM lane, rows 0-63 the first unit of the pair, 64-127
    // the second (or the ones tile: every one of its rows is the bias gradient)
    for (int pr = 0; pr < n_pairs + (pair_ones ? 1 : 0); ++pr) {
      const bool ones_half = pr == n_pairs && q >= 2;
      const int u = u0 + 2 * pr + (q >> 1);
      const int m = (q & 1) * 32 + lane;
      for (int col = 0; col < NT; col += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + pr * NT + col + (static_cast<uint32_t>(q * 32) << 16), r);
        tmem_ld_wait();
        if (!ones_half || (q == 2 && lane == 0)) {
          float4* dst = reinterpret_cast<float4*>(
              ones_half ? wsp + rows * p.Ncols + nt * NT + col
                        : wsp + (static_cast<size_t>(u) * 64 + m) * p.Ncols + nt * NT + col);
#pragma unroll
          for (int x = 0; x < 8; ++x)
            dst[x] = make_float4(__uint_as_float(r[4 * x]), __uint_as_float(r[4 * x + 1]),
                                 __uint_as_float(r[4 * x + 2]), __uint_as_float(r[4 * x + 3]));
        }
      }
    }
    for (int i = 0; i < (pair_ones ? 0 : n_single); ++i) {
      const int u = u0 + 2 * n_pairs + i;
      // M=64 accumulators occupy lanes 0-15 of each TMEM quadrant: row = 16*q + lane.
      const int m = 16 * q + lane;
      for (int col = 0; col < NT; col += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + (n_pairs + i) * NT + col + (static_cast<uint32_t>(q * 32) << 16), r);
        tmem_ld_wait();
        if (lane < 16) {
          float4* dst = reinterpret_cast<float4*>(
              wsp + (static_cast<size_t>(u) * 64 + m) * p.Ncols + nt * NT + col);
#pragma unroll
          for (int x = 0; x < 8; ++x)
            dst[x] = make_float4(__uint_as_float(r[4 * x]), __uint_as_float(r[4 * x + 1]),
                                 __uint_as_float(r[4 * x + 2]), __uint_as_float(r[4 * x + 3]));
        }
      }
    }
    if (do_db && q == 0 && !pair_ones) {
      for (int col = 0; col < NT; col += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + db_slot * NT + col, r);
        tmem_ld_wait();
        if (lane == 0) {
          float* dst = wsp + rows * p.Ncols + nt * NT + col;
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
#ifdef BPB_ROLE_PROFILE
  if (threadIdx.x < 32 && p.dbg != nullptr && blockIdx.x < 8) {
    unsigned v = prof_smem()[threadIdx.x];
    if (threadIdx.x == 15) v = static_cast<unsigned>(clock64() - t_kernel_start);
    if (threadIdx.x == 16) v = static_cast<unsigned>(visits);
    atomicAdd(p.dbg + 1024 + 32 * blockIdx.x + threadIdx.x, v);
  }
#endif
}

// Scatter of one reduced element into the caller's Keras-layout gradient tensors.
struct WgradScatter {
  int mode;  // 0 tower conv, 1 input conv, 2 heads
  int rows, cols;
  int W, Fw, T, H, Cp;
  float *o0, *o1, *o2;  // mode 0: dw, db | mode 1: dw1, db1 | mode 2: dpw, -, dcw
};

__device__ __forceinline__ void wgrad_scatter(const WgradScatter& sc, long long i, float tot) {
  // (rows + 1) * cols < 2^31 for every geometry the planner accepts: 32-bit divisions
  const unsigned n_main = static_cast<unsigned>(sc.rows) * static_cast<unsigned>(sc.cols);
  const unsigned iu = static_cast<unsigned>(i);
  const bool is_b = iu >= n_main;
  const int r = is_b ? 0 : static_cast<int>(iu / static_cast<unsigned>(sc.cols));
  const int n = static_cast<int>(is_b ? iu - n_main : iu - static_cast<unsigned>(r) * sc.cols);
  if (sc.mode == 0) {
    if (!is_b) sc.o0[i] = tot;
    else if (sc.o1) sc.o1[n] = tot;
  } else if (sc.mode == 1) {
    // rows = im2col index k = j*8 + c of the one-hot window, cols = filters
    if (n >= sc.Fw) return;
    if (is_b) {
      if (sc.o1) sc.o1[n] = tot;
    } else {
      const int j = r >> 3, c = r & 7;
      if (j < sc.W && c < 4) sc.o0[(j * 4 + c) * sc.Fw + n] = tot;
    }
  } else if (sc.mode == 3) {
    // transposed heads job: rows = im2col index kk = j'*Cp + ch of d8, cols = tower channel c
    if (is_b || n >= sc.Fw) return;
    const int jp = r / sc.Cp, ch = r % sc.Cp;
    if (jp >= sc.W) return;
    if (ch < sc.T) sc.o0[(static_cast<long long>(sc.W - 1 - jp) * sc.Fw + n) * sc.T + ch] = tot;
    else if (ch < sc.T + sc.H && jp == 0) sc.o2[n * sc.H + (ch - sc.T)] = tot;
  } else {
    // rows = tower channel c, cols = im2col index kk = j'*Cp + ch of the packed loss gradient
    const int jp = n / sc.Cp, ch = n % sc.Cp;
    if (jp >= sc.W) return;
    if (is_b) return;  // the heads' bias gradients are computed in fp32 by bpb_heads_pack_grad
    if (r < sc.Fw) {
      if (ch < sc.T) sc.o0[(static_cast<long long>(sc.W - 1 - jp) * sc.Fw + r) * sc.T + ch] = tot;
      else if (ch < sc.T + sc.H && jp == 0) sc.o2[r * sc.H + (ch - sc.T)] = tot;
    }
  }
}

struct WgradReduceJobs {
  int n_jobs;
  int ksplit[kWgMaxJobs];
  long long ws_off[kWgMaxJobs];
  WgradScatter sc[kWgMaxJobs];  // per-job geometry and outputs
};

// out[i] = sum_s ws[s][i] in a fixed order (deterministic).  One thread per output of job
// blockIdx.y, four independent chains over the splits (s = 0, 4, ... | 1, 5, ... | ...), every load
// of a warp one coalesced 128-byte line.  (The first version gave a block 32 outputs and spread
// the splits over its eight warps: 4 246 blocks of almost no work each for C1, 10.6 us -- a
// tenth of the weight-gradient family -- for 7 MB of partial sums.)
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ ws_all, const WgradReduceJobs jobs) {
  pdl_launch_dependents();
  const int job = blockIdx.y;
  const int ksplit = jobs.ksplit[job];
  const WgradScatter sc = jobs.sc[job];
  const long long n = (static_cast<long long>(sc.rows) + 1) * sc.cols;
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  pdl_wait();  // the partial sums are the predecessor's output
  if (i >= n) return;
  const float* ws = ws_all + jobs.ws_off[job] + i;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int s = 0;
  for (; s + 3 < ksplit; s += 4) {
    s0 += ws[(s + 0) * n];
    s1 += ws[(s + 1) * n];
    s2 += ws[(s + 2) * n];
    s3 += ws[(s + 3) * n];
  }
  if (s < ksplit) s0 += ws[s * n];
  if (s + 1 < ksplit) s1 += ws[(s + 1) * n];
  if (s + 2 < ksplit) s2 += ws[(s + 2) * n];
  wgrad_scatter(sc, i, (s0 + s1) + (s2 + s3));
}

// One contraction: out[u*64 + m][n] = sum over positions of A_u[pos][m] * G[pos][n], u = (tap, chunk).
struct WgradDesc {
  const char* who;
  int B, rows;      // sequences, positions per sequence (the K extent)
  int n_taps, NC;   // units: tap j (A rows shifted by j*dil) x 64-channel chunk c of the A operand
  int dil;
  int Ncols;        // result columns = channels of the G operand (multiple of 64)
  // operands as strided bf16 tensors [B][rows][inner]
  const void* a_ptr[2];
  unsigned long long a_inner, a_rows, a_row_stride_bytes, a_batch_stride_bytes;
  const void* g_ptr[2];
  unsigned long long g_inner, g_rows, g_row_stride_bytes, g_batch_stride_bytes;
  const void* g_mask;  // XF: the G operand is g (.) g_mask ([B][g_rows] uint64, 64 columns) or NULL
  int n_passes, pa_bits, pg_bits, db_mask;
  bool allow_halo;
  bool allow_strip;  // A rows are one 16-byte position apart (n_taps == 1, single A plane)
  float* ws;
  long long ws_bytes;
  WgradScatter sc;
};

struct WgradPlan {
  int NT, UG, n_units, n_groups, n_ntiles, n_items, stages;
  int n_units_j[kWgMaxJobs];
  int ksplit[kWgMaxJobs];
  int halo[kWgMaxJobs], halo_rows[kWgMaxJobs];
  long long ws_off[kWgMaxJobs];
  size_t stage_bytes, smem_bytes, ws_bytes;
  int grid;
};

// Jobs must agree on Ncols; rows / dilation may differ, and so may the unit count (and the STRIP
// staging of the A operand) as long as every job fits one accumulator group.
static bool plan_wgrad(int n_jobs, int B, const int* rows, const int* halo_rows_in,
                       const int* n_units_j, int Ncols, WgradPlan* pl,
                       const bool* strip_j = nullptr) {
  if (n_jobs < 1 || n_jobs > kWgMaxJobs) return false;
  // BPB_WG_NT caps the column tile (256 | 128 | 64; default 128): a 128-column tile leaves TMEM
  // room for three units per CTA, which the paired M = 128 MMAs need (F = 512 weight gradients:
  // 520 TFLOP/s with 256-column tiles and one unit per CTA, 685 with 128)
  static const int nt_cap = [] {
    const char* e = getenv("BPB_WG_NT");
    const int v = e ? atoi(e) : 128;
    return (v == 64 || v == 256) ? v : 128;
  }();
  int NT = (Ncols % 256 == 0) ? 256 : (Ncols % 128 == 0 ? 128 : 64);
  if (NT > nt_cap) NT = nt_cap;
  pl->NT = NT;
  int n_units = 0;
  bool uniform = true;
  for (int j = 0; j < n_jobs; ++j) {
    pl->n_units_j[j] = n_units_j[j];
    if (n_units_j[j] > n_units) n_units = n_units_j[j];
    if (n_units_j[j] != n_units_j[0]) uniform = false;
  }
  pl->n_units = n_units;
  // Accumulator slots of NT columns: one is reserved for the bias-gradient unit, every other one
  // holds a PAIR of units (M = 128 MMAs) when pairing is on.  More units per CTA = more FLOPs per
  // staged G tile: at F = 512 the launch is bound by what one SM can pull out of L2 (80 KB per
  // 6.3 MFLOP visit with three units), so four units (96 KB per 8.4 MFLOP) are worth ~10 %.
  // BPB_WG_UG caps the units per CTA (1..4).
  static const int ug_cap = [] {
    const char* e = getenv("BPB_WG_UG");
    const int v = e ? atoi(e) : 4;
    return (v >= 1 && v <= 4) ? v : 4;
  }();
  static const bool pair_on = [] {
    const char* e = getenv("BPB_WG_PAIR");
    return !(e && e[0] == '0');
  }();
  int ug = (512 / NT - 1) * (pair_on ? 2 : 1);
  if (ug > n_units) ug = n_units;
  if (ug > ug_cap) ug = ug_cap;
  pl->UG = ug;
  pl->n_groups = (n_units + ug - 1) / ug;
  if (!uniform && pl->n_groups != 1) return false;
  pl->n_ntiles = Ncols / NT;
  pl->n_items = pl->n_groups * pl->n_ntiles;
  const int sms = sm_count();
  if (sms <= 0) return false;
  // split-K factors: the SMs are shared between the jobs in proportion to their position tiles
  // (BPB_WG_STRIP_COST weights the STRIP-staged jobs: they were MMA-bound at ~1.7k against ~1.4k
  // cycles per tile with M = 64 MMAs (best weight 1.25); with paired M = 128 MMAs 1.0 measures best)
  static const double strip_cost = [] {
    const char* e = getenv("BPB_WG_STRIP_COST");
    const double v = e ? atof(e) : 0.0;
    return v > 0.1 ? v : 1.0;
  }();
  double total_kb = 0;
  int kb[kWgMaxJobs];
  double cost[kWgMaxJobs];
  for (int j = 0; j < n_jobs; ++j) {
    kb[j] = B * ((rows[j] + 127) / 128);
    cost[j] = kb[j] * ((strip_j && strip_j[j] && n_jobs > 1) ? strip_cost : 1.0);
    total_kb += cost[j];
  }
  int budget = sms / pl->n_items;  // splits available in total
  if (budget < 2 * n_jobs) {
    // Wide models: an (unit group, column tile) item per job already over-subscribes the SMs
    // (F = 512, 11 layers: 264 CTAs of very different lengths on 148 SMs).  Split the long jobs
    // further so that the CTAs pack: ~BPB_WG_WAVES (default 4) waves of CTAs in total (measured
    // at C4, batch 64: 9.5 ms with one split per job, 8.6 at 2.5 waves, 8.0 at 4, 8.2 at 6).
    static const double waves = [] {
      const char* e = getenv("BPB_WG_WAVES");
      const double v = e ? atof(e) : 0.0;
      return v >= 1.0 ? v : 4.0;
    }();
    const int want = static_cast<int>(waves * sms / pl->n_items);
    budget = want > n_jobs ? want : n_jobs;
  }
  int used = 0;
  for (int j = 0; j < n_jobs; ++j) {
    int ks = static_cast<int>(budget * cost[j] / total_kb);
    if (ks < 1) ks = 1;
    if (ks > kb[j]) ks = kb[j];
    pl->ksplit[j] = ks;
    used += ks;
  }
  // hand the remainder to the largest jobs
  for (int guard = 0; used < budget && guard < 4 * kWgMaxJobs; ++guard) {
    int best = -1;
    double best_load = 0;
    for (int j = 0; j < n_jobs; ++j) {
      const double load = cost[j] / pl->ksplit[j];
      if (pl->ksplit[j] < kb[j] && load > best_load) { best_load = load; best = j; }
    }
    if (best < 0) break;
    ++pl->ksplit[best];
    ++used;
  }
  pl->grid = used * pl->n_items;
  // staging: HALO (one box for all taps) where the dilation span allows it, else one tile per unit
  size_t stage = 0;
  for (int j = 0; j < n_jobs; ++j) {
    const int hr = halo_rows_in ? halo_rows_in[j] : 0;
    pl->halo[j] = hr > 0 && hr <= 256 && pl->n_groups == 1;
    pl->halo_rows[j] = pl->halo[j] ? hr : 0;
    const size_t sb = (strip_j && strip_j[j] && pl->n_groups == 1)
                          ? static_cast<size_t>(NT / 64) * kWgTile +
                                ((128 + 8 * n_units_j[j]) * 16 + 1023) / 1024 * 1024
                      : pl->halo[j]
                          ? static_cast<size_t>(NT / 64) * kWgTile + (hr * 128 + 1023) / 1024 * 1024
                          : static_cast<size_t>(NT / 64 + (n_units_j[j] < ug ? n_units_j[j] : ug)) *
                                kWgTile;
    if (sb > stage) stage = sb;
  }
  pl->stage_bytes = stage;
  const size_t fixed = 1024 + kWgTile + (3 * kWgMaxStages + 2) * 8 + 48 +
                       static_cast<size_t>(kWgMaxStages) * kWgMaskStage;
  const int max_smem = max_smem_optin();
  if (max_smem <= 0) return false;
  int stages = static_cast<int>((static_cast<size_t>(max_smem) - fixed) / stage);
  if (stages > kWgMaxStages) stages = kWgMaxStages;
  if (stages < 1) return false;
  pl->stages = stages;
  pl->smem_bytes = fixed + stages * stage;
  long long off = 0;
  for (int j = 0; j < n_jobs; ++j) {
    const long long per_split = (static_cast<long long>(n_units_j[j]) * 64 + 1) * Ncols;
    pl->ws_off[j] = off;
    off += per_split * pl->ksplit[j];
  }
  pl->ws_bytes = static_cast<size_t>(off) * sizeof(float);
  return true;
}

template <int NT, bool XF = false>
static int launch_wgrad(const WgradMaps& maps, const WgradKParams& kp, size_t smem, int grid,
                        cudaStream_t stream) {
  auto kern = wgrad3_tc_kernel<NT, XF>;
  static bool configured = false;
  if (!configured) {
    BP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       max_smem_optin()));
    configured = true;
  }
  BP_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(192), smem, stream, maps, kp));
  return 0;
}

static bool wgrad_strip_enabled() {
  static const bool on = [] {
    const char* e = getenv("BPB_STRIP");
    return !(e && e[0] == '0');
  }();
  return on;
}

// All descs share Ncols, batch and passes; the unit structure (taps x chunks, STRIP staging) and the
// scatter of the result may differ per job; d[0].ws is the workspace.
static int wgrad_run(const WgradDesc* d, int n_jobs, cudaStream_t stream) {
  const WgradDesc& d0 = d[0];
  BP_REQUIRE(n_jobs >= 1 && n_jobs <= kWgMaxJobs, "%s: %d jobs (1..%d supported)", d0.who, n_jobs,
             kWgMaxJobs);
  int rows[kWgMaxJobs], halo_rows[kWgMaxJobs];
  for (int j = 0; j < n_jobs; ++j) {
    BP_REQUIRE(d[j].Ncols == d0.Ncols && d[j].B == d0.B && d[j].n_passes == d0.n_passes &&
                   d[j].pa_bits == d0.pa_bits && d[j].pg_bits == d0.pg_bits &&
                   d[j].db_mask == d0.db_mask,
               "%s: every job of one launch must have the same columns, batch and passes", d0.who);
    rows[j] = d[j].rows;
    halo_rows[j] = (d[j].allow_halo && d[j].NC == 1 && d[j].n_taps > 1)
                       ? 128 + (d[j].n_taps - 1) * d[j].dil : 0;
  }
  const bool strip_enabled = wgrad_strip_enabled();
  bool want_strip[kWgMaxJobs];
  int units[kWgMaxJobs];
  for (int j = 0; j < n_jobs; ++j) {
    want_strip[j] = strip_enabled && d[j].allow_strip && d[j].n_taps == 1 &&
                    d[j].a_row_stride_bytes == 16 && d[j].a_ptr[1] == nullptr && d[j].NC <= 4;
    units[j] = d[j].n_taps * d[j].NC;
    // a job that is not STRIP-staged is addressed through the shared (taps, chunks) decoding
    BP_REQUIRE(want_strip[j] || (d[j].n_taps == d0.n_taps && d[j].NC == d0.NC) || n_jobs == 1,
               "%s: jobs of one launch must share the tap structure unless STRIP-staged", d0.who);
  }
  WgradPlan pl;
  BP_REQUIRE(plan_wgrad(n_jobs, d0.B, rows, halo_rows, units, d0.Ncols, &pl, want_strip),
             "%s: cannot plan (no device?)", d0.who);
  BP_REQUIRE(static_cast<size_t>(d0.ws_bytes) >= pl.ws_bytes, "%s: workspace too small (%lld < %zu)",
             d0.who, d0.ws_bytes, pl.ws_bytes);
  WgradKParams kp{};
  kp.B = d0.B;
  kp.NC = d0.NC;
  kp.n_passes = d0.n_passes;
  kp.pa_bits = d0.pa_bits;
  kp.pg_bits = d0.pg_bits;
  kp.db_mask = d0.db_mask;
  kp.n_units = pl.n_units;
  kp.UG = pl.UG;
  kp.n_groups = pl.n_groups;
  kp.n_ntiles = pl.n_ntiles;
  kp.n_items = pl.n_items;
  kp.stages = pl.stages;
  kp.stage_bytes = static_cast<int>(pl.stage_bytes);
  kp.Ncols = d0.Ncols;
  kp.n_jobs = n_jobs;
  static const bool pair_enabled = [] {
    const char* e = getenv("BPB_WG_PAIR");
    return !(e && e[0] == '0');
  }();
  kp.pair_units = pair_enabled ? 1 : 0;
  kp.ws = d0.ws;
  kp.dbg = debug_buffer();

  static thread_local WgradMaps maps;  // 6 KB: too large for the stack frame of every caller
  memset(&maps, 0, sizeof(maps));
  WgradReduceJobs rj{};
  rj.n_jobs = n_jobs;
  long long n_max = 0;
  int cta = 0;
  for (int j = 0; j < n_jobs; ++j) {
    kp.cta_begin[j] = cta;
    cta += pl.ksplit[j] * pl.n_items;
    kp.tiles_per_seq[j] = (d[j].rows + 127) / 128;
    kp.dil[j] = d[j].dil;
    kp.halo[j] = pl.halo[j];
    kp.halo_bytes[j] = pl.halo_rows[j] * 128;
    kp.ws_off[j] = pl.ws_off[j];
    kp.n_units_j[j] = units[j];
    if (d[j].g_mask != nullptr) {
      BP_REQUIRE(pl.NT == 64 && d[j].Ncols == 64 && d[j].n_passes == 1 &&
                     d[j].g_row_stride_bytes == 128 &&
                     d[j].g_batch_stride_bytes == 128ull * d[j].g_rows &&
                     ((static_cast<unsigned long long>(d[j].B) * d[j].g_rows) & 1ull) == 0,
                 "%s: gz_mask (operand masking in shared memory) needs the bf16 path, F = 64 and "
                 "an even B * L_out", d0.who);
      kp.gmask[j] = static_cast<const uint8_t*>(d[j].g_mask);
      kp.g_rows[j] = static_cast<int>(d[j].g_rows);
    }
    if (want_strip[j] && pl.n_groups == 1) {
      kp.strip[j] = 1;
      kp.strip_bytes[j] = (128 + 8 * units[j]) * 16;
      kp.a_batch_bytes[j] = static_cast<long long>(d[j].a_batch_stride_bytes);
      kp.a_base[j] = static_cast<const uint8_t*>(d[j].a_ptr[0]);
    }
    rj.ksplit[j] = pl.ksplit[j];
    rj.ws_off[j] = pl.ws_off[j];
    rj.sc[j] = d[j].sc;
    rj.sc[j].rows = units[j] * 64;
    rj.sc[j].cols = d[j].Ncols;
    const long long n_j = (static_cast<long long>(rj.sc[j].rows) + 1) * rj.sc[j].cols;
    if (n_j > n_max) n_max = n_j;
    for (int pl_i = 0; pl_i < 2; ++pl_i) {
      if (d[j].a_ptr[pl_i] &&
          !make_tmap_3d_strided(&maps.A[pl_i][j], d[j].a_ptr[pl_i], d[j].a_inner, d[j].a_rows,
                                d[j].B, d[j].a_row_stride_bytes, d[j].a_batch_stride_bytes, 64,
                                pl.halo[j] ? pl.halo_rows[j] : 128))
        return 3;
      if (d[j].g_ptr[pl_i] &&
          !make_tmap_3d_strided(&maps.G[pl_i][j], d[j].g_ptr[pl_i], d[j].g_inner, d[j].g_rows,
                                d[j].B, d[j].g_row_stride_bytes, d[j].g_batch_stride_bytes, 64, 128))
        return 3;
    }
  }
  kp.cta_begin[n_jobs] = cta;
  int rc;
  bool any_xf = false;
  for (int j = 0; j < n_jobs; ++j) any_xf = any_xf || kp.gmask[j] != nullptr;
  if (any_xf) rc = launch_wgrad<64, true>(maps, kp, pl.smem_bytes, pl.grid, stream);  // NT == 64
  else if (pl.NT == 64) rc = launch_wgrad<64>(maps, kp, pl.smem_bytes, pl.grid, stream);
  else if (pl.NT == 128) rc = launch_wgrad<128>(maps, kp, pl.smem_bytes, pl.grid, stream);
  else rc = launch_wgrad<256>(maps, kp, pl.smem_bytes, pl.grid, stream);
  if (rc) return rc;

  dim3 rgrid(static_cast<unsigned>((n_max + 255) / 256), static_cast<unsigned>(n_jobs));
  BP_CHECK_CUDA(launch_pdl(wgrad_reduce_kernel, rgrid, dim3(256), 0, stream,
                           static_cast<const float*>(d0.ws), rj));
  return 0;
}

static int wgrad_run(const WgradDesc& d, cudaStream_t stream) { return wgrad_run(&d, 1, stream); }

// (strip_j must say which jobs the launch will STRIP-stage: the split-K shares, and with them the
// workspace layout, depend on it)
static long long wgrad_ws_bytes(const char* who, int n_jobs, int B, const int* rows,
                                const int* n_units_j, int Ncols, const bool* strip_j = nullptr) {
  WgradPlan pl;
  if (Ncols <= 0 || Ncols % 64 != 0 || n_jobs < 1 || n_jobs > kWgMaxJobs ||
      !plan_wgrad(n_jobs, B, rows, nullptr, n_units_j, Ncols, &pl, strip_j)) {
    set_error("%s: unsupported geometry or no CUDA device", who);
    return -1;
  }
  return static_cast<long long>(pl.ws_bytes);
}

static long long wgrad_ws_bytes(const char* who, int n_jobs, int B, const int* rows, int n_units,
                                int Ncols) {
  int units[kWgMaxJobs];
  for (int j = 0; j < kWgMaxJobs; ++j) units[j] = n_units;
  return wgrad_ws_bytes(who, n_jobs, B, rows, units, Ncols);
}

static int window_kc(int W, int C) { return (W * C + 63) / 64; }

}  // namespace bp

extern "C" long long bpb_wgrad3_ws_bytes(int B, int L_out, int F, int precise) {
  (void)precise;
  return bp::wgrad_ws_bytes("bpb_wgrad3_ws_bytes", 1, B, &L_out, F > 0 ? 3 * (F / 64) : 0, F);
}

static int fill_wgrad3_desc(const char* who, const bpb_wgrad3_args* a, bp::WgradDesc* out) {
  using namespace bp;
  BP_REQUIRE(a != nullptr, "%s: null args", who);
  BP_REQUIRE(a->F > 0 && a->F % 64 == 0, "%s: F=%d must be a positive multiple of 64", who, a->F);
  BP_REQUIRE(a->act_hi && a->gz_hi && a->dw, "%s: NULL tensor", who);
  BP_REQUIRE((a->act_lo != nullptr) == (a->gz_lo != nullptr),
             "%s: act_lo and gz_lo must both be given or both be NULL", who);
  const bool precise = a->act_lo != nullptr;
  WgradDesc d{};
  d.who = who;
  d.B = a->B;
  d.rows = a->L_out;
  d.n_taps = 3;
  d.NC = a->F / 64;
  d.dil = a->dil;
  d.Ncols = a->F;
  d.a_ptr[0] = a->act_hi;
  d.a_ptr[1] = a->act_lo;
  d.a_inner = a->F;
  d.a_rows = a->L_in;
  d.a_row_stride_bytes = static_cast<unsigned long long>(a->F) * 2;
  d.a_batch_stride_bytes = d.a_row_stride_bytes * a->L_in;
  d.g_ptr[0] = a->gz_hi;
  d.g_ptr[1] = a->gz_lo;
  d.g_inner = a->F;
  d.g_rows = a->L_out;
  d.g_row_stride_bytes = static_cast<unsigned long long>(a->F) * 2;
  d.g_batch_stride_bytes = d.g_row_stride_bytes * a->L_out;
  d.g_mask = a->gz_mask;
  BP_REQUIRE(!a->gz_mask || (!precise && a->F == 64), "%s: gz_mask needs the bf16 path and F = 64",
             who);
  // hi*hi, lo*hi, hi*lo; the bias gradient adds the hi and lo planes of gz (passes 0 and 2)
  d.n_passes = precise ? 3 : 1;
  d.pa_bits = 0b010;
  d.pg_bits = 0b100;
  d.db_mask = 0b101;
  d.allow_halo = true;
  d.ws = a->ws;
  d.ws_bytes = a->ws_bytes;
  d.sc.mode = 0;
  d.sc.o0 = a->dw;
  d.sc.o1 = a->db;
  *out = d;
  return 0;
}

extern "C" int bpb_wgrad3_tc(const bpb_wgrad3_args* a, void* stream_) {
  bp::WgradDesc d;
  if (int rc = fill_wgrad3_desc("bpb_wgrad3_tc", a, &d)) return rc;
  BP_REQUIRE(a->ws != nullptr, "bpb_wgrad3_tc: NULL workspace");
  return bp::wgrad_run(d, static_cast<cudaStream_t>(stream_));
}

extern "C" long long bpb_wgrad3_multi_ws_bytes(int n_layers, int B, const int* L_out, int F) {
  if (n_layers < 1 || n_layers > bp::kWgMaxJobs || L_out == nullptr) {
    bp::set_error("bpb_wgrad3_multi_ws_bytes: 1..%d layers", bp::kWgMaxJobs);
    return -1;
  }
  return bp::wgrad_ws_bytes("bpb_wgrad3_multi_ws_bytes", n_layers, B, L_out,
                            F > 0 ? 3 * (F / 64) : 0, F);
}

extern "C" int bpb_wgrad3_multi(const bpb_wgrad3_args* layers, int n_layers, float* ws,
                                long long ws_bytes, void* stream_) {
  using namespace bp;
  BP_REQUIRE(layers != nullptr && ws != nullptr && n_layers >= 1 && n_layers <= kWgMaxJobs,
             "bpb_wgrad3_multi: 1..%d layers and a workspace are required", kWgMaxJobs);
  WgradDesc d[kWgMaxJobs];
  for (int i = 0; i < n_layers; ++i)
    if (int rc = fill_wgrad3_desc("bpb_wgrad3_multi", &layers[i], &d[i])) return rc;
  d[0].ws = ws;
  d[0].ws_bytes = ws_bytes;
  return wgrad_run(d, n_layers, static_cast<cudaStream_t>(stream_));
}

extern "C" long long bpb_input_conv_wgrad_ws_bytes(int B, int L_out, int W, int F) {
  return bp::wgrad_ws_bytes("bpb_input_conv_wgrad_ws_bytes", 1, B, &L_out, bp::window_kc(W, 8), F);
}

static bp::WgradDesc input_conv_wgrad_desc(const char* who, int B, int L_in, int L_out, int W, int Fw,
                                           int F, const bpb_bf16* x8, const bpb_bf16* gz_hi,
                                           const bpb_bf16* gz_lo, float* dw, float* db) {
  using namespace bp;
  WgradDesc d{};
  d.who = who;
  d.B = B;
  d.rows = L_out;
  d.n_taps = 1;
  d.NC = window_kc(W, 8);
  d.dil = 0;
  d.Ncols = F;
  d.a_ptr[0] = x8;
  d.a_inner = static_cast<unsigned long long>(W) * 8;
  d.a_rows = L_out;
  d.a_row_stride_bytes = 16;
  d.a_batch_stride_bytes = static_cast<unsigned long long>(L_in) * 16;
  d.g_ptr[0] = gz_hi;
  d.g_ptr[1] = gz_lo;
  d.g_inner = F;
  d.g_rows = L_out;
  d.g_row_stride_bytes = static_cast<unsigned long long>(F) * 2;
  d.g_batch_stride_bytes = d.g_row_stride_bytes * L_out;
  // the one-hot operand is exact: x*gz_hi (+ x*gz_lo)
  d.n_passes = gz_lo ? 2 : 1;
  d.pa_bits = 0;
  d.pg_bits = 0b10;
  d.db_mask = 0b11;
  d.allow_strip = true;  // x8 carries the spare tail (bpb_x8_elems)
  d.sc.mode = 1;
  d.sc.W = W;
  d.sc.Fw = Fw;
  d.sc.o0 = dw;
  d.sc.o1 = db;
  return d;
}

extern "C" int bpb_input_conv_wgrad(int B, int L_in, int L_out, int W, int Fw, int F,
                                    const bpb_bf16* x8, const bpb_bf16* gz_hi,
                                    const bpb_bf16* gz_lo, float* dw, float* db, float* ws,
                                    long long ws_bytes, void* stream_) {
  using namespace bp;
  BP_REQUIRE(x8 && gz_hi && dw && ws, "bpb_input_conv_wgrad: NULL argument");
  BP_REQUIRE(F > 0 && F % 64 == 0 && Fw <= F && L_out == L_in - W + 1,
             "bpb_input_conv_wgrad: bad geometry");
  WgradDesc d = input_conv_wgrad_desc("bpb_input_conv_wgrad", B, L_in, L_out, W, Fw, F, x8, gz_hi,
                                      gz_lo, dw, db);
  d.ws = ws;
  d.ws_bytes = ws_bytes;
  return wgrad_run(d, static_cast<cudaStream_t>(stream_));
}

// The tower layers of a step AND the input convolution in one launch (bf16 path; the jobs differ
// in their unit structure: 3 tap units HALO/SLAB-staged vs ceil(8W/64) window units STRIP-staged).
extern "C" long long bpb_wgrad_tower_input_ws_bytes(int n_layers, int B, const int* L_out, int F,
                                                    int L_out0, int W, int L_last, int W_out) {
  using namespace bp;
  const int extra = 1 + (W_out > 0 ? 1 : 0);
  if (n_layers < 1 || n_layers + extra > kWgMaxJobs || L_out == nullptr || F <= 0 || F % 64 != 0) {
    set_error("bpb_wgrad_tower_input_ws_bytes: 1..%d layers, F a multiple of 64",
              kWgMaxJobs - extra);
    return -1;
  }
  int rows[kWgMaxJobs], units[kWgMaxJobs];
  bool strip[kWgMaxJobs];
  for (int j = 0; j < n_layers; ++j) {
    rows[j] = L_out[j];
    units[j] = 3 * (F / 64);
    strip[j] = false;
  }
  rows[n_layers] = L_out0;
  units[n_layers] = window_kc(W, 8);
  strip[n_layers] = wgrad_strip_enabled() && units[n_layers] <= 4;
  if (W_out > 0) {
    rows[n_layers + 1] = L_last;
    units[n_layers + 1] = window_kc(W_out, 8);
    strip[n_layers + 1] = wgrad_strip_enabled() && units[n_layers + 1] <= 4;
  }
  // the window jobs must be STRIP-staged to share a launch with the tower's tap structure
  if (!strip[n_layers] || (W_out > 0 && !strip[n_layers + 1])) {
    set_error("bpb_wgrad_tower_input_ws_bytes: window too wide for STRIP staging");
    return -1;
  }
  return wgrad_ws_bytes("bpb_wgrad_tower_input_ws_bytes", n_layers + extra, B, rows, units, F,
                        strip);
}

extern "C" int bpb_wgrad_tower_input(const bpb_wgrad3_args* layers, int n_layers, int L_in0,
                                     int L_out0, int W, int Fw, const bpb_bf16* x8,
                                     const bpb_bf16* gz0_hi, const uint64_t* gz0_mask, float* dw1,
                                     float* db1, int L_last,
                                     int W_out, int T, int H, const bpb_bf16* act_last,
                                     const bpb_bf16* d8, float* dpw, float* dcw, float* ws,
                                     long long ws_bytes, void* stream_) {
  using namespace bp;
  const bool with_heads = d8 != nullptr;
  BP_REQUIRE(layers != nullptr && ws != nullptr && n_layers >= 1 &&
                 n_layers + 1 + (with_heads ? 1 : 0) <= kWgMaxJobs,
             "bpb_wgrad_tower_input: 1..%d layers and a workspace are required", kWgMaxJobs - 2);
  BP_REQUIRE(x8 && gz0_hi && dw1, "bpb_wgrad_tower_input: NULL argument");
  WgradDesc d[kWgMaxJobs];
  for (int i = 0; i < n_layers; ++i) {
    if (int rc = fill_wgrad3_desc("bpb_wgrad_tower_input", &layers[i], &d[i])) return rc;
    BP_REQUIRE(layers[i].act_lo == nullptr, "bpb_wgrad_tower_input: bf16 path only");
  }
  const int F = layers[0].F, B = layers[0].B;
  BP_REQUIRE(Fw <= F && L_out0 == L_in0 - W + 1, "bpb_wgrad_tower_input: bad input-conv geometry");
  d[n_layers] = input_conv_wgrad_desc("bpb_wgrad_tower_input", B, L_in0, L_out0, W, Fw, F, x8,
                                      gz0_hi, nullptr, dw1, db1);
  d[n_layers].g_mask = gz0_mask;
  // single pass: only bit 0 of the plane selectors is read; align them with the tower's
  d[n_layers].pa_bits = d[0].pa_bits;
  d[n_layers].pg_bits = d[0].pg_bits;
  d[n_layers].db_mask = d[0].db_mask;
  int n_jobs = n_layers + 1;
  if (with_heads) {
    // The heads' weight gradients, transposed so that they have the tower's column count: the A
    // operand is the STRIP of the packed loss gradient d8 (8 channels = 16 bytes per position),
    // the G operand the last tower activations:
    //   out[j'*8 + ch][c] = sum_s d8[s + j'][ch] act[s][c]   (= dpw[W-1-j'][c][ch], dcw on j' = 0)
    BP_REQUIRE(act_last && dpw && dcw && W_out >= 1 && T >= 1 && H >= 1 && T + H <= 8,
               "bpb_wgrad_tower_input: the heads job needs T + H <= 8 (one 16-byte d8 row)");
    WgradDesc h{};
    h.who = "bpb_wgrad_tower_input";
    h.B = B;
    h.rows = L_last;
    h.n_taps = 1;
    h.NC = window_kc(W_out, 8);
    h.Ncols = F;
    h.a_ptr[0] = d8;
    h.a_inner = static_cast<unsigned long long>(W_out) * 8;
    h.a_rows = L_last;
    h.a_row_stride_bytes = 16;
    h.a_batch_stride_bytes = static_cast<unsigned long long>(L_last + W_out - 1) * 16;
    h.g_ptr[0] = act_last;
    h.g_inner = F;
    h.g_rows = L_last;
    h.g_row_stride_bytes = static_cast<unsigned long long>(F) * 2;
    h.g_batch_stride_bytes = h.g_row_stride_bytes * L_last;
    h.n_passes = 1;
    h.pa_bits = d[0].pa_bits;
    h.pg_bits = d[0].pg_bits;
    h.db_mask = d[0].db_mask;
    h.allow_strip = true;  // d8 carries the spare tail (bpb_heads_d8_elems)
    h.sc.mode = 3;
    h.sc.W = W_out;
    h.sc.Fw = Fw;
    h.sc.T = T;
    h.sc.H = H;
    h.sc.Cp = 8;
    h.sc.o0 = dpw;
    h.sc.o2 = dcw;
    d[n_jobs++] = h;
  }
  d[0].ws = ws;
  d[0].ws_bytes = ws_bytes;
  return wgrad_run(d, n_jobs, static_cast<cudaStream_t>(stream_));
}

extern "C" long long bpb_heads_wgrad_ws_bytes(int B, int L_in, int W, int F, int Cp) {
  return bp::wgrad_ws_bytes("bpb_heads_wgrad_ws_bytes", 1, B, &L_in, F > 0 ? F / 64 : 0,
                            bp::window_kc(W, Cp) * 64);
}

extern "C" int bpb_heads_wgrad_tc(int B, int L_in, int W, int Fw, int F, int T, int H, int Cp,
                                  const bpb_bf16* act_hi, const bpb_bf16* act_lo,
                                  const bpb_bf16* d8_hi, const bpb_bf16* d8_lo, float* dpw,
                                  float* dcw, float* ws, long long ws_bytes, void* stream_) {
  using namespace bp;
  BP_REQUIRE(act_hi && d8_hi && dpw && dcw && ws, "bpb_heads_wgrad_tc: NULL argument");
  BP_REQUIRE((act_lo != nullptr) == (d8_lo != nullptr),
             "bpb_heads_wgrad_tc: act_lo and d8_lo must both be given or both be NULL");
  BP_REQUIRE(F > 0 && F % 64 == 0 && Fw <= F && Cp % 8 == 0 && Cp >= T + H,
             "bpb_heads_wgrad_tc: bad geometry");
  const bool precise = act_lo != nullptr;
  const int Lp = L_in + W - 1;
  WgradDesc d{};
  d.who = "bpb_heads_wgrad_tc";
  d.B = B;
  d.rows = L_in;
  d.n_taps = 1;
  d.NC = F / 64;
  d.dil = 0;
  d.Ncols = window_kc(W, Cp) * 64;
  d.a_ptr[0] = act_hi;
  d.a_ptr[1] = act_lo;
  d.a_inner = F;
  d.a_rows = L_in;
  d.a_row_stride_bytes = static_cast<unsigned long long>(F) * 2;
  d.a_batch_stride_bytes = d.a_row_stride_bytes * L_in;
  d.g_ptr[0] = d8_hi;
  d.g_ptr[1] = d8_lo;
  d.g_inner = static_cast<unsigned long long>(W) * Cp;
  d.g_rows = L_in;
  d.g_row_stride_bytes = static_cast<unsigned long long>(Cp) * 2;
  d.g_batch_stride_bytes = static_cast<unsigned long long>(Lp) * Cp * 2;
  d.n_passes = precise ? 3 : 1;
  d.pa_bits = 0b010;
  d.pg_bits = 0b100;
  d.db_mask = 0b101;
  d.ws = ws;
  d.ws_bytes = ws_bytes;
  d.sc.mode = 2;
  d.sc.W = W;
  d.sc.Fw = Fw;
  d.sc.T = T;
  d.sc.H = H;
  d.sc.Cp = Cp;
  d.sc.o0 = dpw;
  d.sc.o2 = dcw;
  return wgrad_run(d, static_cast<cudaStream_t>(stream_));
}
