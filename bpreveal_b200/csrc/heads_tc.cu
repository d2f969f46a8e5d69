// K3 / K4 forward on tensor cores: every profile head and every counts head of the model in one
// tcgen05 GEMM (models.py:39-49 of the reference: Conv1D(valid, width W) per head, and
// GlobalAveragePooling1D -> Dense(1) per head).
//
//   logits[b,t,k]   = pb[k] + sum_{j<W} sum_c act[b,t+j,c] pw[j][c][k]          k < T (all heads)
//   logcounts[b,h]  = cb[h] + (1/L_in) sum_t sum_c act[b,t,c] cw[c][h]
//
// GEMM view: M = 128 consecutive positions, N = T + H padded to 16, K = W taps x F channels.  The
// counts heads ride along as H extra output columns whose weights are non-zero on tap 0 only
// (the mean over positions commutes with the dense layer), so the last tower activation is read
// exactly once.  One TMA box of 128 + W - 1 rows per tile and channel chunk feeds all W taps
// (descriptors start at row offsets 0..W-1 of the box: SWIZZLE_128B depends on the absolute
// shared-memory address only); weights stay resident in shared memory.
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-5
// epilogue (TMEM quadrant = warp % 4): bias, coalesced fp32 logits stores, warp-shuffle reduction
// of the counts columns into per-tile partial sums (added in tile order by a tiny second kernel:
// deterministic, no atomics).
#include "common.cuh"
#include "ptx.cuh"

#include <cstring>

namespace bp {

constexpr int kHdStages = 4;

struct HeadsKParams {
  int B, L_in, L_out, tiles_per_seq, W, KC, N, T, H;
  int n_aplanes, n_wplanes, npass, pa_bits, pw_bits;
  int sub_bytes, box_bytes, stages;
  int row_shift;  // the box of tile t0 starts at row t0 + row_shift (negative rows read as zero)
  uint32_t idesc;
  // 1: this launch handles one GROUP of input channels of a model whose weights do not fit shared
  // memory at once (F = 512 with 75-wide heads: 1.2 MB); logits are added to what the previous
  // group's launch left (the bias comes with group 0)
  int accumulate;
  const float* pb;
  const uint8_t* w_image;  // the prepared weights: the shared-memory image, copied in one piece
  float* logits;
  float* cpart;  // [B][tiles_per_seq][H] (of this channel group)
  unsigned* dbg;
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

template <uint32_t TMEM_COLS>
__global__ void __launch_bounds__(192, 1)
heads_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                    const __grid_constant__ CUtensorMap tmW, const HeadsKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int b_tile = p.N * 128;
  const int w_tiles = p.n_wplanes * p.W * p.KC;
  const int stage_bytes = p.KC * p.n_aplanes * p.sub_bytes;
  uint8_t* w_smem = smem;
  uint8_t* ring = w_smem + ((static_cast<size_t>(w_tiles) * b_tile + 1023) & ~static_cast<size_t>(1023));
  float* s_bias = reinterpret_cast<float*>(ring + static_cast<size_t>(p.stages) * stage_bytes);
  float* s_cpart = s_bias + 256;  // [4 warps][16]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_cpart + 64);
  uint64_t* full = bars;
  uint64_t* empty = full + kHdStages;
  uint64_t* acc_full = empty + kHdStages;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* w_full = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.B * p.tiles_per_seq;

  pdl_launch_dependents();
#ifdef BPB_ROLE_PROFILE
  const long long t_kernel_start = clock64();
  if (threadIdx.x < 32) prof_smem()[threadIdx.x] = 0;
#define HD_STAMP(slot) prof_smem()[slot] = static_cast<unsigned>(clock64() - t_kernel_start)
#else
#define HD_STAMP(slot)
#endif
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA0);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    s_bias[i] = (i < p.T && p.pb != nullptr) ? p.pb[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) HD_STAMP(20);

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(w_full, static_cast<uint32_t>(w_tiles) * b_tile);
      // bpb_prep_heads_fwd_weights writes the swizzled shared-memory image, so ONE bulk copy
      // brings every tap in (issuing a tensor load per tap cost ~150 cycles each on this thread,
      // at the head of the kernel's critical path)
      bulk_load_1d(w_smem, p.w_image, static_cast<uint32_t>(w_tiles) * b_tile, w_full);
      HD_STAMP(21);
      pdl_wait();  // the activations come from the previous kernel; the weights do not
      uint32_t stage = 0, phase = 0, ntile = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ntile) {
        const int b = tile / p.tiles_per_seq;
        const int t0 = (tile % p.tiles_per_seq) * 128;
        mbar_wait_wd(&empty[stage], phase ^ 1, p.dbg, 0x2100 + stage, ntile);
        mbar_expect_tx(&full[stage], static_cast<uint32_t>(p.KC * p.n_aplanes * p.box_bytes));
        uint8_t* dst = ring + static_cast<size_t>(stage) * stage_bytes;
        for (int c = 0; c < p.KC; ++c)
          for (int pl = 0; pl < p.n_aplanes; ++pl)
            tma_load_3d(pl ? &tmA1 : &tmA0, dst + static_cast<size_t>(c * p.n_aplanes + pl) * p.sub_bytes,
                        &full[stage], c * 64, t0 + p.row_shift, b);
        if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      mbar_wait_wd(w_full, 0, p.dbg, 0x2300);
      HD_STAMP(22);
      uint32_t stage = 0, phase = 0, as = 0, aphase = 0, ntile = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ntile) {
        mbar_wait_wd(&acc_empty[as], aphase ^ 1, p.dbg, 0x2400 + as, ntile);
        mbar_wait_wd(&full[stage], phase, p.dbg, 0x2500 + stage, ntile);
#ifdef BPB_ROLE_PROFILE
        if (ntile == 0) HD_STAMP(23);
#endif
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * p.N;
        const uint32_t base = smem_u32(ring + static_cast<size_t>(stage) * stage_bytes);
        uint32_t accum = 0;
        for (int pass = 0; pass < p.npass; ++pass) {
          const int ap = (p.pa_bits >> pass) & 1, wp = (p.pw_bits >> pass) & 1;
          // W x 4 MMAs per chunk issued by ONE thread: the descriptor words advance by constants
          // (A: one box row = 128 B per tap, W: one tile per tap), two adds per tap
          constexpr uint32_t DHI = umma_desc_hi(1024);
          for (int c = 0; c < p.KC; ++c) {
            uint32_t a_lo = umma_desc_lo(base + (c * p.n_aplanes + ap) * p.sub_bytes, 0);
            uint32_t b_lo = umma_desc_lo(smem_u32(w_smem) + ((wp * p.W) * p.KC + c) * b_tile, 0);
            const uint32_t b_step = (static_cast<uint32_t>(p.KC) * b_tile) >> 4;
            for (int j = 0; j < p.W; ++j) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_bf16_lohi(d_tmem, a_lo + 2u * k, DHI, b_lo + 2u * k, DHI, p.idesc, accum);
                accum = 1;
              }
              a_lo += 8u;
              b_lo += b_step;
            }
          }
        }
        umma_commit(&empty[stage]);
        umma_commit(&acc_full[as]);
#ifdef BPB_ROLE_PROFILE
        if (ntile == 0) HD_STAMP(24);
#endif
        if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1; }
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
      HD_STAMP(25);
    }
  } else {
    pdl_wait();  // global stores below
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;
    uint32_t as = 0, aphase = 0, ntile = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ntile) {
      const int b = tile / p.tiles_per_seq;
      const int mt = tile % p.tiles_per_seq;
      const int t = mt * 128 + row;
      mbar_wait_wd(&acc_full[as], aphase, p.dbg, 0x2800 + as, ntile);
#ifdef BPB_ROLE_PROFILE
      if (ntile == 0 && et == 0) HD_STAMP(26);
#endif
      tc_fence_after();
      for (int n0 = 0; n0 < p.N; n0 += 16) {
        uint32_t acc[16];
        tmem_ld16(tmem_base + as * p.N + n0 + (static_cast<uint32_t>(q * 32) << 16), acc);
        tmem_ld_wait();
        if (n0 < p.T && t < p.L_out) {
          float* dst = p.logits + (static_cast<size_t>(b) * p.L_out + t) * p.T;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (n0 + i < p.T)
              dst[n0 + i] = __uint_as_float(acc[i]) + (p.accumulate ? dst[n0 + i] : s_bias[n0 + i]);
        }
        // counts columns T .. T+H-1: deterministic sum over the 128 rows of the tile (rows beyond
        // L_in were zero-filled by TMA and add nothing); the column test is warp-uniform
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int h = n0 + i - p.T;
          if (h >= 0 && h < p.H) {
            float v = __uint_as_float(acc[i]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) s_cpart[q * 16 + h] = v;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
      named_bar_sync(1, 128);
      if (et < p.H && et < 16)
        p.cpart[(static_cast<size_t>(b) * p.tiles_per_seq + mt) * p.H + et] =
            (s_cpart[2 * 16 + et] + s_cpart[3 * 16 + et]) + (s_cpart[0 * 16 + et] + s_cpart[1 * 16 + et]);
      named_bar_sync(1, 128);
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
    atomicAdd(p.dbg + 1024 + 32 * blockIdx.x + threadIdx.x, v);
  }
#endif
}

// logcounts[b][h] = cb[h] + (sum over channel groups and tiles of cpart[g][b][tile][h]) / L_in
__global__ void heads_counts_finish_kernel(int B, int tiles, int H, int L_in, int groups,
                                           const float* __restrict__ cpart,
                                           const float* __restrict__ cb,
                                           float* __restrict__ logcounts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  const int b = i / H, h = i % H;
  float s = 0.f;
  for (int g = 0; g < groups; ++g)
    for (int tIdx = 0; tIdx < tiles; ++tIdx)
      s += cpart[((static_cast<size_t>(g) * B + b) * tiles + tIdx) * H + h];
  logcounts[i] = cb[h] + s / static_cast<float>(L_in);
}

// pw [W][Fw][T], cw [Fw][H] fp32 -> op: bf16 image of the kernel's shared-memory operand,
// [planes][W][KC] tiles of N rows x 64 channels
// (planes, taps, KCg chunks) per channel GROUP, groups back to back: one group's image is what one
// launch keeps resident (KCg = K / 64 when everything fits, i.e. a single group)
__global__ void __launch_bounds__(256)
prep_heads_fwd_kernel(int W, int Fw, int K, int N, int T, int H, int planes, int flip_input_conv,
                      int KCg, const float* __restrict__ pw, const float* __restrict__ cw,
                      __nv_bfloat16* __restrict__ op) {
  const long long n_el = static_cast<long long>(W) * N * K;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n_el) return;
  const int c = static_cast<int>(i % K);
  const int n = static_cast<int>((i / K) % N);
  const int j = static_cast<int>(i / (static_cast<long long>(K) * N));
  float v = 0.f;
  if (c < Fw) {
    if (flip_input_conv) {
      // transposed input convolution: pw is the Keras kernel [W][4][Fw]; tap j of this operand is
      // tap W-1-j of the forward filter, column n the input channel, row c the filter
      if (n < T) v = pw[(static_cast<long long>(W - 1 - j) * T + n) * Fw + c];
    } else if (n < T) {
      v = pw[(static_cast<long long>(j) * Fw + c) * T + n];
    } else if (n < T + H && j == 0) {
      v = cw[c * H + (n - T)];
    }
  }
  // destination: tile (plane, tap, 64-channel chunk) of N rows x 128 B, 16-byte chunks XOR-ed
  // with the row (the 128-byte swizzle of a tile whose base is 1024-aligned in shared memory)
  const int cc = c & 63;
  const long long in_tile = static_cast<long long>(n) * 64 + (((cc >> 3) ^ (n & 7)) << 3) + (cc & 7);
  const int grp = (c >> 6) / KCg, cg = (c >> 6) % KCg;
  float r = v;
  for (int pl = 0; pl < planes; ++pl) {
    const __nv_bfloat16 hbits = __float2bfloat16_rn(r);
    const long long tile = ((static_cast<long long>(grp) * planes + pl) * W + j) * KCg + cg;
    op[tile * (static_cast<long long>(N) * 64) + in_tile] = hbits;
    r -= __bfloat162float(hbits);
  }
}

static int heads_n(int T, int H) { return (T + H + 15) / 16 * 16; }

struct HeadsPlan {
  int N, KC, groups, box_rows, sub_bytes, stages;  // KC: chunks of ONE channel group
  size_t w_bytes, smem_bytes;
};

// The largest channel group (a divisor of F / 64 chunks) whose weights stay resident next to at
// least two operand stages; groups > 1 means one launch per group, accumulating into the logits.
static bool plan_heads(int W, int F, int T, int H, int planes, HeadsPlan* pl) {
  pl->N = heads_n(T, H);
  pl->box_rows = 128 + W - 1;
  if (pl->N > 256 || pl->box_rows > 256 || F <= 0 || F % 64 != 0) return false;
  pl->sub_bytes = (pl->box_rows * 128 + 1023) / 1024 * 1024;
  const size_t fixed = 1024 + 256 * 4 + 64 * 4 + (2 * kHdStages + 5) * 8 + 16;
  const int max_smem = max_smem_optin();
  if (max_smem <= 0) return false;
  const int KCall = F / 64;
  for (int groups = 1; groups <= KCall; ++groups) {
    if (KCall % groups != 0) continue;
    const int KC = KCall / groups;
    const size_t w_bytes = (static_cast<size_t>(planes) * W * KC * pl->N * 128 + 1023) / 1024 * 1024;
    const size_t stage = static_cast<size_t>(KC) * planes * pl->sub_bytes;
    if (fixed + w_bytes + 2 * stage > static_cast<size_t>(max_smem) &&
        !(groups == 1 && fixed + w_bytes + stage <= static_cast<size_t>(max_smem)))
      continue;
    int st = static_cast<int>((static_cast<size_t>(max_smem) - fixed - w_bytes) / stage);
    if (st > kHdStages) st = kHdStages;
    pl->KC = KC;
    pl->groups = groups;
    pl->w_bytes = w_bytes;
    pl->stages = st;
    pl->smem_bytes = fixed + w_bytes + st * stage;
    return true;
  }
  return false;
}

}  // namespace bp

extern "C" long long bpb_heads_fwd_wop_elems(int W, int F, int T, int H, int planes) {
  return static_cast<long long>(planes) * W * bp::heads_n(T, H) * F;
}

extern "C" int bpb_prep_heads_fwd_weights(int W, int Fw, int F, int T, int H, const float* pw,
                                          const float* cw, bpb_bf16* w_op, int planes,
                                          void* stream_) {
  using namespace bp;
  BP_REQUIRE(pw && cw && w_op && F % 64 == 0 && Fw <= F && (planes == 1 || planes == 2) && T >= 1 &&
                 H >= 1,
             "bpb_prep_heads_fwd_weights: bad arguments");
  const int N = heads_n(T, H);
  const long long n = static_cast<long long>(W) * N * F;
  HeadsPlan pl;
  BP_REQUIRE(plan_heads(W, F, T, H, planes, &pl),
             "bpb_prep_heads_fwd_weights: W=%d F=%d planes=%d does not fit shared memory", W, F, planes);
  prep_heads_fwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0,
                          static_cast<cudaStream_t>(stream_)>>>(
      W, Fw, F, N, T, H, planes, 0, pl.KC, pw, cw, reinterpret_cast<__nv_bfloat16*>(w_op));
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" long long bpb_input_dgrad_wop_elems(int W, int F, int planes) {
  return bpb_heads_fwd_wop_elems(W, F, 4, 0, planes);
}

extern "C" int bpb_input_conv_dgrad_tc_supported(int W, int F, int planes) {
  bp::HeadsPlan pl;
  return (F > 0 && F % 64 == 0 && bp::plan_heads(W, F, 4, 0, planes, &pl) && pl.groups == 1) ? 0 : 1;
}

extern "C" int bpb_prep_input_dgrad_weights(int W, int Fw, int F, const float* w, bpb_bf16* w_op,
                                            int planes, void* stream_) {
  using namespace bp;
  BP_REQUIRE(w && w_op && F % 64 == 0 && Fw <= F && (planes == 1 || planes == 2),
             "bpb_prep_input_dgrad_weights: bad arguments");
  const int N = heads_n(4, 0);
  const long long n = static_cast<long long>(W) * N * F;
  prep_heads_fwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0,
                          static_cast<cudaStream_t>(stream_)>>>(
      W, Fw, F, N, 4, 0, planes, 1, F / 64, w, nullptr, reinterpret_cast<__nv_bfloat16*>(w_op));
  BP_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// one [B][tiles][H] block of partial sums per channel group, for up to 16 groups (F <= 1024)
extern "C" long long bpb_heads_fwd_ws_bytes(int B, int L_in, int H) {
  return static_cast<long long>(B) * ((L_in + 127) / 128) * H * 4 * 16;
}

// Returns 0 if the tensor-core kernel can run this geometry (the weights of at least one
// 64-channel group + two boxes fit in shared memory, T + H <= 256, H <= 16, W <= 129); callers
// fall back to bpb_heads_fwd otherwise.
extern "C" int bpb_heads_fwd_tc_supported(int W, int F, int T, int H, int planes) {
  bp::HeadsPlan pl;
  return (F > 0 && F % 64 == 0 && F <= 1024 && H <= 16 && T <= 240 &&
          bp::plan_heads(W, F, T, H, planes, &pl)) ? 0 : 1;
}

// Channel groups bpb_heads_fwd_tc runs this geometry in (1 = everything resident; > 1 = one
// launch per group, and the per-tile counts partials are [groups][B][tiles][H]); 0 = unsupported.
extern "C" int bpb_heads_fwd_groups(int W, int F, int T, int H, int planes) {
  bp::HeadsPlan pl;
  if (bpb_heads_fwd_tc_supported(W, F, T, H, planes) != 0) return 0;
  bp::plan_heads(W, F, T, H, planes, &pl);
  return pl.groups;
}

// One launch of the window GEMM.  full = 0: the heads ('valid' correlation, L_out = L_in - W + 1,
// tiles cover L_in because the counts columns sum over every input row); full = 1: 'full'
// correlation of L_in + W - 1 output rows (the box starts W - 1 rows early; rows outside the
// sequence read as zero through TMA's out-of-bounds fill) -- the transposed input convolution.
static int heads_fwd_run(const char* who, int full, int B, int L_in, int W, int F, int T, int H,
                         const bpb_bf16* act_hi, const bpb_bf16* act_lo, const bpb_bf16* w_op,
                         const float* pb, const float* cb, float* logits, float* logcounts,
                         float* ws, long long ws_bytes, void* stream_) {
  using namespace bp;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int planes = act_lo ? 2 : 1;
  HeadsPlan pl;
  BP_REQUIRE(plan_heads(W, F, T, H, planes, &pl),
             "%s: W=%d F=%d planes=%d does not fit shared memory", who, W, F, planes);
  BP_REQUIRE(H == 0 || ws_bytes >= bpb_heads_fwd_ws_bytes(B, L_in, H), "%s: workspace too small", who);
  BP_REQUIRE(pl.groups == 1 || (!full && pl.groups <= 16), "%s: channel groups only for the heads", who);
  HeadsKParams kp{};
  kp.B = B;
  kp.L_in = L_in;
  kp.L_out = full ? L_in + W - 1 : L_in - W + 1;
  kp.row_shift = full ? -(W - 1) : 0;
  kp.tiles_per_seq = ((full ? kp.L_out : L_in) + 127) / 128;
  kp.W = W;
  kp.KC = pl.KC;
  kp.N = pl.N;
  kp.T = T;
  kp.H = H;
  kp.n_aplanes = planes;
  kp.n_wplanes = planes;
  kp.npass = planes == 2 ? 3 : 1;
  kp.pa_bits = 0b010;
  kp.pw_bits = 0b100;
  kp.sub_bytes = pl.sub_bytes;
  kp.box_bytes = pl.box_rows * 128;
  kp.stages = pl.stages;
  kp.idesc = umma_idesc_bf16(128, pl.N, 0, 0);
  kp.pb = pb;
  kp.logits = logits;
  kp.dbg = debug_buffer();
  const int total = B * kp.tiles_per_seq;
  int grid = sm_count();
  if (grid > total) grid = total;
  const uint32_t cols = 2u * pl.N;
  const int Fg = pl.KC * 64;  // channels of one group
  for (int g = 0; g < pl.groups; ++g) {
    kp.accumulate = g > 0 ? 1 : 0;
    kp.w_image = reinterpret_cast<const uint8_t*>(w_op) +
                 static_cast<size_t>(g) * planes * W * pl.KC * pl.N * 128;
    kp.cpart = ws ? ws + static_cast<size_t>(g) * B * kp.tiles_per_seq * H : nullptr;
    CUtensorMap maps[3];
    memset(maps, 0, sizeof(maps));
    // the group's channels of the activation: same rows and strides, inner extent Fg
    const unsigned long long row_b = static_cast<unsigned long long>(F) * 2, bat_b = row_b * L_in;
    if (!make_tmap_3d_strided(&maps[0], act_hi + static_cast<size_t>(g) * Fg, Fg, L_in, B, row_b, bat_b,
                              64, pl.box_rows))
      return 3;
    if (planes == 2 && !make_tmap_3d_strided(&maps[1], act_lo + static_cast<size_t>(g) * Fg, Fg, L_in, B,
                                             row_b, bat_b, 64, pl.box_rows))
      return 3;
    maps[2] = maps[0];  // (the weights arrive by one bulk copy of their prepared image)
#define LAUNCH_HD(C)                                                                            \
  do {                                                                                          \
    static bool configured = false;                                                             \
    if (!configured) {                                                                          \
      BP_CHECK_CUDA(cudaFuncSetAttribute(heads_fwd_tc_kernel<C>,                                \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                                         max_smem_optin()));                                    \
      configured = true;                                                                        \
    }                                                                                           \
    BP_CHECK_CUDA(launch_pdl(heads_fwd_tc_kernel<C>, dim3(grid), dim3(192), pl.smem_bytes,    \
                             stream, maps[0], maps[1], maps[2], kp));                           \
  } while (0)
    if (cols <= 32) LAUNCH_HD(32);
    else if (cols <= 64) LAUNCH_HD(64);
    else if (cols <= 128) LAUNCH_HD(128);
    else if (cols <= 256) LAUNCH_HD(256);
    else LAUNCH_HD(512);
#undef LAUNCH_HD
    BP_CHECK_CUDA(cudaGetLastError());
  }
  if (H > 0 && logcounts != nullptr) {
    heads_counts_finish_kernel<<<(B * H + 127) / 128, 128, 0, stream>>>(
        B, kp.tiles_per_seq, H, L_in, pl.groups, ws, cb, logcounts);
    BP_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

extern "C" int bpb_heads_fwd_tc(int B, int L_in, int W, int F, int T, int H,
                                const bpb_bf16* act_hi, const bpb_bf16* act_lo,
                                const bpb_bf16* w_op, const float* pb, const float* cb,
                                float* logits, float* logcounts, float* ws, long long ws_bytes,
                                void* stream_) {
  // logcounts == NULL: leave the per-tile partial sums in ws for bpb_loss_fwd_bwd_pack to finish
  BP_REQUIRE(act_hi && w_op && pb && cb && logits && ws, "bpb_heads_fwd_tc: NULL argument");
  BP_REQUIRE(F > 0 && F % 64 == 0 && H >= 1 && H <= 16 && T >= 1 && T <= 240 && L_in >= W,
             "bpb_heads_fwd_tc: bad geometry F=%d T=%d H=%d", F, T, H);
  return heads_fwd_run("bpb_heads_fwd_tc", 0, B, L_in, W, F, T, H, act_hi, act_lo, w_op, pb, cb, logits,
                       logcounts, ws, ws_bytes, stream_);
}

// Data gradient of the input convolution wrt the one-hot input (attribution only) as the same
// window GEMM: dx[b,t,c] = sum_j sum_n gz[b,t-j,n] w[j][c][n], a 'full' correlation with the
// tap-reversed, transposed filter (bpb_prep_input_dgrad_weights).
extern "C" int bpb_input_conv_dgrad_tc(int B, int L_in, int L_out, int W, int F,
                                       const bpb_bf16* gz_hi, const bpb_bf16* gz_lo,
                                       const bpb_bf16* w_op, float* dx, void* stream_) {
  BP_REQUIRE(gz_hi && w_op && dx, "bpb_input_conv_dgrad_tc: NULL argument");
  BP_REQUIRE(F > 0 && F % 64 == 0 && L_out == L_in - W + 1 && L_out >= 1,
             "bpb_input_conv_dgrad_tc: bad geometry");
  return heads_fwd_run("bpb_input_conv_dgrad_tc", 1, B, L_out, W, F, 4, 0, gz_hi, gz_lo, w_op, nullptr,
                       nullptr, dx, nullptr, nullptr, 0, stream_);
}
