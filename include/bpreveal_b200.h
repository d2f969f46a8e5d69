/*
 * bpreveal_b200 -- C ABI of the B200-native BPNet hot path.
 *
 * The reference (mmtrebuchet/bpreveal) has NO native/FFI boundary on this path: the model is a
 * Keras graph (src/models.py:53-115, 226-387), the losses are TF ops (src/losses.py:15-63) and
 * attribution is a patched TF gradient registry (src/internal/shap.py:612-678).  The entry points
 * below are therefore the seam a maintainer would bind from Python (ctypes);
 * each one names the reference code whose arithmetic it replaces.  The only pre-existing native
 * signatures near the path are slide/slideChar (src/internal/libslide.c:44-91), mirrored by
 * bpb_slide_*.
 *
 * Conventions: every pointer is a DEVICE pointer unless named host_*; `stream` is a cudaStream_t
 * passed as void*; all functions return 0 on success, non-zero on failure with the message
 * available from bpb_last_error() (thread-local).  No function allocates device memory or
 * synchronises the device.  Activations are channels-last [B][L][F] bf16 ("planes"); the precise
 * path carries every activation as a (hi, lo) pair of bf16 planes whose sum is the value (fp32
 * accumulate everywhere); lo pointers are NULL on the bf16 path.  F must be a multiple of 64
 * (callers zero-pad narrower models).
 */
#ifndef BPREVEAL_B200_H
#define BPREVEAL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint16_t bpb_bf16; /* raw bfloat16 bits */

int bpb_version(void);
const char* bpb_last_error(void);
/* Fills SM count and compute capability of the current device. */
int bpb_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* sizeof() of the argument structs as compiled: 0 bpb_conv3_args, 1 bpb_wgrad3_args,
 * 2 bpb_input_conv_args, 3 bpb_heads_dgrad_args, 4 bpb_prep_table, 5 bpb_peer_args (-1 otherwise);
 * lets a binding check its layout. */
int bpb_abi_struct_size(int which);
/* Pipeline watchdog: kernels whose mbarrier waits time out record
 * records in a host-mapped buffer and trap.  Copies up to n (<= 2048) words: word 0 = number of
 * timed-out waits, words 8.. = records of 5 words {tag, block, thread, parity, payload}.
 * Returns the number of words copied (host_out is a HOST pointer). */
int bpb_debug_dump(unsigned* host_out, int n);

/* ------------------------------------------------------------------------------------------
 * K2/K6/K8: width-3 dilated convolution as a tcgen05 implicit GEMM (M = positions, N = F,
 * K = 3F), used for the forward of models.py:92-104 (conv + bias + ReLU + crop + residual add),
 * for its data gradient, and for the DeepSHAP-rescale backward (shap.py:612-639).
 *
 *   acc[b,t,n] = sum_{j<3} sum_c in[b, t + tap_off[j], c] * w[(j*F + n)*F + c]      (rows outside
 *                [0, L_in) read as zero)
 *   mode 0 (forward):  v = acc + bias[n];
 *                      if zx_in:   mult = rescale(zx_in[b / zx_div, t, n], v)  -> out2
 *                      if z_out:   z_out[b,t,n] = v (fp32)
 *                      if mask_out: bit n%64 of mask_out[(b*L_out + t)*(F/64) + n/64] = (v > 0)
 *                      out[b,t,n] = relu(v) + resid[b, t + resid_off, n]
 *   mode 1 (backward): v = acc + (0 <= t + resid_off < L_res ? resid[b, t + resid_off, n] : 0)
 *                      out[b,t,n] = v
 *                      out2[b,t,n] = v * (mask_in bit | mult_in[b,t,n] | 1)
 *                      with mask_a (bf16 path, F = 64, B * L_in even): `in` is read as
 *                      in[b,s,c] * (bit c of mask_a[b*L_in + s]) -- the ReLU mask of the layer
 *                      above is applied to the staged operand in shared memory, so the training
 *                      backward passes the unmasked gradient g and never stores g (.) mask
 * w is [n_wplanes][3][F][F] bf16 with the contraction index innermost.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int B, L_in, L_out, F;
  int tap_off[3];
  int mode;
  const bpb_bf16 *in_hi, *in_lo;     /* [B][L_in][F] */
  const bpb_bf16 *w_hi, *w_lo;       /* [3][F][F], k innermost */
  const float* bias;                 /* [F] or NULL */
  const bpb_bf16 *resid_hi, *resid_lo; /* [B][L_res][F] or NULL */
  int L_res, resid_off;
  bpb_bf16 *out_hi, *out_lo;         /* [B][L_out][F] or NULL */
  bpb_bf16 *out2_hi, *out2_lo;       /* [B][L_out][F] or NULL */
  uint64_t* mask_out;                /* [B][L_out][F/64] or NULL */
  const uint64_t* mask_in;           /* [B][L_out][F/64] or NULL */
  const bpb_bf16 *mult_hi, *mult_lo; /* [B][L_out][F] or NULL */
  const float* zx_in;                /* [B/zx_div][L_out][F] fp32 or NULL */
  int zx_div;
  float* z_out;                      /* [B][L_out][F] fp32 or NULL */
  const uint64_t* mask_a;            /* [B][L_in][F/64] or NULL (mode 1 only, see above) */
} bpb_conv3_args;
int bpb_conv3_tc(const bpb_conv3_args* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K7: weight / bias gradient of the dilated convolution (tcgen05, MN-major operands, split-K
 * over positions; partial sums go to a caller-provided workspace and are added in a fixed order,
 * so the result is deterministic):
 *   dw[j][c][n] = sum_{b,t} act[b, t + j*dil, c] * gz[b,t,n]      (padded Keras layout (3, F, F))
 *   db[n]       = sum_{b,t} gz[b,t,n]
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int B, L_in, L_out, F, dil;
  const bpb_bf16 *act_hi, *act_lo; /* [B][L_in][F] */
  const bpb_bf16 *gz_hi, *gz_lo;   /* [B][L_out][F] */
  float* dw;                       /* [3][F][F] fp32, overwritten */
  float* db;                       /* [F] fp32, overwritten, or NULL */
  float* ws;                       /* workspace of bpb_wgrad3_ws_bytes() bytes */
  long long ws_bytes;
  const uint64_t* gz_mask;         /* [B][L_out] or NULL.  bf16 path, F = 64, B * L_out even: gz_hi
                                      holds the UNMASKED gradient g and the operand is
                                      g[b,t,n] * (bit n of gz_mask[b*L_out + t]), masked in shared
                                      memory (see bpb_conv3_args.mask_a) */
} bpb_wgrad3_args;
long long bpb_wgrad3_ws_bytes(int B, int L_out, int F, int precise);
int bpb_wgrad3_tc(const bpb_wgrad3_args* a, void* stream);
/* The same for up to 12 layers of one network (same B and F) in ONE launch: the SMs are shared
 * between the layers in proportion to their position tiles, so the split-K partials, the kernel
 * prologue / tail and the reduction are paid once per step instead of once per layer.  The ws /
 * ws_bytes fields of the individual layers are ignored; every layer's dw / db is overwritten. */
long long bpb_wgrad3_multi_ws_bytes(int n_layers, int B, const int* host_L_out, int F);
int bpb_wgrad3_multi(const bpb_wgrad3_args* host_layers, int n_layers, float* ws,
                     long long ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * K1: input convolution on one-hot DNA (models.py:85-90) as a tcgen05 GEMM.
 *   z[b,t,f] = bias[f] + sum_{j<W} sum_c x[b,t+j,c] w[j][c][f];  out = relu(z)
 * The sequence is first expanded to x8 = bf16 [B][L_in][8] (channels 4..7 zero); row t of the
 * implicit im2col matrix is then the contiguous run x8[b, t, 0] .. x8[b, t+W-1, 7], which a TMA
 * tensor map with overlapping rows (row stride 16 bytes) loads directly.  The Keras kernel is
 * prepared once per weight update as w_op = bf16 [w_planes][F][KC*64] with k = j*8 + c
 * (KC = ceil(8 W / 64)); the planes are successive bf16 remainders of the fp32 weight, so with
 * w_planes = 3 the product is exact to fp32 (the one-hot operand is exact in bf16).  Output activations are written with row stride F >= Fw
 * (padding channels are written as zero).
 * Optional outputs: relu mask bits, fp32 z (attribution of the query), rescale multiplier
 * against zx_in (attribution of a shuffled reference).
 * ------------------------------------------------------------------------------------------ */
long long bpb_x8_elems(int B, int L); /* B*L*8 plus the spare tail that whole-strip reads need */
int bpb_onehot_expand(int B, int L, const uint8_t* x, bpb_bf16* x8, void* stream);
long long bpb_input_conv_wop_elems(int W, int F, int planes);
int bpb_prep_input_conv_weights(int W, int Fw, int F, const float* w, bpb_bf16* w_op, int planes,
                                void* stream);
typedef struct {
  int B, L_in, L_out, W, Fw, F;
  const bpb_bf16* x8;   /* [B][L_in][8] */
  const bpb_bf16* w_op; /* [w_planes][F][KC*64] */
  int w_planes;
  const float* bias;    /* [Fw] */
  bpb_bf16 *out_hi, *out_lo;
  uint64_t* mask_out;
  float* z_out;
  const float* zx_in;
  int zx_div;
  bpb_bf16 *mult_hi, *mult_lo;
} bpb_input_conv_args;
int bpb_input_conv_fwd(const bpb_input_conv_args* a, void* stream);
/* bpb_wgrad3_multi and bpb_input_conv_wgrad in ONE launch (bf16 path): the tower layers plus the
 * input convolution as one more job of the same split-K kernel (its window units are STRIP-staged
 * from x8) and, with d8 != NULL, the heads' weight gradients of bpb_heads_wgrad_tc as a further
 * job (transposed: its window units are STRIP-staged from d8 [B][L_last+W_out-1][8], so T + H <= 8;
 * act_last [B][L_last][F]).  _ws_bytes (W_out = 0: no heads job) returns -1 when the geometry
 * cannot share a launch (e.g. F > 64, where the tower needs several accumulator groups); callers
 * then use the separate entry points. */
long long bpb_wgrad_tower_input_ws_bytes(int n_layers, int B, const int* L_out, int F, int L_out0,
                                         int W, int L_last, int W_out);
int bpb_wgrad_tower_input(const bpb_wgrad3_args* layers, int n_layers, int L_in0, int L_out0, int W,
                          int Fw, const bpb_bf16* x8, const bpb_bf16* gz0_hi,
                          const uint64_t* gz0_mask /* NULL, or as bpb_wgrad3_args.gz_mask */,
                          float* dw1, float* db1, int L_last, int W_out, int T, int H,
                          const bpb_bf16* act_last, const bpb_bf16* d8, float* dpw, float* dcw,
                          float* ws, long long ws_bytes, void* stream);

/* dw[j][c][f] = sum_{b,t} x[b,t+j,c] gz[b,t,f]; db[f] = sum gz  (dw is [W][4][Fw], overwritten).
 * Runs on the K7 tensor-core kernel: the A operand is the implicit im2col view of x8. */
long long bpb_input_conv_wgrad_ws_bytes(int B, int L_out, int W, int F);
int bpb_input_conv_wgrad(int B, int L_in, int L_out, int W, int Fw, int F, const bpb_bf16* x8,
                         const bpb_bf16* gz_hi, const bpb_bf16* gz_lo, float* dw, float* db,
                         float* ws, long long ws_bytes, void* stream);
/* dx[b,s,c] = sum_j sum_f gz[b,s-j,f] w[j][c][f]  (fp32 out [B][L_in][4]); attribution only */
int bpb_input_conv_dgrad(int B, int L_in, int L_out, int W, int Fw, int F, const bpb_bf16* gz_hi,
                         const bpb_bf16* gz_lo, const float* w, float* dx, void* stream);

/* ------------------------------------------------------------------------------------------
 * K3/K4: output heads (models.py:20-50).  All heads are evaluated together: T = sum of num-tasks.
 *   logits[b,t,k] = pb[k] + sum_{j<W} sum_c act[b,t+j,c] pw[j][c][k]        (fp32 [B][L_out][T])
 *   gap[b,c]      = mean_t act[b,t,c]                                       (fp32 [B][F])
 *   logcounts[b,h] = cb[h] + sum_c gap[b,c] cw[c][h]                        (fp32 [B][H])
 * pw is the concatenation of the heads' Keras kernels along the task axis; cw likewise [Fw][H].
 * ------------------------------------------------------------------------------------------ */
int bpb_heads_fwd(int B, int L_in, int L_out, int W, int Fw, int F, int T, int H,
                  const bpb_bf16* act_hi, const bpb_bf16* act_lo, const float* pw,
                  const float* pb, const float* cw, const float* cb, float* logits, float* gap,
                  float* logcounts, void* stream);
/* Data gradient of the input convolution wrt the one-hot input (attribution only; TF gradient of
 * models.py:85-90) on the same window-GEMM kernel as bpb_heads_fwd_tc, as a 'full' correlation:
 *   dx[b,t,c] = sum_j sum_n gz[b,t-j,n] w[j][c][n]      dx fp32 [B][L_in][4], gz bf16 [B][L_out][F]
 * w_op (bpb_input_dgrad_wop_elems elements) is the tap-reversed, transposed filter laid out by
 * bpb_prep_input_dgrad_weights from the Keras kernel w [W][4][Fw].  _supported returns 0 when the
 * geometry fits shared memory; otherwise use bpb_input_conv_dgrad (CUDA cores, fp32 weights). */
long long bpb_input_dgrad_wop_elems(int W, int F, int planes);
int bpb_input_conv_dgrad_tc_supported(int W, int F, int planes);
int bpb_prep_input_dgrad_weights(int W, int Fw, int F, const float* w, bpb_bf16* w_op, int planes,
                                 void* stream);
int bpb_input_conv_dgrad_tc(int B, int L_in, int L_out, int W, int F, const bpb_bf16* gz_hi,
                            const bpb_bf16* gz_lo, const bpb_bf16* w_op, float* dx, void* stream);

/* The same forward as ONE tcgen05 GEMM (N = T + H padded to 16; the counts heads are extra output
 * columns with weights on tap 0 only; gap is not materialised).  w_op comes from
 * bpb_prep_heads_fwd_weights (an opaque bf16 image of the kernel's shared-memory operand:
 * [planes][W][F/64] tiles of N x 64, 128-byte swizzled); ws holds per-tile partial sums of the counts
 * columns (logcounts may be NULL: the sums are then left in ws for bpb_loss_fwd_bwd_pack).
 * bpb_heads_fwd_tc_supported returns 0 when the geometry fits shared memory; otherwise use
 * bpb_heads_fwd (CUDA cores). */
long long bpb_heads_fwd_wop_elems(int W, int F, int T, int H, int planes);
int bpb_prep_heads_fwd_weights(int W, int Fw, int F, int T, int H, const float* pw,
                               const float* cw, bpb_bf16* w_op, int planes, void* stream);
long long bpb_heads_fwd_ws_bytes(int B, int L_in, int H);
int bpb_heads_fwd_tc_supported(int W, int F, int T, int H, int planes);
/* Channel groups the geometry runs in: 1 = all weights resident; g > 1 = one launch per group of
 * F / g channels accumulating into the logits (wide models: F = 512 with 75-wide heads is 1.2 MB
 * of weights), the workspace then holds [g][B][tiles][H] partial sums; 0 = unsupported. */
int bpb_heads_fwd_groups(int W, int F, int T, int H, int planes);
int bpb_heads_fwd_tc(int B, int L_in, int W, int F, int T, int H, const bpb_bf16* act_hi,
                     const bpb_bf16* act_lo, const bpb_bf16* w_op, const float* pb,
                     const float* cb, float* logits, float* logcounts, float* ws,
                     long long ws_bytes, void* stream);
/* Parameter gradients of the heads on the K7 tensor-core kernel, from the packed loss gradient
 * d8 (see below) and the last tower activations act [B][L_in][F]:
 *   dpw[j][c][k] = sum_{b,t} act[b,t+j,c] dlogits[b,t,k]
 *   dcw[c][h]    = sum_b mean_t(act[b,t,c]) dlogcounts[b,h]
 * (overwritten; Keras layouts [W][Fw][T], [Fw][H]).  The bias gradients come from
 * bpb_heads_pack_grad. */
long long bpb_heads_wgrad_ws_bytes(int B, int L_in, int W, int F, int Cp);
int bpb_heads_wgrad_tc(int B, int L_in, int W, int Fw, int F, int T, int H, int Cp,
                       const bpb_bf16* act_hi, const bpb_bf16* act_lo, const bpb_bf16* d8_hi,
                       const bpb_bf16* d8_lo, float* dpw, float* dcw, float* ws,
                       long long ws_bytes, void* stream);

/* Data gradient of the heads on tensor cores.  bpb_heads_pack_grad writes
 *   d8[b, r, ch] = dlogits[b, r-(W-1), ch]  (ch < T, zero outside the profile)
 *                = dlogcounts[b, ch-T] / L_in  (T <= ch < T+H, every row)        bf16 planes
 * as [planes][B][L_in+W-1][Cp] (Cp a multiple of 8 >= T+H; each plane is followed by a spare tail,
 * bpb_heads_d8_elems() is the size of the whole buffer and the lo plane starts at its middle); bpb_prep_heads_dgrad_weights lays the
 * profile / counts kernels out as the matching GEMM operand
 * (with dpb != NULL it also writes the heads' bias gradients dpb[k] = sum dlogits[b,t,k] and
 * dcb[h] = sum_b dlogcounts[b,h] in fp32: their terms cancel almost exactly, so they are not
 * derived from the rounded d8); bpb_heads_dgrad_tc then computes
 *   g[b,s,c]  = sum_j sum_k dlogits[b,s-j,k] pw[j][c][k] + sum_h dlogcounts[b,h] cw[c][h] / L_in
 *   gz = g * mask_in / mult   (as bpb_conv3_tc mode 1)
 */
long long bpb_heads_d8_elems(int B, int L_in, int W, int Cp, int planes);
long long bpb_heads_dgrad_wop_elems(int W, int F, int Cp, int planes);
int bpb_heads_pack_grad(int B, int L_in, int L_out, int W, int T, int H, int Cp, int planes,
                        const float* dlogits, const float* dlogcounts, bpb_bf16* d8, float* dpb,
                        float* dcb, void* stream);
int bpb_prep_heads_dgrad_weights(int W, int Fw, int F, int T, int H, int Cp, const float* pw,
                                 const float* cw, bpb_bf16* w_op, int planes, void* stream);
typedef struct {
  int B, L_in, L_out, W, F, Cp;
  const bpb_bf16 *d8_hi, *d8_lo; /* [B][L_in+W-1][Cp] */
  const bpb_bf16* w_op;          /* [planes][F][KC*64], KC = ceil(W*Cp/64) */
  bpb_bf16 *g_hi, *g_lo;         /* [B][L_in][F] or NULL */
  bpb_bf16 *gz_hi, *gz_lo;       /* [B][L_in][F] or NULL */
  const uint64_t* mask_in;
  const bpb_bf16 *mult_hi, *mult_lo;
} bpb_heads_dgrad_args;
int bpb_heads_dgrad_tc(const bpb_heads_dgrad_args* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * K5: losses (losses.py:15-63, training.py:64-65).  Heads are described by task offsets
 * task_off[H+1] into the T axis.  Per head h:
 *   mnll_h = -(1/B) sum_b [lgamma(N_b+1) - sum_j lgamma(k_bj+1) + sum_j k_bj log_softmax(l_b)_j]
 *   mse_h  = (1/B) sum_b (logcounts[b,h] - log(sum_j k_bj))^2  (times lambda_h in the total)
 * Writes losses[0..H) = mnll, losses[H..2H) = lambda*mse (what Keras logs as the loss value of
 * reweightableMse) and, if dlogits != NULL, the gradient of
 *   sum_h profile_weight[h] * mnll_h + sum_h lambda[h] * mse_h
 * wrt logits and logcounts.  counts are fp32 [B][L][T]; logits may carry an additive frozen
 * bias term (combined model) which the caller adds beforehand.  task_off, profile_weight and
 * lambda are device arrays; logcounts_true [B][H] may be NULL (then log(sum_j k_bj) is used, which
 * is what generators.py:139-140 feeds).
 * ------------------------------------------------------------------------------------------ */
int bpb_loss_fwd_bwd(int B, int L, int T, int H, const int* task_off, const float* logits,
                     const float* logcounts, const float* counts, const float* logcounts_true,
                     const float* profile_weight, const float* lambda, float* losses,
                     float* dlogits, float* dlogcounts, void* stream);

/* K5 and bpb_heads_pack_grad in one launch (the training step's path): same losses and gradients,
 * the gradient written straight into the packed tensor d8 (layout of bpb_heads_pack_grad,
 * bpb_heads_d8_elems() elements; the caller zero-fills d8 ONCE -- only its T+H live channels are
 * written here), the heads' bias gradients dpb [T], dcb [H] (nullable) and the loss sums reduced
 * over the batch in sequence order by the last block to finish (deterministic; no atomics on
 * floats).  dlogits may be NULL.  ws: bpb_loss_pack_ws_elems() floats, zero-filled once by the
 * caller and then owned by these calls (one buffer per stream).  host_task_off is the HOST copy of
 * task_off (heads of 1..8 tasks).  counts_partials (nullable): the workspace bpb_heads_fwd_tc was
 * given when it was called with logcounts == NULL ([B][tiles][H], tiles = ceil(L_in / 128)); the
 * kernel then finishes logcounts[b,h] = cb[h] + sum / L_in itself -- in the order of the finishing
 * kernel it replaces -- and stores them into `logcounts`, which is an OUTPUT in that mode. */
long long bpb_loss_pack_ws_elems(int B, int T, int H);
int bpb_loss_fwd_bwd_pack(int B, int L, int T, int H, const int* task_off, const float* logits,
                          const float* logcounts, const float* counts, const float* logcounts_true,
                          const float* profile_weight, const float* lambda, float* losses,
                          float* dlogits, float* dlogcounts, int L_in, int W, int Cp, int planes,
                          bpb_bf16* d8, float* dpb, float* dcb, float* ws, const int* host_task_off,
                          const float* counts_partials, int tiles, const float* cb, void* stream);

/* ------------------------------------------------------------------------------------------
 * K9: Keras-semantics Adam over one flat fp32 bucket:
 *   m += (g-m)(1-b1); v += (g^2-v)(1-b2); p -= lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps)
 * `step_and_lr` is a device buffer {float step, float lr} so that the update can live inside a
 * CUDA graph; grad_scale multiplies g first (1/world_size after an all-reduce sum).
 * ------------------------------------------------------------------------------------------ */
int bpb_adam_step(long long n, float* param, const float* grad, float* m, float* v,
                  const float* step_and_lr, float beta1, float beta2, float eps, float grad_scale,
                  void* stream);

/* K9 and the four bpb_prep_* refreshes in one launch (the training step's path).  The table says
 * where the weight tensors live in the flat bucket (element offsets; Keras layouts, dil_w padded
 * to [layers][3][F][F]) and where their bf16 operand layouts go; every updated weight is written
 * to its operand positions (bit-equal to the bpb_prep_* kernels; padding entries are not touched
 * and must be zero from allocation).  state = {float step, float lr, uint32 ticket (zero)}: the
 * update uses step + 1 and the kernel stores step + 1 back (no separate increment launch). */
typedef struct {
  long long off_dil_w, off_conv1_w, off_prof_w, off_cnt_w;
  int n_layers, F, Fw, W_in, W_out, T, H, Cp;
  int planes;    /* bf16 remainder planes of the tower / heads operands (1 | 2) */
  int in_planes; /* ... of the input convolution operand (bpb_prep_input_conv_weights) */
  int N_heads;   /* padded column count of the heads-forward image (16 * ceil((T+H)/16)) */
  bpb_bf16 *w_fwd, *w_bwd; /* bpb_prep_conv3_weights outputs */
  bpb_bf16* w_in_op;       /* bpb_prep_input_conv_weights output */
  bpb_bf16* w_hd_op;       /* bpb_prep_heads_dgrad_weights output */
  bpb_bf16* w_hf_op;       /* bpb_prep_heads_fwd_weights output, or NULL */
  int hf_kcg;              /* 64-channel chunks per channel group of that image:
                              (F / 64) / bpb_heads_fwd_groups(...) (0 = F / 64, one group) */
} bpb_prep_table;
int bpb_adam_prep_step(long long n, float* param, const float* grad, float* m, float* v,
                       float* state, float beta1, float beta2, float eps, float grad_scale,
                       const bpb_prep_table* table, void* stream);

/* Data-parallel training over NVLink peer memory (one process per GPU, one node): the gradient
 * all-reduce is fused into the optimiser kernel.  Each rank allocates its gradient bucket and a
 * 128-byte flag block with bpb_peer_alloc (cudaMalloc + CUDA IPC handle, zero-filled), the ranks
 * exchange the 64-byte handles by any means and map each other's memory with bpb_peer_open.
 * bpb_adam_prep_step_peer = bpb_adam_prep_step with grad = the SUM over ranks of grads[r], read
 * straight out of peer memory in rank order (bit-identical on every rank), scaled by 1/world;
 * it waits until every peer has reached the same call -- bounded in wall time by flag word 17
 * (seconds, 0 = 30): a thread whose wait expires stores 1 + the missing rank into flag word 18 of
 * this rank's block and carries on, so the host can detect it and recover (no trap, no sticky
 * fault).  Flag block words: 0..7 ready[r], 8..15 done[r], 16 sequence, 17 bound, 18 error.  bpb_peer_wait
 * must be enqueued before anything writes this rank's bucket again (start of the next step): it
 * waits until every peer has finished reading it.  grads[r] / flags[r] are THIS process's
 * mappings of rank r's bucket / flag block (r == rank: its own allocation). */
typedef struct {
  int world, rank;
  const float* grads[8];
  unsigned* flags[8];
} bpb_peer_args;
int bpb_peer_alloc(long long bytes, void** ptr, unsigned char* handle64);
int bpb_peer_open(const unsigned char* handle64, void** ptr);
int bpb_peer_close(void* ptr);
int bpb_peer_free(void* ptr);
int bpb_peer_wait(const bpb_peer_args* peers, void* stream);
int bpb_adam_prep_step_peer(long long n, float* param, float* m, float* v, float* state, float beta1,
                            float beta2, float eps, const bpb_prep_table* table,
                            const bpb_peer_args* peers, void* stream);

/* fp32 Keras conv kernels [layers][3][Fw][Fw] -> padded bf16 operand layouts of bpb_conv3_tc:
 *   fwd[l][p][j][n][c] = w[l][j][c][n]   and   bwd[l][p][j][n][c] = w[l][j][n][c]   (each
 *   [layers][planes][3][F][F], zero padded; planes = 2 (hi then lo) when precise, else 1) */
int bpb_prep_conv3_weights(int n_layers, int Fw, int F, const float* w, bpb_bf16* fwd,
                           bpb_bf16* bwd, int precise, void* stream);

/* Prediction-side helpers (device pointers).
 * bpb_onehot_encode: ASCII bases seq[n] -> onehot [n][4] uint8, columns A C G T, either case,
 *   anything else all-zero (utils.py:420-486 with allowN); *n_other (device, nullable) receives the
 *   number of bytes that were not ACGT, so the caller can enforce allowN = False.
 * bpb_logits_to_profile: profile[b,:,tasks of head h] = softmax over the L x T_h block of the
 *   head's logits times exp(logcounts[b,h]) (utils.py:523-572, for a whole batch). */
int bpb_onehot_encode(long long n, const unsigned char* seq, unsigned char* onehot,
                      unsigned long long* n_other, void* stream);
int bpb_logits_to_profile(int B, int L, int T, int H, const int* task_off, const float* logits,
                          const float* logcounts, float* profile, void* stream);

/* ------------------------------------------------------------------------------------------
 * Epoch jitter / shuffle gather (generators.py:129-138, libslide.c:44-91):
 *   dst[row_idx[r]] = src[r, col_idx[r] : col_idx[r] + dst_cols]
 * ------------------------------------------------------------------------------------------ */
int bpb_slide_u8(const uint8_t* src, uint8_t* dst, int rows, int src_cols, int dst_cols,
                 int depth, const int* row_idx, const int* col_idx, void* stream);
int bpb_slide_f32(const float* src, float* dst, int rows, int src_cols, int dst_cols, int depth,
                  const int* row_idx, const int* col_idx, void* stream);
/* logcounts targets: out[r] = log(sum over (cols, depth) of vals[r]) (generators.py:139-140) */
int bpb_log_row_sums(const float* vals, float* out, int rows, int n_per_row, void* stream);

/* Elementwise helpers used by the transformation / combined models and attribution. */
/* y = sum over terms of  m1 * act(m2*u + b2) + b1  (models.py:118-175); act: 0 linear (m2,b2
 * unused), 1 sigmoid, 2 relu; coeffs is [n_terms][4] = {m1, b1, m2, b2} on the device. */
int bpb_transform_fwd(long long n, const float* u, float* y, int n_terms, const int* act,
                      const float* coeffs, void* stream);
/* dcoeffs[t][4] += sum_i dy[i] * d y_i / d{m1,b1,m2,b2} (trainTransformationModel) */
int bpb_transform_bwd(long long n, const float* u, const float* dy, int n_terms, const int* act,
                      const float* coeffs, float* dcoeffs, void* stream);
int bpb_add_f32(long long n, const float* a, const float* b, float* out, void* stream);
int bpb_mul_f32(long long n, const float* a, const float* b, float* out, void* stream);
/* layers.py:52-75 CountsLogSumExp in fp64 where use_bias[h] != 0, identity on the residual
 * otherwise (models.py:366-380); dresid_scale (optional) = d out / d resid_lc. */
int bpb_counts_lse(int B, int H, const int* use_bias, const float* bias_lc, const float* resid_lc,
                   float* out, float* dresid_scale, void* stream);

/* The combined model's output stage in one launch (models.py:349-380): the frozen bias model's
 * raw outputs [B][L][T] / [B][H] go through the per-head transformation of models.py:118-175
 * (coeffs [H][n_pterms + n_cterms][4], a head's profile terms first; 0 terms = a plain solo bias
 * model), out_logits = transformed bias logits + resid_logits, out_lc = CountsLogSumExp (fp64)
 * of the transformed bias log-counts and resid_lc where use_bias[h], else resid_lc;
 * dresid_scale (optional) = d out_lc / d resid_lc; bias_logits_t / bias_lc_t (optional) receive
 * the transformed bias outputs.  Replaces 2 H bpb_transform_fwd + bpb_add_f32 + bpb_counts_lse. */
int bpb_combine_bias(int B, int L, int T, int H, const int* task_off, const float* bias_logits,
                     const float* bias_lc, int n_pterms, const int* p_act, int n_cterms,
                     const int* c_act, const float* coeffs, const int* use_bias,
                     const float* resid_logits, const float* resid_lc, float* out_logits,
                     float* out_lc, float* dresid_scale, float* bias_logits_t, float* bias_lc_t,
                     void* stream);

/* DeepSHAP through the output stage of a transformation / combined model (shap.py:612-639,
 * 782-793 applied to models.py:118-175, 349-380): rescale rule on Sigmoid / Relu / Exp / Log,
 * ordinary gradients on the linear pieces.  R = Q * n (query, reference) pairs, pair i belongs to
 * query i / n.  *_x are the QUERY outputs ([Q][L][T] / [Q][H]), *_r the reference outputs
 * ([R][L][T] / [R][H]) of the bias network (raw, before the transformation) and of the residual
 * network; dlogits_c / dlc_c the metric's gradient wrt the COMBINED outputs of every pair.
 * Outputs (any may be NULL): the seeds of the two rescale-rule backward passes.  resid_lc_* NULL =
 * a transformation model alone. */
int bpb_combined_shap_seed(int R, int n, int L, int T, int H, const int* task_off,
                           const float* bias_logits_x, const float* bias_logits_r,
                           const float* bias_lc_x, const float* bias_lc_r,
                           const float* resid_lc_x, const float* resid_lc_r, int n_pterms,
                           const int* p_act, int n_cterms, const int* c_act, const float* coeffs,
                           const int* use_bias, const float* dlogits_c, const float* dlc_c,
                           float* dlogits_resid, float* dlc_resid, float* dlogits_bias,
                           float* dlc_bias, void* stream);

/* k-let preserving shuffles of a one-hot sequence (HOST pointers, host code -- the reference's
 * internal/libushuffle.c:59-344 runs on the host too): the DeepSHAP references for kmer-size > 1
 * (ushuffle.shuffleOHE, interpretUtils.py:1223-1225).  host_in [length][alphabet] (alphabet <= 8,
 * any non-zero byte is hot), host_out [numShuffles][length][alphabet] 0/1 bytes.  Draws from a
 * restatement of glibc's random() seeded with `seed` at the start of the call, in the reference's
 * order: the output is bit-identical to the reference's for the same seed. */
int bpb_ushuffle_ohe(const unsigned char* host_in, unsigned char* host_out, int alphabet, int length,
                     int kmerSize, int numShuffles, unsigned seed);

/* ------------------------------------------------------------------------------------------
 * DeepSHAP plumbing (interpretUtils.py:1216-1312, shap.py:26-30).
 * bpb_shuffle_rows: refs[q*n + r, t, :] = x[q, perm[r*L + t], :]  (x u8 [Q][L][4], perm int32 [n][L])
 * bpb_shap_combine: phi[q,t,c] = mean_r mult[q*n+r,t,c] * (x[q,t,c] - refs[q*n+r,t,c]), or the
 *   hypothetical projection phi[q,t,i] = mean_r sum_c (1[c==i] - refs[..,c]) * mult[..,c].
 * ------------------------------------------------------------------------------------------ */
int bpb_shuffle_rows(int Q, int n, int L, const uint8_t* x, const int* perm, uint8_t* refs,
                     void* stream);
/* Seeds of the interpretFlat profile metric (interpretFlat.py:187-237) under the DeepSHAP
 * 2-input rule (shap.py:642-678): for pair (q, r) of query logits_x [Q][L][T] and reference
 * logits_r [Q*n][L][T], over the columns with sel[k] != 0,
 *   dlogits[r] = m - mean(m),  m = (softmax(l~_x) + softmax(l~_r)) / 2  (0 where |l~_x - l~_r| < 1e-7)
 * and the metric values vx [Q], vr [Q*n] (either may be NULL). */
int bpb_flat_profile_seed(int Q, int n, int L, int T, const int* sel, const float* logits_x,
                          const float* logits_r, float* dlogits, float* vx, float* vr,
                          void* stream);
int bpb_shap_combine(int Q, int n, int L, const uint8_t* x, const uint8_t* refs,
                     const float* mult, int hypothetical, float* phi, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BPREVEAL_B200_H */
