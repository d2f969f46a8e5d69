// K2 / K6 / K8 entry point: width-3 dilated residual convolution of the BPNet tower, its data
// gradient and the DeepSHAP-rescale variant of that gradient (models.py:92-104,
// shap.py:612-639 of the reference), on the shared tcgen05 engine of conv_core.cuh.
#include "conv_core.cuh"

extern "C" int bpb_conv3_tc(const bpb_conv3_args* a, void* stream_) {
  using namespace bp;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BP_REQUIRE(a != nullptr, "bpb_conv3_tc: null args");
  BP_REQUIRE(a->F > 0 && a->F % 64 == 0, "bpb_conv3_tc: F=%d must be a positive multiple of 64",
             a->F);
  BP_REQUIRE(a->B > 0 && a->L_in > 0 && a->L_out > 0, "bpb_conv3_tc: bad geometry B=%d L_in=%d L_out=%d",
             a->B, a->L_in, a->L_out);
  BP_REQUIRE(a->in_hi && a->w_hi, "bpb_conv3_tc: in_hi / w_hi must not be NULL");
  BP_REQUIRE(a->resid_hi != nullptr && a->L_res > 0,
             "bpb_conv3_tc: the residual tensor is mandatory (models.py:97-102 always adds it)");
  BP_REQUIRE(a->mode == 0 || a->mode == 1, "bpb_conv3_tc: mode must be 0 or 1");
  const int n_planes = a->in_lo ? 2 : 1;
  if (n_planes == 2)
    BP_REQUIRE(a->w_lo != nullptr, "bpb_conv3_tc: precise path needs w_lo");
  const int has_out = a->out_hi ? 1 : 0, has_out2 = a->out2_hi ? 1 : 0;
  BP_REQUIRE(has_out || has_out2 || a->z_out || a->mask_out, "bpb_conv3_tc: no output requested");
  if (n_planes == 2) {
    BP_REQUIRE(!has_out || a->out_lo, "bpb_conv3_tc: precise path needs out_lo");
    BP_REQUIRE(!has_out2 || a->out2_lo, "bpb_conv3_tc: precise path needs out2_lo");
    BP_REQUIRE(a->resid_lo, "bpb_conv3_tc: precise path needs resid_lo");
    BP_REQUIRE(!a->mult_hi || a->mult_lo, "bpb_conv3_tc: precise path needs mult_lo");
    // the kernel addresses weight plane wp at row offset wp*3*F
    BP_REQUIRE(a->w_lo == a->w_hi + static_cast<size_t>(3) * a->F * a->F,
               "bpb_conv3_tc: w_lo must directly follow w_hi in memory");
  }
  int epi;
  if (a->mode == 0) {
    if (has_out2) {
      BP_REQUIRE(a->zx_in != nullptr && a->zx_div > 0, "bpb_conv3_tc: forward out2 needs zx_in");
      BP_REQUIRE(a->z_out == nullptr && a->mask_out == nullptr,
                 "bpb_conv3_tc: forward out2 cannot be combined with z_out / mask_out");
      epi = EPI_FWD_R;
    } else {
      epi = a->z_out ? EPI_FWD_Z : EPI_FWD;
    }
    BP_REQUIRE(a->mask_in == nullptr && a->mult_hi == nullptr,
               "bpb_conv3_tc: mask_in / mult are backward inputs");
  } else {
    BP_REQUIRE(a->z_out == nullptr && a->mask_out == nullptr && a->zx_in == nullptr,
               "bpb_conv3_tc: z_out / mask_out / zx_in are forward arguments");
    BP_REQUIRE(!(a->mask_in && a->mult_hi), "bpb_conv3_tc: give mask_in or mult, not both");
    epi = a->mult_hi ? EPI_BWD_MULT : EPI_BWD_MASK;
  }
  if (a->mask_a)
    BP_REQUIRE(a->mode == 1 && n_planes == 1 && !a->mult_hi && a->F == 64,
               "bpb_conv3_tc: mask_a needs mode 1, the bf16 path, no multiplier and F = 64");

  ConvDesc d{};
  d.who = "bpb_conv3_tc";
  d.B = a->B;
  d.L_out = a->L_out;
  d.F = a->F;
  d.epi = epi;
  d.planes = n_planes;
  d.a_ptr[0] = a->in_hi;
  d.a_ptr[1] = a->in_lo;
  d.n_aplanes = n_planes;
  d.a_inner = a->F;
  d.a_rows = a->L_in;
  d.a_row_stride_bytes = static_cast<unsigned long long>(a->F) * 2;
  d.a_batch_stride_bytes = static_cast<unsigned long long>(a->F) * 2 * a->L_in;
  d.n_taps = 3;
  d.KC = a->F / 64;
  d.last_nk = 4;
  for (int j = 0; j < 3; ++j) d.tap_off[j] = a->tap_off[j];
  d.w_ptr = a->w_hi;
  d.n_wplanes = n_planes;
  // hi*hi, lo*hi, hi*lo (the lo*lo term is below fp32 resolution)
  d.npass = n_planes == 2 ? 3 : 1;
  d.pa_bits = 0b010;
  d.pw_bits = 0b010000;
  d.bias = a->bias;
  d.bias_n = a->F;
  d.resid[0] = a->resid_hi;
  d.resid[1] = a->resid_lo;
  d.L_res = a->L_res;
  d.resid_off = a->resid_off;
  d.out[0] = a->out_hi;
  d.out[1] = a->out_lo;
  d.out2[0] = a->out2_hi;
  d.out2[1] = a->out2_lo;
  d.mask_in = a->mask_in;
  d.mask_out = a->mask_out;
  d.mask_a = a->mask_a;
  d.mult[0] = a->mult_hi;
  d.mult[1] = a->mult_lo;
  d.zx_in = a->zx_in;
  d.zx_div = a->zx_div;
  d.z_out = a->z_out;
  d.allow_halo = true;
  return conv_run<true>(d, stream);
}
