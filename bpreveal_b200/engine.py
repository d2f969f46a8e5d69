"""Device-side execution of a BPNet ("solo") network: forward, training step, DeepSHAP backward.

This is the B200 replacement for what TensorFlow executes for the Keras graph built by the
reference's ``models.soloModel`` (src/models.py:53-115): parameters live in one flat fp32 bucket
(one Adam launch, one NCCL all-reduce), activations are channels-last bf16 planes that stay
resident in HBM, every layer is one native launch (see ``include/bpreveal_b200.h``) and a whole
step is captured in a CUDA graph.  All arithmetic is done by the CUDA kernels in ``csrc/``.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np
import os

import torch

from . import ops


@dataclasses.dataclass
class NetConfig:
    inputLength: int
    outputLength: int
    numFilters: int
    numLayers: int
    inputFilterWidth: int
    outputFilterWidth: int
    headTasks: list  # num-tasks per head
    precise: bool = False  # hi/lo bf16 planes, ~fp32 accuracy

    @property
    def Fp(self):
        return ((self.numFilters + 63) // 64) * 64

    @property
    def T(self):
        return int(sum(self.headTasks))

    @property
    def H(self):
        return len(self.headTasks)

    def lengths(self):
        """Activation lengths: [after conv1, after dil0, ...] (models.py:87-104)."""
        L = [self.inputLength - self.inputFilterWidth + 1]
        for i in range(self.numLayers):
            L.append(L[-1] - 2 * 2 ** (i + 1))
        return L

    def validate(self):
        L = self.lengths()
        lout = L[-1] - self.outputFilterWidth + 1
        if lout != self.outputLength:
            raise ValueError(
                f"Inconsistent network: input {self.inputLength} gives output {lout}, "
                f"not {self.outputLength}")


class ParamBucket:
    """Flat fp32 parameter bucket with named views in (padded) Keras layouts."""

    def __init__(self, cfg: NetConfig, device):
        self.cfg = cfg
        Fw, Fp, L = cfg.numFilters, cfg.Fp, cfg.numLayers
        W, Wo, T, H = cfg.inputFilterWidth, cfg.outputFilterWidth, cfg.T, cfg.H
        self.spec = [
            ("dil_w", (L, 3, Fp, Fp)), ("dil_b", (L, Fp)),
            ("conv1_w", (W, 4, Fw)), ("conv1_b", (Fw,)),
            ("prof_w", (Wo, Fw, T)), ("prof_b", (T,)),
            ("cnt_w", (Fw, H)), ("cnt_b", (H,)),
        ]
        self.offsets = {}
        off = 0
        for name, shape in self.spec:
            n = int(np.prod(shape))
            self.offsets[name] = (off, n, shape)
            off += (n + 3) // 4 * 4  # keep every view 16-byte aligned
        self.numel = off
        self.flat = torch.zeros((off,), dtype=torch.float32, device=device)

    def view(self, flat, name):
        off, n, shape = self.offsets[name]
        return flat[off:off + n].view(shape)

    def views(self, flat=None):
        flat = self.flat if flat is None else flat
        return {name: self.view(flat, name) for name, _ in self.spec}


class SoloNet:
    """One BPNet tower + heads on one GPU."""

    def __init__(self, cfg: NetConfig, device=None):
        cfg.validate()
        self.cfg = cfg
        self.device = torch.device(device if device is not None else "cuda")
        self.params = ParamBucket(cfg, self.device)
        self.p = self.params.views()
        self.grads_flat = None
        self.adam_m = None
        self.adam_v = None
        # {step, lr, ticket word of bpb_adam_prep_step, pad}
        self.step_and_lr = torch.zeros((4,), dtype=torch.float32, device=self.device)
        self.step = 0
        planes = 2 if cfg.precise else 1
        Fp, L = cfg.Fp, cfg.numLayers
        self.w_fwd = torch.zeros((max(L, 1), planes, 3, Fp, Fp), dtype=torch.bfloat16,
                                 device=self.device)
        self.w_bwd = torch.zeros_like(self.w_fwd)
        # tcgen05 operands of the input convolution and of the heads' data gradient
        self.cp = ops.heads_cp(cfg.T, cfg.H)
        # K1 weights are prepared as 2 (bf16 path) / 3 (precise path) bf16 remainder planes; three
        # are exact to fp32 (the one-hot operand is exact), so the first layer's pre-activations
        # and ReLU mask are fp32-accurate.  PREDICTION on the bf16 path multiplies only the first
        # plane, like every other layer: the second was measured to change no output error of the
        # C1 parity run (logits 1.59e-2 against 1.65e-2, both at the bf16 storage floor of
        # 1.45e-2) while doubling the kernel's MMAs -- 18.6 -> 14.5 us per launch, prediction
        # 505k -> 567k windows/s.  Training and attribution keep both planes: with one, ReLU
        # flips in the first layer move the conv1 weight gradient of small networks by 13 % of
        # its largest entry, and PISA's efficiency identity (a difference of nearly equal sums)
        # from 4.7e-2 to 1.4e-1.  BPB_W1_EXTRA=1 uses both planes everywhere.
        self.w_in_planes = planes + 1
        self.w_in_fast = self.w_in_planes if (cfg.precise or
                                              os.environ.get("BPB_W1_EXTRA", "0") == "1") else 1
        self.w_in_op = ops.input_conv_wop(cfg.inputFilterWidth, Fp, self.w_in_planes, self.device)
        self.w_hd_op = ops.heads_dgrad_wop(cfg.outputFilterWidth, Fp, self.cp, planes, self.device)
        # transposed input convolution (attribution only), refreshed where it is used
        self.w_id_op = ops.input_dgrad_wop(cfg.inputFilterWidth, Fp, planes, self.device)
        # heads forward as one tcgen05 GEMM when its weights fit shared memory (else CUDA cores)
        self.heads_tc = ops.heads_fwd_tc_supported(cfg.outputFilterWidth, Fp, cfg.T, cfg.H, planes)
        self.heads_groups = 1
        if self.heads_tc:
            self.w_hf_op = ops.heads_fwd_wop(cfg.outputFilterWidth, Fp, cfg.T, cfg.H, planes,
                                             self.device)
            # > 1: wide models whose head weights do not fit shared memory at once run one launch
            # per channel group (the per-tile counts partials are then per group as well)
            self.heads_groups = ops.heads_fwd_groups(cfg.outputFilterWidth, Fp, cfg.T, cfg.H, planes)
        self.host_task_off = [int(v) for v in np.concatenate([[0], np.cumsum(cfg.headTasks)])]
        self.task_off = torch.tensor(self.host_task_off, dtype=torch.int32, device=self.device)
        self._ws = {}
        self._graphs = {}

    # ------------------------------------------------------------------ parameters
    def init_random(self, seed=1234, scale=1.0, biasScale=0.0):
        """Glorot-uniform kernels, zero biases (Keras defaults)."""
        cfg = self.cfg
        rng = np.random.default_rng(seed)

        def glorot(shape, fi, fo):
            lim = math.sqrt(6.0 / (fi + fo))
            return rng.uniform(-lim, lim, size=shape).astype(np.float32) * scale

        Fn = cfg.numFilters
        sd = {"conv1_w": glorot((cfg.inputFilterWidth, 4, Fn), cfg.inputFilterWidth * 4,
                                cfg.inputFilterWidth * Fn),
              "conv1_b": (rng.standard_normal(Fn) * biasScale).astype(np.float32)}
        for i in range(cfg.numLayers):
            sd[f"dil{i}_w"] = glorot((3, Fn, Fn), 3 * Fn, 3 * Fn)
            sd[f"dil{i}_b"] = (rng.standard_normal(Fn) * biasScale).astype(np.float32)
        for h, T in enumerate(cfg.headTasks):
            sd[f"prof{h}_w"] = glorot((cfg.outputFilterWidth, Fn, T), cfg.outputFilterWidth * Fn,
                                      cfg.outputFilterWidth * T)
            sd[f"prof{h}_b"] = (rng.standard_normal(T) * biasScale).astype(np.float32)
            sd[f"cnt{h}_w"] = glorot((Fn, 1), Fn, 1)
            sd[f"cnt{h}_b"] = (rng.standard_normal(1) * biasScale).astype(np.float32)
        self.load_state(sd)
        return sd

    def load_state(self, sd):
        """Load a dict of numpy arrays in Keras layouts: conv1_w (W,4,F), dil{i}_w (3,F,F),
        prof{h}_w (Wo,F,T_h), cnt{h}_w (F,1) and the matching *_b biases."""
        cfg = self.cfg
        Fn = cfg.numFilters
        p = self.p
        dev = self.device

        def t(a):
            return torch.as_tensor(np.asarray(a, dtype=np.float32), device=dev)

        self._rf_twin = None  # cached receptive-field twin of interpret.pisa_batch is stale now
        self.params.flat.zero_()
        p["conv1_w"].copy_(t(sd["conv1_w"]))
        p["conv1_b"].copy_(t(sd["conv1_b"]))
        for i in range(cfg.numLayers):
            p["dil_w"][i, :, :Fn, :Fn].copy_(t(sd[f"dil{i}_w"]))
            p["dil_b"][i, :Fn].copy_(t(sd[f"dil{i}_b"]))
        k = 0
        for h, T in enumerate(cfg.headTasks):
            p["prof_w"][:, :, k:k + T].copy_(t(sd[f"prof{h}_w"]))
            p["prof_b"][k:k + T].copy_(t(sd[f"prof{h}_b"]))
            p["cnt_w"][:, h:h + 1].copy_(t(sd[f"cnt{h}_w"]))
            p["cnt_b"][h:h + 1].copy_(t(sd[f"cnt{h}_b"]))
            k += T
        self.refresh_weights()

    def state(self, flat=None):
        """Inverse of ``load_state`` (numpy, unpadded Keras layouts)."""
        cfg = self.cfg
        Fn = cfg.numFilters
        p = self.params.views(flat)
        sd = {"conv1_w": p["conv1_w"].cpu().numpy().copy(),
              "conv1_b": p["conv1_b"].cpu().numpy().copy()}
        for i in range(cfg.numLayers):
            sd[f"dil{i}_w"] = p["dil_w"][i, :, :Fn, :Fn].cpu().numpy().copy()
            sd[f"dil{i}_b"] = p["dil_b"][i, :Fn].cpu().numpy().copy()
        k = 0
        for h, T in enumerate(cfg.headTasks):
            sd[f"prof{h}_w"] = p["prof_w"][:, :, k:k + T].cpu().numpy().copy()
            sd[f"prof{h}_b"] = p["prof_b"][k:k + T].cpu().numpy().copy()
            sd[f"cnt{h}_w"] = p["cnt_w"][:, h:h + 1].cpu().numpy().copy()
            sd[f"cnt{h}_b"] = p["cnt_b"][h:h + 1].cpu().numpy().copy()
            k += T
        return sd

    def refresh_weights(self):
        """fp32 master weights -> bf16 tcgen05 operand layouts (after load / after Adam)."""
        cfg = self.cfg
        planes = 2 if cfg.precise else 1
        if cfg.numLayers > 0:
            ops.prep_conv3_weights(self.p["dil_w"], cfg.Fp, self.w_fwd, self.w_bwd, cfg.precise)
        ops.prep_input_conv_weights(self.p["conv1_w"], cfg.Fp, self.w_in_op, self.w_in_planes)
        ops.prep_heads_dgrad_weights(self.p["prof_w"], self.p["cnt_w"], cfg.Fp, self.cp,
                                     self.w_hd_op, planes)
        if self.heads_tc:
            ops.prep_heads_fwd_weights(self.p["prof_w"], self.p["cnt_w"], cfg.Fp, self.w_hf_op,
                                       planes)

    # ------------------------------------------------------------------ workspaces
    def _workspace(self, B, kind):
        key = (B, kind)
        if key in self._ws:
            return self._ws[key]
        cfg, dev = self.cfg, self.device
        Fp, prec = cfg.Fp, cfg.precise
        Ls = cfg.lengths()
        ws = {"B": B}
        ws["x"] = torch.zeros((B, cfg.inputLength, 4), dtype=torch.uint8, device=dev)
        ws["x8"] = ops.new_x8(B, cfg.inputLength, dev)
        ws["logits"] = torch.empty((B, cfg.outputLength, cfg.T), dtype=torch.float32, device=dev)
        ws["gap"] = torch.empty((B, Fp), dtype=torch.float32, device=dev)
        ws["logcounts"] = torch.empty((B, cfg.H), dtype=torch.float32, device=dev)
        ws["hf_ws"] = ops.heads_fwd_ws(B, Ls[-1], cfg.H, dev)
        if kind == "predict":
            # two ping-pong activation buffers sized for the longest layer
            ws["ping"] = [ops.new_planes(B, Ls[0], Fp, prec, dev) for _ in range(2)]
        else:
            ws["acts"] = [ops.new_planes(B, L, Fp, prec, dev) for L in Ls]
        if kind == "train":
            ws["masks"] = [torch.empty((B, L, Fp // 64), dtype=torch.int64, device=dev)
                           for L in Ls]
            ws["counts"] = torch.zeros((B, cfg.outputLength, cfg.T), dtype=torch.float32,
                                       device=dev)
            ws["losses"] = torch.zeros((2 * cfg.H,), dtype=torch.float32, device=dev)
            # log-count targets handed in by the generator (generators.py:139-140); without them
            # the loss kernels use log(sum of the profile), which is what that generator computes
            ws["lc_true"] = torch.zeros((B, cfg.H), dtype=torch.float32, device=dev)
            ws["dlogits"] = torch.empty_like(ws["logits"])
            ws["dlogcounts"] = torch.empty_like(ws["logcounts"])
            ws["g"] = [ops.new_planes(B, Ls[0], Fp, prec, dev) for _ in range(2)]
            # one gz buffer per layer (no ping-pong): the tower weight gradients are computed
            # together after the data-gradient chain
            ws["gzl"] = [ops.new_planes(B, L, Fp, prec, dev) for L in Ls]
            joint, joint_heads = -1, False
            if not prec and 0 < len(Ls) - 1 <= 10:
                if self.cp == 8:  # the heads' d8 rows are 16 bytes: they can join as well
                    joint = ops.wgrad_tower_input_ws_bytes(B, Ls[1:], Fp, Ls[0],
                                                           cfg.inputFilterWidth, Ls[-1],
                                                           cfg.outputFilterWidth)
                    joint_heads = joint > 0
                if joint <= 0:
                    joint = ops.wgrad_tower_input_ws_bytes(B, Ls[1:], Fp, Ls[0],
                                                           cfg.inputFilterWidth)
            ws["wg_joint"] = joint > 0
            ws["wg_joint_heads"] = joint_heads
            nb = max([max(joint, 0)] + [ops.wgrad3_ws_bytes(B, L, Fp, prec) for L in Ls[1:]] +
                     ([ops.wgrad3_multi_ws_bytes(B, Ls[1:], Fp)] if 0 < len(Ls) - 1 <= 12 else []) +
                     [ops.input_conv_wgrad_ws_bytes(B, Ls[0], cfg.inputFilterWidth, Fp),
                      ops.heads_wgrad_ws_bytes(B, Ls[-1], cfg.outputFilterWidth, Fp, self.cp)])
            ws["scratch"] = torch.empty((nb // 4 + 4,), dtype=torch.float32, device=dev)
            ws["d8"] = ops.heads_d8(B, Ls[-1], cfg.outputFilterWidth, self.cp,
                                    2 if prec else 1, dev)
            ws["loss_ws"] = ops.loss_pack_ws(B, cfg.T, cfg.H, dev)
        if kind == "shap_x":
            ws["z"] = [torch.empty((B, L, Fp), dtype=torch.float32, device=dev) for L in Ls]
        if kind == "shap_r":
            ws["mult"] = [ops.new_planes(B, L, Fp, prec, dev) for L in Ls]
            ws["ping"] = [ops.new_planes(B, Ls[0], Fp, prec, dev) for _ in range(2)]
            del ws["acts"]
            ws["dlogits"] = torch.zeros_like(ws["logits"])
            ws["dlogcounts"] = torch.zeros_like(ws["logcounts"])
            ws["g"] = [ops.new_planes(B, Ls[0], Fp, prec, dev) for _ in range(2)]
            ws["gz"] = [ops.new_planes(B, Ls[0], Fp, prec, dev) for _ in range(2)]
            ws["dx"] = torch.empty((B, cfg.inputLength, 4), dtype=torch.float32, device=dev)
            ws["d8"] = ops.heads_d8(B, Ls[-1], cfg.outputFilterWidth, self.cp,
                                    2 if prec else 1, dev)
        self._ws[key] = ws
        return ws

    @staticmethod
    def _sub(planes, L):
        """View of the first L rows per sequence of an over-allocated plane set as a new
        contiguous [B, L, F] tensor sharing the same storage start (buffers are reused across
        layers of different lengths, so the view is re-shaped over the flat storage)."""
        hi, lo = planes
        B, _, Fp = hi.shape
        h = hi.view(-1)[:B * L * Fp].view(B, L, Fp)
        l_ = lo.view(-1)[:B * L * Fp].view(B, L, Fp) if lo is not None else None
        return (h, l_)

    # ------------------------------------------------------------------ forward
    def _forward(self, ws, x, save_masks=False, z_list=None, zx_list=None, zx_div=1,
                 mult_list=None, defer_counts=False):
        cfg = self.cfg
        p = self.p
        Ls = cfg.lengths()
        keep = "acts" in ws

        def buf(i):
            return ws["acts"][i] if keep else self._sub(ws["ping"][i % 2], Ls[i])

        a = buf(0)
        ops.onehot_expand(x, ws["x8"])
        exact_first = save_masks or bool(z_list) or bool(zx_list)  # training / attribution
        ops.input_conv_fwd(ws["x8"], self.w_in_op,
                           self.w_in_planes if exact_first else self.w_in_fast,
                           cfg.inputFilterWidth, cfg.numFilters, p["conv1_b"], cfg.Fp, a,
                           mask_out=ws["masks"][0] if save_masks else None,
                           z_out=z_list[0] if z_list else None,
                           zx_in=zx_list[0] if zx_list else None, zx_div=zx_div,
                           mult=mult_list[0] if mult_list else None)
        for i in range(cfg.numLayers):
            d = 2 ** (i + 1)
            nxt = buf(i + 1)
            ops.conv3(a, self.w_fwd[i], (0, d, 2 * d), 0, Ls[i + 1], bias=p["dil_b"][i], resid=a,
                      resid_off=d, out=nxt,
                      mask_out=ws["masks"][i + 1] if save_masks else None,
                      z_out=z_list[i + 1] if z_list else None,
                      zx_in=zx_list[i + 1] if zx_list else None, zx_div=zx_div,
                      out2=mult_list[i + 1] if mult_list else None)
            a = nxt
        if self.heads_tc:
            # defer_counts: the fused loss kernel finishes the log-counts from the per-tile sums
            ops.heads_fwd_tc(a, self.w_hf_op, cfg.outputFilterWidth, cfg.T, cfg.H, p["prof_b"],
                             p["cnt_b"], ws["logits"], None if defer_counts else ws["logcounts"],
                             ws["hf_ws"])
        else:
            ops.heads_fwd(a, p["prof_w"], p["prof_b"], p["cnt_w"], p["cnt_b"], ws["logits"],
                          ws["gap"], ws["logcounts"])
        return a

    def predict_into(self, ws):
        """Forward on ws['x'] -> ws['logits'], ws['logcounts'] (graph-capturable)."""
        self._forward(ws, ws["x"])

    def predict(self, x_u8, batchSize=64):
        """x_u8: (N, Lin, 4) uint8 CUDA tensor -> (logits (N, Lout, T), logcounts (N, H))."""
        cfg = self.cfg
        N = x_u8.shape[0]
        out_l = torch.empty((N, cfg.outputLength, cfg.T), dtype=torch.float32, device=self.device)
        out_c = torch.empty((N, cfg.H), dtype=torch.float32, device=self.device)
        for s in range(0, N, batchSize):
            e = min(N, s + batchSize)
            ws = self._workspace(e - s, "predict")
            ws["x"].copy_(x_u8[s:e])
            self._run_graph(("predict", e - s), lambda ws=ws: self.predict_into(ws))
            out_l[s:e].copy_(ws["logits"])
            out_c[s:e].copy_(ws["logcounts"])
        return out_l, out_c

    def predict_host(self, x_host, batchSize=128, out_logits=None, out_logcounts=None):
        """Streaming prediction for host data (utils.BatchPredictor.runBatch, utils.py:986-1051, at
        makePredictions scale): ``x_host`` (N, Lin, 4) uint8 in (ideally pinned) host memory ->
        pinned host tensors (logits (N, Lout, T), logcounts (N, H)).  Three streams: the H2D copy
        of batch i+1 and the D2H copy of batch i-1 overlap the forward pass of batch i (two
        staging slots each way)."""
        cfg, dev = self.cfg, self.device
        x_host = torch.as_tensor(x_host)
        N, B = x_host.shape[0], int(batchSize)
        if out_logits is None:
            out_logits = torch.empty((N, cfg.outputLength, cfg.T), dtype=torch.float32,
                                     pin_memory=True)
        if out_logcounts is None:
            out_logcounts = torch.empty((N, cfg.H), dtype=torch.float32, pin_memory=True)
        n_full = N // B
        if n_full > 0:
            ws = self._workspace(B, "predict")
            st = ws.get("stream_io")
            if st is None:
                st = {"xin": [torch.empty_like(ws["x"]) for _ in range(2)],
                      "lo": [torch.empty_like(ws["logits"]) for _ in range(2)],
                      "co": [torch.empty_like(ws["logcounts"]) for _ in range(2)],
                      "in_ready": [torch.cuda.Event() for _ in range(2)],
                      "in_free": [torch.cuda.Event() for _ in range(2)],
                      "out_ready": [torch.cuda.Event() for _ in range(2)],
                      "out_free": [torch.cuda.Event() for _ in range(2)],
                      "h2d": torch.cuda.Stream(device=dev), "d2h": torch.cuda.Stream(device=dev)}
                ws["stream_io"] = st
            cur = torch.cuda.current_stream(dev)
            for ev in st["in_free"] + st["out_free"]:
                ev.record(cur)
            body = lambda ws=ws: self.predict_into(ws)  # noqa: E731
            for i in range(n_full):
                k, s, e = i & 1, i * B, (i + 1) * B
                st["h2d"].wait_event(st["in_free"][k])
                with torch.cuda.stream(st["h2d"]):
                    st["xin"][k].copy_(x_host[s:e], non_blocking=True)
                    st["in_ready"][k].record(st["h2d"])
                cur.wait_event(st["in_ready"][k])
                ws["x"].copy_(st["xin"][k], non_blocking=True)
                st["in_free"][k].record(cur)
                self._run_graph(("predict", B), body)
                cur.wait_event(st["out_free"][k])
                st["lo"][k].copy_(ws["logits"], non_blocking=True)
                st["co"][k].copy_(ws["logcounts"], non_blocking=True)
                st["out_ready"][k].record(cur)
                st["d2h"].wait_event(st["out_ready"][k])
                with torch.cuda.stream(st["d2h"]):
                    out_logits[s:e].copy_(st["lo"][k], non_blocking=True)
                    out_logcounts[s:e].copy_(st["co"][k], non_blocking=True)
                    st["out_free"][k].record(st["d2h"])
            st["d2h"].synchronize()
        if n_full * B < N:  # ragged tail through the plain path
            s = n_full * B
            tl, tc = self.predict(x_host[s:].to(dev), batchSize=B)
            out_logits[s:].copy_(tl)
            out_logcounts[s:].copy_(tc)
        return out_logits, out_logcounts

    # ------------------------------------------------------------------ training
    def enable_training(self, grads_flat=None):
        """Allocates gradients and Adam moments.  ``grads_flat``: an external bucket of the right
        size (``parallel.PeerGradExchange.grads``: peer-mapped memory for the fused all-reduce)."""
        if self.grads_flat is None:
            if grads_flat is not None:
                assert grads_flat.numel() == self.params.flat.numel() and \
                    grads_flat.dtype == torch.float32
                grads_flat.zero_()
            self.grads_flat = grads_flat if grads_flat is not None else \
                torch.zeros_like(self.params.flat)
            self.adam_m = torch.zeros_like(self.params.flat)
            self.adam_v = torch.zeros_like(self.params.flat)
            self.g = self.params.views(self.grads_flat)
            self.profile_weight = torch.ones((self.cfg.H,), dtype=torch.float32,
                                             device=self.device)
            self.lam = torch.ones((self.cfg.H,), dtype=torch.float32, device=self.device)

    def forward_backward_into(self, ws, bias_logits=None, bias_logcounts=None, use_bias=None,
                              lc_true=False):
        """Forward + losses + backward on ws['x'], ws['counts'] -> self.grads_flat, ws['losses'].

        For the combined model (models.py:363-380) the frozen bias outputs are added before the
        loss: logits by addition, logcounts by CountsLogSumExp where ``use_bias[h]``."""
        cfg, p, g = self.cfg, self.p, self.g
        Ls = cfg.lengths()
        # solo model: losses, packed loss gradient and the heads' bias gradients in ONE launch
        fused = bias_logits is None and max(cfg.headTasks) <= 8
        defer = fused and self.heads_tc and self.heads_groups == 1
        aL = self._forward(ws, ws["x"], save_masks=True, defer_counts=defer)
        logits, lc = ws["logits"], ws["logcounts"]
        if bias_logits is not None:
            if "logits_c" not in ws:
                ws["logits_c"] = torch.empty_like(logits)
                ws["lc_c"] = torch.empty_like(lc)
                ws["dscale"] = torch.empty_like(lc)
            if callable(bias_logits):
                # combined model: ``bias_logits(ws)`` runs the frozen bias model and the fused
                # transformation + add + log-sum-exp launch into ws['logits_c' / 'lc_c' / 'dscale']
                bias_logits(ws)
            else:
                ops.add_f32(logits, bias_logits, ws["logits_c"])
                ops.counts_lse(use_bias, bias_logcounts, lc, ws["lc_c"], ws["dscale"])
            logits, lc = ws["logits_c"], ws["lc_c"]
        nL = cfg.numLayers
        planes = 2 if cfg.precise else 1
        if fused:
            ops.loss_fwd_bwd_pack(self.task_off, self.host_task_off, logits, lc, ws["counts"],
                                  self.profile_weight, self.lam, ws["losses"], Ls[nL],
                                  cfg.outputFilterWidth, self.cp, planes, ws["d8"], ws["loss_ws"],
                                  dlogits=ws["dlogits"], dlogcounts=ws["dlogcounts"],
                                  dpb=g["prof_b"], dcb=g["cnt_b"],
                                  logcounts_true=ws["lc_true"] if lc_true else None,
                                  counts_partials=ws["hf_ws"] if defer else None,
                                  cb=p["cnt_b"] if defer else None)
        else:
            ops.loss_fwd_bwd(self.task_off, logits, lc, ws["counts"], self.profile_weight,
                             self.lam, ws["losses"], ws["dlogits"], ws["dlogcounts"],
                             logcounts_true=ws["lc_true"] if lc_true else None)
        if bias_logits is not None:
            ops.mul_f32(ws["dlogcounts"], ws["dscale"], ws["dlogcounts"])
        # Backward.  The data-gradient chain comes first (heads -> tower layers, each needing the
        # previous one) and leaves one gz buffer per layer; every tower weight gradient is then
        # computed by ONE split-K launch (bpb_wgrad3_multi) instead of one launch + reduction per
        # layer.
        xf = self._xf_backward(ws)
        if not fused:
            ops.heads_pack_grad(ws["dlogits"], ws["dlogcounts"], Ls[nL], cfg.outputFilterWidth,
                                self.cp, planes, ws["d8"], dpb=g["prof_b"], dcb=g["cnt_b"])
        if not ws.get("wg_joint_heads"):
            ops.heads_wgrad(aL, ws["d8"], cfg.outputFilterWidth, cfg.numFilters, cfg.T, cfg.H,
                            self.cp, planes, g["prof_w"], g["cnt_w"], ws["scratch"])
        if xf:
            # XF (bf16 path, F = 64): nothing ever stores gz = g (.) mask.  ws['gzl'][i] holds the
            # UNMASKED gradient g_i; the data-gradient and weight-gradient kernels AND the staged
            # operand with the layer's 1-bit ReLU mask in shared memory (bpb_conv3_args.mask_a,
            # bpb_wgrad3_args.gz_mask).  Per layer the chain reads g and writes g: 2 planes
            # instead of 4 (gz + g in, g + gz out).
            gl, mk = ws["gzl"], ws["masks"]
            ops.heads_dgrad(ws["d8"], self.w_hd_op, ws["B"], Ls[nL], cfg.outputFilterWidth, cfg.Fp,
                            self.cp, planes, g=gl[nL])
            for i in range(nL - 1, -1, -1):
                d = 2 ** (i + 1)
                ops.conv3(gl[i + 1], self.w_bwd[i], (0, -d, -2 * d), 1, Ls[i], resid=gl[i + 1],
                          resid_off=-d, out=gl[i], mask_a=mk[i + 1])
            layers = [(ws["acts"][i], gl[i + 1], 2 ** (i + 1), g["dil_w"][i], g["dil_b"][i],
                       mk[i + 1]) for i in range(nL)]
            ops.wgrad_tower_input(layers, ws["x8"], gl[0], cfg.inputFilterWidth, cfg.numFilters,
                                  g["conv1_w"], g["conv1_b"], ws["scratch"],
                                  heads=(aL, ws["d8"], cfg.outputFilterWidth, cfg.T, cfg.H,
                                         g["prof_w"], g["cnt_w"]) if ws.get("wg_joint_heads")
                                  else None, gz0_mask=mk[0])
            return
        gcur = self._sub(ws["g"][nL % 2], Ls[nL])
        gzcur = ws["gzl"][nL]
        ops.heads_dgrad(ws["d8"], self.w_hd_op, ws["B"], Ls[nL], cfg.outputFilterWidth, cfg.Fp,
                        self.cp, planes, g=gcur, gz=gzcur, mask_in=ws["masks"][nL])
        for i in range(nL - 1, -1, -1):
            d = 2 ** (i + 1)
            gnext = self._sub(ws["g"][i % 2], Ls[i])
            gznext = ws["gzl"][i]
            ops.conv3(gzcur, self.w_bwd[i], (0, -d, -2 * d), 1, Ls[i], resid=gcur, resid_off=-d,
                      out=gnext if i > 0 else None, out2=gznext, mask_in=ws["masks"][i])
            gcur, gzcur = gnext, gznext
        if ws.get("wg_joint"):
            # tower layers + input convolution: one split-K launch and one reduction
            ops.wgrad_tower_input([(ws["acts"][i], ws["gzl"][i + 1], 2 ** (i + 1), g["dil_w"][i],
                                    g["dil_b"][i]) for i in range(nL)], ws["x8"], ws["gzl"][0],
                                  cfg.inputFilterWidth, cfg.numFilters, g["conv1_w"],
                                  g["conv1_b"], ws["scratch"],
                                  heads=(aL, ws["d8"], cfg.outputFilterWidth, cfg.T, cfg.H,
                                         g["prof_w"], g["cnt_w"]) if ws.get("wg_joint_heads")
                                  else None)
            return
        if 0 < nL <= 12:
            ops.wgrad3_multi([(ws["acts"][i], ws["gzl"][i + 1], 2 ** (i + 1), g["dil_w"][i],
                               g["dil_b"][i]) for i in range(nL)], ws["scratch"])
        else:
            for i in range(nL):
                ops.wgrad3(ws["acts"][i], ws["gzl"][i + 1], 2 ** (i + 1), g["dil_w"][i],
                           g["dil_b"][i], ws["scratch"])
        ops.input_conv_wgrad(ws["x8"], ws["gzl"][0], cfg.inputFilterWidth, cfg.numFilters,
                             g["conv1_w"], g["conv1_b"], ws["scratch"])

    def _xf_backward(self, ws):
        """True when the backward masks its operands in shared memory instead of materialising gz
        planes (BPB_XF=1): bf16 path, F = 64, the joint weight-gradient launch, even row counts
        (the mask words travel as 16-byte aligned bulk copies).  Opt-in: it removes ~200 MB of
        DRAM writes per C1 step and is bit-identical, but measured no faster (the
        data gradient is bound by its per-tile pipeline, not by HBM -- the gz planes it re-reads
        are L2 hits -- and the masker warps cost the weight-gradient launch 8 us)."""
        cfg = self.cfg
        if cfg.precise or cfg.Fp != 64 or not ws.get("wg_joint") or cfg.numLayers < 1:
            return False
        if os.environ.get("BPB_XF", "0") != "1":
            return False
        return all((ws["B"] * L) % 2 == 0 for L in cfg.lengths())

    def apply_adam(self, grad_scale=1.0):
        # the optimiser step counter lives on the device so that the update replays inside a graph
        self._rf_twin = None
        # Adam, the step increment and the refresh of every bf16 operand layout: one launch
        ops.adam_prep_step(self.params.flat, self.grads_flat, self.adam_m, self.adam_v,
                           self.step_and_lr, self._prep_table(), grad_scale=grad_scale)

    def _prep_table(self):
        if getattr(self, "_ptab", None) is None:
            from . import _lib
            cfg, off = self.cfg, self.params.offsets
            t = _lib.PrepTable()
            t.off_dil_w, t.off_conv1_w = off["dil_w"][0], off["conv1_w"][0]
            t.off_prof_w, t.off_cnt_w = off["prof_w"][0], off["cnt_w"][0]
            t.n_layers, t.F, t.Fw = cfg.numLayers, cfg.Fp, cfg.numFilters
            t.W_in, t.W_out, t.T, t.H, t.Cp = (cfg.inputFilterWidth, cfg.outputFilterWidth, cfg.T,
                                               cfg.H, self.cp)
            t.planes, t.in_planes = (2 if cfg.precise else 1), self.w_in_planes
            t.N_heads = (cfg.T + cfg.H + 15) // 16 * 16
            t.w_fwd, t.w_bwd = self.w_fwd.data_ptr(), self.w_bwd.data_ptr()
            t.w_in_op, t.w_hd_op = self.w_in_op.data_ptr(), self.w_hd_op.data_ptr()
            t.w_hf_op = self.w_hf_op.data_ptr() if self.heads_tc else None
            t.hf_kcg = (cfg.Fp // 64) // max(self.heads_groups, 1)
            self._ptab = t
        return self._ptab

    def set_step(self, step, lr):
        self.step = step
        self.step_and_lr[:2].copy_(torch.tensor([float(step), float(lr)], dtype=torch.float32),
                                   non_blocking=False)

    def eval_losses(self, x_u8, counts, bias_logits=None, bias_logcounts=None, use_bias=None,
                    logcounts_true=None):
        """Forward + losses only (validation batches): returns the device vector
        [mnll_h ..., lambda_h * mse_h ...]."""
        self.enable_training()
        cfg = self.cfg
        B = x_u8.shape[0]
        ws = self._workspace(B, "train")
        if "x0" in ws:  # (train_step may have left a host-staging slot bound: use the fixed buffers)
            ws["x"], ws["counts"] = ws["x0"], ws["c0"]
        ws["x"].copy_(x_u8, non_blocking=True)
        ws["counts"].copy_(counts, non_blocking=True)
        if logcounts_true is not None:
            ws["lc_true"].copy_(logcounts_true.reshape(B, cfg.H), non_blocking=True)
        self._forward(ws, ws["x"])
        logits, lc = ws["logits"], ws["logcounts"]
        if bias_logits is not None:
            if "logits_c" not in ws:
                ws["logits_c"] = torch.empty_like(logits)
                ws["lc_c"] = torch.empty_like(lc)
                ws["dscale"] = torch.empty_like(lc)
            if callable(bias_logits):
                # combined model: ``bias_logits(ws)`` runs the frozen bias model and the fused
                # transformation + add + log-sum-exp launch into ws['logits_c' / 'lc_c' / 'dscale']
                bias_logits(ws)
            else:
                ops.add_f32(logits, bias_logits, ws["logits_c"])
                ops.counts_lse(use_bias, bias_logcounts, lc, ws["lc_c"], ws["dscale"])
            logits, lc = ws["logits_c"], ws["lc_c"]
        ops.loss_fwd_bwd(self.task_off, logits, lc, ws["counts"], self.profile_weight, self.lam,
                         ws["losses"], None, None,
                         logcounts_true=ws["lc_true"] if logcounts_true is not None else None)
        del cfg
        return ws["losses"]

    def train_step(self, x_u8, counts, lr, use_graph=True, allreduce=None, world=1,
                   bias_logits=None, bias_logcounts=None, use_bias=None, peers=None,
                   logcounts_true=None):
        """One optimisation step.  ``x_u8`` / ``counts`` may live on the device or in (pinned)
        host memory.  With ``allreduce`` (a callable taking the flat gradient bucket, e.g. an NCCL
        all-reduce) the step is data parallel: local forward/backward, one all-reduce of the
        bucket, identical Adam on every rank with the gradient scaled by 1/world."""
        self.enable_training()
        B = x_u8.shape[0]
        ws = self._workspace(B, "train")
        slot = self._stage_inputs(ws, x_u8, counts)
        try:
            return self._train_step_staged(ws, B, slot, lr, use_graph, allreduce, world,
                                           bias_logits, bias_logcounts, use_bias, peers,
                                           logcounts_true)
        finally:
            self._release_inputs(ws, slot)

    def _train_step_staged(self, ws, B, slot, lr, use_graph, allreduce, world, bias_logits,
                           bias_logcounts, use_bias, peers, logcounts_true):
        lct = logcounts_true is not None
        if lct:  # (B, H) log-count targets of the generator; [B] / [B, 1] per head stacked
            ws["lc_true"].copy_(logcounts_true.reshape(B, self.cfg.H), non_blocking=True)
        self.step += 1
        if lr != getattr(self, "_lr_on_device", None):
            self.step_and_lr[1:2].fill_(float(lr))  # stream-ordered scalar fill, only on change
            self._lr_on_device = lr

        hook = bias_logits if callable(bias_logits) else None
        if peers is not None:
            # data parallel over NVLink peer memory: the optimiser kernel sums the ranks' buckets
            # itself -- one graph per step, no collective call, no host involvement
            assert (bias_logits is None or hook is not None) and \
                self.grads_flat.data_ptr() == peers.grads.data_ptr()

            def body_peer():
                ops.peer_wait(peers.args)
                self.forward_backward_into(ws, hook, lc_true=lct)
                self._rf_twin = None
                ops.adam_prep_step_peer(self.params.flat, self.adam_m, self.adam_v,
                                        self.step_and_lr, self._prep_table(), peers.args)
            if use_graph:
                self._run_graph(("train_peer", B, peers.world, lct, id(hook), slot), body_peer)
            else:
                body_peer()
            return ws["losses"]
        if bias_logits is not None and hook is None:
            # bias outputs handed in as tensors (they change every batch): eager
            self.forward_backward_into(ws, bias_logits, bias_logcounts, use_bias, lc_true=lct)
            if allreduce is not None:
                allreduce(self.grads_flat)
            self.apply_adam(grad_scale=1.0 / world)
        elif allreduce is None:
            def body():
                self.forward_backward_into(ws, hook, lc_true=lct)
                self.apply_adam()
            if use_graph:
                self._run_graph(("train", B, lct, id(hook), slot), body)
            else:
                body()
        else:
            def body_a():
                self.forward_backward_into(ws, hook, lc_true=lct)

            def body_b():
                self.apply_adam(grad_scale=1.0 / world)
            if use_graph:
                self._run_graph(("train_fb", B, lct, id(hook), slot), body_a)
                allreduce(self.grads_flat)
                self._run_graph(("train_opt", B, world), body_b)
            else:
                body_a()
                allreduce(self.grads_flat)
                body_b()
        return ws["losses"]

    def _stage_inputs(self, ws, x_u8, counts):
        """Inputs -> device buffers the step's CUDA graph reads; returns the staging slot (None
        for device inputs), which is part of the graph's key.

        Host inputs go through two staging slots filled on a copy stream, so the host-to-device
        copy of step i+1 overlaps the kernels of step i (the call returns as soon as the work is
        enqueued).  The step READS THE SLOT ITSELF -- ``ws['x'] / ws['counts']`` are re-bound to
        it and there is one graph per slot -- instead of moving it into a fixed buffer first:
        those two device copies were 5 us of every 390 us step.  ``_release_inputs`` hands the
        slot back once the step's work is enqueued."""
        if "x0" not in ws:
            ws["x0"], ws["c0"] = ws["x"], ws["counts"]
        if x_u8.is_cuda and counts.is_cuda:
            ws["x"], ws["counts"] = ws["x0"], ws["c0"]
            ws["x"].copy_(x_u8, non_blocking=True)
            ws["counts"].copy_(counts, non_blocking=True)
            return None
        st = ws.get("h2d")
        if st is None:
            st = {"x": [torch.empty_like(ws["x0"]) for _ in range(2)],
                  "c": [torch.empty_like(ws["c0"]) for _ in range(2)],
                  "ready": [torch.cuda.Event() for _ in range(2)],
                  "free": [torch.cuda.Event() for _ in range(2)],
                  "k": 0, "stream": torch.cuda.Stream(device=self.device)}
            ws["h2d"] = st
            cur0 = torch.cuda.current_stream(self.device)
            for ev in st["free"]:
                ev.record(cur0)
        k = st["k"]
        st["k"] = k ^ 1
        cs, cur = st["stream"], torch.cuda.current_stream(self.device)
        cs.wait_event(st["free"][k])  # the step that read this slot last has finished
        with torch.cuda.stream(cs):
            st["x"][k].copy_(x_u8, non_blocking=True)
            st["c"][k].copy_(counts, non_blocking=True)
            st["ready"][k].record(cs)
        cur.wait_event(st["ready"][k])
        ws["x"], ws["counts"] = st["x"][k], st["c"][k]
        return k

    def _release_inputs(self, ws, slot):
        if slot is not None:
            ws["h2d"]["free"][slot].record(torch.cuda.current_stream(self.device))

    # ------------------------------------------------------------------ CUDA graphs
    def _run_graph(self, key, body):
        g = self._graphs.get(key)
        if g is None:
            # warm up once eagerly (configures kernel attributes outside of capture)
            body()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                body()
            self._graphs[key] = g
            # The eager warm-up already produced this call's result and, for training, already
            # applied the update, so the first call does not replay.
            return
        g.replay()

    # ------------------------------------------------------------------ DeepSHAP (PISA / flat)
    def shap_batch(self, x_u8, refs_u8, dlogits_fn):
        """DeepSHAP-rescale multipliers for Q queries against n references each.

        x_u8: (Q, Lin, 4); refs_u8: (Q*n, Lin, 4), the n references of query q at rows
        q*n .. q*n+n-1.  ``dlogits_fn(ws_x, ws_r)`` must fill ws_r['dlogits'] / ws_r['dlogcounts']
        with the gradient of the metric wrt the outputs of every (query, reference) pair.
        Returns (dx (Q*n, Lin, 4) fp32 multipliers, ws_x, ws_r).  shap.py:347-394, 612-639.
        """
        wx, wr = self.shap_forward(x_u8, refs_u8)
        dlogits_fn(wx, wr)
        return self.shap_backward(wx, wr), wx, wr

    def shap_forward(self, x_u8, refs_u8):
        """The two forward passes of ``shap_batch``: the queries (fp32 pre-activations kept) and
        the references (rescale multipliers against the queries' pre-activations kept).  Separate
        so that a combined model can look at both networks' outputs before seeding the backward
        passes."""
        Q = x_u8.shape[0]
        R = refs_u8.shape[0]
        n = R // Q
        wx = self._workspace(Q, "shap_x")
        wr = self._workspace(R, "shap_r")
        wx["x"].copy_(x_u8)
        wr["x"].copy_(refs_u8)
        self._forward(wx, wx["x"], z_list=wx["z"])
        self._forward(wr, wr["x"], zx_list=wx["z"], zx_div=n, mult_list=wr["mult"])
        return wx, wr

    def shap_backward(self, wx, wr):
        """The multiplier-weighted backward pass from wr['dlogits'] / wr['dlogcounts'] to the
        one-hot input: returns dx (R, Lin, 4) fp32."""
        cfg, p = self.cfg, self.p
        Ls = cfg.lengths()
        R = wr["B"]
        nL = cfg.numLayers
        gcur = self._sub(wr["g"][nL % 2], Ls[nL])
        gzcur = self._sub(wr["gz"][nL % 2], Ls[nL])
        planes = 2 if cfg.precise else 1
        ops.heads_pack_grad(wr["dlogits"], wr["dlogcounts"], Ls[nL], cfg.outputFilterWidth,
                            self.cp, planes, wr["d8"])
        ops.heads_dgrad(wr["d8"], self.w_hd_op, R, Ls[nL], cfg.outputFilterWidth, cfg.Fp, self.cp,
                        planes, g=gcur, gz=gzcur, mult=wr["mult"][nL])
        for i in range(nL - 1, -1, -1):
            d = 2 ** (i + 1)
            gnext = self._sub(wr["g"][i % 2], Ls[i])
            gznext = self._sub(wr["gz"][i % 2], Ls[i])
            ops.conv3(gzcur, self.w_bwd[i], (0, -d, -2 * d), 1, Ls[i], resid=gcur, resid_off=-d,
                      out=gnext if i > 0 else None, out2=gznext, mult=wr["mult"][i])
            gcur, gzcur = gnext, gznext
        planes_id = 2 if cfg.precise else 1
        if ops.input_conv_dgrad_tc_supported(cfg.inputFilterWidth, cfg.Fp, planes_id):
            # transposed input convolution on tensor cores (its operand is rebuilt here: it is
            # only needed for attribution, 3 us, and always matches the current weights)
            ops.prep_input_dgrad_weights(p["conv1_w"], cfg.Fp, self.w_id_op, planes_id)
            ops.input_conv_dgrad_tc(gzcur, self.w_id_op, cfg.inputFilterWidth, wr["dx"])
        else:
            ops.input_conv_dgrad(gzcur, p["conv1_w"], cfg.inputLength, wr["dx"])
        return wr["dx"]
