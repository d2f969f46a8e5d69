#!/usr/bin/env python
"""DRAM traffic of ONE training step, kernel by kernel, with the caches left alone.

  # on the GPU box (under ncu; --cache-control none keeps L2 contents between kernels, the two
  # DRAM counters and the duration fit one pass, so nothing is replayed):
  ncu --profile-from-start off --cache-control none --clock-control none \
      --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
      --csv --log-file gpurun_out/step_traffic_c1.csv python tools/step_traffic.py run c1
  # here: aggregate per kernel family -> profiles/step_traffic_r2_c1.json (read by bench.py)
  python tools/step_traffic.py summarize gpurun_out/step_traffic_c1.csv profiles/step_traffic_r2_c1.json

The profiled region is one eager (un-graphed) step after three warm-up steps on the same
buffers, i.e. with L2 in the state the previous step left it, which is what a step of the timed
loop sees.  Round 1 quoted a constant taken from a cache-flushed replay instead.
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(config):
    import torch
    import bench
    from bpreveal_b200.engine import NetConfig, SoloNet
    cfgd = {"c1": bench.C1, "c4": bench.C4, "c3": bench.C3R}[config]
    batch = 64 if config != "c4" else 16
    dev = torch.device("cuda:0")
    net = SoloNet(NetConfig(**cfgd), dev)
    net.init_random(seed=1234)
    net.enable_training()
    xs, cs = bench.synth(batch * 2, cfgd, 20261019)
    x, c = torch.tensor(xs, device=dev), torch.tensor(cs, device=dev)
    kw = {}
    if config == "c3":
        hook, keep = bench.c3_bias_hook(net, batch, dev)
        kw["bias_logits"] = hook
    for i in range(3):
        net.train_step(x[(i % 2) * batch:(i % 2 + 1) * batch], c[(i % 2) * batch:(i % 2 + 1) * batch],
                       0.004, use_graph=False, **kw)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    net.train_step(x[batch:], c[batch:], 0.004, use_graph=False, **kw)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


def family(name):
    """Kernel name (demangled, with template arguments) -> the family names of bench.py."""
    if "conv_tc_kernel" in name:
        targs = [a.strip().split(")")[-1] for a in name.split("<", 1)[1].split(">", 1)[0].split(",")]
        epi, resid = targs[0], targs[6] == "1"  # (EPI, NT, RESIDENT, PLANES, HALO, EG, RESID[, XF])
        if not resid:
            return "input_conv_fwd" if epi in ("0", "1", "2") else "heads_dgrad"
        return "conv3_fwd" if epi in ("0", "1", "2") else "conv3_dgrad"
    for key, fam in (("wgrad3_tc_kernel", "wgrad_tower_input"), ("wgrad_reduce", "wgrad_tower_input"),
                     ("heads_fwd_tc", "heads_fwd_tc"), ("loss_pack", "loss_fwd_bwd_pack"),
                     ("loss_kernel", "loss_fwd_bwd"), ("adam", "adam_prep_step"),
                     ("onehot_expand", "onehot_expand"), ("pack_grad", "heads_pack_grad"),
                     ("combine_bias", "combine_bias")):
        if key in name:
            return fam
    return name.split("<")[0].split("(")[0]


def summarize(src, dst):
    rows = [r for r in csv.reader(open(src, encoding="utf-8")) if len(r) > 5]
    hdr = rows[0]
    ik, im, iv, iid = (hdr.index("Kernel Name"), hdr.index("Metric Name"),
                       hdr.index("Metric Value"), hdr.index("ID"))
    per = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        e = per.setdefault(r[iid], {"kernel": r[ik]})
        e[r[im]] = v
    fams = collections.OrderedDict()
    for e in per.values():
        f = fams.setdefault(family(e["kernel"]), {"launches": 0, "read": 0.0, "write": 0.0, "ns": 0.0})
        f["launches"] += 1
        f["read"] += e.get("dram__bytes_read.sum", 0.0)
        f["write"] += e.get("dram__bytes_write.sum", 0.0)
        f["ns"] += e.get("gpu__time_duration.sum", 0.0)
    tot_ns = sum(f["ns"] for f in fams.values())
    out = {"source": os.path.basename(src),
           "how": "ncu --profile-from-start off --cache-control none --clock-control none, one "
                  "eager step after three warm-up steps (tools/step_traffic.py)",
           "dram_bytes_per_step": {k: f["read"] + f["write"] for k, f in fams.items()},
           "families": {k: {"launches": f["launches"], "dram_read": f["read"], "dram_write": f["write"],
                            "us_under_ncu": f["ns"] / 1e3, "share": f["ns"] / tot_ns}
                        for k, f in fams.items()},
           "total_dram_bytes": sum(f["read"] + f["write"] for f in fams.values())}
    with open(dst, "w", encoding="utf-8") as fp:
        json.dump(out, fp, indent=1)
    for k, f in sorted(fams.items(), key=lambda kv: -kv[1]["ns"]):
        print(f"{k:22s} n={f['launches']:3d} {f['ns'] / 1e3:8.1f} us  read {f['read'] / 1e6:8.1f} MB  "
              f"write {f['write'] / 1e6:8.1f} MB")
    print("total", out["total_dram_bytes"] / 1e6, "MB")


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run(sys.argv[2] if len(sys.argv) > 2 else "c1")
    else:
        summarize(sys.argv[2], sys.argv[3])
