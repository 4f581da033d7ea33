"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total time, share.
usage: python tools/summarize_launches.py profiles/bench_launches_r01.csv"""
import collections, csv, sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
ix = {h: i for i, h in enumerate(rows[0])}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) < len(ix):
        continue
    v = float(r[ix["Metric Value"]]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r[ix["Metric Unit"]]]
    name = r[ix["Kernel Name"]]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
OURS = ("scan_", "dwconv", "causal_conv1d", "pooled_attn", "local_attn", "layernorm_", "in_sums", "in_apply", "in_finalize",
        "in_param", "colsum", "avgpool", "linattn", "walk_", "residual_scale", "silu_gate", "diff_lambda", "copy_rows", "bias_add_cl")
ours = {k: v for k, v in agg.items() if ("mlagg::" in k or any(t in k for t in OURS)) and "at::native" not in k}
GROUPS = (("tcgen05 projection GEMMs", ("gemm_tc",)), ("selective scan", ("scan_fwd", "scan_bwd")),
          ("walk pack / unpack", ("walk_pack", "walk_unpack")), ("depthwise conv", ("dwconv", "dw3x3_")), ("local attention", ("local_attn",)),
          ("pooled attention", ("pooled_attn", "avgpool")), ("LayerNorm", ("layernorm_",)),
          ("instance norm", ("in_sums", "in_apply", "in_finalize", "in_param")),
          ("element-wise seams / copies", ("residual_scale", "silu_gate", "diff_lambda", "copy_rows", "bias_add_cl", "colsum")))
print(f"{sum(v[0] for v in agg.values())} launches, {tot:.2f} ms summed device time (ncu lists: cold-cache, serialised -- compare "
      f"SHARES; CUPTI lists from bench.py --trace-step: in-stream durations)")
print(f"libmlagg_b200.so kernels: {sum(v[0] for v in ours.values())} launches, {sum(v[1] for v in ours.values()):.2f} ms = "
      f"{100 * sum(v[1] for v in ours.values()) / tot:.1f} % of the step")
for gname, pats in GROUPS:
    sel = [v for k, v in ours.items() if any(t in k for t in pats)]
    if sel:
        print(f"   {gname:30s} {sum(v[0] for v in sel):5d}x {sum(v[1] for v in sel):7.3f} ms {100 * sum(v[1] for v in sel) / tot:5.1f} %")
glue = [v for k, v in agg.items() if "at::native" in k and any(t in k for t in ("direct_copy", "CUDAFunctor_add", "FillFunctor", "CatArray"))]
print(f"   {'torch copies / adds / fills / cat':30s} {sum(v[0] for v in glue):5d}x {sum(v[1] for v in glue):7.3f} ms {100 * sum(v[1] for v in glue) / tot:5.1f} %")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{v[1]:8.3f} ms {v[0]:5d}x {100 * v[1] / tot:5.1f} %  {'*' if k in ours else ' '} {k[:110]}")
