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
ours = {k: v for k, v in agg.items() if any(t in k for t in OURS) and "at::native" not in k}
print(f"{sum(v[0] for v in agg.values())} launches, {tot:.2f} ms summed device time (cold-cache, serialised: compare SHARES)")
print(f"libmlagg_b200.so kernels: {sum(v[0] for v in ours.values())} launches, {sum(v[1] for v in ours.values()):.2f} ms = "
      f"{100 * sum(v[1] for v in ours.values()) / tot:.1f} % of the step")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{v[1]:8.3f} ms {v[0]:5d}x {100 * v[1] / tot:5.1f} %  {'*' if k in ours else ' '} {k[:110]}")
