"""Join gpurun_out/unet_seq.json (ordered ops with shapes) with the ncu launch list (per-launch durations):
prints time by kernel, and by (op, shape)."""
import csv, json, sys, collections, re
seq = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/unet_seq.json"))
path = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/unet_launches_r01b.csv"
lines = open(path).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(lines[start:]))
launches = []
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1000 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000)
    launches.append((r["Kernel Name"], r["Grid Size"], us))
tot = sum(l[2] for l in launches)
print(f"launches {len(launches)}  total {tot / 1000:.2f} ms")
byk = collections.defaultdict(lambda: [0, 0.0])
for n, g, us in launches:
    k = re.sub(r"\(.*", "", n)
    byk[k][0] += 1
    byk[k][1] += us
for k, v in sorted(byk.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{v[1] / 1000:8.3f} ms {100 * v[1] / tot:5.1f}%  n={v[0]:4d}  {k[:110]}")
# ---- attach shapes: walk ops in order, consuming the expected kernels
EXPECT = {"gemm": ["gemm_bf16_tcgen05_kernel", "gemv_bf16_kernel"], "conv_nhwc": ["gemm_bf16_tcgen05_kernel"],
          "conv_nhwc_direct": ["conv_direct"], "layernorm": ["rownorm"], "groupnorm_nhwc": ["gn_stats", "gn_apply", "gn_fused"],
          "attention": ["flash_attn"], "attention_short": ["attn_short"], "add_rowgroup": ["add_rowgroup"],
          "upsample2x_nhwc": ["upsample"], "add": ["add_"], "cfg_combine": ["cfg_"], "rmsnorm": ["rownorm"]}
i = 0
agg = collections.defaultdict(lambda: [0, 0.0])
other = 0.0
for op in seq:
    pats = EXPECT[op["op"]]
    need = 2 if op["op"] == "groupnorm_nhwc" else 1
    got = 0
    t = 0.0
    while i < len(launches) and got < need:
        n, g, us = launches[i]
        i += 1
        if any(p in n for p in pats):
            got += 1
            t += us
            if "gn_fused" in n:
                break
        else:
            other += us
    key = f'{op["op"]} {op["shapes"][:2]} {op["pos"]} {op["kw"]}'
    agg[key][0] += 1
    agg[key][1] += t
while i < len(launches):
    other += launches[i][2]
    i += 1
print(f"torch-side / unmatched kernels: {other / 1000:.2f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:50]:
    print(f"{v[1] / 1000:8.3f} ms  n={v[0]:4d}  {v[1] / v[0]:8.1f} us  {k}")
