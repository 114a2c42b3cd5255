"""Per-launch view of an `ncu --metrics gpu__time_duration.sum --csv` launch list: prints the kernels of ONE pass in
launch order (optionally starting at the n-th launch of a marker kernel) and the per-kernel totals of that pass.
This is what exposed the 274 us GELU epilogue and the one-CTA column-mean finalize in the FocalNet forward.

  python tools/launch_breakdown.py gpurun_out/focal_launches.csv --marker im2col --pass 1 [--skip-torch]"""
import argparse, collections, csv, re


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--marker", default="", help="substring of the kernel that opens a pass (default: whole file)")
    ap.add_argument("--pass", dest="npass", type=int, default=0, help="which occurrence of the marker opens the pass")
    ap.add_argument("--skip-torch", action="store_true", help="drop at:: (torch) kernels")
    ap.add_argument("--top", type=int, default=40)
    a = ap.parse_args()
    lines = open(a.csv).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = [r for r in csv.DictReader(lines[start:]) if r["Metric Name"] == "gpu__time_duration.sum"]
    names = [re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "") for r in rows]
    lo, hi = 0, len(rows)
    if a.marker:
        marks = [i for i, n in enumerate(names) if a.marker in n]
        lo = marks[a.npass]
        hi = marks[a.npass + 1] if a.npass + 1 < len(marks) else len(rows)
    agg = collections.defaultdict(lambda: [0, 0.0])
    seq = []
    for r, n in zip(rows[lo:hi], names[lo:hi]):
        if a.skip_torch and n.startswith("at::"):
            continue
        v = float(r["Metric Value"].replace(",", ""))
        us = v / 1000 if r["Metric Unit"].startswith("n") else v
        agg[n][0] += 1
        agg[n][1] += us
        seq.append((n, us, r.get("Grid Size", "")))
    tot = sum(v[1] for v in agg.values())
    print(f"pass of {len(seq)} launches, {tot / 1e3:.3f} ms of kernel time (ncu: cold-cache, serialised)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} {v[0]:5d} {v[1]:9.1f} us {v[1] / tot:6.3f}")
    print("--- first launches in order")
    for n, us, g in seq[:a.top]:
        print(f"{us:8.1f} us  {g:14s} {n[:80]}")


if __name__ == "__main__":
    main()
