"""Summaries of ncu captures for profiles/: (1) per-kernel totals of a `--metrics gpu__time_duration.sum` launch
list, (2) the key `--set full` metrics of each captured launch of a .ncu-rep (read with `ncu -i --page raw --csv`)."""
import csv, json, re, subprocess, sys, collections


def _us(v, u):
    return v / 1000 if u.startswith("n") else (v if u.startswith("u") else v * 1000)


def _bytes(v, u):
    u = u.lower()
    return v * (1e9 if u.startswith("g") else 1e6 if u.startswith("m") else 1e3 if u.startswith("k") else 1.0)


def launch_list(path, out, last=0):
    """Per-kernel totals of a launch list; `last` > 0 keeps only the last `last` launches (one pass of a two-pass driver).
    When the list also holds dram__bytes_read/write.sum, their per-kernel totals are added."""
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    per = collections.OrderedDict()
    for r in csv.DictReader(lines[start:]):
        d = per.setdefault(r["ID"], {"k": re.sub(r"\(.*", "", r["Kernel Name"]), "us": 0.0, "rd": 0.0, "wr": 0.0})
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Name"] == "gpu__time_duration.sum":
            d["us"] = _us(v, r["Metric Unit"])
        elif r["Metric Name"] == "dram__bytes_read.sum":
            d["rd"] = _bytes(v, r["Metric Unit"])
        elif r["Metric Name"] == "dram__bytes_write.sum":
            d["wr"] = _bytes(v, r["Metric Unit"])
    rows = list(per.values())
    if last:
        rows = rows[-last:]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for d in rows:
        a = agg[d["k"]]
        a[0] += 1; a[1] += d["us"]; a[2] += d["rd"]; a[3] += d["wr"]
    tot = sum(v[1] for v in agg.values())
    has_dram = any(v[2] or v[3] for v in agg.values())
    with open(out, "w") as f:
        f.write("kernel,launches,total_us,share" + (",dram_read_MB,dram_write_MB" if has_dram else "") + "\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'"{k}",{v[0]},{v[1]:.1f},{v[1] / tot:.4f}' + (f",{v[2] / 1e6:.1f},{v[3] / 1e6:.1f}" if has_dram else "") + "\n")
    print(out, "total ms", tot / 1000, "launches", len(rows))
    return agg


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "sm__cycles_elapsed.max"]


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": re.sub(r"\(.*", "", r[hdr.index("Kernel Name")])}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                d[k] = f"{r[i]} {units[i]}".strip()
        res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    print(out, len(res), "launches")


if __name__ == "__main__":
    if sys.argv[1] == "list":
        launch_list(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 0)
    else:
        full(sys.argv[2], sys.argv[3])
