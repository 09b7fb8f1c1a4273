"""Folds an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of
`bench.py --steps 1 --warmup 3 --no-profile` into per-kernel time / DRAM traffic of the timed step
(-> profiles/r2_ncu_traffic_pipeline_b64.json, read by bench.py for roofline.traffic).
usage: ncu_traffic.py launches.csv launches_per_step out.json"""
import collections, csv, json, re, sys

src, per_step, dst = sys.argv[1], int(sys.argv[2]), sys.argv[3]
lines = [l for l in open(src) if not l.startswith("==")]
by = collections.OrderedDict()
for row in csv.DictReader(lines):
    by.setdefault(int(row["ID"]), {"name": row["Kernel Name"]})[row["Metric Name"]] = (
        float(row["Metric Value"].replace(",", "")), row["Metric Unit"])
ids = sorted(by)
step = [by[i] for i in ids[3 * per_step:4 * per_step]]  # warm-up steps 0..2, timed step 3
assert len(step) == per_step, (len(ids), per_step)
unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for k in step:
    nm = re.sub(r"<.*", "", k["name"].replace("void ", "").replace("cvb::", "")).split("(")[0]
    a = agg[nm]
    a[0] += 1
    for j, m in enumerate(("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum")):
        v, u = k[m]
        a[j + 1] += v * unit[u]
tot = sum(a[1] for a in agg.values())
alias = {"k_attn_win": "k_attn_tc", "k_ln_rows_t": "k_ln_rows"}
out = {}
for nm, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    out[alias.get(nm, nm)] = {"launches": a[0], "ms": a[1] * 1e3, "share": a[1] / tot, "dram_read_bytes": a[2],
                              "dram_write_bytes": a[3], "traffic_bytes_per_launch": (a[2] + a[3]) / a[0],
                              "dram_GBps": (a[2] + a[3]) / a[1] / 1e9}
    print(f"{nm:26s} n={a[0]:3d} {a[1]*1e3:8.3f} ms share {a[1]/tot:.3f} dram {(a[2]+a[3])/1e9:7.3f} GB -> {(a[2]+a[3])/a[1]/1e9:6.0f} GB/s")
json.dump({"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                     "on `bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-extras` (cfg2 pipeline, tiny, B=64); the timed step",
           "first_kernel": step[0]["name"][:60], "last_kernel": step[-1]["name"][:60],
           "launches_per_step": per_step, "total_ms": tot * 1e3, "kernels": out}, open(dst, "w"), indent=1)
