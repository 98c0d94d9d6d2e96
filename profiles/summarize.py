"""Turn the ncu CSV logs of one fitness evaluation (tests/profile_step.py --pop 64 --evals 1) into the summaries
bench.py and DESIGN.md cite.

    python profiles/summarize.py launches <launches.csv> <out.txt>          # per-kernel share of the step
    python profiles/summarize.py traffic  <conv_dram.csv> <out.json>         # DRAM bytes per conv_tc launch
    python profiles/summarize.py metrics  <metrics.csv> <out.txt>            # per-launch time / DRAM / tensor pipe

The launch list comes from `ncu --metrics gpu__time_duration.sum --clock-control none --csv`; its per-launch times are
cold-cache and serialised, so only the SHARES are meaningful.
"""
import collections
import csv
import json
import re
import sys


def read(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    return rows[h + 1:]


def to_ms(v, unit):
    v = float(v.replace(",", ""))
    return v / 1e6 if unit.startswith("n") else v / 1e3 if unit.startswith("u") else v if unit.startswith("m") else v * 1e3


def metrics(src, dst):
    """Long-format ncu CSV (one row per launch and metric) -> one line per launch: time, DRAM bytes and GB/s, tensor
    pipe / LSU pipe / issue-slot utilisation; plus per-kernel-family totals.  Times are cold-cache and serialised."""
    rows = read(src)
    per = collections.OrderedDict()
    for r in rows:
        key = int(r[0])
        d = per.setdefault(key, dict(name=re.sub(r"^void ", "", r[4]).split("(")[0].split("::")[-1], grid=r[8]))
        d[r[12]] = float(r[14].replace(",", ""))
    fam = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    with open(dst, "w") as f:
        f.write("# one step at P=64 (tests/profile_step.py --pop 64 --evals 1) under ncu, per launch (cold, serialised)\n")
        f.write("#  id  time_us   dram_MB  dram_GB/s  tensor%  lsu%  issue%  kernel  grid\n")
        for k, d in per.items():
            t = d.get("gpu__time_duration.sum", 0.0) / 1e3
            b = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
            tp = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
            f.write(f"{k:5d} {t:9.1f} {b / 1e6:9.1f} {b / max(t, 1e-9) / 1e3:9.1f} {tp:7.1f} "
                    f"{d.get('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 0.0):6.1f} "
                    f"{d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0.0):6.1f}  {d['name'][:48]}  {d['grid']}\n")
            name = "conv_tc_kernel<*>" if "conv_tc_kernel" in d["name"] else d["name"].split("<")[0]
            fam[name][0] += 1
            fam[name][1] += t
            fam[name][2] += b
            fam[name][3] += tp * t
        tot = sum(v[1] for v in fam.values())
        f.write(f"# per kernel family: launches, total time (us), share, DRAM GB, time-weighted tensor pipe %; total {tot / 1e3:.2f} ms\n")
        for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1]):
            f.write(f"# {v[0]:4d} {v[1]:10.1f} {100 * v[1] / tot:6.1f}% {v[2] / 1e9:8.2f} GB {v[3] / max(v[1], 1e-9):6.1f}%  {k}\n")


def launches(src, dst):
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in read(src):
        name = re.sub(r"^void ", "", r[4]).split("(")[0]
        name = "conv_tc_kernel<*>" if "conv_tc_kernel" in name else name.split("::")[-1]
        agg[name][0] += 1
        agg[name][1] += to_ms(r[-1], r[-2])
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# one step at P=64 (tests/profile_step.py), ncu gpu__time_duration.sum per launch (cold, serialised): "
                f"total {tot:.2f} ms, {sum(v[0] for v in agg.values())} launches\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[1]:9.3f} ms {v[0]:5d} launches {100 * v[1] / tot:6.1f} %  {k}\n")
    print(open(dst).read())


def traffic(src, dst):
    rd = wr = ms = 0.0
    ids = set()
    for r in read(src):
        metric, unit, val = r[-3], r[-2], r[-1]
        ids.add(r[0])
        if metric == "dram__bytes_read.sum":
            rd += float(val.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        elif metric == "dram__bytes_write.sum":
            wr += float(val.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        elif metric == "gpu__time_duration.sum":
            ms += to_ms(val, unit)
    n = len(ids)
    out = dict(bytes_per_launch=(rd + wr) / n, launches=n, dram_read_bytes_per_step=rd, dram_write_bytes_per_step=wr,
               kernel_ms_under_ncu=ms,
               source=f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:conv_tc, "
                      f"tests/profile_step.py --pop 64 --evals 1 ({src})")
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"launches": launches, "traffic": traffic, "metrics": metrics}[sys.argv[1]](sys.argv[2], sys.argv[3])
