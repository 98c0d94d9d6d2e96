"""Text summary of an `ncu --set full --import-source on` report: per kernel the key metrics, the warp-stall mix and
the SASS instructions that collected the most stall samples.
    python profiles/ncu_summary.py gpurun_out/hot3_r02v.ncu-rep > profiles/r02_ncu_hot3.txt
(reads the report with `ncu -i ... --page raw|source --csv`; no GPU needed)"""
import csv
import io
import subprocess
import sys

KEYS = ["launch__grid_size", "launch__registers_per_thread", "gpu__time_duration.sum", "sm__cycles_elapsed.max",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def ncu(*args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def main(rep):
    rows = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    print(f"# {rep}: {len(rows) - 2} kernel launch(es); ncu --set full --clock-control none (cold, serialised)")
    for r in rows[2:]:
        print("\n== " + r[hdr.index("Kernel Name")][:110])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"   {k:70s} {r[i]:>16s} {units[i]}")
        st = []
        for i, h in enumerate(hdr):
            if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                v = float(r[i].replace(",", "")) if r[i] else 0.0
                st.append((v, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        print("   warp stalls per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
    # per-instruction stall samples (the source page lists every kernel in launch order)
    text = ncu("-i", rep, "--page", "source", "--csv")
    blocks = text.split('"Kernel Name",')[1:]
    seen = set()
    for b in blocks:
        lines = list(csv.reader(io.StringIO('"Kernel Name",' + b)))
        name, hdr2 = lines[0][1], lines[1]
        if (name, len(lines)) in seen:      # (the page lists a kernel once per view)
            continue
        seen.add((name, len(lines)))
        isrc, iall, iex = hdr2.index("Source"), hdr2.index("Warp Stall Sampling (All Samples)"), hdr2.index("Instructions Executed")
        data = []
        for r in lines[2:]:
            try:
                data.append((int(r[iall]), int(r[iex]), r[isrc].strip()))
            except (ValueError, IndexError):
                pass
        tot = sum(d[0] for d in data) or 1
        print(f"\n== stall samples by SASS instruction: {name[:100]}  (total {tot})")
        for i in sorted(sorted(range(len(data)), key=lambda i: -data[i][0])[:12]):
            print(f"   #{i:5d} {100.0 * data[i][0] / tot:5.1f} %  executed {data[i][1]:>9d}  {data[i][2][:90]}")


if __name__ == "__main__":
    main(sys.argv[1])
