"""Summarise ncu outputs: a launch-list CSV (per-kernel device time) and/or a
full .ncu-rep capture (key metrics + hottest SASS by executed instructions)."""
import csv
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, {}
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            name = d["Kernel Name"].split("(")[0][:44]
            agg.setdefault(name, []).append(float(d["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    for k, v in agg.items():
        print(f"{k:46s} n={len(v):3d} mean={sum(v)/len(v)/1e3:9.1f} us share={100*sum(v)/tot:5.1f}%")


def report(path, top=0.006):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, v = rows[0], rows[2]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]
    for k in want:
        if k in h:
            i = h.index(k)
            print(f"{k:86s} {rows[1][i]:>14s} {v[i]}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr, data = rows[1], rows[2:]
    ia, isrc, ist = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
    tot = sum(int(r[ia]) for r in data if r[ia].isdigit())
    print("total warp instructions", tot)
    for i, r in enumerate(data):
        if r[ia].isdigit() and int(r[ia]) > tot * top:
            print(f"{i:5d} {int(r[ia]):12d} {r[ist]:>7s}  {r[isrc][:100]}")


def traffic_json(rep, out):
    """profiles/r02_raster_traffic.json: DRAM bytes of one launch of the captured kernel (bench.py's
    roofline.traffic), tied to the kernel sources by bench.kernel_sources_sha256()."""
    import json
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units, v = rows[0], rows[1], rows[2]

    def val(name):
        i = h.index(name)
        x = float(v[i].replace(",", ""))
        return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[i], 1)

    kernel = v[h.index("Kernel Name")].split("(")[0].split("::")[-1]
    json.dump({"kernel": kernel, "dram_bytes_read": int(val("dram__bytes_read.sum")), "dram_bytes_write": int(val("dram__bytes_write.sum")),
               "duration_ms_under_ncu": val("gpu__time_duration.sum"), "workload": "c2 (tests/golden/c2_4k.rvct.xz via tools/quickbench.py)",
               "capture": os.path.basename(rep), "sources_sha256": bench.kernel_sources_sha256()}, open(out, "w"), indent=1)
    print(open(out).read())


if __name__ == "__main__":
    if len(sys.argv) == 4 and sys.argv[1] == "--traffic-json":
        traffic_json(sys.argv[3], sys.argv[2])
        sys.exit(0)
    for p in sys.argv[1:]:
        print("==", p)
        if p.endswith(".csv"):
            launches(p)
        else:
            report(p)
