"""Per-source-line instruction counts of one kernel from a .ncu-rep (cuda,sass view)."""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
agg = []
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        ia = hdr.index("Instructions Executed")
        ist = hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit() and r[ia].isdigit():
        agg.append((int(r[ia]), int(r[ist]) if r[ist].isdigit() else 0, cur_file, int(r[0]), r[1].strip()[:110]))
tot = sum(a[0] for a in agg)
tots = sum(a[1] for a in agg)
print(f"total warp instructions {tot}, samples {tots}")
for a in sorted(agg, reverse=True)[:top]:
    print(f"{100*a[0]/tot:5.1f}% inst {100*a[1]/max(tots,1):5.1f}% stall  {a[2]}:{a[3]}  {a[4]}")
if len(sys.argv) > 3:
    # ranges: file:lo-hi=name,...
    buckets = {}
    spec = []
    for item in sys.argv[3].split(","):
        rng, name = item.split("=")
        f, lh = rng.split(":")
        lo, hi = lh.split("-")
        spec.append((f, int(lo), int(hi), name))
    for a in agg:
        name = "other"
        for f, lo, hi, nm in spec:
            if a[2] == f and lo <= a[3] <= hi:
                name = nm
                break
        b = buckets.setdefault(name, [0, 0])
        b[0] += a[0]
        b[1] += a[1]
    for k, v in sorted(buckets.items(), key=lambda kv: -kv[1][0]):
        print(f"{k:20s} {100*v[0]/tot:5.1f}% inst ({v[0]/1e6:8.1f} M) {100*v[1]/max(tots,1):5.1f}% stall samples")
