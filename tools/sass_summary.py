"""profiles/rNN_sass_summary.txt + rNN_sass_hot_kernels.txt.xz: cuobjdump -sass of the kernels of the
C2 frame, with an opcode histogram per kernel (what the issue slots go to) and the mnemonics that
prove bulk asynchronous copies / shared-memory atomics. usage: sass_summary.py <round tag, e.g. r02>"""
import collections
import lzma
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = os.path.join(ROOT, "rive-runtime_b200", "_build", "librivecuda.so")
want = ["raster_spans_kernel", "raster_tiles_kernel", "raster_tiles_exact_kernel", "setup_patches_kernelILb0", "tessellate_kernel", "sort_tiles_kernel", "scatter_kernel"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks, cur, name = {}, None, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        cur = blocks.setdefault(name, [])
    if cur is not None:
        cur.append(line)
out = [f"# cuobjdump -sass of librivecuda.so (sm_100a): the kernels of the C2 frame (span rasteriser) and the two other tile rasterisers.",
       "# Opcode histogram per kernel (what the issue slots go to), then the mnemonics of interest: UBLKCP = cp.async.bulk (TMA bulk copy of the",
       "# tile's id list), SYNCS = mbarrier, ATOMS / REDS = shared-memory atomics (the span rasteriser's planes), VOTE / SHFL = warp collectives,",
       "# CCTL / prefetch. No tcgen05: nothing on this path is a dense contraction (DESIGN.md section 4).", ""]
full = []
for w in want:
    for fn, lines in blocks.items():
        if w in fn:
            ops = collections.Counter()
            for l in lines:
                m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
                if m:
                    ops[m.group(1)] += 1
            total = sum(ops.values())
            out.append(f"== {fn}: {total} instructions; " + ", ".join(f"{k} {v}" for k, v in ops.most_common(24)))
            interest = {k: v for k, v in ops.items() if k in ("UBLKCP", "SYNCS", "ATOMS", "REDS", "RED", "ATOMG", "VOTE", "SHFL", "CCTL", "LDGSTS", "REDUX", "DFMA", "DMUL", "DADD", "STG", "LDG", "LDS", "STS", "BAR")}
            out.append("   of interest: " + ", ".join(f"{k} {v}" for k, v in sorted(interest.items())))
            out.append("")
            full += lines + [""]
out.append(f"# Full listings: profiles/{tag}_sass_hot_kernels.txt.xz (xz -dc to read).")
open(os.path.join(ROOT, "profiles", f"{tag}_sass_summary.txt"), "w").write("\n".join(out) + "\n")
with lzma.open(os.path.join(ROOT, "profiles", f"{tag}_sass_hot_kernels.txt.xz"), "wt") as f:
    f.write("\n".join(full))
print("\n".join(out[:14]))
