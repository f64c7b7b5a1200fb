"""Warp-stall samples per CUDA source line from `ncu --page source --csv --print-source cuda,sass` output.
usage: python tools/ncu_lines.py report.ncu-rep kernel_regex [launch_index]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Function Name"]
blocks = []
for n, i in enumerate(starts):
    j = starts[n + 1] - 1 if n + 1 < len(starts) else len(rows)
    blocks.append((rows[i][1], rows[i + 1], rows[i + 2:j]))
name, hdr, body = blocks[which]
print(name[:120])
si = hdr.index("Warp Stall Sampling (All Samples)")
ii = hdr.index("Instructions Executed")
agg = {}
for r in body:
    if len(r) <= max(si, ii) or not r[0].isdigit() or r[2] != "-":
        continue
    agg[int(r[0])] = (int(r[si] or 0), int(r[ii] or 0), r[1])
tot = sum(v[0] for v in agg.values()) or 1
toti = sum(v[1] for v in agg.values()) or 1
print(f"total samples {tot}, instructions {toti}")
for ln, (s, ins, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{ln:5d} {100*s/tot:5.1f}% stall  {100*ins/toti:5.1f}% inst  {src.strip()[:110]}")
