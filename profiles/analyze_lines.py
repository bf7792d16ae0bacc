"""Per-source-line totals from `ncu --page source --csv --print-source cuda,sass`:
warp-instructions executed and stall samples by file and line (needs -lineinfo)."""
import collections
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr = None, None
lines = collections.OrderedDict()
for r in csv.reader(open(path)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_s, i_e, i_src = r.index("# Samples"), r.index("Instructions Executed"), 1
        continue
    if cur is None or hdr is None or r[0] in ("", "Function Name"):
        continue
    try:
        ln, s, e = int(r[0]), int(r[i_s]), int(r[i_e])
    except ValueError:
        continue
    lines[(cur, ln)] = (e, s, r[i_src].strip()[:90])
tot_e = sum(v[0] for v in lines.values())
tot_s = sum(v[1] for v in lines.values())
print(f"total executed {tot_e / 1e6:.1f}M warp-instructions, {tot_s} samples")
by_file = collections.Counter()
by_file_s = collections.Counter()
for (f, ln), (e, s, _) in lines.items():
    by_file[f] += e
    by_file_s[f] += s
for f, e in by_file.most_common():
    print(f"{f:28s} executed {100 * e / tot_e:5.1f}%  samples {100 * by_file_s[f] / tot_s:5.1f}%")
print("--- top lines by executed instructions")
for (f, ln), (e, s, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{ln:<5d} exec {100 * e / tot_e:5.2f}%  samp {100 * s / tot_s:5.2f}%  {src}")
