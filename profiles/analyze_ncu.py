"""Summarise an ncu report exported with `--page raw --csv` and `--page source --csv`."""
import collections
import csv
import re
import sys

raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ",
        "launch__registers_per_thread", "sm__warps_active.avg.pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct", "smsp__issue_active.avg.pct",
        "smsp__inst_executed.sum ", "sm__cycles_elapsed.max ", "smsp__pcsamp_warps_issue_stalled",
        "smsp__pcsamp_sample_count", "smsp__cycles_active.avg", "launch__grid_size",
        "launch__block_size", "sm__throughput.avg.pct", "gpu__dram_throughput.avg.pct",
        "lts__t_bytes.sum ", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ",
        "sass__inst_executed_register_spilling"]
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(w.strip()) for w in want) and "not_issued" not in h:
        try:
            if "pcsamp_warps" in h and float(v) < 1500:
                continue
        except ValueError:
            pass
        print(f"{h} [{u}] = {v}")
rows = list(csv.reader(open(src)))
hdr = rows[1]
ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops, samp, data = collections.Counter(), collections.Counter(), []
for r in rows[2:]:
    try:
        n, s = int(r[ie]), int(r[isamp])
    except (ValueError, IndexError):
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia])
    op = m.group(2).split(".")[0] if m else r[ia][:10]
    ops[op] += n
    samp[op] += s
    data.append((s, n, r[ia]))
tot, totsamp = sum(ops.values()), sum(samp.values())
print("total warp-inst", tot, "samples", totsamp)
for op, n in ops.most_common(16):
    print(f"{op:12s} {n:12d} {100 * n / tot:5.1f}%  samples {100 * samp[op] / totsamp:5.1f}%")
reg = collections.OrderedDict()
for i, d in enumerate(data):
    reg.setdefault(i // 1000, [0, 0])
    reg[i // 1000][0] += d[0]
    reg[i // 1000][1] += d[1]
for k, v in reg.items():
    print(f"instr {k * 1000:6d}: samples {100 * v[0] / totsamp:5.1f}%  executed {v[1] / 1e6:8.1f}M")
for i in sorted(range(len(data)), key=lambda i: -data[i][0])[:12]:
    print(i, data[i][0], data[i][1], data[i][2][:80])
