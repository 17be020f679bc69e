"""Markdown digest of ncu --set full reports for profiles/: python scripts/ncu_report.py out.md rep1.ncu-rep [rep2 ...]"""
import csv, io, subprocess, sys

KEYS = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs/thread"), ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "HMMA subpipe %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("smsp__inst_executed.sum", "warp instructions")]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(txt)))
    return r[0], r[1], r[2]


def top_sass(rep, n=10):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hi]
    a, s = h.index("Warp Stall Sampling (All Samples)"), h.index("Source")
    data = []
    for r in rows[hi + 1:]:
        try: data.append((int(r[a]), r[s].strip()))
        except (ValueError, IndexError): pass
    tot = sum(d[0] for d in data) or 1
    return tot, sorted(data, reverse=True)[:n]


out = open(sys.argv[1], "w")
out.write("# ncu --set full digests (B200, --clock-control none; cold caches, one launch each)\n")
for rep in sys.argv[2:]:
    h, u, v = raw(rep)
    d = dict(zip(h, zip(u, v)))
    out.write(f"\n## {rep.split('/')[-1]}: `{d['Kernel Name'][1][:110]}`\n\n| metric | value |\n|---|---|\n")
    for k, label in KEYS:
        if k in d: out.write(f"| {label} | {d[k][1]} {d[k][0]} |\n")
    st = sorted(((float(d[k][1]), k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for k in h
                 if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio')), reverse=True)[:6]
    out.write("\nwarp stalls per issue: " + ", ".join(f"{n} {x:.2f}" for x, n in st) + "\n\n")
    tot, top = top_sass(rep)
    out.write(f"top sampled SASS ({tot} samples):\n\n```\n")
    for c, src in top: out.write(f"{100.0 * c / tot:5.1f}%  {src[:110]}\n")
    out.write("```\n")
out.close()
