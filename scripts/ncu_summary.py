"""One-paragraph summary of an ncu --set full report: python scripts/ncu_summary.py rep.ncu-rep"""
import csv, io, subprocess, sys
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(txt)))
    h, v = r[0], r[2]
    d = dict(zip(h, v))
    g = lambda k: d.get(k, "?")
    print(f"== {rep}: {g('Kernel Name')[:70]}")
    print(f"   time {g('gpu__time_duration.sum')} us  grid {g('launch__grid_size')}  dram {g('dram__throughput.avg.pct_of_peak_sustained_elapsed')}%  "
          f"lts {g('lts__throughput.avg.pct_of_peak_sustained_elapsed')}%  sm {g('sm__throughput.avg.pct_of_peak_sustained_elapsed')}%  "
          f"dramR {g('dram__bytes_read.sum')} dramW {g('dram__bytes_write.sum')}")
    tens = [k for k in h if 'pipe_tensor' in k and 'pct' in k]
    print("   " + "  ".join(f"{k.split('.')[0][-28:]}={d[k]}" for k in tens[:6]))
    st = sorted(((float(d[k]), k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for k in h
                 if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio')), reverse=True)[:5]
    print("   stalls: " + ", ".join(f"{n}={x:.2f}" for x, n in st))
