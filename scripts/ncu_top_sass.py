"""Top stalled SASS instructions of an ncu report: python scripts/ncu_top_sass.py rep.ncu-rep [n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
a, s, e = h.index("Warp Stall Sampling (All Samples)"), h.index("Source"), h.index("Instructions Executed")
data = []
for idx, r in enumerate(rows[hi + 1:]):
    try: data.append((int(r[a]), idx, r[s].strip(), int(r[e])))
    except (ValueError, IndexError): pass
tot = sum(d[0] for d in data)
print(f"{rep}: {tot} samples, {len(data)} instructions")
for d in sorted(data, reverse=True)[:n]:
    print(f"{d[0]:6d} {100.0 * d[0] / tot:5.1f}%  #{d[1]:5d} exec={d[3]:8d}  {d[2][:120]}")
