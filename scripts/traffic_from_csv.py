#!/usr/bin/env python
"""Builds profiles/traffic.json (DRAM bytes per launch, per kernel function of the executor's profiler tags) from an
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` launch list of bench.py.
Only the last COMPLETE step (patch_embed_fwd ... next patch_embed_fwd) is used."""
import csv
import json
import re
import sys
from collections import defaultdict

EPI = {"0": "store", "1": "gelu", "2": "resid", "3": "pixshuf", "4": "split2", "5": "dgelu", "6": "head", "7": "head_bwd",
       "8": "rowscale", "9": "dgelu", "10": "ln_bwd", "11": "store+ln", "12": "resid+ln"}


def tag(name):
    n = re.sub(r"^void |<unnamed>::|\(anonymous namespace\)::", "", name)
    m = re.match(r"gemm_nt_(tc05|mma)_kernel<\(?(?:int\))?(\d+), \(?(?:int\))?(\d+)(?:, \(?(?:int\))?\d+)*>", n)
    if m:
        return f"gemm_nt<{EPI.get(m.group(3) if m.group(1) == 'tc05' else m.group(2), '?')}>"
    for k, t in (("gemm_tn_", "gemm_tn"), ("win_attn_fwd", "win_attn_fwd"), ("win_attn_bwd", "win_attn_bwd"),
                 ("layernorm_fwd", "layernorm_fwd"), ("layernorm_bwd", "layernorm_bwd"), ("patch_embed_fwd", "patch_embed_fwd"),
                 ("patch_embed_bwd", "patch_embed_bwd"), ("pack_weights", "pack_weights"), ("l1_loss", "l1_loss"),
                 ("mlp_block_fwd", "mlp_block_fwd"), ("wmsa_block_fwd", "wmsa_block_fwd"), ("scale_rows", "elementwise"), ("add_inplace", "elementwise"), ("sum_copies", "elementwise"), ("permute_bias", "misc")):
        if k in n:
            return t
    return None


def main(path, out):
    lines = [l for l in open(path) if l.startswith('"')]
    launches = defaultdict(dict)
    order = []
    for r in csv.DictReader(lines):
        i = int(r["ID"])
        if i not in launches:
            order.append(i)
            launches[i]["name"] = r["Kernel Name"]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"].lower()
        mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(unit, 1)
        launches[i][r["Metric Name"]] = v * mult
    seq = [launches[i] for i in order]
    # a step starts at PatchEmbed forward (the weight repack only runs when the parameters changed)
    packs = [i for i, l in enumerate(seq) if "patch_embed_fwd" in l["name"]]
    if len(packs) < 2:
        raise SystemExit("need at least one complete step in the capture")
    step = seq[packs[-2]:packs[-1]]
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for l in step:
        t = tag(l["name"])
        if t is None:
            continue
        agg[t][0] += 1
        agg[t][1] += l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0)
        agg[t][2] += l.get("gpu__time_duration.sum", 0)
    res = {t: {"launches_per_step": n, "dram_bytes_per_launch": round(b / n), "dram_bytes_per_step": round(b),
               "us_per_step_under_ncu": round(us, 1)} for t, (n, b, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])}
    res["_total_dram_bytes_per_step"] = round(sum(v[1] for v in agg.values()))
    res["_source"] = "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none, python bench.py, last complete step"
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1)[:1500])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
