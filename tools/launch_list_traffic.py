"""ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) of one bench
step -> per-kernel-class launches / time / DRAM bytes, written as JSON for bench.py's roofline.traffic and as text for
profiles/.   usage: python tools/launch_list_traffic.py profiles/r1_launches_X.csv 256 profiles/r1_conv_traffic.json"""
import io
import json
import re
import sys

import pandas as pd

src, crops, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
lines = [l for l in open(src) if not l.startswith("==")]
d = pd.read_csv(io.StringIO("".join(lines)))
d["Metric Value"] = pd.to_numeric(d["Metric Value"].astype(str).str.replace(",", ""), errors="coerce")
w = d.pivot_table(index=["ID", "Kernel Name"], columns="Metric Name", values="Metric Value", aggfunc="first").reset_index()
unit = d[d["Metric Name"] == "gpu__time_duration.sum"]["Metric Unit"].iloc[0]
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(unit, 1e-6)
bunit = d[d["Metric Name"] == "dram__bytes_read.sum"]["Metric Unit"].iloc[0].lower()
bscale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(bunit, 1.0)


def short(n):
    m = re.search(r"(\w+)(<[^>]*>)?\(", n)
    if not m:
        return n
    return m.group(1) + (re.sub(r"\(\w+\)", "", m.group(2)) if m.group(2) else "")


w["cls"] = w["Kernel Name"].map(short)
w["ms"] = w["gpu__time_duration.sum"] * scale
w["bytes"] = (w["dram__bytes_read.sum"] + w["dram__bytes_write.sum"]) * bscale
g = w.groupby("cls").agg(launches=("ms", "size"), ms=("ms", "sum"), dram_bytes=("bytes", "sum")).sort_values("ms", ascending=False)
g["share_pct"] = 100 * g.ms / g.ms.sum()
conv = g[g.index.str.startswith("conv")]
blob = {"crops_per_step": crops, "source": src + " (ncu, one timed step; cold-cache serialised launches: shares, not absolutes)",
        "conv_dram_bytes_per_step": float(conv.dram_bytes.sum()), "conv_launches": int(conv.launches.sum()),
        "classes": {k: {"launches": int(r.launches), "ms": float(r.ms), "dram_bytes": float(r.dram_bytes), "share_pct": float(r.share_pct)}
                    for k, r in g.iterrows()}}
json.dump(blob, open(out, "w"), indent=1)
print(g.to_string())
