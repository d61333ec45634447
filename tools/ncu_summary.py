#!/usr/bin/env python
"""Summarise Nsight Compute artefacts into the small JSON files kept under profiles/.

  python tools/ncu_summary.py full  out.json  name=file.ncu-rep [name=file.ncu-rep ...]
      key metrics (duration, grid, registers, DRAM bytes, tensor-pipe / issue / DRAM utilisation, L2 hit rate) of every
      launch inside `ncu --set full` reports
  python tools/ncu_summary.py shares out.json launches.csv
      per-kernel share of the summed gpu__time_duration of an `ncu --metrics gpu__time_duration.sum --csv` launch list
"""
import collections
import csv
import json
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max.per_second", "smsp__inst_executed.sum",
]


def full(out, pairs):
    res = {}
    for pair in pairs:
        name, path = pair.split("=", 1)
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader([l for l in txt.splitlines() if l.startswith('"')]))
        if len(rows) < 3:
            res[name] = []
            continue
        hdr, units = rows[0], rows[1]
        items = []
        for r in rows[2:]:
            d = {"Kernel Name": r[hdr.index("Kernel Name")]}
            for k in KEEP:
                if k in hdr:
                    i = hdr.index(k)
                    d[k] = f"{r[i]} {units[i]}".strip()
            items.append(d)
        res[name] = items
    json.dump(res, open(out, "w"), indent=1)


def shares(out, path):
    rows = list(csv.reader([l for l in open(path) if l.startswith('"')]))
    hdr = rows[0]
    i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[i_val].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[i_unit], 1e-6)
        name = re.sub(r"\(.*$", "", r[i_name]).replace("void ", "").replace("blim::", "")
        a = agg.setdefault(name, [0.0, 0])
        a[0] += v
        a[1] += 1
    # weight initialisation (torch RNG / casts) and the engine's one-off weight repacking run before the timed step
    setup = lambda k: k.startswith(("native::", "at::")) or "repack_rows" in k or "elementwise" in k
    tot = sum(a[0] for k, a in agg.items() if not setup(k))
    ks = [{"kernel": k, "share_pct": round(100 * a[0] / tot, 3), "ms": round(a[0], 3), "launches": a[1]} for k, a in
          sorted(agg.items(), key=lambda kv: -kv[1][0]) if not setup(k)]
    skipped = [{"kernel": k[:80], "ms": round(a[0], 3), "launches": a[1]} for k, a in agg.items() if setup(k)]
    json.dump({"source": path, "total_ms": tot, "kernels": ks, "setup_kernels_excluded": skipped}, open(out, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "full":
        full(sys.argv[2], sys.argv[3:])
    else:
        shares(sys.argv[2], sys.argv[3])
