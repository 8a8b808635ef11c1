#!/usr/bin/env python
"""Per-kernel instruction / DRAM counters per ray from an ncu metrics CSV of tools/render_once.py:

    ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \\
        --clock-control none --csv --log-file ncu.csv python tools/render_once.py --workload W --spp S --out run.json
    tools/ncu_counters.py ncu.csv run.json            -> one JSON object on stdout (merge into profiles/r2_kernel_counters.json)

The CSV holds every launch of the process; render_once.py renders `renders` identical frames, so the totals are
divided by that count.  Units per kernel class: rays of the class (closest-hit rays for k_extend and k_shade's
bounce, occlusion rays for k_shadow)."""
import collections
import csv
import json
import sys


def main():
    rows = [r for r in csv.reader(l for l in open(sys.argv[1], errors="replace") if l.startswith('"'))]
    run = json.load(open(sys.argv[2]))
    hdr = rows[0]
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.defaultdict(lambda: collections.defaultdict(float))
    launches = collections.Counter()
    for r in rows[1:]:
        name = r[ki]
        cls = next((c for c in ("k_extend", "k_shade", "k_shadow", "k_resolve", "k_init") if c + "<" in name or c + "(" in name), None)
        if cls is None:
            cls = "cub_sort" if "cub" in name or "RadixSort" in name or "Onesweep" in name else "other"
        v = float(r[vi].replace(",", "") or 0)
        unit = r[ui]
        if r[mi] == "gpu__time_duration.sum":
            v *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(unit, 1e-6)
            launches[cls] += 1
        elif r[mi].startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        tot[cls][r[mi]] += v
    n = run["renders"]
    units = {"k_extend": run["rays_closest"], "k_shade": run["hits"], "k_shadow": run["rays_shadow"]}
    out = {"workload": run["workload"], "spp": run["spp"], "resolution": [run["width"], run["height"]], "depth": run["depth"],
           "rays_closest": run["rays_closest"], "rays_shadow": run["rays_shadow"], "hits": run["hits"], "samples": run["samples"],
           "note": "ncu --clock-control none, one wavefront pool, production instantiation of the kernels", "kernels": {}}
    for cls, m in tot.items():
        warp_inst, lane_inst = m.get("smsp__inst_executed.sum", 0) / n, m.get("smsp__thread_inst_executed.sum", 0) / n
        e = {"launches": launches[cls] // n, "ms_total": m.get("gpu__time_duration.sum", 0) / n,
             "warp_inst": warp_inst, "lane_inst": lane_inst, "active_lanes_per_inst": lane_inst / warp_inst if warp_inst else None,
             "dram_bytes": (m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)) / n}
        if cls in units and units[cls]:
            e["unit"] = {"k_extend": "closest-hit ray", "k_shade": "shaded hit", "k_shadow": "occlusion ray"}[cls]
            e["lane_inst_per_unit"] = lane_inst / units[cls]
            e["warp_inst_per_unit"] = warp_inst / units[cls]
            e["dram_bytes_per_unit"] = e["dram_bytes"] / units[cls]
            e["dram_bytes_per_launch"] = e["dram_bytes"] / max(1, e["launches"])
        out["kernels"][cls] = e
    print(json.dumps(out))


if __name__ == "__main__":
    main()
