#!/usr/bin/env python
"""profiles/r02_dram_traffic.json from the committed ncu CSV captures under profiles/ (targeted-metric
passes: `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,... --clock-control none --csv`).
Bytes per launch, mean over the captured launches of each kernel. bench.py reads the JSON for
roofline.traffic and distances.dram_*.   python tools/make_traffic_json.py"""
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ids = {}
    for r in rows[1:]:
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        ids.setdefault(d["ID"], {"kernel": d["Kernel Name"]})[d["Metric Name"]] = v
    return list(ids.values())


def entry(ls, source, note=None):
    n = len(ls)
    mean = lambda k: sum(l.get(k, 0.0) for l in ls) / n
    e = {"dram_read_bytes": mean("dram__bytes_read.sum"), "dram_write_bytes": mean("dram__bytes_write.sum"),
         "lts_bytes": mean("lts__t_bytes.sum"), "l2_hit_pct": mean("lts__t_sector_hit_rate.pct"),
         "warp_instructions": mean("smsp__inst_executed.sum"), "duration_us_under_ncu": mean("gpu__time_duration.sum") / 1e3,
         "launches_averaged": n, "kernel": ls[0]["kernel"][:70],
         "source": f"profiles/{source} (ncu targeted metrics, --clock-control none, default cache control)"}
    if note:
        e["note"] = note
    return e


def main():
    out = {}
    f = "r02_l2_core_mut_cfg2_selection.csv"
    out["core_mut_kernel"] = entry(launches(os.path.join(PROF, f)), f,
                                   "the bench workload (cfg2: selection + competition): few distinct parents, duplicate reads hit in L2")
    f = "r02_l2_core_mut_rng.csv"
    out["core_mut_kernel_neutral"] = entry(launches(os.path.join(PROF, f)), f,
                                           "same shape, neutral (~630 distinct parents of 1000): the read side in full + the random donor sectors of recombination")
    f = "r02_l2_core_mut_copy_control.csv"
    out["core_mut_kernel_copy_control"] = entry(launches(os.path.join(PROF, f)), f, "no events (core_mu = 0): pure gather-by-parent copy, neutral parents")
    f = "r02_l2_pair.csv"
    by = {}
    for l in launches(os.path.join(PROF, f)):
        by.setdefault(l["kernel"].split("(")[0].split("::")[-1], []).append(l)
    names = {"pair_tile2_kernel": "pair_core_kernel", "core_planes_kernel": "core_planes_kernel", "pair_acc_kernel": "pair_acc_kernel"}
    for k, ls in by.items():
        if k in names:
            out[names[k]] = entry(ls, f)
    json.dump(out, open(os.path.join(PROF, "r02_dram_traffic.json"), "w"), indent=1)
    for k, v in out.items():
        print(k.ljust(30), "read %.1f MB  write %.1f MB  L2 %.2f GB  %.1f us" % (v["dram_read_bytes"] / 1e6, v["dram_write_bytes"] / 1e6, v["lts_bytes"] / 1e9, v["duration_us_under_ncu"]))


if __name__ == "__main__":
    main()
