#!/usr/bin/env python
"""profiles/r02_dram_traffic.json from the committed ncu CSV launch captures (targeted-metric passes,
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,... --csv`): bytes per launch, mean over the
captured launches of each kernel. bench.py reads the JSON for roofline.traffic / distances.dram_*.

    python tools/make_traffic_json.py profiles/r02_l2_core_mut_rng.csv:core_mut_kernel \
                                      profiles/r02_l2_pair_core.csv:pair_core_kernel
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def per_launch(path):
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


def main():
    out_path = os.path.join(ROOT, "profiles", "r02_dram_traffic.json")
    out = json.load(open(out_path)) if os.path.exists(out_path) else {}
    for spec in sys.argv[1:]:
        path, key = spec.split(":")
        launches = per_launch(path)
        n = len(launches)
        # a pass made of several launches (column chunks): sum over the launches of one pass when asked with key+N
        per = 1
        if "+" in key:
            key, per = key.split("+")
            per = int(per)
        rd = sum(l["dram__bytes_read.sum"] for l in launches) / n * per
        wr = sum(l["dram__bytes_write.sum"] for l in launches) / n * per
        out[key] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "launches_averaged": n, "launches_per_pass": per,
                    "duration_us_under_ncu": sum(l.get("gpu__time_duration.sum", 0.0) for l in launches) / n / 1e3 * per,
                    "kernel": launches[0]["kernel"],
                    "source": f"{os.path.relpath(path, ROOT)} (ncu targeted metrics, --clock-control none, default cache control)"}
    json.dump(out, open(out_path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
