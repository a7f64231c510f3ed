#!/usr/bin/env python
"""Record the DRAM traffic of a workload's kernel from an `ncu --set full` capture into profiles/traffic.json,
together with what ties the number to the code: kernel name, capture file, git commit, hash of the kernel sources.

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > profiles/r2_ncu_full_<tag>.csv
    python tools/record_traffic.py <workload name> profiles/r2_ncu_full_<tag>.csv

bench.py reports `roofline.traffic` from this file and nulls an entry whose source hash no longer matches."""
import csv
import datetime
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    workload, path = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    vals, names = [], set()
    for r in data:
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[col[k]].replace(",", "")) * UNIT[units[col[k]]]
        vals.append(tot)
        names.add(r[col["Kernel Name"]])
    import bench
    entry = {
        "traffic": int(round(sum(vals) / len(vals))), "launches_averaged": len(vals), "kernel": sorted(names)[0] if len(names) == 1 else sorted(names),
        "capture": os.path.relpath(path, ROOT),
        "git": subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip(),
        "when": datetime.datetime.now(datetime.timezone.utc).strftime("%Y-%m-%dT%H:%MZ"),
        "sources_sha256": bench.kernel_sources_sha256(),
    }
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        db = json.load(open(p))
    except Exception:
        db = {}
    db["_comment"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch from ncu --set full, written by tools/record_traffic.py; "
                      "bench.py nulls an entry whose sources_sha256 differs from the kernel sources in the tree")
    db[workload] = entry
    json.dump(db, open(p, "w"), indent=1)
    print(workload, entry)


if __name__ == "__main__":
    main()
