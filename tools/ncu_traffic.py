"""profiles/ncu_traffic.json from `ncu -i <rep> --page raw --csv` exports: DRAM bytes per launch
of the kernel bench.py times (roofline.traffic).

    python tools/ncu_traffic.py c3_msmarco_doc_maxp=profiles/r2_score_tma_full_raw.csv [workload=csv ...]
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        out = json.load(open(path))
    except Exception:
        out = {}
    sources = []
    for arg in sys.argv[1:]:
        workload, raw = arg.split("=", 1)
        rows = list(csv.reader(open(raw)))
        head, units, vals = rows[0], rows[1], rows[2]
        col = {h: i for i, h in enumerate(head)}
        total = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            total += float(vals[col[name]].replace(",", "")) * UNIT[units[col[name]]]
        out[workload] = total
        sources.append(f"{workload}: {raw}, dram__bytes_read.sum + dram__bytes_write.sum of one launch of "
                       f"`{vals[col['Kernel Name']]}` ({vals[col['gpu__time_duration.sum']]} "
                       f"{units[col['gpu__time_duration.sum']]}), ncu --set full --clock-control none")
    keep = [s for s in out.get("_source", "").split(" | ") if s and not any(s.startswith(a.split("=")[0] + ":") for a in sys.argv[1:])]
    out["_source"] = " | ".join(keep + sources)
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
