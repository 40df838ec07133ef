"""Summarise `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none
--clock-control none --csv` logs of bench.py runs into profiles/traffic.json: DRAM bytes per k_step launch, averaged over
the captured consecutive launches (caches are NOT flushed between them, so the write-back of earlier launches' dirty
lines is part of each launch's figure: the steady state of the timed loop, which a single cold launch is not).
usage: python scripts/ncu_traffic.py <tag> <env_id>|<n_envs>=<csv> ..."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def parse(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        rows.append(r)
    per = {}
    for r in rows:
        if "k_step" not in r.get("Kernel Name", ""):
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"].lower()
        scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "nsecond": 1e-3, "msecond": 1e3}.get(unit, 1.0)
        per.setdefault(r["ID"], {})[r["Metric Name"]] = val * scale
    return list(per.values())


def main():
    tag = sys.argv[1]
    out_path = os.path.join(ROOT, "profiles", "traffic.json")
    table = json.load(open(out_path)) if os.path.exists(out_path) else {}
    for arg in sys.argv[2:]:
        key, path = arg.split("=", 1)
        launches = parse(path)
        if len(launches) < 4:
            print("too few k_step launches in", path, len(launches))
            continue
        launches = launches[2:]  # the first launches still see a cold L2
        rd = sum(x["dram__bytes_read.sum"] for x in launches) / len(launches)
        wr = sum(x["dram__bytes_write.sum"] for x in launches) / len(launches)
        us = sum(x["gpu__time_duration.sum"] for x in launches) / len(launches)
        table[key] = {"bytes_per_launch": rd + wr, "read": rd, "write": wr, "launches_averaged": len(launches), "us_per_launch_under_ncu": us,
                      "source": f"profiles/{tag}_traffic_{key.split('|')[0]}.csv: ncu dram__bytes_read.sum + dram__bytes_write.sum per k_step launch, "
                                f"mean of {len(launches)} consecutive launches, --cache-control none (steady state: write-back included)"}
        print(key, table[key])
    json.dump(table, open(out_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
