"""profiles/traffic.json from the ncu launch lists `tools/one_step.py` was captured with (one file per shape):

    python tools/traffic_from_csv.py "C4=gpurun_out/traffic_c4.csv:gpurun_out/traffic_c4.log" "C5=..."

Each csv holds dram__bytes_read.sum / dram__bytes_write.sum / gpu__time_duration.sum per launch of exactly one forward +
backward; the log is the script's stdout (how many launches belong to the forward).  Writes "<shape> fwd" and
"<shape> both" = DRAM bytes of the forward launches / of all launches, and prints the per-launch table.
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(path):
    with open(path) as handle:
        lines = [line for line in handle if line.startswith('"')]
    table = {}
    for row in csv.DictReader(lines):
        entry = table.setdefault(int(row["ID"]), {"kernel": row["Kernel Name"]})
        value = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3,
                 "msecond": 1e3}.get(unit, 1)
        entry[row["Metric Name"]] = value * scale
    return [table[key] for key in sorted(table)]


def main(specs):
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as handle:
            result = json.load(handle)
    except OSError:
        result = {}
    for spec in specs:
        name, files = spec.split("=")
        table_path, log_path = files.split(":")
        with open(log_path) as handle:
            counts = re.search(r"forward_launches (\d+) step_launches (\d+)", handle.read())
        n_forward = int(counts.group(1))
        rows = launches(table_path)
        assert len(rows) == int(counts.group(2)), (len(rows), counts.group(2))
        print("%s: %d launches, %d of them the forward" % (name, len(rows), n_forward))
        for i, row in enumerate(rows):
            print("  %-8s %-90s %9.1f us  read %8.3f GB  write %8.3f GB" % (
                "forward" if i < n_forward else "backward", re.sub(r"\(.*", "", row["kernel"])[:90], row["gpu__time_duration.sum"],
                row["dram__bytes_read.sum"] / 1e9, row["dram__bytes_write.sum"] / 1e9))
        total = [row["dram__bytes_read.sum"] + row["dram__bytes_write.sum"] for row in rows]
        result["%s fwd" % name] = int(sum(total[:n_forward]))
        result["%s both" % name] = int(sum(total))
    with open(path, "w") as handle:
        json.dump(result, handle, indent=1, sort_keys=True)
        handle.write("\n")


if __name__ == "__main__":
    main(sys.argv[1:])
