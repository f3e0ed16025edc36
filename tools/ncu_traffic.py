"""Write profiles/r02_traffic.json (what bench.py's roofline.traffic reads) from ncu captures of the shipped kernels.
usage: ncu_traffic.py <key> <report.ncu-rep> <units in that launch> <source note> [<key> <report> <units> <note> ...]
key: c3 | c3_msa | c2 | c4. dram_bytes_per_unit = (dram__bytes_read.sum + dram__bytes_write.sum) / units of the captured launch."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r02_traffic.json")


def scaled(value, unit):
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value) * mult[unit]


def read_bytes(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    rd = scaled(*d["dram__bytes_read.sum"])
    wr = scaled(*d["dram__bytes_write.sum"])
    ms = d.get("gpu__time_duration.sum", ("0", "ms"))
    return rd, wr, d.get("Kernel Name", ("?", ""))[0], ms


def main():
    cur = {}
    if os.path.exists(OUT):
        cur = json.load(open(OUT))
    a = sys.argv[1:]
    for i in range(0, len(a), 4):
        key, rep, units, note = a[i], a[i + 1], int(a[i + 2]), a[i + 3]
        rd, wr, kernel, ms = read_bytes(rep)
        cur[key] = {"dram_bytes_per_unit": (rd + wr) / units, "dram_read_bytes": rd, "dram_write_bytes": wr, "units_in_capture": units,
                    "kernel": kernel, "capture_duration": "%s %s (under ncu, not a bench value)" % ms,
                    "source": "%s (%s)" % (note, os.path.basename(rep))}
    json.dump(cur, open(OUT, "w"), indent=1, sort_keys=True)
    print(json.dumps(cur, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
