"""Print the headline fields of bench JSON lines. usage: show_bench.py <file> [...]"""
import json
import sys

for f in sys.argv[1:]:
    for line in open(f):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        r = d.get("roofline") or {}
        print(f, "|", d.get("impl", "ours"), d.get("metric"), "value %.4g %s" % (d.get("value", 0), d.get("unit")), "| ms/step %.4g" % d.get("ms_per_step", 0),
              "| e2e", (d.get("e2e") or {}).get("value"), "| frac", r.get("frac"), "| traffic", r.get("traffic"), "| cpu", (d.get("cpu_baseline") or {}).get("value"),
              "| gpu_ref", (d.get("gpu_reference") or {}), "| launches", d.get("gpu_launches"), "| clocks", d.get("clocks"))
