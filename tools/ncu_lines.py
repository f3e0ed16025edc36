"""Join an `ncu --page source --csv` SASS dump with nvdisasm line info: per-source-line stall samples / instructions.
usage: ncu_lines.py <src.csv> <cubin> <kernel-substring> [topN]"""
import csv
import re
import subprocess
import sys

src_csv, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# walk: find function section for kernel, then lines '//## File "x", line N' and '/*addr*/ instr'
addr2line = {}
cur_line = None
in_fn = False
for ln in dis.split("\n"):
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        in_fn = kern in m.group(1)
        continue
    if not in_fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", ln)
    if m:
        addr2line[int(m.group(1), 16)] = cur_line
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ai, si, ii = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = None
agg = {}
tot_s = tot_i = 0
for r in rows[hi + 1:]:
    if len(r) <= max(si, ii):
        continue
    try:
        a = int(r[ai], 16) if r[ai].startswith("0x") else int(r[ai])
    except ValueError:
        continue
    if base is None:
        base = a
    key = addr2line.get(a - base)
    s = int(r[si] or 0)
    i = int(r[ii] or 0)
    tot_s += s
    tot_i += i
    d = agg.setdefault(key, [0, 0])
    d[0] += s
    d[1] += i
print("total samples %d, warp instructions %d" % (tot_s, tot_i))
srcs = {}
for (k, v) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if k:
        f = k[0]
        if f not in srcs:
            try:
                srcs[f] = open("/root/repo/genomeworks_b200/csrc/" + f).read().split("\n")
            except Exception:
                srcs[f] = []
        if 0 < k[1] <= len(srcs[f]):
            text = srcs[f][k[1] - 1].strip()[:100]
    print("%-22s samp %5.1f%%  inst %5.1f%%  %s" % ("%s:%d" % k if k else "?", 100.0 * v[0] / max(tot_s, 1), 100.0 * v[1] / max(tot_i, 1), text))
