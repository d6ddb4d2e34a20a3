"""Turn ncu output brought back from the GPU box into the text summaries kept under profiles/.

  python tools/ncu_summary.py raw   <file.ncu-rep> "<header line>"   > profiles/xxx_ncu.txt
  python tools/ncu_summary.py list  <launches.csv> "<header line>"   > profiles/xxx_launches_summary.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(dram__bytes_(read|write)\.sum(\.per_second|\.pct_of_peak_sustained_elapsed)?|"
    r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|gpu__time_duration\.sum|"
    r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|"
    r"l1tex__data_pipe_lsu_wavefronts(_mem_shared)?\.(sum|avg\.pct_of_peak_sustained_elapsed)|"
    r"launch__(block_size|grid_size|registers_per_thread|occupancy_limit_\w+)|"
    r"sm__cycles_(active|elapsed)\.avg|sm__inst_executed\.avg\.per_cycle_(elapsed|active)|"
    r"sm__inst_executed_pipe_\w+\.avg\.pct_of_peak_sustained_active|"
    r"sm__warps_active\.avg\.pct_of_peak_sustained_active|"
    r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|"
    r"smsp__inst_executed\.sum|smsp__issue_active\.avg\.pct_of_peak_sustained_active)$")


def raw(rep, header):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    names, units = rows[0], rows[1]
    print(header)
    for row in rows[2:]:
        kn = row[names.index("Kernel Name")]
        print("kernel:", kn)
        for n, u, v in sorted(zip(names, units, row)):
            if KEEP.match(n) and "not_issued" not in n:
                print("%-96s%s%s" % (n, v, ("  [%s]" % u) if u else ""))


def launches(path, header):
    rows = list(csv.reader(open(path)))
    start = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[start]
    iK, iV = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= iV:
            continue
        k = re.sub(r"\(.*", "", r[iK]).replace("void ", "")
        ns = float(r[iV].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
    unit = rows[start + 1][h.index("Metric Unit")] if "Metric Unit" in h else "ns"
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1e-3)
    tot = sum(a[1] for a in agg.values())
    print(header)
    print("(cold-cache, serialised launches: compare SHARES, not absolutes)\n")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-90s launches %3d  total %10.1f us  avg %8.1f us  share %5.1f%%" % (
            k[:90], n, t * scale, t * scale / n, 100 * t / tot))


if __name__ == "__main__":
    {"raw": raw, "list": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
