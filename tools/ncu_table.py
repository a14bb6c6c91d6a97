#!/usr/bin/env python
"""One line per profiled launch from the text written by tools/ncu_summary.py: duration, DRAM bytes and GB/s
(dram__bytes_read + dram__bytes_write over gpu__time_duration), fraction of the measured HBM peak, L2 / SM / issue
utilisation, registers.  Usage: tools/ncu_table.py summary.txt [hbm_peak_gbs]"""
import json
import os
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6,
        "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def val(d, key):
    if key not in d:
        return None
    v, u = d[key]
    return float(v.replace(",", "")) * UNIT.get(u, 1.0)


def main(path, peak):
    blocks = open(path).read().split("kernel:")[1:]
    print("%-34s %9s %9s %9s %6s %6s %6s %6s %5s" % ("kernel", "time_us", "dram_MB", "GB/s", "%HBM", "L2%", "SM%", "issue%", "regs"))
    for b in blocks:
        lines = b.strip().splitlines()
        name = lines[0].split("(")[0].split("::")[-1].strip()[:34]
        d = {}
        for l in lines[1:]:
            p = l.split()
            if len(p) >= 2:
                d[p[0]] = (p[1], p[2] if len(p) > 2 else "")
        t = val(d, "gpu__time_duration.sum")
        by = (val(d, "dram__bytes_read.sum") or 0.0) + (val(d, "dram__bytes_write.sum") or 0.0)
        gbs = by / t / 1e9 if t else 0.0
        pct = lambda k: ("%.1f" % float(d[k][0])) if k in d else "-"
        print("%-34s %9.1f %9.1f %9.0f %6.1f %6s %6s %6s %5s" % (
            name, t * 1e6, by / 1e6, gbs, 100.0 * gbs / peak, pct("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
            pct("sm__throughput.avg.pct_of_peak_sustained_elapsed"), pct("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            d.get("launch__registers_per_thread", ("-",))[0]))


if __name__ == "__main__":
    peak = float(sys.argv[2]) if len(sys.argv) > 2 else None
    if peak is None:
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
        peak = json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6550.0
    main(sys.argv[1], peak)
