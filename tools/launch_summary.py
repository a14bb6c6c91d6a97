#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`): launches, total ms, share.
Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes.
Usage: tools/launch_summary.py launches.csv ["header comment"]"""
import collections
import csv
import re
import sys


def main(path, note=""):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    ki, vi, ui = rows[hdr].index("Kernel Name"), rows[hdr].index("Metric Value"), rows[hdr].index("Metric Unit")
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
    acc, cnt = collections.Counter(), collections.Counter()
    for r in rows[hdr + 1:]:
        if len(r) <= vi or r[ui] not in scale:
            continue
        name = re.sub(r"^void |wcx::|\(anonymous namespace\)::", "", r[ki].split("(")[0]).strip()
        name = re.sub(r"\((bool|int)\)", "", name)
        acc[name] += float(r[vi].replace(",", "")) * scale[r[ui]]
        cnt[name] += 1
    total = sum(acc.values())
    if note:
        print("# " + note)
    print("# per-launch times are cold-cache and serialised by ncu: compare shares, not absolutes")
    print("%-62s %8s %12s %7s" % ("kernel", "launches", "total_ms", "share"))
    for name, v in acc.most_common():
        print("%-62s %8d %12.3f %6.1f%%" % (name[:62], cnt[name], v, 100.0 * v / total))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
