#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of
`bench.py --steps 1 --warmup 1`: per-launch time and DRAM bytes of the LAST step (one step = one embed_rows launch), and
the decoder's totals (to_planar_f16 .. umma_mrf / conv_post).   python tools/parse_launches.py gpurun_out/launches_traffic.csv"""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    d = collections.OrderedDict()
    for row in csv.DictReader(lines):
        e = d.setdefault(row["ID"], {"name": row["Kernel Name"]})
        v = float(row["Metric Value"].replace(",", ""))
        u, m = row["Metric Unit"], row["Metric Name"]
        if m == "gpu__time_duration.sum":
            v = v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(u, 1)
        else:
            v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        e[m] = v
    return list(d.values())


def main():
    L = load(sys.argv[1])
    starts = [i for i, x in enumerate(L) if "embed_rows" in x["name"]]
    step = L[starts[-2]:starts[-1]] if len(starts) >= 2 else L[starts[-1]:]
    in_dec, dec_b, dec_t, dec_n, tot_t, tot_b = False, 0.0, 0.0, 0, 0.0, 0.0
    print("idx,kernel,us,dram_MB,dram_GBps")
    for i, x in enumerate(step):
        t = x["gpu__time_duration.sum"]
        b = x.get("dram__bytes_read.sum", 0.0) + x.get("dram__bytes_write.sum", 0.0)
        nm = re.sub(r"\(.*", "", x["name"]).replace(", ", ";").replace(",", ";")[:56]
        if "to_planar" in nm:
            in_dec = True
        if in_dec:
            dec_b += b; dec_t += t; dec_n += 1
        tot_t += t; tot_b += b
        print("%d,%s,%.1f,%.1f,%.0f" % (i, nm, t, b / 1e6, b / t / 1e3 if t else 0))
        if "conv_post" in nm or "umma_mrf" in nm:       # the decoder's last launch (one-kernel last MRF stage, or conv_post)
            in_dec = False
    print("# step: %d launches, %.1f us serialised, %.2f GB DRAM" % (len(step), tot_t, tot_b / 1e9))
    print("# decoder: %d launches, %.1f us, %.3f GB DRAM (read+write)" % (dec_n, dec_t, dec_b / 1e9))


if __name__ == "__main__":
    main()
