"""ncu launch list (csv: gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch) of ONE whole search
step -> per-kernel-family time / DRAM-byte table (stdout) and profiles/r02_step_traffic.json, which bench.py reads for
`roofline.traffic`.   usage: launch_traffic.py <launches.csv> <unrolled|first_order> [out.json]"""
import collections
import csv
import json
import re
import sys

MIXED = ("fwdA", "fwdB", "combine", "node_stats", "bwdB", "bwdA", "wgrad", "SourceGrad", "ArchGrads")
NAMES = {"KFwdA4": "fwdA", "KFwdB4": "fwdB", "KFwdA": "fwdA", "KFwdB": "fwdB", "KCombine": "combine", "KNodeStats": "node_stats",
         "KBwdB2": "bwdB", "KBwdB4": "bwdB", "KBwdA2": "bwdA", "KBwdA4": "bwdA", "KWgrad2": "wgrad", "KWgrad4": "wgrad",
         "KSourceGrad": "SourceGrad", "KArchGrads": "ArchGrads"}


def family(kname):
    m = re.search(r"pcd::(K\w+)", kname)
    if m:
        return NAMES.get(m.group(1), m.group(1)[1:])
    m = re.search(r"(\w+)\s*[<(]", kname)
    return (m.group(1) if m else kname)[:40]


def main():
    path, mode = sys.argv[1], sys.argv[2]
    out = sys.argv[3] if len(sys.argv) > 3 else None
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        rows.append(r)
    per = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])        # launches, us, read, write
    for r in rows:
        k = (r["ID"], r["Kernel Name"])
        v = float(str(r["Metric Value"]).replace(",", ""))
        unit = r["Metric Unit"]
        e = per[k]
        if r["Metric Name"] == "gpu__time_duration.sum":
            e[1] = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
            e[0] = 1
        else:
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            e[2 if "read" in r["Metric Name"] else 3] = v * mult
    fam = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for (_, kn), e in per.items():
        f = fam[family(kn)]
        for i in range(4):
            f[i] += e[i]
    tot_us = sum(f[1] for f in fam.values())
    print(f"{'family':24s} {'launches':>8s} {'ms':>9s} {'share':>7s} {'read MB':>10s} {'write MB':>10s}")
    for k, f in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:24s} {f[0]:8d} {f[1] / 1e3:9.3f} {100 * f[1] / tot_us:6.1f}% {f[2] / 1e6:10.1f} {f[3] / 1e6:10.1f}")
    mixed = [f for k, f in fam.items() if k in MIXED]
    mb = sum(f[2] + f[3] for f in mixed)
    print(f"MixedOp group: {sum(f[1] for f in mixed) / 1e3:.3f} ms (cold-cache, serialised), {mb / 1e9:.3f} GB DRAM traffic; all: {tot_us / 1e3:.3f} ms")
    if out:
        try:
            with open(out) as fh:
                doc = json.load(fh)
        except Exception:
            doc = {}
        doc.update(batch=64, source=path.split("profiles/")[-1] if "profiles/" in path else path)
        doc[mode] = {"mixedop_group_bytes": mb, "mixedop_group_ms_serialised": sum(f[1] for f in mixed) / 1e3,
                     "by_family": {k: {"launches": f[0], "ms": f[1] / 1e3, "dram_read": f[2], "dram_write": f[3]} for k, f in fam.items()}}
        with open(out, "w") as fh:
            json.dump(doc, fh, indent=1)


if __name__ == "__main__":
    main()
