"""DRAM traffic per launch of the roofline kernels from an `ncu --set full` report of this round.

    python tools/ncu_traffic.py gpurun_out/r2a_target.ncu-rep [more.ncu-rep ...] > profiles/ncu_traffic.json

Reads the raw page of every report (`ncu -i REP --page raw --csv`), sums dram__bytes_read.sum + dram__bytes_write.sum per
kernel launch and averages over the launches of each kernel (base name before '<' / '(').  bench.py reads the result
for `roofline.traffic` (no literals in bench.py)."""
import csv
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "bytes": 1.0}


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    start = out.find('"ID"')
    rows = list(csv.reader(io.StringIO(out[start:])))
    header, units = rows[0], rows[1]
    return header, units, rows[2:]


def main(reps):
    acc = defaultdict(list)
    dur = defaultdict(list)
    for rep in reps:
        header, units, rows = raw_rows(rep)
        col = {h: i for i, h in enumerate(header)}
        kn = col.get("Kernel Name")
        rd, wr, du = col.get("dram__bytes_read.sum"), col.get("dram__bytes_write.sum"), col.get("gpu__time_duration.sum")
        if kn is None or rd is None or wr is None:
            continue
        for r in rows:
            name = r[kn].split("<")[0].split("(")[0].strip()
            name = name.split("::")[-1]
            try:
                b = float(r[rd].replace(",", "")) * UNIT.get(units[rd], 1.0) + float(r[wr].replace(",", "")) * UNIT.get(units[wr], 1.0)
            except ValueError:
                continue
            acc[name].append(b)
            if du is not None:
                try:
                    dur[name].append(float(r[du].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[du], 1.0))
                except ValueError:
                    pass
    res = {k: sum(v) / len(v) for k, v in acc.items()}
    res["launches"] = {k: len(v) for k, v in acc.items()}
    res["us_under_ncu"] = {k: round(sum(v) / len(v), 2) for k, v in dur.items()}
    res["source"] = "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, mean per launch: " + ", ".join(os.path.basename(r) for r in reps)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1:])
