"""DRAM traffic per launch of the roofline kernels from this round's `ncu --set full` captures.

    python tools/ncu_traffic.py RAW_OR_REP [RAW_OR_REP ...] > profiles/ncu_traffic.json

Arguments are .ncu-rep reports (read through `ncu -i REP --page raw --csv`) or raw-page CSV exports of them.  Every
launch is listed (kernel with template arguments, grid, dram__bytes_read.sum + dram__bytes_write.sum, duration under
ncu), and two aggregates are named for bench.py's `roofline.traffic` (no literals in bench.py):
  decode_attn_kernel      mean over the CROSS-attention launches (those that stream the encoder K/V: > 50 MB) at 24
                          decode rows; `decode_attn_kernel_<rows>_rows` for captures at other row counts
  gemm_bf16_2cta_kernel   mean over the four GEMMs of one encoder layer (qkv, out+res, fc1+GELU, fc2+res = the launches
                          bench.py's GEMM probe times), i.e. the GEMM launches between the first and the third LayerNorm
"""
import csv
import io
import json
import os
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "bytes": 1.0}
TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}


def rows_of(path):
    if path.endswith(".ncu-rep"):
        text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    else:
        text = open(path).read()
    start = text.find('"ID"')
    rows = list(csv.reader(io.StringIO(text[start:])))
    return rows[0], rows[1], rows[2:]


def num(s):
    return float(s.replace(",", ""))


def main(paths):
    launches = []
    for path in paths:
        header, units, rows = rows_of(path)
        col = {h: i for i, h in enumerate(header)}
        kn, rd, wr = col["Kernel Name"], col["dram__bytes_read.sum"], col["dram__bytes_write.sum"]
        du, gs = col.get("gpu__time_duration.sum"), col.get("launch__grid_size")
        for r in rows:
            name = r[kn].split("(")[0].strip()
            for ns in ("tw::", "gemm::", "attn::", "dec::", "ln::", "logmel::", "void "):
                name = name.replace(ns, "")
            launches.append({"kernel": name, "grid": int(num(r[gs])) if gs is not None else None,
                             "dram_bytes": num(r[rd]) * UNIT.get(units[rd], 1.0) + num(r[wr]) * UNIT.get(units[wr], 1.0),
                             "us_under_ncu": round(num(r[du]) * TIME.get(units[du], 1.0), 2) if du is not None else None,
                             "report": os.path.basename(path)})
    res = {}
    # cross-attention launches by decode rows (grid = splits x heads x rows = 80 CTAs per row at 20 heads x 4 splits):
    # `decode_attn_kernel` stays the 24-row figure, wider captures (tools/ncu_target.py 96 ...) add `..._<rows>_rows`
    by_rows = {}
    for l in launches:
        if l["kernel"].startswith("decode_attn_kernel") and l["dram_bytes"] > 50e6 and l["grid"]:
            by_rows.setdefault(l["grid"] // 80, []).append(l["dram_bytes"])
    for rows, vals in sorted(by_rows.items()):
        res["decode_attn_kernel" if rows == 24 else "decode_attn_kernel_%d_rows" % rows] = sum(vals) / len(vals)
    ln_seen, layer = 0, []
    for l in launches:
        if l["kernel"].startswith("layernorm_kernel"):
            ln_seen += 1
        elif l["kernel"].startswith("gemm_bf16") and 1 <= ln_seen <= 2:
            layer.append(l["dram_bytes"])
    if len(layer) == 4:
        res["gemm_bf16_2cta_kernel"] = sum(layer) / 4
        res["gemm_layer_shapes"] = dict(zip(("qkv", "out+res", "fc1+gelu", "fc2+res"), layer))
    res["source"] = ("ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum per launch, "
                     "large-v3-turbo, tools/ncu_target.py at 24 decode rows (and at the row count a key names): " + ", ".join(os.path.basename(p) for p in paths))
    res["launches"] = launches
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1:])
