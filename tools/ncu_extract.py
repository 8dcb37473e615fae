"""Turns an `ncu --set full` report of bench.py into the tracked evidence under profiles/:

    python tools/ncu_extract.py <report.ncu-rep> <streams in the capture> <out prefix, e.g. profiles/r02> ["header note"]

  <prefix>_ncu_full.csv   one row per profiled launch: kernel class, grid, registers, duration, DRAM bytes and
                          throughput, SM / issue / warp activity, fp64 and DMMA pipe activity, warp instructions
  profiles/traffic.json   dram__bytes_read.sum + dram__bytes_write.sum per stream per launch (largest launch of a class)
  profiles/inst.json      smsp__inst_executed.sum (warp instructions) per stream per launch, same choice; bench.py
                          divides it by the measured launch time for the issue-slot roofline of the issue-bound kernels

Runs here (CPU): `ncu -i` only reads the report.  Kernel launches map to bench.py's kernel classes by name, template
argument and order inside a step (klt: temporal, stereo, new; be_gram: lost-feature update, prune update)."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

COLS = [
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
    ("gpu__time_duration.sum", "time_us"), ("dram__bytes_read.sum", "dram_read_MB"), ("dram__bytes_write.sum", "dram_write_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_inst_pct"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("smsp__pipe_tensor_subpipe_dmma_cycles_active.avg", "dmma_cycles_active"),
    ("sm__cycles_active.avg", "sm_cycles_active"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu.sum", "smem_bank_conflicts"),
]


def find(hdr, name):
    for i, h in enumerate(hdr):
        if h == name or h.endswith("." + name):
            return i
    return -1


def to_unit(v, unit, want):
    """ncu prints byte counts in a per-row unit (byte, Kbyte, Mbyte, Gbyte) and times in ns/us/ms."""
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3,
             "nsecond": 1e-3, "second": 1e6}
    if want == "MB":
        return x * scale.get(unit, 1.0) / 1e6
    if want == "us":
        return x * scale.get(unit, 1.0)
    return x


def classify(names):
    """Kernel name list (launch order) -> class per launch."""
    out, klt_i, gram_seen = [], 0, {}
    for n, grid in names:
        base = n.split("(")[0]
        c = base
        if "pyr_down_bulk" in base:
            c = "pyr_down_l1"
        elif "pyr_down" in base or "pyr_tail" in base:
            c = "pyr_down_ln"
        elif "klt" in base:
            c = ("klt_temporal", "klt_stereo", "klt_new")[klt_i % 3]
            klt_i += 1
        elif "detect_kernel" in base:
            c = "detect"
        elif "be_gemm_kernel" in base:
            m = re.search(r"<\(?(?:int\))?(\d)>", base)
            c = "be_gemm_" + ("pht", "s", "w", "pupd")[int(m.group(1))] if m else "be_gemm"
        elif "be_gram" in base:
            c = "be_gram" if grid > min(g for nn, g in names if "be_gram" in nn) else "be_gram_prune"
        elif "be_feature_jac_prune" in base:
            c = "be_feature_jac_prune"
        else:
            c = re.sub(r"_kernel.*", "", base.replace("void ", "").replace("mskf::", ""))
            if c.startswith("fe_"):
                c = "fe_bookkeeping"
        out.append(c)
    return out


def main():
    rep, streams, prefix = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    note = sys.argv[4] if len(sys.argv) > 4 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ik = hdr.index("Kernel Name")
    idx = [(find(hdr, m), short, m) for m, short in COLS]
    names = []
    ig = find(hdr, "launch__grid_size")
    for r in data:
        names.append((r[ik].replace("mskf::", ""), int(float(r[ig].replace(",", "")))))
    classes = classify(names)
    table = []
    for r, (n, _), c in zip(data, names, classes):
        ent = {"class": c, "kernel": n.split("(")[0].replace("void ", "").replace(", ", " ")}
        for i, short, m in idx:
            if i < 0:
                ent[short] = None
                continue
            want = "MB" if short.endswith("_MB") else "us" if short == "time_us" else ""
            ent[short] = to_unit(r[i], units[i], want)
        table.append(ent)
    out_csv = prefix + "_ncu_full.csv"
    with open(out_csv, "w") as f:
        if note:
            f.write("# " + note + "\n")
        f.write("# columns: " + ", ".join(f"{s} = {m}" for m, s in COLS) + "\n")
        keys = ["class", "kernel"] + [s for _, s in COLS]
        f.write(",".join(keys) + "\n")
        for e in table:
            f.write(",".join("" if e[k] is None else (f"{e[k]:.4g}" if isinstance(e[k], float) else str(e[k])) for k in keys) + "\n")
    traffic, inst = {}, {}
    for e in table:
        if e["dram_read_MB"] is None or e["warp_inst"] is None:
            continue
        b = (e["dram_read_MB"] + e["dram_write_MB"]) * 1e6 / streams
        if b > traffic.get(e["class"], -1):
            traffic[e["class"]] = round(b)
        i = e["warp_inst"] / streams
        if i > inst.get(e["class"], -1):
            inst[e["class"]] = round(i)
    src = f"{os.path.relpath(out_csv, ROOT)} (ncu --set full, {streams} streams in the launch; largest launch of each class)"
    for fn, d, what in (("traffic.json", traffic, "dram__bytes_read.sum + dram__bytes_write.sum"), ("inst.json", inst, "smsp__inst_executed.sum (warp instructions)")):
        d = dict(sorted(d.items()))
        d["_comment"] = f"{what} per stream per launch, from {src}; bench.py multiplies by its stream count"
        with open(os.path.join(ROOT, "profiles", fn), "w") as f:
            json.dump(d, f, indent=2)
            f.write("\n")
    print(f"{len(table)} launches -> {out_csv}, profiles/traffic.json, profiles/inst.json")


if __name__ == "__main__":
    main()
