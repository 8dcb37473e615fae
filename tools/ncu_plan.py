"""Helper of tools/capture_evidence.sh: reads the launch list of a bench.py run (ncu --metrics gpu__time_duration.sum
--csv) and prints `skip count` for an `ncu -k regex:FILTER --launch-skip skip --launch-count count` capture of ONE
engine step (step index `step`, counted in detect_kernel launches), or, with --summary, the per-kernel time table of
the steps from `step` on of the first engine (shares of the serialised, cold-cache launches)."""
import csv
import re
import sys


def main():
    path, filt, step = sys.argv[1], re.compile(sys.argv[2]), int(sys.argv[3])
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    iu = hdr.index("Metric Unit")
    for r in rd:
        if len(r) > iv and r[im] == "gpu__time_duration.sum":
            v = float(r[iv].replace(",", ""))
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[iu], 1e-3)
            rows.append((r[ik], v))
    names = [n for n, _ in rows]
    det = [i for i, n in enumerate(names) if "detect_kernel" in n]
    # a step starts at the first filtered launch after the previous step's detect... simpler: from the launch after
    # detect #(step-1)'s bookkeeping to detect #step is not a whole step; use pyramid launches as step starts
    starts = [i for i, n in enumerate(names) if "pyr_down_bulk" in n and (i == 0 or "pyr_down" not in names[i - 1])]
    if not starts:
        starts = det
    s0 = starts[min(step, len(starts) - 2)]
    s1 = starts[min(step, len(starts) - 2) + 1]
    if "--summary" in sys.argv:
        # steps [step, step + 6) of the first engine
        e = starts[min(step + 6, len(starts) - 1)]
        agg = {}
        for n, v in rows[s0:e]:
            b = re.sub(r"\(.*", "", n).replace("void ", "").replace("mskf::", "")
            if "synth" in b or "elementwise" in b or "copy" in b.lower():
                continue
            a = agg.setdefault(b, [0, 0.0])
            a[0] += 1
            a[1] += v
        tot = sum(a[1] for a in agg.values())
        nsteps = max(1, min(step + 6, len(starts) - 1) - min(step, len(starts) - 2))
        print(f"# launches {s0}..{e} of the list ({nsteps} engine steps from step {step}); cold-cache serialised times: compare SHARES")
        print("kernel,launches_per_step,us_per_step,share")
        for b, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"{b},{c / nsteps:.1f},{v / nsteps:.1f},{v / tot:.4f}")
        return
    skip = sum(1 for n in names[:s0] if filt.search(n))
    count = sum(1 for n in names[s0:s1] if filt.search(n))
    print(skip, count)


if __name__ == "__main__":
    main()
