#!/usr/bin/env python
"""Warp instructions per CUDA source line of one kernel of an ncu report (needs -lineinfo and --import-source on).

    python tools/ncu_lines.py report.ncu-rep kernel_name [launch_index=0] [min_share=0.003] [function_substring]

kernel_name is ncu's base name ("be_gemm_kernel"); function_substring picks a template instance ("<(int)3>").
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    min_share = float(sys.argv[4]) if len(sys.argv) > 4 else 0.003
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', kern],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # every launch of the kernel is one block that starts with a "File Path" row
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'File Path']
    if len(sys.argv) > 5:
        starts = [i for i in starts if sys.argv[5] in rows[i + 1][1]]
    starts = [i for i in starts if len(rows) > i + 3 and len(rows[i + 2]) > 5]  # blocks that carry metrics
    if not starts:
        print('no source page for', kern)
        return
    b = starts[min(which, len(starts) - 1)]
    e = starts[which + 1] if which + 1 < len(starts) else len(rows)
    hdr = rows[b + 2]
    ii, ti = hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
    si = hdr.index('# Samples')
    agg, tot, stot = {}, 0, 0
    for r in rows[b + 3:e]:
        if len(r) <= ii or r[2] != '-':
            continue
        try:
            v, tv, sm = int(r[ii]), int(r[ti]), int(r[si])
        except ValueError:
            continue
        agg[int(r[0])] = (v, tv, sm, r[1])
        tot += v
        stot += sm
    print('launches in report: %d; launch %d: %d warp instructions, %d samples' % (len(starts), which, tot, stot))
    for ln in sorted(agg):
        v, tv, sm, s = agg[ln]
        if v > tot * min_share or sm > stot * min_share:
            print('%5d %6.2f%% inst  %6.2f%% samples  lanes %4.1f  %s' % (ln, 100.0 * v / tot, 100.0 * sm / max(stot, 1), tv / max(v, 1), s[:100]))


if __name__ == '__main__':
    main()
