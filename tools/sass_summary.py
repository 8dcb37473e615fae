#!/usr/bin/env python
"""Static SASS instruction counts per kernel of the in-tree library -> profiles/<tag>_sass_summary.txt.

    python tools/sass_summary.py [tag=r02]

DMMA proves the fp64 tensor pipe, UBLKCP the 1-D bulk TMA copies, LDGSTS the cp.async staging, SYNCS the
mbarriers, REDUX / CREDUX the warp reductions (names: /opt/skills/guides/B200_PROFILING.md).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ['DMMA', 'DFMA', 'DADD', 'DMUL', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'LDGSTS', 'SYNCS', 'REDUX', 'CREDUX', 'IMAD', 'LDG', 'STG',
       'LDS', 'STS', 'SHFL', 'BAR', 'MUFU']


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r02'
    so = os.path.join(ROOT, 'msckf_stereo_c_b200', 'libmsckf_b200.so')
    out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True, check=True).stdout
    cur, counts, tot = None, collections.OrderedDict(), {}
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = cur.split('(')[0].replace('void ', '').replace(', ', ' ')
            counts[cur] = collections.Counter()
            tot[cur] = 0
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m and cur:
            tot[cur] += 1
            op = m.group(1).split('.')[0]
            if op in OPS:
                counts[cur][op] += 1
    path = os.path.join(ROOT, 'profiles', tag + '_sass_summary.txt')
    with open(path, 'w') as f:
        f.write('# cuobjdump -sass msckf_stereo_c_b200/libmsckf_b200.so (sm_100a): static instruction counts per kernel\n')
        f.write('# DMMA = fp64 tensor pipe (mma.sync.m8n8k4.f64); UBLKCP = 1-D bulk TMA (cp.async.bulk); LDGSTS = cp.async;\n')
        f.write('# SYNCS = mbarrier; REDUX / CREDUX = warp reduce (add / min-max).  No UTMALDG/UTMASTG: no tensor-map TMA; no tcgen05: the path has no\n')
        f.write('# low-precision GEMM (fp64 has no tcgen05 form).\n')
        f.write('kernel,total,' + ','.join(OPS) + '\n')
        for k, c in counts.items():
            f.write(k + ',' + str(tot[k]) + ',' + ','.join(str(c[o]) for o in OPS) + '\n')
    print(path)


if __name__ == '__main__':
    main()
