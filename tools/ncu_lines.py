"""Attribute the warp-stall samples of one kernel in an ncu report to SOURCE LINES (no GPU needed).

    python tools/ncu_lines.py gpurun_out/prof_r02.ncu-rep k_suIdLi32 "k_su<double" profiles/ncu_r02_ksu_lines.md

ncu's source page is exported per SASS instruction (--page source --csv); the line table comes from nvdisasm -g of
the cubin inside rda_planner_b200/librda_b200.so, which must be the build that was profiled (same sources)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, mangled, pretty, out = sys.argv[1:5]
page = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{pretty.split("<")[0]}'],
                      capture_output=True, text=True).stdout
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(['cuobjdump', '-xelf', 'all', os.environ.get('RDA_LINES_SO', os.path.join(ROOT, 'rda_planner_b200', 'librda_b200.so'))], cwd=tmp, check=True,
                   capture_output=True)
    cub = [f for f in os.listdir(tmp) if f.startswith('rda_kernels.') and f.endswith('.cubin')][0]
    dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(dis) if l.startswith('.text.') and mangled in l][0]
cur, off2line = None, {}
for l in dis[start + 1:]:
    if l.startswith('.text.'):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m:
        off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(io.StringIO(page)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]
si, ii = hdr.index('# Samples'), hdr.index('Instructions Executed')
base, agg, tot, ninst = None, collections.defaultdict(lambda: [0, 0]), 0, 0
for r in rows[hi + 1:]:
    try:
        a, s, n = int(r[0], 16), int(r[si]), int(r[ii])
    except (ValueError, IndexError):
        continue
    if base is None:
        base = a
    ln = off2line.get(a - base)
    agg[ln][0] += s
    agg[ln][1] += n
    tot += s
    ninst += n
src = {}


def text(f, n):
    p = os.path.join(ROOT, 'rda_planner_b200', 'csrc', f)
    if f not in src and os.path.exists(p):
        src[f] = open(p).read().splitlines()
    return src[f][n - 1].strip()[:100] if f in src and n - 1 < len(src[f]) else ''


with open(out, 'w') as f:
    f.write(f'# {pretty}: warp-stall samples per source line ({os.path.basename(rep)}; {tot} samples, {ninst} warp instructions)\n\n')
    f.write('| samples % | warp instructions | line | source |\n|---|---|---|---|\n')
    for k, (s, n) in sorted(agg.items(), key=lambda x: -x[1][0])[:40]:
        if k:
            f.write(f'| {100 * s / tot:.2f} | {n} | {k[0]}:{k[1]} | `{text(*k)}` |\n')
print('wrote', out)
