"""Static SASS attribution: instructions of a kernel per source region (no GPU needed).

    python tools/sass_by_source.py [out.md]

Extracts the sm_100a cubin from rda_planner_b200/librda_b200.so (cuobjdump -xelf), disassembles it with
line information (nvdisasm -g; the library is built with -lineinfo) and counts instructions per source
region for the two kernels that matter.  Static counts: for straight-line, fully unrolled code (the lean
cell pass) they are the executed counts of the path taken; for loops (the su-QP) they are per trip."""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'rda_planner_b200', 'librda_b200.so')
CSRC = os.path.join(ROOT, 'rda_planner_b200', 'csrc')


def disassemble():
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(['cuobjdump', '-xelf', 'all', SO], cwd=tmp, check=True, capture_output=True)
        cub = [f for f in os.listdir(tmp) if f.startswith('rda_kernels.') and f.endswith('.cubin')][0]
        return subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cub)], capture_output=True, text=True, check=True).stdout.splitlines()


def attribute(lines, needle):
    start = [i for i, ln in enumerate(lines) if ln.startswith('.text.') and needle in ln][0]
    cur, cnt, ops = None, collections.Counter(), collections.defaultdict(collections.Counter)
    for ln in lines[start + 1:]:
        if ln.lstrip().startswith('.section') or ln.startswith('//-----'):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        mm = re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', ln)
        if mm:
            cnt[cur] += 1
            ops[cur][mm.group(1)] += 1
    return cnt, ops


def line_of(fname, text, nth=0):
    hits = [i + 1 for i, ln in enumerate(open(os.path.join(CSRC, fname)).read().splitlines()) if text in ln]
    return hits[nth]


def region(cnt, ops, fname, a, b):
    tot = sum(v for k, v in cnt.items() if k and k[0] == fname and a <= k[1] < b)
    fp64 = sum(c for k in cnt if k and k[0] == fname and a <= k[1] < b for o, c in ops[k].items() if o in ('DFMA', 'DMUL', 'DADD'))
    return tot, fp64


def main():
    lines = disassemble()
    out = ['# Static SASS attribution by source region (`tools/sass_by_source.py`, nvdisasm -g, no GPU)\n']
    # ---- lean cell pass ----
    cnt, ops = attribute(lines, 'k_cells_fastILi4ELi4')
    f = 'cell_lean.cuh'
    marks = [('robot in the world frame, set-up', line_of(f, 'RDA_HD bool cell_lean')),
             ('disc obstacle branch (not taken for polygons)', line_of(f, 'if (kind == RDA_OBS_CIRCLE)')),
             ('rows -> unit normals, offsets, vertices', line_of(f, '// ---- polygon rows')),
             ('closest pair + separating-axis test (32 point-segment tests)', line_of(f, '// ---- closest pair')),
             ('obstacle-side support vertex and LP-vertex multipliers', line_of(f, '    if (sep) {', 1)),
             ('robot-side support vertex and multipliers', line_of(f, '// ---- robot side')),
             ('z, zeta, coefficients', line_of(f, '// ---- multipliers and updates')),
             ('end', 10 ** 6)]
    out.append(f'## k_cells_fast<4,4>: {sum(cnt.values())} instructions static (ncu: 2 336 executed per cell on the polygon path)\n')
    out.append('| region | instructions |\n|---|---|')
    for (name, a), (_, b) in zip(marks[:-1], marks[1:]):
        out.append(f'| `cell_lean.cuh` {name} | {region(cnt, ops, f, a, b)[0]} |')
    byfile = collections.Counter()
    for k, v in cnt.items():
        byfile[k[0] if k else '?'] += v
    out.append(f"| `rda_hd.h` helpers inlined into the above (clamp / min / max / rsqrt / reciprocal) | {byfile['rda_hd.h']} |")
    out.append(f"| `rda_kernels.cu` wrapper: index arithmetic, loads, stores, residual reductions, worklist, division slow paths | {byfile['rda_kernels.cu']} |")
    out.append(f"| CUDA intrinsics headers (shuffles, atomics) | {sum(v for k, v in byfile.items() if k.endswith('.hpp'))} |\n")
    # ---- su-QP ----
    cnt, ops = attribute(lines, 'k_suIdLi32')
    f = 'su_solver.cuh'
    ric, fwd, sol = line_of(f, 'RDA_HD void su_riccati'), line_of(f, '// forward sweep'), line_of(f, 'RDA_HD int su_solve')
    out.append(f'## k_su<double,32>: {sum(cnt.values())} instructions static\n')
    out.append('| region | instructions | of which DFMA/DMUL/DADD | trips per interior point iteration |\n|---|---|---|---|')
    for name, a, b, trips in (('Riccati backward sweep, one stage (factorising and solve-only variants together)', ric, fwd, 'T stages x 2 sweeps'),
                              ('Riccati forward sweep, one stage', fwd, sol, 'T stages x 2 sweeps'),
                              ('`su_solve`: set-up, gradient/Hessian assembly, step lengths, update (lane = stage)', sol, 10 ** 6, 'N hinges x 5 passes per lane')):
        t, d = region(cnt, ops, f, a, b)
        out.append(f'| {name} | {t} | {d} | {trips} |')
    text = '\n'.join(out) + '\n'
    if len(sys.argv) > 1:
        open(sys.argv[1], 'w').write(text)
    print(text)


if __name__ == '__main__':
    main()
