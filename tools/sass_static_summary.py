"""Static per-kernel summary of the shipped library (no GPU): registers / stack / shared memory (cuobjdump -res-usage), SASS
instruction count and the counts of a few mnemonics (FP64 arithmetic, global loads / stores, shuffles, atomics, and the bulk
TMA copies UBLKCP with their mbarrier SYNCS) — VERDICT r1 item 7 asked for the UBLKCP / UTMALDG line counts either way.
    python tools/sass_static_summary.py [out.md]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'rda_planner_b200', 'librda_b200.so')
MNEMONICS = ['DFMA', 'DMUL', 'DADD', 'MUFU', 'FFMA', 'LDG', 'STG', 'LDS', 'STS', 'LDL', 'STL', 'SHFL', 'ATOMG', 'RED', 'UBLKCP', 'UTMALDG',
             'SYNCS', 'BAR']


def demangle(names):
    out = subprocess.run(['c++filt'] + names, capture_output=True, text=True).stdout.splitlines()
    clean = []
    for o in out:
        o = o.replace('void ', '').replace('(anonymous namespace)::', '').replace('rda::', '')
        clean.append(re.sub(r'\(.*$', '', o))
    return clean


def main(out=None):
    res = subprocess.run(['cuobjdump', '-res-usage', SO], capture_output=True, text=True, check=True).stdout
    usage = {}
    for m in re.finditer(r'Function (\S+):\n\s+REG:(\d+) STACK:(\d+) SHARED:(\d+)', res):
        usage[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(4)))
    sass = subprocess.run(['cuobjdump', '-sass', SO], capture_output=True, text=True, check=True).stdout.splitlines()
    counts, cur = {}, None
    for ln in sass:
        m = re.search(r'Function : (\S+)', ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', ln)
        if m and cur:
            op = m.group(1)
            counts[cur]['_all'] += 1
            for k in MNEMONICS:
                if op == k or op.startswith(k + '.') or (k in ('LDG', 'STG', 'LDS', 'STS', 'LDL', 'STL') and op.startswith(k)):
                    counts[cur][k] += 1
    names = sorted(counts, key=lambda n: -counts[n]['_all'])
    pretty = dict(zip(names, demangle(names)))
    rows = ['# Static summary of `librda_b200.so` (sm_100a, `tools/sass_static_summary.py`, no GPU)', '',
            'Registers / stack / static shared memory from `cuobjdump -res-usage`; instruction and mnemonic counts from `cuobjdump -sass`',
            '(static: loops count once).  `UBLKCP` = `cp.async.bulk` (1-D bulk TMA copy), `SYNCS` = mbarrier operations; `UTMALDG` (tensor-map',
            'TMA loads) are not used — the staged blocks are contiguous per-instance arrays.', '',
            '| kernel | regs | stack B | static smem B | instructions | ' + ' | '.join(MNEMONICS) + ' |', '|---|---|---|---|---|' + '---|' * len(MNEMONICS)]
    for n in names:
        u = usage.get(n, ('?', '?', '?'))
        c = counts[n]
        rows.append(f"| `{pretty[n]}` | {u[0]} | {u[1]} | {u[2]} | {c['_all']} | " + ' | '.join(str(c[k]) for k in MNEMONICS) + ' |')
    text = '\n'.join(rows) + '\n'
    if out:
        open(out, 'w').write(text)
    print(text)


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else None)
