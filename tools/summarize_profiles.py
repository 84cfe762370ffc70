"""Turn the raw ncu outputs under gpurun_out/ into the committed summaries under profiles/.

    python tools/summarize_profiles.py <round-tag>      # e.g. r01
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
out_dir = os.path.join(ROOT, 'profiles')
os.makedirs(out_dir, exist_ok=True)

# ---- launch list (gpu__time_duration per launch) ----
lp = os.path.join(ROOT, 'gpurun_out', f'launches_{tag}.csv')
if os.path.exists(lp):
    rows = list(csv.reader(open(lp)))
    hdr, agg = None, collections.defaultdict(list)
    for r in rows:
        if len(r) > 5 and r[0] == 'ID':
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                agg[d['Kernel Name'].split('(')[0].replace('<unnamed>::', '').replace('void ', '')].append(
                    float(d['Metric Value'].replace(',', '')))
            except ValueError:
                pass
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(out_dir, f'launches_{tag}.md'), 'w') as f:
        f.write(f'# ncu launch list ({tag}): gpu__time_duration.sum per launch, --clock-control none\n\n')
        f.write('command: `ncu --metrics gpu__time_duration.sum --clock-control none -s <warm-up> -c <N> --csv python bench.py ...`\n')
        f.write('(cold-cache, serialised launches: compare SHARES, not absolutes)\n\n')
        f.write('| kernel | launches | mean us | min us | max us | share of step |\n|---|---|---|---|---|---|\n')
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f'| {k} | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {min(v) / 1e3:.1f} | {max(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f}% |\n')
    print('wrote launches summary')

# ---- full capture: selected raw metrics per kernel ----
rp = os.path.join(ROOT, 'gpurun_out', f'prof_{tag}.ncu-rep')
if os.path.exists(rp):
    raw = subprocess.run(['ncu', '-i', rp, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
            'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
            'smsp__inst_executed.sum', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
            'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores']
    traffic = {}
    with open(os.path.join(out_dir, f'ncu_{tag}_summary.md'), 'w') as f:
        f.write(f'# ncu --set full --clock-control none --import-source on ({tag}), one launch per kernel\n\n')
        for r in rows[2:]:
            name = r[idx['Kernel Name']].split('(')[0].replace('<unnamed>::', '').replace('void ', '')
            f.write(f'## {name}\n\n| metric | value | unit |\n|---|---|---|\n')
            for w in want:
                if w in idx:
                    f.write(f'| {w} | {r[idx[w]]} | {units[idx[w]]} |\n')
            f.write('\n')

            def gb(key):
                v, u = float(r[idx[key]].replace(',', '')), units[idx[key]]
                return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}[u]
            key = name.split('<')[0]
            traffic[key] = gb('dram__bytes_read.sum') + gb('dram__bytes_write.sum')
    meta = {}
    mp = os.path.join(ROOT, 'gpurun_out', f'prof_{tag}_meta.json')
    if not os.path.exists(mp):
        mp = os.path.join(ROOT, 'gpurun_out', 'prof_r01_meta.json')      # written by tools/profile_target.py (same target every round)
    if os.path.exists(mp):
        meta = json.load(open(mp))
    traffic['k_cells'] = sum(v for k, v in traffic.items() if k.startswith('k_cells_'))      # all cell passes of one iteration
    traffic['_note'] = 'dram__bytes_read.sum + dram__bytes_write.sum per launch (bytes), ncu --set full, ' + json.dumps(meta)
    json.dump(traffic, open(os.path.join(out_dir, 'roofline_traffic.json'), 'w'), indent=1)
    print('wrote ncu summary + roofline_traffic.json', traffic)
