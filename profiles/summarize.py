#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

  python profiles/summarize.py launches gpurun_out/launches_X.csv  > profiles/X_launches.md
  python profiles/summarize.py full gpurun_out/prof_X.ncu-rep     > profiles/X_full.md
  python profiles/summarize.py traffic gpurun_out/prof_X.ncu-rep PARTICLES > profiles/traffic.json   (read by bench.py)
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_static', 'launch__grid_size',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']


def short(name):
    return re.sub(r'\(.*', '', name).replace('<unnamed>::', '').replace('void ', '')


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    L = [(short(r[ki]), float(r[vi].replace(',', ''))) for r in rows[1:]]
    starts = [i for i, (k, _) in enumerate(L) if k.startswith('k_predict')]
    print("# ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`), %d launches captured\n" % len(L))
    print("Per-launch times are cold-cache and serialised; compare SHARES, not absolutes.\n")
    if len(starts) >= 3:
        s, e = starts[-3], starts[-2]
        tot = sum(v for _, v in L[s:e])
        agg = collections.OrderedDict()
        for k, v in L[s:e]:
            agg.setdefault(k, [0.0, 0])
            agg[k][0] += v
            agg[k][1] += 1
        print("One whole step (%d launches, %.1f us serialised):\n" % (e - s, tot / 1e3))
        print("| kernel | launches | total us | share |\n|---|---|---|---|")
        for k, (v, c) in agg.items():
            print("| %s | %d | %.1f | %.1f%% |" % (k, c, v / 1e3, 100 * v / tot))
    print("\nAll launches (kernel, us):\n")
    print("```")
    for k, v in L:
        print("%-28s %10.1f" % (k, v / 1e3))
    print("```")


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# ncu --set full --clock-control none: %s\n" % path)
    for r in rows[2:]:
        print("## %s\n" % short(r[idx['Kernel Name']]))
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in idx:
                print("| %s | %s | %s |" % (k, r[idx[k]], units[idx[k]]))
        print()


def traffic(path, particles):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the captured launches of each kernel."""
    import json
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    acc = collections.OrderedDict()
    for r in rows[2:]:
        full_name = short(r[idx['Kernel Name']])
        name = re.sub(r'<.*', '', full_name)
        if name == 'k_delta_p' and not full_name.startswith('k_delta_p<0'):
            name = 'k_delta_p_update'          # the last iteration's launch also does update.glsl: kept apart
        b = sum(float(r[idx[k]].replace(',', '')) * scale[units[idx[k]]] for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
        t = float(r[idx['gpu__time_duration.sum']].replace(',', ''))
        acc.setdefault(name, []).append((b, t, units[idx['gpu__time_duration.sum']]))
    res = {"capture": path.replace('gpurun_out/', 'profiles/ summary of '), "particles": int(particles), "kernels": {}}
    for k, v in acc.items():
        res["kernels"][k] = {"dram_bytes_per_launch": sum(x[0] for x in v) / len(v), "launches": len(v),
                             "ncu_duration": sum(x[1] for x in v) / len(v), "ncu_duration_unit": v[0][2]}
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    {'launches': launches, 'full': full, 'traffic': traffic}[sys.argv[1]](*sys.argv[2:])
