#!/usr/bin/env python
"""Turns `ncu -i X.ncu-rep --page raw --csv` pages into the per-kernel summaries under profiles/ and
profiles/counters.json (read by bench.py: DRAM traffic and warp instructions per launch).
    python profiles/summarize_ncu.py <tag>=<raw.csv>:<envs per launch> ...   e.g. 2v2=profiles/r02_k_step_2v2_raw.csv:16384"""
import csv
import json
import os
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.max', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum',
        'smsp__sass_inst_executed_op_shared_ld.sum', 'smsp__sass_inst_executed_op_shared_st.sum', 'smsp__sass_inst_executed_op_global_ld.sum',
        'smsp__sass_inst_executed_op_global_st.sum']
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3}


def read(path):
    rows = list(csv.reader(open(path, errors='replace')))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    hdr, units = rows[hi], rows[hi + 1]
    out = []
    for r in rows[hi + 2:]:
        if len(r) < len(hdr):
            continue
        d = {}
        for k, u, v in zip(hdr, units, r):
            if k in KEEP:
                try:
                    d[k] = float(v.replace(',', '')) * SCALE.get(u, 1.0)
                except ValueError:
                    pass
        d['kernel'] = r[hdr.index('Kernel Name')]
        out.append(d)
    return out


def main():
    root = os.path.dirname(os.path.abspath(__file__))
    cpath = os.path.join(root, 'counters.json')
    counters = json.load(open(cpath)) if os.path.exists(cpath) else {}
    for spec in sys.argv[1:]:
        tag, rest = spec.split('=')
        path, envs = rest.rsplit(':', 1)
        L = read(path)
        mean = {k: sum(d[k] for d in L if k in d) / max(1, sum(1 for d in L if k in d)) for k in KEEP}
        s = {'source': os.path.basename(path), 'kernel': L[0]['kernel'], 'launches_captured': len(L), 'envs_per_launch': int(envs),
             'units': 'durations us, bytes, ratios as ncu prints them', 'mean': mean}
        json.dump(s, open(os.path.join(root, os.path.basename(path).replace('_raw.csv', '_summary.json')), 'w'), indent=1)
        counters[tag] = {'source': 'profiles/' + os.path.basename(path), 'kernel': L[0]['kernel'], 'envs_per_launch': int(envs),
                         'dram_bytes_per_launch': mean['dram__bytes_read.sum'] + mean['dram__bytes_write.sum'],
                         'warp_inst_per_launch': mean['smsp__inst_executed.sum'],
                         'active_lanes_per_inst': mean['smsp__thread_inst_executed_per_inst_executed.ratio'],
                         'issue_slots_busy_pct': mean['smsp__issue_active.avg.pct_of_peak_sustained_active'],
                         'duration_us_under_ncu': mean['gpu__time_duration.sum']}
        print(tag, json.dumps(counters[tag]))
        for k in KEEP:
            if 'stalled' in k or k in ('smsp__average_warp_latency_per_inst_issued.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size'):
                print('   %-90s %.3f' % (k, mean[k]))
    json.dump(counters, open(cpath, 'w'), indent=1)


if __name__ == '__main__':
    main()
