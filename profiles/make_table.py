#!/usr/bin/env python
"""Regenerates the measured table of DESIGN.md section 8 (between the MEASURED_TABLE markers) from
profiles/bench_r02_<workload>.json and profiles/counters.json:  python profiles/make_table.py"""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W = ['2v2', '1v1_heal_only', 'ffa', 'ffa_lidar']


def load(w):
    return json.loads(open(os.path.join(ROOT, 'profiles', f'bench_r02_{w}.json')).read().strip().splitlines()[-1])


def main():
    out = ['<!-- MEASURED_TABLE begin (profiles/make_table.py) -->',
           '| workload (`bench.py --workload`) | envs/GPU | ms/step | agent-steps/s (device) | `e2e` (rewards+dones) | `concurrent_groups` device / e2e | `e2e_obs` sync / pipelined | k_step / k_obs2 / k_lidar ms | k_step GB/s (frac of 6 534) | C port, 16 threads | GPU/CPU |',
           '|---|---|---|---|---|---|---|---|---|---|---|']
    for w in W:
        d = load(w); r = d['roofline']; k = r['kernel_ms_all']
        cpu = d.get('cpu_baseline', {}).get('value')
        out.append(f"| {w} | {d['config']['envs_per_gpu']} | {d['ms_per_step']:.4f} | {d['value']:.3e} | {d['e2e']['value']:.3e} | {d['concurrent_groups']['value']:.3e} / {d['concurrent_groups']['e2e']['value']:.3e} | "
                   f"{d['e2e_obs']['value']:.3e} / {d['e2e_obs']['pipelined']['value']:.3e} | {k['k_step']:.4f} / {k['k_obs']:.4f} / {k['k_lidar']:.4f} | "
                   f"{r['achieved']:.0f} ({r['frac']:.4f}) | {cpu:.3e} | {d['value'] / cpu:.0f}x |" if cpu else '| %s | n/a |' % w)
    d = load('2v2')
    ph = d.get('ms_per_step_by_episode_phase') or {}
    if ph:
        out.append('')
        out.append('2v2 by episode phase (all envs reset together, then timed windows): ' + ', '.join(f'{k}: {v:.3f} ms' for k, v in ph.items()) + '.')
    rp = d.get('cpu_baseline_ref_python')
    if rp:
        out.append(f"Reference Python on the shim (`cpu_baseline_ref_python`, build container, {rp['cores']} cores): {rp['value']:.3e} agent-steps/s "
                   f"= {rp['env_steps_per_sec_per_core']:.0f} env-steps/s per core.")
    cp = os.path.join(ROOT, 'profiles', 'counters.json')
    if os.path.exists(cp):
        c = json.load(open(cp))
        for tag in ('2v2', 'ffa'):
            if tag in c:
                x = c[tag]
                out.append(f"ncu `{x['kernel'].split('(')[0]}` ({x['envs_per_launch']} envs): {x['warp_inst_per_launch'] / 1e6:.1f} M warp instructions, "
                           f"{x['active_lanes_per_inst']:.1f} active lanes, issue slots {x['issue_slots_busy_pct']:.1f} % busy, DRAM {x['dram_bytes_per_launch'] / 1e6:.1f} MB per launch.")
    out.append('<!-- MEASURED_TABLE end -->')
    p = os.path.join(ROOT, 'DESIGN.md')
    s = open(p).read()
    block = '\n'.join(out)
    if 'MEASURED_TABLE begin' in s:
        s = re.sub(r'<!-- MEASURED_TABLE begin.*?MEASURED_TABLE end -->', lambda m: block, s, flags=re.S)
    else:
        s = s.replace('MEASURED_TABLE', block, 1)
    open(p, 'w').write(s)
    print(block)


if __name__ == '__main__':
    main()
