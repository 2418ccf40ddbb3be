"""Turns the ncu outputs a gpurun call brought back (gpurun_out/) into the tracked summaries under profiles/.

    python tools/summarize_ncu.py --tag r01b --launches gpurun_out/launches.csv --rep gpurun_out/prof.ncu-rep

--launches : csv log of `ncu --metrics gpu__time_duration.sum --clock-control none ... --csv --log-file ...`
--rep      : report of `ncu --set full --clock-control none --import-source on ...` (read with `ncu -i ... --page raw --csv`)
Writes profiles/<tag>_launches.md, profiles/<tag>_kernels.md and (from the full capture) profiles/traffic.json, the
per-launch dram bytes bench.py reports as roofline.traffic.
"""
import argparse
import collections
import csv
import io
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
    'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
    'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum',
]
ENTRY = {'tmpl_ll_fwd': 'scae_tmpl_ll_fwd', 'tmpl_ll_bwd': 'scae_tmpl_ll_bwd', 'caps_ll_fwd': 'scae_caps_ll_fwd',
         'caps_ll_bwd': 'scae_caps_ll_bwd', 'caps2_fwd': 'scae_caps_ll_fwd', 'caps2_bwd': 'scae_caps_ll_bwd',
         'caps3_fwd': 'scae_caps_ll_fwd', 'caps3_bwd': 'scae_caps_ll_bwd'}


def to_bytes(value, unit):
    v = float(value.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)


def launches(path, tag, title):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg, total, n = collections.OrderedDict(), 0.0, 0
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        v = v / 1e3 if row['Metric Unit'] == 'ns' else (v * 1e3 if row['Metric Unit'] == 'ms' else v)
        short = re.sub(r'\(.*', '', row['Kernel Name'])[:110]
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
        n += 1
    out = os.path.join(ROOT, 'profiles', f'{tag}_launches.md')
    with open(out, 'w') as f:
        f.write(f'# ncu launch list -- {title}\n\n')
        f.write('Command (B200, under gpurun): `ncu --metrics gpu__time_duration.sum --clock-control none '
                '--profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_step.py`\n\n')
        f.write(f'{n} kernel launches in one train step, {total / 1e3:.2f} ms total device time (serialised, cold-cache: '
                'compare shares, not absolutes).\n\n| us | share | launches | kernel |\n|---:|---:|---:|---|\n')
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
            f.write(f'| {t:.1f} | {100 * t / total:.1f}% | {c} | `{k}` |\n')
    print('wrote', out)


def kernels(rep, tag, title, batch=1024, write_json=True):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = os.path.join(ROOT, 'profiles', f'{tag}_kernels.md')
    traffic, issue = {}, {}
    with open(out, 'w') as f:
        f.write(f'# ncu --set full, hot-path kernels -- {title}\n\n')
        f.write('Command (B200, under gpurun): `ncu --set full --clock-control none --import-source on --profile-from-start '
                "off -k regex:'tmpl_ll|caps' -o gpurun_out/prof python tools/profile_step.py`; read "
                'here with `ncu -i ... --page raw --csv`.\n\n')
        for d in data:
            name = d[idx['Kernel Name']]
            f.write(f'## `{name[:120]}`\n\n| metric | value | unit |\n|---|---:|---|\n')
            for m in METRICS:
                if m in idx:
                    f.write(f'| {m} | {d[idx[m]]} | {units[idx[m]]} |\n')
            stalls = [(h, float(d[i].replace(',', '') or 0)) for h, i in idx.items()
                      if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
            for h, v in sorted(stalls, key=lambda t: -t[1])[:6]:
                f.write(f'| {h} | {v:.3f} | ratio |\n')
            f.write('\n')
            for key, entry in ENTRY.items():
                if key in name:
                    b = to_bytes(d[idx['dram__bytes_read.sum']], units[idx['dram__bytes_read.sum']]) + \
                        to_bytes(d[idx['dram__bytes_write.sum']], units[idx['dram__bytes_write.sum']])
                    traffic[entry] = traffic.get(entry, 0) + int(b)
                    # lane-instructions the kernel actually executes per launch (ncu): bench.py divides by the work units
                    # of its batch for the issue-roof figures of the path-1 kernels
                    warp_inst = float(d[idx['smsp__inst_executed.sum']].replace(',', ''))
                    lanes = float(d[idx['smsp__thread_inst_executed_per_inst_executed.ratio']].replace(',', ''))
                    issue[entry] = dict(lane_instr_per_launch=warp_inst * lanes, warp_instr_per_launch=warp_inst,
                                        kernel=name[:80])
    if not write_json:
        print('wrote', out)
        return
    with open(os.path.join(ROOT, 'profiles', 'traffic.json'), 'w') as f:
        json.dump(traffic, f, indent=1, sort_keys=True)
    with open(os.path.join(ROOT, 'profiles', 'issue_counts.json'), 'w') as f:
        json.dump(dict(batch=batch, source=f'ncu --set full, {tag}', kernels=issue), f, indent=1, sort_keys=True)
    print('wrote', out, 'and profiles/traffic.json', traffic)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--tag', required=True)
    ap.add_argument('--title', default='one SCAE train step, B=1024, MNIST config (O=32)')
    ap.add_argument('--launches')
    ap.add_argument('--rep')
    ap.add_argument('--no-json', action='store_true', help='only the markdown summary (keep traffic.json / issue_counts.json)')
    ap.add_argument('--batch', type=int, default=1024, help='images per launch in the --rep capture')
    a = ap.parse_args()
    if a.launches:
        launches(a.launches, a.tag, a.title)
    if a.rep:
        kernels(a.rep, a.tag, a.title, a.batch, not a.no_json)
