# -*- coding: utf-8 -*-
"""Turns the raw ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.
    python scripts/summarize_profiles.py <tag> [file ...]     e.g. r01c launches_r4_bench.csv prof_tree_attn_r2.ncu-rep
Without file arguments every launches*.csv / *.ncu-rep in gpurun_out/ is summarised.  Also writes
profiles/<tag>_traffic.json (DRAM bytes per launch of each fully captured kernel), which bench.py reads for
roofline.traffic."""
import json
import collections
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'profiles')
GO = os.path.join(ROOT, 'gpurun_out')
tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
os.makedirs(OUT, exist_ok=True)


def short(n):
    n = re.sub(r'\(.*', '', n)
    n = n.replace('pia::', '').replace('void ', '')
    return n[:64]


def launch_list(path, name):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = [(int(r['ID']), r['Kernel Name'], float(r['Metric Value'].replace(',', '')))
            for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
    gets = [i for i, r in enumerate(rows) if 'k_get' in r[1]]
    cmd = ('bench.py --steps 2 --warmup 3 --no-cpu-baseline' if 'bench' in name else
           'scripts/profile_step.py --batch 8 (batched loop, 8 request slots)' if 'batch' in name else 'scripts/profile_step.py')
    out = [f'# ncu launch list ({name}): `ncu --metrics gpu__time_duration.sum --clock-control none` on '
           f'`{cmd}` (Llama-2-7B shape)', '',
           'Per-launch times are cold-cache and serialised: read the SHARES, not the absolutes.', '',
           f'{len(rows)} launches captured; one decode step = the launches between two consecutive `k_get`.', '']
    if len(gets) >= 2:
        seg = rows[gets[-2]:gets[-1]]
        tot = sum(r[2] for r in seg)
        agg = collections.defaultdict(lambda: [0, 0.0])
        for r in seg:
            agg[short(r[1])][0] += 1
            agg[short(r[1])][1] += r[2]
        out += [f'## one decode step: {len(seg)} launches, sum of kernel durations {tot / 1e3:.0f} us', '',
                '| kernel | launches | total us | share |', '|---|---:|---:|---:|']
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.append(f'| `{k}` | {v[0]} | {v[1] / 1e3:.1f} | {100 * v[1] / tot:.1f}% |')
        own = sum(r[2] for r in seg if re.search(r'\b(gemm|attn|fused|trie|accept)::k_|\bk_(moe_combine|l2_prefetch)\b', r[1]))
        blas = sum(r[2] for r in seg if 'nvjet' in r[1] or 'cutlass' in r[1] or 'gemv' in r[1])
        out += ['', f'own kernels (libpia_b200, `pia::*`): {100 * own / tot:.1f}% of the step; cuBLAS GEMMs (`nvjet_*`): '
                    f'{100 * blas / tot:.1f}%; other (torch elementwise): {100 * (tot - own - blas) / tot:.1f}%']
    open(os.path.join(OUT, f'{tag}_{name}.md'), 'w').write('\n'.join(out) + '\n')


WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'sm__cycles_active.avg', 'sm__cycles_elapsed.max',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum']


TRAFFIC = {}


def full(rep, name):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(raw)))
    if len(rd) < 3:
        return
    hdr, units = rd[0], rd[1]
    out = [f'# ncu --set full summary: {name}', '', f'source: `{os.path.basename(rep)}` (kept in gpurun_out/, not tracked)', '']
    for row in rd[2:]:
        d = dict(zip(hdr, row))
        out += [f"## launch {d.get('ID')}: `{short(d.get('Kernel Name', ''))}`", '', '| metric | value | unit |', '|---|---:|---|']
        for w in WANT:
            if w in d:
                out.append(f'| {w} | {d[w]} | {units[hdr.index(w)]} |')
        try:
            scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
            tb = sum(float(d[m].replace(',', '')) * scale.get(units[hdr.index(m)], 1.0)
                     for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
            TRAFFIC.setdefault(name + ':' + short(d.get('Kernel Name', '')), []).append(tb)
            out.append(f'| dram bytes read + written | {tb / 1e6:.2f} | MB |')
        except (KeyError, ValueError):
            pass
        out.append('')
    open(os.path.join(OUT, f'{tag}_{name}.md'), 'w').write('\n'.join(out) + '\n')


only = sys.argv[2:]
for f in sorted(os.listdir(GO)):
    if only and f not in only:
        continue
    p = os.path.join(GO, f)
    if f.startswith('launches') and f.endswith('.csv'):
        launch_list(p, f[:-4])
    elif f.endswith('.ncu-rep'):
        full(p, f[:-8])
if TRAFFIC:
    tp = os.path.join(OUT, f'{tag}_traffic.json')
    merged = json.load(open(tp)) if os.path.exists(tp) else {}   # captures summarised earlier stay
    merged.update({k: float(sum(v) / len(v)) for k, v in TRAFFIC.items()})
    json.dump(merged, open(tp, 'w'), indent=1, sort_keys=True)
print(sorted(os.listdir(OUT)))
