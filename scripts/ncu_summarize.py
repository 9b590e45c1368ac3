"""Turn an `ncu --set full` report (.ncu-rep) and/or a launch list (.csv from
`ncu --metrics gpu__time_duration.sum --csv`) into the small text summaries committed under profiles/.

    python scripts/ncu_summarize.py --rep gpurun_out/prof_gram.ncu-rep --launches gpurun_out/launches.csv \
        --out profiles/r01_gram_hist --note "20k x 512, fp16x3, cta_group 2"
"""
import argparse
import csv
import io
import json
import re
import subprocess
from collections import OrderedDict
from pathlib import Path

RAW_KEEP = re.compile(
    r'^(gpu__time_duration\.sum|sm__cycles_elapsed\.avg|smsp__cycles_active\.avg|dram__bytes_read\.sum|dram__bytes_write\.sum|'
    r'dram__bytes_read\.sum\.per_second|lts__t_bytes\.sum|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|'
    r'l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|'
    r'sm__inst_executed\.sum|smsp__inst_executed\.sum|smsp__issue_active\.avg\.pct_of_peak_sustained_active|'
    r'smsp__thread_inst_executed_per_inst_executed\.ratio|sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|'
    r'sm__pipe_tensor_subpipe_hmma_cycles_active\.avg\.pct_of_peak_sustained_active|'
    r'sm__pipe_(alu|fma|fmaheavy|shared)_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|'
    r'sm__pipe_tma_cycles_active\.avg\.pct_of_peak_sustained_elapsed|'
    r'l1tex__data_bank_conflicts_pipe_lsu_mem_shared(_op_(ld|st|atom))?\.sum|'
    r'smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|smsp__warps_eligible\.avg\.per_cycle_active|'
    r'launch__(registers_per_thread|shared_mem_per_block_dynamic|grid_size|block_size|cluster_size|waves_per_multiprocessor)|'
    r'sm__inst_executed_pipe_(uniform|tensor\w*|lsu|alu|fma\w*|xu)\.sum|smsp__inst_executed_op_\w+\.sum)$')


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(['ncu', '-i', str(rep), '--page', page, '--csv', *extra], capture_output=True, text=True)
    if out.returncode != 0:
        raise SystemExit(out.stderr)
    return out.stdout


def summarize_raw(rep):
    rows = list(csv.reader(io.StringIO(ncu_csv(rep, 'raw'))))
    hdr, units = rows[0], rows[1]
    kernels = []
    for r in rows[2:]:
        d = OrderedDict()
        d['kernel'] = r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''
        for i, h in enumerate(hdr):
            if RAW_KEEP.match(h) and r[i] != '':
                d[h] = '%s %s' % (r[i], units[i]) if units[i] else r[i]
        kernels.append(d)
    return kernels


def summarize_source(rep, top=40):
    text = ncu_csv(rep, 'source', ('--print-source', 'sass'))
    rows = list(csv.reader(io.StringIO(text)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
    out = []
    for n, s in enumerate(starts):
        e = starts[n + 1] if n + 1 < len(starts) else len(rows)
        hdr = rows[s + 1]
        ix = {h: i for i, h in enumerate(hdr)}
        body = rows[s + 2:e]
        stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
        tot = {h: sum(int(r[ix[h]] or 0) for r in body) for h in stall_cols}
        hot = sorted(body, key=lambda r: -int(r[ix['# Samples']] or 0))[:top]
        out.append({'kernel': rows[s][1], 'sass_instructions': len(body),
                    'instructions_executed': sum(int(r[ix['Instructions Executed']] or 0) for r in body),
                    'samples': sum(int(r[ix['# Samples']] or 0) for r in body),
                    'stall_samples': tot,
                    'hot': [{'idx': body.index(r), 'sass': r[ix['Source']].strip(), 'samples': int(r[ix['# Samples']] or 0),
                             'executed': int(r[ix['Instructions Executed']] or 0)} for r in hot]})
    return out


def summarize_launches(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    hdr = rows[start]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= ix['Metric Value'] or r[ix['Metric Name']] != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', r[ix['Kernel Name']].split('<')[0]).strip()
        if 'gram_kernel' in r[ix['Kernel Name']]:
            name = r[ix['Kernel Name']].split('(CUtensorMap')[0]
        a = agg.setdefault(name, {'launches': 0, 'ns': 0.0, 'grid': r[ix['Grid Size']], 'block': r[ix['Block Size']]})
        a['launches'] += 1
        a['ns'] += float(r[ix['Metric Value']].replace(',', ''))
    total = sum(a['ns'] for a in agg.values()) or 1.0
    return [{'kernel': k, **v, 'share': v['ns'] / total} for k, v in sorted(agg.items(), key=lambda kv: -kv[1]['ns'])]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rep')
    ap.add_argument('--launches')
    ap.add_argument('--out', required=True)
    ap.add_argument('--note', default='')
    ap.add_argument('--top', type=int, default=40)
    a = ap.parse_args()
    out = Path(a.out)
    out.parent.mkdir(parents=True, exist_ok=True)
    doc = {'note': a.note}
    lines = ['# ncu summary: %s' % out.name, '', a.note, '']
    if a.launches:
        doc['launches'] = summarize_launches(a.launches)
        lines += ['## launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised)', '',
                  '| kernel | launches | total us | share | grid | block |', '|---|---:|---:|---:|---|---|']
        for k in doc['launches']:
            lines.append('| `%s` | %d | %.1f | %.1f%% | %s | %s |' % (k['kernel'][:90], k['launches'], k['ns'] / 1e3, 100 * k['share'], k['grid'], k['block']))
        lines.append('')
    if a.rep:
        doc['raw'] = summarize_raw(a.rep)
        doc['source'] = summarize_source(a.rep, a.top)
        for n, k in enumerate(doc['raw']):
            lines += ['## kernel %d: `%s` (--set full)' % (n, k['kernel'][:120]), '']
            for key, v in k.items():
                if key != 'kernel':
                    lines.append('- `%s` = %s' % (key, v))
            lines.append('')
        for n, k in enumerate(doc['source'][:1]):
            lines += ['## SASS hot spots, kernel %d (%d SASS instructions, %d warp-instructions executed, %d samples)' %
                      (n, k['sass_instructions'], k['instructions_executed'], k['samples']), '',
                      'stall samples: ' + ', '.join('%s=%d' % (h, v) for h, v in k['stall_samples'].items() if v), '',
                      '| idx | samples | executed | SASS |', '|---:|---:|---:|---|']
            for h in k['hot']:
                lines.append('| %d | %d | %d | `%s` |' % (h['idx'], h['samples'], h['executed'], h['sass']))
            lines.append('')
    out.with_suffix('.md').write_text('\n'.join(lines))
    out.with_suffix('.json').write_text(json.dumps(doc, indent=1))
    print('wrote', out.with_suffix('.md'), out.with_suffix('.json'))


if __name__ == '__main__':
    main()
