"""Summarise an ncu report exported with `--page raw --csv` and `--page source --csv --print-source sass`.

    ncu -i X.ncu-rep --page raw --csv > X_raw.csv
    ncu -i X.ncu-rep --page source --csv --print-source sass > X_sass.csv
    python tools/ncu_summary.py X_raw.csv [X_sass.csv]
"""
import csv
import sys

raw, sass = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
rows = list(csv.reader(open(raw)))
hdr = rows[0]; units = rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_xu.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sectors_op_red.sum',
        'lts__t_sectors_op_atom.sum', 'sm__cycles_elapsed.avg.per_second', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('=' * 100)
    for k in keys:
        if k in d:
            print('%-76s %s %s' % (k, d[k], units[hdr.index(k)]))


def num(x):
    try:
        return float(x.replace(',', ''))
    except ValueError:
        return 0.0


if sass:
    allrows = list(csv.reader(open(sass)))
    sections, cur = [], None
    for r in allrows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1] if len(r) > 1 else '?', 'rows': []}; sections.append(cur)
        elif cur is not None:
            cur['rows'].append(r)
    for sec in sections:
        rows = sec['rows']
        hdr = rows[0]; ix = {k: i for i, k in enumerate(hdr)}
        data = [r for r in rows[1:] if len(r) == len(hdr)]
        print('=' * 100)
        print('SASS profile of', sec['name'][:120])
        tot_inst = sum(num(r[ix['Instructions Executed']]) for r in data); tot_samp = sum(num(r[ix['# Samples']]) for r in data)
        st = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
        print('total warp inst %.4g, samples %d' % (tot_inst, tot_samp))
        print('stall mix (all samples): ' + ' '.join('%s=%.1f%%' % (k[6:], 100 * sum(num(r[ix[k]]) for r in data) / max(tot_samp, 1))
                                                     for k in sorted(st, key=lambda k: -sum(num(r[ix[k]]) for r in data))[:8]))
        blk = 32
        for b in range(0, len(data), blk):
            seg = data[b:b + blk]
            inst = sum(num(r[ix['Instructions Executed']]) for r in seg); th = sum(num(r[ix['Thread Instructions Executed']]) for r in seg)
            samp = sum(num(r[ix['# Samples']]) for r in seg)
            if inst / max(tot_inst, 1) < 0.004:
                continue
            stalls = {k: sum(num(r[ix[k]]) for r in seg) for k in st}
            top = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
            ops = ' '.join(sorted(set(r[ix['Source']].split()[0] for r in seg if r[ix['Source']])))
            print('%4d-%4d inst %5.1f%% thr/inst %5.1f samp %5.1f%% %s | %s' % (b, b + blk, 100 * inst / tot_inst, th / max(inst, 1), 100 * samp / max(tot_samp, 1),
                                                                               ' '.join('%s=%.0f%%' % (k[6:], 100 * v / max(samp, 1)) for k, v in top), ops[:90]))
