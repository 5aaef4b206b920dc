"""Hot SASS listing from an ncu source page: ncu -i X.ncu-rep --page source --csv --print-source sass > X_sass.csv ;
python tools/ncu_hot_sass.py X_sass.csv KERNEL_SUBSTRING [min_share_percent] -> every instruction above the share, in address order."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]; thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.15
sections, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1] if len(r) > 1 else '?', 'rows': []}; sections.append(cur)
    elif cur is not None:
        cur['rows'].append(r)
def num(x):
    try: return float(x.replace(',', ''))
    except ValueError: return 0.0
for sec in sections:
    if want not in sec['name']: continue
    hdr = sec['rows'][0]; ix = {k: i for i, k in enumerate(hdr)}
    data = [r for r in sec['rows'][1:] if len(r) == len(hdr)]
    tot = sum(num(r[ix['Instructions Executed']]) for r in data)
    print('SASS of %s: every instruction that is >= %.2f %% of the %.4g executed warp instructions, in address order.' % (sec['name'][:100], thr, tot))
    print('Columns: address, warp instructions executed, share, average active lanes, stall samples, instruction')
    for r in data:
        n = num(r[ix['Instructions Executed']])
        if n / max(tot, 1) * 100 >= thr:
            t = num(r[ix['Thread Instructions Executed']])
            print('%s %12d %5.2f%% thr %4.1f samp %7d | %s' % (r[ix['Address']][-6:] if 'Address' in ix else '', n, 100 * n / tot, t / max(n, 1), num(r[ix['# Samples']]), r[ix['Source']]))
    break
