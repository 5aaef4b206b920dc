"""Per-source-line summary of an ncu report:  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X_cs.csv ; python tools/ncu_lines.py X_cs.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
cur = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr and r[0] not in ('', '-') and len(r) >= 10:
        try:
            ln = int(r[0])
        except ValueError:
            continue
        def col(name):
            i = hdr.index(name) - len(hdr)      # from the right: a source line may contain the CSV quote character
            try:
                return float(r[i].replace(',', '')) if r[i] not in ('-', '') else 0.0
            except ValueError:
                return 0.0
        out.append((cur, ln, r[1].strip()[:110], col('Instructions Executed'), col('Thread Instructions Executed'), col('# Samples')))
ti = sum(o[3] for o in out); ts = sum(o[5] for o in out)
print('total warp inst %.4g samples %d' % (ti, ts))
for o in sorted(out, key=lambda o: -o[5])[:top]:
    print('%-18s %4d inst %5.1f%% lanes %4.1f samp %5.1f%% | %s' % (o[0][:18], o[1], 100 * o[3] / ti, o[4] / max(o[3], 1), 100 * o[5] / ts, o[2]))
