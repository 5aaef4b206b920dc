"""Phase breakdown of k_forward_grid from an ncu cuda,sass source export (see tools/ncu_lines.py): python tools/ncu_phases.py X_cs.csv
Source lines are mapped to phases by the marker comments of render_kernels.cu at the time of the capture (pass the .cu used)."""
import csv, sys, re
src = open(sys.argv[2] if len(sys.argv) > 2 else 'nlos_surface_optimization_b200/csrc/render_kernels.cu').read().split('\n')
def find(pat, start=0):
    for i in range(start, len(src)):
        if pat in src[i]: return i + 1
    return 10**9
k0 = find('K1g forward, perspective grid')
marks = [(find('K1g forward, perspective grid'), 'prologue/scan-util'), (find('// ---------------- frame', k0), 'frame'), (find('pass 0: project', k0), 'pass0'),
         (find('pass 1: count', k0), 'pass1'), (find('const uint2 r = trect[p]', k0) - 1, 'pass2'), (find('pass 3: samples', k0), 'p3 load+cull'),
         (find('for (int k = 0; k < P.spp; ++k) {', k0), 'p3 generate'), (find('bool occ = false;', k0), 'p3 cell lookup'), (find('const int maxg', k0), 'p3 scan'),
         (find('pooled exact tests', k0), 'p3 pool build'), (find('for (int i = lane; i < total', k0), 'p3 exact'), (find('const bool visible = need', k0), 'p3 splat/vis'),
         (find('K3 residual'), 'other')]
core = open('nlos_surface_optimization_b200/csrc/nlos_core.cuh').read().split('\n')
def cfind(p):
    for i, l in enumerate(core):
        if p in l: return i + 1
    return 10**9
c_proj, c_rect, c_entry, c_ray, c_pre, c_frame = cfind('NLOS_HD void pg_project'), cfind('NLOS_HD void pg_tri_rect'), cfind('NLOS_HD unsigned pg_entry'), cfind('NLOS_HD void pg_ray'), cfind('NLOS_HD bool pg_precheck'), cfind('NLOS_HD void pg_init_frame')
c_occ0, c_occ1 = cfind('NLOS_HD bool tri_occludes_od'), cfind('NLOS_HD bool tri_occludes_fast')
c_quant = cfind('NLOS_HD int pg_quant')
def phase(f, l):
    if f == 'render_kernels.cu':
        name = 'before'
        for ln, nm in marks:
            if l >= ln: name = nm
        return name
    if f == 'nlos_core.cuh':
        if c_proj <= l < c_quant: return 'pass0'
        if c_quant <= l < c_entry: return 'pass1'
        if c_entry <= l < c_ray: return 'pass2'
        if c_ray <= l < c_pre: return 'p3 cell lookup'
        if c_pre <= l < c_frame: return 'p3 scan'
        if c_occ0 <= l < c_occ1 or 150 <= l <= 162: return 'p3 exact'
        if 66 <= l <= 110 or l >= cfind('struct ShadeTri'): return 'p3 generate'
        return 'vector algebra (shared)'
    return 'intrinsics'
rows = list(csv.reader(open(sys.argv[1])))
cur = None; hdr = None; agg = {}; ti = ts = 0
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr and r[0] not in ('', '-') and len(r) >= 10:
        try: ln = int(r[0])
        except ValueError: continue
        def col(n):
            i = hdr.index(n) - len(hdr)
            try: return float(r[i].replace(',', ''))
            except ValueError: return 0.0
        i_, t_, s_ = col('Instructions Executed'), col('Thread Instructions Executed'), col('# Samples')
        a = agg.setdefault(phase(cur, ln), [0, 0, 0]); a[0] += i_; a[1] += t_; a[2] += s_; ti += i_; ts += s_
print('total warp inst %.4g' % ti)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print('%-28s inst %5.1f%% (%.2e) lanes %4.1f samples %5.1f%%' % (k, 100 * a[0] / ti, a[0], a[1] / max(a[0], 1), 100 * a[2] / ts))
