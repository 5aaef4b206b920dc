"""Registers / spills of the instantiations of one kernel from the ptxas log: python tools/ptxas_regs.py k_forward_group"""
import re, sys
s = open('nlos_surface_optimization_b200/csrc/render_kernels.ptxas.log').read()
blocks = re.split(r"ptxas info\s+: Compiling entry function '([^']+)'", s)
seen = {}
for i in range(1, len(blocks), 2):
    name, body = blocks[i], blocks[i + 1]
    if sys.argv[1] in name:
        m = re.search(r'Used (\d+) registers', body); sp = re.search(r'(\d+) bytes spill stores', body); st = re.search(r'(\d+) bytes stack', body)
        key = (m.group(1), sp.group(1), st.group(1)); seen[key] = seen.get(key, 0) + 1
for k, n in seen.items(): print('regs %s spill %s stack %s  x%d' % (k + (n,)))
