"""Executed warp-instructions per SASS opcode for one kernel launch of an ncu report.
usage: ncu_opcodes.py <report.ncu-rep> <kernel regex> [launch-skip] [top]"""
import collections, csv, io, re, subprocess, sys
rep, kre = sys.argv[1:3]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kre, '--launch-skip', skip, '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
h = rows[hi]; I = {n: i for i, n in enumerate(h)}
E = collections.Counter(); T = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) != len(h) or r[I['Address']] == 'Address': continue
    src = r[I['Source']].strip()
    src = re.sub(r'^@!?U?P\d+\s+', '', src)
    op = src.split()[0].rstrip(';') if src else '?'
    base = op.split('.')[0]
    E[base] += int(r[I['Instructions Executed']] or 0)
    T[op] += int(r[I['Instructions Executed']] or 0)
tot = sum(E.values())
print('total warp instr', tot)
for k, v in E.most_common(top): print(f'{100*v/tot:5.1f}%  {v:>10}  {k}')
xu = sum(v for k, v in E.items() if k in ('F2F', 'F2I', 'I2F', 'FRND', 'MUFU', 'I2FP', 'F2FP', 'POPC', 'FLO', 'BREV'))
d = sum(v for k, v in E.items() if k in ('DADD', 'DMUL', 'DFMA', 'DSETP', 'DMNMX'))
print(f'XU-class {xu} ({100*xu/tot:.1f}%)   FP64-class {d} ({100*d/tot:.1f}%)')
print('detail:', [(k, v) for k, v in T.most_common(60) if k.split('.')[0] in ('F2F', 'F2I', 'I2F', 'FRND', 'MUFU', 'POPC', 'FLO', 'BREV', 'I2FP')])
