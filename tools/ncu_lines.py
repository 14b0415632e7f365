"""Attribute ncu SASS-level samples/instruction counts to CUDA source lines.
usage: ncu_lines.py <report.ncu-rep> <kernel regex> <launch-skip> <mangled-substr> [top]"""
import collections, csv, io, re, subprocess, sys
rep, kre, skip, sub = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
sass = open('/tmp/sass/all.sass').read().split('\n')
# line map for the function: ordered list of (line, file) per instruction
start = next(i for i, l in enumerate(sass) if l.startswith('.text.') and sub in l)
cur = None; order = []
for l in sass[start + 1:]:
    if l.startswith('\t.section') or (l.startswith('.text.') ): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l): order.append(cur)
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kre, '--launch-skip', skip, '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if '# Samples' in r)
h = rows[hi]; I = {n: i for i, n in enumerate(h)}
body = [r for r in rows[hi + 1:] if len(r) == len(h) and r[I['Address']] != 'Address']
body = body[:len(order)] if len(body) > len(order) else body
print('sass rows', len(body), 'line-mapped instrs', len(order))
S = collections.Counter(); E = collections.Counter()
for r, ln in zip(body, order):
    S[ln] += int(r[I['# Samples']] or 0); E[ln] += int(r[I['Instructions Executed']] or 0)
ts, te = sum(S.values()), sum(E.values())
src = {}
for (f, n) in S:
    if f not in src:
        try: src[f] = open('/root/repo/grid_ndt_b200/csrc/' + f).read().split('\n')
        except Exception: src[f] = []
print('total samples', ts, 'warp instr', te)
for ln, c in sorted(S.items(), key=lambda kv: -kv[1])[:top]:
    f, n = ln if ln else ('?', 0)
    text = src.get(f, [])[n - 1].strip()[:80] if ln and n - 1 < len(src.get(f, [])) else ''
    print(f'{100*c/ts:5.1f}% stall  {100*E[ln]/te:5.1f}% instr  {f}:{n}  {text}')
