"""Hot regions of a kernel from an ncu source-page csv: runs of instructions with the same execution count, with their
share of executed warp-instructions and of stall samples.  usage: python profiles/ncu_hot.py <source.csv> <kernel substring>"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]
kern = []; cur = None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}; kern.append(cur)
    elif r and r[0].startswith('0x') and cur is not None:
        cur['rows'].append(r)
k = [k for k in kern if want in k['name']][0]
R = k['rows']
tot = sum(int(r[5]) for r in R); st = sum(int(r[2]) for r in R) or 1
print(k['name'][:90], '\nwarp instr', tot, 'static', len(R), 'stall samples', st)
byop = collections.Counter()
for r in R:
    op = [o for o in r[1].strip().split() if not o.startswith('@')][0].split('.')[0]
    byop[op] += int(r[5])
print('  mix:', ', '.join(f"{op} {100*c/tot:.1f}%" for op, c in byop.most_common(22)))
segs = []; start = 0; last = None
for i, r in enumerate(R):
    c = int(r[5])
    if last is None or abs(c - last) > 0.03 * max(c, last, 1):
        if last is not None: segs.append((start, i - 1))
        start = i; last = c
segs.append((start, len(R) - 1))
for a, b in segs:
    n = sum(int(r[5]) for r in R[a:b + 1]); s_ = sum(int(r[2]) for r in R[a:b + 1])
    if n > 0.015 * tot or s_ > 0.03 * st:
        ops = collections.Counter([[o for o in r[1].strip().split() if not o.startswith('@')][0].split('.')[0] for r in R[a:b + 1]])
        print(f"  [{a:5d}-{b:5d}] {b-a+1:4d} instrs x {int(R[a][5]):9d}  exec {100*n/tot:5.1f}%  stalls {100*s_/st:5.1f}%  {dict(ops.most_common(6))}")
