"""Summarise an `ncu --page source --csv` dump: contiguous SASS regions with similar execution counts, their share of the
executed warp instructions and of the stall samples, and their most frequent opcodes.  python sass_regions.py dump.csv [kernel name substring]"""
import csv, sys
allrows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"] + [len(allrows)]
sec = [k for k in range(len(starts) - 1) if want in allrows[starts[k]][1]][0]
rows = allrows[starts[sec]:starts[sec + 1]]
print(rows[0][1])
h = rows[1]
data = [r for r in rows[2:] if len(r) == len(h)]
iS, iE, iN = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
stall = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
tot = sum(int(r[iE]) for r in data); totS = sum(int(r[iN]) for r in data)
print("total warp instructions", tot, "samples", totS, "sass lines", len(data))
seg, cur = [], None
for k, r in enumerate(data):
    e = int(r[iE])
    if cur is None or not (0.7 * cur['e0'] <= e <= 1.4 * cur['e0']):
        if cur: seg.append(cur)
        cur = {'start': k, 'e0': max(e, 1), 'sumE': 0, 'sumS': 0, 'n': 0, 'ops': {}, 'st': {}}
    cur['sumE'] += e; cur['sumS'] += int(r[iN]); cur['n'] += 1
    t = r[iS].split()
    op = t[1] if t[0].startswith('@') else t[0]
    cur['ops'][op] = cur['ops'].get(op, 0) + 1
    for i in stall:
        v = int(r[i] or 0)
        if v: cur['st'][h[i]] = cur['st'].get(h[i], 0) + v
seg.append(cur)
for s in seg:
    if s['sumE'] > 0.01 * tot or s['sumS'] > 0.01 * totS:
        top = sorted(s['ops'].items(), key=lambda x: -x[1])[:6]
        st = sorted(s['st'].items(), key=lambda x: -x[1])[:3]
        print(f"sass[{s['start']:5d}+{s['n']:4d}] exec/instr={s['e0']:9d} instr%={100*s['sumE']/tot:5.1f} samples%={100*s['sumS']/totS:5.1f} {top} {st}")
