"""Hot spots of one kernel from an .ncu-rep source page (SASS view): python tools/ncu_hot.py <rep> [kernel-index] [top]
Prints total samples, instructions executed, and the SASS lines with the most stall samples (with neighbours)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == 'Kernel Name':
        cur = dict(name=row[1], hdr=None, rows=[])
        blocks.append(cur)
    elif cur is not None and cur['hdr'] is None:
        cur['hdr'] = row
    elif cur is not None and row:
        cur['rows'].append(row)
b = blocks[which]
h = b['hdr']
iS, iI, iT, iSrc = h.index('# Samples'), h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('Source')
stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
rows = b['rows']
tot_s = sum(int(r[iS]) for r in rows)
tot_i = sum(int(r[iI]) for r in rows)
print(b['name'], 'SASS lines', len(rows), 'samples', tot_s, 'warp-instr', tot_i, 'lane-instr', sum(int(r[iT]) for r in rows))
agg = {}
for r in rows:
    for i, n in stall_cols:
        agg[n] = agg.get(n, 0) + int(r[i])
print('stalls:', ', '.join(f'{n[6:]}={v}' for n, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
order = sorted(range(len(rows)), key=lambda i: -int(rows[i][iS]))[:top]
for i in sorted(order):
    r = rows[i]
    st = ', '.join(f'{n[6:]}={r[j]}' for j, n in stall_cols if int(r[j]) > 0.15 * max(1, int(r[iS])))
    print(f'{i:5d} {int(r[iS]):6d} {int(r[iI]):9d}  {r[iSrc].strip():60s} {st}')
if len(sys.argv) > 4:   # instruction histogram in chunks of N SASS lines
    n = int(sys.argv[4])
    for s in range(0, len(rows), n):
        chunk = rows[s:s + n]
        ins = sum(int(r[iI]) for r in chunk)
        smp = sum(int(r[iS]) for r in chunk)
        if ins > 0.005 * tot_i:
            print(f'lines {s:5d}-{s + n:5d}: instr {ins / tot_i * 100:5.1f}%  samples {smp / tot_s * 100:5.1f}%   first: {chunk[0][iSrc].strip()[:50]}')
