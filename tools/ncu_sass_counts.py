"""Per-SASS-instruction execution counts of one kernel from an ncu report (source page): histogram of the execution-count
classes (loop nest levels) with their share of the instruction stream, and the memory / vote / branch instructions.

    python tools/ncu_sass_counts.py <rep> [kernel regex] [min_count for the listing]"""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
minc = int(sys.argv[3]) if len(sys.argv) > 3 else 10**12
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
sec, hdr, data, done = None, None, [], False
for r in csv.reader(out.splitlines()):
    if r and r[0] == 'Kernel Name':
        if data:
            break
        sec, hdr = r[1], None
        continue
    if r and r[0] == 'Address':
        hdr = r
        continue
    if sec and hdr and len(r) == len(hdr) and (pat is None or pat.search(sec)):
        data.append((r[0][-5:], r[hdr.index('Source')].strip(), int(r[hdr.index('Instructions Executed')]),
                     r[hdr.index('Avg. Predicated-On Threads Executed')], r[hdr.index('L1 Wavefronts Shared')]))
tot = sum(d[2] for d in data)
print(sec[:100] if sec else None, 'total', tot, 'static', len(data))
hist = collections.Counter()
for d in data:
    hist[d[2]] += 1
print('exec count x static instructions = share')
for k, v in sorted(hist.items(), key=lambda kv: -kv[0] * kv[1])[:16]:
    print(f'{k:>10} x {v:>4} = {100.0 * k * v / tot:5.1f} %')
for d in data:
    if d[2] >= minc and re.search(r'MATCH|CREDUX|VOTE|BAR.SYNC|MUFU.EX2|STS|LDS|BRA|LDG|STG', d[1]):
        print(d[0], d[1][:46].ljust(46), str(d[2]).rjust(10), d[3].rjust(5), d[4].rjust(10))
