"""Per-SASS-instruction execution counts of one kernel from an ncu report (source page): memory, vote, branch lines and
the totals per region between them.   python tools/ncu_sass_counts.py <rep> [min_count]"""
import csv, re, subprocess, sys
rep = sys.argv[1]
minc = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
isrc, iex, ith, iwf = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Avg. Predicated-On Threads Executed'), hdr.index('L1 Wavefronts Shared')
tot = 0
run = 0
for r in rows[2:]:
    s = r[isrc]
    n = int(r[iex])
    tot += n
    run += n
    if re.search(r'MATCH|CREDUX|VOTE|BAR.SYNC|MUFU.EX2|STS|LDS|BRA|LDG|STG', s) and n >= minc:
        print(r[0][-5:], s.strip()[:46].ljust(46), r[iex].rjust(10), r[ith].rjust(5), r[iwf].rjust(10), ' cum', run)
        run = 0
print('total', tot)
