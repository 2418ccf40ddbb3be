"""Per-source-line executed-instruction and stall-sample shares of one kernel, from an `ncu --set full --import-source on`
report and the in-tree library built with -lineinfo.

    python tools/ncu_lines.py --rep gpurun_out/prof.ncu-rep --kernel caps2_fwd [--top 40]

ncu's `--page source --csv` gives per-SASS-instruction counters keyed by address; `nvdisasm -gi` of the matching cubin
gives the source line (innermost inlined frame) of every SASS offset.  This joins the two.
"""
import argparse
import collections
import csv
import glob
import io
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'torch_scae_b200', 'csrc', 'libscae_b200.so')


def sass_line_map(kernel_regex):
    """{sass offset: (file, line, sass text)} for the first function whose mangled name matches."""
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', LIB], cwd=tmp, check=True, capture_output=True)
    for cubin in sorted(glob.glob(os.path.join(tmp, '*.cubin'))):
        out = subprocess.run(['nvdisasm', '-gi', '-c', cubin], capture_output=True, text=True).stdout
        fn, cur, table, hit = None, None, {}, False
        for line in out.splitlines():
            g = re.match(r'\s*\.text\.(\S+):', line)
            if g:
                if hit:
                    return table
                fn = g.group(1)
                hit = re.search(kernel_regex, fn) is not None
                table, cur = {}, None
                continue
            if not hit:
                continue
            g = re.search(r'//## File "([^"]+)", line (\d+)', line)
            if g:
                cur = (os.path.basename(g.group(1)), int(g.group(2)))
                continue
            g = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
            if g:
                table[int(g.group(1), 16)] = (cur, g.group(2).strip())
        if hit:
            return table
    raise SystemExit(f'no function matching {kernel_regex!r} in {LIB}')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rep', required=True)
    ap.add_argument('--kernel', required=True, help='regex on the kernel name')
    ap.add_argument('--top', type=int, default=40)
    ap.add_argument('--launch', type=int, default=0, help='which matching launch in the report')
    args = ap.parse_args()
    table = sass_line_map(args.kernel)
    txt = subprocess.run(['ncu', '-i', args.rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{args.kernel}'],
                         capture_output=True, text=True).stdout
    # one block per launch: a "Kernel Name" row, a header row, then the instructions
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(txt)):
        if row and row[0] == 'Kernel Name':
            cur = []
            blocks.append(cur)
        elif cur is not None and row:
            cur.append(row)
    rows = blocks[args.launch]
    hdr = rows[0]
    ia, iex, ism = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    base = int(rows[1][ia], 16)
    by_line, by_op = collections.Counter(), collections.Counter()
    smp_line = collections.Counter()
    tot = tots = 0
    for r in rows[1:]:
        off = int(r[ia], 16) - base
        ex, sm = int(r[iex]), int(r[ism])
        loc, sass = table.get(off, (None, '?'))
        by_line[loc] += ex
        smp_line[loc] += sm
        op = sass.split()[1 if sass.startswith('@') else 0].split('.')[0] if sass != '?' else '?'
        by_op[op] += ex
        tot += ex
        tots += sm
    print(f'{args.kernel}: {tot} warp instructions executed, {tots} stall samples')
    src_cache = {}
    for loc, ex in by_line.most_common(args.top):
        text = ''
        if loc:
            path = os.path.join(ROOT, 'torch_scae_b200', 'csrc', loc[0])
            if path not in src_cache:
                src_cache[path] = open(path).read().splitlines() if os.path.exists(path) else []
            lines = src_cache[path]
            text = lines[loc[1] - 1].strip()[:90] if 0 < loc[1] <= len(lines) else ''
        print(f'{ex:10d} {100 * ex / tot:5.1f}%  stall {100 * smp_line[loc] / max(tots, 1):5.1f}%  {loc}  {text}')
    print('--- by opcode')
    for op, ex in by_op.most_common(25):
        print(f'{ex:10d} {100 * ex / tot:5.1f}%  {op}')


if __name__ == '__main__':
    main()
