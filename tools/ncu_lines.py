#!/usr/bin/env python3
"""Join ncu per-SASS-instruction counts with nvdisasm line info: instructions executed per source line.
usage: ncu_lines.py <report.ncu-rep> <disasm from `nvdisasm -g cubin`> <mangled-name substring> <source file> [elements]"""
import collections, csv, re, subprocess, sys
rep, dis, fnsub, srcf = sys.argv[1:5]
elems = float(sys.argv[5]) if len(sys.argv) > 5 else None
fn = None; line = None; seq = []
for l in open(dis):
    m = re.match(r'\s*\.text\.(\S+):', l)
    if m: fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: line = int(m.group(2)); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m and fn and fnsub in fn: seq.append((int(m.group(1), 16), line, m.group(2).strip()))
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
body = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name': break
    body.append(r)
base = int(body[0][0], 16)
cnt = {int(r[0], 16) - base: (int(r[5]) if r[5].isdigit() else 0) for r in body}
stall = {int(r[0], 16) - base: (int(r[2]) if r[2].isdigit() else 0) for r in body}
per = collections.Counter(); pst = collections.Counter()
for off, ln, txt in seq:
    per[ln] += cnt.get(off, 0); pst[ln] += stall.get(off, 0)
tot = sum(per.values()); tst = sum(pst.values())
src = open(srcf).read().splitlines()
print('sass', len(seq), 'ncu', len(body), 'total warp insts', tot, ('thread-inst/elem %.2f' % (tot * 32 / elems)) if elems else '')
for ln, c in per.most_common(int(sys.argv[6]) if len(sys.argv)>6 else 36):
    print(f'{c:9d} {100*c/tot:5.1f}% stall {100*pst[ln]/max(tst,1):5.1f}%  L{ln}: {src[ln-1].strip()[:95] if ln and ln <= len(src) else ""}')
