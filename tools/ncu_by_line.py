"""Aggregates an ncu SASS source page (--page source --csv --print-source sass) by CUDA source
line, using nvdisasm -g line markers of the same cubin.  Usage:
  python tools/ncu_by_line.py sass.csv dis.txt <kernel mangled substring> [top]"""
import csv, re, sys, collections
sass_csv, dis, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# 1) instruction -> (file,line) in order, from nvdisasm
lines = open(dis, errors='ignore').read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('\t.section\t.text.') and kname in l)
cur = ('?', 0); seq = []
inline_stack = ''
for l in lines[start + 1:]:
    if l.startswith('\t.section') and seq: break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: seq.append((int(m.group(1), 16), cur, m.group(2)))
# 2) metrics from ncu
rows = list(csv.reader(open(sass_csv)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
print('sass rows', len(data), 'disasm instrs', len(seq))
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
tot = [0, 0, 0]
n = min(len(data), len(seq))
for r, (addr, loc, text) in zip(data[:n], seq[:n]):
    ie = int(r[col['Instructions Executed']] or 0); te = int(r[col['Thread Instructions Executed']] or 0)
    sm = int(r[col['# Samples']] or 0)
    a = agg[loc]; a[0] += ie; a[1] += te; a[2] += sm; a[3] += 1
    tot[0] += ie; tot[1] += te; tot[2] += sm
print('total warp-instrs %d thread-instrs %d avg threads %.2f samples %d' % (tot[0], tot[1], tot[1] / max(1, tot[0]), tot[2]))
print('%-22s %6s %8s %8s %7s %5s' % ('file:line', 'sass', 'inst%', 'samp%', 'thr/w', ''))
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print('%-22s %6d %7.2f%% %7.2f%% %7.2f' % ('%s:%d' % loc, a[3], 100 * a[0] / tot[0], 100 * a[2] / max(1, tot[2]), a[1] / max(1, a[0])))
