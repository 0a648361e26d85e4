"""Aggregates an ncu SASS source page by the OUTERMOST source line (the call site in the kernel
body) and, one level down, by the line inside the callee, using `nvdisasm -gi` inline chains.
Usage: python tools/ncu_by_phase.py sass.csv dis_gi.txt <kernel substring> [depth]"""
import csv, re, sys, collections
sass_csv, dis, kname = sys.argv[1], sys.argv[2], sys.argv[3]
depth = int(sys.argv[4]) if len(sys.argv) > 4 else 2
lines = open(dis, errors='ignore').read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('\t.section\t.text.') and kname in l)
seq, chain = [], []
last_chain = []
for l in lines[start + 1:]:
    if l.startswith('\t.section') and seq: break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        chain.append((m.group(1).split('/')[-1], int(m.group(2)), m.group(3) is not None)); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        if chain:
            # chain is innermost ... outermost; cut at the first frame that is not inlined (outermost)
            frames = []
            for f in chain:
                frames.append((f[0], f[1]))
                if not f[2]: break
            last_chain = frames[::-1]  # outermost first
            chain = []
        seq.append((last_chain, m.group(2)))
rows = list(csv.reader(open(sass_csv)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hdr_i]; col = {n: i for i, n in enumerate(hdr)}
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
assert len(data) == len(seq), (len(data), len(seq))
agg = collections.defaultdict(lambda: [0, 0, 0, 0]); tot = [0, 0, 0]
stall_cols = [c for c in hdr if c.startswith('stall_') and 'Not Issued' not in c]
stalls = collections.defaultdict(lambda: collections.Counter())
for r, (frames, text) in zip(data, seq):
    ie = int(r[col['Instructions Executed']] or 0); te = int(r[col['Thread Instructions Executed']] or 0); sm = int(r[col['# Samples']] or 0)
    key = tuple(frames[:depth])
    a = agg[key]; a[0] += ie; a[1] += te; a[2] += sm; a[3] += 1
    tot[0] += ie; tot[1] += te; tot[2] += sm
    for c in stall_cols:
        v = int(r[col[c]] or 0)
        if v: stalls[key][c] += v
print('total warp-instrs %d  avg threads %.2f  samples %d' % (tot[0], tot[1] / max(1, tot[0]), tot[2]))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:int(sys.argv[5]) if len(sys.argv) > 5 else 40]:
    top = ', '.join('%s %.0f%%' % (k.replace('stall_', ''), 100 * v / max(1, a[2])) for k, v in stalls[key].most_common(3))
    print('%-46s sass %5d inst %6.2f%% samp %6.2f%% thr/w %5.2f  | %s' % (' > '.join('%s:%d' % f for f in key), a[3], 100 * a[0] / tot[0], 100 * a[2] / max(1, tot[2]), a[1] / max(1, a[0]), top))
