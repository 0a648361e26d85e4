"""Summarises an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch
list: the launches from the LAST occurrence of <first kernel> up to the first of <stop kernels>.
Usage: ncu_launch_summary.py launches.csv first_kernel n_primitives cuda_event_ms"""
import collections, csv, sys
path, first, n_prims, event_ms = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
rows = list(csv.reader(open(path)))
hdr = None; data = []
for r in rows:
    if r and r[0] == 'ID': hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(dict(zip(hdr, r)))
by = collections.OrderedDict()
for d in data:
    k = d['ID']; by.setdefault(k, {'name': d['Kernel Name']})
    by[k][d['Metric Name']] = (float(d['Metric Value'].replace(',', '')), d['Metric Unit'])
ids = list(by.keys())
start = ids.index([i for i in ids if first in by[i]['name']][-1])
agg = collections.OrderedDict(); tot = 0; totb = 0
for i in ids[start:]:
    b = by[i]; nm = b['name'].split('(')[0].split('<')[0].replace('void ', '')[:44]
    if any(s in nm for s in ('render_kernel', 'aov_kernel', 'resolve_kernel')): break
    u = b['gpu__time_duration.sum'][0] / 1e3
    mb = lambda k: b[k][0] * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}[b[k][1]]
    r, w = mb('dram__bytes_read.sum'), mb('dram__bytes_write.sum')
    a = agg.setdefault(nm, [0, 0, 0, 0]); a[0] += u; a[1] += r; a[2] += w; a[3] += 1
    tot += u; totb += r + w
print("ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none")
print(f"last upload of the run; {n_prims} primitives")
print(f"per-launch times under ncu are cold-cache and serialised; CUDA-event time of the same build not under ncu: {event_ms} ms\n")
print(f"{'kernel':46s} {'n':>2s} {'us':>8s} {'DRAM rd MB':>11s} {'DRAM wr MB':>11s} {'GB/s':>7s} {'share':>6s}")
for nm, a in agg.items():
    print(f"{nm:46s} {a[3]:2d} {a[0]:8.1f} {a[1]:11.1f} {a[2]:11.1f} {(a[1] + a[2]) / a[0] * 1e3:7.0f} {100 * a[0] / tot:5.1f}%")
print(f"{'total':46s}    {tot:8.1f} {'':>11s} {'':>11s} {totb / tot * 1e3:7.0f}")
print("\nDRAM traffic per primitive: %.0f B; HBM peak (MEASURED_PEAKS.json): 6550 GB/s" % (totb * 1e6 / n_prims))
