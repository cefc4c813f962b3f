"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line: python tools/ncu_lines.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Instructions Executed' in r][0]
hdr = rows[hi]; ix = {}
for i, h in enumerate(hdr): ix.setdefault(h, i)
agg = []
for r in rows[hi + 1:]:
    if len(r) > ix['Thread Instructions Executed'] and r[0] != '' and r[2] == '-':
        try:
            agg.append((int(r[ix['# Samples']]), int(r[ix['Instructions Executed']]), int(r[ix['Thread Instructions Executed']]), r[0], r[1].strip()[:110],
                        int(r[ix['stall_long_sb']] or 0), int(r[ix['stall_barrier']] or 0), int(r[ix['stall_wait']] or 0), int(r[ix['stall_membar']] or 0), int(r[ix['stall_no_inst']] or 0), int(r[ix['stall_short_sb']] or 0)))
        except Exception:
            pass
ts = sum(a[0] for a in agg); ti = sum(a[1] for a in agg); tt = sum(a[2] for a in agg)
print('samples', ts, 'warp inst', ti, 'lanes/inst %.1f' % (tt / ti))
for name, k in (('long_sb', 5), ('barrier', 6), ('wait', 7), ('membar', 8), ('no_inst', 9), ('short_sb', 10)):
    print('  %-8s %5.1f%%' % (name, 100 * sum(a[k] for a in agg) / ts))
agg.sort(reverse=True)
for a in agg[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{100*a[0]/ts:5.1f}% (lsb {100*a[5]/ts:4.1f} bar {100*a[6]/ts:4.1f} wait {100*a[7]/ts:4.1f} mb {100*a[8]/ts:4.1f}) {100*a[1]/ti:5.1f}% inst lanes {a[2]/max(a[1],1):5.1f} L{a[3]:>4} {a[4]}")
