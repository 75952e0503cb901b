"""SASS listing of an ncu source page (`--page source --csv --print-source cuda,sass`) in address order with the
source line, samples, executed count and the two top stall reasons of every instruction."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if len(r) > 3 and r[0] == 'Line No']
ins = []
for n, st0 in enumerate(hi):
    end = hi[n + 1] - 2 if n + 1 < len(hi) else len(rows)
    h = rows[st0]; fn = rows[st0 - 2][1].split('/')[-1]
    iS = h.index('# Samples'); iI = h.index('Instructions Executed')
    st = [(h.index(c), c[6:]) for c in h if c.startswith('stall_') and 'Not Issued' not in c]
    ln = 0
    for r in rows[st0 + 1:end]:
        if len(r) != len(h): continue
        if r[0]:
            try: ln = int(r[0])
            except Exception: pass
            continue
        try: addr = int(r[2], 16); s = int(r[iS]); ex = int(r[iI])
        except Exception: continue
        top = sorted(((int(r[i] or 0), c) for i, c in st), reverse=True)[:2]
        ins.append((addr, fn, ln, r[3].strip(), s, ex, [(c, v) for v, c in top if v > 0]))
ins.sort()
base = ins[0][0] if ins else 0
for a, fn, ln, sass, s, ex, top in ins:
    print("%05x %-14s L%-4d %-64s smp %7d ex %10d %s" % (a - base, fn[:14], ln, sass[:64], s, ex, top))
