"""Per-section breakdown of an ncu source page of k_mma (`ncu -i rep --page source --csv --print-source cuda,sass`):
samples, stall reasons, instructions and shared-memory wavefronts per cell, by the lambdas of gm_mma_kernel.cuh."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
ncells = float(sys.argv[2]) if len(sys.argv) > 2 else 16777216.0
srcfile = sys.argv[3] if len(sys.argv) > 3 else '/root/repo/gemotion.jl_b200/csrc/gm_mma_kernel.cuh'
import os
hi = [i for i, r in enumerate(rows) if len(r) > 3 and r[0] == 'Line No']
tables = []  # (file, start, end)
for n, st_ in enumerate(hi):
    tables.append((rows[st_ - 2][1] if st_ >= 2 else '', st_, hi[n + 1] - 2 if n + 1 < len(hi) else len(rows)))
h = rows[hi[0]]
def col(name):
    return h.index(name) if name in h else None
iS, iI, iW = col('# Samples'), col('Instructions Executed'), col('L1 Wavefronts Shared')
stalls = [c for c in h if c.startswith('stall_')]
def I(x):
    try: return int(x)
    except Exception: return 0
src = open(srcfile).read().split('\n')
pat = r'auto (issue_gd|issue_val|issue_lists|lists_parity|p_entries|write_out|select|tile|state_task|residual_task) = |for \(int X = X0 - 4|if \(is_aux\) \{|// main warps: lane|// ---- epilogue: nine entries|// ---- constant tables|unsigned phase = 0;'
marks = [(i + 1, l.strip()[:44]) for i, l in enumerate(src) if re.search(pat, l)]
def sec(ln):
    cur = 'pre'
    for m, name in marks:
        if ln >= m: cur = "%d %s" % (m, name)
    return cur
smp = collections.Counter(); ins = collections.Counter(); wf = collections.Counter(); st = collections.defaultdict(collections.Counter)
other = collections.Counter()
for fname, start, end in tables:
    mine = os.path.basename(fname) == os.path.basename(srcfile)
    for r in rows[start + 1:end]:
        try: ln = int(r[0])
        except Exception: continue
        k = sec(ln) if mine else "0 helpers in " + os.path.basename(fname)
        smp[k] += I(r[iS]); ins[k] += I(r[iI]); wf[k] += I(r[iW]) if iW is not None else 0
        for c in stalls: st[k][c] += I(r[h.index(c)])
T = sum(smp.values()) or 1
for k in sorted(smp, key=lambda x: int(x.split()[0]) if x != 'pre' else 0):
    top = ", ".join("%s %.1f" % (c[6:], 100 * v / T) for c, v in st[k].most_common(8) if v > 0.003 * T and 'Not Issued' not in c)
    print("%-50s smp %5.1f%%  inst/cell %6.1f  wf/cell %6.1f  | %s" % (k, 100 * smp[k] / T, ins[k] / ncells, wf[k] / ncells, top))
print("total inst/cell %.1f  shared wavefronts/cell %.1f" % (sum(ins.values()) / ncells, sum(wf.values()) / ncells))
tot = collections.Counter()
for k in st:
    tot.update(st[k])
print("stall reasons (%% of samples): " + ", ".join("%s %.1f" % (c[6:], 100 * v / T) for c, v in tot.most_common(24) if 'Not Issued' not in c))
