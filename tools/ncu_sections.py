import csv, collections, re, sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if len(r)>3 and r[0]=='Line No']
start=hi[0]; end=hi[1]-2
h=rows[start]; iS=h.index('# Samples'); iI=h.index('Instructions Executed'); iW=h.index('L1 Wavefronts Shared')
ib=h.index('stall_barrier')
def I(x):
    try: return int(x)
    except: return 0
src=open('/root/repo/gemotion.jl_b200/csrc/gm_gath_kernel.cuh').read().split('\n')
marks=[(i+1,l.strip()[:40]) for i,l in enumerate(src) if re.search(r'auto (issue_gd|issue_val|state|issue_lists|lists_parity|p_entries|residual|write_out) = |for \(int X = X0 - 4|if \(is_aux\) \{|// ---- role of a matrix lane|rank records of the four rounds|four independent chains|if \(info & \(1u << 6\)\)',l)]
def sec(ln):
    cur='pre'
    for m,name in marks:
        if ln>=m: cur="%d %s"%(m,name)
    return cur
reg=collections.Counter(); ins=collections.Counter(); wf=collections.Counter(); bar=collections.Counter()
for r in rows[start+1:end]:
    try: ln=int(r[0])
    except: continue
    k=sec(ln)
    reg[k]+=I(r[iS]); ins[k]+=I(r[iI]); wf[k]+=I(r[iW]); bar[k]+=I(r[ib])
T=sum(reg.values())
for k in sorted(reg,key=lambda x:int(x.split()[0]) if x!='pre' else 0): print("%-46s smp %5.1f%% (barrier %4.1f) inst/c %6.1f wf/c %6.1f"%(k,100*reg[k]/T,100*bar[k]/T, ins[k]/16777216, wf[k]/16777216))
print("total inst/cell", sum(ins.values())/16777216, "wf/cell", sum(wf.values())/16777216)
