import csv, re, sys, collections
# SASS listing with line markers
cur=None; order=[]   # (addr_offset, line, inlined_chain)
for l in open('/tmp/cub/solve.sass'):
    m=re.search(r'//## File "([^"]+)", line (\d+)(.*)',l)
    if m:
        f=m.group(1); ln=int(m.group(2)); rest=m.group(3)
        cur=(f.split('/')[-1],ln,rest.strip()); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);',l)
    if m: order.append((int(m.group(1),16),cur,m.group(2)))
rows=list(csv.reader(open(sys.argv[1])))
H=rows[1]; ai=H.index('Address'); ii=H.index('Instructions Executed'); ss=H.index('# Samples'); si=H.index('Source')
base=None; agg=collections.defaultdict(lambda:[0,0]); tot=[0,0]
sass=[]
for r in rows[2:]:
    a=int(r[ai],16)
    if base is None: base=a
    sass.append((a-base,int(r[ii] or 0),int(r[ss] or 0),r[si]))
addr2line={o:c for o,c,_ in order}
for off,ins,smp,txt in sass:
    c=addr2line.get(off)
    key=(c[0],c[1]) if c else ('?',0)
    agg[key][0]+=ins; agg[key][1]+=smp; tot[0]+=ins; tot[1]+=smp
src={}
for fn in ['nls_kernels.cu']:
    for i,l in enumerate(open('/root/repo/nlp_filter_b200/csrc/'+fn),1): src[(fn,i)]=l.rstrip()
print('total warp-inst',tot[0],'samples',tot[1])
top=sorted(agg.items(),key=lambda kv:-kv[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 40]
for k,(ins,smp) in top:
    print('%5.1f%% samp %5.1f%% inst  %s:%d  %s'%(100*smp/tot[1],100*ins/tot[0],k[0],k[1],src.get(k,'')[:100].strip()))
# by phase (line ranges)
