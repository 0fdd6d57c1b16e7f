import csv,re,collections,sys
src=sys.argv[1]
sass=open('/tmp/t/elf2/all.sass').read().split('\n')
key=sys.argv[2] if len(sys.argv)>2 else 'attn_fwd4_kernelI13__nv_bfloat16S1_Li96ELb0'
inside=False; cur=('?',0); m={}; order=[]
for ln in sass:
    if ln.startswith('//--------------------- .text.'):
        inside = key in ln; continue
    if not inside: continue
    mm=re.search(r'//## File "([^"]+)", line (\d+)',ln)
    if mm: cur=(mm.group(1).split('/')[-1],int(mm.group(2))); continue
    mm=re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);',ln)
    if mm: m[int(mm.group(1),16)]=(cur,mm.group(2)); order.append(int(mm.group(1),16))
b=[o for o in order if 'USETMAXREG' in m[o][1]]
rows=list(csv.reader(open(src)))
hdr=rows[1]; idx={n:i for i,n in enumerate(hdr)}
data=rows[2:]; base=int(data[0][0],16)
names=['common prologue','softmax+epilogue','staging','mma/producer/tail']
def role(off):
    for i,x in enumerate(b):
        if off<x: return names[i]
    return names[len(b)]
stall=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg=collections.defaultdict(collections.Counter); ni=collections.Counter(); top=collections.defaultdict(list)
for r in data:
    off=int(r[0],16)-base; ro=role(off)
    ni[ro]+=int(r[idx['Instructions Executed']])
    s=int(r[idx['# Samples']]); agg[ro]['samples']+=s
    top[ro].append((s,off,m.get(off,(('?',0),'?'))))
    for h in stall:
        v=r[idx[h]]
        if v and v!='0': agg[ro][h]+=int(v)
tot=sum(c['samples'] for c in agg.values())
for ro,c in agg.items():
    print('%s: warp-instr %d samples %d (%.1f%%)'%(ro,ni[ro],c['samples'],100*c['samples']/tot))
    print('     '+', '.join('%s %d'%(h.replace('stall_',''),v) for h,v in c.most_common(8) if h!='samples'))
want=sys.argv[3] if len(sys.argv)>3 else 'staging'
for s,off,((f,l),ins) in sorted(top[want],reverse=True)[:int(sys.argv[4]) if len(sys.argv)>4 else 16]:
    print('   %5d  %6x %s:%d  %s'%(s,off,f,l,ins[:60]))
