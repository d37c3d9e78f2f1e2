import csv,io,subprocess,sys,re
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
cur=None;hdr=None;rows=[]
for r in csv.reader(io.StringIO(out)):
    if len(r)==2 and r[0]=="File Path": cur=r[1];hdr=None
    elif r and r[0]=="Line No": hdr=r
    elif hdr and len(r)>=9 and r[0]:
        try: rows.append((cur.split('/')[-1],int(r[0]),int(r[7] or 0),int(r[6] or 0),int(r[8] or 0)))
        except ValueError: pass
src=open('slacken_b200/csrc/slk_core.h').read().split('\n')
def find(pat):
    for i,l in enumerate(src):
        if pat in l: return i+1
    raise Exception(pat)
marks=[('hdr/encode',1),('compress',find('slk_compress_fast(uint64_t x)')-1),('hash',find('SLK_HD uint32_t slk_mulhi32')),('next_bucket',find('SLK_HD uint64_t slk_next_bucket')-4),('bucket match',find('struct slk_bucket')),('probe_rest',find('slk_probe_rest')-1),('lca/tax',find('// LowestCommonAncestor.apply')),('scanner',find('SLK_HD uint64_t slk_min64')),('read_src..',find('struct slk_read_src')-5),('resolve_tree',find('SLK_HD uint32_t slk_resolve_tree')-3),('fast_hist',find('struct slk_fast_hist')-2),('store',find('struct slk_store_local')-6),('hist',find('SLK_HD uint32_t size() const { return nk; }')),('spill',find('SLK_HD_NOINLINE void spill')-1),('resolve_slow',find('SLK_HD_NOINLINE uint32_t resolve_slow')-1),('run init',find('SLK_HD void run(')-1),('push_hit',find('auto push_hit')),('close: match pass',find('auto close = [&]')),('close: issue pass',find('s < n_cur; s += SLK_LANES')),('close: pending pass',find('for (uint32_t q = lane; q < n_pend; q += SLK_LANES)')),('close: merge pass',find('uint32_t s = head_prev;')),('append',find('auto append = ')),('mate loop',find('for (int mt = 0; mt <')),('step',find('auto step = [&]')),('fstep',find('auto fstep = [&]')),('scan_block',find('auto scan_block = [&]')),('packed load',find('if (PACKED) {')),('ascii load',find('// 16-byte chunks over')),('mate end',find('// mate end:')),('tail',find('for (int e = 0; e < 2; e++) close();')),('pack_read etc',find('// K1 as a stand-alone step'))]
tot=sum(r[2] for r in rows); ts=sum(r[3] for r in rows)
agg={}
for f,ln,ie,sm,ti in rows:
    if f!='slk_core.h': reg=f
    else:
        c=[m for m in sorted(marks,key=lambda m:m[1]) if m[1]<=ln]; reg=c[-1][0] if c else 'hdr'
    a=agg.setdefault(reg,[0,0,0]); a[0]+=ie;a[1]+=sm;a[2]+=ti
nw=float(sys.argv[2]) if len(sys.argv)>2 else 62500
for k,(ie,sm,ti) in sorted(agg.items(),key=lambda x:-x[1][0]):
    print(f"{k:22s} {100*ie/tot:5.1f}% inst {ie/nw:8.0f}/warp  lanes {ti/max(ie,1):5.1f}  {100*sm/ts:5.1f}% smp")
print('total per warp',tot/nw)
