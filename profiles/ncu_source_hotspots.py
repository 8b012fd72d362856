import csv,subprocess,io,re,sys,collections
rep=sys.argv[1]; cubin=sys.argv[2]; kern=sys.argv[3]; src=sys.argv[4]
ncu_kern=sys.argv[5] if len(sys.argv)>5 else kern   # demangled-name regex for ncu when `kern` is a mangled fragment
out=subprocess.run("ncu -i %s --page source --csv --kernel-name regex:%s" % (rep, ncu_kern),shell=True,capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
# one block of rows per captured launch of the kernel ("Kernel Name" line, header line, SASS rows); argv[6] = which block (default 0)
starts=[i for i,r in enumerate(rows) if r and r[0]=="Kernel Name"]+[len(rows)]
blk=int(sys.argv[6]) if len(sys.argv)>6 else 0
rows=rows[starts[blk]:starts[blk+1]] if len(starts)>1 else rows
hi=[i for i,r in enumerate(rows) if "Source" in r and "Address" in r][0]
hdr=rows[hi]; si=hdr.index("Warp Stall Sampling (All Samples)"); so=hdr.index("Source"); ie=hdr.index("Instructions Executed"); te=hdr.index("Thread Instructions Executed")
body=[r for r in rows[hi+1:] if len(r)>si]
# nvdisasm with line info
dis=subprocess.run("nvdisasm -g -c %s"%cubin,shell=True,capture_output=True,text=True).stdout
lines=dis.splitlines()
# find the function section for kern
insts=[]; cur=None; infn=False
for l in lines:
    if l.startswith(".text.") : infn = kern in l
    if not infn: continue
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
    if m: insts.append((cur,m.group(2)))
print("sass insts in ncu:",len(body)," in nvdisasm:",len(insts))
agg=collections.Counter(); aggi=collections.Counter(); aggt=collections.Counter()
n=min(len(body),len(insts))
for k in range(n):
    s=int(body[k][si] or 0); agg[insts[k][0]]+=s; aggi[insts[k][0]]+=int(body[k][ie] or 0); aggt[insts[k][0]]+=int(body[k][te] or 0)
tot=sum(agg.values()); toti=sum(aggi.values())
srcl=open(src).read().splitlines()
print("total samples",tot,"total warp insts",toti)
for (f,ln),v in agg.most_common(45):
    txt=srcl[ln-1].strip()[:100] if f and f.endswith(src.split('/')[-1]) and ln<=len(srcl) else ""
    print("%5.1f%% stall  %5.1f%% inst  thr/inst %4.1f  %s:%d  %s"%(100*v/tot,100*aggi[(f,ln)]/max(toti,1),aggt[(f,ln)]/max(aggi[(f,ln)],1),f,ln,txt))
