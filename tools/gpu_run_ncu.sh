# usage: bash tools/gpu_run_ncu.sh <kernel regex> <skip> <count> <out name> [batch]
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B=${5:-64}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$1" -s $2 -c $3 -o /tmp/$4 python tools/profile_target.py 2 $B > gpurun_out/prof_$4.log 2>&1; tail -n 1 gpurun_out/prof_$4.log
python tools/ncu_summary.py /tmp/$4.ncu-rep > gpurun_out/$4_ncu_selected.txt
ncu -i /tmp/$4.ncu-rep --page source --csv > /tmp/$4_src.csv 2>/dev/null
python - <<PY
import csv, collections
rows=list(csv.reader(open("/tmp/$4_src.csv")))
# several kernels may follow each other: split at "Kernel Name" rows
out=open("gpurun_out/$4_source_top.txt","w")
i=0
while i < len(rows):
    if rows[i] and rows[i][0]=="Kernel Name":
        name=rows[i][1]; hdr=rows[i+1]; j=i+2
        data=[]
        while j < len(rows) and not (rows[j] and rows[j][0]=="Kernel Name"):
            data.append(rows[j]); j+=1
        isrc=hdr.index("Source"); isamp=hdr.index("Warp Stall Sampling (All Samples)"); iex=hdr.index("Instructions Executed")
        tot=sum(int(d[isamp]) for d in data if d[isamp].isdigit()); totex=sum(int(d[iex]) for d in data if d[iex].isdigit())
        out.write("== %s\n   SASS instructions %d, warp-instructions executed %d, stall samples %d\n" % (name[:120], len(data), totex, tot))
        top=sorted(range(len(data)), key=lambda k:-int(data[k][isamp]) if data[k][isamp].isdigit() else 0)[:30]
        for k in sorted(top):
            out.write("   %5d %6s %9s  %s\n" % (k, data[k][isamp], data[k][iex], data[k][isrc][:100]))
        ops=collections.Counter()
        for d in data:
            t=[x for x in d[isrc].split() if not x.startswith("@")]
            if t and d[iex].isdigit(): ops[t[0].split(".")[0]]+=int(d[iex])
        out.write("   executed by opcode: %s\n" % ", ".join("%s %d" % kv for kv in ops.most_common(14)))
        i=j
    else:
        i+=1
out.close()
PY
head -c 6000 gpurun_out/$4_ncu_selected.txt; head -n 60 gpurun_out/$4_source_top.txt
