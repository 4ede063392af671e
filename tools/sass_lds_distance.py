"""Offline check of software pipelining in the unrolled DP bodies: for every branch-free run of a kernel that holds >= 32 LDS.64
(one steady-state column), the distance in instructions between each row-entry load and the first use of its result.
usage: python tools/sass_lds_distance.py <cubin or .so> <substring of the mangled kernel name>"""
import re, sys, subprocess
# usage: lds_dist.py cubin function-substring
out = subprocess.run(["cuobjdump","-sass",sys.argv[1]],capture_output=True,text=True).stdout
parts = out.split("Function : ")
for part in parts[1:]:
    name = part.split("\n")[0]
    if sys.argv[2] not in name: continue
    ins=[]
    for l in part.split("\n"):
        m=re.match(r"^\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m: ins.append((int(m.group(1),16), m.group(2).strip()))
    run=[]; bodies=[]
    for a,t in ins:
        op=t.split()[0] if not t.startswith('@') else t.split()[1]
        if op.startswith(('BRA','BRX','BSSY','BSYNC','BREAK','EXIT','RET','CALL')):
            if len(run)>=250 and sum(1 for _,x in run if 'LDS.64' in x)>=32: bodies.append(run)
            run=[]
        else: run.append((a,t))
    print(name[:70], "instr", len(ins))
    for body in bodies:
        d=[]
        for idx,(a,t) in enumerate(body):
            if 'LDS.64' in t:
                m=re.search(r"LDS\.64 (R\d+),", t)
                r=int(m.group(1)[1:]); regs=[f"R{r}", f"R{r+1}"]
                for j in range(idx+1,len(body)):
                    ops=body[j][1].split(',',1)
                    if len(ops)>1 and any(re.search(r"\b%s\b"%x, ops[1]) for x in regs):
                        d.append(j-idx); break
        h={}
        for _,t in body:
            op=t.split()[0]; h[op]=h.get(op,0)+1
        print("  steady body len", len(body), "mean LDS->use", sum(d)/max(len(d),1), "min", min(d), "moves", sum(v for k,v in h.items() if 'MOV' in k))
