import sys, subprocess, re
sys.path.insert(0,'/root/repo')
from sterne_b200 import build as b
def mk(name, defs):
    out='/root/repo/build_variants/lib_%s.so'%name
    cmd=[b.nvcc_path()]+b.NVCC_FLAGS+['-Xptxas','-v']+['-D%s'%d for d in defs]+['-o',out,b.SRC]
    r=subprocess.run(cmd,capture_output=True,text=True)
    txt=r.stderr
    if r.returncode: print(name,'FAILED',txt[-2000:]); return
    # find R big kernel info
    blocks=re.split(r'ptxas info\s+: Compiling entry function',txt)
    for blk in blocks:
        m=re.search(r"stn_loglike_fast_kernelILi(\d)ELi1E",blk)
        if m and int(m.group(1))>=2:
            u=re.search(r'Used (\d+) registers',blk); sp=re.search(r'(\d+) bytes spill stores, (\d+) bytes spill loads',blk)
            print(name,'R=%s'%m.group(1),'regs',u.group(1),'spill',sp.group(1),sp.group(2))
if __name__=='__main__':
    import json
    for name,defs in json.loads(sys.argv[1]).items(): mk(name,defs)
