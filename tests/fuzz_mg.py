"""One-off fuzz (test infrastructure): the sliced multi-rank tail over gloo on the kernel emulator.
    python tests/fuzz_mg.py <seed> <world>
Ten random genome sets per run; every rank's result, the concatenation of the sliced parts and the shared host buffer must
equal the single-engine result (checks inside tests/mg_worker.py)."""
import sys,ctypes,os,random,json,subprocess,socket,tempfile
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/oracle')
import numpy as np
import dbg_oracle, digests
from superpang_b200 import _capi
from superpang_b200.assembler import Assembler
lib=_capi._declare(ctypes.CDLL('/root/repo/tests/emu/libspb_emu.so'))
rnd=random.Random(int(sys.argv[1])); world=int(sys.argv[2])
cases=[]; want=[]
while len(cases)<10:
    k=rnd.choice([21,31,45])
    n=rnd.randint(3,6); L=rnd.randint(1500,5000); p=rnd.choice([0.0,0.01,0.03,0.08])
    rng=np.random.default_rng(rnd.randint(0,1<<30))
    anc=rng.integers(0,4,L); sd={}
    for g in range(n):
        x=anc.copy(); m=rng.random(L)<p; x[m]=(x[m]+rng.integers(1,4,int(m.sum())))&3
        if rnd.random()<0.3: x=(3-x)[::-1]
        if rnd.random()<0.3:
            a=rnd.randint(0,L-50); x=np.concatenate([x[a:],x[:a]])
        if rnd.random()<0.2: x=x[:rnd.randint(5,k)]              # a sequence shorter than k in between
        sd[f"g{g}"]="".join("ACGT"[v] for v in x)
    try: dbg_oracle.run_oracle(sd,k)
    except AssertionError: continue
    A=Assembler(sd,k,engine=_capi.Engine(lib=lib))
    want.append(digests.summarize(A.to_result())); cases.append(dict(records=list(sd.items()),k=k))
td=tempfile.mkdtemp(); cp=os.path.join(td,'cases.json'); json.dump(cases,open(cp,'w')); out=os.path.join(td,'out.json')
s=socket.socket(); s.bind(('127.0.0.1',0)); port=str(s.getsockname()[1]); s.close()
env=dict(os.environ, SPB_MG_PROTOCOL="part", SPB_PART_AVG=str(rnd.choice([20,40,100])), SPB_MG_ALL_LINKS="2", SPB_EXPECT_SLICED="1", SPB_EMU_ORDER=rnd.choice(["","reverse","shuffle"]))
procs=[subprocess.Popen([sys.executable,'/root/repo/tests/mg_worker.py',str(r),str(world),port,cp,out],env=env,stderr=subprocess.DEVNULL) for r in range(world)]
rcs=[p.wait(timeout=900) for p in procs]
ok = rcs==[0]*world and all(json.load(open(f"{out}.{r}"))==want for r in range(world))
print('seed',sys.argv[1],'world',world,'rcs',rcs,'ok',ok)
