"""One-off fuzz (test infrastructure): optimistic links / verification / repair on the kernel emulator against the oracle.
    python tests/fuzz_repair.py <seed> <n_cases>
Random genome sets (SNPs, reverse-complemented and rotated copies), stored hash bits masked at random, random bucket sizes
and thread orders.  A seeded subset runs in the suite (tests/test_emu_parity.py::test_emu_optimistic_links_fuzz)."""
import sys,ctypes,os,random
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/oracle')
import numpy as np
import dbg_oracle, digests
from superpang_b200 import _capi
from superpang_b200.assembler import Assembler
lib=_capi._declare(ctypes.CDLL('/root/repo/tests/emu/libspb_emu.so'))
rnd=random.Random(int(sys.argv[1]))
fails=0; ran=0; deg=0
for it in range(int(sys.argv[2])):
    k=rnd.choice([17,21,31,33,45])
    n=rnd.randint(2,5); L=rnd.randint(200,2500); p=rnd.choice([0.0,0.01,0.03,0.1])
    rng=np.random.default_rng(rnd.randint(0,1<<30))
    anc=rng.integers(0,4,L)
    sd={}
    for g in range(n):
        x=anc.copy(); m=rng.random(L)<p; x[m]=(x[m]+rng.integers(1,4,int(m.sum())))&3
        if rnd.random()<0.3: x=(3-x)[::-1]                      # reverse complement of a genome
        if rnd.random()<0.3:
            a=rnd.randint(0,L-50); x=np.concatenate([x[a:],x[:a]])   # rotation
        sd[f"g{g}"]="".join("ACGT"[v] for v in x)
    env=dict(SPB_DEDUP="part", SPB_PART_AVG=str(rnd.choice([20,60,300])), SPB_D_HMASK=rnd.choice(["FFFFFFFF","FFF00000","FF000000","FFFF0000","F8000000"]), SPB_EMU_ORDER=rnd.choice(["","reverse","shuffle"]))
    for a,b in env.items(): os.environ[a]=b
    try:
        want=dbg_oracle.run_oracle(sd,k)
    except AssertionError:
        deg+=1
        try:
            Assembler(sd,k,engine=_capi.Engine(lib=lib)).to_result(); print("ACCEPTED DEGENERATE", it, env); fails+=1
        except _capi.DegenerateInputError: pass
        continue
    try:
        got=Assembler(sd,k,engine=_capi.Engine(lib=lib)).to_result()
        digests.assert_same(want,got); ran+=1
    except Exception as e:
        fails+=1; print("FAIL", it, k, n, L, p, env, type(e).__name__, str(e)[:200])
print("ran",ran,"degenerate",deg,"fails",fails)
