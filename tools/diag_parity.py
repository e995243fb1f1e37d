import sys, os, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
from oracle import oracle as O
from ggufb200 import synth
from ggufb200.model import Engine
P=[1,300,301,302,303]
for preset,ftype in [("tiny","Q4_K_M"),("small","Q4_K_M"),("medium","Q4_K_M"),("tiny","Q8_0")]:
    path=f"/tmp/{preset}-{ftype}.gguf"; synth.write_gguf(path,preset,ftype,0xB200)
    ref=O.OracleLlama(path,n_ctx=128); rt,rl=ref.greedy(P,48,return_logits=True)
    for g,p in ((False,False),(True,True)):
        eng=Engine(path,n_ctx=128,use_graph=g,use_pdl=p); eng.warmup(); eng.reset(); eng.prefill(P)
        errs=[];toks=[]
        for i in range(48):
            lg=eng.last_logits(); toks.append(eng.tokens(i+1)[i])
            errs.append(float(np.abs(lg-rl[i]).max()/np.abs(rl[i]).max()))
            if i<47: eng.decode(1)
        eng.close()
        nm=sum(a!=b for a,b in zip(toks,rt))
        print(preset,ftype,'graph' if g else 'eager', 'mismatch',nm, 'err first8', ['%.1e'%e for e in errs[:8]], 'max %.2e'%max(errs), 'median %.2e'%np.median(errs), flush=True)
