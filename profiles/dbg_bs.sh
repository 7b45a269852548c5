python profiles/dbg_bs.py gpurun_out/dbg_768.json
SMCB_LIB=$PWD/particles_b200/variants/libsmcb_bs512.so python profiles/dbg_bs.py gpurun_out/dbg_512.json
python - <<P
import json, torch, numpy as np
a=json.load(open("gpurun_out/dbg_768.json")); b=json.load(open("gpurun_out/dbg_512.json"))
ra=np.array(a["rows"]); rb=np.array(b["rows"])
d=np.abs(ra[:,1]-rb[:,1]); print("first step with |dlogLt|>1e-9:", int(np.argmax(d>1e-9)) if (d>1e-9).any() else None, d[:12])
print("rs flags equal", np.array_equal(ra[:,2], rb[:,2]), "first rs", int(np.argmax(ra[:,2]>0)))
i=int(np.argmax(d>1e-12)); print("first diverging step", i, ra[max(0,i-2):i+3].tolist(), rb[max(0,i-2):i+3].tolist())
xa=np.array(a["xsum"]); xb=np.array(b["xsum"]); j=int(np.argmax(np.abs(xa-xb)>1e-6)); print("first xsum diff at", j, xa[j-1:j+2], xb[j-1:j+2])
P
