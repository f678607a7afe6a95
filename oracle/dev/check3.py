import sys, time, pickle
sys.path.insert(0, '/root/repo/oracle')
import numpy as np
import api, chacha_air as ca
from stwo_core import *
key = bytes(range(32)); nonce = bytes([0,0,0,9,0,0,0,0x4a,0,0,0,0])
pt = bytes((i*7) & 0xff for i in range(64))
ks = ca.chacha20_keystream_bytes(key, nonce, 1, 1)
ct = bytes(a ^ b for a, b in zip(pt, ks))
dbg={}
out = api.generate_chacha20_proof(key, nonce, 1, pt, ct, debug=dbg)
dd={k:dbg[k] for k in ('sampled','random_coeff','oods','quot')}; dd['comp_evals']=dbg['scheme'].trees[2].evals; pickle.dump(dd, open('/tmp/dbg.pkl','wb'))
