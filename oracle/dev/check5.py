import sys, time, base64, os
sys.path.insert(0, '/root/repo/oracle')
import numpy as np
import api, chacha_air as ca, ref_wasm as r
rng = np.random.default_rng(1)
for nb, seed in [(2,0),(16,1),(17,2),(40,3),(64,4)]:
    rng = np.random.default_rng(seed)
    key = rng.bytes(32); nonce = rng.bytes(12); counter = int(rng.integers(0, 2**31))
    pt = rng.bytes(64*nb)
    ks = ca.chacha20_keystream_bytes(key, nonce, counter, nb)
    ct = bytes(a ^ b for a, b in zip(pt, ks))
    t=time.time(); ref = r.generate_chacha20_proof(key, nonce, counter, pt, ct); t1=time.time()-t
    t=time.time(); mine = api.generate_chacha20_proof(key, nonce, counter, pt, ct); t2=time.time()-t
    if 'error' in ref:
        print(nb, 'ref error', ref, 'mine', mine.get('error')); continue
    print(nb, 'ref %.2fs mine %.2fs'%(t1,t2), 'equal', mine['proof']==ref['proof'], len(ref['proof']))
