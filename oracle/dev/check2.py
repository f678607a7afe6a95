import sys, time
sys.path.insert(0, '/root/repo/oracle')
import api, chacha_air as ca
key = bytes(range(32)); nonce = bytes([0,0,0,9,0,0,0,0x4a,0,0,0,0])
pt = bytes((i*7) & 0xff for i in range(64))
ks = ca.chacha20_keystream_bytes(key, nonce, 1, 1)
ct = bytes(a ^ b for a, b in zip(pt, ks))
ref = open('/tmp/cc1.bin','rb').read()
t=time.time(); dbg={}
out = api.generate_chacha20_proof(key, nonce, 1, pt, ct, debug=dbg)
print(time.time()-t, out.get('error'))
mine = out['proof_bytes']
print(len(mine), len(ref), mine == ref)
if mine != ref:
    for i,(a,b) in enumerate(zip(mine,ref)):
        if a!=b: print('first diff at', i); break
    print('nonce', dbg['nonce'], 'queries', dbg['queries'])
import struct
def find(b, name):
    i = mine.find(b); print(name, 'in mine at', i)
print('ref first layer commitment a772000f... in mine?', mine.find(bytes.fromhex('a772000f8c12ae32e246ba5dfba85e64bdcac0089e8265fe4e06f78b629a0e9c')))
print('ref inner 9b90 in mine?', mine.find(bytes.fromhex('9b90b19c31f47c850ae5ad85c1e9c4d77215b2eac4bdb788bcc14bdc0c81642e')))
print('ref last coeffs', mine.find(struct.pack('<4I',438163532, 55359895, 754968638, 1795397974)))
