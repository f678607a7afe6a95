import struct
def parse(b, hdr):
    pos = [hdr]
    def u64():
        v = struct.unpack_from('<Q', b, pos[0])[0]; pos[0] += 8; return v
    out = {}
    out['config'] = b[pos[0]:pos[0]+25]; pos[0] += 25
    n = u64(); out['commitments'] = [b[pos[0]+32*i:pos[0]+32*i+32] for i in range(n)]; pos[0] += 32*n
    nt = u64(); sv = []
    for t in range(nt):
        nc = u64(); tv = []
        for c in range(nc):
            k = u64(); tv.append(b[pos[0]:pos[0]+16*k]); pos[0] += 16*k
        sv.append(tv)
    out['sampled'] = sv
    nt = u64(); dec = []
    for t in range(nt):
        k = u64(); dec.append(b[pos[0]:pos[0]+32*k]); pos[0] += 32*k
    out['decommit'] = dec
    nt = u64(); qv = []
    for t in range(nt):
        nc = u64(); tv = []
        for c in range(nc):
            k = u64(); tv.append(b[pos[0]:pos[0]+4*k]); pos[0] += 4*k
        qv.append(tv)
    out['queried'] = qv
    out['pow'] = u64()
    def layer():
        k = u64(); w = b[pos[0]:pos[0]+16*k]; pos[0] += 16*k
        h = u64(); d = b[pos[0]:pos[0]+32*h]; pos[0] += 32*h
        c = b[pos[0]:pos[0]+32]; pos[0] += 32
        return dict(witness=w, decommit=d, commitment=c)
    out['first'] = layer()
    ni = u64(); out['inner'] = [layer() for _ in range(ni)]
    k = u64(); out['last'] = b[pos[0]:pos[0]+16*k]; pos[0] += 16*k
    out['end'] = pos[0] + 4
    return out
