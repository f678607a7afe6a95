import sys, pickle, itertools
sys.path.insert(0, '/root/repo/oracle')
import numpy as np
from stwo_core import *
d = pickle.load(open('/tmp/dbg.pkl','rb'))
lde = np.load('/tmp/cc1_lde.npy')
cols = np.concatenate([lde, np.stack(d['comp_evals'])], axis=0)
vals = [v for tv in d['sampled'] for cv in tv for v in cv]
assert len(vals) == cols.shape[0]
ref = {13:(2006172302, 1044852000, 1260148445, 1833503757), 14:(1913690960, 978855406, 1146440054, 1086044268), 21:(40711492, 80681998, 917949800, 573445707)}
rc = d['random_coeff']; px, py = d['oods']
dom = canonic_domain(5); xs, ys = dom.points_bitrev()
row = 13
x, y = int(xs[row]), int(ys[row])
c = py.conj() - py
terms = []
for k, val in enumerate(vals):
    a = val.conj() - val
    b = val * c - a * py
    terms.append(c * int(cols[k,row]) - (a * y + b))
# denominators
prx, pix = (px.v[0], px.v[1]), (px.v[2], px.v[3])
pry, piy = (py.v[0], py.v[1]), (py.v[2], py.v[3])
den = c_sub(c_mul(c_sub(prx,(x,0)), piy), c_mul(c_sub(pry,(y,0)), pix))
di = c_inv(den)
def mulcm(q, cm):
    lo = c_mul(q.v[:2], cm); hi = c_mul(q.v[2:], cm); return QM31(lo[0],lo[1],hi[0],hi[1])
n = len(terms)
def comb(exps):
    acc = QM31(0)
    return acc
# variant A: alpha^(k+1); B: alpha^k ; C: alpha^(n-k) ; D: alpha^(n-1-k)
pw = [QM31(1)]
for i in range(n+1): pw.append(pw[-1]*rc)
for name, f in [('k+1', lambda k: k+1), ('k', lambda k: k), ('n-k', lambda k: n-k), ('n-1-k', lambda k: n-1-k)]:
    acc = QM31(0)
    for k,t in enumerate(terms): acc = acc + pw[f(k)]*t
    print(name, mulcm(acc, di).v, ref[row])
