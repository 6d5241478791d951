import numpy as np, sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.abspath(__file__)))
from direct_prototype import run
from multiprocessing import Pool
g = np.linspace(-10,10,1024)
def work(i0):
    X, Y = np.meshgrid(g[i0:i0+16], g, indexing='ij')
    chis = np.stack([X.ravel(), Y.ravel()],1)
    o = run(20, chis)
    return {k: v for k,v in o.items()}
if __name__=="__main__":
    with Pool(8) as p:
        outs = p.map(work, range(0,1024,16))
    out = {k: np.concatenate([o[k] for o in outs]) for k in outs[0]}
    np.savez('/tmp/proto/full.npz', **out)
    for k in ("orth","res","dtr","gaps"):
        v = out[k]; print(k, "max %.3e min %.3e p999 %.3e p9999 %.3e median %.3e" % (v.max(), v.min(), np.quantile(v,.999), np.quantile(v,.9999), np.median(v)))
    for thr in (1e-9,1e-10,1e-11,1e-12):
        print("dtr>%g: %d" % (thr, (out["dtr"]>thr).sum()))
    for thr in (1,1e-1,1e-2,1e-3,1e-4,1e-6):
        print("gap<%g: %d" % (thr, (out["gaps"]<thr).sum()))
    i = np.argsort(-out["dtr"])[:10]
    for j in i: print(j//1024, j%1024, g[j//1024], g[j%1024], out["dtr"][j], out["gaps"][j], out["orth"][j])
