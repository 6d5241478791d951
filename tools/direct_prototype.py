import numpy as np, sys, time
sys.path.insert(0, __import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), '..'))
from oracle import qtet_oracle as O

def direct_vecs(a, b, lam):
    """a [B,d], b [d-1], lam [B,d] -> V [B,d(k),d(i)] by two-sided recurrences joined at argmax|p q|"""
    B, d = a.shape
    p = np.zeros((B, d, d)); q = np.zeros((B, d, d))   # [B,k,i]
    p[:,0,:] = 1.0
    p[:,1,:] = (lam - a[:,0:1]) / b[0]
    for k in range(1, d-1):
        p[:,k+1,:] = ((lam - a[:,k:k+1]) * p[:,k,:] - b[k-1]*p[:,k-1,:]) / b[k]
    q[:,d-1,:] = 1.0
    q[:,d-2,:] = (lam - a[:,d-1:d]) / b[d-2]
    for k in range(d-2, 0, -1):
        q[:,k-1,:] = ((lam - a[:,k:k+1]) * q[:,k,:] - b[k]*q[:,k+1,:]) / b[k-1]
    pq = np.abs(p*q)
    r = np.argmax(pq, axis=1)     # [B,i]
    kk = np.arange(d)[None,:,None]
    pr = np.take_along_axis(p, r[:,None,:], 1); qr = np.take_along_axis(q, r[:,None,:], 1)
    z = np.where(kk <= r[:,None,:], p/pr, q/qr)
    z /= np.linalg.norm(z, axis=1, keepdims=True)
    return z, r

def run(N, chis, T=30, max_t=25):
    const = {"max_N": N, "sites": 2, "max_t": max_t, "omegas": [-3,3], "chis":[0,0], "coupling":1, "timesteps":T}
    states = O.fock_basis(N,2)
    off = O.offdiag_matrix(states, 1.0)
    b = np.array([off[k+1,k] for k in range(N)])
    a = O.diag_energies(states, [-3,3], chis)
    a = a - a.mean(axis=1, keepdims=True)
    Hb = np.zeros((len(chis), N+1, N+1)); idx=np.arange(N+1)
    Hb[:,idx,idx]=a; Hb[:,idx[:-1],idx[1:]]=b; Hb[:,idx[1:],idx[:-1]]=b
    lam, Vref = np.linalg.eigh(Hb)
    V, r = direct_vecs(a, b, lam)
    # orthogonality
    G = np.einsum('bki,bkj->bij', V, V) - np.eye(N+1)
    orth = np.abs(G).max(axis=(1,2))
    res = np.abs(np.einsum('bkl,bli->bki', Hb, V) - V*lam[:,None,:]).max(axis=(1,2))
    # trace using V
    t = np.linspace(0, max_t, T)
    def trace(V, lam):
        c = V[:,0,:]
        ph = np.exp(-1j*lam[:,None,:]*t[None,:,None])   # [B,T,i]
        psi = np.einsum('bki,bti->btk', V, c[:,None,:]*ph)
        return (np.abs(psi)**2 * states[None,None,:,1]).sum(-1)
    tr = trace(V, lam); tr0 = trace(Vref, lam)
    gaps = np.diff(lam, axis=1).min(axis=1)
    # weighted gap: only pairs with c_i c_j significant
    c = Vref[:,0,:]
    w = np.abs(c[:,1:]*c[:,:-1]) / np.diff(lam,axis=1)
    return dict(orth=orth, res=res, dtr=np.abs(tr-tr0).max(axis=1), gaps=gaps, w=w.max(axis=1), norm=np.abs(a).max(axis=1))

if __name__ == "__main__":
    N = 20
    g = np.linspace(-10,10,1024)
    # sample: a 96x96 subgrid + special lines
    ii = np.arange(0,1024,11)
    X, Y = np.meshgrid(g[ii], g[ii], indexing='ij')
    chis = np.stack([X.ravel(), Y.ravel()],1)
    t0=time.time(); out = run(N, chis); print("time", time.time()-t0, "n", len(chis))
    for k in ("orth","res","dtr"):
        v = out[k]; print(k, "max %.3e  p99 %.3e p999 %.3e median %.3e" % (v.max(), np.quantile(v,.99), np.quantile(v,.999), np.median(v)))
    bad = out["dtr"] > 2e-9
    print("frac dtr>2e-9:", bad.mean(), " frac >1e-10:", (out["dtr"]>1e-10).mean())
    for thr in (1e-1,1e-2,1e-3,1e-4,1e-6):
        print("gap<%g frac %.4f ; of bad points covered: %.3f" % (thr, (out["gaps"]<thr).mean(), (out["gaps"][bad]<thr).mean() if bad.any() else -1))
    np.save('/tmp/proto/out.npy', out)
