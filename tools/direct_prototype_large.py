import numpy as np, sys, time
sys.path.insert(0, __import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), '..'))
from oracle import qtet_oracle as O
from scipy.linalg import hessenberg, eigh_tridiagonal
np.seterr(all='ignore')

def direct_vecs_1(a, b, lam):
    d = len(a)
    p = np.zeros((d, d)); q = np.zeros((d, d))   # [k,i]
    sp = np.zeros((d,d)); 
    p[0] = 1.0
    p[1] = (lam - a[0]) / b[0]
    for k in range(1, d-1):
        p[k+1] = ((lam - a[k]) * p[k] - b[k-1]*p[k-1]) / b[k]
    q[d-1] = 1.0
    q[d-2] = (lam - a[d-1]) / b[d-2]
    for k in range(d-2, 0, -1):
        q[k-1] = ((lam - a[k]) * q[k] - b[k]*q[k+1]) / b[k-1]
    pq = np.abs(p*q)
    r = np.argmax(pq, axis=0)
    kk = np.arange(d)[:,None]
    pr = np.take_along_axis(p, r[None,:], 0); qr = np.take_along_axis(q, r[None,:], 0)
    z = np.where(kk <= r[None,:], p/pr, q/qr)
    z /= np.linalg.norm(z, axis=0, keepdims=True)
    return z, (np.abs(p).max(), np.abs(q).max())

def gs_fix(z, lam, thr):
    d = len(lam); n=0
    start = 0
    for i in range(1, d):
        if lam[i]-lam[i-1] < thr:
            for j in range(start, i):
                z[:,i] -= (z[:,j]@z[:,i])*z[:,j]
            z[:,i] /= np.linalg.norm(z[:,i]); n+=1
        else:
            start = i
    return n

def study(const, chis, thr_rel):
    states = O.fock_basis(const["max_N"], const["sites"])
    f = const["sites"]; N=const["max_N"]
    t = np.linspace(0, 25, 30)
    res=[]
    for x in chis:
        H = O.hamiltonian(const, x)
        d = H.shape[0]
        H = H - np.eye(d)*np.trace(H)/d
        if f == 2:
            a = np.diag(H).copy(); b = np.diag(H,1).copy(); Q = np.eye(d)
        else:
            Tm, Q = hessenberg(H, calc_q=True)
            a = np.diag(Tm).copy(); b = np.diag(Tm,1).copy()
        lam = eigh_tridiagonal(a, b, eigvals_only=True)
        nrm = np.abs(a).max() + 2*np.abs(b).max()
        z, growth = direct_vecs_1(a, b, lam)
        ok = np.isfinite(z).all()
        gap = np.diff(lam).min()
        V0 = Q @ z
        ngs = gs_fix(z, lam, thr_rel*nrm)
        V1 = Q @ z
        lam_ref, Vref = np.linalg.eigh(H)
        def trace(V, lam):
            c = V[0,:]
            ph = np.exp(-1j*lam[None,:]*t[:,None])
            psi = (c[None,:]*ph) @ V.T
            return (np.abs(psi)**2 * states[None,:,f-1]).sum(-1)
        tr = trace(Vref, lam_ref)
        e0 = np.abs(trace(V0, lam)-tr).max(); e1 = np.abs(trace(V1, lam)-tr).max()
        res.append((gap/nrm, e0, e1, ngs, ok, growth[0], growth[1], np.abs(b).min()/nrm))
    return np.array(res, dtype=float)

if __name__=="__main__":
    rng = np.random.default_rng(0)
    which = sys.argv[1]
    if which == "trimer":
        const = {"max_N": 10, "sites": 3, "max_t": 25, "omegas": [-3,3,3], "chis":[0,0,0], "coupling":1, "timesteps":30}
        X = rng.uniform(-10,10,(1500,3))
    else:
        const = {"max_N": 100, "sites": 2, "max_t": 25, "omegas": [-3,3], "chis":[0,0], "coupling":1, "timesteps":30}
        X = rng.uniform(-10,10,(600,2))
    r = study(const, X, 2.0**-14)
    print(which, "n", len(r), "finite", r[:,4].mean())
    print("relgap quantiles", np.quantile(r[:,0],[0,.001,.01,.1,.5]))
    print("err no-GS: max %.3e p99 %.3e med %.3e" % (np.nanmax(r[:,1]), np.nanquantile(r[:,1],.99), np.nanmedian(r[:,1])))
    print("err GS   : max %.3e p99 %.3e med %.3e" % (np.nanmax(r[:,2]), np.nanquantile(r[:,2],.99), np.nanmedian(r[:,2])))
    print("points with GS:", (r[:,3]>0).mean(), "mean pairs", r[:,3].mean())
    print("growth max p %.3e q %.3e ; min |b|/nrm %.3e" % (np.nanmax(r[:,5]), np.nanmax(r[:,6]), r[:,7].min()))
    i = np.argsort(-np.nan_to_num(r[:,2], nan=1e9))[:8]
    for j in i: print(X[j], r[j])
