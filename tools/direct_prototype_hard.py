import numpy as np, sys
sys.path.insert(0, __import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), '..')); sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.abspath(__file__)))
from oracle import qtet_oracle as O
from direct_prototype_large import direct_vecs_1, gs_fix
from scipy.linalg import eigh_tridiagonal
np.seterr(all='ignore')
N=20
states = O.fock_basis(N,2); off = O.offdiag_matrix(states,1.0)
t = np.linspace(0,25,30)
def trace(V, lam, site=1):
    c = V[0,:]; ph = np.exp(-1j*lam[None,:]*t[:,None]); psi = (c[None,:]*ph) @ V.T
    return (np.abs(psi)**2 * states[None,:,site]).sum(-1)
import mpmath as mp
def mp_trace(a,b,site=1):
    mp.mp.dps=60
    d=len(a); H = mp.zeros(d,d)
    for k in range(d): H[k,k]=mp.mpf(float(a[k]))
    for k in range(d-1): H[k,k+1]=H[k+1,k]=mp.mpf(float(b[k]))
    E,Q = mp.eigsy(H)
    out=[]
    for tt in t:
        psi=[sum(Q[j,i]*Q[0,i]*mp.expj(-E[i]*mp.mpf(float(tt))) for i in range(d)) for j in range(d)]
        out.append(float(sum(states[j,site]*abs(psi[j])**2 for j in range(d))))
    return np.array(out)
rows=[]
for om in ([3,3],[-3,3]):
  for chi in [(0.05,0.05),(0.2,0.2),(0.5,0.5),(1,1),(2,2),(5,5),(10,10),(0.3,-0.3),(0.6,0.0),(3,3.0000001),(1.0,1.0+1e-9),(7,-7),(0,0)]:
    a = O.diag_energies(states, om, np.array(chi)); a=a-a.mean(); b=np.array([off[k+1,k] for k in range(N)])
    lam = eigh_tridiagonal(a,b,eigvals_only=True)
    nrm = np.abs(lam).max()
    z,_ = direct_vecs_1(a,b,lam)
    z0=z.copy()
    ngs = gs_fix(z, lam, 2.0**-12*nrm)
    Hd = np.diag(a)+np.diag(b,1)+np.diag(b,-1)
    lr,Vr = np.linalg.eigh(Hd)
    tm = mp_trace(a,b)
    print(om, chi, "relgap %.2e ngs %d | err direct %.2e  direct+GS %.2e  numpy-eigh %.2e | orth0 %.1e orthGS %.1e" % (np.diff(lam).min()/nrm, ngs,
        np.abs(trace(z0,lam)-tm).max(), np.abs(trace(z,lam)-tm).max(), np.abs(trace(Vr,lr)-tm).max(),
        np.abs(z0.T@z0-np.eye(21)).max(), np.abs(z.T@z-np.eye(21)).max()))
