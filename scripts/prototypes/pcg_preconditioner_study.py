"""CPU study (scipy) of preconditioners for the Poisson system the oracle assembles on a periodic Kuhn box:
Jacobi (what the reference and the device PCG use) against Chebyshev polynomial preconditioning.

    python scripts/prototypes/pcg_preconditioner_study.py 16
"""
import sys, time, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as sla
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle
from vlasovtucker_b200 import synthetic
h = int(sys.argv[1]) if len(sys.argv)>1 else 16
nodes,tets,tris,ents = synthetic.kuhn_box(h,h,h,(1.0,1.0,1.0))
t=time.time()
m = oracle.Mesh.from_arrays(nodes,tets,tris,ents,[(1,2),(3,4),(5,6)])
po = oracle.Poisson(m); po.initialize()
ip, ix, va = po.csr()[:3]
n = len(ip)-1
A = sp.csr_matrix((va,ix,ip),shape=(n,n))
print("n",n,"nnz",A.nnz,"setup %.1fs"%(time.time()-t))
print("symmetric?", abs(A-A.T).max())
U = sp.triu(A).tocsr()
S = (U + sp.triu(A,1).T).tocsr()
rng = np.random.default_rng(0)
x = np.sin(2*np.pi*m.tetCentroid[:,0]) + 0.3*rng.standard_normal(n)
b = S @ x
dinv = 1.0/S.diagonal()
def run(M, name):
    it=[0]
    def cb(_): it[0]+=1
    t=time.time()
    sol, info = sla.cg(S, b, rtol=1e-12, atol=0, maxiter=20000, M=M, callback=cb)
    print("%-28s iters %5d  relres %.1e  (%.2fs)"%(name, it[0], np.linalg.norm(S@sol-b)/np.linalg.norm(b), time.time()-t), flush=True)
    return it[0]
run(sla.LinearOperator((n,n), matvec=lambda r: dinv*r), "Jacobi")
# Chebyshev polynomial preconditioner on D^-1 S, spectrum estimate by power iteration
def lam_max():
    v = rng.standard_normal(n)
    for _ in range(50):
        v = dinv*(S@v); l = np.linalg.norm(v); v/=l
    return l
lmax = 1.05*lam_max()
for deg in (2,4,8):
    for ratio in (10.0, 30.0, 100.0):
        lmin = lmax/ratio
        th, de = (lmax+lmin)/2, (lmax-lmin)/2
        def cheb(r, deg=deg, th=th, de=de):
            # standard Chebyshev iteration for S z = r with Jacobi scaling, zero initial guess
            z = np.zeros_like(r); res = r.copy()
            sigma = th/de; rho = 1.0/sigma
            dvec = dinv*res/th
            for k in range(deg):
                z = z + dvec
                res = res - S@dvec
                rho_new = 1.0/(2*sigma - rho)
                dvec = rho_new*rho*dvec + 2*rho_new/de*(dinv*res)
                rho = rho_new
            return z
        it = run(sla.LinearOperator((n,n), matvec=cheb), "Chebyshev deg %d ratio %g"%(deg,ratio))
        print("      SpMV total ~", it*(deg+1))
