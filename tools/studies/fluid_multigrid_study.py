"""Numerical prototype (numpy/scipy, oracle matrices) of the round-2 fluid preconditioner:
velocity block: geometric multigrid V-cycle on A = M_r + K (Galerkin coarse operators through the nested-P2
prolongation, Chebyshev-Jacobi smoothing); pressure block: M_p^-1 (diag) + V-cycle / exact on a P1
variable-coefficient Laplacian L1 = int (1/r) grad.grad (Darcy part of the Schur complement)."""
import sys, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))); sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.abspath(__file__)))
from oracle.fluid_oracle import OracleFluidSolver
from oracle.fem_oracle import StructuredMesh, evaluate_field
from fluid_minres_prototype import pminres

def p2_prolongation(mc, mf):
    """scalar P2 prolongation coarse->fine by evaluating coarse basis functions at fine nodes"""
    X, Y = np.meshgrid(mf.xl, mf.yl, indexing='xy')
    xs, ys = X.ravel(), Y.ravel()
    cols=[]
    # evaluate_field works on vector P2 (degree 2) fields of length nu: use component 0
    P = sp.lil_matrix((mf.n2, mc.n2))
    # exploit locality: a coarse basis function is supported on <= 6 coarse cells; brute force columns in blocks
    I = np.eye(mc.n2)
    for j in range(mc.n2):
        vals = np.zeros(mc.nu); vals[2*j] = 1.0
        out = evaluate_field(mc, vals, 2, xs, ys)
        col = np.asarray(out)[:,0] if np.ndim(out)==2 else np.asarray(out)[0::2]
        nz = np.flatnonzero(np.abs(col) > 1e-14)
        P[nz, j] = col[nz]
    return P.tocsr()

def cheb_jacobi(A, dinv, b, x, lmax, steps, ratio=30.0):
    lmin = lmax/ratio; theta=0.5*(lmax+lmin); delta=0.5*(lmax-lmin)
    sigma=theta/delta; rho_old=1.0/sigma
    r=b-A@x; d=dinv*r/theta; x=x+d
    for _ in range(steps-1):
        rho=1.0/(2*sigma-rho_old)
        r=b-A@x
        d=rho*rho_old*d+2*rho/delta*(dinv*r)
        x=x+d; rho_old=rho
    return x

class MG:
    def __init__(self, A, Ps, coarse_steps=3, fine_steps=2):
        self.A=[A]; self.P=Ps
        for P in Ps: self.A.append((P.T@self.A[-1]@P).tocsr())
        self.dinv=[1.0/a.diagonal() for a in self.A]
        self.lmax=[]
        for a,d in zip(self.A,self.dinv):
            v=np.random.default_rng(0).random(a.shape[0])
            for _ in range(30):
                v=d*(a@v); v/=np.linalg.norm(v)
            self.lmax.append(1.1*float(v@(d*(a@v))))
        self.lu=spla.splu(self.A[-1].tocsc())
        self.steps=[fine_steps]+[coarse_steps]*(len(self.A)-1)
    def vcycle(self,b,l=0):
        if l==len(self.A)-1: return self.lu.solve(b)
        A,d=self.A[l],self.dinv[l]
        x=cheb_jacobi(A,d,b,np.zeros_like(b),self.lmax[l],self.steps[l])
        r=b-A@x
        x=x+self.P[l]@self.vcycle(self.P[l].T@r,l+1)
        x=x+cheb_jacobi(A,d,b-A@x,np.zeros_like(b),self.lmax[l],self.steps[l])
        return x

def study(N, rho_fn, label):
    s=OracleFluidSolver(N,__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))+'/designs/diffuser.json'); pr=s.problem; pr.set_penalization(0.1)
    m=s.mesh; nu,n1,n2=m.nu,m.n1,m.n2
    rho=rho_fn(s)
    full=(pr.A0+pr._brinkman(rho)).tocsr(); A=full[:nu,:nu]; D=full[nu:,:nu]
    interior=np.ones(nu,bool); interior[pr.bc_dofs]=False
    g=np.zeros(nu); g[pr.bc_dofs]=pr.bc_vals
    Pi=sp.diags(interior.astype(float))
    A0=(Pi@A@Pi+sp.diags((~interior).astype(float))).tocsr(); B=-(D@Pi)
    K=sp.bmat([[A0,B.T],[B,None]],format='csr')
    fu=-(Pi@(A@g)); fp=D@g; fp-=fp.mean(); b=np.concatenate([fu,fp])
    # scalar velocity operator (x component) on the node lattice with Dirichlet identity
    As=A0[0::2][:,0::2].tocsr()
    # hierarchy of meshes
    meshes=[m]; 
    while meshes[-1].nx%2==0 and meshes[-1].ny%2==0 and meshes[-1].nx>2:
        mm=meshes[-1]; meshes.append(StructuredMesh(mm.W,mm.H,mm.nx//2,mm.ny//2))
    Ps=[]
    for mf,mc in zip(meshes[:-1],meshes[1:]):
        P=p2_prolongation(mc,mf)
        # Dirichlet: boundary nodes of both levels are fixed -> zero those rows/cols
        def bmask(mm):
            Xl,Yl=np.meshgrid(np.arange(mm.Lx),np.arange(mm.Ly),indexing='xy')
            return ((Xl==0)|(Xl==mm.Lx-1)|(Yl==0)|(Yl==mm.Ly-1)).ravel()
        bf,bc=bmask(mf),bmask(mc)
        P=sp.diags((~bf).astype(float))@P@sp.diags((~bc).astype(float))
        # keep identity coupling for boundary so Galerkin operator stays nonsingular
        P=P+sp.csr_matrix((np.ones(bc.sum()),(np.flatnonzero(bf)[np.searchsorted(np.flatnonzero(bf),np.flatnonzero(bf))][:0],[])),shape=P.shape) if False else P
        Ps.append(P.tocsr())
    # Galerkin with zeroed boundary columns gives singular coarse ops on boundary dofs: add identity there
    mg=MG.__new__(MG); mg.A=[As]; mg.P=Ps
    for P,mc in zip(Ps,meshes[1:]):
        Ac=(P.T@mg.A[-1]@P).tolil()
        Xl,Yl=np.meshgrid(np.arange(mc.Lx),np.arange(mc.Ly),indexing='xy')
        bc=np.flatnonzero(((Xl==0)|(Xl==mc.Lx-1)|(Yl==0)|(Yl==mc.Ly-1)).ravel())
        for i in bc: Ac[i,i]=1.0
        mg.A.append(Ac.tocsr())
    mg.dinv=[1.0/a.diagonal() for a in mg.A]; mg.lmax=[]
    for a,d in zip(mg.A,mg.dinv):
        v=np.random.default_rng(0).random(a.shape[0])
        for _ in range(40):
            v=d*(a@v); v/=np.linalg.norm(v)
        mg.lmax.append(1.1*float(v@(d*(a@v))))
    mg.lu=spla.splu(mg.A[-1].tocsc()); mg.steps=[2]+[3]*(len(mg.A)-1)
    def Aprec(v):
        out=np.empty_like(v); out[0::2]=mg.vcycle(v[0::2]); out[1::2]=mg.vcycle(v[1::2]); return out
    # pressure: Mp diag + L1 = int (1/r) grad.grad on P1 (vertex-averaged 1/r per triangle)
    K1rows,K1cols,K1vals=[],[],[]
    for t in ('A','B'):
        area,gl=m.geom[t]; conn=m.tri_v[t]
        w=(1.0/pr.r(rho[conn])).mean(axis=1)*area
        Ke=gl@gl.T
        K1rows.append(np.repeat(conn,3,axis=1).ravel()); K1cols.append(np.tile(conn,(1,3)).ravel())
        K1vals.append((w[:,None]*Ke.ravel()[None,:]).ravel())
    L1=sp.csr_matrix((np.concatenate(K1vals),(np.concatenate(K1rows),np.concatenate(K1cols))),shape=(n1,n1))
    L1=L1+1e-8*L1.diagonal().max()*sp.identity(n1)
    L1lu=spla.splu(L1.tocsc()); dMp=pr.M1.diagonal(); Mplu=spla.splu(pr.M1.tocsc())
    dA=A0.diagonal()
    def run(name,prec):
        t=time.time(); x,its,rr=pminres(lambda v:K@v,prec,b,1e-10,5000)
        print(f'  {label} N={N} {name:44s} its={its:5d}  ({time.time()-t:.1f}s)')
    run('V-cycle(A) | Mp^-1 + L1^-1 (P1 Darcy Laplacian)', lambda v: np.concatenate([Aprec(v[:nu]), Mplu.solve(v[nu:])+L1lu.solve(v[nu:])]))
    run('V-cycle(A) | diag(Mp)^-1 + L1^-1', lambda v: np.concatenate([Aprec(v[:nu]), v[nu:]/dMp+L1lu.solve(v[nu:])]))
    run('V-cycle(A) | diag(Mp)^-1', lambda v: np.concatenate([Aprec(v[:nu]), v[nu:]/dMp]))

sd=OracleFluidSolver(16,__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))+'/designs/diffuser.json'); rd=sd.solve(); rho16=rd['rho']
study(16, lambda s: s.rho, 'uniform')
study(16, lambda s: rho16, 'late   ')
study(32, lambda s: s.rho, 'uniform')
