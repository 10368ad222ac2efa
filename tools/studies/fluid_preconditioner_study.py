import sys, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))); sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.abspath(__file__)))
from oracle.fluid_oracle import OracleFluidSolver
from fluid_minres_prototype import pminres
def build(N,rho=None):
    s=OracleFluidSolver(N,__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))+'/designs/diffuser.json'); pr=s.problem; pr.set_penalization(0.1)
    m=s.mesh; nu,n1=m.nu,m.n1
    if rho is None: rho=s.rho
    full=(pr.A0+pr._brinkman(rho)).tocsr(); A=full[:nu,:nu]; D=full[nu:,:nu]
    interior=np.ones(nu,bool); interior[pr.bc_dofs]=False; I=np.flatnonzero(interior)
    g=np.zeros(nu); g[pr.bc_dofs]=pr.bc_vals
    AII=A[I][:,I].tocsc(); B=(-D[:,I]).tocsr()
    fu=-(A[I]@g); fp=D@g; fp-=fp.mean()
    Mr=pr._brinkman(rho).tocsr()[:nu,:nu][I][:,I]
    return s,pr,AII,B,np.concatenate([fu,fp]),Mr
def study(N,rho=None,label=''):
    s,pr,AII,B,b,Mr=build(N,rho)
    nI=AII.shape[0]; n1=B.shape[0]
    K=sp.bmat([[AII,B.T],[B,None]],format='csr')
    dA=AII.diagonal()
    dS_as=np.asarray((B.multiply(B))@(1/dA)).ravel()
    Alu=spla.splu(AII)
    Mp=pr.M1.tocsc(); Mplu=spla.splu(Mp); dMp=Mp.diagonal()
    # Cahouet-Chabard-like: S^-1 ~ Mp^-1 (viscous) + (B diag(Mr)^-1 B^T)^-1 (Darcy), regularised
    dMr=Mr.diagonal()
    Lp=(B@sp.diags(1/dMr)@B.T).tocsc()+1e-10*sp.identity(n1,format='csc')*Lp_scale if False else None
    L=(B@sp.diags(1/dMr)@B.T).tocsc(); L=L+1e-8*L.diagonal().max()*sp.identity(n1,format='csc'); Llu=spla.splu(L)
    S=(B@sp.csc_matrix(Alu.solve(B.T.toarray()))) if n1<2000 else None
    def run(name,prec):
        x,its,rr=pminres(lambda v:K@v,prec,b,1e-10,20000)
        print(f'  {label} N={N} {name:34s} its={its}')
    run('diag(A) | diag S (assembled)', lambda v: np.concatenate([v[:nI]/dA, v[nI:]/dS_as]))
    run('exact A | diag S (assembled)', lambda v: np.concatenate([Alu.solve(v[:nI]), v[nI:]/dS_as]))
    run('exact A | diag(Mp)', lambda v: np.concatenate([Alu.solve(v[:nI]), v[nI:]/dMp]))
    run('exact A | Mp^-1 + L^-1 (C-C)', lambda v: np.concatenate([Alu.solve(v[:nI]), Mplu.solve(v[nI:])+Llu.solve(v[nI:])]))
    run('diag(A) | Mp^-1 + L^-1 (C-C)', lambda v: np.concatenate([v[:nI]/dA, Mplu.solve(v[nI:])+Llu.solve(v[nI:])]))
s=OracleFluidSolver(20,__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))+'/designs/diffuser.json'); r=s.solve(); rho20=r['rho']
study(20,None,'uniform'); study(20,rho20,'late   '); study(40,None,'uniform')
