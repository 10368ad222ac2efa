import sys, numpy as np, scipy.sparse as sp, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle.fluid_oracle import OracleFluidSolver

def pminres(apply, prec, b, tol, maxit):
    n=len(b); x=np.zeros(n)
    v_old=np.zeros(n); v=b.copy(); z=prec(v)
    gamma=np.sqrt(z@v); gamma_old=1.0
    eta=gamma; eta0=gamma
    s_old=s=0.0; c_old=c=1.0
    w_old=np.zeros(n); w=np.zeros(n)
    for j in range(1,maxit+1):
        z=z/gamma
        Az=apply(z)
        delta=Az@z
        v_new=Az-(delta/gamma)*v-(gamma/gamma_old)*v_old
        z_new=prec(v_new)
        gamma_new=np.sqrt(z_new@v_new)
        a0=c*delta-c_old*s*gamma
        a1=np.sqrt(a0*a0+gamma_new*gamma_new)
        a2=s*delta+c_old*c*gamma
        a3=s_old*gamma
        c_new=a0/a1; s_new=gamma_new/a1
        w_new=(z-a3*w_old-a2*w)/a1
        x=x+c_new*eta*w_new
        eta=-s_new*eta
        v_old,v=v,v_new; z=z_new; gamma_old,gamma=gamma,gamma_new
        c_old,c=c,c_new; s_old,s=s,s_new; w_old,w=w,w_new
        if abs(eta)<=tol*eta0: return x,j,abs(eta)/eta0
    return x,maxit,abs(eta)/eta0

def run(N,rho=None,tol=1e-10):
    s=OracleFluidSolver(N,__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))+'/designs/diffuser.json'); pr=s.problem; pr.set_penalization(0.1)
    m=s.mesh; nu,n1=m.nu,m.n1
    if rho is None: rho=s.rho
    Afull=(pr.A0+pr._brinkman(rho)).tocsr()
    A=Afull[:nu,:nu]; D=Afull[nu:,:nu]
    interior=np.ones(nu,bool); interior[pr.bc_dofs]=False
    g=np.zeros(nu); g[pr.bc_dofs]=pr.bc_vals
    P=sp.diags(interior.astype(float))
    A0=P@A@P + sp.diags((~interior).astype(float))   # identity on boundary
    B=-(D@P)
    K=sp.bmat([[A0,B.T],[B,None]],format='csr')
    fu=-(P@(A@g)); fp=D@g; fp-=fp.mean()
    b=np.concatenate([fu,fp])
    dA=A0.diagonal()
    # un-assembled Schur diagonal: sum over triangles of Dloc^2/dA
    dS=np.zeros(n1)
    for t in ('A','B'):
        pts,w,phi,dphi,gl=pr._table(t,6)
        for d in (0,1):
            Dl=np.einsum('q,qc,ql->cl',w,pts,dphi[:,:,d])  # (3,6)
            nodes=m.tri_n[t]; dof=2*nodes+d               # (nt,6)
            inv=np.where(interior[dof],1.0/dA[dof],0.0)     # (nt,6)
            contrib=(Dl**2)[None,:,:]*inv[:,None,:]        # (nt,3,6)
            for cidx in range(3):
                dS+=np.bincount(m.tri_v[t][:,cidx],weights=contrib[:,cidx,:].sum(1),minlength=n1)
    dM=np.concatenate([dA,dS])
    t0=time.time()
    x,its,rr=pminres(lambda v:K@v, lambda v:v/dM, b, tol, 100000)
    u=g+x[:nu]
    uo,po=pr.forward(rho)
    print(N,'its',its,'est',rr,'true relres',np.linalg.norm(b-K@x)/np.linalg.norm(b),'u diff',np.abs(u-uo).max()/np.abs(uo).max(), 'time',time.time()-t0)
    return s
if __name__=="__main__": s=run(20)
if __name__=="__main__": r=s.solve(); run(20,r["rho"])
if __name__=="__main__": run(40)
