"""CPU two-grid experiments behind DESIGN.md (hourglass hierarchy): PCG iterations of a two-grid cycle on the oracle's
plate matrices with and without a checkerboard-modulated coarse space.  TEST INFRASTRUCTURE / analysis only (numpy, scipy).

    python oracle/experiments/hourglass_two_grid.py
"""
import sys, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as sla
import os
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), 'tests'))
import femol_oracle as o

def prolong_1d(n):
    nc=(n+1)//2
    P=np.zeros((n+1,nc+1))
    for i in range(n+1):
        if i%2==0: P[i,i//2]=1
        elif i==n: P[i,(i+1)//2]=1
        else: P[i,i//2]=.5; P[i,i//2+1]=.5
    return P

def run(n, t=0.5, levels=2):
    case=dict(mesh=('Q4', 1., 1., n, n), physics='displacement', model='plate', materials=[('iso', 10920, 0.3, 1)], tensors=[t],
              fixed=[(([0], [[0, 1.]]), None)], forces=[([0,0,-1.,0,0,0], ([[0.001,0.999]], [[0.001,0.999]]))])
    arr=o.case_arrays(case)
    K=o.global_matrices(arr,'K').tocsr()
    F=arr['F']
    ntot=K.shape[0]
    fixed=np.zeros(ntot,bool); fixed[arr['fixed_idx']]=True
    Px=prolong_1d(n); Py=prolong_1d(n)
    P2=sp.kron(sp.csr_matrix(Py), sp.csr_matrix(Px))          # nodes x-fastest: index j*npx+i -> kron(Py,Px)
    P=sp.kron(P2, sp.identity(6)).tocsr()
    # zero rows of fixed fine dofs; coarse fixed dofs: those whose injection fine dof is fixed
    Dm=sp.diags((~fixed).astype(float))
    P=Dm@P
    nc=(n+1)//2; npc=nc+1
    cf=np.zeros(P.shape[1],bool)
    for J in range(npc):
        for I in range(npc):
            fi=min(2*I,n); fj=min(2*J,n)
            for d in range(6):
                if fixed[(fj*(n+1)+fi)*6+d]: cf[(J*npc+I)*6+d]=True
    keepc=np.where(~cf)[0]
    P=P[:,keepc]
    par=np.add.outer(np.arange(n+1),np.arange(n+1)).reshape(-1)%2
    S=sp.diags(np.repeat(1.0-2.0*par,6))
    # block-Jacobi
    Kd=K.toarray()
    Dinv=np.zeros_like(Kd)
    for i in range((n+1)**2):
        sl=slice(6*i,6*i+6); Dinv[sl,sl]=np.linalg.inv(Kd[sl,sl])
    Dinv=sp.csr_matrix(Dinv)
    Kc=(P.T@K@P).toarray(); Kci=np.linalg.inv(Kc)
    Ph=S@P
    Kh=(Ph.T@K@Ph).toarray(); Khi=np.linalg.inv(Kh)
    om=0.7
    def twogrid(r, hour):
        # symmetric: pre-smooth, coarse correction(s), post-smooth
        e=om*(Dinv@r)
        res=r-K@e
        e=e+P@(Kci@(P.T@res))
        if hour:
            res=r-K@e
            e=e+Ph@(Khi@(Ph.T@res))
            # symmetrise: apply first coarse correction again
            res=r-K@e
            e=e+P@(Kci@(P.T@res))
        res=r-K@e
        e=e+om*(Dinv@res)
        return e
    def additive(r):
        return twogrid(r, False) + Ph@(Khi@(Ph.T@r))
    def additive_smoothed(r):
        # hourglass space with its own Jacobi-smoothed two-grid cycle on S K S  (what a second V-cycle would do)
        rs=S@r
        Ks=S@K@S
        e=om*(Dinv@rs); res=rs-Ks@e; e=e+P@(Khi@(P.T@res)); res=rs-Ks@e; e=e+om*(Dinv@res)
        return twogrid(r, False) + S@e
    out=[]
    for name,fn in (('additive',additive),('additive, smoothed second cycle',additive_smoothed)):
        its=[0]
        def cb(x): its[0]+=1
        M=sla.LinearOperator(K.shape, matvec=fn)
        x,info=sla.cg(K,F,rtol=1e-8,M=M,maxiter=2000,callback=cb)
        print('   n=%d %s: %d'%(n,name,its[0]))
    for hour in (False, True):
        its=[0]
        def cb(x): its[0]+=1
        M=sla.LinearOperator(K.shape, matvec=lambda r: twogrid(r,hour))
        x,info=sla.cg(K,F,rtol=1e-8,M=M,maxiter=2000,callback=cb)
        out.append(its[0])
    print('n=%d t=%g: two-grid PCG iterations: standard %d, with hourglass-modulated coarse space %d'%(n,t,out[0],out[1]))

for n in (16,32):
    run(n)
