"""Independent numpy restatement of the 5-point essential-matrix problem (orthonormal null space, dictionary
polynomials, numpy.roots) used to pin oracle/gcr_oracle_e.cpp: same equations, different arithmetic."""
import numpy as np
from numpy.polynomial import polynomial as Pn
from scipy.linalg import null_space


def pmul(a,b):
    c={}
    for ea,va in a.items():
        for eb,vb in b.items():
            e=(ea[0]+eb[0],ea[1]+eb[1],ea[2]+eb[2]); c[e]=c.get(e,0)+va*vb
    return c
def padd(a,b,s=1):
    c=dict(a)
    for e,v in b.items(): c[e]=c.get(e,0)+s*v
    return c
M3=[(3,0,0),(0,3,0),(2,1,0),(1,2,0),(2,0,1),(2,0,0),(0,2,1),(0,2,0),(1,1,1),(1,1,0),(1,0,2),(1,0,1),(1,0,0),(0,1,2),(0,1,1),(0,1,0),(0,0,3),(0,0,2),(0,0,1),(0,0,0)]
mono1=[(1,0,0),(0,1,0),(0,0,1),(0,0,0)]
def five_point(P):
    A=np.array([[p[0]*p[2],p[0]*p[3],p[0],p[1]*p[2],p[1]*p[3],p[1],p[2],p[3],1] for p in P])
    N=null_space(A).T
    E=[[{mono1[b]:N[b][3*j+i] for b in range(4)} for j in range(3)] for i in range(3)]
    EET=[[{} for j in range(3)] for i in range(3)]
    for i in range(3):
        for j in range(3):
            for k in range(3): EET[i][j]=padd(EET[i][j],pmul(E[i][k],E[j][k]))
    tr={}
    for i in range(3): tr=padd(tr,EET[i][i])
    for i in range(3): EET[i][i]=padd(EET[i][i],tr,-0.5)
    rows=[]
    for i in range(3):
        for j in range(3):
            r={}
            for k in range(3): r=padd(r,pmul(EET[i][k],E[k][j]))
            rows.append(r)
    d=padd(pmul(pmul(E[0][1],E[1][2]),E[2][0]),pmul(pmul(E[0][2],E[1][1]),E[2][0]),-1)
    d=padd(d,pmul(padd(pmul(E[0][2],E[1][0]),pmul(E[0][0],E[1][2]),-1),E[2][1]))
    d=padd(d,pmul(padd(pmul(E[0][0],E[1][1]),pmul(E[0][1],E[1][0]),-1),E[2][2]))
    rows.append(d)
    C=np.array([[r.get(m,0) for m in M3] for r in rows])
    X=np.linalg.solve(C[:,:10],C[:,10:])
    px=np.zeros((3,4));py=np.zeros((3,4));pc=np.zeros((3,5))
    for i in range(3):
        ra=X[4+2*i]; rb=X[5+2*i]
        px[i]=[ra[2],ra[1]-rb[2],ra[0]-rb[1],-rb[0]]
        py[i]=[ra[5],ra[4]-rb[5],ra[3]-rb[4],-rb[3]]
        pc[i]=[ra[9],ra[8]-rb[9],ra[7]-rb[8],ra[6]-rb[7],-rb[6]]
    c=np.zeros(11)
    for p,sg in zip([(0,1,2),(1,2,0),(2,0,1),(0,2,1),(1,0,2),(2,1,0)],[1,1,1,-1,-1,-1]):
        t=Pn.polymul(Pn.polymul(px[p[0]],py[p[1]]),pc[p[2]]); c[:len(t)]+=sg*t
    r=np.roots(c[::-1]); rr=np.sort(r[np.abs(r.imag)<1e-9*np.maximum(1,np.abs(r.real))].real)
    out=[]
    for z in rr:
        B=np.array([[Pn.polyval(z,px[i]),Pn.polyval(z,py[i])] for i in range(3)]); b=np.array([Pn.polyval(z,pc[i]) for i in range(3)])
        xz=np.linalg.solve(B[:2],b[:2]); x,y=-xz
        e=N[0]*x+N[1]*y+N[2]*z+N[3]; Em_=e.reshape(3,3).T
        out.append(Em_/np.linalg.norm(Em_))
    return out, c


# GrevLex column order of solver_essential_matrix_five_point_stewenius.h:264-309
S3=[(3,0,0),(2,1,0),(1,2,0),(0,3,0),(2,0,1),(1,1,1),(0,2,1),(1,0,2),(0,1,2),(0,0,3),(2,0,0),(1,1,0),(0,2,0),(1,0,1),(0,1,1),(0,0,2),(1,0,0),(0,1,0),(0,0,1),(0,0,0)]
def stewenius(P, w=None):
    """Independent numpy statement of EssentialMatrixFivePointSteweniusSolver::estimateModel for n > 5 points followed
    by EssentialMatrixEstimator::estimateModelNonminimal's normalisation: pivoted QR of A^T A (LAPACK dgeqp3 via scipy),
    dictionary polynomials, numpy.linalg.solve / eig.  Returns the normalised models (E33 >= 0), any order."""
    from scipy.linalg import qr
    P=np.asarray(P,dtype=float); w=np.ones(len(P)) if w is None else np.asarray(w,dtype=float)
    A=np.array([[p[0]*p[2],p[0]*p[3],p[0],p[1]*p[2],p[1]*p[3],p[1],p[2],p[3],1] for p in P])*w[:,None]
    Q,_,_=qr(A.T@A,pivoting=True)
    N=Q[:,5:]                      # 9 x 4
    E=[[{mono1[b]:N[i+3*j][b] for b in range(4)} for j in range(3)] for i in range(3)]
    EET=[[{} for j in range(3)] for i in range(3)]
    for i in range(3):
        for j in range(3):
            for k in range(3): EET[i][j]=padd(EET[i][j],pmul(E[i][k],E[j][k]),2)
    tr={}
    for i in range(3): tr=padd(tr,EET[i][i])
    rows=[]
    for i in range(3):
        for j in range(3):
            r={}
            for k in range(3): r=padd(r,pmul(EET[i][k],E[k][j]))
            rows.append(padd(r,pmul(tr,E[i][j]),-0.5))
    d=padd(pmul(pmul(E[0][1],E[1][2]),E[2][0]),pmul(pmul(E[0][2],E[1][1]),E[2][0]),-1)
    d=padd(d,pmul(padd(pmul(E[0][2],E[1][0]),pmul(E[0][0],E[1][2]),-1),E[2][1]))
    d=padd(d,pmul(padd(pmul(E[0][0],E[1][1]),pmul(E[0][1],E[1][0]),-1),E[2][2]))
    rows.append(d)
    C=np.array([[r.get(m,0) for m in S3] for r in rows])
    X=np.linalg.solve(C[:,:10],C[:,10:])
    Am=np.zeros((10,10)); Am[:3]=X[:3]; Am[3]=X[4]; Am[4]=X[5]; Am[5]=X[7]
    Am[6,0]=Am[7,1]=Am[8,3]=Am[9,6]=-1
    lam,V=np.linalg.eig(Am)
    out=[]
    for i in range(10):
        if lam[i].imag!=0: continue
        e=N@V[6:,i].real; Em_=e.reshape(3,3).T
        Em_=Em_/np.linalg.norm(Em_)
        out.append(-Em_ if Em_[2,2]<0 else Em_)
    return out
