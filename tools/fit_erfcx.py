"""Provenance of the rational approximation used by the narrow Gaussian engine (od_narrow.cu):
R(u) = P5(u)/Q6(u) ~ erfcx(u/sqrt 2) on [0, 8.45], fitted by Loeb iteration in 60-digit arithmetic with the
weight 0.5 exp(-u^2/2), i.e. minimising the ABSOLUTE error of the normal CDF it is used for (5.4e-15).
Run: python tools/fit_erfcx.py"""
import mpmath as mp, numpy as np, sys
mp.mp.dps = 60
UMAX = mp.mpf('8.45')
def f(u): return mp.erfc(u/mp.sqrt(2))*mp.exp(u*u/2)      # erfcx(u/sqrt2)
def w(u): return mp.mpf('0.5')*mp.exp(-u*u/2)
N=240
nodes=[UMAX*(1-mp.cos(mp.pi*(i+mp.mpf('0.5'))/N))/2 for i in range(N)]
fv=[f(u) for u in nodes]; wv=[w(u) for u in nodes]
def fit(m,n,iters=12):
    den=[mp.mpf(1)]*N
    for it in range(iters):
        A=mp.matrix(N,m+1+n); b=mp.matrix(N,1)
        for i,u in enumerate(nodes):
            W=wv[i]/den[i]
            for k in range(m+1): A[i,k]=W*u**k
            for k in range(n): A[i,m+1+k]=-W*fv[i]*u**k
            b[i]=W*fv[i]*u**n
        c=mp.lu_solve(A.T*A, A.T*b) if False else mp.qr_solve(A,b)[0]
        p=[c[k] for k in range(m+1)]; q=[c[m+1+k] for k in range(n)]+[mp.mpf(1)]
        den=[abs(mp.polyval(q[::-1],u)) for u in nodes]
    return p,q
def check(p,q,npts=2000):
    # evaluate in DOUBLE with Horner, compare to mp reference, weighted
    pd=[float(x) for x in p]; qd=[float(x) for x in q]
    worst=0; worst_mp=0
    for i in range(npts+1):
        u=float(UMAX)*i/npts
        P=0.0
        for c in pd[::-1]: P=P*u+c
        Q=0.0
        for c in qd[::-1]: Q=Q*u+c
        ref=f(mp.mpf(u)); ww=w(mp.mpf(u))
        worst=max(worst, float(ww*abs(mp.mpf(P)/mp.mpf(Q)-ref)))
        worst_mp=max(worst_mp, float(ww*abs(mp.polyval(p[::-1],mp.mpf(u))/mp.polyval(q[::-1],mp.mpf(u))-ref)))
    return worst, worst_mp
for (m,n) in []:
    p,q=fit(m,n)
    e,em=check(p,q)
    print(m,n,"double-eval weighted err %.2e  exact-coef err %.2e"%(e,em), "signs p", ''.join('+' if x>0 else '-' for x in p), "q", ''.join('+' if x>0 else '-' for x in q))
    sys.stdout.flush()
p,q=fit(5,6,iters=20)
print("P", [mp.nstr(x,20) for x in p])
print("Q", [mp.nstr(x,20) for x in q])
print(check(p,q,4000))
