"""Rounding of the trajectory formulations against 40-digit arithmetic (mpmath): the literal updateA recursion (Core.hs:239-240)
with and without a fused multiply-add, the closed form of rc_traj_kernel (u0 x cumulative product), and a cancellation-free
variant.  Prints the largest relative error of a(s) over the steps.    python tools/traj_rounding.py"""
import numpy as np, mpmath as mp, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rarecoal_b200 as R
T = np.array(R.defaultTimes()); dts = np.diff(np.concatenate([[0.0],T]))
mp.mp.dps = 40
for a0,N,S in ((100.0,1.0,440),(100.0,1.0,3043),(1000.0,1.0,440),(10000.0,0.5,440),(100.0,0.05,440)):
    EA = np.exp(-(0.5*dts/N))
    # literal (no fma)
    a=a0; lit=[]
    for s in range(S):
        a = 1.0/(1.0+((1.0/a)-1.0)*EA[s]); lit.append(a)
    # literal with fma emulate: fma(u,EA,1) in extended precision via mp
    a=a0; litf=[]
    for s in range(S):
        u=(1.0/a)-1.0
        d=float(mp.mpf(u)*mp.mpf(EA[s])+1)   # single rounding
        a=1.0/d; litf.append(a)
    # closed form u0*cumprod
    u0=(1.0/a0)-1.0; C=np.cumprod(EA[:S]); cf=1.0/(1.0+u0*C)
    # closed form with fma: fma(u0,C,1)
    cff=np.array([1.0/float(mp.mpf(u0)*mp.mpf(c)+1) for c in C])
    # positive form: (1-C)+C/a0 with 1-C from -expm1(-cumsum)
    cs=np.cumsum(0.5*dts[:S]/N); D=-np.expm1(-cs); Cx=np.exp(-cs); pf=1.0/(D+Cx/a0)
    # exact (given the double EA values as exact numbers, i.e. what the literal recursion would give without rounding)
    ex=[]; u=mp.mpf(1)/mp.mpf(a0)-1
    for s in range(S):
        u=u*mp.mpf(EA[s]); ex.append(1/(1+u))
    # exact with exact exp
    ex2=[]; acc=mp.mpf(0)
    for s in range(S):
        acc+=mp.mpf(0.5)*mp.mpf(dts[s])/mp.mpf(N); ex2.append(1/(1+(mp.mpf(1)/mp.mpf(a0)-1)*mp.e**(-acc)))
    def err(x,ref): return max(abs(float((mp.mpf(float(v))-r)/r)) for v,r in zip(x,ref))
    print(f"a0={a0} N={N} S={S}: lit {err(lit,ex):.2e} litfma {err(litf,ex):.2e} closed {err(cf,ex):.2e} closedfma {err(cff,ex):.2e} positive(vs exact-exp) {err(pf,ex2):.2e}  lit vs exact-exp {err(lit,ex2):.2e}")
    print("   lit vs litfma", max(abs(x-y)/y for x,y in zip(lit,litf)), " closedfma vs lit", max(abs(x-y)/y for x,y in zip(cff,lit)), " positive vs lit", max(abs(x-y)/y for x,y in zip(pf,lit)))
