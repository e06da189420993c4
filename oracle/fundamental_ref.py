"""oracle/fundamental_ref.py -- numpy restatement of cv::findFundamentalMat(pts1, pts2, FM_RANSAC, threshold, 0.99, status)
(TEST INFRASTRUCTURE ONLY), the outlier gate of Tracking::trackReferenceFrame (ic_gvins/ic_gvins/tracking/tracking.cc:546-555; only
`status` is consumed).  OpenCV is an un-vendored dependency of the reference; its published algorithm (calib3d: fundam.cpp /
ptsetreg.cpp) is restated here -- cv::RNG(-1) multiply-with-carry stream, 7-point minimal subsets with duplicate / collinearity
rejection, null space of the 7 x 9 system, cubic in lambda (cv::solveCubic), up to three models per subset, symmetric epipolar error
in double rounded to float, strict `goodCount > max(best, 6)` acceptance, RANSACUpdateNumIters -- and pinned against cv2 4.13.0:
tests/golden/fundamental_golden.npz holds inlier masks that this file reproduces bit for bit (tests/test_fundamental.py)."""
import math

import numpy as np
class RNG:
    def __init__(s, state=0xffffffffffffffff): s.state=state
    def next(s):
        s.state=((s.state & 0xffffffff)*4164903690 + (s.state>>32)) & 0xffffffffffffffff
        return s.state & 0xffffffff
    def uniform(s,a,b):
        return a + s.next() % (b-a) if a!=b else a
def collinear(pts, count):
    # haveCollinearPoints: check that the i-th selected point does not belong to a line connecting some previously selected points
    i=count-1
    for j in range(i):
        dx1=pts[j][0]-pts[i][0]; dy1=pts[j][1]-pts[i][1]
        for k in range(j):
            dx2=pts[k][0]-pts[i][0]; dy2=pts[k][1]-pts[i][1]
            if abs(dx2*dy1-dy2*dx1) <= 1.1920929e-07*(abs(dx1)+abs(dy1)+abs(dx2)+abs(dy2)): return True
    return False
def solve_cubic(c):
    a0,a1,a2,a3=c  # a0 x^3 + a1 x^2 + a2 x + a3
    if a0==0:
        if a1==0:
            if a2==0: return []
            return [-a3/a2]
        d=a2*a2-4*a1*a3
        if d>=0:
            d=math.sqrt(d); q1=(-a2+d)*0.5; q2=(a2+d)*-0.5
            if abs(q1)>abs(q2): return [q1/a1, a3/q1]
            else:
                return [q2/a1, (a3/q2 if q2!=0 else q2/a1)]
        return []
    a0=1./a0; a1*=a0; a2*=a0; a3*=a0
    Q=(a1*a1-3*a2)*(1./9); R=(2*a1*a1*a1-9*a1*a2+27*a3)*(1./54); Qcubed=Q*Q*Q; d=Qcubed-R*R
    if d>0:
        theta=math.acos(R/math.sqrt(Qcubed)); sqrtQ=math.sqrt(Q); t0=-2*sqrtQ; t1=theta*(1./3); t2=a1*(1./3)
        return [t0*math.cos(t1)-t2, t0*math.cos(t1+(2.*math.pi/3))-t2, t0*math.cos(t1+(4.*math.pi/3))-t2]
    elif d==0:
        if R>=0: 
            x0=-2*pow(R,1./3)-a1/3; x1=pow(R,1./3)-a1/3
        else:
            x0=2*pow(-R,1./3)-a1/3; x1=-pow(-R,1./3)-a1/3
        return [x0,x1]
    else:
        d=math.sqrt(-d); e=pow(d+abs(R),1./3)
        if R>0: e=-e
        return [(e+Q/e)-a1*(1./3)]
def run7(m1,m2):
    a=np.zeros((7,9))
    for i in range(7):
        x0,y0=map(float,m1[i]); x1,y1=map(float,m2[i])
        a[i]=[x1*x0,x1*y0,x1,y1*x0,y1*y0,y1,x0,y0,1]
    u,w,vt=np.linalg.svd(a,full_matrices=True)
    f1=vt[7].copy(); f2=vt[8].copy()
    f1=f1-f2
    t0=f2[4]*f2[8]-f2[5]*f2[7]; t1=f2[3]*f2[8]-f2[5]*f2[6]; t2=f2[3]*f2[7]-f2[4]*f2[6]
    c=[0]*4
    c[3]=f2[0]*t0-f2[1]*t1+f2[2]*t2
    c[2]=(f1[0]*t0-f1[1]*t1+f1[2]*t2-f1[3]*(f2[1]*f2[8]-f2[2]*f2[7])+f1[4]*(f2[0]*f2[8]-f2[2]*f2[6])-f1[5]*(f2[0]*f2[7]-f2[1]*f2[6])+f1[6]*(f2[1]*f2[5]-f2[2]*f2[4])-f1[7]*(f2[0]*f2[5]-f2[2]*f2[3])+f1[8]*(f2[0]*f2[4]-f2[1]*f2[3]))
    t0=f1[4]*f1[8]-f1[5]*f1[7]; t1=f1[3]*f1[8]-f1[5]*f1[6]; t2=f1[3]*f1[7]-f1[4]*f1[6]
    c[0]=f1[0]*t0-f1[1]*t1+f1[2]*t2
    c[1]=(f2[0]*t0-f2[1]*t1+f2[2]*t2-f2[3]*(f1[1]*f1[8]-f1[2]*f1[7])+f2[4]*(f1[0]*f1[8]-f1[2]*f1[6])-f2[5]*(f1[0]*f1[7]-f1[1]*f1[6])+f2[6]*(f1[1]*f1[5]-f1[2]*f1[4])-f2[7]*(f1[0]*f1[5]-f1[2]*f1[3])+f2[8]*(f1[0]*f1[4]-f1[1]*f1[3]))
    roots=solve_cubic(c)
    Fs=[]
    for r in roots:
        lam=r; mu=1.0; s=f1[8]*r+f2[8]
        F=np.zeros(9)
        if abs(s)>2.220446049250313e-16: mu=1./s; lam*=mu; F[8]=1.
        else: F[8]=0.
        F[:8]=f1[:8]*lam+f2[:8]*mu
        Fs.append(F)
    return Fs
def errors(m1,m2,F):
    x1=m1[:,0].astype(np.float64); y1=m1[:,1].astype(np.float64); x2=m2[:,0].astype(np.float64); y2=m2[:,1].astype(np.float64)
    a=F[0]*x1+F[1]*y1+F[2]; b=F[3]*x1+F[4]*y1+F[5]; c=F[6]*x1+F[7]*y1+F[8]
    s2=1./(a*a+b*b); d2=x2*a+y2*b+c
    a=F[0]*x2+F[3]*y2+F[6]; b=F[1]*x2+F[4]*y2+F[7]; c=F[2]*x2+F[5]*y2+F[8]
    s1=1./(a*a+b*b); d1=x1*a+y1*b+c
    return np.maximum(d1*d1*s1,d2*d2*s2).astype(np.float32)
def update_iters(p,ep,mp,maxIters):
    p=min(max(p,0.),1.); ep=min(max(ep,0.),1.)
    num=max(1.-p,2.2250738585072014e-308); denom=1.-pow(1.-ep,mp)
    if denom<2.2250738585072014e-308: return 0
    num=math.log(num); denom=math.log(denom)
    return maxIters if (denom>=0 or -num>=maxIters*(-denom)) else int(np.rint(num/denom))
def find_fm_ransac(m1,m2,thresh,conf,maxIters=1000):
    m1=np.asarray(m1,np.float32); m2=np.asarray(m2,np.float32); count=len(m1)
    rng=RNG(); niters=maxIters; best=None; maxGood=0; t=np.float32(thresh*thresh)
    it=0
    while it<niters:
        # getSubset
        found=False
        iters=0
        while iters<10000:
            idx=[]; i=0
            while i<7 and iters<10000:
                while True:
                    v=rng.uniform(0,count)
                    if v not in idx: break
                idx.append(v); i+=1
            # checkSubset after full subset
            ms1=m1[idx]; ms2=m2[idx]
            if i==7 and (collinear(ms1,7) or collinear(ms2,7)):
                iters+=1; continue
            found=True; break
        if not found:
            if it==0: return None
            break
        Fs=run7(ms1,ms2)
        for F in Fs:
            err=errors(m1,m2,F); mask=(err<=t); good=int(mask.sum())
            if good>max(maxGood,6):
                best=mask.copy(); maxGood=good
                niters=update_iters(conf,(count-good)/count,7,niters)
        it+=1
    return best


def triangulate_point(T0, T1, pc0, pc1):
    """Tracking::triangulatePoint (tracking.cc:796-808): T = 3 x 4 T_c_w; pw = V[:, -1] of the 4 x 4 design matrix, dehomogenised"""
    T0, T1 = np.asarray(T0, float).reshape(3, 4), np.asarray(T1, float).reshape(3, 4)
    D = np.stack([pc0[0] * T0[2] - T0[0], pc0[1] * T0[2] - T0[1], pc1[0] * T1[2] - T1[0], pc1[1] * T1[2] - T1[1]])
    v = np.linalg.svd(D)[2][-1]
    return v[:3] / v[3]
