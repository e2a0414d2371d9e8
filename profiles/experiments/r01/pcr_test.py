import numpy as np
def pcr(a, b, c, d):
    n = len(b); a=a.copy(); b=b.copy(); c=c.copy(); d=d.copy()
    s = 1
    while s < n:
        an, bn, cn, dn = a.copy(), b.copy(), c.copy(), d.copy()
        for i in range(n):
            al = -a[i]/b[i-s] if i-s >= 0 else 0.0
            ga = -c[i]/b[i+s] if i+s < n else 0.0
            an[i] = al*a[i-s] if i-s >= 0 else 0.0
            cn[i] = ga*c[i+s] if i+s < n else 0.0
            bn[i] = b[i] + (al*c[i-s] if i-s >= 0 else 0.0) + (ga*a[i+s] if i+s < n else 0.0)
            dn[i] = d[i] + (al*d[i-s] if i-s >= 0 else 0.0) + (ga*d[i+s] if i+s < n else 0.0)
        a, b, c, d = an, bn, cn, dn
        s *= 2
    return d/b
def thomas(a, b, c, d):
    n=len(b); b=b.copy(); d=d.copy()
    for i in range(1,n):
        w=a[i]/b[i-1]; b[i]-=w*c[i-1]; d[i]-=w*d[i-1]
    x=np.zeros(n); x[-1]=d[-1]/b[-1]
    for i in range(n-2,-1,-1): x[i]=(d[i]-c[i]*x[i+1])/b[i]
    return x
rng=np.random.default_rng(0)
for n, lo, hi in ((500, 1e-5, 20.), (1000, 1e-5, 10.), (250, 5e-6, 1.), (200, 1e-3, 0.5)):
    x = np.logspace(np.log10(lo), np.log10(hi), n); y = np.sin(3*np.log(x)) * x**-1.5 + rng.normal(size=n)*1e-3
    h = np.diff(x)
    # netlib system (Spline.cpp:211-257)
    b = np.zeros(n); a = np.zeros(n); c = np.zeros(n); r = np.zeros(n)
    b[1:-1] = 2*(h[:-1]+h[1:]); a[1:] = h; c[:-1] = h
    sl = np.diff(y)/h; r[1:-1] = sl[1:]-sl[:-1]
    b[0] = -h[0]; b[-1] = -h[-1]
    r[0] = (r[2]/(x[3]-x[1]) - r[1]/(x[2]-x[0]))*h[0]**2/(x[3]-x[0])
    r[-1] = -(r[-2]/(x[-1]-x[-3]) - r[-3]/(x[-2]-x[-4]))*h[-1]**2/(x[-1]-x[-4])
    xt = thomas(a,b,c,r); xp = pcr(a,b,c,r)
    print('netlib n=%d max rel diff %.2e (scale %.2e)' % (n, np.max(np.abs(xt-xp))/np.max(np.abs(xt)), np.max(np.abs(xt))), ' pointwise max rel %.2e'%np.max(np.abs(xt-xp)/np.maximum(np.abs(xt),1e-300)))
