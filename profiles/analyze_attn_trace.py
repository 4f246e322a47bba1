import numpy as np, sys
for name in sys.argv[1:]:
    a=np.fromfile(name,dtype=np.uint64).reshape(3,4096)
    ev=[]
    for reg in range(3):
        r=a[reg]; r=r[r!=0]
        for x in r: ev.append((int(x>>np.uint64(8)), reg, int(x&np.uint64(0xff))))
    ev.sort()
    print('====',name,'events',len(ev))
    for tile in (0,1):
        e=[(t,i) for t,r,i in ev if r==tile]
        its=[]; cur={}
        for t,i in e:
            cur[i]=t
            if i==5: its.append(cur); cur={}
        its=its[100:300]
        def avg(f): return np.mean([f(x) for x in its])
        period=np.mean([its[k+1][1]-its[k][1] for k in range(len(its)-1)])
        print(f'tile{tile}: period {period:.0f} | ld {avg(lambda x:x[2]-x[1]):.0f} max+resc {avg(lambda x:x[3]-x[2]):.0f} exp+st-issue {avg(lambda x:x[4]-x[3]):.0f} waitst+arrive {avg(lambda x:x[5]-x[4]):.0f} | busy {avg(lambda x:x[5]-x[1]):.0f} | wait-for-S {np.mean([its[k+1][1]-its[k][5] for k in range(len(its)-1)]):.0f}')
    e=[(t,i) for t,r,i in ev if r==2]
    its=[]; cur={}
    for t,i in e:
        cur[i]=t
        if i==15: its.append(cur); cur={}
    its=its[100:300]
    def avg(f): return np.mean([f(x) for x in its])
    print('issuer: P0obs->PV0 issued %.0f, ->QK0 issued %.0f, wait P1 %.0f, PV1 issue %.0f, QK1 issue %.0f, wait P0(next) %.0f'%(
        avg(lambda x:x[12]-x[10]), avg(lambda x:x[14]-x[12]), avg(lambda x:x[11]-x[14]), avg(lambda x:x[13]-x[11]), avg(lambda x:x[15]-x[13]),
        np.mean([its[k+1][10]-its[k][15] for k in range(len(its)-1)])))
    s0=[t for t,r,i in ev if r==0 and i==5]; i10=[t for t,r,i in ev if r==2 and i==10]
    s0f=[t for t,r,i in ev if r==0 and i==1]; i14=[t for t,r,i in ev if r==2 and i==14]
    n=min(len(s0),len(i10)); print('tile0 arrive->issuer obs: %.0f'%np.mean([i10[k]-s0[k] for k in range(100,min(300,n))]))
    n=min(len(s0f)-1,len(i14)); print('QK0 issued+commit -> softmax0 sees S(next): %.0f'%np.mean([s0f[k+1]-i14[k] for k in range(100,min(300,n))]))
    s1=[t for t,r,i in ev if r==1 and i==5]; i11=[t for t,r,i in ev if r==2 and i==11]
    n=min(len(s1),len(i11)); print('tile1 arrive->issuer obs: %.0f'%np.mean([i11[k]-s1[k] for k in range(100,min(300,n))]))
    s1f=[t for t,r,i in ev if r==1 and i==1]; i15=[t for t,r,i in ev if r==2 and i==15]
    n=min(len(s1f)-1,len(i15)); print('QK1 issued+commit -> softmax1 sees S(next): %.0f'%np.mean([s1f[k+1]-i15[k] for k in range(100,min(300,n))]))
    n=min(len(s0),len(s1)); print('tile1 P-ready minus tile0 P-ready: %.0f'%np.mean([s1[k]-s0[k] for k in range(100,min(300,n))]))
