import numpy as np,sys
raw=open(sys.argv[1],'rb').read()
NT,TPC,T,RB=np.frombuffer(raw[:16],np.int32)
tr=np.frombuffer(raw[16:],np.uint64).reshape(NT,TPC,8).astype(np.float64)
t0=tr[tr>1e12].min(); tr=np.where(tr>1e12,(tr-t0)/1e3,np.nan)
d=tr[:,0,:]; h=tr[:,T+1,:]
j=np.arange(30,240)
pc=lambda x: np.round(np.nanpercentile(x,[5,25,50,75,95]),2)
base=d[j,2]
print("period", pc(np.diff(d[30:240,7])))
for k,n in enumerate(["tail: follow done","Ppre in","T tile in","GEMM done","Dpre next tried","potrf tail done","inverse done","potrf head done"]):
    print(" %-16s"%n, pc(h[j,k]-base))
print(" barrier passed  ", pc(d[j,3]-base), " X", pc(d[j,5]-d[j,4]), " Dpre wait", pc(d[j,6]-d[j,5]), " update||stores", pc(d[j,7]-d[j,6]))
