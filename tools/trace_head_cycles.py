import numpy as np,sys
raw=open(sys.argv[1],'rb').read()
NT,TPC,T,RB=np.frombuffer(raw[:16],np.int32)
tr=np.frombuffer(raw[16:],np.uint64).reshape(NT,TPC,8)
h=tr[:,T+1,:]; j=np.arange(30,240)
print(sys.argv[2], "head cycles median", np.median(h[j,1].astype(np.float64)))

