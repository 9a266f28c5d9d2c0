import sys, time; sys.path.insert(0, '.')
import numpy as np
import ugemm_b200 as u
M,N,K=1023,1000,1023
for (ta,tb) in (("N","N"),("N","T"),("T","N"),("T","T")):
    for pad in (0,1):
        ar,ac = (M,K) if ta=="N" else (K,M)
        br,bc = (K,N) if tb=="N" else (N,K)
        lda, ldb = ac+pad if (ac+pad)%4 else ac+pad, bc
        lda = ac + pad; ldb = bc + pad*0
        dA=u.DeviceBuffer(ar*lda+8).fill_uniform(1); dB=u.DeviceBuffer(br*ldb+8).fill_uniform(2); dC=u.DeviceBuffer(M*N)
        for mode in ("auto","simt"):
            try:
                avg,best=u.sgemm_cuda_time_dev(mode, 10, 2, "R",ta,tb,M,N,K,1.0,dA,lda,dB,ldb,0.0,dC,N)
                print(ta,tb,"lda",lda,"ldb",ldb,mode,"kernel",u.last_kernel(),"repacked",u.last_repacked(),f"avg {avg*1e3:.1f} us best {best*1e3:.1f} us", flush=True)
            except Exception as e: print(ta,tb,mode,"ERR",e)
# host path timing
A=np.arange(1,K*M+1,dtype=np.float32); B=np.arange(1,K*N+1,dtype=np.float32); C=np.ones(M*N,np.float32)
for name,fn in (("rnn",u.sgemm_rnn),("rnt",u.sgemm_rnt),("rtn",u.sgemm_rtn)):
    fn(M,N,K,1.0,A,B,0.0,C)
    t=time.perf_counter()
    for _ in range(10): fn(M,N,K,1.0,A,B,0.0,C)
    print(name, "host path", (time.perf_counter()-t)/10*1e3, "ms", u.last_kernel(), u.last_repacked())
