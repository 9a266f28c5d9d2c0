import sys; sys.path.insert(0, '.')
import ugemm_b200 as u
def t(M,N,K,beta,iters=6):
    dA=u.DeviceBuffer(M*K).fill_uniform(1); dB=u.DeviceBuffer(K*N).fill_uniform(2); dC=u.DeviceBuffer(M*N).fill_uniform(3)
    avg,best=u.sgemm_cuda_time_dev("3xtf32", iters, 2, "R","N","N",M,N,K,1.0,dA,K,dB,N,beta,dC,N)
    print(f"M={M} N={N} K={K} beta={beta}: avg {avg:.3f} ms min {best:.3f} ms  {2*M*N*K/avg/1e9:.1f} / {2*M*N*K/best/1e9:.1f} TF", flush=True)
    for d in (dA,dB,dC): d.free()
for beta in (0.0, 1.0):
    t(16384, 8192, 4096, beta)
    t(16384, 8192, 2048, beta)
    t(16384, 8192, 32768, beta, 3)
    t(16384, 32768, 4096, beta)
t(8192,8192,8192,0.0,10)
t(8192,8192,8192,1.0,10)
