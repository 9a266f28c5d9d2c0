import sys, time, ctypes as C; sys.path.insert(0, '.')
import ugemm_b200 as u
M=N=K=8192
L=u.lib(); u.sgemm_cuda_init(0)
hA,hB,hC=(L.ugemm_cuda_malloc_host(n*4) for n in (M*K,K*N,M*N))
u.lib().ugemm_fill_uniform_host(hA, M*K, 1, 0.0, 1.0); u.lib().ugemm_fill_uniform_host(hB, K*N, 2, 0.0, 1.0)
for _ in range(2): u.sgemm_cuda("R","N","N",M,N,K,1.0,hA,K,hB,N,0.0,hC,N)
t=time.perf_counter()
for _ in range(8): u.sgemm_cuda("R","N","N",M,N,K,1.0,hA,K,hB,N,0.0,hC,N)
dt=(time.perf_counter()-t)/8
print("e2e ms", dt*1e3, "TF", 2*M*N*K/dt/1e12, flush=True)
