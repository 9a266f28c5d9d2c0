import sys; sys.path.insert(0, '.')
import ugemm_b200 as u
def t(mode, cg, M, N, K, iters=20):
    u.set_k1_tuning(cta_group=cg)
    dA=u.DeviceBuffer(M*K).fill_uniform(1); dB=u.DeviceBuffer(K*N).fill_uniform(2); dC=u.DeviceBuffer(M*N)
    avg,best=u.sgemm_cuda_time_dev(mode, iters, 3, "R","N","N",M,N,K,1.0,dA,K,dB,N,0.0,dC,N)
    for d in (dA,dB,dC): d.free()
    return best*1e3, 2*M*N*K/best/1e9
for (M,N,K) in [(256,256,256),(512,512,512),(1024,1024,1024),(1536,1536,1536),(2048,2048,2048),(3072,3072,3072),(4096,4096,4096),(4096,3000,2048),(1024,1024,8192),(8192,1024,1024),(512,8192,512)]:
    r = {("3xtf32",1): t("3xtf32",1,M,N,K), ("3xtf32",2): t("3xtf32",2,M,N,K), ("simt",2): t("simt",2,M,N,K)}
    print(f"{M}x{N}x{K}: " + "  ".join(f"{m}/cg{c}: {us:8.1f} us {tf:6.1f} TF" for (m,c),(us,tf) in r.items()), flush=True)
