import os, sys, itertools, subprocess
sys.path.insert(0, os.getcwd())
if len(sys.argv) > 1:
    import torch
    from multimodal_clinical_b200 import _lib
    lib = _lib.load()
    def run(A,B,M,N,K,a_mn,b_mn,bn):
        out = torch.full((M,N), float('nan'), device='cuda')
        rc = lib.lf_debug_tc_gemm(A.data_ptr(), B.data_ptr(), None, out.data_ptr(), M,N,K, A.stride(0), B.stride(0), N, a_mn,b_mn,bn,1,M*N, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize(); return out
    torch.manual_seed(0)
    M,N,K=128,32,8
    A=torch.randn(M,K,device='cuda'); B=torch.zeros(K,N,device='cuda')
    for k in range(K): B[k,k]=1.0; B[k,8+k]=2.0; B[k, 16+k] = 3.0; B[k,24+k]=4.0
    out=run(A,B,M,N,K,0,1,32)
    ref=(A.double()@B.double()).float()
    err=((out-ref).norm()/ref.norm()).item()
    print("CFG", sys.argv[1], "err %.4f"%err, "nz", int((out!=0).sum()), "row0", [round(x,2) for x in out[0,:12].tolist()], "A0", [round(x,2) for x in A[0].tolist()])
else:
    for lt, swz, sbo in itertools.product([1,2,0], [4,3,0], [512,1024]):
        env=dict(os.environ, LF_TC_LT=str(lt), LF_TC_SWZ=str(swz), LF_TC_SBO=str(sbo))
        r=subprocess.run([sys.executable, __file__, f"lt={lt},swz={swz},sbo={sbo}"], env=env, capture_output=True, text=True, timeout=120)
        print((r.stdout.strip() or r.stderr.strip()[-300:]))
