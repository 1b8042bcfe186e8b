import ctypes, torch, sys
sys.path.insert(0,'/root/repo')
from speech_separation_b200 import _lib
lib=_lib.load()
dev=torch.device('cuda:0')
def P(t): return None if t is None else ctypes.c_void_p(t.data_ptr())
for (epi,NOUT,K,M,use_res) in [(1,256,128,777,True),(1,256,128,777,False),(1,128,128,256,True),(1,128,128,256,False)]:
    torch.manual_seed(0)
    A=torch.randn(M,K,device=dev).half(); W=(torch.randn(NOUT,K,device=dev)/K**0.5).half()
    bias=torch.randn(NOUT,device=dev); res=torch.randn(M,NOUT,device=dev)
    out=torch.full((M,NOUT),float('nan'),device=dev)
    rc=lib.vatss_tc_gemm(epi,P(A),K,P(W),P(bias),P(res) if use_res else None,NOUT,None,None,P(out),NOUT,None,0,0,None,M,NOUT,K,None)
    torch.cuda.synchronize()
    want=A.float()@W.float().t()+bias+(res if use_res else 0)
    err=(out-want).abs()
    bad=(err>1e-3)|torch.isnan(out)
    print(epi,NOUT,K,M,use_res,"bad",int(bad.sum()),"of",out.numel(), "max",float(err[~torch.isnan(err)].max()))
    if bad.any():
        idx=bad.nonzero()
        print(" rows",idx[:,0].unique()[:20].tolist()," cols",idx[:,1].unique()[:40].tolist())
