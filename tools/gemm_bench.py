"""Time the three GEMM stages of a DPTN sub-block alone at the cfg-2 size (M = 1 358 400 tokens, L2 flushed between
launches) and check them against torch on the first rows: QKV (fp16 out), out-projection + LayerNorm (fp32 residual),
FFN + LayerNorm (fp16 residual)."""
import ctypes, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_separation_b200 import _lib
lib = _lib.load()
dev = torch.device('cuda:0')
def P(t): return ctypes.c_void_p(t.data_ptr()) if t is not None else None
torch.manual_seed(0)
M = 32 * 283 * 150
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, NOUT, K in (("qkv F16", 384, 128), ("outproj LN", 128, 128), ("ffn LN16", 128, 256)):
    A = torch.randn(M, K, device=dev).half()
    W = (torch.randn(NOUT, K, device=dev) / K ** 0.5).half()
    bias = torch.randn(NOUT, device=dev)
    res = torch.randn(M, NOUT, device=dev) if name.startswith("outproj") else None
    res16 = torch.randn(M, NOUT, device=dev).half() if name.startswith("ffn") else None
    lw = torch.rand(NOUT, device=dev) + 0.5; lb = torch.randn(NOUT, device=dev)
    o32 = torch.zeros(M, NOUT, device=dev) if not name.startswith("qkv") else None
    o16 = torch.zeros(M, NOUT, device=dev, dtype=torch.float16)
    def run():
        if name.startswith("qkv"):
            return lib.vatss_tc_gemm(0, P(A), K, P(W), P(bias), None, 0, None, None, None, 0, P(o16), NOUT, 0, None, M, NOUT, K, None)
        if name.startswith("outproj"):
            return lib.vatss_tc_gemm(2, P(A), K, P(W), P(bias), P(res), NOUT, P(lw), P(lb), None, NOUT, P(o16), NOUT, 0, None, M, NOUT, K, None)
        return lib.vatss_tc_gemm_ln16(P(A), K, P(W), P(bias), P(res16), NOUT, P(lw), P(lb), P(o32), NOUT, P(o16), NOUT, 0, None, M, NOUT, K, None)
    ts = []
    for it in range(7):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); _lib.check(run(), name); b.record(); torch.cuda.synchronize()
        if it >= 2: ts.append(a.elapsed_time(b))
    n = 4096
    ref = A[:n].float() @ W.float().t() + bias
    if res is not None: ref = ref + res[:n]
    if res16 is not None: ref = ref + res16[:n].float()
    if not name.startswith("qkv"): ref = torch.nn.functional.layer_norm(ref, (NOUT,), lw, lb, 1e-5)
    err = ((o16[:n].float() - ref).norm() / ref.norm()).item()
    gb = M * (K * 2 + NOUT * 2 + (NOUT * 4 if res is not None else 0) + (NOUT * 2 if res16 is not None else 0) + (NOUT * 4 if (o32 is not None and name.startswith("ffn")) else 0)) / 1e9
    print(f"{name:12s} {min(ts):.3f} ms (median {sorted(ts)[len(ts)//2]:.3f})  {gb:.2f} GB -> {gb / min(ts):.2f} TB/s  rel err vs torch {err:.2e}", flush=True)
    del A, res, res16, o32, o16
