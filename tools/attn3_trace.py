"""clock64 timeline of CTA 0 of the attention kernel (tc_attn3.cu): 8 consecutive jobs of softmax warpgroup 0 (its
softmax warp of quadrant 0 and the MMA issuer of its pipeline)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_separation_b200 import _lib
lib = _lib.load()
dev = torch.device('cuda:0')
def P(t): return ctypes.c_void_p(t.data_ptr())
NAMES = ["s_full seen", "softmax done", "P arrived", "issuer: waits P", "issuer: P seen", "issuer: O free", "issuer: PV issued",
         "issuer: next S issued", "softmax: waits S"]
for mode, B, S, C in [(0, 32, 283, 150), (1, 32, 283, 150)]:
    N, heads = 128, 4
    torch.manual_seed(0)
    qkv = torch.randn(B * S * C, 3 * N, device=dev).half()
    out = torch.empty(B * S * C, N, dtype=torch.float16, device=dev)
    trace = torch.zeros(8 * 16, dtype=torch.int64, device=dev)
    for it in range(3):
        lib.vatss_debug_lstm_trace(P(trace) if it == 2 else None)
        _lib.check(lib.vatss_tc_attention(P(qkv), P(out), mode, B, S, C, N, heads, 0, None), "attn")
    torch.cuda.synchronize()
    lib.vatss_debug_lstm_trace(None)
    t = trace.cpu().reshape(8, 16)
    t0 = int(t[0, 0])
    print("mode", mode, "(cycles relative to the first traced job's s_full)")
    for s in range(8):
        a = [int(v) - t0 if int(v) else None for v in t[s, :15]]
        print(f" job {s}: " + " | ".join(f"{n} {v}" for n, v in zip(NAMES, a) if n != "-" and v is not None))
