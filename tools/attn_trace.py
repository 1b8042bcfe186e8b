"""Print a clock64 timeline of CTA 0 of the attention kernel (softmax warp 0 and MMA warp 0) for 4 groups."""
import ctypes, sys, torch
sys.path.insert(0, '/root/repo')
from speech_separation_b200 import _lib
lib = _lib.load()
dev = torch.device('cuda:0')
def P(t): return ctypes.c_void_p(t.data_ptr())
for mode, B, S, C in [(0, 32, 283, 150), (1, 32, 283, 150)]:
    N, heads = 128, 4
    torch.manual_seed(0)
    qkv = torch.randn(B * S * C, 3 * N, device=dev).half()
    out = torch.empty(B * S * C, N, dtype=torch.float16, device=dev)
    trace = torch.zeros(128, dtype=torch.int64, device=dev)
    for it in range(3):
        lib.vatss_debug_lstm_trace(P(trace) if it == 2 else None)
        _lib.check(lib.vatss_tc_attention(P(qkv), P(out), mode, B, S, C, N, heads, 0, None), "attn")
    torch.cuda.synchronize()
    lib.vatss_debug_lstm_trace(None)
    t = trace.cpu().reshape(4, 32)
    t0 = int(t[0, 0])
    nb = 3 if mode == 0 else 5
    print("mode", mode)
    for s in range(4):
        a = [int(v) - t0 for v in t[s, :16]]
        print(f" group {8+s}: start@{a[0]} | max pass {a[1]-a[0]} | bar {a[2]-a[1]} | exp pass {a[3]-a[2]} (blocks done@" + ",".join(str(a[8+j]) for j in range(nb)) + f") | deferred read-out {a[4]-a[3]} end@{a[4]}")
