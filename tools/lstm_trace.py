"""Print a clock64 timeline of CTA 0 of the LSTM kernel (MMA thread and one gate warp) for 4 time steps."""
import ctypes, sys, torch
sys.path.insert(0, '/root/repo')
from speech_separation_b200 import _lib
lib = _lib.load()
dev = torch.device('cuda:0')
def P(t): return ctypes.c_void_p(t.data_ptr())
for mode, B, S, C in [(0, 32, 283, 150), (1, 32, 283, 150)]:
    N, H, ndir = 128, 128, 2
    torch.manual_seed(0)
    rnn = torch.nn.LSTM(N, H, batch_first=True, bidirectional=True)
    names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
    keep = [getattr(rnn, n + suf).detach().to(dev).contiguous() for suf in ("", "_reverse") for n in names]
    table = (ctypes.c_void_p * 8)(*[t.data_ptr() for t in keep])
    x = torch.randn(B, S, C, N, device=dev).half()
    out = torch.empty(B * S * C, 2 * H, dtype=torch.float16, device=dev)
    wpack = torch.empty(ndir * 512 * (N + H), dtype=torch.float16, device=dev)
    bpack = torch.empty(ndir * 512, dtype=torch.float32, device=dev)
    trace = torch.zeros(128, dtype=torch.int64, device=dev)
    for it in range(3):
        lib.vatss_debug_lstm_trace(P(trace) if it == 2 else None)
        _lib.check(lib.vatss_tc_lstm(P(x), None, table, P(out), mode, B, S, C, N, ndir, 1, P(wpack), P(bpack), None), "lstm")
    torch.cuda.synchronize()
    lib.vatss_debug_lstm_trace(None)
    t = trace.cpu().reshape(4, 32)
    t0 = int(t[0, 0])
    print("mode", mode)
    for s in range(4):
        m = [int(v) - t0 for v in t[s, :7]]
        g = [int(v) - t0 for v in t[s, 8:21]]
        print(f" step {8+s}: MMA start {m[0]} xfull+{m[1]-m[0]} | X012 done@{m[2]} hfull wait {m[3]-m[2]} | H012 issued@{m[4]} accempty3 wait {m[5]-m[4]} end@{m[6]}")
        print("          gate warp: " + " ".join(f"c{c}[wait {g[3*c+1]-g[3*c]} @{g[3*c+1]} math {g[3*c+2]-g[3*c+1]}]" for c in range(4)) + f" harrive@{g[12]}")
