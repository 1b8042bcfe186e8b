import ctypes, sys, torch
sys.path.insert(0, '/root/repo')
from speech_separation_b200 import _lib
lib = _lib.load()
dev = torch.device('cuda:0')
def P(t): return ctypes.c_void_p(t.data_ptr())
lib.vatss_debug_lstm_pingpong(1)
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
        m = [int(v) - t0 for v in t[s, :8]]
        g = [int(v) - t0 for v in t[s, 8:24]]
        print(f" step {8+s}: MMA start@{m[0]} | X(A) done@{m[1]} hfullA wait->{m[2]} H(A) issued@{m[3]} | X(B) done@{m[4]} hfullB wait->{m[5]} H(B) issued@{m[6]}")
        print("          gates: " + " ".join(f"{'AB'[i//2]}{i%2}[wait {g[3*i+1]-g[3*i]} @{g[3*i+1]} math {g[3*i+2]-g[3*i+1]}]" for i in range(4)) + f" hA@{g[12]} hB@{g[13]}")
lib.vatss_debug_lstm_pingpong(0)
