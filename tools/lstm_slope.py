"""Per-step time and fixed overhead of the LSTM kernels from launches with different numbers of time steps."""
import ctypes, sys, torch
sys.path.insert(0, '/root/repo')
from speech_separation_b200 import _lib
lib = _lib.load()
dev = torch.device('cuda:0')
def P(t): return ctypes.c_void_p(t.data_ptr())
N, H, ndir = 128, 128, 2
torch.manual_seed(0)
rnn = torch.nn.LSTM(N, H, batch_first=True, bidirectional=True)
names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
keep = [getattr(rnn, n + suf).detach().to(dev).contiguous() for suf in ("", "_reverse") for n in names]
table = (ctypes.c_void_p * 8)(*[t.data_ptr() for t in keep])
wpack = torch.empty(ndir * 512 * (N + H), dtype=torch.float16, device=dev)
bpack = torch.empty(ndir * 512, dtype=torch.float32, device=dev)
for pp in (0, 1):
    lib.vatss_debug_lstm_pingpong(pp)
    res = []
    for C in (50, 150, 300, 600):
        B, S = 32, 283          # 9056 sequences = 144 CTAs, intra mode: time steps = C
        x = torch.randn(B, S, C, N, device=dev).half()
        out = torch.empty(B * S * C, 2 * H, dtype=torch.float16, device=dev)
        for _ in range(2):
            _lib.check(lib.vatss_tc_lstm(P(x), None, table, P(out), 0, B, S, C, N, ndir, 1, P(wpack), P(bpack), None), "l")
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            _lib.check(lib.vatss_tc_lstm(P(x), None, table, P(out), 0, B, S, C, N, ndir, 1, P(wpack), P(bpack), None), "l")
        b.record(); torch.cuda.synchronize()
        res.append((C, a.elapsed_time(b) / 5))
        del x, out
    (c1, t1), (c2, t2) = res[1], res[3]
    slope = (t2 - t1) / (c2 - c1)
    print(f"pingpong={pp}:", [(c, round(t, 3)) for c, t in res], f"per step {slope*1e3:.2f} us = {slope*1e3*1965:.0f} cycles @1965 MHz, fixed {t1 - slope*c1:.3f} ms")
lib.vatss_debug_lstm_pingpong(0)
