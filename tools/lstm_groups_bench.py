"""Time of one LSTM launch at the production shapes for every row-group count of k_tc_lstm_pp (4 = dense 128-row
tiles, 3 / 2 = a group duplicated over both slots of a half tile, 0 = automatic), and bit-equality of the outputs."""
import ctypes, sys, torch
sys.path.insert(0, '/root/repo')
from speech_separation_b200 import _lib
lib = _lib.load()
dev = torch.device('cuda:0')
def P(t): return ctypes.c_void_p(t.data_ptr())
H, ndir = 128, 2
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
shapes = [("cfg-2 intra", 0, 32, 283, 150, 128), ("cfg-2 inter", 1, 32, 283, 150, 128),
          ("cfg-5 intra (batch 16)", 0, 16, 283, 150, 128), ("cfg-5 inter (batch 16)", 1, 16, 283, 150, 128),
          ("cfg-1 intra (batch 4, N=64)", 0, 4, 283, 150, 64), ("cfg-1 inter (batch 4, N=64)", 1, 4, 283, 150, 64),
          ("cfg-3 inter (16 x 10 s)", 1, 16, 710, 150, 128)]
for name, mode, B, S, C, N in shapes:
    torch.manual_seed(0)
    rnn = torch.nn.LSTM(N, H, batch_first=True, bidirectional=True)
    names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
    keep = [getattr(rnn, n + suf).detach().to(dev).contiguous() for suf in ("", "_reverse") for n in names]
    table = (ctypes.c_void_p * 8)(*[t.data_ptr() for t in keep])
    wpack = torch.empty(ndir * 512 * (N + H), dtype=torch.float16, device=dev)
    bpack = torch.empty(ndir * 512, dtype=torch.float32, device=dev)
    x = torch.randn(B, S, C, N, device=dev).half()
    ref, line = None, []
    for g in (4, 3, 2, 0):
        lib.vatss_debug_lstm_groups(g)
        out = torch.empty(B * S * C, 2 * H, dtype=torch.float16, device=dev)
        ts = []
        for it in range(6):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(lib.vatss_tc_lstm(P(x), None, table, P(out), mode, B, S, C, N, ndir, 1, P(wpack), P(bpack), None), "l")
            b.record(); torch.cuda.synchronize()
            if it >= 2: ts.append(a.elapsed_time(b))
        if ref is None: ref = out.clone()
        line.append(f"groups={g}: {sum(ts)/len(ts):.3f} ms{'' if torch.equal(ref, out) else ' MISMATCH'}")
    print(f"{name:30s}", "  ".join(line), flush=True)
lib.vatss_debug_lstm_groups(0)
