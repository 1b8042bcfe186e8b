"""Random shapes: the half-tile (ping-pong) LSTM kernel must be bit-identical to the one-tile kernel."""
import ctypes, random, sys, torch
sys.path.insert(0, '/root/repo')
from speech_separation_b200 import _lib
lib = _lib.load()
dev = torch.device('cuda:0')
def P(t): return ctypes.c_void_p(t.data_ptr()) if t is not None else None
H = 128
rng = random.Random(5)
bad = 0
for it in range(40):
    mode = rng.choice([0, 1]); N = rng.choice([64, 128]); ndir = rng.choice([1, 2]); act = rng.choice([0, 1])
    precise = (N == 64 and rng.random() < 0.4)
    B = rng.choice([1, 2, 3, 5, 8]); S = rng.choice([1, 2, 3, 7, 20, 61]); C = rng.choice([1, 2, 5, 33, 64, 65, 150])
    torch.manual_seed(it)
    rnn = torch.nn.LSTM(N, H, batch_first=True, bidirectional=(ndir == 2))
    names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
    keep = [getattr(rnn, n + suf).detach().to(dev).contiguous() for suf in (["", "_reverse"][:ndir]) for n in names]
    table = (ctypes.c_void_p * 8)(*([t.data_ptr() for t in keep] + [0] * (8 - len(keep))))
    xf = torch.randn(B, S, C, N, device=dev)
    x = xf.half(); xlo = (xf - x.float()).half() if precise else None
    nx = 2 * N if precise else N
    wpack = torch.empty(ndir * 512 * (nx + H), dtype=torch.float16, device=dev)
    bpack = torch.empty(ndir * 512, dtype=torch.float32, device=dev)
    outs = []
    for pp in (0, 1):
        lib.vatss_debug_lstm_pingpong(pp)
        out = torch.full((B * S * C, ndir * H), float('nan'), dtype=torch.float16, device=dev)
        _lib.check(lib.vatss_tc_lstm(P(x), P(xlo), table, P(out), mode, B, S, C, N, ndir, act, P(wpack), P(bpack), None), "lstm")
        torch.cuda.synchronize(); outs.append(out)
    ok = torch.equal(outs[0], outs[1]) and bool(torch.isfinite(outs[0].float()).all())
    if not ok:
        bad += 1
        print("MISMATCH", dict(mode=mode, N=N, ndir=ndir, act=act, precise=precise, B=B, S=S, C=C))
lib.vatss_debug_lstm_pingpong(1)
print("fuzz done, mismatches:", bad)
