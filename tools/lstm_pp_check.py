"""Validate and time the ping-pong LSTM kernel against torch.nn.LSTM and the production kernel."""
import ctypes, sys, torch
sys.path.insert(0, '/root/repo')
from speech_separation_b200 import _lib
lib = _lib.load()
dev = torch.device('cuda:0')
def P(t): return ctypes.c_void_p(t.data_ptr()) if t is not None else None
H = 128
def run(mode, B, S, C, N, ndir, act, pp, reps=1, time_it=False):
    torch.manual_seed(mode * 100 + B + S + C + N)
    rnn = torch.nn.LSTM(N, H, batch_first=True, bidirectional=(ndir == 2))
    with torch.no_grad():
        for p_ in rnn.parameters():
            p_.copy_(p_.half().float() if p_.dim() == 2 else p_)
    x = torch.randn(B, S, C, N).half()
    names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
    keep = [getattr(rnn, n + suf).detach().to(dev).contiguous() for suf in (["", "_reverse"][:ndir]) for n in names]
    table = (ctypes.c_void_p * 8)(*([t.data_ptr() for t in keep] + [0] * (8 - len(keep))))
    xd = x.to(dev)
    out = torch.full((B * S * C, ndir * H), float('nan'), dtype=torch.float16, device=dev)
    wpack = torch.empty(ndir * 512 * (N + H), dtype=torch.float16, device=dev)
    bpack = torch.empty(ndir * 512, dtype=torch.float32, device=dev)
    lib.vatss_debug_lstm_pingpong(pp)
    ms = None
    for _ in range(reps):
        _lib.check(lib.vatss_tc_lstm(P(xd), None, table, P(out), mode, B, S, C, N, ndir, act, P(wpack), P(bpack), None), "lstm")
    torch.cuda.synchronize()
    if time_it:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            _lib.check(lib.vatss_tc_lstm(P(xd), None, table, P(out), mode, B, S, C, N, ndir, act, P(wpack), P(bpack), None), "lstm")
        b.record(); torch.cuda.synchronize(); ms = a.elapsed_time(b) / 5
    lib.vatss_debug_lstm_pingpong(0)
    return rnn, x, out.clone(), ms
CASES = [(0, 1, 3, 5, 128, 1, 0), (0, 2, 70, 12, 128, 2, 1), (1, 3, 9, 150, 128, 2, 1), (1, 32, 6, 150, 128, 2, 0), (0, 5, 77, 10, 64, 2, 1), (1, 2, 7, 250, 64, 1, 0)]
for (mode, B, S, C, N, ndir, act) in CASES:
    rnn, x, got_pp, _ = run(mode, B, S, C, N, ndir, act, 1)
    _, _, got_old, _ = run(mode, B, S, C, N, ndir, act, 0)
    xf = x.float()
    seqs = xf.reshape(B * S, C, N) if mode == 0 else xf.permute(0, 2, 1, 3).reshape(B * C, S, N)
    with torch.no_grad(): ref = rnn(seqs)[0]
    ref = ref.reshape(B, S, C, ndir * H) if mode == 0 else ref.reshape(B, C, S, ndir * H).permute(0, 2, 1, 3)
    if act: ref = torch.relu(ref)
    g = got_pp.float().cpu().reshape(B, S, C, ndir * H)
    err = float((g - ref).norm() / ref.norm()) if torch.isfinite(g).all() else float('nan')
    same = torch.equal(got_pp, got_old)
    print(f"case mode={mode} B={B} S={S} C={C} N={N} ndir={ndir}: ping-pong rel err {err:.3e}, bit-identical to production kernel: {same}")
for mode in (0, 1):
    _, _, _, ms0 = run(mode, 32, 283, 150, 128, 2, 1, 0, reps=2, time_it=True)
    _, _, o1, ms1 = run(mode, 32, 283, 150, 128, 2, 1, 1, reps=2, time_it=True)
    print(f"cfg-2 mode {mode}: production {ms0:.3f} ms, ping-pong {ms1:.3f} ms")
