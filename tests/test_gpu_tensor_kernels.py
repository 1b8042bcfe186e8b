"""Unit tests of the tcgen05 TENSOR-engine kernels through their stand-alone C-ABI entry points.

Each kernel is compared with a plain PyTorch fp32 evaluation of the same op on the same
fp16-rounded operands (so the only difference is fp32 accumulation order)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    assert torch.cuda.is_available()
    from speech_separation_b200 import _lib

    return _lib.load()


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def ln(x, w, b):
    return torch.nn.functional.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


GEMM_CASES = [
    # epi, NOUT, K, M
    (0, 384, 128, 128), (0, 384, 128, 1000), (0, 384, 128, 150 * 283 + 5),
    (1, 256, 128, 777), (1, 128, 128, 4097),
    (2, 128, 128, 300), (2, 128, 256, 12345), (3, 128, 256, 513),
    (0, 192, 64, 999), (2, 64, 64, 1025), (2, 64, 256, 2050), (3, 64, 256, 400), (1, 128, 64, 640), (1, 64, 64, 130),
]


@pytest.mark.parametrize("epi,NOUT,K,M", GEMM_CASES)
def test_tc_gemm_matches_torch(lib, epi, NOUT, K, M):
    from speech_separation_b200 import _lib

    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(M + NOUT + K + epi)
    A = torch.randn(M, K, generator=g).to(dev).half()
    W = (torch.randn(NOUT, K, generator=g) / K ** 0.5).to(dev).half()
    bias = torch.randn(NOUT, generator=g).to(dev)
    res = torch.randn(M, NOUT, generator=g).to(dev)
    lw = (1 + 0.2 * torch.randn(NOUT, generator=g)).to(dev)
    lb = (0.1 * torch.randn(NOUT, generator=g)).to(dev)
    slope = torch.tensor([0.25], device=dev)
    out32 = torch.full((M, NOUT), float("nan"), device=dev)
    out16 = torch.full((M, NOUT), float("nan"), device=dev, dtype=torch.float16)
    act16 = 2 if epi == 2 else 0
    rc = lib.vatss_tc_gemm(epi, _p(A), K, _p(W), _p(bias), _p(res), NOUT, _p(lw), _p(lb), _p(out32), NOUT, _p(out16),
                           NOUT, act16, _p(slope), M, NOUT, K, None)
    _lib.check(rc, "vatss_tc_gemm")
    torch.cuda.synchronize()
    base = A.float() @ W.float().t() + bias
    if epi == 0:
        assert torch.allclose(out16.float(), base, atol=2e-2, rtol=2e-3)
        err = (out16.float() - base).norm() / base.norm()
        assert err < 1e-3
    elif epi == 1:
        want = base + res
        assert (out32 - want).norm() / want.norm() < 1e-5
    elif epi == 2:
        want = ln(base + res, lw, lb)
        assert (out32 - want).norm() / want.norm() < 1e-5
        want16 = torch.where(want >= 0, want, 0.25 * want)
        assert (out16.float() - want16).norm() / want16.norm() < 1e-3
    else:
        want = ln(base, lw, lb) + res
        assert (out32 - want).norm() / want.norm() < 1e-5


@pytest.mark.parametrize("NOUT,K,M,with32", [(128, 128, 4321, True), (128, 256, 2000, True), (128, 256, 999, False),
                                               (64, 64, 1500, True), (64, 256, 640, False), (128, 128, 128, False)])
def test_tc_gemm_layernorm_with_fp16_residual(lib, NOUT, K, M, with32):
    """LayerNorm epilogue whose residual tile is the fp16 tensor (DPTN sub-blocks); fp32 output optional."""
    from speech_separation_b200 import _lib

    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(M + NOUT + K)
    A = torch.randn(M, K, generator=g).to(dev).half()
    W = (torch.randn(NOUT, K, generator=g) / K ** 0.5).to(dev).half()
    bias = torch.randn(NOUT, generator=g).to(dev)
    res16 = torch.randn(M, NOUT, generator=g).to(dev).half()
    lw = (1 + 0.2 * torch.randn(NOUT, generator=g)).to(dev)
    lb = (0.1 * torch.randn(NOUT, generator=g)).to(dev)
    out32 = torch.full((M, NOUT), float("nan"), device=dev)
    out16 = torch.full((M, NOUT), float("nan"), device=dev, dtype=torch.float16)
    rc = lib.vatss_tc_gemm_ln16(_p(A), K, _p(W), _p(bias), _p(res16), NOUT, _p(lw), _p(lb),
                                _p(out32) if with32 else None, NOUT, _p(out16), NOUT, 1, None, M, NOUT, K, None)
    _lib.check(rc, "vatss_tc_gemm_ln16")
    torch.cuda.synchronize()
    want = ln(A.float() @ W.float().t() + bias + res16.float(), lw, lb)
    if with32:
        assert (out32 - want).norm() / want.norm() < 1e-5
    else:
        assert torch.isnan(out32).all()          # untouched
    want16 = torch.relu(want)
    assert (out16.float() - want16).norm() / want16.norm() < 1e-3


@pytest.mark.parametrize("M", [42450, 84900, 37905])
def test_tc_gemm_layernorm_epilogue_is_race_free(lib, M):
    """Repeated launches must be bit-identical.  Regression test for a cross-proxy write-after-read race: the
    epilogue released the TMA-written residual tile right after its shared-memory reads without a proxy fence, and
    the next tile's TMA load could overwrite rows that were still being read (a few runs in sixty differed)."""
    from speech_separation_b200 import _lib

    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(M)
    A = torch.randn(M, 128, generator=g).to(dev).half()
    W = (torch.randn(128, 128, generator=g) * 0.1).to(dev).half()
    bias = torch.randn(128, generator=g).to(dev)
    res = torch.randn(M, 128, generator=g).to(dev)
    res16 = res.half()
    lw, lb = torch.ones(128, device=dev), torch.zeros(128, device=dev)
    o16 = torch.empty(M, 128, dtype=torch.float16, device=dev)
    for variant in ("fp32 residual", "fp16 residual"):
        ref = None
        for _ in range(40):
            if variant == "fp32 residual":
                rc = lib.vatss_tc_gemm(2, _p(A), 128, _p(W), _p(bias), _p(res), 128, _p(lw), _p(lb), None, 128, _p(o16), 128,
                                       0, None, M, 128, 128, None)
            else:
                rc = lib.vatss_tc_gemm_ln16(_p(A), 128, _p(W), _p(bias), _p(res16), 128, _p(lw), _p(lb), None, 128, _p(o16),
                                            128, 0, None, M, 128, 128, None)
            _lib.check(rc, "gemm")
            torch.cuda.synchronize()
            if ref is None:
                ref = o16.clone()
            else:
                assert torch.equal(ref, o16), variant
        want = ln(A.float() @ W.float().t() + bias + (res if variant == "fp32 residual" else res16.float()), lw, lb)
        assert (ref.float() - want).norm() / want.norm() < 1e-3


def test_tc_gemm_strided_operand(lib):
    """A taken as a column slice of a wider matrix (the per-speaker halves of the overlap-add output)."""
    from speech_separation_b200 import _lib

    dev = torch.device("cuda:0")
    M, K, NOUT = 1000, 128, 128
    big = torch.randn(M, 2 * K, device=dev).half()
    W = (torch.randn(NOUT, K, device=dev) / K ** 0.5).half()
    bias = torch.randn(NOUT, device=dev)
    for j in range(2):
        A = big[:, j * K:(j + 1) * K]
        out32 = torch.empty(M, NOUT, device=dev)
        rc = lib.vatss_tc_gemm(1, ctypes.c_void_p(A.data_ptr()), 2 * K, _p(W), _p(bias), None, 0, None, None, _p(out32),
                               NOUT, None, 0, 0, None, M, NOUT, K, None)
        _lib.check(rc, "vatss_tc_gemm")
        want = A.float() @ W.float().t() + bias
        assert (out32 - want).norm() / want.norm() < 1e-5


LSTM_CASES = [
    # mode, B, S, C, N, ndir, act
    (0, 1, 3, 5, 128, 1, 0),       # one partial tile, forward only, tiny
    (0, 2, 70, 12, 128, 2, 1),     # 140 sequences -> 2 tiles (1 pair), both directions
    (1, 3, 9, 150, 128, 2, 1),     # inter: sequences (b,k), time = chunk index
    (1, 32, 6, 150, 128, 2, 0),    # inter with the production tile shape (30 x 4)
    (0, 5, 77, 10, 64, 2, 1),      # N = 64 models, 385 sequences -> 4 tiles (2 pairs)
    (1, 2, 7, 250, 64, 1, 0),      # DPRNN-like chunk, unidirectional inter
]


@pytest.mark.parametrize("pingpong", [1, 0], ids=["half-tiles", "one-tile"])
@pytest.mark.parametrize("mode,B,S,C,N,ndir,act", LSTM_CASES)
def test_tc_lstm_matches_torch(lib, mode, B, S, C, N, ndir, act, pingpong):
    """Both LSTM kernels: k_tc_lstm_pp (two interleaved 64-row half tiles per CTA, the default) and k_tc_lstm."""
    from speech_separation_b200 import _lib

    lib.vatss_debug_lstm_pingpong(pingpong)

    dev = torch.device("cuda:0")
    torch.manual_seed(mode * 100 + B + S + C + N)
    H = 128
    rnn = torch.nn.LSTM(N, H, batch_first=True, bidirectional=(ndir == 2))
    with torch.no_grad():
        for p_ in rnn.parameters():
            p_.copy_(p_.half().float() if p_.dim() == 2 else p_)  # fp16-representable weights
    x = torch.randn(B, S, C, N).half()
    xf = x.float()
    seqs = xf.reshape(B * S, C, N) if mode == 0 else xf.permute(0, 2, 1, 3).reshape(B * C, S, N)
    with torch.no_grad():
        ref = rnn(seqs)[0]  # (G, len, ndir*H), CPU fp32
    ref = ref.reshape(B, S, C, ndir * H) if mode == 0 else ref.reshape(B, C, S, ndir * H).permute(0, 2, 1, 3)
    if act:
        ref = torch.relu(ref)
    names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
    keep = [getattr(rnn, n + suf).detach().to(dev).contiguous() for suf in (["", "_reverse"][:ndir]) for n in names]
    table = (ctypes.c_void_p * 8)(*([t.data_ptr() for t in keep] + [None] * (8 - len(keep))))
    xd = x.to(dev).contiguous()
    out = torch.full((B * S * C, ndir * H), float("nan"), dtype=torch.float16, device=dev)
    wpack = torch.empty(ndir * 512 * (N + H), dtype=torch.float16, device=dev)
    bpack = torch.empty(ndir * 512, dtype=torch.float32, device=dev)
    rc = lib.vatss_tc_lstm(_p(xd), None, table, _p(out), mode, B, S, C, N, ndir, act, _p(wpack), _p(bpack), None)
    _lib.check(rc, "vatss_tc_lstm")
    torch.cuda.synchronize()
    got = out.float().cpu().reshape(B, S, C, ndir * H)
    assert torch.isfinite(got).all()
    err = (got - ref).norm() / ref.norm()
    print(f"tc_lstm mode={mode} B={B} S={S} C={C} N={N} ndir={ndir}: rel err {err:.3e}, max abs {(got - ref).abs().max():.3e}")
    lib.vatss_debug_lstm_pingpong(1)
    assert err < 3e-3  # fp16 h feedback + fp16 output + tanh.approx


ATT_CASES = [
    # mode, B, S, C, N, heads
    (0, 1, 2, 150, 128, 4),      # intra, production chunk, single kv block
    (0, 2, 5, 150, 128, 4),
    (1, 2, 141, 6, 128, 4),      # inter 2 s: len 141 (two query tiles, one kv block)
    (1, 1, 283, 4, 128, 4),      # inter 4 s: len 283 -> two kv blocks of 144 (two-pass softmax)
    (0, 3, 4, 40, 128, 4),       # short sequences
    (0, 2, 3, 150, 64, 4),       # N = 64: head dim 16, one group of four heads
    (1, 2, 283, 3, 64, 4),
    (0, 1, 2, 250, 128, 4),      # longer chunk: two kv blocks of 128
    (1, 1, 710, 3, 128, 4),      # inter 10 s: K / V streamed through the shared-memory ring
    (1, 2, 400, 2, 64, 4),
    # ragged last query tile (<= 32 rows: replicated-quadrant / 8-column-strip mode) and its boundaries
    (0, 3, 2, 129, 128, 4),      # 1 ragged row, 3 kv blocks (resident)
    (0, 2, 2, 160, 128, 4),      # 32 ragged rows
    (0, 2, 2, 161, 128, 4),      # 33 rows in the last tile: regular mode
    (0, 5, 3, 20, 128, 4),       # the only tile is ragged
    (0, 4, 2, 33, 128, 4),
    (0, 2, 2, 128, 128, 4),      # exactly one full tile
    (0, 2, 2, 192, 128, 4),      # 3 full kv blocks, last query tile 64 rows
    (0, 2, 1, 193, 128, 4),      # 4 kv blocks (two-pass), last kv block 1 row
    (1, 2, 257, 3, 128, 4),      # 3 query tiles, last with 1 row; 5 kv blocks
    (1, 3, 150, 5, 64, 4),       # head dim 16 with a ragged tile
]


def _attention_reference(qkv, mode, B, S, C, N, heads, dev):
    hd = N // heads
    x = qkv.float().to(dev)
    seq = x.reshape(B * S, C, 3 * N) if mode == 0 else x.permute(0, 2, 1, 3).reshape(B * C, S, 3 * N)
    G, Ls = seq.shape[0], seq.shape[1]
    q, k, v = (seq[..., i * N:(i + 1) * N].reshape(G, Ls, heads, hd).transpose(1, 2) for i in range(3))
    p = torch.softmax((q @ k.transpose(-1, -2)) * 0.6931471805599453, dim=-1)  # exp2 scores
    ref = (p @ v).transpose(1, 2).reshape(G, Ls, N)
    return ref.reshape(B, S, C, N) if mode == 0 else ref.reshape(B, C, S, N).permute(0, 2, 1, 3)


@pytest.mark.parametrize("version", [3])
@pytest.mark.parametrize("mode,B,S,C,N,heads", [(0, 2, 3, 150, 128, 4), (1, 2, 283, 3, 128, 4), (1, 1, 710, 2, 128, 4),
                                                (0, 2, 2, 150, 64, 4), (1, 2, 200, 2, 64, 4)])
def test_tc_attention_growing_logits_force_the_rescale_path(lib, mode, B, S, C, N, heads, version):
    """Keys late in the sequence get much larger logits than the first chunk: the kernel accumulates O in TMEM with the
    first chunk's row maximum as reference and must take its rescale path (difference > 2^8)."""
    from speech_separation_b200 import _lib

    dev = torch.device("cuda:0")
    torch.manual_seed(11 * mode + S + C + N)
    hd = N // heads
    qkv = torch.randn(B, S, C, 3 * N)
    qkv[..., :N] *= 1.4426950408889634 / hd ** 0.5
    L = C if mode == 0 else S
    ramp = torch.linspace(0.3, 6.0, L)                       # key scale grows along the sequence: logits up to +-30
    if mode == 0:
        qkv[..., N:2 * N] *= ramp[None, None, :, None]
    else:
        qkv[..., N:2 * N] *= ramp[None, :, None, None]
    qkv = qkv.half()
    ref = _attention_reference(qkv, mode, B, S, C, N, heads, dev)
    qd = qkv.to(dev).contiguous()
    out = torch.full((B * S * C, N), float("nan"), dtype=torch.float16, device=dev)
    lib.vatss_debug_attention_version(version)
    rc = lib.vatss_tc_attention(_p(qd), _p(out), mode, B, S, C, N, heads, 0, None)
    lib.vatss_debug_attention_version(3)
    _lib.check(rc, "vatss_tc_attention")
    torch.cuda.synchronize()
    got = out.float().reshape(B, S, C, N)
    assert torch.isfinite(got).all()
    err = ((got - ref).norm() / ref.norm()).item()
    print(f"attention v{version} rescale case mode={mode} S={S} C={C} N={N}: rel err {err:.3e}")
    assert err < 2e-3


@pytest.mark.parametrize("force_simt", [0, 1, -1], ids=["v3_tmem", "simt", "v1_smem"])
@pytest.mark.parametrize("mode,B,S,C,N,heads", ATT_CASES)
def test_tc_attention_matches_torch(lib, mode, B, S, C, N, heads, force_simt):
    """force_simt 0: tcgen05 kernel v3 (P and O in TMEM, tc_attn3.cu, the default); 1: SIMT fallback; -1: round-1 tcgen05
    kernel (P through shared memory)."""
    from speech_separation_b200 import _lib

    lib.vatss_debug_attention_version(1 if force_simt < 0 else 3)
    force_simt = max(force_simt, 0)

    dev = torch.device("cuda:0")
    torch.manual_seed(7 * mode + B + S + C + N)
    hd = N // heads
    qkv = torch.randn(B, S, C, 3 * N)
    qkv[..., :N] *= 1.4426950408889634 / hd ** 0.5 * 2.0  # pre-scaled q (x2 for a peakier softmax)
    qkv = qkv.half()
    x = qkv.float().to(dev)
    seq = x.reshape(B * S, C, 3 * N) if mode == 0 else x.permute(0, 2, 1, 3).reshape(B * C, S, 3 * N)
    G, Ls = seq.shape[0], seq.shape[1]
    q, k, v = (seq[..., i * N:(i + 1) * N].reshape(G, Ls, heads, hd).transpose(1, 2) for i in range(3))
    p = torch.softmax((q @ k.transpose(-1, -2)) * 0.6931471805599453, dim=-1)  # exp2 scores
    ref = (p @ v).transpose(1, 2).reshape(G, Ls, N)
    ref = ref.reshape(B, S, C, N) if mode == 0 else ref.reshape(B, C, S, N).permute(0, 2, 1, 3)
    qd = qkv.to(dev).contiguous()
    out = torch.full((B * S * C, N), float("nan"), dtype=torch.float16, device=dev)
    rc = lib.vatss_tc_attention(_p(qd), _p(out), mode, B, S, C, N, heads, force_simt, None)
    lib.vatss_debug_attention_version(3)
    _lib.check(rc, "vatss_tc_attention")
    torch.cuda.synchronize()
    got = out.float().reshape(B, S, C, N)
    assert torch.isfinite(got).all()
    err = ((got - ref).norm() / ref.norm()).item()
    print(f"attention simt={force_simt} mode={mode} B={B} S={S} C={C} N={N}: rel err {err:.3e}")
    assert err < 2e-3


def test_tc_lstm_kernels_agree_on_random_shapes(lib):
    """The half-tile (ping-pong) kernel and the one-tile kernel do the same arithmetic: their outputs must be
    bit-identical on random shapes (single time step, partial tiles, one direction, hi/lo split, both widths)."""
    import random

    from speech_separation_b200 import _lib

    dev = torch.device("cuda:0")
    rng = random.Random(5)
    H = 128
    try:
        for it in range(24):
            mode, N, ndir, act = rng.choice([0, 1]), rng.choice([64, 128]), rng.choice([1, 2]), rng.choice([0, 1])
            precise = N == 64 and rng.random() < 0.4
            B, S, C = rng.choice([1, 2, 3, 5, 8]), rng.choice([1, 2, 3, 7, 20, 61]), rng.choice([1, 2, 5, 33, 64, 65, 150])
            torch.manual_seed(it)
            rnn = torch.nn.LSTM(N, H, batch_first=True, bidirectional=(ndir == 2))
            names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
            keep = [getattr(rnn, n + suf).detach().to(dev).contiguous() for suf in (["", "_reverse"][:ndir]) for n in names]
            table = (ctypes.c_void_p * 8)(*([t.data_ptr() for t in keep] + [0] * (8 - len(keep))))
            xf = torch.randn(B, S, C, N, device=dev)
            x = xf.half()
            xlo = (xf - x.float()).half() if precise else None
            wpack = torch.empty(ndir * 512 * ((2 * N if precise else N) + H), dtype=torch.float16, device=dev)
            bpack = torch.empty(ndir * 512, dtype=torch.float32, device=dev)
            outs = []
            # one-tile kernel, then the half-tile kernel with 4 / 3 / 2 row groups per CTA and its automatic choice
            for pingpong, groups in ((0, 0), (1, 4), (1, 3), (1, 2), (1, 0)):
                lib.vatss_debug_lstm_pingpong(pingpong)
                lib.vatss_debug_lstm_groups(groups)
                out = torch.full((B * S * C, ndir * H), float("nan"), dtype=torch.float16, device=dev)
                rc = lib.vatss_tc_lstm(_p(x), _p(xlo) if precise else None, table, _p(out), mode, B, S, C, N, ndir, act,
                                       _p(wpack), _p(bpack), None)
                _lib.check(rc, "vatss_tc_lstm")
                torch.cuda.synchronize()
                outs.append(out)
            assert torch.isfinite(outs[0].float()).all()
            for k, o in enumerate(outs[1:]):
                assert torch.equal(outs[0], o), dict(variant=k + 1, mode=mode, N=N, ndir=ndir, precise=precise, B=B, S=S, C=C)
    finally:
        lib.vatss_debug_lstm_pingpong(1)
        lib.vatss_debug_lstm_groups(0)


@pytest.mark.parametrize("groups", [2, 3])
@pytest.mark.parametrize("mode,B,S,C,N,ndir,act", LSTM_CASES)
def test_tc_lstm_spread_groups_match_dense_tiles(lib, mode, B, S, C, N, ndir, act, groups):
    """k_tc_lstm_pp with 2 / 3 row groups per CTA (a group duplicated over both 32-row slots of a half tile, so that the
    inter-chunk recurrence spreads over more SMs) is bit-identical to the dense 4-group tiles, partial groups included."""
    from speech_separation_b200 import _lib

    dev = torch.device("cuda:0")
    torch.manual_seed(7 + mode + B + S + C)
    H = 128
    rnn = torch.nn.LSTM(N, H, batch_first=True, bidirectional=(ndir == 2))
    names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
    keep = [getattr(rnn, n + suf).detach().to(dev).contiguous() for suf in (["", "_reverse"][:ndir]) for n in names]
    table = (ctypes.c_void_p * 8)(*([t.data_ptr() for t in keep] + [None] * (8 - len(keep))))
    x = torch.randn(B, S, C, N, device=dev).half()
    wpack = torch.empty(ndir * 512 * (N + H), dtype=torch.float16, device=dev)
    bpack = torch.empty(ndir * 512, dtype=torch.float32, device=dev)
    outs = []
    try:
        for g in (4, groups):
            lib.vatss_debug_lstm_groups(g)
            out = torch.full((B * S * C, ndir * H), float("nan"), dtype=torch.float16, device=dev)
            rc = lib.vatss_tc_lstm(_p(x), None, table, _p(out), mode, B, S, C, N, ndir, act, _p(wpack), _p(bpack), None)
            _lib.check(rc, "vatss_tc_lstm")
            torch.cuda.synchronize()
            outs.append(out)
    finally:
        lib.vatss_debug_lstm_groups(0)
    assert torch.isfinite(outs[0].float()).all()
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("mode,B,S,C", [(0, 2, 40, 25), (1, 3, 11, 250)])
def test_tc_lstm_hi_lo_split_variant(lib, mode, B, S, C):
    """PRECISE variant (DPRNN): fp32 weights and inputs, hi/lo fp16 splits inside; must track the fp32 LSTM much
    more closely than the plain fp16 kernel (whose error is dominated by the fp16 rounding of W_ih and x)."""
    from speech_separation_b200 import _lib

    dev = torch.device("cuda:0")
    torch.manual_seed(3 + mode)
    N, H, ndir = 64, 128, 2
    rnn = torch.nn.LSTM(N, H, batch_first=True, bidirectional=True)
    x = 3.0 * torch.randn(B, S, C, N)   # un-normalised magnitudes like DPRNN's residual stream
    seqs = x.reshape(B * S, C, N) if mode == 0 else x.permute(0, 2, 1, 3).reshape(B * C, S, N)
    with torch.no_grad():
        ref = rnn(seqs)[0]
    ref = ref.reshape(B, S, C, ndir * H) if mode == 0 else ref.reshape(B, C, S, ndir * H).permute(0, 2, 1, 3)
    names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
    keep = [getattr(rnn, n + suf).detach().to(dev).contiguous() for suf in ("", "_reverse") for n in names]
    table = (ctypes.c_void_p * 8)(*[t.data_ptr() for t in keep])
    xd = x.to(dev)
    hi = xd.half()
    lo = (xd - hi.float()).half()
    errs = []
    for use_lo in (True, False):
        out = torch.full((B * S * C, ndir * H), float("nan"), dtype=torch.float16, device=dev)
        wpack = torch.empty(ndir * 512 * (2 * N + H), dtype=torch.float16, device=dev)
        bpack = torch.empty(ndir * 512, dtype=torch.float32, device=dev)
        rc = lib.vatss_tc_lstm(_p(hi), _p(lo) if use_lo else None, table, _p(out), mode, B, S, C, N, ndir, 0, _p(wpack),
                               _p(bpack), None)
        _lib.check(rc, "vatss_tc_lstm")
        torch.cuda.synchronize()
        got = out.float().cpu().reshape(B, S, C, ndir * H)
        errs.append(((got - ref).norm() / ref.norm()).item())
    print(f"lstm hi/lo mode={mode}: rel err precise {errs[0]:.3e}, plain fp16 {errs[1]:.3e}")
    assert errs[0] < 4e-4           # output rounding to fp16 (2^-11) dominates
    assert errs[0] < 0.7 * errs[1]
