"""Time the attention core alone at the production shapes (CUDA events, L2 flushed between launches), kernel v1
(round 1, P through shared memory) against v2 (P in TMEM), and check both against torch on a subsample.

    python tools/attn_bench.py [--reps 5] [--shapes cfg2|all]
"""
import argparse
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speech_separation_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--shapes", default="cfg2")
ap.add_argument("--versions", default="1,3")
args = ap.parse_args()
lib = _lib.load()
dev = torch.device("cuda:0")


def P(t):
    return ctypes.c_void_p(t.data_ptr())


SHAPES = [("intra 4s", 0, 32, 283, 150, 128), ("inter 4s", 1, 32, 283, 150, 128)]
if args.shapes == "all":
    SHAPES += [("intra 10s", 0, 16, 710, 150, 128), ("inter 10s", 1, 16, 710, 150, 128),
               ("intra 4s N64", 0, 32, 283, 150, 64), ("inter 4s N64", 1, 32, 283, 150, 64)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, mode, B, S, C, N in SHAPES:
    heads, hd = 4, N // 4
    torch.manual_seed(0)
    qkv = torch.randn(B * S * C, 3 * N, device=dev)
    qkv[:, :N] *= 1.4426950408889634 / hd ** 0.5
    qkv = qkv.half()
    outs = {}
    for ver in [int(v) for v in args.versions.split(",")]:
        lib.vatss_debug_attention_version(ver)
        out = torch.full((B * S * C, N), float("nan"), dtype=torch.float16, device=dev)
        ms = []
        for it in range(args.reps + 2):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(lib.vatss_tc_attention(P(qkv), P(out), mode, B, S, C, N, heads, 0, None), "attn")
            b.record()
            torch.cuda.synchronize()
            if it >= 2:
                ms.append(a.elapsed_time(b))
        outs[ver] = out
        seqs, L = (B * S, C) if mode == 0 else (B * C, S)
        exps = seqs * heads * L * L
        # reference on the first utterance
        x = qkv[: S * C].float().reshape(S, C, 3 * N)
        seq = x if mode == 0 else x.permute(1, 0, 2)
        q, k, v = (seq[..., i * N:(i + 1) * N].reshape(seq.shape[0], seq.shape[1], heads, hd).transpose(1, 2) for i in range(3))
        ref = (torch.softmax((q @ k.transpose(-1, -2)) * 0.6931471805599453, dim=-1) @ v).transpose(1, 2).reshape(seq.shape[0], seq.shape[1], N)
        ref = ref if mode == 0 else ref.permute(1, 0, 2)
        got = out[: S * C].float().reshape(S, C, N)
        err = ((got - ref).norm() / ref.norm()).item()
        best = min(ms)
        print(f"{name:14s} v{ver}: {best:.3f} ms (median {sorted(ms)[len(ms) // 2]:.3f}) | {exps / best / 1e6:.0f} G true exp/s "
              f"= {exps / (best * 1e-3) / (148 * 16 * 1.9e9):.2f} of the MUFU floor @1.9 GHz | rel err vs torch {err:.2e}", flush=True)
    if len(outs) >= 2:
        vs = sorted(outs)
        d = (outs[vs[0]].float() - outs[vs[-1]].float()).abs().max().item()
        print(f"{name:14s} max |v{vs[0]} - v{vs[-1]}| = {d:.3e}")
lib.vatss_debug_attention_version(3)
