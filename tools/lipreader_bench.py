"""Lipreader front end on the GPU: parity of both engines against the CPU oracle on a small clip and device time per
batch of frames (CUDA events, L2 flushed between repetitions).  Usage: python tools/lipreader_bench.py [B T reps]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import lipreader_oracle as LO  # noqa: E402  (checker only)
from speech_separation_b200 import Lipreading  # noqa: E402


def flops_per_frame(hc=88, wc=88):
    def od(n, k, s, p):
        return (n + 2 * p - k) // s + 1
    h1, w1 = od(hc, 7, 2, 3), od(wc, 7, 2, 3)
    total = 2 * h1 * w1 * 64 * 245
    h, w, inpl = od(h1, 3, 2, 1), od(w1, 3, 2, 1), 64
    for li, pl in enumerate((64, 128, 256, 512)):
        for b in range(2):
            s = 2 if (li > 0 and b == 0) else 1
            ho, wo = od(h, 3, s, 1), od(w, 3, s, 1)
            total += 2 * ho * wo * pl * 9 * inpl + 2 * ho * wo * pl * 9 * pl
            if s == 2:
                total += 2 * ho * wo * pl * inpl
            h, w, inpl = ho, wo, pl
    return total


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    dev = torch.device("cuda:0")
    relu_type = "swish"
    sd = LO.make_state_dict(relu_type)
    net = Lipreading(relu_type=relu_type, extract_feats=True)
    net.load_state_dict(sd)
    net = net.to(dev)
    out = {"flops_per_frame": flops_per_frame()}
    # parity on a small clip
    frames = LO.make_frames(2, 9, seed=1)
    x = torch.from_numpy(LO.preprocess(frames).astype(np.float32))[:, None]
    with torch.no_grad():
        ref = LO.forward(sd, x, relu_type).numpy().astype(np.float64)
    for eng in ("f32", "tensor"):
        y = net.set_engine(eng)(x.to(dev), lengths=[9, 9]).cpu().numpy()
        out[f"rel_l2_{eng}"] = float(np.linalg.norm(y - ref) / np.linalg.norm(ref))
    print(json.dumps(out), flush=True)
    # timing
    vid = torch.from_numpy(LO.make_frames(1, T, seed=2)).to(dev).repeat(B, 1, 1, 1).contiguous()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    from speech_separation_b200 import extract_embeddings
    for eng in ("f32", "tensor"):
        net.set_engine(eng)
        for _ in range(2):
            extract_embeddings(net, vid)
        torch.cuda.synchronize()
        ms = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            extract_embeddings(net, vid)
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        m = float(np.median(ms))
        out[f"ms_{eng}"] = m
        out[f"frames_per_s_{eng}"] = B * T / m * 1e3
        out[f"tflops_{eng}"] = B * T * out["flops_per_frame"] / m / 1e9
    out["frames"] = B * T
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
