"""Two forwards of the cfg-2 workload (batch 32 x 4 s DPTN-AV) for ncu: the first warms up, profile the second.

    python tools/profile_forward.py [--batch 32] [--seconds 4] [--forwards 2]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import speech_separation_b200 as V  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--seconds", type=float, default=4.0)
ap.add_argument("--forwards", type=int, default=2)
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(42)
net = V.DPTNAVWavEncDec(**bench.MODEL_KW).eval().to(dev)
mix, s1, s2, e1, e2 = (t.to(dev) for t in bench.make_batch(args.batch, int(args.seconds * 16000), 1234))
for _ in range(args.forwards):
    out = net(mix=mix, s1_embedding=e1, s2_embedding=e2)
    rows, _, summary = V.pit_sisnr_all(out["s1_pred"], out["s2_pred"], s1, s2, mix)
torch.cuda.synchronize()
print("ok", float(summary[4]))
