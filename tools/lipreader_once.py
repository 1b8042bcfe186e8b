"""One lipreader forward per engine on synthetic frames (for ncu launch lists).  Usage: lipreader_once.py [B T engine]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_separation_b200 import Lipreading, extract_embeddings  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
engines = sys.argv[3:] or ["f32", "tensor"]
torch.manual_seed(0)
net = Lipreading(relu_type="swish", extract_feats=True).to("cuda:0")
vid = (torch.rand(B, T, 96, 96, device="cuda:0") * 255).round()
for eng in engines:
    net.set_engine(eng)
    emb = extract_embeddings(net, vid)
    torch.cuda.synchronize()
    print(eng, tuple(emb.shape), float(emb.abs().mean()))
