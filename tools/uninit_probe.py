"""Poison the workspace before a forward: any dependence of the outputs on unwritten workspace bytes shows up."""
import sys, torch
sys.path.insert(0, '/root/repo')
import bench
import speech_separation_b200 as V
from speech_separation_b200 import _lib
dev = torch.device('cuda:0')
mix, s1, s2, e1, e2 = (t.to(dev) for t in bench.make_batch(4, 64000, 1234))
torch.manual_seed(42)
net = V.DPTNAVWavEncDec(**bench.MODEL_KW).eval().to(dev)
for B in (1, 2):
    args = dict(mix=mix[:B], s1_embedding=e1[:B], s2_embedding=e2[:B])
    ref = net(**args)["s1_pred"].clone()
    ws = next(iter(net._workspaces.values()))
    n = ws.numel()
    print(f"B={B}: workspace {n/2**20:.1f} MiB")
    for pattern, label in ((0x00, "zeros"), (0xFF, "0xFF (NaN)"), (0x3C, "0x3C3C (fp16 ~1.06 / fp32 0.0115)")):
        ws.fill_(pattern)
        out = net(**args)["s1_pred"]
        d = float((out - ref).norm() / ref.norm()) if torch.isfinite(out).all() else float('nan')
        print(f"   poison {label}: finite={bool(torch.isfinite(out).all())} rel diff vs first run {d:.2e}")
    # locate: poison one 1/16 slice at a time with NaN over a zero background
    ws.fill_(0); base = net(**args)["s1_pred"].clone()
    step = (n + 15) // 16
    hits = []
    for k in range(16):
        ws.fill_(0); ws[k*step:(k+1)*step].fill_(0xFF)
        out = net(**args)["s1_pred"]
        if not torch.isfinite(out).all() or not torch.equal(out, base): hits.append(k)
    print("   slices (of 16) whose poisoning changes the output:", hits)
