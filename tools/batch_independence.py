"""Batch independence at cfg-2: utterance i of a batch of 32 must equal the same utterance run alone."""
import sys
import torch
sys.path.insert(0, '/root/repo')
import bench
import speech_separation_b200 as V
dev = torch.device('cuda:0')
torch.manual_seed(42)
net = V.DPTNAVWavEncDec(**bench.MODEL_KW).eval().to(dev)
mix, s1, s2, e1, e2 = (t.to(dev) for t in bench.make_batch(32, 64000, 1234))
full = net(mix=mix, s1_embedding=e1, s2_embedding=e2)["s1_pred"].clone()
again = net(mix=mix, s1_embedding=e1, s2_embedding=e2)["s1_pred"].clone()
print("deterministic:", torch.equal(full, again))
for n in (1, 2, 4, 8, 16):
    sub = net(mix=mix[:n], s1_embedding=e1[:n], s2_embedding=e2[:n])["s1_pred"]
    d = ((sub - full[:n]).norm(dim=1) / full[:n].norm(dim=1))
    print(f"batch {n}: rel-L2 per utterance vs batch of 32:", [f"{x:.1e}" for x in d.tolist()][:4])
