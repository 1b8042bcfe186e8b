"""Many forwards of the same input at several batch sizes: every output must be bit-identical to the first."""
import sys, torch
sys.path.insert(0, '/root/repo')
import bench
import speech_separation_b200 as V
dev = torch.device('cuda:0')
torch.manual_seed(42)
net = V.DPTNAVWavEncDec(**bench.MODEL_KW).eval().to(dev)
mix, s1, s2, e1, e2 = (t.to(dev) for t in bench.make_batch(32, 64000, 1234))
total_bad = 0
for B, reps in ((32, 40), (1, 60), (3, 60), (7, 40)):
    ref = None; bad = 0
    for _ in range(reps):
        out = net(mix=mix[:B], s1_embedding=e1[:B], s2_embedding=e2[:B])
        cur = torch.stack([out["s1_pred"], out["s2_pred"]])
        if ref is None: ref = cur.clone()
        elif not torch.equal(ref, cur): bad += 1
    print(f"B={B}: {bad}/{reps - 1} repeats differ")
    total_bad += bad
print("TOTAL", total_bad)
