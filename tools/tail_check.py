"""Checksum of the separated waveforms of one forward (cfg-2 shape by default): run with VATSS_TAIL_STAGED=0 and =1 -
the staged tail kernel must reproduce the gather kernel bit for bit (same operands, same order of additions)."""
import hashlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import speech_separation_b200 as V
dev = torch.device("cuda:0")
for model, kw, B, T in (("dptn_av", bench.MODEL_KW, 32, 64000), ("dptn_av", bench.MODEL_KW, 3, 16000 + 37), ("dptn_av", bench.MODEL_KW, 1, 160000)):
    torch.manual_seed(42)
    net = V.DPTNAVWavEncDec(**kw).eval().to(dev)
    mix, s1, s2, e1, e2 = (t.to(dev) for t in bench.make_batch(B, T, 1234))
    out = net(mix=mix, s1_embedding=e1, s2_embedding=e2)
    torch.cuda.synchronize()
    h = hashlib.sha1(out["s1_pred"].cpu().numpy().tobytes() + out["s2_pred"].cpu().numpy().tobytes()).hexdigest()
    print(model, B, T, h, float(out["s1_pred"].abs().mean()))
