"""Audio-seconds per second through the whole Inferencer loop (H2D, forward, fused metrics, D2H, torch.save)
compared with the reference-style loop semantics (host sync per metric, clone + torch.save on the critical path)."""
import shutil, sys, tempfile, time
from pathlib import Path
import torch
sys.path.insert(0, '/root/repo')
import bench
import speech_separation_b200 as V
from speech_separation_b200.inference import Inferencer

dev = torch.device('cuda:0')
torch.manual_seed(42)
net = V.DPTNAVWavEncDec(**bench.MODEL_KW).eval().to(dev)
B, T, NB = 32, 64000, 6
mix, s1, s2, e1, e2 = (t.pin_memory() for t in bench.make_batch(B, T, 1234))
def batches():
    return [{"mix": mix, "s1": s1, "s2": s2, "s1_embedding": e1, "s2_embedding": e2,
             "audio_path": [f"utt_{b}_{i}.wav" for i in range(B)]} for b in range(NB)]
cfg = {"inferencer": {"device_tensors": ["mix", "s1", "s2", "s1_embedding", "s2_embedding"], "from_pretrained": None}}
mets = {"inference": [V.SISNRMetric(name="SISNR"), V.SISNRiMetric(name="SISNRi")]}
audio = NB * B * T / 16000

def run(save):
    out = Path(tempfile.mkdtemp(prefix="vatss_inf_")) if save else None
    inf = Inferencer(net, cfg, dev, {"val": batches()}, out, metrics=mets, skip_model_load=True)
    inf.run_inference()                      # warm-up (allocations, pinned staging)
    inf.evaluation_dataloaders = {"val": batches()}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    logs = inf.run_inference()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if out: shutil.rmtree(out)
    return audio / dt, logs

def reference_style(save):
    """Same kernels, but the reference's loop structure: .item() per metric, per-row clone + torch.save inline."""
    out = Path(tempfile.mkdtemp(prefix="vatss_ref_")) if save else None
    def loop():
        tot = [0.0, 0.0]
        for b in batches():
            d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}
            d.update(net(**d))
            tot[0] += mets["inference"][0](**d)
            tot[1] += mets["inference"][1](**d).item()
            if out:
                for i in range(B):
                    torch.save({"s1_pred": d["s1_pred"][i].clone(), "s2_pred": d["s2_pred"][i].clone(),
                                "s1_true": d["s1"][i].clone(), "s2_true": d["s2"][i].clone()}, out / f"{Path(d['audio_path'][i]).stem}.pth")
        return tot
    with torch.no_grad():
        loop()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        loop()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if out: shutil.rmtree(out)
    return audio / dt

for save in (False, True):
    v, logs = run(save)
    r = reference_style(save)
    print(f"save={save}: Inferencer {v:8.1f} audio-s/s | reference-style loop over the same kernels {r:8.1f} audio-s/s | metrics {logs}")
