"""Inferencer / MetricTracker: host logic on CPU, the real loop on the GPU.

Reference behaviour: src/trainer/inferencer.py:98-202, src/metrics/tracker.py:4-72.
"""
import json
import os
from pathlib import Path

import pytest
import torch

from speech_separation_b200.inference import Inferencer, MetricTracker

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "metric_tracker.json")


def test_metric_tracker_matches_reference_golden():
    cases = json.load(open(GOLDEN))["cases"]
    for c in cases:
        tr = MetricTracker(*c["keys"], writer=None)
        for k, v, n in c["updates"]:
            tr.update(k, v, n=n)
        for k in c["keys"]:
            assert tr.avg(k) == pytest.approx(c["avg"][k], rel=1e-12, abs=1e-12)
        res = tr.result()
        assert set(res) == set(c["result"])
        for k in c["keys"]:
            assert res[k] == pytest.approx(c["result"][k], rel=1e-12, abs=1e-12)
        tr.reset()
        assert tr.result() == c["after_reset"]
        tr.update("loss", 1.5)
        assert tr.result() == c["after_reset_then_loss_1.5"]
        assert list(tr.keys()) == c["keys_listed"]


def test_metric_tracker_rejects_unknown_key_and_accepts_tensors():
    tr = MetricTracker("a", "b")
    with pytest.raises(KeyError):
        tr.update("c", 1.0)
    tr.update("a", torch.tensor(2.0))          # 0-d host tensor, as SISNRiMetric returns in the reference
    tr.update("a", torch.tensor(4.0), n=3)
    assert tr.avg("a") == pytest.approx((2.0 + 12.0) / 4)
    assert tr.avg("b") == 0.0


class _FakeModel:
    """Host stand-in with the model's calling convention (the CUDA model is exercised in the gpu test)."""

    def eval(self):
        return self

    def __call__(self, mix, **batch):
        return {"s1_pred": 0.5 * mix, "s2_pred": -0.25 * mix}


class _MeanAbs:
    name = "mean_abs"

    def __call__(self, s1_pred, **batch):
        return float(s1_pred.abs().mean())


def _batches(n_batches, B, T, with_gt=True, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for b in range(n_batches):
        s1, s2 = torch.randn(B, T, generator=g) * 0.1, torch.randn(B, T, generator=g) * 0.1
        out.append({"mix": s1 + s2, "s1": s1 if with_gt else None, "s2": s2 if with_gt else None,
                    "s1_embedding": torch.randn(B, 512, 25 * T // 16000, generator=g),
                    "s2_embedding": torch.randn(B, 512, 25 * T // 16000, generator=g),
                    "audio_path": [f"/data/mix/utt_{b}_{i}.wav" for i in range(B)]})
    return out


def test_inferencer_host_logic_cpu(tmp_path):
    cfg = {"inferencer": {"device_tensors": ["mix", "s1", "s2"], "from_pretrained": None}}
    batches = _batches(3, 2, 800)
    expect = sum(float((0.5 * b["mix"]).abs().mean()) for b in batches) / 3
    inf = Inferencer(_FakeModel(), cfg, "cpu", {"val": batches}, tmp_path, metrics={"inference": [_MeanAbs()]},
                     batch_transforms={"inference": {"mix": lambda x: x}}, skip_model_load=True)
    logs = inf.run_inference()
    assert logs["val"]["mean_abs"] == pytest.approx(expect, rel=1e-6)
    files = sorted((tmp_path / "val").iterdir())
    assert [f.name for f in files] == sorted(f"utt_{b}_{i}.pth" for b in range(3) for i in range(2))
    rec = torch.load(tmp_path / "val" / "utt_1_1.pth")
    assert set(rec) == {"s1_pred", "s2_pred", "s1_true", "s2_true"}
    assert torch.equal(rec["s1_pred"], 0.5 * batches[1]["mix"][1])
    assert torch.equal(rec["s2_true"], batches[1]["s2"][1])
    # no ground truth: predictions only, no metric updates (inferencer.py:150-166)
    nogt = _batches(1, 2, 800, with_gt=False, seed=3)
    cfg2 = {"inferencer": {"device_tensors": ["mix"], "from_pretrained": None}}
    inf2 = Inferencer(_FakeModel(), cfg2, "cpu", {"test": nogt}, tmp_path, metrics={"inference": [_MeanAbs()]},
                      skip_model_load=True)
    assert inf2.run_inference()["test"] == {"mean_abs": 0.0}
    assert set(torch.load(tmp_path / "test" / "utt_0_0.pth")) == {"s1_pred", "s2_pred"}


def test_inferencer_requires_checkpoint_unless_skipped(tmp_path):
    with pytest.raises(AssertionError):
        Inferencer(_FakeModel(), {"inferencer": {"device_tensors": []}}, "cpu", {}, tmp_path)


@pytest.mark.gpu
def test_inferencer_gpu_matches_per_batch_calls(tmp_path):
    import speech_separation_b200 as V
    dev = torch.device("cuda:0")
    torch.manual_seed(42)
    kw = dict(num_features=128, video_emb_size=512, hidden_video=128, kernel_size_enc=7, hidden_dim=128, num_blocks=2,
              chunk_size=150, step_size=75, num_heads=4, dropout=0.1, bidir=True)
    net = V.DPTNAVWavEncDec(**kw).eval().to(dev)
    ckpt = tmp_path / "model_best.pth"
    torch.save({"state_dict": net.state_dict()}, ckpt)
    net2 = V.DPTNAVWavEncDec(**kw).eval().to(dev)      # different random init, replaced by the checkpoint
    cfg = {"inferencer": {"device_tensors": ["mix", "s1", "s2", "s1_embedding", "s2_embedding"],
                          "from_pretrained": str(ckpt)}}
    batches = _batches(3, 2, 32000, seed=11)
    mets = [V.SISNRMetric(name="SISNR"), V.SISNRiMetric(name="SISNRi")]
    inf = Inferencer(net2, cfg, dev, {"val": [dict(b) for b in batches]}, tmp_path / "out", metrics={"inference": mets},
                     skip_model_load=False)
    logs = inf.run_inference()["val"]
    # the reference's aggregation: one batch-level value per batch, plain mean over batches
    want = {"SISNR": 0.0, "SISNRi": 0.0}
    for b in batches:
        d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}
        out = net(**d)
        d.update(out)
        want["SISNR"] += mets[0](**d) / 3
        want["SISNRi"] += float(mets[1](**d)) / 3
        rec = torch.load(tmp_path / "out" / "val" / f"{Path(b['audio_path'][1]).stem}.pth")
        assert torch.equal(rec["s1_pred"], out["s1_pred"][1].cpu())
        assert torch.equal(rec["s2_pred"], out["s2_pred"][1].cpu())
        assert torch.equal(rec["s1_true"], b["s1"][1])
    assert logs["SISNR"] == pytest.approx(want["SISNR"], abs=1e-4)
    assert logs["SISNRi"] == pytest.approx(want["SISNRi"], abs=1e-4)
    assert len(list((tmp_path / "out" / "val").iterdir())) == 6


class _FreshBatches:
    """A dataloader stand-in that builds every batch dict on demand and keeps no reference to it (what a real
    torch DataLoader does): the previous batch is freed before the next one is moved to the device, so the caching
    allocator hands the same device addresses to consecutive batches."""

    def __init__(self, n, B, T, seed, scale):
        self.n, self.B, self.T, self.seed, self.scale = n, B, T, seed, scale

    def make(self, i):
        b = _batches(1, self.B, self.T, seed=self.seed + i)[0]
        b["s1"] = b["s1"] * self.scale[i]      # very different levels per batch: a stale summary is far off
        b["mix"] = b["s1"] + b["s2"]
        return b

    def __iter__(self):
        for i in range(self.n):
            yield self.make(i)


@pytest.mark.gpu
def test_shared_sisnr_summary_never_outlives_its_batch():
    """Round-1 advisor finding: the shared SI-SNR cache was keyed on device addresses, which the caching allocator
    recycles from batch to batch (and raw-pointer kernels leave `_version` at 0) -> stale metric values."""
    import speech_separation_b200 as V
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    kw = dict(num_features=128, video_emb_size=512, hidden_video=128, kernel_size_enc=7, hidden_dim=128, num_blocks=1,
              chunk_size=150, step_size=75, num_heads=4, dropout=0.1, bidir=True)
    net = V.DPTNAVWavEncDec(**kw).eval().to(dev)
    cfg = {"inferencer": {"device_tensors": ["mix", "s1", "s2", "s1_embedding", "s2_embedding"], "from_pretrained": None}}
    loader = _FreshBatches(5, 2, 16000, seed=100, scale=[1.0, 8.0, 0.1, 3.0, 0.5])
    mets = [V.SISNRMetric(name="SISNR"), V.SISNRiMetric(name="SISNRi")]
    inf = Inferencer(net, cfg, dev, {"val": loader}, None, metrics={"inference": mets}, skip_model_load=True)
    logs = inf.run_inference()["val"]
    want = {"SISNR": 0.0, "SISNRi": 0.0}
    per_batch = []
    for i in range(loader.n):
        d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in loader.make(i).items()}
        d.update(net(**d))
        per_batch.append(float(mets[1](**d)))
        want["SISNR"] += mets[0](**d) / loader.n
        want["SISNRi"] += per_batch[-1] / loader.n
    assert max(per_batch) - min(per_batch) > 1.0     # the batches really differ
    assert logs["SISNR"] == pytest.approx(want["SISNR"], abs=1e-4)
    assert logs["SISNRi"] == pytest.approx(want["SISNRi"], abs=1e-4)


def test_shared_sisnr_cache_is_identity_keyed():
    from speech_separation_b200.inference import _SharedSisnr
    import speech_separation_b200.inference as I

    calls = []
    orig = I.pit_sisnr_all
    I.pit_sisnr_all = lambda *ts: (None, None, calls.append(ts) or len(calls))
    try:
        sh = _SharedSisnr()
        a = {k: torch.zeros(2, 8) for k in ("s1_pred", "s2_pred", "s1", "s2", "mix")}
        assert sh.summary(a) == 1 and sh.summary(a) == 1            # same tensors: one pass
        b = {k: v.clone() for k, v in a.items()}
        assert sh.summary(b) == 2                                   # equal values, other objects: recomputed
        b["s1"].add_(1.0)
        assert sh.summary(b) == 3                                   # in-place change (version counter)
        sh.reset()
        assert sh.summary(b) == 4                                   # a new batch never hits
    finally:
        I.pit_sisnr_all = orig
