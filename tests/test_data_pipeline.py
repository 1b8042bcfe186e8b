"""Input pipeline: wav / npz formats, index layout and collate semantics of the reference
(src/datasets/ss_dataset.py:48-116, base_dataset.py:60-135, collate.py:4-46)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle.gen_collate_golden import items as golden_items  # noqa: E402  (seeded generator shared with the fixture)
from speech_separation_b200.data import SSDataset, collate_fn, load_object, make_dataloader  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "collate.json")


def test_collate_matches_reference_golden():
    for c in json.load(open(GOLDEN))["cases"]:
        batch = collate_fn(golden_items(*c["args"]))
        assert list(batch.keys()) == c["key_order"]
        for k, want in c["batch"].items():
            if want is None:
                assert batch[k] is None
            elif isinstance(want, dict):
                assert list(batch[k].shape) == want["shape"]
                assert float(batch[k].double().sum()) == pytest.approx(want["sum"], rel=1e-12, abs=1e-12)
            else:
                assert batch[k] == want


def _write_corpus(root, n=5, T=8000, E=512, Tv=12, with_gt=True):
    from scipy.io import wavfile
    rng = np.random.default_rng(0)
    split = root / "audio" / "val"
    for sub in ["mix"] + (["s1", "s2"] if with_gt else []):
        (split / sub).mkdir(parents=True)
    (root / "embedding").mkdir()
    truth = {}
    for i in range(n):
        a, b = f"spk{i}a", f"spk{i}b"
        s1 = (rng.standard_normal(T) * 3000).astype(np.int16)
        s2 = (rng.standard_normal(T) * 3000).astype(np.int16)
        mix = (s1.astype(np.int32) + s2).clip(-32768, 32767).astype(np.int16)
        wavfile.write(split / "mix" / f"{a}_{b}.wav", 16000, mix)
        if with_gt:
            wavfile.write(split / "s1" / f"{a}_{b}.wav", 16000, s1)
            wavfile.write(split / "s2" / f"{a}_{b}.wav", 16000, s2)
        for spk in (a, b):
            emb = rng.standard_normal((E, Tv)).astype(np.float32)
            np.savez(root / "embedding" / f"{spk}.npz", emb)
            truth[spk] = emb
        truth[f"{a}_{b}"] = (mix, s1, s2)
    return truth


def test_ssdataset_index_items_and_batches(tmp_path):
    truth = _write_corpus(tmp_path)
    ds = SSDataset(part="val", audio_dir=tmp_path / "audio", embedding_dir=tmp_path / "embedding",
                   video_dir=tmp_path / "mouth")
    assert len(ds) == 5 and not ds.contains_video and ds.contains_embedding
    index = json.load(open(tmp_path / "audio" / "val_index.json"))      # cached like the reference
    assert set(index[0]) == {"mix_wav_path", "s1_wav_path", "s2_wav_path", "s1_video_path", "s2_video_path",
                             "s1_embedding_path", "s2_embedding_path", "audio_len"}
    assert index[0]["audio_len"] == pytest.approx(0.5) and index[0]["s1_video_path"] is None
    item = ds[2]
    stem = os.path.basename(item["audio_path"])[:-4]
    mix, s1, s2 = truth[stem]
    assert item["mix"].shape == (1, 8000) and item["mix"].dtype == torch.float32
    np.testing.assert_allclose(item["mix"][0].numpy(), mix / 32768.0, atol=1e-7)
    np.testing.assert_allclose(item["s2"][0].numpy(), s2 / 32768.0, atol=1e-7)
    a, b = stem.split("_")
    assert item["s1_embedding"].shape == (1, 512, 12)
    np.testing.assert_array_equal(item["s1_embedding"][0].numpy(), truth[a])
    np.testing.assert_array_equal(item["s2_embedding"][0].numpy(), truth[b])
    assert item["s1_video"] is None and "mix_spectrogram" not in item
    dl = make_dataloader(ds, batch_size=2, num_workers=0, pin_memory=False)
    batches = list(dl)
    assert [b["mix"].shape[0] for b in batches] == [2, 2, 1]
    assert batches[0]["s1_embedding"].shape == (2, 512, 12) and batches[0]["s1_video"] is None
    assert isinstance(batches[0]["audio_path"], list) and len(batches[0]["audio_path"]) == 2
    # second construction reuses the cached index; `limit` truncates it
    assert len(SSDataset(part="val", audio_dir=tmp_path / "audio", embedding_dir=tmp_path / "embedding", limit=3)) == 3


def test_ssdataset_accepts_the_reference_config_keywords(tmp_path):
    """src/configs/datasets/ss_dataset.yaml passes `instance_transforms`; BaseDataset also takes `encoder`,
    `shuffle_index` and `limit` (base_dataset.py:23-54): same shuffle (python random, seed 42), same transform rules."""
    import random

    _write_corpus(tmp_path)
    kw = dict(part="val", audio_dir=tmp_path / "audio", embedding_dir=tmp_path / "embedding")
    plain = SSDataset(**kw)
    paths = [d["mix_wav_path"] for d in plain._index]
    want = list(paths)
    random.seed(42)
    random.shuffle(want)
    shuf = SSDataset(shuffle_index=True, limit=3, encoder=object(), **kw)
    assert [d["mix_wav_path"] for d in shuf._index] == want[:3] and shuf.encoder is None
    tf = {"mix": lambda x: 2.0 * x, "s1": lambda x: x + 1.0, "get_spectrogram": lambda x: x.abs()[..., :4] + 1.0}
    ds = SSDataset(instance_transforms=tf, **kw)
    a, b = plain[1], ds[1]
    assert torch.equal(b["mix"], 2.0 * a["mix"])                 # applied exactly once (single_key, then skipped)
    assert torch.equal(b["s1"], a["s1"] + 1.0) and torch.equal(b["s2"], a["s2"])
    # the spectrogram is taken from the un-augmented waveform the item loader read (base_dataset.py:113-116)
    assert torch.allclose(b["mix_spectrogram"], torch.log((a["mix"].abs()[..., :4] + 1.0).clamp(1e-5)))
    assert "s1_spectrogram" in b and "s2_spectrogram" in b


def test_ssdataset_without_ground_truth_and_object_formats(tmp_path):
    _write_corpus(tmp_path, n=2, with_gt=False)
    ds = SSDataset(part="val", audio_dir=tmp_path / "audio", embedding_dir=tmp_path / "embedding")
    item = ds[0]
    assert item["s1"] is None and item["s2"] is None
    batch = collate_fn([ds[0], ds[1]])
    assert batch["s1"] is None and batch["mix"].shape == (2, 8000)
    x = torch.arange(6.0).reshape(2, 3)
    np.save(tmp_path / "a.npy", x.numpy())
    torch.save(x, tmp_path / "a.pt")
    assert torch.equal(load_object(tmp_path / "a.npy"), x[None]) and torch.equal(load_object(tmp_path / "a.pt"), x[None])
    with pytest.raises(ValueError):
        load_object(tmp_path / "a.txt")


@pytest.mark.gpu
def test_inferencer_over_ssdataset_end_to_end(tmp_path):
    """wav + npz on disk -> SSDataset -> pinned DataLoader -> Inferencer -> .pth per utterance + SI-SNRi."""
    import speech_separation_b200 as V
    truth = _write_corpus(tmp_path, n=4, T=16000, Tv=25)
    ds = SSDataset(part="val", audio_dir=tmp_path / "audio", embedding_dir=tmp_path / "embedding")
    dl = make_dataloader(ds, batch_size=2, num_workers=0, pin_memory=True)
    dev = torch.device("cuda:0")
    torch.manual_seed(42)
    net = V.DPTNAVWavEncDec(num_features=128, video_emb_size=512, hidden_video=128, kernel_size_enc=7, hidden_dim=128,
                            num_blocks=1, chunk_size=150, step_size=75, num_heads=4).eval().to(dev)
    cfg = {"inferencer": {"device_tensors": ["mix", "s1", "s2", "s1_embedding", "s2_embedding"], "from_pretrained": None}}
    mets = {"inference": [V.SISNRiMetric(name="SISNRi")]}
    logs = V.Inferencer(net, cfg, dev, {"val": dl}, tmp_path / "out", metrics=mets, skip_model_load=True).run_inference()
    assert np.isfinite(logs["val"]["SISNRi"])
    files = sorted(p.name for p in (tmp_path / "out" / "val").iterdir())
    assert files == sorted(f"{k}.pth" for k in truth if "_" in k)
    rec = torch.load(tmp_path / "out" / "val" / files[0])
    assert rec["s1_pred"].shape == (16000,) and torch.isfinite(rec["s1_pred"]).all()
    np.testing.assert_allclose(rec["s1_true"].numpy(), truth[files[0][:-4]][1] / 32768.0, atol=1e-7)
