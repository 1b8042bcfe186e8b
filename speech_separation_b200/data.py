"""Input side of the evaluation loop: mixture / source wav files and lip-embedding arrays -> batches.

Mirrors the reference's data formats and batch keys:
`SSDataset` = src/datasets/ss_dataset.py:12-116 (directory layout, `<part>_index.json`, index entry keys) with the
item loader of src/datasets/base_dataset.py:60-135,143-149,188-205, and `collate_fn` = src/datasets/collate.py:4-46
(same key lists, `None` passthrough, `torch.cat` along dim 0, `audio_path` as a list).

What is dropped on purpose: the per-item MelSpectrogram / log (`base_dataset.py:115-123`) that every reference batch
pays and no waveform model reads (`mix_spectrogram`, `s1_spectrogram`, `s2_spectrogram` are only consumed by the
spectrogram models, which are out of scope).  Pass `spectrogram_fn` to get those keys back.  Audio decoding uses
torchaudio when it is installed (as the reference does) and scipy's wav reader otherwise.
"""
import json
import random
import os
import wave
from pathlib import Path

import numpy as np
import torch

TENSOR_KEYS = ["mix_spectrogram", "complex_spectrogram", "s1_spectrogram", "s2_spectrogram", "s1_video", "s2_video",
               "s1_embedding", "s2_embedding", "mix", "s1", "s2"]
LIST_KEYS = ["audio_path"]


def collate_fn(dataset_items):
    """list of items -> batch dict (src/datasets/collate.py:4-46)."""
    batch = {}
    for key in TENSOR_KEYS + LIST_KEYS:
        if key not in dataset_items[0]:
            continue
        if dataset_items[0][key] is None:
            batch[key] = None
            continue
        values = [item[key] for item in dataset_items]
        batch[key] = torch.cat(values, dim=0) if key in TENSOR_KEYS else values
    return batch


def _wav_info(path):
    """(frames, sample_rate) from the header."""
    try:
        with wave.open(str(path), "rb") as w:
            return w.getnframes(), w.getframerate()
    except wave.Error:   # float / extensible wav: fall back to decoding
        from scipy.io import wavfile
        sr, data = wavfile.read(str(path))
        return data.shape[0], sr


def load_audio(path, target_sr=16000):
    """First channel as a (1, T) float32 tensor at `target_sr` (src/datasets/base_dataset.py:143-149)."""
    try:
        import torchaudio
        audio, sr = torchaudio.load(str(path))
        audio = audio[0:1, :]
        if sr != target_sr:
            audio = torchaudio.functional.resample(audio, sr, target_sr)
        return audio
    except ImportError:
        from scipy.io import wavfile
        sr, data = wavfile.read(str(path))
        if sr != target_sr:
            raise RuntimeError(f"{path}: sample rate {sr} != {target_sr} and torchaudio (resampling) is not installed")
        if data.ndim > 1:
            data = data[:, 0]
        if np.issubdtype(data.dtype, np.integer):
            data = data.astype(np.float32) / float(np.iinfo(data.dtype).max + 1)
        return torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32))[None, :]


def load_object(path):
    """.npy / .npz (first array) / .pt / .pth -> tensor with a leading batch axis (base_dataset.py:188-205)."""
    path = str(path)
    if path.endswith(".npy"):
        obj = torch.from_numpy(np.load(path))
    elif path.endswith(".npz"):
        with np.load(path) as data:
            obj = torch.from_numpy(data[next(iter(data))])
    elif path.endswith(".pt") or path.endswith(".pth"):
        obj = torch.load(path)
    else:
        raise ValueError(f"unsupported object file: {path}")
    return obj.unsqueeze(0)


def load_video(path):
    data = np.load(str(path))
    return torch.FloatTensor(data["data"]).unsqueeze(0)


class SSDataset(torch.utils.data.Dataset):
    """`<audio_dir>/<part>/{mix,s1,s2}/<id1>_<id2>.wav`, `<embedding_dir>/<id>.npz`, `<video_dir>/<id>.npz`.

    Accepts every keyword the reference's dataset configs pass (src/configs/datasets/ss_dataset.yaml ->
    src/datasets/ss_dataset.py:12-46 -> base_dataset.py:23-54): `limit`, `target_sr`, `encoder`, `shuffle_index`
    (python `random` with seed 42, then the limit, base_dataset.py:295-314) and `instance_transforms`, applied per key
    like `preprocess_data` (base_dataset.py:207-233: the "mix" transform first, then every other key except the
    special "get_spectrogram").  Deliberate divergences, all on keys no waveform model reads:
    * the per-item log-MelSpectrogram (`mix_spectrogram`, base_dataset.py:115-123) is only computed when an
      `instance_transforms["get_spectrogram"]` (or `spectrogram_fn`) is given - the reference fails without one;
    * mouth-crop videos are loaded only with `load_videos=True` (default: when the model needs them, i.e. never on
      the embedding path); the reference loads them whenever the files exist;
    * `encoder` (an STFT front-end of the spectrogram models) is accepted and ignored, as in the reference
      (`self.encoder = None`, base_dataset.py:49).
    """

    def __init__(self, part="train", audio_dir=None, video_dir=None, embedding_dir=None, target_sr=16000, limit=None,
                 load_videos=False, spectrogram_fn=None, root=None, encoder=None, shuffle_index=False,
                 instance_transforms=None):
        root = Path(root) if root is not None else Path.cwd() / "data"
        self._audio_dir = Path(audio_dir) if audio_dir is not None else root / "audio"
        self._video_dir = Path(video_dir) if video_dir is not None else root / "mouth"
        self._embedding_dir = Path(embedding_dir) if embedding_dir is not None else root / "embedding"
        self.contains_video = self._video_dir.exists()
        self.contains_embedding = self._embedding_dir.exists()
        self.target_sr = target_sr
        self.load_videos = load_videos
        self.encoder = None
        self.instance_transforms = instance_transforms
        if spectrogram_fn is None and instance_transforms is not None and "get_spectrogram" in instance_transforms:
            mel = instance_transforms["get_spectrogram"]
            spectrogram_fn = lambda audio: torch.log(mel(audio).clamp(1e-5))  # noqa: E731  (base_dataset.py:161)
        self.spectrogram_fn = spectrogram_fn
        index = self._get_or_load_index("custom" if part is None else part)
        self._index = self._shuffle_and_limit_index(index, limit, shuffle_index)

    @staticmethod
    def _shuffle_and_limit_index(index, limit, shuffle_index):
        if shuffle_index:
            random.seed(42)
            random.shuffle(index)
        return index if limit is None else index[:limit]

    def preprocess_data(self, instance_data, special_keys=("get_spectrogram",), single_key=None):
        if self.instance_transforms is None:
            return instance_data
        if single_key is not None:
            if single_key in self.instance_transforms:
                instance_data[single_key] = self.instance_transforms[single_key](instance_data[single_key])
            return instance_data
        for name in self.instance_transforms.keys():
            if name in special_keys:
                continue
            instance_data[name] = self.instance_transforms[name](instance_data[name])
        return instance_data

    # ---- index (ss_dataset.py:48-116) ----------------------------------------------------------------
    def _get_or_load_index(self, part):
        index_path = self._audio_dir / f"{part}_index.json"
        if index_path.exists():
            with index_path.open() as f:
                return json.load(f)
        index = self._create_index(part)
        with index_path.open("w") as f:
            json.dump(index, f, indent=2)
        return index

    def _create_index(self, part):
        split_dir = self._audio_dir if part == "custom" else self._audio_dir / part
        mix_dir = split_dir / "mix"
        has_gt = (split_dir / "s1").exists()
        index = []
        for wavname in sorted(os.listdir(mix_dir)):
            if not wavname.endswith(".wav"):
                continue
            id1, id2 = wavname.replace(".wav", "").split("_")
            resolve = lambda p: str(Path(p).absolute().resolve())  # noqa: E731
            frames, sr = _wav_info(mix_dir / wavname)
            index.append({
                "mix_wav_path": resolve(mix_dir / wavname),
                "s1_wav_path": resolve(split_dir / "s1" / wavname) if has_gt else None,
                "s2_wav_path": resolve(split_dir / "s2" / wavname) if has_gt else None,
                "s1_video_path": resolve(self._video_dir / f"{id1}.npz") if self.contains_video else None,
                "s2_video_path": resolve(self._video_dir / f"{id2}.npz") if self.contains_video else None,
                "s1_embedding_path": resolve(self._embedding_dir / f"{id1}.npz") if self.contains_embedding else None,
                "s2_embedding_path": resolve(self._embedding_dir / f"{id2}.npz") if self.contains_embedding else None,
                "audio_len": frames / sr,
            })
        return index

    # ---- items (base_dataset.py:60-135) --------------------------------------------------------------
    def __len__(self):
        return len(self._index)

    def __getitem__(self, ind):
        d = self._index[ind]
        item = {"mix": load_audio(d["mix_wav_path"], self.target_sr), "s1": None, "s2": None, "s1_video": None,
                "s2_video": None, "s1_embedding": None, "s2_embedding": None, "audio_path": d["mix_wav_path"]}
        if d["s1_wav_path"] is not None:
            item["s1"] = load_audio(d["s1_wav_path"], self.target_sr)
            item["s2"] = load_audio(d["s2_wav_path"], self.target_sr)
        if self.load_videos and d["s1_video_path"] is not None:
            item["s1_video"] = load_video(d["s1_video_path"])
            item["s2_video"] = load_video(d["s2_video_path"])
        if d["s1_embedding_path"] is not None:
            item["s1_embedding"] = load_object(d["s1_embedding_path"])
            item["s2_embedding"] = load_object(d["s2_embedding_path"])
        raw_mix = item["mix"]
        item = self.preprocess_data(item, single_key="mix")           # wav augmentation of the mixture first ...
        if self.spectrogram_fn is not None:
            item["mix_spectrogram"] = self.spectrogram_fn(raw_mix)    # ... the spectrogram still sees the loaded audio
            if item["s1"] is not None:
                item["s1_spectrogram"] = self.spectrogram_fn(item["s1"])
                item["s2_spectrogram"] = self.spectrogram_fn(item["s2"])
        return self.preprocess_data(item, special_keys=("get_spectrogram", "mix"))


def make_dataloader(dataset, batch_size, num_workers=2, pin_memory=True, drop_last=False):
    """The reference's evaluation dataloader (src/configs/dataloader/example.yaml) with pinned staging, so the
    Inferencer's host->device copies are asynchronous."""
    return torch.utils.data.DataLoader(dataset, batch_size=batch_size, shuffle=False, num_workers=num_workers,
                                       collate_fn=collate_fn, pin_memory=pin_memory, drop_last=drop_last)
