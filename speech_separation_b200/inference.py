"""Evaluation loop around the separation hot path: `Inferencer` and `MetricTracker`.

Drop-in for src/trainer/inferencer.py:9-202 and src/metrics/tracker.py:4-72 (the
immediate caller of the forward pass).  Same constructor arguments, methods, on-disk format
(`<save_path>/<part>/<stem>.pth` holding {"s1_pred","s2_pred"[,"s1_true","s2_true"]}) and metric
aggregation (every metric value is a *batch-level* number and every batch counts once).

What changes is where the host waits.  The reference's loop stalls the GPU four to six times per batch
(`.item()` inside every metric call, src/metrics/base_metric.py:53-56) and then clones and `torch.save`s every
row on the critical path.  Here

* metric values stay on the device: metrics exposing `device_value(**batch)` return a 0-d CUDA tensor, the
  tracker accumulates them with device adds and reads them back once, in `result()` / `avg()`;
* SI-SNR and SI-SNRi share one pass over the waveforms (a per-batch cache of the `vatss_pit_sisnr` result);
* predictions leave through pinned staging buffers on a copy stream and are written by a background thread, so
  file I/O overlaps the next batch's forward.
"""
import queue
import threading
from pathlib import Path

import torch

from .loss import pit_sisnr_all


class MetricTracker:
    """Aggregates batch-level metric values (src/metrics/tracker.py:4-72): total, counts, average per key."""

    def __init__(self, *keys, writer=None):
        self.writer = writer
        self._keys = list(keys)
        self.reset()

    def reset(self):
        self._total = {k: 0.0 for k in self._keys}
        self._counts = {k: 0 for k in self._keys}
        self._pending = {k: None for k in self._keys}   # device-side partial totals (no host sync on update)

    def update(self, key, value, n=1):
        if key not in self._total:
            raise KeyError(key)
        if isinstance(value, torch.Tensor) and value.is_cuda:
            v = value.detach().to(torch.float64).reshape(()) * n
            self._pending[key] = v if self._pending[key] is None else self._pending[key] + v
        else:
            self._total[key] += float(value) * n
        self._counts[key] += n

    def _flush(self):
        live = [k for k in self._keys if self._pending[k] is not None]
        if live:
            host = torch.stack([self._pending[k] for k in live]).cpu()   # the one device -> host read
            for k, v in zip(live, host.tolist()):
                self._total[k] += v
                self._pending[k] = None

    def avg(self, key):
        self._flush()
        return self._total[key] / self._counts[key] if self._counts[key] else 0.0

    def result(self):
        self._flush()
        return {k: (self._total[k] / self._counts[k] if self._counts[k] else 0.0) for k in self._keys}

    def keys(self):
        return list(self._keys)


class _SharedSisnr:
    """One `vatss_pit_sisnr` pass per batch, shared by every SI-SNR-family metric of the loop.

    The cache lives for ONE batch: `Inferencer.process_batch` calls `reset()` before the metrics of a batch are
    evaluated, and a hit requires the very same tensor objects (identity, not addresses: the caching allocator hands
    the addresses of a freed batch to the next one, and tensors written by raw-pointer kernels keep `_version` 0).
    The cache holds references to its key tensors, so their ids cannot be recycled while it is valid.
    """

    def __init__(self):
        self.reset()

    def reset(self):
        self._key = None
        self._val = None

    def summary(self, batch):
        ts = (batch["s1_pred"], batch["s2_pred"], batch["s1"], batch["s2"], batch.get("mix"))
        hit = self._key is not None and all(a is b for a, b in zip(ts, self._key[0])) and \
            self._key[1] == tuple(t._version if t is not None else -1 for t in ts)
        if not hit:
            self._val = pit_sisnr_all(*ts)[2]
            self._key = (ts, tuple(t._version if t is not None else -1 for t in ts))
        return self._val


class _AsyncWriter:
    """Pinned staging + background `torch.save`: the compute stream never waits for the file system.

    On-disk divergence from the reference (src/trainer/inferencer.py:146-166): the saved `.pth` dicts hold CPU
    tensors (clones of the pinned staging rows); the reference saves clones of the device tensors.  Same keys, shapes,
    dtype and values; `torch.load(..., map_location=...)` reads both."""

    def __init__(self, device, depth=3):
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.copy_stream = torch.cuda.Stream(self.device) if self.cuda else None
        self.free = queue.Queue()
        for _ in range(depth):
            self.free.put({})
        self.jobs = queue.Queue()
        self.error = None
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _stage(self, slot, name, t):
        buf = slot.get(name)
        if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
            buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=self.cuda)
            slot[name] = buf
        buf.copy_(t, non_blocking=True)
        return buf

    def submit(self, tensors, paths):
        """tensors: name -> (B, T) tensor (device or host); paths: one file per row."""
        if self.error is not None:
            raise self.error
        slot = self.free.get()   # back-pressure: at most `depth` batches in flight
        event = None
        if self.cuda:
            self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.copy_stream):
                staged = {k: self._stage(slot, k, v) for k, v in tensors.items()}
                event = torch.cuda.Event()
                event.record(self.copy_stream)
            for v in tensors.values():   # keep the sources alive until the copy stream is done with them
                if v.is_cuda:
                    v.record_stream(self.copy_stream)
        else:
            staged = {k: self._stage(slot, k, v) for k, v in tensors.items()}
        self.jobs.put((slot, staged, event, list(paths)))

    def _run(self):
        while True:
            job = self.jobs.get()
            if job is None:
                return
            slot, staged, event, paths = job
            try:
                if event is not None:
                    event.synchronize()
                for i, path in enumerate(paths):
                    torch.save({k: v[i].clone() for k, v in staged.items()}, path)
            except Exception as e:  # surfaced on the next submit / drain
                self.error = e
            finally:
                self.free.put(slot)
                self.jobs.task_done()

    def drain(self):
        self.jobs.join()
        if self.error is not None:
            raise self.error

    def close(self):
        """Always stops the worker; a stored writer error is raised after the thread has been joined."""
        try:
            self.jobs.join()
        finally:
            self.jobs.put(None)
            self.thread.join()
        if self.error is not None:
            raise self.error


def _cfg_get(cfg, name, default=None):
    if cfg is None:
        return default
    if isinstance(cfg, dict):
        return cfg.get(name, default)
    getter = getattr(cfg, "get", None)
    if getter is not None:
        try:
            return getter(name, default)
        except TypeError:
            pass
    return getattr(cfg, name, default)


class Inferencer:
    """Runs the model over the evaluation dataloaders, aggregates metrics, saves the separated waveforms.

    Mirrors src/trainer/inferencer.py:9-202 (constructor, `run_inference`, `process_batch`, `_inference_part`) and
    the two BaseTrainer helpers it uses (`move_batch_to_device`, `transform_batch`, `_from_pretrained`:
    src/trainer/base_trainer.py:343-381,539-560).
    """

    def __init__(self, model, config, device, dataloaders, save_path, metrics=None, batch_transforms=None,
                 skip_model_load=False):
        self.config = config
        self.cfg_trainer = _cfg_get(config, "inferencer")
        assert skip_model_load or _cfg_get(self.cfg_trainer, "from_pretrained") is not None, \
            "Provide checkpoint or set skip_model_load=True"
        self.device = device
        self.model = model
        self.batch_transforms = batch_transforms
        self.evaluation_dataloaders = {k: v for k, v in dataloaders.items()}
        self.save_path = Path(save_path) if save_path is not None else None
        self.metrics = metrics
        if self.metrics is not None:
            self.evaluation_metrics = MetricTracker(*[m.name for m in self.metrics["inference"]], writer=None)
        else:
            self.evaluation_metrics = None
        self.is_train = False
        self._shared = _SharedSisnr()
        self._writer = None
        if not skip_model_load:
            self._from_pretrained(_cfg_get(self.cfg_trainer, "from_pretrained"))

    # ---- BaseTrainer helpers -------------------------------------------------------------------------
    def _from_pretrained(self, pretrained_path):
        pretrained_path = str(pretrained_path)
        print(f"Loading model weights from: {pretrained_path} ...")
        checkpoint = torch.load(pretrained_path, map_location=self.device)
        if isinstance(checkpoint, dict) and checkpoint.get("state_dict") is not None:
            self.model.load_state_dict(checkpoint["state_dict"])
        else:
            self.model.load_state_dict(checkpoint)

    def move_batch_to_device(self, batch):
        for name in _cfg_get(self.cfg_trainer, "device_tensors", []):
            batch[name] = batch[name].to(self.device, non_blocking=True)
        return batch

    def transform_batch(self, batch):
        if self.batch_transforms is None:
            return batch
        transforms = self.batch_transforms.get("train" if self.is_train else "inference")
        if transforms is not None:
            for name in transforms.keys():
                batch[name] = transforms[name](batch[name])
        return batch

    # ---- the loop ------------------------------------------------------------------------------------
    def run_inference(self):
        part_logs = {}
        for part, dataloader in self.evaluation_dataloaders.items():
            part_logs[part] = self._inference_part(part, dataloader)
        return part_logs

    def _metric_value(self, met, batch):
        fused = getattr(met, "device_value", None)
        if fused is not None:
            return fused(shared=self._shared, **batch)   # 0-d device tensor, no host sync
        return met(**batch)

    def process_batch(self, batch_idx, batch, metrics, part):
        self._shared.reset()   # the shared SI-SNR summary never outlives its batch
        batch = self.move_batch_to_device(batch)
        batch = self.transform_batch(batch)
        outputs = self.model(**batch)
        batch.update(outputs)

        has_gt = batch.get("s1") is not None
        if has_gt and metrics is not None:
            for met in self.metrics["inference"]:
                metrics.update(met.name, self._metric_value(met, batch))
        # (as in the reference, a batch with ground truth but no tracker is not written)
        if self.save_path is not None and (not has_gt or metrics is not None):
            tensors = {"s1_pred": batch["s1_pred"], "s2_pred": batch["s2_pred"]}
            if has_gt:
                tensors["s1_true"] = batch["s1"]
                tensors["s2_true"] = batch["s2"]
            paths = [self.save_path / part / f"{Path(p).stem}.pth" for p in batch["audio_path"]]
            self._writer.submit(tensors, paths)
        return batch

    def _inference_part(self, part, dataloader):
        self.is_train = False
        self.model.eval()
        if self.evaluation_metrics is not None:
            self.evaluation_metrics.reset()
        if self.save_path is not None:
            (self.save_path / part).mkdir(exist_ok=True, parents=True)
            self._writer = _AsyncWriter(self.device)
        try:
            with torch.no_grad():
                for batch_idx, batch in enumerate(dataloader):
                    self.process_batch(batch_idx=batch_idx, batch=batch, part=part, metrics=self.evaluation_metrics)
        except BaseException as loop_error:
            if self._writer is not None:
                writer, self._writer = self._writer, None
                try:
                    writer.close()
                except Exception as writer_error:   # keep the loop's exception, chain the writer's
                    raise loop_error from writer_error
            raise
        if self._writer is not None:
            writer, self._writer = self._writer, None
            writer.close()
        return self.evaluation_metrics.result() if self.evaluation_metrics is not None else {}


def _load_wav(path):
    """First channel of a PCM / float wav file as a (1, T) float32 tensor (src/utils/eval_si_snri.py:11-17)."""
    try:
        import torchaudio
        audio, _ = torchaudio.load(path)
        return audio[0:1, :]
    except ImportError:
        import numpy as np
        from scipy.io import wavfile
        _, data = wavfile.read(path)
        if data.ndim > 1:
            data = data[:, 0]
        if np.issubdtype(data.dtype, np.integer):
            data = data.astype(np.float32) / float(np.iinfo(data.dtype).max + 1)
        return torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32))[None, :]


def calculate(pred_dir, gt_dir, sr=16000, device="cuda"):
    """Offline SI-SNRi over saved predictions: mean over files of the per-file PIT SI-SNRi
    (src/utils/eval_si_snri.py:19-39; `gt_dir` holds mix/, s1/, s2/ wav folders, `pred_dir` the .pth files)."""
    from .metrics import SISNRiMetric
    metric = SISNRiMetric()
    values = []
    for s1_path in sorted(Path(gt_dir, "s1").iterdir()):
        pred = torch.load(Path(pred_dir, s1_path.stem + ".pth"), map_location="cpu")
        wav = s1_path.stem + ".wav"
        args = dict(mix=_load_wav(Path(gt_dir, "mix", wav)), s1=_load_wav(s1_path), s2=_load_wav(Path(gt_dir, "s2", wav)),
                    s1_pred=pred["s1_pred"][None, :], s2_pred=pred["s2_pred"][None, :])
        values.append(metric.device_value(**{k: v.to(device) for k, v in args.items()}).to(torch.float64))
    if not values:
        raise ValueError(f"no ground-truth files under {Path(gt_dir, 's1')}")
    return {"SiSNRi": float(torch.stack(values).mean().item())}
