"""SI-SNR / SI-SNRi metrics, drop-in for src/metrics/si_snr.py:6-12 and src/metrics/si_snri.py:7-30.

The reference wraps torchmetrics' ScaleInvariantSignalNoiseRatio, calls it 4 (+2) times and
`.item()`s each result; PIT is batch-level (`max` over the two permutations of the batch means,
src/metrics/base_metric.py:53-60).  Here one kernel pass produces all six per-utterance values
and the batch-level decisions (`vatss_pit_sisnr`).
"""
import torch

from .loss import pit_sisnr_all


class SS2BaseMetric:
    def __init__(self, name=None, lower_better=False, *args, **kwargs):
        self.name = name if name is not None else type(self).__name__
        if lower_better:
            raise ValueError("the fused SI-SNR kernel implements higher-is-better PIT only")
        self.device = kwargs.get("device", None)

    def _summary(self, s1_pred, s2_pred, s1, s2, mix=None):
        if self.device is not None and self.device != "auto":
            dev = torch.device(self.device)
            s1_pred, s2_pred, s1, s2 = (t.to(dev) for t in (s1_pred, s2_pred, s1, s2))
            mix = mix.to(dev) if mix is not None else None
        return pit_sisnr_all(s1_pred, s2_pred, s1, s2, mix)

    def _shared_summary(self, shared, s1_pred, s2_pred, s1, s2, mix):
        """summary[8] of the batch; `shared` (inference._SharedSisnr) lets several metrics reuse one kernel pass."""
        if shared is not None and (self.device is None or self.device == "auto"):
            return shared.summary({"s1_pred": s1_pred, "s2_pred": s2_pred, "s1": s1, "s2": s2, "mix": mix})
        return self._summary(s1_pred, s2_pred, s1, s2, mix)[2]


class SISNRMetric(SS2BaseMetric):
    """__call__(**batch) -> float: batch-level PIT SI-SNR in dB."""

    def __call__(self, s1_pred, s2_pred, s1, s2, **batch):
        _, _, summary = self._summary(s1_pred, s2_pred, s1, s2)
        return float(summary[3].item())

    def device_value(self, s1_pred, s2_pred, s1, s2, mix=None, shared=None, **batch):
        """Same value as __call__, as a 0-d float32 device tensor (no host synchronisation)."""
        return self._shared_summary(shared, s1_pred, s2_pred, s1, s2, mix)[3].float()


class SISNRiMetric(SS2BaseMetric):
    """__call__(**batch) -> 0-d tensor: PIT SI-SNR minus the mean SI-SNR of the mixture."""

    def __call__(self, s1_pred, s2_pred, s1, s2, mix, **batch):
        _, _, summary = self._summary(s1_pred, s2_pred, s1, s2, mix)
        return summary[4].float()

    def device_value(self, s1_pred, s2_pred, s1, s2, mix, shared=None, **batch):
        return self._shared_summary(shared, s1_pred, s2_pred, s1, s2, mix)[4].float()

    def per_utterance(self, s1_pred, s2_pred, s1, s2, mix, **batch):
        """Per-utterance PIT SI-SNRi (B,) f64 - what src/utils/eval_si_snri.py:31-39 computes file by file."""
        rows, _, _ = self._summary(s1_pred, s2_pred, s1, s2, mix)
        sep = torch.maximum((rows[:, 0] + rows[:, 1]) / 2, (rows[:, 2] + rows[:, 3]) / 2)
        return sep - (rows[:, 4] + rows[:, 5]) / 2
