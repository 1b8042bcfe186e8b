"""Utterance sharding across GPUs and the one exchange step of the path.

Utterances are independent in the forward pass, so rank r of W processes the contiguous range
[r*n/W, (r+1)*n/W) with replicated weights and no data-path collective.  The only exchange is
an all-reduce(sum) of a 12-element fp64 vector (six per-pair SI-SNR sums, the per-utterance-PIT
SI-SNRi sum, the count, four per-pair loss sums: 96 bytes, one `ncclAllReduce` over NVLink on GPUs,
gloo in the CPU tests).  Host-side logic only; the tensors it reduces are produced by
`vatss_pit_sisnr`.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced split of n utterances; earlier ranks take the remainder."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def separate_in_micro_batches(net, mix, s1_embedding=None, s2_embedding=None, micro_batch=32):
    """Run `net` over a large shard in micro-batches (bounded workspace: ~0.22 GB per 4-s utterance).

    cfg-3 of BASELINE.json gives every rank 128..512 ten-second utterances; the activations of one micro-batch of
    32 x 10 s are ~17 GB, so the shard is streamed through the same workspace.  Returns the concatenated predictions.
    """
    outs1, outs2 = [], []
    for lo in range(0, mix.shape[0], micro_batch):
        hi = min(mix.shape[0], lo + micro_batch)
        kw = {"mix": mix[lo:hi]}
        if s1_embedding is not None:
            kw["s1_embedding"], kw["s2_embedding"] = s1_embedding[lo:hi], s2_embedding[lo:hi]
        out = net(**kw)
        outs1.append(out["s1_pred"])
        outs2.append(out["s2_pred"])
    return {"s1_pred": torch.cat(outs1), "s2_pred": torch.cat(outs2)}


def sisnr_sums(rows, rows_loss=None):
    """Local sums to be reduced: six per-pair SI-SNR sums, per-utterance-PIT SI-SNRi sum, count, loss sums x4."""
    rows = rows.double()
    sep = torch.maximum((rows[:, 0] + rows[:, 1]) / 2, (rows[:, 2] + rows[:, 3]) / 2)
    snri = sep - (rows[:, 4] + rows[:, 5]) / 2
    parts = [rows.sum(0), snri.sum().reshape(1), torch.tensor([rows.shape[0]], dtype=torch.float64, device=rows.device)]
    if rows_loss is not None:
        parts.append(rows_loss.double().sum(0))
    else:
        parts.append(torch.zeros(4, dtype=torch.float64, device=rows.device))
    return torch.cat(parts)


def all_reduce_sisnr_sums(sums, group=None):
    """The exchange step alone: all-reduce(sum) of the 12-element vector of `sisnr_sums`, result left ON THE DEVICE.

    The collective is stream-ordered (the current stream waits for it, the host does not), so an evaluation loop can
    issue it every step and read the metrics once per partition (`metrics_from_sums(t.tolist())`) instead of
    synchronising the host with the GPU after every batch."""
    sums = sums.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def reduce_sisnr(sums, group=None):
    """all-reduce(sum) the vector of `sisnr_sums` and derive the global metrics (one host read).

    Returns dict with: si_snri_batch_pit (reference batch-level PIT over the GLOBAL batch),
    si_snri_utt_pit (mean of per-utterance PIT SI-SNRi), si_snr_batch_pit, loss_batch_pit, count.
    """
    return metrics_from_sums(all_reduce_sisnr_sums(sums, group).tolist())


def metrics_from_sums(s):
    """Global metrics from the reduced 12-element vector (a Python list)."""
    n = s[7]
    m = [v / n for v in s[:6]]
    sep = max((m[0] + m[1]) / 2, (m[2] + m[3]) / 2)
    l = [v / n for v in s[8:12]]
    l1, l2 = (l[0] + l[1]) / 2, (l[2] + l[3]) / 2
    return {
        "si_snr_batch_pit": sep,
        "si_snri_batch_pit": sep - (m[4] + m[5]) / 2,
        "si_snri_utt_pit": s[6] / n,
        "loss_batch_pit": l2 if l2 < l1 else l1,
        "count": int(round(n)),
    }
