"""Multi-GPU path: utterance shards on separate GPUs, NCCL all-reduce of the SI-SNR
sums.  Needs >= 2 CUDA devices (skipped otherwise); the host-side logic is covered on CPU with gloo in
tests/test_cpu_host.py."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

KW = dict(num_features=128, video_emb_size=512, hidden_video=128, kernel_size_enc=7, hidden_dim=128, num_blocks=2,
          chunk_size=150, step_size=75, num_heads=4, dropout=0.1, bidir=True)


def _inputs(n, T, Tv):
    g = torch.Generator().manual_seed(2024)
    s1 = 0.1 * torch.randn(n, T, generator=g)
    s2 = 0.1 * torch.randn(n, T, generator=g)
    return s1 + s2, s1, s2, torch.randn(n, 512, Tv, generator=g), torch.randn(n, 512, Tv, generator=g)


def _nccl_worker(rank, world, port, n, T, Tv, micro, q):
    import torch.distributed as dist

    import speech_separation_b200 as V
    from speech_separation_b200.sharding import reduce_sisnr, separate_in_micro_batches, shard_range, sisnr_sums

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    torch.manual_seed(42)
    net = V.DPTNAVWavEncDec(**KW).eval().to(dev)
    mix, s1, s2, e1, e2 = _inputs(n, T, Tv)
    lo, hi = shard_range(n, rank, world)
    out = separate_in_micro_batches(net, mix[lo:hi].to(dev), e1[lo:hi].to(dev), e2[lo:hi].to(dev), micro_batch=micro)
    rows, rows_loss, _ = V.pit_sisnr_all(out["s1_pred"], out["s2_pred"], s1[lo:hi].to(dev), s2[lo:hi].to(dev), mix[lo:hi].to(dev))
    red = reduce_sisnr(sisnr_sums(rows, rows_loss))
    loss = V.SiSNRWavLoss()(s1_pred=out["s1_pred"], s2_pred=out["s2_pred"], s1=s1[lo:hi].to(dev), s2=s2[lo:hi].to(dev))["loss"]
    lsum = loss.double() * (hi - lo)               # cfg-5: data-parallel mean of the per-rank (batch-level PIT) losses
    dist.all_reduce(lsum)
    if rank == 0:
        q.put((red, float(lsum) / n, out["s1_pred"].cpu()))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_nccl_reduction_equals_single_gpu():
    import torch.multiprocessing as mp

    import speech_separation_b200 as V
    from speech_separation_b200.sharding import reduce_sisnr, sisnr_sums

    n, T, Tv, micro = 6, 16000, 25, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 2000
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, n, T, Tv, micro, q)) for r in range(2)]
    for p in procs:
        p.start()
    red, loss_dp, s1p_rank0 = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # single GPU, whole batch at once
    dev = torch.device("cuda:0")
    torch.manual_seed(42)
    net = V.DPTNAVWavEncDec(**KW).eval().to(dev)
    mix, s1, s2, e1, e2 = _inputs(n, T, Tv)
    out = net(mix=mix.to(dev), s1_embedding=e1.to(dev), s2_embedding=e2.to(dev))
    rows, rows_loss, _ = V.pit_sisnr_all(out["s1_pred"], out["s2_pred"], s1.to(dev), s2.to(dev), mix.to(dev))
    one = reduce_sisnr(sisnr_sums(rows, rows_loss))
    assert torch.equal(out["s1_pred"][:3].cpu(), s1p_rank0)        # rank 0's shard: bitwise the same utterances
    assert red["count"] == one["count"] == n
    for k in ("si_snri_batch_pit", "si_snri_utt_pit", "si_snr_batch_pit", "loss_batch_pit"):
        assert abs(red[k] - one[k]) < 1e-9, (k, red[k], one[k])
    assert abs(float(V.SISNRiMetric()(s1_pred=out["s1_pred"], s2_pred=out["s2_pred"], s1=s1.to(dev), s2=s2.to(dev),
                                      mix=mix.to(dev))) - red["si_snri_batch_pit"]) < 1e-4
    assert abs(loss_dp) < 1e3 and loss_dp == loss_dp
