"""Pin the numpy oracle against fixtures produced by the reference's own modules.

The fixtures come from oracle/gen_golden.py (reference imported unmodified, CPU fp32).
"""
import glob
import os

import numpy as np
import pytest

from oracle import vatss_oracle as O

KINDS = ["dptn_av", "dptn_wav", "dptn_mask", "dprnn"]


def load_tiny(golden_dir, kind):
    z = np.load(os.path.join(golden_dir, f"tiny_{kind}.npz"))
    kw = {k: v for k, v in zip(z["cfg_keys"], z["cfg_vals"])}
    cfg = O.PathConfig(
        kind=kind,
        num_features=int(kw["num_features"]),
        kernel_size_enc=int(kw["kernel_size_enc"]),
        hidden_dim=int(kw["hidden_dim"]),
        num_blocks=int(kw["num_blocks"]),
        chunk_size=int(kw["chunk_size"]),
        step_size=int(kw["step_size"]),
        num_heads=int(kw.get("num_heads", 4)),
        bidir=bool(kw["bidir"]),
        video_emb_size=int(kw.get("video_emb_size", 0)),
    )
    sd = {k[3:]: z[k] for k in z.files if k.startswith("sd.")}
    return z, cfg, sd


def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))


@pytest.mark.parametrize("kind", KINDS)
def test_oracle_matches_reference_forward(golden_dir, kind):
    z, cfg, sd = load_tiny(golden_dir, kind)
    P = O.to_numpy_state(sd, np.float64)
    taps = {}
    e1 = z["e1"] if "e1" in z.files else None
    e2 = z["e2"] if "e2" in z.files else None
    s1p, s2p = O.forward(P, cfg, z["mix"], e1, e2, taps=taps)
    # the reference ran in fp32, the oracle in fp64: agreement at fp32 round-off level
    assert rel_l2(s1p, z["s1_pred"]) < 2e-5
    assert rel_l2(s2p, z["s2_pred"]) < 2e-5
    # stage taps: encoded is (B,N,L) in the reference, token-major here
    assert rel_l2(taps["encoded"].transpose(0, 2, 1), z["tap.encoded"]) < 1e-5
    last = taps[f"block{cfg.num_blocks - 1}"]  # (B,S,C,N) -> reference (B,N,S,C)
    assert rel_l2(last.transpose(0, 3, 1, 2), z["tap.blocks_out"]) < 2e-5


@pytest.mark.parametrize("kind", KINDS)
def test_oracle_fp32_mode_close(golden_dir, kind):
    z, cfg, sd = load_tiny(golden_dir, kind)
    P = O.to_numpy_state(sd, np.float32)
    e1 = z["e1"] if "e1" in z.files else None
    e2 = z["e2"] if "e2" in z.files else None
    s1p, s2p = O.forward(P, cfg, z["mix"], e1, e2)
    assert s1p.dtype == np.float32
    assert rel_l2(s1p, z["s1_pred"]) < 5e-5
    assert rel_l2(s2p, z["s2_pred"]) < 5e-5


def test_segmentation_and_overlap_add_bit_exact(golden_dir):
    z = np.load(os.path.join(golden_dir, "segola.npz"))
    n = len([k for k in z.files if k.endswith(".meta")])
    assert n >= 5
    for i in range(n):
        B, N, L, C, P = (int(v) for v in z[f"c{i}.meta"])
        seg = O.segment_channel_major(z[f"c{i}.x"], C, P)
        assert seg.shape == z[f"c{i}.seg"].shape
        assert np.array_equal(seg, z[f"c{i}.seg"])
        ola = O.overlap_add_channel_major(z[f"c{i}.y"], P)
        assert np.array_equal(ola, z[f"c{i}.ola"])
        # token-major view agrees with the channel-major one
        tm = O.segment_token_major(z[f"c{i}.x"].transpose(0, 2, 1), C, P)
        assert np.array_equal(tm.transpose(0, 3, 1, 2), z[f"c{i}.seg"])


def test_loss_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "loss.npz"))
    for i in range(3):
        s1, s2, s1p, s2p = (z[f"c{i}.{k}"].astype(np.float64) for k in ("s1", "s2", "s1p", "s2p"))
        pair = [O.sisnr_loss_rows(a, b).mean() for a, b in [(s1p, s1), (s2p, s2), (s1p, s2), (s2p, s1)]]
        np.testing.assert_allclose(pair, z[f"c{i}.pair"], rtol=2e-5)
        np.testing.assert_allclose(O.pit_sisnr_loss(s1p, s2p, s1, s2), z[f"c{i}.loss"], rtol=2e-5)
        # the restated torchmetrics SI-SNR is -loss/2 up to the eps terms (SURVEY.md §8c cross-check)
        m = np.mean(O.si_snr_metric_rows(s1p, s1))
        np.testing.assert_allclose(m, -0.5 * z[f"c{i}.pair"][0], rtol=1e-4)
    # case 1 was built with swapped speakers: batch-level PIT must choose permutation 2
    pair = z["c1.pair"]
    assert (pair[2] + pair[3]) / 2 < (pair[0] + pair[1]) / 2
    np.testing.assert_allclose(z["c1.loss"], (pair[2] + pair[3]) / 2, rtol=1e-6)


@pytest.mark.parametrize("kind", KINDS)
def test_tiny_loss_golden(golden_dir, kind):
    z, cfg, sd = load_tiny(golden_dir, kind)
    got = O.pit_sisnr_loss(*(z[k].astype(np.float64) for k in ("s1_pred", "s2_pred", "s1", "s2")))
    np.testing.assert_allclose(got, z["loss"], rtol=1e-4)


def test_prod_goldens_present(golden_dir):
    files = sorted(glob.glob(os.path.join(golden_dir, "prod_*.npz")))
    assert len(files) >= 5
    for f in files:
        z = np.load(f)
        assert z["s1_pred"].shape == (int(z["B"]), int(z["T"]))
        assert np.isfinite(z["s1_pred"]).all() and np.isfinite(z["s2_pred"]).all()


@pytest.mark.parametrize("kind", KINDS)
def test_torch_port_reproduces_reference_goldens_exactly(golden_dir, kind):
    """oracle/torch_port.forward (the CPU baseline / `--impl reference` arm of bench.py) drives the same torch.nn
    modules in the same order as the reference: on the production 1-s goldens its outputs must EQUAL the reference's
    (same torch build, same thread count independent kernels: bitwise or within 1 ulp-level noise of reduction order)."""
    import torch

    import speech_separation_b200 as V
    from oracle import torch_port
    from oracle.gen_golden import PROD, make_inputs

    cls = {"dptn_av": V.DPTNAVWavEncDec, "dptn_wav": V.DPTNWavEncDec, "dptn_mask": V.DPTNEncDec,
           "dprnn": V.DPRNNEncDec}[kind]
    z = np.load(os.path.join(golden_dir, f"prod_{kind}_B2_T16000.npz"))
    torch.manual_seed(int(z["weight_seed"]))
    net = cls(**PROD[kind]).eval()
    Tv = int(z["Tv"]) if int(z["Tv"]) > 0 else None
    mix, s1, s2, e1, e2 = make_inputs(2, 16000, Tv=Tv, E=PROD[kind].get("video_emb_size"), seed=int(z["input_seed"]))
    out = torch_port.forward(net, mix, e1, e2)
    for k in ("s1_pred", "s2_pred"):
        got = out[k].numpy()
        assert got.shape == z[k].shape
        # thread-count dependent reduction order inside ATen is the only freedom: far below the 1e-3 budget
        assert rel_l2(got, z[k]) < 2e-6, (kind, k, rel_l2(got, z[k]))


def test_oracle_loss_gradient_matches_reference_autograd(golden_dir):
    """oracle.pit_sisnr_loss_grad (closed form) against torch autograd through the reference's own SiSNRWavLoss
    (oracle/gen_loss_grad_golden.py): the gradient `batch["loss"].backward()` leaves on the predictions."""
    z = np.load(os.path.join(golden_dir, "loss.npz"))
    g = np.load(os.path.join(golden_dir, "loss_grad.npz"))
    for i in range(3):
        a = [z[f"c{i}.{k}"].astype(np.float64) for k in ("s1p", "s2p", "s1", "s2")]
        g1, g2 = O.pit_sisnr_loss_grad(*a)
        up = float(g[f"c{i}.upstream"])
        assert rel_l2(up * g1, g[f"c{i}.g1"]) < 2e-6
        assert rel_l2(up * g2, g[f"c{i}.g2"]) < 2e-6
