"""CPU restatement of the reference lipreader's feature path.

TEST INFRASTRUCTURE ONLY: imported by tests/, by oracle/gen_lipreader_golden.py and by tools/; the product package
never imports it (the CUDA path in speech_separation_b200/lipreader.py has no CPU fallback).

What it follows (reference, read-only):
  * `Lipreading.forward` with modality="video", backbone_type="resnet", extract_feats=True
    - src/lipreader/lipreading/model.py:252-273; frontend3D :180-207; threeD_to_2D_tensor :19-22
  * ResNet-18 trunk of BasicBlocks - src/lipreader/lipreading/models/resnet.py:31-84 (block), :87-145 (stages, avgpool)
  * Swish - src/lipreader/lipreading/models/swish.py:10
  * the "val"/"test" video preprocessing - src/lipreader/lipreading/dataloaders.py:13-29 with
    Normalize / CenterCrop of src/lipreader/lipreading/preprocess.py:62-103

Parity pin: `tests/golden/lipreader_*.npz` are outputs of the reference's own `Lipreading` class (imported unmodified
by oracle/gen_lipreader_golden.py) on the weights `make_state_dict` generates; tests/test_lipreader.py checks both
restatements against them: `forward` (torch CPU ops in the reference's order, fp32 / fp64) and `forward_numpy` (numpy
only, fp64: sliding windows + einsum - none of torch's convolution, pooling or BatchNorm code).  The reference ships no pretrained lipreader weights and has no tests for this path, so the
weights are synthetic: numpy PCG64 streams per tensor, BatchNorm running statistics and affine terms randomised so that
the folding of BatchNorm into the convolutions is actually exercised.
"""
import zlib

import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-5   # nn.BatchNorm default


def state_dict_names(relu_type):
    """Keys of the reference state_dict on the feature path, with their shapes."""
    out = []

    def bn(prefix, c):
        out.extend([(prefix + ".weight", (c,)), (prefix + ".bias", (c,)), (prefix + ".running_mean", (c,)),
                    (prefix + ".running_var", (c,))])

    inpl = 64
    for li, pl in enumerate((64, 128, 256, 512), start=1):
        for b in range(2):
            p = f"trunk.layer{li}.{b}"
            out.append((p + ".conv1.weight", (pl, inpl, 3, 3)))
            bn(p + ".bn1", pl)
            if relu_type == "prelu":
                out.append((p + ".relu1.weight", (pl,)))
            out.append((p + ".conv2.weight", (pl, pl, 3, 3)))
            bn(p + ".bn2", pl)
            if relu_type == "prelu":
                out.append((p + ".relu2.weight", (pl,)))
            if li > 1 and b == 0:
                out.append((p + ".downsample.0.weight", (pl, inpl, 1, 1)))
                bn(p + ".downsample.1", pl)
            inpl = pl
    out.append(("frontend3D.0.weight", (64, 1, 5, 7, 7)))
    bn("frontend3D.1", 64)
    if relu_type == "prelu":
        out.append(("frontend3D.2.weight", (64,)))
    return out


def make_state_dict(relu_type, seed=2024):
    """Deterministic synthetic weights (platform-independent: numpy PCG64, one stream per tensor name)."""
    sd = {}
    for name, shape in state_dict_names(relu_type):
        rng = np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))
        if name.endswith("running_var"):
            v = rng.uniform(0.5, 1.5, shape)
        elif name.endswith("running_mean"):
            v = rng.normal(0.0, 0.2, shape)
        elif ".bn" in name or "downsample.1" in name or "frontend3D.1" in name:
            v = rng.uniform(0.6, 1.4, shape) if name.endswith("weight") else rng.normal(0.0, 0.1, shape)
        elif "relu" in name or name == "frontend3D.2.weight":
            v = rng.uniform(0.1, 0.4, shape)
        else:   # convolution: the reference's own init scale, model.py:281-293
            n = float(np.prod(shape[2:]) * shape[0])
            v = rng.normal(0.0, np.sqrt(2.0 / n), shape)
        sd[name] = torch.from_numpy(v.astype(np.float32))
    return sd


def make_frames(B, T, H=96, W=96, seed=7):
    """Synthetic mouth crops with values in [0, 255] (uint8-valued, smooth in space and time plus noise)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    yy, xx = np.meshgrid(np.linspace(-1, 1, H), np.linspace(-1, 1, W), indexing="ij")
    out = np.empty((B, T, H, W), np.float32)
    for b in range(B):
        ph = rng.uniform(0, 6.28, 3)
        for t in range(T):
            img = 120 + 60 * np.sin(3 * xx + ph[0] + 0.3 * t) * np.cos(2 * yy + ph[1]) + 30 * np.sin(5 * yy * xx + ph[2] - 0.2 * t)
            img = img + rng.normal(0, 12, (H, W))
            out[b, t] = np.clip(np.rint(img), 0, 255)
    return out


def preprocess(frames):
    """dataloaders.py:24-27 on (T, H, W) or (B, T, H, W) arrays: Normalize(0, 255), CenterCrop(88, 88),
    Normalize(0.421, 0.165) - in the reference's operation order and dtype promotion (numpy float32 in, python floats)."""
    x = (frames - 0.0) / 255.0
    h, w = x.shape[-2:]
    th, tw = 88, 88
    dw = int(round((w - tw)) / 2.0)
    dh = int(round((h - th)) / 2.0)
    x = x[..., dh:dh + th, dw:dw + tw]
    return (x - 0.421) / 0.165


def _act(x, relu_type, slope):
    if relu_type == "relu":
        return F.relu(x)
    if relu_type == "prelu":
        return F.prelu(x, slope)
    return x * torch.sigmoid(x)


def _bn(x, sd, prefix):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], training=False, eps=EPS)


def forward(sd, x, relu_type, dtype=torch.float32):
    """x (B, 1, T, H, W) preprocessed -> (B, T, 512)."""
    sd = {k: v.to(dtype) for k, v in sd.items()}
    x = x.to(dtype)
    B = x.shape[0]
    y = F.conv3d(x, sd["frontend3D.0.weight"], stride=(1, 2, 2), padding=(2, 3, 3))
    y = _act(_bn(y, sd, "frontend3D.1"), relu_type, sd.get("frontend3D.2.weight"))
    y = F.max_pool3d(y, kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1))
    Tn = y.shape[2]
    y = y.transpose(1, 2).reshape(B * Tn, y.shape[1], y.shape[3], y.shape[4])
    for li in range(1, 5):
        for b in range(2):
            p = f"trunk.layer{li}.{b}"
            stride = 2 if (li > 1 and b == 0) else 1
            o = F.conv2d(y, sd[p + ".conv1.weight"], stride=stride, padding=1)
            o = _act(_bn(o, sd, p + ".bn1"), relu_type, sd.get(p + ".relu1.weight"))
            o = _bn(F.conv2d(o, sd[p + ".conv2.weight"], stride=1, padding=1), sd, p + ".bn2")
            r = y
            if (p + ".downsample.0.weight") in sd:
                r = _bn(F.conv2d(y, sd[p + ".downsample.0.weight"], stride=stride), sd, p + ".downsample.1")
            y = _act(o + r, relu_type, sd.get(p + ".relu2.weight"))
    y = y.mean(dim=(2, 3))
    return y.view(B, Tn, y.shape[1])


# --------------------------------------------------------------------------------------
# numpy-only restatement (fp64): independent of torch's convolution / pooling / BatchNorm kernels
# --------------------------------------------------------------------------------------
def _np_conv(x, w, stride, pad):
    """x (N, C, *spatial), w (O, C, *kernel): cross-correlation with zero padding, as nn.Conv2d / nn.Conv3d compute it."""
    nd = w.ndim - 2
    x = np.pad(x, [(0, 0), (0, 0)] + [(p, p) for p in pad])
    win = np.lib.stride_tricks.sliding_window_view(x, w.shape[2:], axis=tuple(range(2, 2 + nd)))
    win = win[(slice(None), slice(None)) + tuple(slice(None, None, s) for s in stride)]   # (N, C, *out, *kernel)
    k = "xyz"[:nd]
    o = "pqr"[:nd]
    return np.einsum(f"ni{o}{k},mi{k}->nm{o}", win, w, optimize=True)


def _np_bn(x, sd, prefix):
    shape = (1, -1) + (1,) * (x.ndim - 2)
    g, b = sd[prefix + ".weight"].reshape(shape), sd[prefix + ".bias"].reshape(shape)
    m, v = sd[prefix + ".running_mean"].reshape(shape), sd[prefix + ".running_var"].reshape(shape)
    return (x - m) / np.sqrt(v + EPS) * g + b


def _np_act(x, relu_type, slope):
    if relu_type == "relu":
        return np.maximum(x, 0.0)
    if relu_type == "prelu":
        return np.where(x >= 0, x, x * slope.reshape((1, -1) + (1,) * (x.ndim - 2)))
    return x / (1.0 + np.exp(-x))


def forward_numpy(sd, x, relu_type):
    """Same contract as `forward`, numpy fp64 only: x (B, 1, T, H, W) -> (B, T, 512)."""
    sd = {k: np.asarray(v, dtype=np.float64) for k, v in sd.items()}
    x = np.asarray(x, dtype=np.float64)
    B = x.shape[0]
    y = _np_conv(x, sd["frontend3D.0.weight"], (1, 2, 2), (2, 3, 3))
    y = _np_act(_np_bn(y, sd, "frontend3D.1"), relu_type, sd.get("frontend3D.2.weight"))
    y = np.pad(y, [(0, 0), (0, 0), (0, 0), (1, 1), (1, 1)], constant_values=-np.inf)      # MaxPool3d(1x3x3, s 1x2x2, p 0x1x1)
    y = np.lib.stride_tricks.sliding_window_view(y, (3, 3), axis=(3, 4))[:, :, :, ::2, ::2].max(axis=(-1, -2))
    Tn = y.shape[2]
    y = y.transpose(0, 2, 1, 3, 4).reshape(B * Tn, y.shape[1], y.shape[3], y.shape[4])
    for li in range(1, 5):
        for b in range(2):
            p = f"trunk.layer{li}.{b}"
            stride = 2 if (li > 1 and b == 0) else 1
            o = _np_conv(y, sd[p + ".conv1.weight"], (stride, stride), (1, 1))
            o = _np_act(_np_bn(o, sd, p + ".bn1"), relu_type, sd.get(p + ".relu1.weight"))
            o = _np_bn(_np_conv(o, sd[p + ".conv2.weight"], (1, 1), (1, 1)), sd, p + ".bn2")
            r = y
            if (p + ".downsample.0.weight") in sd:
                r = _np_bn(_np_conv(y, sd[p + ".downsample.0.weight"], (stride, stride), (0, 0)), sd, p + ".downsample.1")
            y = _np_act(o + r, relu_type, sd.get(p + ".relu2.weight"))
    return y.mean(axis=(2, 3)).reshape(B, Tn, -1)
