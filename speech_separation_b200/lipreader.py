"""Lipreader front end, executed by libvatss_b200.so.

Host-side mirror of the reference's video feature extractor - the network that turns 96 x 96 mouth crops into the
(512, Tv) lip embeddings `DPTNAVWavEncDec.forward` consumes:

    Lipreading                 <- src/lipreader/lipreading/model.py:144-273   (modality="video", backbone_type="resnet")
    init_lipreader             <- src/utils/init_utils.py:168-207
    get_preprocessing_pipelines <- src/lipreader/lipreading/dataloaders.py:8-29  ("val" / "test" video pipeline)
    extract_embeddings         <- make_embeddings.py:52-68 (one mouth-crop clip -> (512, T) embedding)
    make_embeddings            <- make_embeddings.py:11-73 (directory of mouth-crop .npz -> directory of embedding .npz)

Scope: the feature path the reference actually runs (`extract_feats=True`: frontend3D -> ResNet-18 trunk -> (B, T, 512),
model.py:253-273).  The word-classification head (`self.tcn`, model.py:209-244) is never executed by the separation
system (init_utils.py:205 always passes extract_feats=True); checkpoints that contain `tcn.*` tensors load with those
keys ignored, as the reference's `load_model(..., allow_size_mismatch=True)` does with `strict=False`
(lipreading/utils.py:173-185).  Constructing with `extract_feats=False`, `modality="audio"`, the shufflenet backbone
or `use_boundary=True` raises NotImplementedError instead of silently doing something else.

The torch.nn sub-modules are parameter containers (reference state_dict keys and shapes: `frontend3D.0.weight`,
`frontend3D.1.*`, `trunk.layer<l>.<b>.{conv1,bn1,relu1,conv2,bn2,relu2,downsample.0,downsample.1}.*`); their `forward`s
are never called.  BatchNorm runs on its running statistics (the reference only runs the lipreader in eval(),
make_embeddings.py:47, profiler.py:15); calling `.train()` raises.  No CPU path, no autograd.
"""
import ctypes
import json
import math

import torch
from torch import nn

from . import _lib

RELU_TYPE = {"relu": 1, "prelu": 2, "swish": 3}
ENGINE = {"f32": 0, "tensor": 1}


class _Act(nn.Module):
    """Placeholder for ReLU / Swish (no parameters); PReLU slots use nn.PReLU for its `weight`."""


def _act(relu_type, channels):
    return nn.PReLU(num_parameters=channels) if relu_type == "prelu" else _Act()


class _BlockParams(nn.Module):
    """Parameters of BasicBlock (models/resnet.py:31-62)."""

    def __init__(self, inplanes, planes, stride, relu_type):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu1 = _act(relu_type, planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.relu2 = _act(relu_type, planes)
        self.downsample = None
        if stride != 1 or inplanes != planes:   # downsample_basic_block, models/resnet.py:14-18
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes, kernel_size=1, stride=stride, bias=False),
                                            nn.BatchNorm2d(planes))


class _TrunkParams(nn.Module):
    """Parameters of ResNet(BasicBlock, [2, 2, 2, 2]) (models/resnet.py:87-135)."""

    def __init__(self, relu_type):
        super().__init__()
        inplanes = 64
        for li, planes in enumerate((64, 128, 256, 512), start=1):
            blocks = []
            for b in range(2):
                stride = 2 if (li > 1 and b == 0) else 1
                blocks.append(_BlockParams(inplanes, planes, stride, relu_type))
                inplanes = planes
            setattr(self, f"layer{li}", nn.Sequential(*blocks))


def _conv_param_names(relu_type):
    """Reference state_dict keys in the order of the C parameter table (include/vatss.h, VATSS_LIP_*)."""
    prelu = relu_type == "prelu"

    def conv(w, bn, act):
        return [w + ".weight", bn + ".weight", bn + ".bias", bn + ".running_mean", bn + ".running_var",
                (act + ".weight") if (prelu and act) else None]

    names = conv("frontend3D.0", "frontend3D.1", "frontend3D.2")
    for li in range(1, 5):
        for b in range(2):
            p = f"trunk.layer{li}.{b}"
            names += conv(p + ".conv1", p + ".bn1", p + ".relu1")
            names += conv(p + ".conv2", p + ".bn2", p + ".relu2")
            if li > 1 and b == 0:
                names += conv(p + ".downsample.0", p + ".downsample.1", None)
            else:
                names += [None] * 6
    return names


class Lipreading(nn.Module):
    """Drop-in for `src.lipreader.lipreading.model.Lipreading` on its feature-extraction path."""

    def __init__(self, modality="video", hidden_dim=256, backbone_type="resnet", num_classes=500, relu_type="prelu",
                 tcn_options={}, densetcn_options={}, width_mult=1.0, use_boundary=False, extract_feats=False):
        super().__init__()
        if modality != "video" or backbone_type != "resnet":
            raise NotImplementedError("speech_separation_b200.Lipreading covers modality='video', backbone_type='resnet' "
                                      "(the lipreader VAT-SS uses: src/lipreader/configs/lrw_resnet18_mstcn.json)")
        if use_boundary:
            raise NotImplementedError("use_boundary=True is not supported")
        if not extract_feats:
            raise NotImplementedError("only extract_feats=True (the embedding path, model.py:273) is implemented; the "
                                      "word-classification TCN head is never run by the separation system")
        if relu_type not in RELU_TYPE:
            raise ValueError(f"relu_type must be one of {sorted(RELU_TYPE)}")
        self.extract_feats = True
        self.backbone_type = backbone_type
        self.modality = modality
        self.use_boundary = False
        self.relu_type = relu_type
        self.frontend_nout = 64
        self.backend_out = 512
        self.trunk = _TrunkParams(relu_type)
        self.frontend3D = nn.Sequential(
            nn.Conv3d(1, 64, kernel_size=(5, 7, 7), stride=(1, 2, 2), padding=(2, 3, 3), bias=False),
            nn.BatchNorm3d(64), _act(relu_type, 64),
            nn.MaxPool3d(kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1)))
        self._initialize_weights_randomly()
        self._engine = ENGINE["tensor"]
        self._cache_key = None
        self._table = self._packed = self._keep = None
        self._ws = {}
        self.eval()

    def _initialize_weights_randomly(self):
        # model.py:275-309: conv weights ~ N(0, sqrt(2 / (prod(kernel) * out_channels))), BatchNorm weight 1 / bias 0
        for m in self.modules():
            if isinstance(m, (nn.Conv3d, nn.Conv2d)):
                n = math.prod(m.kernel_size) * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / float(n)))
            elif isinstance(m, (nn.BatchNorm3d, nn.BatchNorm2d)):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def train(self, mode=True):
        if mode:
            raise NotImplementedError("the lipreader runs on BatchNorm running statistics only (eval mode)")
        return super().train(False)

    def set_engine(self, name):
        """'tensor' (default: tcgen05 trunk, fp16 activations) or 'f32' (fp32 FMA-pipe kernels, ~1e-6 of the reference)."""
        self._engine = ENGINE[name]
        return self

    def load_state_dict(self, state_dict, strict=True, **kw):
        # checkpoints of the full reference model also carry the classification head (`tcn.*`), which has no counterpart
        own = {k: v for k, v in state_dict.items() if not k.startswith("tcn.")}
        return super().load_state_dict(own, strict=strict, **kw)

    # -- C parameter table -------------------------------------------------------------------
    def _prepare(self, device):
        lib = _lib.load()
        tensors = dict(self.state_dict(keep_vars=True))
        names = _conv_param_names(self.relu_type)
        used = [tensors[n] for n in names if n is not None]
        key = tuple((t.data_ptr(), t._version) for t in used) + (str(device),)
        if key == self._cache_key:
            return lib
        table = (ctypes.c_void_p * len(names))()
        keep = []
        for i, n in enumerate(names):
            if n is None:
                table[i] = None
                continue
            t = tensors[n].detach()
            if not t.is_cuda or t.device != device:
                raise RuntimeError(f"{n} is on {t.device}, input on {device}: move the module with .to(device); "
                                   "speech_separation_b200 has no CPU path")
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
            keep.append(t)
            table[i] = t.data_ptr()
        packed = torch.empty(int(lib.vatss_lipreader_packed_bytes()), dtype=torch.uint8, device=device)
        _lib.check(lib.vatss_lipreader_pack_weights(table, len(names), RELU_TYPE[self.relu_type], packed.data_ptr(),
                                                    packed.numel(), _lib.stream_ptr()), "vatss_lipreader_pack_weights")
        self._table, self._packed, self._keep, self._cache_key = table, packed, keep, key
        return lib

    def _run(self, video, crop=None, pre_scale=1.0, pre_shift=0.0):
        """video (B, T, H, W) f32 on CUDA; crop = (y0, x0, Hc, Wc) or None -> (B, T, 512)."""
        video = _lib.f32c(video, "video")
        B, T, H, W = video.shape
        y0, x0, Hc, Wc = crop if crop is not None else (0, 0, H, W)
        with torch.cuda.device(video.device):
            lib = self._prepare(video.device)
            k = (B, T, Hc, Wc, str(video.device))
            ws = self._ws.get(k)
            if ws is None:
                nbytes = int(lib.vatss_lipreader_workspace_bytes(B, T, Hc, Wc))
                if nbytes == 0:
                    raise ValueError(f"unsupported lipreader input: {B} x {T} frames of {Hc} x {Wc}")
                self._ws.clear()
                ws = torch.empty(nbytes, dtype=torch.uint8, device=video.device)
                self._ws[k] = ws
            out = torch.empty(B, T, self.backend_out, dtype=torch.float32, device=video.device)
            _lib.check(lib.vatss_lipreader_forward(self._packed.data_ptr(), self._packed.numel(),
                                                   RELU_TYPE[self.relu_type], video.data_ptr(), B, T, H, W, y0, x0, Hc,
                                                   Wc, float(pre_scale), float(pre_shift), out.data_ptr(),
                                                   ws.data_ptr(), ws.numel(), self._engine, _lib.stream_ptr()),
                       "vatss_lipreader_forward")
        return out

    def forward(self, x, lengths, boundaries=None):
        """x (B, 1, T, H, W) preprocessed mouth crops -> (B, T, 512) features (model.py:252-273; `lengths` is only used
        by the classification head and is ignored here, as in the reference when extract_feats=True)."""
        if x.dim() != 5 or x.shape[1] != 1:
            raise ValueError(f"x must be (batch, 1, frames, height, width), got {tuple(x.shape)}")
        return self._run(x[:, 0])

    def __str__(self):
        n = sum(p.numel() for p in self.parameters())
        return f"{super().__str__()}\nAll parameters: {n}"


def init_lipreader(config, path=None):
    """Mirror of src/utils/init_utils.py:168-207: build the lipreader described by the JSON config and, when `path`
    is given, load `checkpoint["model_state_dict"]` (classification-head tensors are skipped)."""
    with open(config) as fp:
        args_loaded = json.load(fp)
    model = Lipreading(modality="video", num_classes=args_loaded.get("num_classes", 500),
                       backbone_type=args_loaded["backbone_type"], relu_type=args_loaded["relu_type"],
                       width_mult=args_loaded["width_mult"], use_boundary=args_loaded.get("use_boundary", False),
                       extract_feats=True)
    if path is not None:
        checkpoint = torch.load(path, map_location="cpu")
        model.load_state_dict(checkpoint["model_state_dict"], strict=False)
    return model


# the "val" / "test" video pipeline of lipreading/dataloaders.py:13-29 as numbers
CROP_SIZE = (88, 88)
PRE_MEAN, PRE_STD = 0.421, 0.165


def center_crop_window(h, w, size=CROP_SIZE):
    """CenterCrop.__call__ (lipreading/preprocess.py:98-103): (y0, x0, th, tw)."""
    th, tw = size
    delta_w = int(round((w - tw)) / 2.0)
    delta_h = int(round((h - th)) / 2.0)
    return delta_h, delta_w, th, tw


def extract_embeddings(lipreader, mouth_frames):
    """make_embeddings.py:58-66 for a batch of clips: raw mouth crops (B, T, H, W) or (T, H, W) with values in
    [0, 255] -> embeddings (B, 512, T) (the layout the datasets store and DPTNAVWavEncDec reads).

    Normalize(0, 255) -> CenterCrop(88, 88) -> Normalize(0.421, 0.165) collapses to one affine map and a crop window,
    both applied inside the first kernel's load: nothing preprocessed is materialised."""
    x = mouth_frames
    if x.dim() == 3:
        x = x[None]
    crop = center_crop_window(x.shape[2], x.shape[3])
    feats = lipreader._run(x, crop=crop, pre_scale=1.0 / (255.0 * PRE_STD), pre_shift=-PRE_MEAN / PRE_STD)
    return feats.transpose(1, 2)


def make_embeddings(lipreader, mouths_dir, embeds_dir, device="cuda", max_clips_per_call=64):
    """Mirror of make_embeddings.py:28-73 on top of `extract_embeddings`: every file of `mouths_dir` is an `.npz` whose
    `data` array holds a (T, H, W) mouth-crop clip with values in [0, 255]; the embedding (512, T) is written to
    `embeds_dir/<same file name>` with `np.savez_compressed(path, embedding=...)` - the files `SSDataset` reads back
    (src/datasets/base_dataset.py:143-149).  Clips of equal shape are batched (the reference runs them one by one; the
    results do not depend on the batch composition, bit for bit).  Returns the number of files written."""
    import os

    import numpy as np

    if not os.path.isdir(mouths_dir):
        raise NotADirectoryError(f"Input directory not found: {mouths_dir}")
    os.makedirs(embeds_dir, exist_ok=True)
    names = sorted(n for n in os.listdir(mouths_dir) if os.path.isfile(os.path.join(mouths_dir, n)))
    groups = {}
    for n in names:   # group by clip shape without keeping the pixels in memory
        with np.load(os.path.join(mouths_dir, n)) as z:
            groups.setdefault(tuple(z["data"].shape), []).append(n)
    written = 0
    for shape, members in groups.items():
        for i in range(0, len(members), max_clips_per_call):
            part = members[i:i + max_clips_per_call]
            clips = np.stack([np.load(os.path.join(mouths_dir, n))["data"] for n in part]).astype(np.float32)
            emb = extract_embeddings(lipreader, torch.from_numpy(clips).to(device))      # (len(part), 512, T)
            emb = emb.detach().cpu().numpy()
            for n, e in zip(part, emb):
                np.savez_compressed(os.path.join(embeds_dir, n), embedding=e)
                written += 1
    return written
