"""Time the tensor-engine lipreader with parts of k_lip_conv_tc switched off (vatss_debug_lipreader): where does a tile's
time go?  Results are wrong by construction with any flag set.  Usage: python tools/lipreader_ablate.py [B T]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_separation_b200 import Lipreading, _lib, extract_embeddings  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
torch.manual_seed(0)
net = Lipreading(relu_type="swish", extract_feats=True).to("cuda:0").set_engine("tensor")
vid = (torch.rand(B, T, 96, 96, device="cuda:0") * 255).round()
lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
for flags, what in ((-2, "persistent kernel (vatss_debug_lipreader_kernel(2))"), (-4, "persistent kernel, no epilogue"), (0, "production (one tile per CTA)"), (256, "64-channel layers with two CTAs per SM / three slabs in flight instead of three CTAs / two slabs"), (64, "relaxed mbarrier arrival (timing experiment)"), (128, "no fence.proxy.async (timing only, results invalid)"), (192, "relaxed arrival, no fence"), (8, "no epilogue stores"), (16, "no activation math"), (32, "no residual loads"), (56, "epilogue = TMEM read + BN only"), (1, "no gather loads"), (2, "no epilogue"), (4, "no weight TMA"), (5, "no gather, no weight TMA"),
                    (7, "no gather / epilogue / TMA: launch + barriers + MMA only")):
    lib.vatss_debug_lipreader_kernel(2 if flags < 0 else 1)
    lib.vatss_debug_lipreader(2 if flags == -4 else max(flags, 0))
    for _ in range(2):
        extract_embeddings(net, vid)
    ms = []
    for _ in range(3):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        extract_embeddings(net, vid)
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    print(f"flags {flags} ({what}): {np.median(ms):.3f} ms per {B * T} frames", flush=True)
lib.vatss_debug_lipreader(0)
lib.vatss_debug_lipreader_kernel(1)
ref = extract_embeddings(net, vid)
lib.vatss_debug_lipreader(256)
alt = extract_embeddings(net, vid)
lib.vatss_debug_lipreader(0)
print("two-CTA variant of the 64-channel kernel bit-identical to production:", bool(torch.equal(ref, alt)), flush=True)
