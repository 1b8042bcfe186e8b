"""clock64 / globaltimer stamps of the CTAs of one tcgen05 convolution launch (layer 1, 9 K slabs): where a tile's time goes.
Calls the C ABI of one convolution through the whole forward with a trace buffer set; the LAST conv launch that has
>= 1024 tiles in grid x wins, so the forward is cut after layer 1 by using a tiny trunk... simpler: trace every launch and
read the buffer after the forward (it then holds the last launch whose CTAs wrote it: layer 4 has < 1024 tiles, earlier
rows keep the values of the launches before).  Usage: python tools/lipreader_trace.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_separation_b200 import Lipreading, _lib, extract_embeddings  # noqa: E402

torch.manual_seed(0)
net = Lipreading(relu_type="swish", extract_feats=True).to("cuda:0").set_engine("tensor")
vid = (torch.rand(10, 100, 96, 96, device="cuda:0") * 255).round()
lib = _lib.load()
extract_embeddings(net, vid)
torch.cuda.synchronize()
buf = torch.zeros(8192, dtype=torch.int64, device="cuda:0")
lib.vatss_debug_lipreader_trace(buf.data_ptr())
extract_embeddings(net, vid)
torch.cuda.synchronize()
lib.vatss_debug_lipreader_trace(None)
t = buf.cpu().numpy().reshape(1024, 8)
# rows 400..1023 were last written by a launch with >= 1024 tiles in x: the last of those is layer 2 (946)? no: layer 1 (3782)
rows = t[950:1024]     # only launches with > 950 x-tiles wrote these: the front GEMM (15125) and layer 1 (3782): last = layer 1 conv 4
g0 = rows[:, 0].min()
print("CTA  sm  start_us  prologue  slabs_written  acc_done  epilogue  teardown  total_cycles")
for i, r in enumerate(rows[:40]):
    print(950 + i, int(r[7]), f"{(r[0] - g0) / 1e3:8.2f}", r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], r[6] - r[5], r[6] - r[1])
d = rows
print("median cycles: prologue", np.median(d[:, 2] - d[:, 1]), "slabs", np.median(d[:, 3] - d[:, 2]), "wait acc", np.median(d[:, 4] - d[:, 3]),
      "epilogue", np.median(d[:, 5] - d[:, 4]), "teardown", np.median(d[:, 6] - d[:, 5]), "total", np.median(d[:, 6] - d[:, 1]))
# per SM: gaps between consecutive CTAs
t2 = t[t[:, 0] > 0]
print("start-time span of traced CTAs (us):", (t2[:, 0].max() - t2[:, 0].min()) / 1e3)
