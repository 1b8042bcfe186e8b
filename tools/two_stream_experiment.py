"""Experiment: does running two half-batches on two CUDA streams fill the SMs left idle by the LSTM / attention phases?"""
import sys, time, torch
sys.path.insert(0, '/root/repo')
import bench
import speech_separation_b200 as V
dev = torch.device('cuda:0')
torch.manual_seed(42)
def mk():
    return V.DPTNAVWavEncDec(**bench.MODEL_KW).eval().to(dev)
B, T = 32, 64000
mix, s1, s2, e1, e2 = (t.to(dev) for t in bench.make_batch(B, T, 1234))
def timed(fn, n=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
net = mk()
print("1 stream, B=32:", timed(lambda: net(mix=mix, s1_embedding=e1, s2_embedding=e2)))
from speech_separation_b200 import _lib
lib = _lib.load()
for nsplit, cap in ((2, 0), (2, 74), (2, 100), (2, 60), (4, 37), (4, 74)):
    lib.vatss_debug_cta_limit(cap)
    nets = [mk() for _ in range(nsplit)]
    for n_ in nets: n_.load_state_dict(net.state_dict())
    streams = [torch.cuda.Stream() for _ in range(nsplit)]
    h = B // nsplit
    def run():
        cur = torch.cuda.current_stream()
        for i, (n_, st) in enumerate(zip(nets, streams)):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                n_(mix=mix[i*h:(i+1)*h], s1_embedding=e1[i*h:(i+1)*h], s2_embedding=e2[i*h:(i+1)*h])
        for st in streams: cur.wait_stream(st)
    print(f"{nsplit} streams, B={h} each, CTA cap {cap}:", timed(run))
lib.vatss_debug_cta_limit(0)
