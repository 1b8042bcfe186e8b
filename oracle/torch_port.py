"""CPU port of the reference forward on stock torch ops (TEST / BASELINE INFRASTRUCTURE ONLY).

The reference's CPU path is "whatever ATen / oneDNN kernels torch.nn dispatches to".  This port
drives the very same torch.nn modules - the parameter containers that
speech_separation_b200.model builds are real nn.Conv1d / nn.MultiheadAttention / nn.LSTM /
nn.LayerNorm objects - in the reference's order and tensor layouts, so timing it on the GPU
box's host cores is timing the reference's CPU path without needing /root/reference there.
Used by bench.py (`cpu_baseline`, `--impl reference`) and by tests as a second oracle.

Follows: src/model/dptn_wav.py:35-61,171-194; src/model/dptn.py:36-52,62-79,117-143,183-195;
src/model/dprnn.py:24-47,65-89,122-163,200-227,263-276.
"""
import torch
import torch.nn.functional as F


def _split(x, C, P):
    B, N, L = x.shape
    y = F.unfold(x.reshape(B, N, L, 1), kernel_size=(C, 1), stride=(P, 1))
    return y.reshape(B, N, C, -1).permute(0, 1, 3, 2).contiguous()


def _overlap_add(y, C, P):
    B, N, S, _ = y.shape
    L = (S - 1) * P + C
    z = y.permute(0, 1, 3, 2).reshape(B, N * C, S)
    return F.fold(z, output_size=(L, 1), kernel_size=(C, 1), stride=(P, 1)).squeeze(3)


def _centre_pad(x, target):
    d = target - x.shape[-1]
    return F.pad(x, (d // 2, d - d // 2))


def _transformer(m, z):
    a = m.mha(z, z, z, need_weights=False)[0] + z
    a = m.ln1(a)
    r = m.rnn(a)[0]
    return m.ln2(m.ffn(r) + a)


def _chunk_rnn(m, z):
    return m.norm1d(m.fc(m.rnn(z)[0])) + z


@torch.no_grad()
def forward(net, mix, s1_embedding=None, s2_embedding=None):
    """net: a speech_separation_b200.model module living on the CPU, in eval() mode."""
    kind = net.KIND
    d = net._desc
    C, P, N = d.C, d.P, d.N
    x = mix.unsqueeze(1)
    enc = net.encoder(x)
    if kind == "dptn_av":
        v = torch.cat([net.visual_compression(s1_embedding.permute(0, 2, 1)),
                       net.visual_compression(s2_embedding.permute(0, 2, 1))], -1)
        v = F.interpolate(v.permute(0, 2, 1), size=enc.shape[-1], mode="linear", align_corners=False).permute(0, 2, 1)
        enc = enc + net.gate.tanh() * net.video_ln(v).permute(0, 2, 1)
    B, _, L = enc.shape
    h = _split(enc, C, P)
    S = h.shape[2]
    sub = _chunk_rnn if kind == "dprnn" else _transformer
    for blk in net.dprnn.model:
        z = h.permute(0, 2, 3, 1).reshape(B * S, C, N)
        z = sub(blk.intra_chunk_block, z)
        z = z.reshape(B, S, C, N).permute(0, 2, 1, 3).reshape(B * C, S, N)
        z = sub(blk.inter_chunk_block, z)
        h = z.reshape(B, C, S, N).permute(0, 3, 2, 1).contiguous()
    h = net.dprnn.speakers_separation(h)
    h = _centre_pad(_overlap_add(h, C, P), L).view(B, 2, N, L).transpose(0, 1)
    outs = []
    for o in h:
        if kind == "dptn_mask":
            u = torch.relu(net.dprnn.output(o) * net.dprnn.output_gate(o)) * enc
        else:
            u = net.dprnn.postprocessing(o) + enc
        outs.append(_centre_pad(net.decoder(u), mix.shape[-1]).squeeze(1))
    return {"s1_pred": outs[0], "s2_pred": outs[1]}
