"""CPU ORACLE (test infrastructure, not product code) -- IMFNet descriptor forward, restated.

A functional, CPU-only restatement of the reference hot path
    ResUNet2.forward(x, image)            /root/reference/model/resunet.py:163-235
driven by a plain `state_dict` with the reference's key names, so it can run on the GPU box where
/root/reference does not exist.  Every function cites the reference lines it follows.

How it is pinned (see DESIGN.md "Oracle"):
  * oracle/make_golden.py runs the reference's UNMODIFIED model/*.py on the MinkowskiEngine stand-in
    (oracle/standin) in the build container and checks this restatement against it (max abs diff
    reported in tests/golden/MANIFEST.json), then commits input/output vectors under tests/golden/.
  * tests/test_oracle_golden.py re-checks this file against those vectors on every run.
MinkowskiEngine semantics themselves are restated from ME 0.5.4 (not in /root/reference); the only
reference-held known answer for them is files/3D_head_map.ply (quantisation order).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import sparse_ops as ops

# model/resunet.py:276-326 -- (CHANNELS, TR_CHANNELS) per BN variant; block norm 'BN' only here.
VARIANTS = {
    "ResUNetBN2": ([None, 32, 64, 128, 256], [None, 32, 64, 64, 128]),
    "ResUNetBN2B": ([None, 32, 64, 128, 256], [None, 64, 64, 64, 64]),
    "ResUNetBN2C": ([None, 32, 64, 128, 256], [None, 64, 64, 64, 128]),
    "ResUNetBN2D": ([None, 32, 64, 128, 256], [None, 64, 64, 128, 128]),
    "ResUNetBN2E": ([None, 128, 128, 128, 256], [None, 64, 128, 128, 128]),
}


def _bn(x, sd, p, eps=1e-5):
    """Eval-mode BatchNorm (model/common.py:6 -> ME.MinkowskiBatchNorm -> torch BatchNorm1d; resnet.py BN2d)."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, eps)


def _ln(x, sd, p):
    """nn.LayerNorm(dim), eps 1e-5 (model/attention_fusion.py:36-37)."""
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


# ----------------------------------------------------------------------------------------------
# image encoder: model/Img_Encoder.py:15-18 -> model/resnet.py:195-216 (ResNet-34 cut after layer2)
# ----------------------------------------------------------------------------------------------
def _basic_block2d(x, sd, p, stride, has_down):
    """model/resnet.py:59-75 BasicBlock.forward."""
    out = F.conv2d(x, sd[p + ".conv1.weight"], None, stride, 1)
    out = torch.relu(_bn(out, sd, p + ".bn1"))
    out = F.conv2d(out, sd[p + ".conv2.weight"], None, 1, 1)
    out = _bn(out, sd, p + ".bn2")
    idt = x
    if has_down:
        idt = _bn(F.conv2d(x, sd[p + ".downsample.0.weight"], None, stride, 0), sd, p + ".downsample.1")
    return torch.relu(out + idt)


def image_encoder(sd, image, prefix="img_encoder.backbone"):
    p = prefix
    x = F.conv2d(image, sd[p + ".conv1.weight"], None, 2, 3)            # resnet.py:198
    x = torch.relu(_bn(x, sd, p + ".bn1"))                                 # :199-200
    x = F.max_pool2d(x, 3, 2, 1)                                           # :203
    for i in range(3):                                                     # layer1, :205
        x = _basic_block2d(x, sd, f"{p}.layer1.{i}", 1, False)
    for i in range(4):                                                     # layer2, :207
        x = _basic_block2d(x, sd, f"{p}.layer2.{i}", 2 if i == 0 else 1, i == 0)
    return x                                                               # I1, :216


# ----------------------------------------------------------------------------------------------
# attention fusion: model/attention_fusion.py:132-154 (depth=0, cross_heads=1, mask=None)
# ----------------------------------------------------------------------------------------------
def attention_fusion(sd, data, queries, prefix="attention_fusion"):
    """data [1,L,128] image tokens, queries [1,M,256] point tokens -> [1,M,256]."""
    p0, p1 = prefix + ".cross_attend_blocks.0", prefix + ".cross_attend_blocks.1"
    x = queries
    xn = _ln(x, sd, p0 + ".norm")                                          # PreNorm, :38
    cn = _ln(data, sd, p0 + ".norm_context")                               # :41-42
    q = xn @ sd[p0 + ".fn.to_q.weight"].t()                                # :79
    kv = cn @ sd[p0 + ".fn.to_kv.weight"].t()                              # :81
    k, v = kv.chunk(2, dim=-1)                                             # :82
    scale = q.shape[-1] ** -0.5                                            # heads=1 -> dim_head = inner_dim, :69
    sim = torch.einsum("bid,bjd->bij", q, k) * scale                       # :84
    attn = sim.softmax(dim=-1)                                             # :92
    out = torch.einsum("bij,bjd->bid", attn, v)                            # :93
    out = out @ sd[p0 + ".fn.to_out.weight"].t() + sd[p0 + ".fn.to_out.bias"]   # :95
    x = out + x                                                            # :143
    h = _ln(x, sd, p1 + ".norm") @ sd[p1 + ".fn.net.0.weight"].t() + sd[p1 + ".fn.net.0.bias"]   # :57
    a, gates = h.chunk(2, dim=-1)                                          # GEGLU :50
    h = a * F.gelu(gates)                                                  # :51
    h = h @ sd[p1 + ".fn.net.2.weight"].t() + sd[p1 + ".fn.net.2.bias"]    # :59
    return h + x                                                           # :144


def transformer(sd, images, Fs8, C8):
    """model/resunet.py:237-273: split stride-8 rows by batch id, fuse each with its image tokens."""
    ps, start = [], 0
    nb = int(C8[:, 0].max()) + 1 if len(C8) else 0
    for b in range(nb):
        length = int((C8[:, 0] == b).sum())
        P = Fs8[start:start + length][None]
        img = images[b][None]
        B, C, H, W = img.shape
        tok = img.reshape(B, C, H * W).permute(0, 2, 1)
        ps.append(attention_fusion(sd, tok, P)[0])
        start += length
    return torch.cat(ps, dim=0)


# ----------------------------------------------------------------------------------------------
# sparse ResUNet
# ----------------------------------------------------------------------------------------------
def _conv(cm, x, W, t_in, t_out, K, transposed=False):
    """ME.MinkowskiConvolution(+Transpose).forward on the cached kernel map of (t_in -> t_out)."""
    return ops.conv_forward(x, W, cm.table(t_in, t_out, K, transposed), pairs=cm.pairs(t_in, t_out, K, transposed))


def _block(sd, cm, t, x, p):
    """model/residual_block.py:37-53 BasicBlockBN.forward (conv3 -> BN -> ReLU -> conv3 -> BN -> +x -> ReLU)."""
    out = _conv(cm, x, sd[p + ".conv1.kernel"], t, t, 3)
    out = torch.relu(_bn(out, sd, p + ".norm1.bn"))
    out = _conv(cm, out, sd[p + ".conv2.kernel"], t, t, 3)
    out = _bn(out, sd, p + ".norm2.bn")
    return torch.relu(out + x)


def forward(sd, coords, feats, image, normalize_feature=True, conv1_kernel_size=5, return_intermediates=False):
    """model/resunet.py:163-235.  coords int32 [N,4] (b,x,y,z) unique rows, feats [N,Cin] fp32,
    image [B,3,H,W] fp32.  Returns [N,Cout] descriptors in input row order (+ dict of activations)."""
    sd = {k: (v.detach().float() if v.is_floating_point() else v) for k, v in sd.items()}
    C = coords.detach().cpu().numpy() if isinstance(coords, torch.Tensor) else np.asarray(coords)
    cm = ops.CoordinateManager(C.astype(np.int32))
    x = feats.detach().float().cpu()
    acts = {}
    with torch.no_grad():
        img = image_encoder(sd, image.detach().float().cpu())                       # :166
        acts["image"] = img

        out = _conv(cm, x, sd["conv1.kernel"], 1, 1, conv1_kernel_size)   # :168
        out = _bn(out, sd, "norm1.bn")                                               # :169
        out_s1 = _block(sd, cm, 1, out, "block1")                                    # :170 (relu :171 idempotent)
        acts["out_s1"] = out_s1

        cm.stride(1, 2)
        out = _conv(cm, out_s1, sd["conv2.kernel"], 1, 2, 3)  # :173
        out_s2 = _block(sd, cm, 2, _bn(out, sd, "norm2.bn"), "block2")               # :174-175
        acts["out_s2"] = out_s2

        cm.stride(2, 2)
        out = _conv(cm, out_s2, sd["conv3.kernel"], 2, 4, 3)  # :178
        out_s4 = _block(sd, cm, 4, _bn(out, sd, "norm3.bn"), "block3")               # :179-180
        acts["out_s4"] = out_s4

        cm.stride(4, 2)
        out = _conv(cm, out_s4, sd["conv4.kernel"], 4, 8, 3)  # :183
        out_s8 = _block(sd, cm, 8, _bn(out, sd, "norm4.bn"), "block4")               # :184-185
        acts["out_s8"] = out_s8

        fused = transformer(sd, img, out_s8, cm.get(8).C)                            # :189
        acts["fused"] = fused

        out = _conv(cm, fused, sd["conv4_tr.kernel"], 8, 4, 3, True)  # :191
        out = _block(sd, cm, 4, _bn(out, sd, "norm4_tr.bn"), "block4_tr")            # :192-194
        acts["out_s4_tr"] = out
        out = torch.cat([out, out_s4], dim=1)                                        # :197

        out = _conv(cm, out, sd["conv3_tr.kernel"], 4, 2, 3, True)  # :202
        out = _block(sd, cm, 2, _bn(out, sd, "norm3_tr.bn"), "block3_tr")            # :203-205
        acts["out_s2_tr"] = out
        out = torch.cat([out, out_s2], dim=1)                                        # :208

        out = _conv(cm, out, sd["conv2_tr.kernel"], 2, 1, 3, True)  # :213
        out = _block(sd, cm, 1, _bn(out, sd, "norm2_tr.bn"), "block2_tr")            # :214-216
        acts["out_s1_tr"] = out
        out = torch.cat([out, out_s1], dim=1)                                        # :219

        out = torch.relu(out @ sd["conv1_tr.kernel"])                                # :224-225 (1x1 conv = matmul)
        out = out @ sd["final.kernel"] + sd["final.bias"]                            # :226
        acts["final"] = out
        if normalize_feature:
            out = out / torch.norm(out, p=2, dim=1, keepdim=True)                    # :228-233 (no eps)
    acts["levels"] = {t: cm.get(t).C for t in (1, 2, 4, 8)}
    acts["cm"] = cm
    return (out, acts) if return_intermediates else out
