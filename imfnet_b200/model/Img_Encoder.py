"""ImageEncoder().forward(x): [B,3,H,W] -> [B,128,H/8,W/8]  (/root/reference/model/Img_Encoder.py:9-18).

The module keeps the reference's parameter tree (`backbone.*` = the torchvision-style ResNet-34 key set, so checkpoints load
strictly); the arithmetic of the executed prefix (conv7x7/2 -> BN -> ReLU -> maxpool -> layer1 -> layer2, model/resnet.py:195-216)
runs on the sm_100a kernels: pixel-major fp16 hi/lo ("h2") matrices through the tensor-core sparse-convolution kernel with
closed-form neighbour tables (csrc/image_ops.cu), BatchNorm folded into the epilogues.  cuDNN's fp32 path took 1.65 ms for a 640x480
frame on B200 (it was the longest item of the forward); TF32 cuDNN is not an option because single-pass TF32 breaks the 1e-4 parity.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .. import _lib
from . import resnet


def _fold_bn(bn: nn.BatchNorm2d):
    scale = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
    return scale.contiguous(), (bn.bias.detach() - bn.running_mean * scale).contiguous()


class _Conv:
    """One packed convolution: weights W[K, Cin, Cout] in the tensor-core layout + folded BatchNorm.  act_scale: the branch's activation
    scale A (engine.act_scale_from_bn; stored = true * A): the shift carries it; `first` (a layer reading unscaled input) also the scale."""

    def __init__(self, W: torch.Tensor, bn: nn.BatchNorm2d, kc_in: int, act_scale: float = 1.0, first: bool = False):
        L = _lib.lib()
        K, cin, cout = W.shape
        W = W.contiguous().float()
        wmax = float(W.abs().max())
        wmul = 2.0 ** math.floor(math.log2(2048.0 / wmax)) if wmax > 0 else 1.0
        self.K, self.cin, self.cout, self.kc_in = K, cin, cout, kc_in
        self.packed = torch.empty(int(L.imf_sparse_conv_h2_packed_bytes(K, cin, cout, kc_in)), dtype=torch.uint8, device=W.device)
        _lib.check(L.imf_sparse_conv_h2_pack(W.data_ptr(), K, cin, cout, kc_in, wmul, self.packed.data_ptr(), _lib.cur_stream()))
        scale, shift = _fold_bn(bn)
        self.shift = (shift * act_scale).contiguous()
        self.scale = (scale * (act_scale if first else 1.0) / wmul).contiguous()


def _w3(conv: nn.Conv2d) -> torch.Tensor:
    """torch [Cout, Cin, kh, kw] -> [kh*kw, Cin, Cout] with tap index kx + kw*ky (csrc/image_ops.cu::k_image_conv_table)."""
    w = conv.weight.detach()
    return w.permute(2, 3, 1, 0).reshape(w.shape[2] * w.shape[3], w.shape[1], w.shape[0])


class ImagePlan:
    """Static launch sequence of the encoder for one image size (tables and buffers are built once)."""

    STEM_K = 160      # 7*7*3 = 147 im2col columns padded to a multiple of 32

    def __init__(self, backbone: resnet.ResNet, H: int, W: int, split_small: bool = False, err: torch.Tensor | None = None):
        """err: device int32 status word the convolutions report fp16-range overflows / watchdogs into -- the owning forward plan's, so
        that a failure in the image branch raises like one in the point branch (the encoder's own plans use a private word that
        FusedPlan.run merges)."""
        L = _lib.lib()
        self.split_small = split_small        # see FusedPlan.split_small (engine.py)
        p = next(backbone.parameters())
        _lib.require_cuda(p, "image-encoder weights")
        dev = self.device = p.device
        self.H, self.W = H, W
        c1 = backbone.conv1
        k, s, pd = c1.kernel_size[0], c1.stride[0], c1.padding[0]
        self.stem_geom = (k, s, pd)
        self.H1, self.W1 = (H + 2 * pd - k) // s + 1, (W + 2 * pd - k) // s + 1
        self.H2, self.W2 = (self.H1 + 2 - 3) // 2 + 1, (self.W1 + 2 - 3) // 2 + 1          # maxpool 3x3/2, padding 1
        self.H3, self.W3 = (self.H2 + 2 - 3) // 2 + 1, (self.W2 + 2 - 3) // 2 + 1          # layer2 stride 2
        self.P0, self.P1, self.P2 = self.H1 * self.W1, self.H2 * self.W2, self.H3 * self.W3
        if c1.in_channels * k * k > self.STEM_K:
            raise NotImplementedError("stem convolution larger than 7x7x3")
        with torch.cuda.device(dev):
            st = _lib.cur_stream()

            def table(Hin, Win, ks, stride, pad, n_out):
                ld_n = (n_out + 127) // 128 * 128
                nbr_t = torch.empty((ks * ks, ld_n), dtype=torch.int32, device=dev)
                mask = torch.empty(ld_n // 128 + 1, dtype=torch.int32, device=dev)
                _lib.check(L.imf_image_conv_table(Hin, Win, ks, stride, pad, nbr_t.data_ptr(), ld_n, mask.data_ptr(), st))
                return nbr_t, ld_n, mask

            self.t_id0 = table(self.H1, self.W1, 1, 1, 0, self.P0)                 # identity: the stem is a GEMM on the im2col matrix
            self.t1 = table(self.H2, self.W2, 3, 1, 1, self.P1)
            self.t12 = table(self.H2, self.W2, 3, 2, 1, self.P2)
            self.t12d = table(self.H2, self.W2, 1, 2, 0, self.P2)
            self.t2 = table(self.H3, self.W3, 3, 1, 1, self.P2)
            # weights
            w0 = torch.zeros((1, self.STEM_K, c1.out_channels), dtype=torch.float32, device=dev)
            w0[0, : c1.in_channels * k * k] = c1.weight.detach().permute(2, 3, 1, 0).reshape(-1, c1.out_channels)
            from ..engine import act_scale_from_bn
            bns = [backbone.bn1] + [bn for b in list(backbone.layer1) + list(backbone.layer2) for bn in (b.bn1, b.bn2)]
            A = self.act_scale = act_scale_from_bn(bns)          # stored activations of this branch = true * A (power of two)
            self.stem = _Conv(w0, backbone.bn1, 32, A, first=True)
            # the ResNet stem proper (3 channels, 7x7 / 2 / 3 -> 64) as a fused implicit GEMM (csrc/stem_fused.cu): kernel laid out as
            # [pair of kernel rows][2 x (8 columns kx = -1..6 x 4 channels)][Cout] with zeros at kx = -1 and channel 3; other stems keep the im2col route
            import os
            self.stem_fused = None
            if (k, s, pd) == (7, 2, 3) and c1.in_channels == 3 and c1.out_channels == 64 and os.environ.get("IMFNET_B200_STEM", "fused") != "im2col":
                w7 = torch.zeros((8, 8, 4, 64), dtype=torch.float32, device=dev)          # (an 8th, all-zero kernel row completes the last pair)
                w7[:7, 1:, :3, :] = c1.weight.detach().permute(2, 3, 1, 0)         # [ky, kx, c, o]
                self.stem_fused = _Conv(w7.reshape(4, 64, 64), backbone.bn1, 64, A, first=True)      # slabs of two kernel rows
            self.blocks1 = [(_Conv(_w3(b.conv1), b.bn1, 64, A), _Conv(_w3(b.conv2), b.bn2, 64, A)) for b in backbone.layer1]
            self.blocks2 = []
            for b in backbone.layer2:
                down = None if b.downsample is None else _Conv(_w3(b.downsample[0]), b.downsample[1], 64, A)
                self.blocks2.append((_Conv(_w3(b.conv1), b.bn1, 64, A), _Conv(_w3(b.conv2), b.bn2, 64, A), down))
            self.C1, self.C2 = backbone.layer1[0].conv1.out_channels, backbone.layer2[0].conv1.out_channels
            if self.C1 % 64 or self.C2 % 64 or backbone.layer1[0].downsample is not None or self.blocks2[0][2] is None:
                raise NotImplementedError("unexpected ResNet prefix shape")
            f32 = dict(dtype=torch.float32, device=dev)
            self._alloc_stem(1)
            self.s0 = torch.zeros((self.P0, self.C1), **f32)
            # layer1 (3x3 / 1, 64 -> 64) on the plane-layout implicit GEMM (csrc/image_conv_p8.cu); IMFNET_B200_LAYER1=g4 keeps the gather route
            import os
            self.layer1_p8 = self.C1 == 64 and all(c.K == 9 for pair in self.blocks1 for c in pair) and os.environ.get("IMFNET_B200_LAYER1", "p8") != "g4"
            self._alloc_layer1(1)
            self.l2 = [torch.zeros((self.P2, self.C2), **f32) for _ in range(3)]
            self.tokens = torch.zeros((self.P2, self.C2), **f32)
            self.ws_bytes = int(L.imf_sparse_conv_g4_workspace_bytes(self.C2))
            self.ws = torch.zeros(self.ws_bytes, dtype=torch.uint8, device=dev)        # head = arrival counters (zero)
            self.err = err if err is not None else torch.zeros(1, dtype=torch.int32, device=dev)

    def _alloc_stem(self, B: int):
        """Scratch of the stem for B images: the pre-split image set (fused stem) or the im2col matrix."""
        L = _lib.lib()
        self.col = self.stem_ws = None
        if self.stem_fused is not None:
            self.stem_ws_bytes = int(L.imf_image_stem_workspace_bytes(self.H, self.W, B))
            self.stem_ws = torch.zeros(self.stem_ws_bytes, dtype=torch.uint8, device=self.device)
        else:
            self.col = torch.zeros((B * self.P0, self.STEM_K), dtype=torch.float32, device=self.device)   # h2 footprint = fp32 [n, C]

    def _alloc_layer1(self, B: int):
        """Activations of layer1 for B images: three P8 plane buffers (zero borders) + the pixel-major output layer2 reads, or three
        pixel-major h2 matrices (gather route)."""
        L = _lib.lib()
        f32 = dict(dtype=torch.float32, device=self.device)
        if self.layer1_p8:
            nbytes = int(L.imf_image_p8_bytes(self.H2, self.W2, B))
            self.l1p = [torch.zeros(nbytes, dtype=torch.uint8, device=self.device) for _ in range(3)]
            self.l1 = [torch.zeros((B * self.P1, self.C1), **f32)]
        else:
            self.l1 = [torch.zeros((B * self.P1, self.C1), **f32) for _ in range(3)]

    def _layer1(self, L, B: int, s):
        """maxpool + layer1 on self.s0 -> pixel-major h2 matrix [B * P1, C1] (returned)."""
        n1 = B * self.P1
        if not self.layer1_p8:
            _lib.check(L.imf_image_maxpool_h2_batch(self.s0.data_ptr(), 2 * self.C1, 64, self.C1, self.H1, self.W1, 3, 2, 1,
                                                    self.l1[0].data_ptr(), 2 * self.C1, B, s))
            x, tmp, out = self.l1
            for c1, c2 in self.blocks1:
                self._conv(L, c1, x, self.t1, n1, None, True, tmp, s)
                self._conv(L, c2, tmp, self.t1, n1, x, True, out, s)
                x, out = out, x
            return x
        _lib.check(L.imf_image_maxpool_p8(self.s0.data_ptr(), 2 * self.C1, 64, self.H1, self.W1, 3, 2, 1, self.l1p[0].data_ptr(), B, s))
        x, tmp, out = self.l1p
        err = self.err.data_ptr()
        for i, (c1, c2) in enumerate(self.blocks1):
            last = i == len(self.blocks1) - 1
            _lib.check(L.imf_image_conv3x3_p8_fwd(x.data_ptr(), self.H2, self.W2, B, c1.packed.data_ptr(), c1.scale.data_ptr(), c1.shift.data_ptr(),
                                                  None, 1, tmp.data_ptr(), 0, 0, err, s))
            dst = self.l1[0] if last else out          # the last block writes the pixel-major matrix the strided convolutions of layer2 gather from
            _lib.check(L.imf_image_conv3x3_p8_fwd(tmp.data_ptr(), self.H2, self.W2, B, c2.packed.data_ptr(), c2.scale.data_ptr(), c2.shift.data_ptr(),
                                                  x.data_ptr(), 1, dst.data_ptr(), 1 if last else 0, 2 * self.C1, err, s))
            x, out = out, x
        return self.l1[0]

    def _stem(self, L, images, B: int, s):
        """conv1 -> bn1 -> relu of the ResNet prefix for B images -> self.s0 (h2, pixel-major)."""
        if self.stem_fused is not None:
            c = self.stem_fused
            _lib.check(L.imf_image_stem_h2_fwd(images.data_ptr(), self.H, self.W, B, c.packed.data_ptr(), c.scale.data_ptr(), c.shift.data_ptr(),
                                               self.stem_ws.data_ptr(), self.stem_ws_bytes, self.s0.data_ptr(), 2 * self.C1, self.err.data_ptr(), s))
            return
        k, st, pd = self.stem_geom
        _lib.check(L.imf_image_im2col_h2_batch(images.data_ptr(), 3, self.H, self.W, k, st, pd, self.STEM_K, self.col.data_ptr(),
                                               2 * self.STEM_K, B, s))                               # all images, one launch
        self._conv(L, self.stem, self.col, self.t_id0, B * self.P0, None, True, self.s0, s)

    def _conv(self, L, c: _Conv, X, tab, n_out, R, relu, Y, s):
        nbr_t, ld_n, mask = tab
        split = self.split_small and n_out < 128 * _lib.sm_count()
        _lib.check(L.imf_sparse_conv_g4_fwd(X.data_ptr(), 2 * c.cin, c.kc_in, c.packed.data_ptr(), nbr_t.data_ptr(), ld_n, mask.data_ptr(),
                                            None, n_out, c.K, c.cin, c.cout, c.scale.data_ptr(), c.shift.data_ptr(), _lib.ptr(R),
                                            0 if R is None else 2 * c.cout, 64, 1 if relu else 0, Y.data_ptr(), 2 * c.cout, n_out, 64,
                                            self.ws.data_ptr() if split else None, self.ws_bytes if split else 0,
                                            self.err.data_ptr(), s))

    def enqueue(self, image: torch.Tensor) -> torch.Tensor:
        """image fp32 [3,H,W] (contiguous, on the plan's device) -> fp32 tokens [H/8*W/8, 128] (a buffer owned by the plan)."""
        L = _lib.lib()
        s = _lib.cur_stream()
        self._stem(L, image, 1, s)
        x = self._layer1(L, 1, s)
        y, tmp, out = self.l2
        for i, (c1, c2, down) in enumerate(self.blocks2):
            if down is not None:          # first block: stride 2, 1x1 projection of the skip path
                self._conv(L, c1, x, self.t12, self.P2, None, True, tmp, s)
                self._conv(L, down, x, self.t12d, self.P2, None, False, out, s)
                self._conv(L, c2, tmp, self.t2, self.P2, out, True, y, s)
            else:
                self._conv(L, c1, y, self.t2, self.P2, None, True, tmp, s)
                self._conv(L, c2, tmp, self.t2, self.P2, y, True, out, s)
                y, out = out, y
        _lib.check(L.imf_h2_unpack_scaled_n(y.data_ptr(), 2 * self.C2, self.P2, None, self.C2, 64, 1.0 / self.act_scale, self.tokens.data_ptr(),
                                            self.C2, s))
        return self.tokens


class ImageEncoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.backbone = resnet.resnet34(in_channels=3, pretrained=False, progress=False)
        self.low_latency = False              # set by the owning model's plan (ResUNet2.low_latency)
        self._plans = {}

    def _apply(self, fn, *a, **k):
        self._plans = {}
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._plans = {}
        return super().load_state_dict(*a, **k)

    def _weights_key(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.backbone.parameters()) + list(self.backbone.buffers()))

    def plan(self, H: int, W: int) -> ImagePlan:
        key = self._weights_key()
        key = key + (self.low_latency,)
        hit = self._plans.get((H, W))
        if hit is None or hit[0] != key:
            if self.training:
                raise NotImplementedError("imfnet_b200 implements the eval-mode image encoder (BatchNorm running statistics)")
            hit = self._plans[(H, W)] = (key, ImagePlan(self.backbone, H, W, self.low_latency))
        return hit[1]

    def tokens(self, image: torch.Tensor) -> torch.Tensor:
        """image [3,H,W] -> fp32 [H/8*W/8, 128] pixel-major feature tokens (plan-owned buffer, valid until the next call)."""
        image = image.float().contiguous()
        _lib.require_cuda(image, "image")
        with torch.cuda.device(image.device):
            return self.plan(int(image.shape[1]), int(image.shape[2])).enqueue(image)

    def forward(self, x):
        if torch.is_grad_enabled() and x.requires_grad:
            raise NotImplementedError("imfnet_b200 implements the inference forward only (no autograd)")
        L = _lib.lib()
        outs = []
        for b in range(x.shape[0]):
            tok = self.tokens(x[b])
            plan = self.plan(int(x.shape[2]), int(x.shape[3]))
            o = torch.empty((plan.C2, plan.H3, plan.W3), dtype=torch.float32, device=tok.device)
            with torch.cuda.device(tok.device):
                _lib.check(L.imf_transpose_tokens(tok.data_ptr(), plan.P2, plan.C2, o.data_ptr(), _lib.cur_stream()))
            outs.append(o)
        return torch.stack(outs, dim=0)
