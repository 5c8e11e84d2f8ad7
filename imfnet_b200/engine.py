"""Fused execution plan of ResUNet2.forward (/root/reference/model/resunet.py:163-235) on one B200.

What the reference runs as ~60 MinkowskiEngine / torch module calls becomes:
  * side stream : image encoder (torch/cuDNN, fp32) -> LayerNorm_c + to_kv projection of the image tokens;
  * main stream : coordinate pyramid (1 host read-back of the level sizes) -> neighbour tables ->
                  23 convolution launches with BatchNorm / ReLU / residual folded into their epilogues ->
                  attention fusion at stride 8 -> decoder -> fused 1x1 tail (conv1_tr, ReLU, final, L2 norm).
Channel concatenations (ME.cat, lines 197/208/219) are column windows of one buffer: the producer of each
half writes straight into it, so no concat kernel and no extra copy exists.
"""
from __future__ import annotations

import torch

from . import _lib
from .arena import Arena


class FusedPlan:
    def __init__(self, model):
        self.m = model
        p = next(model.parameters())
        _lib.require_cuda(p, "model parameters")
        self.device = p.device
        self.side = torch.cuda.Stream(device=self.device)
        self.CH, self.TR = model.CHANNELS, model.TR_CHANNELS
        self._key = None
        self.debug = None      # set to a dict to capture intermediate activations (tests only)
        self.conv_impl = "tc"  # "tc" = tcgen05 3xTF32 implicit GEMM; "simt" = fp32 SIMT tier
        self.arena = Arena(self.device)
        self.pack()

    # -- weights -------------------------------------------------------------------------------
    def _weights_key(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.m.parameters()) + list(self.m.buffers()))

    def pack(self):
        m = self.m
        for t in m.parameters():
            if t.dtype != torch.float32:
                raise TypeError("imfnet_b200 computes in float32; cast the model with .float()")
        self.bn = {}
        for name in ("norm1", "norm2", "norm3", "norm4", "norm4_tr", "norm3_tr", "norm2_tr"):
            self.bn[name] = getattr(m, name).folded()
        for b in ("block1", "block2", "block3", "block4", "block4_tr", "block3_tr", "block2_tr"):
            blk = getattr(m, b)
            self.bn[b + ".norm1"] = blk.norm1.folded()
            self.bn[b + ".norm2"] = blk.norm2.folded()
        self.final_bias = None if m.final.bias is None else m.final.bias.detach().reshape(-1).contiguous()
        # tensor-core weight slabs (hi/lo TF32 split, swizzled) for every 3x3x3 convolution
        L = _lib.lib()
        self.packed = {}
        with torch.cuda.device(self.device):
            s = torch.cuda.current_stream().cuda_stream
            for name, mod in m.named_modules():
                if hasattr(mod, "kernel") and getattr(mod, "kernel_volume", 1) == 27 and mod.in_channels % 32 == 0:
                    nbytes = int(L.imf_sparse_conv_tc_packed_bytes(27, mod.in_channels, mod.out_channels))
                    buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
                    _lib.check(L.imf_sparse_conv_tc_pack(mod.kernel.detach().contiguous().data_ptr(), 27, mod.in_channels,
                                                         mod.out_channels, buf.data_ptr(), s))
                    self.packed[id(mod)] = buf
        self.err = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._key = self._weights_key()

    # -- launch helpers ------------------------------------------------------------------------
    def _conv(self, L, X, ldx, conv, nbr, n_out, bn, R, ldr, relu, Y, ldy, stream):
        scale, shift = (None, None) if bn is None else self.bn[bn]
        if self.conv_impl == "simt":          # fp32 SIMT tier (kept for A/B measurements)
            _lib.check(L.imf_sparse_conv_fwd(X, ldx, conv.kernel.data_ptr(), nbr.data_ptr(), None, n_out, conv.kernel_volume,
                                             conv.in_channels, conv.out_channels, _lib.ptr(scale), _lib.ptr(shift), R, ldr,
                                             1 if relu else 0, Y, ldy, stream))
            return
        ws, ws_bytes = None, 0
        if n_out < 12800:                     # few row tiles: let the kernel split a tile's offsets over several CTAs
            ws_bytes = int(L.imf_sparse_conv_tc_workspace_bytes(n_out, conv.out_channels))
            ws = self.arena.take(ws_bytes)
        _lib.check(L.imf_sparse_conv_tc_fwd(X, ldx, self.packed[id(conv)].data_ptr(), nbr.data_ptr(), None, n_out,
                                            conv.kernel_volume, conv.in_channels, conv.out_channels, _lib.ptr(scale),
                                            _lib.ptr(shift), R, ldr, 1 if relu else 0, Y, ldy, _lib.ptr(ws), ws_bytes,
                                            self.err.data_ptr(), stream))

    def _block(self, L, name, X, ldx, nbr, n, C, tmp, Y, ldy, stream):
        """BasicBlockBN (model/residual_block.py:37-53): X -> tmp = relu(bn1(conv1 X)) -> Y = relu(bn2(conv2 tmp) + X)."""
        blk = getattr(self.m, name)
        self._conv(L, X, ldx, blk.conv1, nbr, n, name + ".norm1", None, 0, True, tmp.data_ptr(), C, stream)
        self._conv(L, tmp.data_ptr(), C, blk.conv2, nbr, n, name + ".norm2", X, ldx, True, Y, ldy, stream)

    # -- forward -------------------------------------------------------------------------------
    @torch.no_grad()
    def run(self, x, image: torch.Tensor) -> torch.Tensor:
        m, dev = self.m, self.device
        if self._key != self._weights_key():
            self.pack()
        L = _lib.lib()
        CH, TR = self.CH, self.TR
        F0 = x.F
        _lib.require_cuda(F0, "input features")
        if F0.device != dev:
            raise RuntimeError(f"input is on {F0.device} but the model is on {dev}")
        F0 = F0.float().contiguous()
        image = image.to(device=dev, dtype=torch.float32)
        cm = x.coordinate_manager
        t0 = x.coordinate_map_key.tensor_stride
        if t0 != 1:
            raise NotImplementedError("the fused plan expects an input at tensor stride 1")
        N0 = len(F0)
        f4 = 4  # bytes per float

        with torch.cuda.device(dev):
            main = torch.cuda.current_stream()
            s = main.cuda_stream
            # ---- side stream: image branch (independent of the point branch until the fusion) ----
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                img = m.img_encoder(image)                                   # [B,128,H/8,W/8]   resunet.py:166
                B, Ci, Hp, Wp = img.shape
                kvs = [m.attention_fusion.project_context(img[b].reshape(Ci, Hp * Wp), True) for b in range(B)]
                ev_img = torch.cuda.Event()
                ev_img.record(self.side)

            # ---- coordinates ----
            cm.build_pyramid([2, 4, 8])
            lv = {t: cm.level(t) for t in (1, 2, 4, 8)}
            n1, n2, n4, n8 = lv[1].n, lv[2].n, lv[4].n, lv[8].n
            nb = {t: cm.table(t, t, 3, False) for t in (1, 2, 4, 8)}
            dn = {(1, 2): cm.table(1, 2, 3, False), (2, 4): cm.table(2, 4, 3, False), (4, 8): cm.table(4, 8, 3, False)}
            up = {(8, 4): cm.table(8, 4, 3, True), (4, 2): cm.table(4, 2, 3, True), (2, 1): cm.table(2, 1, 3, True)}

            self.arena.reset()
            buf = self.arena.floats

            # concat buffers: [decoder half | encoder skip half]
            cat1, cat2, cat4 = buf(n1, TR[2] + CH[1]), buf(n2, TR[3] + CH[2]), buf(n4, TR[4] + CH[3])
            ld1, ld2, ld4 = cat1.shape[1], cat2.shape[1], cat4.shape[1]
            s1_ptr, s2_ptr, s4_ptr = cat1.data_ptr() + TR[2] * f4, cat2.data_ptr() + TR[3] * f4, cat4.data_ptr() + TR[4] * f4

            # ---- encoder ----
            a0, a1 = buf(n1, CH[1]), buf(n1, CH[1])
            sc, sh = self.bn["norm1"]
            _lib.check(L.imf_conv_first_fwd(F0.data_ptr(), F0.shape[1], m.conv1.in_channels, m.conv1.kernel.data_ptr(),
                                            lv[1].coords.data_ptr(), None, n1, lv[1].table.data_ptr(), lv[1].capacity,
                                            m.conv1.kernel_size, 1, CH[1], sc.data_ptr(), sh.data_ptr(), 0, a0.data_ptr(),
                                            CH[1], s))                                         # conv1 + norm1   :168-169
            self._block(L, "block1", a0.data_ptr(), CH[1], nb[1], n1, CH[1], a1, s1_ptr, ld1, s)   # out_s1 -> cat1[:, TR2:]

            b0, b1 = buf(n2, CH[2]), buf(n2, CH[2])
            self._conv(L, s1_ptr, ld1, m.conv2, dn[(1, 2)], n2, "norm2", None, 0, False, b0.data_ptr(), CH[2], s)   # :173-174
            self._block(L, "block2", b0.data_ptr(), CH[2], nb[2], n2, CH[2], b1, s2_ptr, ld2, s)   # out_s2 -> cat2[:, TR3:]

            c0, c1 = buf(n4, CH[3]), buf(n4, CH[3])
            self._conv(L, s2_ptr, ld2, m.conv3, dn[(2, 4)], n4, "norm3", None, 0, False, c0.data_ptr(), CH[3], s)   # :178-179
            self._block(L, "block3", c0.data_ptr(), CH[3], nb[4], n4, CH[3], c1, s4_ptr, ld4, s)   # out_s4 -> cat4[:, TR4:]

            d0, d1, d2 = buf(n8, CH[4]), buf(n8, CH[4]), buf(n8, CH[4])
            self._conv(L, s4_ptr, ld4, m.conv4, dn[(4, 8)], n8, "norm4", None, 0, False, d0.data_ptr(), CH[4], s)   # :183-184
            self._block(L, "block4", d0.data_ptr(), CH[4], nb[8], n8, CH[4], d1, d2.data_ptr(), CH[4], s)          # out_s8

            # ---- attention fusion at stride 8 (resunet.py:189, 237-273) ----
            main.wait_event(ev_img)
            fused = buf(n8, CH[4])
            seg = cm.batch_segments(8, B) if B > 1 else [0, n8]
            for b in range(B):
                lo, hi = seg[b], seg[b + 1]
                if hi > lo:
                    m.attention_fusion.fuse(d2[lo:hi], kvs[b], out=fused[lo:hi], arena=self.arena)
            if seg[B] != n8:
                raise ValueError("coordinates reference more batch items than images were given")
            if self.debug is not None:
                self.debug.update(image=img.clone(), out_s1=cat1[:, TR[2]:].clone(), out_s2=cat2[:, TR[3]:].clone(),
                                  out_s4=cat4[:, TR[4]:].clone(), out_s8=d2.clone(), fused=fused.clone(), conv1=a0.clone(),
                                  levels={t: lv[t].coords.clone() for t in (1, 2, 4, 8)})
            for kv in kvs:
                kv.record_stream(main)
            img.record_stream(main)

            # ---- decoder ----
            e0, e1 = buf(n4, TR[4]), buf(n4, TR[4])
            self._conv(L, fused.data_ptr(), CH[4], m.conv4_tr, up[(8, 4)], n4, "norm4_tr", None, 0, False, e0.data_ptr(), TR[4], s)
            self._block(L, "block4_tr", e0.data_ptr(), TR[4], nb[4], n4, TR[4], e1, cat4.data_ptr(), ld4, s)   # -> cat4[:, :TR4]

            g0, g1 = buf(n2, TR[3]), buf(n2, TR[3])
            self._conv(L, cat4.data_ptr(), ld4, m.conv3_tr, up[(4, 2)], n2, "norm3_tr", None, 0, False, g0.data_ptr(), TR[3], s)
            self._block(L, "block3_tr", g0.data_ptr(), TR[3], nb[2], n2, TR[3], g1, cat2.data_ptr(), ld2, s)   # -> cat2[:, :TR3]

            h0, h1 = buf(n1, TR[2]), buf(n1, TR[2])
            self._conv(L, cat2.data_ptr(), ld2, m.conv2_tr, up[(2, 1)], n1, "norm2_tr", None, 0, False, h0.data_ptr(), TR[2], s)
            self._block(L, "block2_tr", h0.data_ptr(), TR[2], nb[1], n1, TR[2], h1, cat1.data_ptr(), ld1, s)   # -> cat1[:, :TR2]

            # ---- tail: conv1_tr -> ReLU -> final(+bias) -> L2 norm (resunet.py:224-233) ----
            out = torch.empty((n1, m.out_channels), dtype=torch.float32, device=dev)
            _lib.check(L.imf_pointwise_tail_fwd(cat1.data_ptr(), ld1, ld1, m.conv1_tr.kernel.data_ptr(), TR[1],
                                                m.final.kernel.data_ptr(), _lib.ptr(self.final_bias), m.out_channels, None,
                                                n1, 1 if m.normalize_feature else 0, out.data_ptr(), m.out_channels, s))
            if self.debug is not None:
                self.debug.update(out_s4_tr=cat4[:, :TR[4]].clone(), out_s2_tr=cat2[:, :TR[3]].clone(),
                                  out_s1_tr=cat1[:, :TR[2]].clone())
        return out
