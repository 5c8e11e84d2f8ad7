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

import math

import torch

from . import _lib
from .arena import Arena


def act_scale_from_bn(bns) -> float:
    """Power-of-two scale A under which the h2 activations of a branch are STORED (stored = true * A), chosen so that typical stored
    magnitudes are O(1): the fp16 hi/lo format keeps 22 mantissa bits only for values between its subnormal range (~6e-5) and its
    overflow (6e4), whereas the reference computes in fp32.  After an eval-mode BatchNorm an activation is gamma * N(0, 1) + beta, so the
    geometric mean over the layers of rms_c sqrt(gamma^2 + beta^2) is the typical magnitude; a checkpoint whose BatchNorm parameters
    sit near 1e-6 or 1e4 is brought back into range.  Because A is a power of two and enters only through the folded BatchNorm shift
    (and the conversions at the branch's borders), A = 1 reproduces the unscaled arithmetic bit for bit."""
    logs = []
    for bn in bns:
        g = bn.weight.detach().double()
        b = bn.bias.detach().double()
        m = float(torch.sqrt((g * g + b * b).mean()))
        if m > 0 and math.isfinite(m):
            logs.append(math.log2(m))
    if not logs:
        return 1.0
    e = -round(sum(logs) / len(logs))
    return float(2.0 ** max(-100, min(100, e)))


def _kc(*channels) -> int:
    """Chunk width of an h2 buffer section holding these channel counts (see csrc/sparse_conv_h2.cu)."""
    return 64 if all(c % 64 == 0 for c in channels) else 32


class FusedPlan:
    # (convolution, the BatchNorm folded into its epilogue) for every 3x3x3 layer of the graph
    CONV_BN = [("conv2", "norm2"), ("conv3", "norm3"), ("conv4", "norm4"), ("conv4_tr", "norm4_tr"), ("conv3_tr", "norm3_tr"),
               ("conv2_tr", "norm2_tr")] + [(f"{b}.conv{i}", f"{b}.norm{i}") for b in
                                            ("block1", "block2", "block3", "block4", "block4_tr", "block3_tr", "block2_tr")
                                            for i in (1, 2)]

    def __init__(self, model):
        self.m = model
        p = next(model.parameters())
        _lib.require_cuda(p, "model parameters")
        self.device = p.device
        self.side = torch.cuda.Stream(device=self.device)
        self.CH, self.TR = model.CHANNELS, model.TR_CHANNELS
        self._key = None
        self.debug = None      # set to a dict to capture intermediate activations (tests only)
        self.arena = Arena(self.device)
        # small levels (fewer 128-row tiles than SMs): splitting a tile's offsets over several CTAs + a reduce launch shortens ONE
        # fragment's latency (1.79 vs 2.05 ms at C2) but costs SM time (every split CTA pays the kernel's fixed cost), i.e.
        # throughput when several fragments are in flight (43.6 -> 48.2 M voxels/s); off unless the model asks for low latency
        self.split_small = bool(getattr(model, "low_latency", False))
        model.img_encoder.low_latency = self.split_small
        self.pack()

    # -- weights -------------------------------------------------------------------------------
    def _weights_key(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.m.parameters()) + list(self.m.buffers()))

    def pack(self):
        """Fold every BatchNorm into (scale, shift), split + swizzle the 3x3x3 kernels for the fp16 hi/lo tensor-core tier."""
        m = self.m
        for t in m.parameters():
            if t.dtype != torch.float32:
                raise TypeError("imfnet_b200 computes in float32; cast the model with .float()")
        mods = dict(m.named_modules())
        CH, TR = self.CH, self.TR
        # activation scale of the point branch (see act_scale_from_bn); un-normalised outputs would carry it, so it stays 1 then
        self.act_scale = act_scale_from_bn([mods[b].bn for _c, b in self.CONV_BN] + [m.norm1.bn]) if m.normalize_feature else 1.0
        A = self.act_scale
        # chunk width of the buffer each convolution reads
        kc_in = {"conv2": _kc(CH[1]), "conv3": _kc(TR[3], CH[2]), "conv4": _kc(TR[4], CH[3]), "conv4_tr": _kc(CH[4]),
                 "conv3_tr": _kc(TR[4], CH[3]), "conv2_tr": _kc(TR[3], CH[2])}
        L = _lib.lib()
        self.conv = {}
        with torch.cuda.device(self.device):
            s = torch.cuda.current_stream().cuda_stream
            for cname, bname in self.CONV_BN:
                conv, bn = mods[cname], mods[bname]
                if conv.kernel_volume != 27 or conv.in_channels % 32 or conv.out_channels not in (32, 64, 128, 256):
                    raise NotImplementedError(f"{cname}: unsupported shape for the tensor-core tier")
                kci = kc_in.get(cname, _kc(conv.in_channels))
                wmax = float(conv.kernel.detach().abs().max())
                wmul = 2.0 ** math.floor(math.log2(2048.0 / wmax)) if wmax > 0 else 1.0
                nbytes = int(L.imf_sparse_conv_h2_packed_bytes(27, conv.in_channels, conv.out_channels, kci))
                buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
                _lib.check(L.imf_sparse_conv_h2_pack(conv.kernel.detach().contiguous().data_ptr(), 27, conv.in_channels,
                                                     conv.out_channels, kci, wmul, buf.data_ptr(), s))
                scale, shift = bn.folded()
                self.conv[cname] = (conv, buf, (scale / wmul).contiguous(), (shift * A).contiguous(), kci)          # stored out = A * true out
        sc, sh = m.norm1.folded()
        sc, sh = (sc * A).contiguous(), (sh * A).contiguous()          # conv1 reads the (unscaled) input features
        self.norm1 = (sc, sh)
        # conv1 with one input channel (the IMFNet configuration: a column of ones, util/misc.py:76-77) runs on the tensor cores as a
        # dense product over the K^3 neighbour features (csrc/conv_first_tc.cu): its kernel packed as ONE offset with K^3 "channels"
        self.conv1_tc = None
        if m.conv1.in_channels == 1 and m.conv1.kernel_size in (1, 3, 5) and CH[1] in (32, 64, 128):
            with torch.cuda.device(self.device):
                K3 = m.conv1.kernel_size ** 3
                KP = int(L.imf_conv_first_tc_columns(m.conv1.kernel_size))
                w1 = torch.zeros((1, KP, CH[1]), dtype=torch.float32, device=self.device)
                w1[0, :K3] = m.conv1.kernel.detach().reshape(K3, CH[1])
                wmax = float(w1.abs().max())
                wmul = 2.0 ** math.floor(math.log2(2048.0 / wmax)) if wmax > 0 else 1.0
                buf = torch.empty(int(L.imf_sparse_conv_h2_packed_bytes(1, KP, CH[1], 64)), dtype=torch.uint8, device=self.device)
                _lib.check(L.imf_sparse_conv_h2_pack(w1.data_ptr(), 1, KP, CH[1], 64, wmul, buf.data_ptr(), torch.cuda.current_stream().cuda_stream))
                torch.cuda.current_stream().synchronize()          # w1 is a temporary
                self.conv1_tc = (buf, (sc / wmul).contiguous(), sh)
        self.final_bias = None if m.final.bias is None else (m.final.bias.detach().reshape(-1) * A).contiguous()          # logits are stored * A
        # tail conv1_tr -> ReLU -> final (+ bias) as two one-offset convolutions on the tensor-core kernel (the captured plans then
        # write the concatenation [decoder | skip] with 32-channel chunks throughout, so that it is ONE h2 matrix of chunk width 32)
        self.tail_tc = None
        self.tail_fused = False
        c0, c1, c2 = m.conv1_tr.in_channels, m.conv1_tr.out_channels, m.final.out_channels
        if (m.conv1_tr.kernel_volume == 1 and m.final.kernel_volume == 1 and c0 % 32 == 0 and c1 in (64, 128, 256) and c2 in (32, 64, 128)
                and m.conv1_tr.bias is None and TR[2] % 32 == 0):
            with torch.cuda.device(self.device):
                st = torch.cuda.current_stream().cuda_stream
                packs = []
                # The logits get their own power-of-two scale A_f: `final` has a bias instead of a BatchNorm, so their magnitude is
                # sqrt(|bias|^2 + |W2 . hidden|^2) whatever the branch's A is (a checkpoint with tiny activations and an O(1) bias would
                # overflow under A); the L2 normalisation removes A_f like it removes A.
                b2 = m.final.bias.detach().double().reshape(-1) if m.final.bias is not None else torch.zeros(1, dtype=torch.float64)
                est = math.sqrt(float((b2 * b2).mean()) + float((m.final.kernel.detach().double() ** 2).mean()) * c1 / (A * A))
                self.final_scale = Af = (2.0 ** -round(math.log2(est)) if est > 0 and math.isfinite(est) else A) if m.normalize_feature else 1.0
                for W, cin, cout, kci in ((m.conv1_tr.kernel, c0, c1, 32), (m.final.kernel, c1, c2, 64)):
                    W = W.detach().reshape(1, cin, cout).contiguous()
                    wmax = float(W.abs().max())
                    wmul = 2.0 ** math.floor(math.log2(2048.0 / wmax)) if wmax > 0 else 1.0
                    buf = torch.empty(int(L.imf_sparse_conv_h2_packed_bytes(1, cin, cout, kci)), dtype=torch.uint8, device=self.device)
                    _lib.check(L.imf_sparse_conv_h2_pack(W.data_ptr(), 1, cin, cout, kci, wmul, buf.data_ptr(), st))
                    mul = (Af / A) if len(packs) == 1 else 1.0          # (second product) hidden is stored * A, the logits * A_f
                    packs.append((buf, torch.full((cout,), mul / wmul, dtype=torch.float32, device=self.device)))
                torch.cuda.current_stream().synchronize()
                zero1 = torch.zeros(c1, dtype=torch.float32, device=self.device)
                bias2 = (m.final.bias.detach().reshape(-1) * Af) if m.final.bias is not None else torch.zeros(c2, dtype=torch.float32, device=self.device)
                self.tail_tc = ((packs[0][0], packs[0][1], zero1), (packs[1][0], packs[1][1], bias2.float().contiguous()))
                # ... and, for the IMFNet shapes, as ONE kernel (IMFNET_B200_TAIL=split keeps the two-launch form, for comparison)
                import os
                self.tail_fused = c0 <= 96 and c1 == 64 and c2 == 32 and os.environ.get("IMFNET_B200_TAIL", "fused") != "split"
        self.err = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.err_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        # split-mode workspace of the convolution kernel (its head holds arrival counters that must start, and are left, zero)
        self.conv_ws_bytes = int(L.imf_sparse_conv_g4_workspace_bytes(max(max(CH[1:]), max(TR[1:]))))
        self.conv_ws = torch.zeros(self.conv_ws_bytes, dtype=torch.uint8, device=self.device)
        self._key = self._weights_key()

    # -- launch helpers ------------------------------------------------------------------------
    def _conv(self, L, cname, X, ldx, tab, n_out, R, ldr, kc_r, relu, Y, ldy, kc_out, stream):
        """One 3x3x3 convolution + folded BatchNorm (+ residual) (+ ReLU); X, R, Y are h2 matrices (ld in halves);
        tab = CoordinateManager.table_t(...) (offset-major neighbour table, its row stride, tile masks)."""
        conv, packed, scale, shift, kci = self.conv[cname]
        nbr_t, ld_n, tile_mask = tab
        split = self.split_small and n_out < 128 * _lib.sm_count()      # fewer row tiles than SMs: the kernel may split a tile's offsets over several CTAs
        _lib.check(L.imf_sparse_conv_g4_fwd(X, ldx, kci, packed.data_ptr(), nbr_t.data_ptr(), ld_n, tile_mask.data_ptr(), None, n_out,
                                            27, conv.in_channels, conv.out_channels, scale.data_ptr(), shift.data_ptr(), R, ldr, kc_r,
                                            1 if relu else 0, Y, ldy, n_out, kc_out, self.conv_ws.data_ptr() if split else None,
                                            self.conv_ws_bytes if split else 0, self.err.data_ptr(), stream))

    def _block(self, L, name, X, ldx, kc_x, nbr, n, C, tmp, Y, ldy, kc_y, stream):
        """BasicBlockBN (model/residual_block.py:37-53): X -> tmp = relu(bn1(conv1 X)) -> Y = relu(bn2(conv2 tmp) + X)."""
        kt = _kc(C)
        self._conv(L, name + ".conv1", X, ldx, nbr, n, None, 0, 0, True, tmp.data_ptr(), 2 * C, kt, stream)
        self._conv(L, name + ".conv2", tmp.data_ptr(), 2 * C, nbr, n, X, ldx, kc_x, True, Y, ldy, kc_y, stream)

    def _unpack(self, ptr, ldh, n, C, kc):
        out = torch.empty((n, C), dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib().imf_h2_unpack_scaled_n(ptr, ldh, n, None, C, kc, 1.0 / self.act_scale, out.data_ptr(), C,
                                                     torch.cuda.current_stream().cuda_stream))
        return out

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

        with torch.cuda.device(dev):
            main = torch.cuda.current_stream()
            s = main.cuda_stream
            # ---- side stream: image branch (independent of the point branch until the fusion) ----
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                # image encoder (resunet.py:166) as pixel-major tokens [H/8*W/8, 128] = the `data` layout of the fusion module
                B = image.shape[0]
                img_plan = m.img_encoder.plan(int(image.shape[2]), int(image.shape[3]))
                img_plan.err.zero_()
                kvs = [m.attention_fusion.project_context(m.img_encoder.tokens(image[b]), False) for b in range(B)]
                img = m.img_encoder(image) if self.debug is not None else None
                ev_img = torch.cuda.Event()
                ev_img.record(self.side)

            # ---- coordinates ----
            self.err.zero_()
            cm.build_pyramid([2, 4, 8])
            lv = {t: cm.level(t) for t in (1, 2, 4, 8)}
            n1, n2, n4, n8 = lv[1].n, lv[2].n, lv[4].n, lv[8].n
            nb = {t: cm.table_t(t, t, 3, False) for t in (1, 2, 4, 8)}
            dn = {(1, 2): cm.table_t(1, 2, 3, False), (2, 4): cm.table_t(2, 4, 3, False), (4, 8): cm.table_t(4, 8, 3, False)}
            up = {(8, 4): cm.table_t(8, 4, 3, True), (4, 2): cm.table_t(4, 2, 3, True), (2, 1): cm.table_t(2, 1, 3, True)}

            self.arena.reset()
            buf = self.arena.floats        # an h2 matrix of C channels occupies exactly the bytes of an fp32 [n, C] matrix

            # concat buffers: [decoder half | encoder skip half]; ld in halves; byte offset of the skip half = 4*TR
            cat1, cat2, cat4 = buf(n1, TR[2] + CH[1]), buf(n2, TR[3] + CH[2]), buf(n4, TR[4] + CH[3])
            ld1, ld2, ld4 = 2 * cat1.shape[1], 2 * cat2.shape[1], 2 * cat4.shape[1]
            s1_ptr, s2_ptr, s4_ptr = cat1.data_ptr() + TR[2] * 4, cat2.data_ptr() + TR[3] * 4, cat4.data_ptr() + TR[4] * 4
            kc1a, kc1b = _kc(TR[2]), _kc(CH[1])
            kc2, kc4 = _kc(TR[3], CH[2]), _kc(TR[4], CH[3])

            # ---- encoder ----
            a0, a1 = buf(n1, CH[1]), buf(n1, CH[1])
            sc, sh = self.norm1
            _lib.check(L.imf_conv_first_h2_fwd(F0.data_ptr(), F0.shape[1], m.conv1.in_channels, m.conv1.kernel.data_ptr(),
                                               lv[1].coords.data_ptr(), None, n1, lv[1].table.data_ptr(), lv[1].capacity,
                                               m.conv1.kernel_size, 1, CH[1], sc.data_ptr(), sh.data_ptr(), 0, a0.data_ptr(),
                                               2 * CH[1], _kc(CH[1]), s))                          # conv1 + norm1   :168-169
            self._block(L, "block1", a0.data_ptr(), 2 * CH[1], _kc(CH[1]), nb[1], n1, CH[1], a1, s1_ptr, ld1, kc1b, s)   # out_s1

            b0, b1 = buf(n2, CH[2]), buf(n2, CH[2])
            self._conv(L, "conv2", s1_ptr, ld1, dn[(1, 2)], n2, None, 0, 0, False, b0.data_ptr(), 2 * CH[2], _kc(CH[2]), s)   # :173-174
            self._block(L, "block2", b0.data_ptr(), 2 * CH[2], _kc(CH[2]), nb[2], n2, CH[2], b1, s2_ptr, ld2, kc2, s)        # out_s2

            c0, c1 = buf(n4, CH[3]), buf(n4, CH[3])
            self._conv(L, "conv3", s2_ptr, ld2, dn[(2, 4)], n4, None, 0, 0, False, c0.data_ptr(), 2 * CH[3], _kc(CH[3]), s)   # :178-179
            self._block(L, "block3", c0.data_ptr(), 2 * CH[3], _kc(CH[3]), nb[4], n4, CH[3], c1, s4_ptr, ld4, kc4, s)        # out_s4

            d0, d1, d2 = buf(n8, CH[4]), buf(n8, CH[4]), buf(n8, CH[4])
            k8 = _kc(CH[4])
            self._conv(L, "conv4", s4_ptr, ld4, dn[(4, 8)], n8, None, 0, 0, False, d0.data_ptr(), 2 * CH[4], k8, s)          # :183-184
            self._block(L, "block4", d0.data_ptr(), 2 * CH[4], k8, nb[8], n8, CH[4], d1, d2.data_ptr(), 2 * CH[4], k8, s)    # out_s8

            # ---- attention fusion at stride 8 (resunet.py:189, 237-273), fp32 tokens ----
            P8 = buf(n8, CH[4])
            _lib.check(L.imf_h2_unpack_scaled_n(d2.data_ptr(), 2 * CH[4], n8, None, CH[4], k8, 1.0 / self.act_scale, P8.data_ptr(), CH[4], s))
            main.wait_event(ev_img)
            fused32 = buf(n8, CH[4])
            seg = cm.batch_segments(8, B) if B > 1 else [0, n8]
            for b in range(B):
                lo, hi = seg[b], seg[b + 1]
                if hi > lo:
                    m.attention_fusion.fuse(P8[lo:hi], kvs[b], out=fused32[lo:hi], arena=self.arena)
            if seg[B] != n8:
                raise ValueError("coordinates reference more batch items than images were given")
            fused = buf(n8, CH[4])
            _lib.check(L.imf_h2_pack_scaled_n(fused32.data_ptr(), CH[4], n8, None, CH[4], k8, self.act_scale, fused.data_ptr(), 2 * CH[4],
                                              self.err.data_ptr(), s))
            if self.debug is not None:
                self.debug.update(image=img.clone(), out_s1=self._unpack(s1_ptr, ld1, n1, CH[1], kc1b),
                                  out_s2=self._unpack(s2_ptr, ld2, n2, CH[2], kc2), out_s4=self._unpack(s4_ptr, ld4, n4, CH[3], kc4),
                                  out_s8=P8.clone(), fused=fused32.clone(),
                                  conv1=self._unpack(a0.data_ptr(), 2 * CH[1], n1, CH[1], _kc(CH[1])),
                                  levels={t: lv[t].coords.clone() for t in (1, 2, 4, 8)})
            for kv in kvs:
                kv.record_stream(main)

            # ---- decoder ----
            e0, e1 = buf(n4, TR[4]), buf(n4, TR[4])
            self._conv(L, "conv4_tr", fused.data_ptr(), 2 * CH[4], up[(8, 4)], n4, None, 0, 0, False, e0.data_ptr(), 2 * TR[4],
                       _kc(TR[4]), s)
            self._block(L, "block4_tr", e0.data_ptr(), 2 * TR[4], _kc(TR[4]), nb[4], n4, TR[4], e1, cat4.data_ptr(), ld4, kc4, s)

            g0, g1 = buf(n2, TR[3]), buf(n2, TR[3])
            self._conv(L, "conv3_tr", cat4.data_ptr(), ld4, up[(4, 2)], n2, None, 0, 0, False, g0.data_ptr(), 2 * TR[3], _kc(TR[3]), s)
            self._block(L, "block3_tr", g0.data_ptr(), 2 * TR[3], _kc(TR[3]), nb[2], n2, TR[3], g1, cat2.data_ptr(), ld2, kc2, s)

            h0, h1 = buf(n1, TR[2]), buf(n1, TR[2])
            self._conv(L, "conv2_tr", cat2.data_ptr(), ld2, up[(2, 1)], n1, None, 0, 0, False, h0.data_ptr(), 2 * TR[2], _kc(TR[2]), s)
            self._block(L, "block2_tr", h0.data_ptr(), 2 * TR[2], _kc(TR[2]), nb[1], n1, TR[2], h1, cat1.data_ptr(), ld1, kc1a, s)

            # ---- tail: conv1_tr -> ReLU -> final(+bias) -> L2 norm (resunet.py:224-233) ----
            out = torch.empty((n1, m.out_channels), dtype=torch.float32, device=dev)
            _lib.check(L.imf_pointwise_tail_h2_fwd(cat1.data_ptr(), ld1, TR[2] + CH[1], TR[2], kc1a, kc1b,
                                                   m.conv1_tr.kernel.data_ptr(), TR[1], m.final.kernel.data_ptr(),
                                                   _lib.ptr(self.final_bias), m.out_channels, None, n1,
                                                   1 if m.normalize_feature else 0, None, out.data_ptr(), m.out_channels, s))
            if self.debug is not None:
                self.debug.update(out_s4_tr=self._unpack(cat4.data_ptr(), ld4, n4, TR[4], kc4),
                                  out_s2_tr=self._unpack(cat2.data_ptr(), ld2, n2, TR[3], kc2),
                                  out_s1_tr=self._unpack(cat1.data_ptr(), ld1, n1, TR[2], kc1a))
            # numeric status of THIS forward (fp16-range overflow, pipeline watchdog), including the image branch's convolutions:
            # checked before the descriptors are handed out -- the eager plan is the slow path (batches, capacity fallback), one
            # stream synchronisation per forward is not what bounds it
            self.err.bitwise_or_(img_plan.err)
            self.err_host.copy_(self.err, non_blocking=True)
            main.synchronize()
            self._raise_on_status(int(self.err_host[0]))
        return out

    @staticmethod
    def _raise_on_status(v: int):
        if v & 0x10000:
            raise FloatingPointError("an activation left the fp16 hi/lo range (|v| > 60000) in the tensor-core tier")
        if v & 0xFFFF:
            raise RuntimeError(f"in-kernel pipeline watchdog fired (code {v & 0xFFFF})")

    def check_numeric_status(self):
        """Raises if any kernel of earlier forwards reported an fp16-range overflow or a pipeline watchdog (one host sync)."""
        self._raise_on_status(int(self.err.item()))


class PlanCache:
    """Captured plans of a model, least-recently-used first, bounded by a byte budget.

    A plan preallocates everything a forward touches (~0.6 GB per 50 k voxels: activations, 10 neighbour tables, 4 hash tables, image
    plan, attention workspace), and a dataset run meets many (row bucket, image size) keys times `streams` pool slots: without a
    bound the GPU memory only grows.  Keys map to one plan or to a pool (list) of plans; `trim` drops whole entries, oldest first,
    never the entry in use.  A dropped plan that is still in flight stays alive through the caller's reference until it is retired.
    Budget: IMFNET_B200_PLAN_CACHE_GB (default 32 of the B200's 180 GB)."""

    def __init__(self, budget_bytes: int | None = None):
        import collections
        import os
        self._d = collections.OrderedDict()
        gb = float(os.environ.get("IMFNET_B200_PLAN_CACHE_GB", "32"))
        self.budget = int(gb * (1 << 30)) if budget_bytes is None else int(budget_bytes)
        self.evictions = 0

    def _touch(self, key):
        self._d.move_to_end(key)

    def get(self, key, default=None):
        if key in self._d:
            self._touch(key)
            return self._d[key]
        return default

    def setdefault(self, key, default):
        if key not in self._d:
            self._d[key] = default
        self._touch(key)
        return self._d[key]

    def __setitem__(self, key, value):
        self._d[key] = value
        self._touch(key)

    def __getitem__(self, key):
        self._touch(key)
        return self._d[key]

    def __delitem__(self, key):
        del self._d[key]

    def __contains__(self, key):
        return key in self._d

    def __len__(self):
        return len(self._d)

    def keys(self):
        return self._d.keys()

    def items(self):
        return self._d.items()

    def clear(self):
        self._d.clear()

    @staticmethod
    def _entry_bytes(entry) -> int:
        plans = entry if isinstance(entry, list) else [entry]
        total = 0
        for g in plans:
            if getattr(g, "_nbytes", None) is None:
                g._nbytes = g.nbytes()
            total += g._nbytes
        return total

    def nbytes(self) -> int:
        return sum(self._entry_bytes(e) for e in self._d.values())

    def trim(self, keep=None):
        """Evict least-recently-used entries until the cache fits its budget (the entry `keep` always stays)."""
        total = self.nbytes()
        for key in list(self._d.keys()):
            if total <= self.budget:
                break
            if key == keep:
                continue
            total -= self._entry_bytes(self._d.pop(key))
            self.evictions += 1


class PlanCapacityError(RuntimeError):
    """A level of the fragment is larger than the static plan was sized for (the caller falls back to the eager plan)."""


class GraphPlan:
    """The same forward as FusedPlan.run for ONE fragment (batch size 1), with every size read on the device and the
    whole launch sequence (coordinate pyramid, 10 neighbour tables, 23 convolutions, image encoder on a forked stream,
    attention fusion, tail) captured once in a CUDA graph per (row bucket, image size).

    Nothing in the sequence depends on a host-side row count: the kernels take `n_dev` pointers into `meta`, the persistent
    convolution kernel partitions its rows on the device, and the buffers are allocated for `rows` voxels (`cap8` tokens at
    stride 8).  A replay therefore costs one graph launch instead of ~150 Python-side launches and one host read-back,
    which is what bounded the eager plan (2.7 ms of host enqueue for 3.4 ms of device time at 50 k voxels)."""

    replayed_launches = 0      # kernels of this library launched through graph replays (bench.py adds it to imf_launch_count)
    ROW_SLACK = 4096           # a plan of `rows` voxels only ever serves fragments of more than rows - ROW_SLACK voxels

    ERR_ITEM_CAPACITY = 0x20000      # status bits of imf_batch_segments_n
    ERR_BATCH_INDEX = 0x40000

    def __init__(self, fused: FusedPlan, rows: int, H: int, W: int, cap8: int, B: int = 1):
        """rows: voxel capacity (all items together); cap8: token capacity of the fusion module at stride 8 (all items together);
        B: batch items the coordinates may name (column 0 < B), each with its own image."""
        self.f, self.rows, self.H, self.W, self.cap8 = fused, int(rows), int(H), int(W), int(cap8)
        self.B = self.num_items = int(B)
        m, dev = fused.m, fused.device
        self.m, self.device = m, dev
        L = _lib.lib()
        CH, TR = fused.CH, fused.TR
        rows = self.rows
        i32 = dict(dtype=torch.int32, device=dev)
        u8 = dict(dtype=torch.uint8, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        self.coords = {1: torch.zeros((rows, 4), **i32)}
        self.feats = torch.zeros((rows, m.conv1.in_channels), **f32)
        self.image = torch.zeros((self.B, 3, self.H, self.W), **f32)
        self.n1 = torch.zeros(1, **i32)
        self.meta = torch.zeros(16, **i32)                  # [0] status, [2] n(stride 2), [3] n(stride 4), [4] n(stride 8)
        self.err = torch.zeros(1, **i32)
        self.meta_host = torch.zeros(17, dtype=torch.int32).pin_memory()
        self.cap = int(L.imf_hash_capacity(rows))
        self.tables = {t: torch.empty(int(L.imf_hash_bytes(self.cap)), **u8) for t in (1, 2, 4, 8)}
        for t in (2, 4, 8):
            self.coords[t] = torch.zeros((rows, 4), **i32)
        self.sm_ws_bytes = int(L.imf_stride_map_workspace_bytes(rows))
        self.sm_ws = torch.empty(self.sm_ws_bytes, **u8)
        self.ldn = (rows + 127) // 128 * 128
        self.perm = {t: torch.zeros(self.ldn, **i32) for t in (1, 2, 4)}
        self.perm_ws_bytes = int(L.imf_parity_perm_workspace_bytes(rows))
        self.perm_ws = torch.empty(self.perm_ws_bytes, **u8)
        self.nbr = {}
        for key in [(1, 1, False), (2, 2, False), (4, 4, False), (8, 8, False), (1, 2, False), (2, 4, False), (4, 8, False),
                    (8, 4, True), (4, 2, True), (2, 1, True)]:
            self.nbr[key] = (torch.empty((27, self.ldn), **i32), self.ldn, torch.zeros(self.ldn // 128 + 1, **i32))

        def h2(c, n=rows):      # an h2 matrix of c channels has the footprint of an fp32 [n, c] matrix
            return torch.zeros((n, c), **f32)

        self.cat1, self.cat2, self.cat4 = h2(TR[2] + CH[1]), h2(TR[3] + CH[2]), h2(TR[4] + CH[3])
        self.a0, self.a1 = h2(CH[1]), h2(CH[1])
        self.b0, self.b1 = h2(CH[2]), h2(CH[2])
        self.c0, self.c1 = h2(CH[3]), h2(CH[3])
        self.d0, self.d1, self.d2, self.fused = h2(CH[4]), h2(CH[4]), h2(CH[4]), h2(CH[4])
        self.e0, self.e1 = h2(TR[4]), h2(TR[4])
        self.g0, self.g1 = h2(TR[3]), h2(TR[3])
        self.h0, self.h1 = h2(TR[2]), h2(TR[2])
        self.P8, self.fused32 = torch.zeros((self.cap8, CH[4]), **f32), torch.zeros((self.cap8, CH[4]), **f32)
        self.side = torch.cuda.Stream(device=dev)
        with torch.cuda.device(dev):
            self.image_plan = self._make_image_plan(m, fused)          # private buffers: plans run concurrently
        self.n_tok = self.image_plan.P2                     # image tokens per image: conv arithmetic, not H/8 * W/8, for odd sizes
        af = m.attention_fusion
        if af.inner != 128:
            raise NotImplementedError("the captured plans implement the IMFNet fusion head (one cross head of 128 channels)")
        # fusion module: all items through one chain of launches (imf_attention_fusion_fwd_batched); per-item row ranges stay on the device
        self.seg = torch.zeros(self.B + 1, **i32)
        self.cnt = torch.zeros(self.B, **i32)
        self.kv = torch.empty(int(L.imf_attention_kv_batched_bytes(self.n_tok, self.B)), **u8)
        self.kv_ws_bytes = int(L.imf_attention_kv_batched_workspace_bytes(self.n_tok, af.dim, af.inner, self.B))
        self.kv_ws = torch.empty(self.kv_ws_bytes, **u8)
        self.att_ws_bytes = int(L.imf_attention_batched_workspace_bytes(self.cap8, self.n_tok, af.latent_dim, af.inner, self.B))
        self.att_ws = torch.empty(max(self.att_ws_bytes, 1), **u8)
        self.conv_ws_bytes = int(L.imf_sparse_conv_g4_workspace_bytes(max(max(CH[1:]), max(TR[1:]))))
        self.conv_ws = torch.zeros(self.conv_ws_bytes, **u8)          # head = arrival counters, zero on entry / left zero
        self.ident = self.ident_mask = None
        if fused.tail_tc is not None:
            self.ident = torch.zeros(self.ldn, **i32)
            self.ident_mask = torch.zeros(self.ldn // 128 + 1, **i32)
        self.cf_ws = self.cf_grid = None
        if fused.conv1_tc is not None:
            self.cf_ws_bytes = int(L.imf_conv_first_tc_workspace_bytes(rows, m.conv1.kernel_size))
            # zero once: the plan never clears the dense grid, it un-scatters the occupied cells after every use (imf_conv_first_tc_release)
            self.cf_ws = torch.zeros(self.cf_ws_bytes, **u8)
            import ctypes as C
            gm, gc = C.c_void_p(), C.c_void_p()
            _lib.check(L.imf_conv_first_tc_grid(self.cf_ws.data_ptr(), rows, m.conv1.kernel_size, C.byref(gm), C.byref(gc)))
            self.cf_grid = (gm.value, gc.value)          # header / cells of the stride-1 row-index grid (also read by the neighbour tables)
        self.out = torch.zeros((rows, m.out_channels), **f32)
        self.graph = None
        self.launches_per_replay = 0
        self._tl = None      # layer-timing records while time_layers() runs, else None

    def _make_image_plan(self, m, fused):
        from .model.Img_Encoder import ImagePlan
        if self.B != 1:
            raise NotImplementedError("several images per replay: imfnet_b200.batched.BatchGraphPlan")
        return ImagePlan(m.img_encoder.backbone, self.H, self.W, fused.split_small, err=self.err)

    # -- footprint (the model's plan cache evicts by bytes) ------------------------------------------
    def nbytes(self) -> int:
        seen, total = set(), 0

        def walk(o):
            nonlocal total
            if isinstance(o, torch.Tensor):
                st = o.untyped_storage()
                if st.data_ptr() not in seen:
                    seen.add(st.data_ptr())
                    total += st.nbytes()
            elif isinstance(o, dict):
                for v in o.values():
                    walk(v)
            elif isinstance(o, (list, tuple)):
                for v in o:
                    walk(v)
            elif hasattr(o, "__dict__") and type(o).__name__ in ("ImagePlan", "BatchedImagePlan", "_Conv"):
                walk(vars(o))

        walk({k: v for k, v in vars(self).items() if k not in ("f", "m")})
        return total

    # -- layer timing (bench.py's roofline leg) ---------------------------------------------------------
    def _tl_begin(self, name, **info):
        if self._tl is None:
            return None
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        return (name, info, e0)

    def _tl_end(self, tok):
        if tok is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            self._tl.append(tok + (e1,))

    def time_layers(self):
        """Runs the plan's launch sequence ONCE eagerly on the current stream, on whatever inputs the last launch left in the static
        buffers, with a CUDA-event pair around every sparse-convolution launch (conv1, the 3x3x3 layers, the fused 1x1 tail); the
        image branch is not forked meanwhile, so nothing else runs beside the timed kernels.  Returns [(layer, info, ms)] in launch
        order: every kernel is timed in the cache state the previous layer leaves behind, as in a replay."""
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            self._tl = []
            try:
                self._enqueue()
                torch.cuda.synchronize(self.device)
                return [(name, info, e0.elapsed_time(e1)) for name, info, e0, e1 in self._tl]
            finally:
                self._tl = None

    # -- the launch sequence -------------------------------------------------------------------
    def _n(self, t):
        return self.n1.data_ptr() if t == 1 else self.meta.data_ptr() + 4 * {2: 2, 4: 3, 8: 4}[t]

    def _conv(self, L, cname, X, ldx, key, t_out, R, ldr, kc_r, relu, Y, ldy, kc_out, s):
        conv, packed, scale, shift, kci = self.f.conv[cname]
        nbr_t, ld_n, tile_mask = self.nbr[key]
        # a stride-1 level of this bucket always has >= 128 * _lib.sm_count() rows when the bucket is large enough: row mode is certain, so no
        # split workspace (and no reduce launch) is needed there
        split = self.f.split_small and not (t_out == 1 and self.rows - self.ROW_SLACK >= 128 * _lib.sm_count())
        out_row = self.perm[t_out].data_ptr() if key[2] else None          # transposed: the table is in parity-grouped row order
        tok = self._tl_begin(cname, key=key, t_out=t_out, cin=conv.in_channels, cout=conv.out_channels, K=27, residual=R is not None)
        _lib.check(L.imf_sparse_conv_g4_fwd_perm(X, ldx, kci, packed.data_ptr(), nbr_t.data_ptr(), ld_n, tile_mask.data_ptr(),
                                                 self._n(t_out), self.rows, 27, conv.in_channels, conv.out_channels, scale.data_ptr(),
                                                 shift.data_ptr(), R, ldr, kc_r, 1 if relu else 0, Y, ldy, self.rows, kc_out, out_row,
                                                 self.conv_ws.data_ptr() if split else None, self.conv_ws_bytes if split else 0,
                                                 self.err.data_ptr(), s))
        self._tl_end(tok)

    def _block(self, L, name, X, ldx, kc_x, t, C, tmp, Y, ldy, kc_y, s):
        kt = _kc(C)
        self._conv(L, name + ".conv1", X, ldx, (t, t, False), t, None, 0, 0, True, tmp.data_ptr(), 2 * C, kt, s)
        self._conv(L, name + ".conv2", tmp.data_ptr(), 2 * C, (t, t, False), t, X, ldx, kc_x, True, Y, ldy, kc_y, s)

    def _enqueue(self):
        m, f, L = self.m, self.f, _lib.lib()
        CH, TR, rows = f.CH, f.TR, self.rows
        main = torch.cuda.current_stream()
        s = main.cuda_stream
        status = self.meta.data_ptr()
        self._enqueue_image(m, main)
        # ---- coordinates: hash, pyramid, neighbour tables (all sizes stay on the device) ----
        self.meta.zero_()
        self.err.zero_()
        _lib.check(L.imf_hash_build(self.coords[1].data_ptr(), self._n(1), rows, self.tables[1].data_ptr(), self.cap, status, s))
        prev = 1
        for t in (2, 4, 8):
            _lib.check(L.imf_stride_map(self.coords[prev].data_ptr(), self._n(prev), rows, t, self.tables[t].data_ptr(), self.cap,
                                        self.coords[t].data_ptr(), self._n(t), None, self.sm_ws.data_ptr(), self.sm_ws_bytes, status, s))
            prev = t
        # transposed convolutions: output rows grouped by coordinate parity, so a tile walks 1-8 offsets instead of 27
        for t in (1, 2, 4):
            _lib.check(L.imf_parity_perm(self.coords[t].data_ptr(), self._n(t), rows, t, self.perm[t].data_ptr(), self.perm_ws.data_ptr(),
                                         self.perm_ws_bytes, s))
        # ---- encoder ----
        ld1, ld2, ld4 = 2 * self.cat1.shape[1], 2 * self.cat2.shape[1], 2 * self.cat4.shape[1]
        s1, s2, s4 = self.cat1.data_ptr() + TR[2] * 4, self.cat2.data_ptr() + TR[3] * 4, self.cat4.data_ptr() + TR[4] * 4
        kc1a, kc1b = _kc(TR[2]), _kc(CH[1])
        if f.tail_tc is not None:
            kc1a = kc1b = 32          # the tail reads [decoder | skip] as one h2 matrix of chunk width 32
        kc2, kc4, k8 = _kc(TR[3], CH[2]), _kc(TR[4], CH[3]), _kc(CH[4])
        sc, sh = f.norm1
        tok = self._tl_begin("conv1", t_out=1, cin=m.conv1.in_channels, cout=CH[1], K=m.conv1.kernel_size ** 3, residual=False)
        if self.cf_ws is not None:          # leaves the stride-1 dense grid populated for the neighbour tables below
            packed1, sc1, sh1 = f.conv1_tc
            _lib.check(L.imf_conv_first_tc_h2_fwd_keep(self.feats.data_ptr(), self.feats.shape[1], packed1.data_ptr(), self.coords[1].data_ptr(),
                                                       self._n(1), rows, self.num_items, self.tables[1].data_ptr(), self.cap, m.conv1.kernel_size,
                                                       CH[1], sc1.data_ptr(), sh1.data_ptr(), 0, self.a0.data_ptr(), 2 * CH[1], _kc(CH[1]),
                                                       self.cf_ws.data_ptr(), self.cf_ws_bytes, self.err.data_ptr(), s))
        else:
            _lib.check(L.imf_conv_first_h2_fwd(self.feats.data_ptr(), self.feats.shape[1], m.conv1.in_channels, m.conv1.kernel.data_ptr(),
                                               self.coords[1].data_ptr(), self._n(1), rows, self.tables[1].data_ptr(), self.cap,
                                               m.conv1.kernel_size, 1, CH[1], sc.data_ptr(), sh.data_ptr(), 0, self.a0.data_ptr(),
                                               2 * CH[1], _kc(CH[1]), s))
        self._tl_end(tok)
        # ---- neighbour tables: all 10 in one launch; those over the stride-1 set read the dense grid conv1 left behind ----
        jobs = (_lib.KmapJob * len(self.nbr))()
        for i, ((t_in, t_out, tr), (nbr_t, ld_n, mask)) in enumerate(self.nbr.items()):
            gm, gc = self.cf_grid if (self.cf_grid is not None and t_in == 1 and not tr) else (None, None)
            jobs[i] = _lib.KmapJob(self.coords[t_out].data_ptr(), self._n(t_out), self.tables[t_in].data_ptr(), nbr_t.data_ptr(),
                                   mask.data_ptr(), self.perm[t_out].data_ptr() if tr else None, -t_out if tr else t_in, gm, gc)
        _lib.check(L.imf_kernel_map_t_batch(jobs, len(self.nbr), rows, self.cap, 3, self.ldn, s))
        if self.cf_ws is not None:
            _lib.check(L.imf_conv_first_tc_release(self.coords[1].data_ptr(), self._n(1), rows, self.num_items, m.conv1.kernel_size,
                                                   self.cf_ws.data_ptr(), self.cf_ws_bytes, s))
        self._block(L, "block1", self.a0.data_ptr(), 2 * CH[1], _kc(CH[1]), 1, CH[1], self.a1, s1, ld1, kc1b, s)
        self._conv(L, "conv2", s1, ld1, (1, 2, False), 2, None, 0, 0, False, self.b0.data_ptr(), 2 * CH[2], _kc(CH[2]), s)
        self._block(L, "block2", self.b0.data_ptr(), 2 * CH[2], _kc(CH[2]), 2, CH[2], self.b1, s2, ld2, kc2, s)
        self._conv(L, "conv3", s2, ld2, (2, 4, False), 4, None, 0, 0, False, self.c0.data_ptr(), 2 * CH[3], _kc(CH[3]), s)
        self._block(L, "block3", self.c0.data_ptr(), 2 * CH[3], _kc(CH[3]), 4, CH[3], self.c1, s4, ld4, kc4, s)
        self._conv(L, "conv4", s4, ld4, (4, 8, False), 8, None, 0, 0, False, self.d0.data_ptr(), 2 * CH[4], k8, s)
        self._block(L, "block4", self.d0.data_ptr(), 2 * CH[4], k8, 8, CH[4], self.d1, self.d2.data_ptr(), 2 * CH[4], k8, s)
        self._enqueue_fusion(L, m, CH[4], k8, main, s)
        # ---- decoder ----
        self._conv(L, "conv4_tr", self.fused.data_ptr(), 2 * CH[4], (8, 4, True), 4, None, 0, 0, False, self.e0.data_ptr(), 2 * TR[4],
                   _kc(TR[4]), s)
        self._block(L, "block4_tr", self.e0.data_ptr(), 2 * TR[4], _kc(TR[4]), 4, TR[4], self.e1, self.cat4.data_ptr(), ld4, kc4, s)
        self._conv(L, "conv3_tr", self.cat4.data_ptr(), ld4, (4, 2, True), 2, None, 0, 0, False, self.g0.data_ptr(), 2 * TR[3], _kc(TR[3]), s)
        self._block(L, "block3_tr", self.g0.data_ptr(), 2 * TR[3], _kc(TR[3]), 2, TR[3], self.g1, self.cat2.data_ptr(), ld2, kc2, s)
        self._conv(L, "conv2_tr", self.cat2.data_ptr(), ld2, (2, 1, True), 1, None, 0, 0, False, self.h0.data_ptr(), 2 * TR[2], _kc(TR[2]), s)
        self._block(L, "block2_tr", self.h0.data_ptr(), 2 * TR[2], _kc(TR[2]), 1, TR[2], self.h1, self.cat1.data_ptr(), ld1, kc1a, s)
        # ---- tail ----
        tok = self._tl_begin("conv1_tr+final", t_out=1, cin=TR[2] + CH[1], mid=TR[1], cout=m.out_channels, K=1, residual=False)
        if f.tail_tc is not None:
            (p1, s1, z1), (p2, s2, b2) = f.tail_tc
            c0, c1, c2 = TR[2] + CH[1], TR[1], m.out_channels
            kmid, kout = _kc(c1), _kc(c2)
            if f.tail_fused:          # one kernel: the hidden layer and the logits stay on the SM (csrc/tail_fused.cu)
                _lib.check(L.imf_tail_fused_h2_fwd(self.cat1.data_ptr(), ld1, rows, self._n(1), c0, c1, c2, p1.data_ptr(), s1.data_ptr(),
                                                   z1.data_ptr(), p2.data_ptr(), s2.data_ptr(), b2.data_ptr(), 1 if m.normalize_feature else 0,
                                                   self.out.data_ptr(), c2, self.err.data_ptr(), s))
                self._tl_end(tok)
                return
            _lib.check(L.imf_identity_table(self._n(1), rows, self.ident.data_ptr(), self.ldn, self.ident_mask.data_ptr(), s))
            # hidden = relu(cat1 . W1) -> h1 (free after block2_tr);  logits = hidden . W2 + b2 -> h0;  out = logits / ||logits||
            _lib.check(L.imf_sparse_conv_g4_fwd(self.cat1.data_ptr(), ld1, 32, p1.data_ptr(), self.ident.data_ptr(), self.ldn,
                                                self.ident_mask.data_ptr(), self._n(1), rows, 1, c0, c1, s1.data_ptr(), z1.data_ptr(), None, 0, 0,
                                                1, self.h1.data_ptr(), 2 * c1, rows, kmid, None, 0, self.err.data_ptr(), s))
            _lib.check(L.imf_sparse_conv_g4_fwd(self.h1.data_ptr(), 2 * c1, kmid, p2.data_ptr(), self.ident.data_ptr(), self.ldn,
                                                self.ident_mask.data_ptr(), self._n(1), rows, 1, c1, c2, s2.data_ptr(), b2.data_ptr(), None, 0, 0,
                                                0, self.h0.data_ptr(), 2 * c2, rows, kout, None, 0, self.err.data_ptr(), s))
            _lib.check(L.imf_h2_unpack_l2norm(self.h0.data_ptr(), 2 * c2, rows, self._n(1), c2, kout, 1 if m.normalize_feature else 0, None,
                                              self.out.data_ptr(), c2, s))
        else:
            _lib.check(L.imf_pointwise_tail_h2_fwd(self.cat1.data_ptr(), ld1, TR[2] + CH[1], TR[2], kc1a, kc1b, m.conv1_tr.kernel.data_ptr(),
                                                   TR[1], m.final.kernel.data_ptr(), _lib.ptr(f.final_bias), m.out_channels, self._n(1), rows,
                                                   1 if m.normalize_feature else 0, None, self.out.data_ptr(), m.out_channels, s))
        self._tl_end(tok)

    def _enqueue_image(self, m, main):
        """Image branch on a forked stream: encoder over all images of the replay, then K / V of their tokens in one projection
        (joined again in _enqueue_fusion)."""
        L = _lib.lib()
        side = main if self._tl is not None else self.side          # (layer timing: nothing runs beside the timed kernels)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            tokens = self.image_plan.enqueue(self.image if self.B > 1 else self.image[0])
            _lib.check(L.imf_attention_kv_batched(m.attention_fusion.packed(), m.attention_fusion.packed_h2(), tokens.data_ptr(), self.n_tok, self.B, self.kv.data_ptr(),
                                                  self.kv_ws.data_ptr(), self.kv_ws_bytes, self.err.data_ptr(), side.cuda_stream))

    def _enqueue_fusion(self, L, m, C8, k8, main, s):
        """Attention fusion at stride 8 (model/resunet.py:189, 237-273) for all items at once: d2 (h2) -> P8 (fp32) -> fused32 ->
        fused (h2).  The per-item row ranges (seg, cnt) are computed on the device; item b attends to the tokens of image b."""
        af = m.attention_fusion
        _lib.check(L.imf_batch_segments_n(self.coords[8].data_ptr(), self._n(8), self.rows, self.B, self.cap8, self.seg.data_ptr(),
                                          self.cnt.data_ptr(), self.err.data_ptr(), s))
        _lib.check(L.imf_h2_unpack_scaled_n(self.d2.data_ptr(), 2 * C8, self.cap8, self._n(8), C8, k8, 1.0 / self.f.act_scale, self.P8.data_ptr(), C8, s))
        main.wait_stream(self.side)
        _lib.check(L.imf_attention_fusion_fwd_batched(af.packed(), af.packed_h2(), self.P8.data_ptr(), C8, self.cap8, self._n(8), self.seg.data_ptr(),
                                                      self.cnt.data_ptr(), self.B, self.kv.data_ptr(), self.n_tok, self.fused32.data_ptr(), C8,
                                                      self.att_ws.data_ptr(), self.att_ws_bytes, self.err.data_ptr(), s))
        _lib.check(L.imf_h2_pack_scaled_n(self.fused32.data_ptr(), C8, self.cap8, self._n(8), C8, k8, self.f.act_scale, self.fused.data_ptr(), 2 * C8,
                                          self.err.data_ptr(), s))

    def capture(self):
        """Warm the sequence up once (lazy initialisation inside torch / cuDNN must not happen during capture), then record it."""
        L = _lib.lib()
        with torch.cuda.device(self.device):
            warm = torch.cuda.Stream(device=self.device)
            warm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(warm):
                self._enqueue()
                self._enqueue()
            torch.cuda.current_stream().wait_stream(warm)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            l0 = L.imf_launch_count()
            with torch.cuda.graph(g):
                self._enqueue()
            self.launches_per_replay = int(L.imf_launch_count() - l0)
            self.graph = g

    # -- one forward ---------------------------------------------------------------------------
    @torch.no_grad()
    def launch(self, coords: torch.Tensor, feats: torch.Tensor, image: torch.Tensor, stream=None, out_host: torch.Tensor | None = None):
        """Enqueue one forward (input copies, graph replay, result clone, status read-back) on `stream` (default: the current
        stream) without waiting for it; finish() returns the descriptors.  Plans launched on different streams overlap.
        The inputs may be pinned host tensors (the copies into the plan's static buffers are then the host->device transfers);
        with `out_host` (pinned [N, C]) the result goes straight to the host instead of into a device clone."""
        N = int(coords.shape[0])
        if N > self.rows:
            raise PlanCapacityError(f"{N} voxels > plan rows {self.rows}")
        with torch.cuda.device(self.device):
            if self.graph is None:
                self.capture()
            cur = torch.cuda.current_stream()
            st = cur if stream is None else stream
            if st is not cur:
                st.wait_stream(cur)                 # the inputs were produced on the caller's stream
            with torch.cuda.stream(st):
                self.coords[1][:N].copy_(coords, non_blocking=True)
                self.feats[:N].copy_(feats, non_blocking=True)
                self.image.copy_(image.reshape(self.image.shape), non_blocking=True)          # (B == 1 here)
                self.n1.fill_(N)
                self.graph.replay()
                GraphPlan.replayed_launches += self.launches_per_replay
                if out_host is None:
                    out = self.out[:N].clone()
                else:
                    out = out_host[:N]
                    out.copy_(self.out[:N], non_blocking=True)
                self.meta_host[:16].copy_(self.meta, non_blocking=True)
                self.meta_host[16:].copy_(self.err, non_blocking=True)
                self._done = torch.cuda.Event()
                self._done.record(st)
            self._pending = (N, out, st, cur)

    def finish(self) -> torch.Tensor:
        N, out, st, cur = self._pending
        self._pending = None
        self._done.synchronize()
        if st is not cur and out.is_cuda:
            out.record_stream(cur)
        mh = self.meta_host.tolist()
        if mh[0]:
            from .sparse import _raise_status
            _raise_status(mh[0])
        self._check_fusion_status(mh)
        self.levels = {1: N, 2: mh[2], 4: mh[3], 8: mh[4]}
        return out

    def _check_fusion_status(self, mh):
        if mh[16] & self.ERR_BATCH_INDEX:
            raise ValueError("coordinates reference more batch items than images were given")
        if mh[4] > self.cap8 or (mh[16] & self.ERR_ITEM_CAPACITY):
            raise PlanCapacityError(f"{mh[4]} stride-8 voxels > plan capacity {self.cap8}")
        FusedPlan._raise_on_status(mh[16] & ~(self.ERR_BATCH_INDEX | self.ERR_ITEM_CAPACITY))

    def run(self, coords: torch.Tensor, feats: torch.Tensor, image: torch.Tensor) -> torch.Tensor:
        self.launch(coords, feats, image)
        return self.finish()
