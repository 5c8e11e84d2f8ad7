"""ResUNet2 family with the reference's constructor, attribute names and forward(x, image) signature
(/root/reference/model/resunet.py:14-326).  Sub-modules only hold parameters under the reference's names (so
`load_state_dict` of a reference checkpoint is strict-clean); `forward` hands the whole graph to the fused CUDA
plan in imfnet_b200/engine.py.
"""
from __future__ import annotations

import torch

from .. import me as ME
import os

from ..engine import FusedPlan, GraphPlan, PlanCache, PlanCapacityError
from .attention_fusion import AttentionFusion
from .common import get_norm
from .Img_Encoder import ImageEncoder
from .residual_block import get_block


class ResUNet2(ME.MinkowskiNetwork):
    NORM_TYPE = None
    BLOCK_NORM_TYPE = 'BN'
    CHANNELS = [None, 32, 64, 128, 256]
    TR_CHANNELS = [None, 32, 64, 64, 128]
    IMG_CHANNELS = [None, 0, 0, 0, 0]

    def __init__(self, in_channels=3, out_channels=32, bn_momentum=0.1, normalize_feature=None, conv1_kernel_size=None,
                 D=3, config=None):
        super().__init__(D)
        if self.NORM_TYPE != 'BN' or self.BLOCK_NORM_TYPE != 'BN':
            raise NotImplementedError(f"{type(self).__name__}: only the batch-norm variants (ResUNetBN2*) are implemented; "
                                      "the configured IMFNet model is ResUNetBN2C (config_3dmatch.py:66)")
        CH, TR, IMG = self.CHANNELS, self.TR_CHANNELS, self.IMG_CHANNELS
        self.normalize_feature = normalize_feature
        self.in_channels, self.out_channels = in_channels, out_channels
        self.conv1_kernel_size = conv1_kernel_size

        def conv(cin, cout, k, s, tr=False, bias=False):
            cls = ME.MinkowskiConvolutionTranspose if tr else ME.MinkowskiConvolution
            return cls(in_channels=cin, out_channels=cout, kernel_size=k, stride=s, dilation=1, bias=bias, dimension=D)

        def norm(c):
            return get_norm(self.NORM_TYPE, c, bn_momentum=bn_momentum, D=D)

        def block(c):
            return get_block(self.BLOCK_NORM_TYPE, c, c, bn_momentum=bn_momentum, D=D)

        # encoder (reference lines 42-89)
        self.conv1, self.norm1, self.block1 = conv(in_channels, CH[1], conv1_kernel_size, 1), norm(CH[1]), block(CH[1])
        self.conv2, self.norm2, self.block2 = conv(CH[1], CH[2], 3, 2), norm(CH[2]), block(CH[2])
        self.conv3, self.norm3, self.block3 = conv(CH[2], CH[3], 3, 2), norm(CH[3]), block(CH[3])
        self.conv4, self.norm4, self.block4 = conv(CH[3], CH[4], 3, 2), norm(CH[4]), block(CH[4])
        # fusion (91-99): Q from stride-8 point tokens, K/V from 128-channel image tokens, one head of CH[4]/2
        self.attention_fusion = AttentionFusion(dim=128, depth=0, latent_dim=CH[4], cross_heads=1, latent_heads=8,
                                                cross_dim_head=int(CH[4] / 2), latent_dim_head=int(CH[4] / 2))
        # decoder (101-158)
        self.conv4_tr, self.norm4_tr, self.block4_tr = conv(CH[4], TR[4], 3, 2, tr=True), norm(TR[4]), block(TR[4])
        self.conv3_tr, self.norm3_tr, self.block3_tr = conv(CH[3] + TR[4] + IMG[1], TR[3], 3, 2, tr=True), norm(TR[3]), block(TR[3])
        self.conv2_tr, self.norm2_tr, self.block2_tr = conv(CH[2] + TR[3] + IMG[2], TR[2], 3, 2, tr=True), norm(TR[2]), block(TR[2])
        self.conv1_tr = conv(CH[1] + TR[2] + IMG[3], TR[1], 1, 1)
        self.final = conv(TR[1], out_channels, 1, 1, bias=True)
        self.img_encoder = ImageEncoder()
        self._plan = None

    def _apply(self, fn, *a, **k):
        self._plan = None                      # weights moved / cast: re-pack lazily (and drop the captured graphs)
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._plan = None
        return super().load_state_dict(*a, **k)

    def forward(self, x, image):
        """x: sparse tensor exposing .F [N,Cin] fp32, .C [N,4] int32 (batch,x,y,z); image: [B,3,H,W] fp32.
        Returns a SparseTensor on x's coordinate map with .F = [N,out_channels] (L2-normalised rows if
        normalize_feature), rows in x's order (reference lines 163-235)."""
        if self.training:
            raise NotImplementedError("imfnet_b200 implements the eval-mode forward (BatchNorm running statistics); "
                                      "call model.eval() as util/misc.py:44-45 does")
        self._ensure_plan()
        if not isinstance(x, ME.SparseTensor):       # duck-typed foreign container (.F / .C), e.g. a real ME tensor
            x = ME.SparseTensor(x.F, coordinates=x.C)
        image = torch.as_tensor(image)
        F = None
        if (image.dim() == 4 and image.shape[0] == 1 and self.use_cuda_graph and x.coordinate_map_key.tensor_stride == 1
                and self._plan.debug is None):
            F = self._forward_graph(x, image)
        if F is None:
            F = self._plan.run(x, image)
        return ME.SparseTensor(F, coordinate_map_key=x.coordinate_map_key, coordinate_manager=x.coordinate_manager)

    # one fragment per call (what util/misc.py:extract_features and scripts/generate_desc.py do): captured CUDA graph per
    # (row bucket, image size); larger batches and oversized stride-8 levels take the eager plan
    use_cuda_graph = os.environ.get("IMFNET_B200_GRAPH", "1") != "0"
    # False (default): tuned for throughput with several fragments in flight (forward_many / forward_many_host); True: small levels
    # split their work over more CTAs, which shortens a lone forward() by ~13 % and costs ~10 % throughput.  Both settings give the
    # same descriptors up to fp32 summation order.  Read when the plans are built; changing it rebuilds them.
    low_latency = os.environ.get("IMFNET_B200_LOW_LATENCY", "0") == "1"

    def _ensure_plan(self):
        if self._plan is None or self._plan.split_small != bool(self.low_latency):
            self._plan = FusedPlan(self)
            self._graphs, self._cap8_scale = PlanCache(), {}
    ROW_BUCKET = GraphPlan.ROW_SLACK

    @torch.no_grad()
    def forward_many(self, items, streams: int = 2):
        """[(x, image), ...] -> [SparseTensor, ...]: independent fragments (what scripts/generate_desc.py iterates over) run
        through `streams` captured plans on as many CUDA streams, so the latency-bound parts of one fragment (coordinate
        pyramid, small deep levels, attention) overlap the others' work.  Results are identical to forward() one by one."""
        if self.training:
            raise NotImplementedError("imfnet_b200 implements the eval-mode forward")
        self._ensure_plan()
        plan = self._plan
        if plan._key != plan._weights_key():
            plan.pack()
            self._graphs.clear()
        outs = [None] * len(items)
        inflight = []                                     # (index, x, GraphPlan)

        def retire():
            i, xi, g = inflight.pop(0)
            try:
                F = g.finish()
            except PlanCapacityError:
                F = None
            if F is None:
                F = self(xi, items[i][1]).F
            outs[i] = ME.SparseTensor(F, coordinate_map_key=xi.coordinate_map_key, coordinate_manager=xi.coordinate_manager)

        for i, (x, image) in enumerate(items):
            if not isinstance(x, ME.SparseTensor):
                x = ME.SparseTensor(x.F, coordinates=x.C)
            image = torch.as_tensor(image)
            N = len(x.F)
            rows = (max(N, 1) + self.ROW_BUCKET - 1) // self.ROW_BUCKET * self.ROW_BUCKET
            key = (rows, int(image.shape[2]), int(image.shape[3]))
            ok = (image.dim() == 4 and image.shape[0] == 1 and self.use_cuda_graph and N > 0 and self._cap8_scale.get(key, 1) == 1
                  and x.coordinate_map_key.tensor_stride == 1)
            if not ok:
                while inflight:
                    retire()
                outs[i] = self(x, image)
                continue
            pool = self._graphs.setdefault(("pool",) + key, [])
            slot = i % max(1, streams)
            while len(pool) <= slot:
                cap8 = self._cap8(rows, 1)
                g = GraphPlan(plan, rows, key[1], key[2], cap8)
                g.stream = torch.cuda.Stream(device=plan.device)
                pool.append(g)
                self._graphs.trim(keep=("pool",) + key)
            g = pool[slot]
            while any(e[2] is g for e in inflight):
                retire()
            g.launch(x.C, x.F.to(device=plan.device, dtype=torch.float32), image.to(device=plan.device, dtype=torch.float32), g.stream)
            inflight.append((i, x, g))
        while inflight:
            retire()
        return outs

    @staticmethod
    def _cap8(rows: int, scale: int) -> int:
        """Token capacity of a plan at stride 8.  Indoor fragments have ~1/46 of their voxels left at stride 8 (real and synthetic
        3DMatch data); the plan is sized for 1/28 (the dense GEMMs of the fusion module pick their tiling from this
        capacity, so a loose bound costs time) and a fragment that exceeds it falls back to the eager plan once and gets a 4x plan."""
        return min(rows, max(512, (rows // 28 + 255) // 256 * 256) * scale)

    @torch.no_grad()
    def forward_many_host(self, frags, streams: int = 2, out=None):
        """End-to-end form of forward_many for host-resident fragments: frags = [(coords int32 [N,4], feats fp32 [N,Cin],
        image fp32 [1,3,H,W]), ...] as (ideally pinned) CPU tensors; returns pinned CPU tensors [N, out_channels].  Every fragment's
        host->device copies, graph replay and device->host copy run on its plan's stream, so the transfers of one fragment overlap
        the compute of the others (the reference does `feature.detach().cpu()` synchronously per fragment,
        scripts/generate_desc.py:118-123).  `out` may provide the destination tensors."""
        if self.training:
            raise NotImplementedError("imfnet_b200 implements the eval-mode forward")
        self._ensure_plan()
        plan = self._plan
        if plan._key != plan._weights_key():
            plan.pack()
            self._graphs.clear()
        outs = [None] * len(frags)
        inflight = []

        def eager(i):
            c, f, im = frags[i]
            x = ME.SparseTensor(f.to(plan.device, non_blocking=True), coordinates=c.to(plan.device, non_blocking=True))
            F = self(x, im.to(plan.device, non_blocking=True)).F
            dst = out[i] if out is not None else torch.empty(F.shape, dtype=F.dtype, pin_memory=True)
            dst[: len(F)].copy_(F)
            return dst[: len(F)]

        def retire():
            i, g = inflight.pop(0)
            try:
                outs[i] = g.finish()
            except PlanCapacityError:
                outs[i] = eager(i)

        for i, (c, f, im) in enumerate(frags):
            N = int(c.shape[0])
            rows = (max(N, 1) + self.ROW_BUCKET - 1) // self.ROW_BUCKET * self.ROW_BUCKET
            key = (rows, int(im.shape[2]), int(im.shape[3]))
            if not (self.use_cuda_graph and N > 0 and im.shape[0] == 1 and self._cap8_scale.get(key, 1) == 1):
                while inflight:
                    retire()
                outs[i] = eager(i)
                continue
            pool = self._graphs.setdefault(("pool",) + key, [])
            slot = i % max(1, streams)
            while len(pool) <= slot:
                g = GraphPlan(plan, rows, key[1], key[2], self._cap8(rows, 1))
                g.stream = torch.cuda.Stream(device=plan.device)
                pool.append(g)
                self._graphs.trim(keep=("pool",) + key)
            g = pool[slot]
            while any(e[1] is g for e in inflight):
                retire()
            dst = out[i] if out is not None else torch.empty((N, self.out_channels), dtype=torch.float32, pin_memory=True)
            g.launch(c, f.float(), im.float(), g.stream, out_host=dst)
            inflight.append((i, g))
        while inflight:
            retire()
        return outs

    def forward_batches(self, frags, batch: int, streams: int = 2, out=None, carry=None):
        """Throughput form for many fragments: groups of `batch` fragments per captured-graph replay (imfnet_b200/batched.py).
        carry: a dict kept by the caller for streaming use (the call does not wait for its last groups; `drain_batches(carry)` does)."""
        from ..batched import forward_batches
        return forward_batches(self, frags, batch, streams, out, carry)

    @staticmethod
    def drain_batches(carry):
        from ..batched import drain_batches
        drain_batches(carry)

    def _forward_graph(self, x, image):
        plan = self._plan
        if plan._key != plan._weights_key():
            plan.pack()
            self._graphs.clear()
        N = len(x.F)
        if N == 0:
            return None
        rows = (N + self.ROW_BUCKET - 1) // self.ROW_BUCKET * self.ROW_BUCKET
        key = (rows, int(image.shape[2]), int(image.shape[3]))
        scale = self._cap8_scale.get(key, 1)
        if scale > 16:
            return None
        g = self._graphs.get(key)
        if g is None:
            cap8 = self._cap8(rows, scale)
            g = self._graphs[key] = GraphPlan(plan, rows, key[1], key[2], cap8)
            self._graphs.trim(keep=key)
        feats = x.F.to(device=plan.device, dtype=torch.float32)
        try:
            return g.run(x.C, feats, image.to(device=plan.device, dtype=torch.float32))
        except PlanCapacityError:
            self._cap8_scale[key] = scale * 4          # unusually dense stride-8 level: bigger plan next time, eager now
            del self._graphs[key]
            return None


class ResUNetBN2(ResUNet2):
    NORM_TYPE = 'BN'


class ResUNetBN2B(ResUNet2):
    NORM_TYPE = 'BN'
    CHANNELS = [None, 32, 64, 128, 256]
    TR_CHANNELS = [None, 64, 64, 64, 64]


class ResUNetBN2C(ResUNet2):
    NORM_TYPE = 'BN'
    CHANNELS = [None, 32, 64, 128, 256]
    TR_CHANNELS = [None, 64, 64, 64, 128]


class ResUNetBN2D(ResUNet2):
    NORM_TYPE = 'BN'
    CHANNELS = [None, 32, 64, 128, 256]
    TR_CHANNELS = [None, 64, 64, 128, 128]


class ResUNetBN2E(ResUNet2):
    NORM_TYPE = 'BN'
    CHANNELS = [None, 128, 128, 128, 256]
    TR_CHANNELS = [None, 64, 128, 128, 128]


class ResUNetIN2(ResUNet2):
    NORM_TYPE = 'BN'
    BLOCK_NORM_TYPE = 'IN'


class ResUNetIN2B(ResUNetBN2B):
    NORM_TYPE = 'BN'
    BLOCK_NORM_TYPE = 'IN'


class ResUNetIN2C(ResUNetBN2C):
    NORM_TYPE = 'BN'
    BLOCK_NORM_TYPE = 'IN'


class ResUNetIN2D(ResUNetBN2D):
    NORM_TYPE = 'BN'
    BLOCK_NORM_TYPE = 'IN'


class ResUNetIN2E(ResUNetBN2E):
    NORM_TYPE = 'BN'
    BLOCK_NORM_TYPE = 'IN'
