"""Batched captured plan: B fragments (+ their B images) through ONE CUDA-graph replay.

The reference accepts batches -- `collate_pair_fn` builds batched coordinates (/root/reference/lib/data_loaders.py:28-91) and
`ResUNet2.forward` / `transformer` split the stride-8 rows per batch item against that item's image
(/root/reference/model/resunet.py:163-273) -- while `scripts/generate_desc.py:65-123` feeds one fragment per call.  For
throughput the batched form is the better shape on a B200: every sparse convolution of the forward becomes ONE persistent
launch over all items' rows, so the per-CTA fixed cost of the convolution kernel (prologue + epilogue, ~13 k cycles, DESIGN.md
section 4.1) and the under-filled small levels (1-4 k rows per fragment at strides 4 and 8, 4 800 pixels in the image encoder's
layer2: 9-38 CTAs on 148 SMs) are paid once per batch instead of once per fragment.  `GraphPlan` (engine.py) keeps several
single-fragment graphs in flight instead; this module is the alternative the bench can switch to (`bench.py --batched B`).

Everything arithmetic runs in the same kernels as the single-fragment plan; per output row the summation order of every
convolution is unchanged (offset order, chunk order), so the descriptors are expected to be bit-identical to forward() fragment by
fragment in the throughput setting (tests/test_gpu_batched.py).

What is specific to a batch:
  * coordinates carry the item index in column 0 (set on the device after the host->device copy);
  * the image encoder runs once over B*H*W pixel rows with neighbour tables replicated per image (`BatchedImagePlan`);
  * the fusion module runs ONCE over all items' stride-8 rows (everything but the attention core is row-wise); the attention
    kernel takes the per-item row ranges, computed on the device (csrc/batched.cu: imf_batch_segments_n), and lets item b attend
    to image b's tokens (csrc/flash_fusion.cu).  The single-fragment plan is the same code with B = 1 (engine.GraphPlan).

This is the execution mode bench.py measures; tests/test_gpu_batched.py checks it against the CPU oracle, the reference's golden
outputs and forward() (bit-identical) on the GPU, tests/test_plan_emulated.py its host logic on the C-ABI emulator.
"""
from __future__ import annotations

import torch

from . import _lib
from .engine import FusedPlan, GraphPlan, PlanCapacityError
from .model.Img_Encoder import ImagePlan


class BatchedImagePlan(ImagePlan):
    """ImagePlan over B images of one size: pixel rows of image b are rows [b*P, (b+1)*P) of every matrix."""

    def __init__(self, backbone, H: int, W: int, B: int, err=None):
        super().__init__(backbone, H, W, False, err=err)
        self.B = B = int(B)
        dev = self.device
        with torch.cuda.device(dev):
            self.t_id0 = self._replicate(self.t_id0, self.P0, self.P0)
            self.t1 = self._replicate(self.t1, self.P1, self.P1)
            self.t12 = self._replicate(self.t12, self.P1, self.P2)
            self.t12d = self._replicate(self.t12d, self.P1, self.P2)
            self.t2 = self._replicate(self.t2, self.P2, self.P2)
            f32 = dict(dtype=torch.float32, device=dev)
            self._alloc_stem(B)
            self.s0 = torch.zeros((B * self.P0, self.C1), **f32)
            self._alloc_layer1(B)
            self.l2 = [torch.zeros((B * self.P2, self.C2), **f32) for _ in range(3)]
            self.tokens = torch.zeros((B * self.P2, self.C2), **f32)

    def _replicate(self, tab, p_in: int, p_out: int):
        """Single-image offset-major table [K, ld] -> table of B images (input rows shifted by b*p_in) + its 128-row tile masks."""
        nbr_t, _ld, _mask = tab
        B, K = self.B, nbr_t.shape[0]
        t = nbr_t[:, :p_out]
        n = B * p_out
        ld = (n + 127) // 128 * 128
        out = torch.full((K, ld), -1, dtype=torch.int32, device=nbr_t.device)
        for b in range(B):
            out[:, b * p_out:(b + 1) * p_out] = torch.where(t >= 0, t + b * p_in, t)
        valid = (out >= 0).view(K, ld // 128, 128).any(dim=2).to(torch.int32)                     # [K, tiles]
        bits = (valid << torch.arange(K, dtype=torch.int32, device=out.device).view(K, 1)).sum(dim=0).to(torch.int32)
        mask = torch.zeros(ld // 128 + 1, dtype=torch.int32, device=out.device)
        mask[: ld // 128] = bits
        return out, ld, mask

    def enqueue(self, images: torch.Tensor) -> torch.Tensor:
        """images fp32 [B,3,H,W] (contiguous, on the plan's device) -> fp32 tokens [B * H/8*W/8, 128] (plan-owned)."""
        L = _lib.lib()
        s = _lib.cur_stream()
        B = self.B
        self._stem(L, images, B, s)
        n2 = B * self.P2
        x = self._layer1(L, B, s)
        y, tmp, out = self.l2
        for c1, c2, down in self.blocks2:
            if down is not None:
                self._conv(L, c1, x, self.t12, n2, None, True, tmp, s)
                self._conv(L, down, x, self.t12d, n2, None, False, out, s)
                self._conv(L, c2, tmp, self.t2, n2, out, True, y, s)
            else:
                self._conv(L, c1, y, self.t2, n2, None, True, tmp, s)
                self._conv(L, c2, tmp, self.t2, n2, y, True, out, s)
                y, out = out, y
        _lib.check(L.imf_h2_unpack_scaled_n(y.data_ptr(), 2 * self.C2, n2, None, self.C2, 64, 1.0 / self.act_scale, self.tokens.data_ptr(), self.C2, s))
        return self.tokens


class BatchGraphPlan(GraphPlan):
    """GraphPlan for exactly B fragments of at most `rows` voxels in total (each sized for `item_cap8` rows at stride 8: the fusion
    module's buffers hold B * item_cap8 tokens, shared by the items)."""

    def __init__(self, fused: FusedPlan, rows: int, H: int, W: int, B: int, item_cap8: int):
        if fused.split_small:
            raise NotImplementedError("the batched plan implements the throughput setting (model.low_latency = False)")
        self.item_cap8 = int(item_cap8)
        super().__init__(fused, rows, H, W, cap8=min(int(rows), int(B) * int(item_cap8)), B=B)

    def _make_image_plan(self, m, fused):
        return BatchedImagePlan(m.img_encoder.backbone, self.H, self.W, self.B, err=self.err)

    # -- one batch ---------------------------------------------------------------------------------
    @torch.no_grad()
    def launch_batch(self, frags, stream=None, out_hosts=None):
        """frags = exactly B tuples (coords int32 [N_b,4], feats fp32 [N_b,Cin], image fp32 [1,3,H,W] or [3,H,W]); device or
        (pinned) host tensors.  Column 0 of the coordinates is overwritten with the item index on the device.  Enqueues the
        copies, one graph replay and the result copies on `stream` without waiting; finish_batch() returns the descriptors."""
        if len(frags) != self.B:
            raise ValueError(f"the plan was built for batches of {self.B} fragments, got {len(frags)}")
        sizes = [int(c.shape[0]) for c, _f, _im in frags]
        total = sum(sizes)
        if total > self.rows:
            raise PlanCapacityError(f"{total} voxels > plan rows {self.rows}")
        with torch.cuda.device(self.device):
            if self.graph is None:
                self.capture()
            cur = torch.cuda.current_stream()
            st = cur if stream is None else stream
            if st is not cur:
                st.wait_stream(cur)
            outs = []
            with torch.cuda.stream(st):
                off = 0
                for b, (c, ft, im) in enumerate(frags):
                    n = sizes[b]
                    self.coords[1][off:off + n].copy_(c, non_blocking=True)
                    self.coords[1][off:off + n, 0].fill_(b)
                    self.feats[off:off + n].copy_(ft, non_blocking=True)
                    self.image[b].copy_(im.reshape(self.image.shape[1:]), non_blocking=True)
                    off += n
                self.n1.fill_(total)
                self.graph.replay()
                GraphPlan.replayed_launches += self.launches_per_replay
                off = 0
                for b, n in enumerate(sizes):
                    if out_hosts is None:
                        o = self.out[off:off + n].clone()
                    else:
                        o = out_hosts[b][:n]
                        o.copy_(self.out[off:off + n], non_blocking=True)
                    outs.append(o)
                    off += n
                self.meta_host[:16].copy_(self.meta, non_blocking=True)
                self.meta_host[16:].copy_(self.err, non_blocking=True)
                self._done = torch.cuda.Event()
                self._done.record(st)
            self._pending = (total, outs, st, cur)

    def finish_batch(self):
        total, outs, st, cur = self._pending
        self._pending = None
        self._done.synchronize()
        if st is not cur:
            for o in outs:
                if o.is_cuda:
                    o.record_stream(cur)
        mh = self.meta_host.tolist()
        if mh[0]:
            from .sparse import _raise_status
            _raise_status(mh[0])
        self._check_fusion_status(mh)
        self.levels = {1: total, 2: mh[2], 4: mh[3], 8: mh[4]}
        return outs

    def launch(self, *a, **k):
        raise NotImplementedError("use launch_batch / finish_batch")

    run = launch


@torch.no_grad()
def forward_batches(model, frags, batch: int, streams: int = 2, out=None, carry=None):
    """frags = [(coords int32 [N,4], feats fp32 [N,Cin], image fp32 [1,3,H,W]), ...], all on the host (ideally pinned) or all on
    the model's device.  Consecutive groups of `batch` fragments go through one `BatchGraphPlan` replay each, `streams` plans in
    flight (the copies of one group overlap the compute of the other); a trailing group smaller than `batch`, groups with images
    of different sizes and groups that overflow a plan's capacity take `forward_many_host` / `forward_many`.
    Returns the descriptors [N, out_channels] per fragment in order: device tensors for device inputs, pinned host tensors
    (`out[i]` when given) for host inputs.

    carry (a dict the caller keeps, initially empty): streaming use -- the call returns WITHOUT waiting for the groups still in flight,
    so the next call's host->device copies and first kernels overlap this call's last groups and their device->host copies; the
    entries of the returned list that belong to such groups are filled in (the same list object) when a later call or
    `drain_batches(carry)` retires them.  A plan (and with `out` its host buffers) is only reused after its previous group finished."""
    from . import me as ME
    if model.training:
        raise NotImplementedError("imfnet_b200 implements the eval-mode forward")
    model._ensure_plan()
    plan = model._plan
    if plan._key != plan._weights_key():
        plan.pack()
        model._graphs.clear()
    B = int(batch)
    if B < 1 or B > 255:
        raise ValueError("batch must be in [1, 255]")
    on_device = len(frags) > 0 and frags[0][0].is_cuda
    outs = [None] * len(frags)
    inflight = [] if carry is None else carry.setdefault("inflight", [])
    bucket = model.ROW_BUCKET

    def sequential(idx, fr=frags, dst=outs, o=out):
        if fr[idx[0]][0].is_cuda:
            items = [(ME.SparseTensor(fr[i][1], coordinates=fr[i][0]), fr[i][2]) for i in idx]
            res = [r.F for r in model.forward_many(items, streams=max(2, streams))]
        else:
            res = model.forward_many_host([fr[i] for i in idx], streams=max(2, streams), out=None if o is None else [o[i] for i in idx])
        for i, r in zip(idx, res):
            dst[i] = r

    def retire():
        idx, g, fr, dst, o = inflight.pop(0)          # (a group of an earlier streaming call carries its own lists)
        try:
            res = g.finish_batch()
        except PlanCapacityError:
            sequential(idx, fr, dst, o)
            return
        for i, r in zip(idx, res):
            dst[i] = r

    groups = len(frags) // B
    for gi in range(groups):
        idx = list(range(gi * B, (gi + 1) * B))
        sizes = [int(frags[i][0].shape[0]) for i in idx]
        shapes = {tuple(frags[i][2].shape[-2:]) for i in idx}
        if min(sizes) == 0 or len(shapes) != 1 or not model.use_cuda_graph:
            while inflight:
                retire()
            sequential(idx)
            continue
        H, W = shapes.pop()
        rows = (sum(sizes) + bucket - 1) // bucket * bucket
        item_rows = (max(sizes) + bucket - 1) // bucket * bucket
        key = ("batch", rows, int(H), int(W), B, item_rows)
        pool = model._graphs.setdefault(key, [])
        slot = gi % max(1, streams)
        while len(pool) <= slot:
            g = BatchGraphPlan(plan, rows, int(H), int(W), B, model._cap8(item_rows, 1))
            g.stream = torch.cuda.Stream(device=plan.device)
            pool.append(g)
        g = pool[slot]
        while any(e[1] is g for e in inflight):
            retire()
        dsts = None
        if not on_device:
            dsts = [out[i] if out is not None else torch.empty((sizes[j], model.out_channels), dtype=torch.float32, pin_memory=True)
                    for j, i in enumerate(idx)]
        g.launch_batch([(frags[i][0], frags[i][1].float(), frags[i][2].float()) for i in idx], g.stream, out_hosts=dsts)
        model._last_batch_plan = g          # (bench.py times this plan's layers in place)
        inflight.append((idx, g, frags, outs, out))
    tail = list(range(groups * B, len(frags)))
    if carry is None or tail:
        while inflight:
            retire()
    if tail:
        sequential(tail)
    if carry is not None:
        carry["retire"] = retire
    return outs


def drain_batches(carry):
    """Waits for every group a streaming `forward_batches(..., carry=carry)` sequence left in flight (their entries of the lists returned
    by those calls are filled in)."""
    inflight = carry.get("inflight", [])
    while inflight:
        carry["retire"]()
