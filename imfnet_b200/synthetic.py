"""Seeded synthetic inputs and weights for tests and bench (no datasets / checkpoints offline).

* `make_fragment`      -- 3DMatch-like voxel set (SURVEY.md section 8d recipe: random rectangles in a room,
                           3 mm noise, first-occurrence quantisation, truncated to exactly `target` voxels).
* `state_dict_spec`    -- names/shapes of the reference `ResUNet*` state_dict (model/resunet.py:25-161,
                           model/residual_block.py:23-35, model/attention_fusion.py:99-130, model/resnet.py:120-160).
* `make_state_dict`    -- numpy-PCG64 seeded values per tensor name (platform independent, unlike torch init),
                           BN statistics randomised so folding errors are visible.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict

import numpy as np
import torch

CONFIGS = {
    # name: (target voxels, voxel size [m], image W, image H)   -- BASELINE.json configs / SURVEY.md 8d
    "C1": (5000, 0.05, 160, 120),
    "C2": (50000, 0.025, 640, 480),
    "C3": (120000, 0.025, 1226, 370),
    "C5": (200000, 0.025, 640, 480),
}

VARIANTS = {
    "ResUNetBN2": ([None, 32, 64, 128, 256], [None, 32, 64, 64, 128]),
    "ResUNetBN2B": ([None, 32, 64, 128, 256], [None, 64, 64, 64, 64]),
    "ResUNetBN2C": ([None, 32, 64, 128, 256], [None, 64, 64, 64, 128]),
    "ResUNetBN2D": ([None, 32, 64, 128, 256], [None, 64, 64, 128, 128]),
    "ResUNetBN2E": ([None, 128, 128, 128, 256], [None, 64, 128, 128, 128]),
}


def _first_unique(keys: np.ndarray) -> np.ndarray:
    _, first = np.unique(keys, return_index=True)
    return np.sort(first)


def make_points(target: int, voxel: float, seed: int = 0) -> np.ndarray:
    """Raw float64 xyz points whose quantisation at `voxel` has at least `target` distinct voxels."""
    rng = np.random.default_rng(seed)
    ext = np.array([3.0, 3.0, 2.6])
    chunks, seen = [], np.zeros((0,), dtype=np.int64)
    while len(seen) < target:
        o = rng.uniform(0, 1, 3) * ext
        if rng.uniform() < 0.7:
            a = np.zeros(3)
            a[rng.integers(3)] = 1.0
        else:
            a = rng.normal(size=3)
            a /= np.linalg.norm(a)
        b = np.cross(a, rng.normal(size=3))
        b /= np.linalg.norm(b)
        la, lb = rng.uniform(0.4, 2.0, 2)
        n = int(la * lb * 6 / voxel ** 2)
        uv = rng.uniform(0, 1, (n, 2))
        p = o + uv[:, :1] * la * a + uv[:, 1:] * lb * b + rng.normal(0, 0.003, (n, 3))
        chunks.append(p)
        q = np.floor(p / voxel).astype(np.int64) + (1 << 15)
        keys = (q[:, 0] << 32) | (q[:, 1] << 16) | q[:, 2]
        seen = np.union1d(seen, keys)
    return np.concatenate(chunks, axis=0)


def make_fragment(target: int, voxel: float, seed: int = 0, batch_index: int = 0):
    """-> (coords int32 [target,4] (b,x,y,z) in first-occurrence order, xyz float64 [target,3] representative points)."""
    pts = make_points(target, voxel, seed)
    q = np.floor(pts / voxel).astype(np.int64)
    qb = q + (1 << 15)
    keys = (qb[:, 0] << 32) | (qb[:, 1] << 16) | qb[:, 2]
    idx = _first_unique(keys)[:target]
    coords = np.concatenate([np.full((len(idx), 1), batch_index, dtype=np.int64), q[idx]], axis=1).astype(np.int32)
    return coords, pts[idx]


def make_image(W: int, H: int, seed: int = 0, batch: int = 1) -> torch.Tensor:
    """Uniform [0,1) float32 image like a PNG read by matplotlib (scripts/generate_desc.py:87-97)."""
    rng = np.random.default_rng(seed + 7919)
    return torch.from_numpy(rng.random((batch, 3, H, W), dtype=np.float32))


def make_config(name: str, seed: int = 0):
    target, voxel, W, H = CONFIGS[name]
    coords, _ = make_fragment(target, voxel, seed)
    feats = torch.ones((len(coords), 1), dtype=torch.float32)
    return torch.from_numpy(coords), feats, make_image(W, H, seed)


# ----------------------------------------------------------------------------------------------
def _bn_spec(p, c):
    return [(p + ".weight", (c,), "bn_w"), (p + ".bias", (c,), "bn_b"), (p + ".running_mean", (c,), "bn_m"),
            (p + ".running_var", (c,), "bn_v"), (p + ".num_batches_tracked", (), "nbt")]


def _resnet34_spec(p, in_ch=3):
    spec = [(p + ".conv1.weight", (64, in_ch, 7, 7), "conv2d")] + _bn_spec(p + ".bn1", 64)
    inpl = 64
    for li, (planes, n) in enumerate([(64, 3), (128, 4), (256, 6), (512, 3)], start=1):
        for bi in range(n):
            q = f"{p}.layer{li}.{bi}"
            spec += [(q + ".conv1.weight", (planes, inpl if bi == 0 else planes, 3, 3), "conv2d")] + _bn_spec(q + ".bn1", planes)
            spec += [(q + ".conv2.weight", (planes, planes, 3, 3), "conv2d")] + _bn_spec(q + ".bn2", planes)
            if bi == 0 and li > 1:
                spec += [(q + ".downsample.0.weight", (planes, inpl, 1, 1), "conv2d")] + _bn_spec(q + ".downsample.1", planes)
        inpl = planes
    spec += [(p + ".fc.weight", (1000, 512), "linear"), (p + ".fc.bias", (1000,), "linear_b:512")]
    return spec


def state_dict_spec(model: str = "ResUNetBN2C", in_channels: int = 1, out_channels: int = 32, conv1_kernel_size: int = 5):
    """[(name, shape, kind)] of the reference model's state_dict."""
    CH, TR = VARIANTS[model]
    spec = []

    def conv(p, k, cin, cout, tr=False):
        kv = k ** 3
        shape = (cin, cout) if kv == 1 else (kv, cin, cout)
        spec.append((p + ".kernel", shape, f"meconv:{(cout if tr else cin) * kv}"))

    def block(p, c):
        conv(p + ".conv1", 3, c, c)
        spec.extend(_bn_spec(p + ".norm1.bn", c))
        conv(p + ".conv2", 3, c, c)
        spec.extend(_bn_spec(p + ".norm2.bn", c))

    conv("conv1", conv1_kernel_size, in_channels, CH[1]); spec.extend(_bn_spec("norm1.bn", CH[1])); block("block1", CH[1])
    conv("conv2", 3, CH[1], CH[2]); spec.extend(_bn_spec("norm2.bn", CH[2])); block("block2", CH[2])
    conv("conv3", 3, CH[2], CH[3]); spec.extend(_bn_spec("norm3.bn", CH[3])); block("block3", CH[3])
    conv("conv4", 3, CH[3], CH[4]); spec.extend(_bn_spec("norm4.bn", CH[4])); block("block4", CH[4])
    a = "attention_fusion.cross_attend_blocks"
    lat, dim, inner = CH[4], 128, int(CH[4] / 2)
    spec += [(f"{a}.0.fn.to_q.weight", (inner, lat), "linear_q"), (f"{a}.0.fn.to_kv.weight", (2 * inner, dim), "linear"),
             (f"{a}.0.fn.to_out.weight", (lat, inner), "linear"), (f"{a}.0.fn.to_out.bias", (lat,), f"linear_b:{inner}"),
             (f"{a}.0.norm.weight", (lat,), "ln_w"), (f"{a}.0.norm.bias", (lat,), "ln_b"),
             (f"{a}.0.norm_context.weight", (dim,), "ln_w"), (f"{a}.0.norm_context.bias", (dim,), "ln_b"),
             (f"{a}.1.fn.net.0.weight", (lat * 8, lat), "linear"), (f"{a}.1.fn.net.0.bias", (lat * 8,), f"linear_b:{lat}"),
             (f"{a}.1.fn.net.2.weight", (lat, lat * 4), "linear"), (f"{a}.1.fn.net.2.bias", (lat,), f"linear_b:{lat * 4}"),
             (f"{a}.1.norm.weight", (lat,), "ln_w"), (f"{a}.1.norm.bias", (lat,), "ln_b")]
    conv("conv4_tr", 3, CH[4], TR[4], True); spec.extend(_bn_spec("norm4_tr.bn", TR[4])); block("block4_tr", TR[4])
    conv("conv3_tr", 3, CH[3] + TR[4], TR[3], True); spec.extend(_bn_spec("norm3_tr.bn", TR[3])); block("block3_tr", TR[3])
    conv("conv2_tr", 3, CH[2] + TR[3], TR[2], True); spec.extend(_bn_spec("norm2_tr.bn", TR[2])); block("block2_tr", TR[2])
    conv("conv1_tr", 1, CH[1] + TR[2], TR[1])
    conv("final", 1, TR[1], out_channels)
    spec.append(("final.bias", (1, out_channels), f"linear_b:{TR[1]}"))
    spec += _resnet34_spec("img_encoder.backbone")
    return spec


def make_state_dict(seed: int = 0, model: str = "ResUNetBN2C", in_channels: int = 1, out_channels: int = 32,
                    conv1_kernel_size: int = 5) -> "OrderedDict[str, torch.Tensor]":
    sd = OrderedDict()
    for name, shape, kind in state_dict_spec(model, in_channels, out_channels, conv1_kernel_size):
        rng = np.random.default_rng([seed, zlib.crc32(name.encode())])
        if kind == "nbt":
            sd[name] = torch.tensor(0, dtype=torch.int64)
            continue
        if kind.startswith("meconv:"):          # ME reset_parameters: U(-1/sqrt(n), 1/sqrt(n))
            b = 1.0 / np.sqrt(float(kind.split(":")[1]))
            v = rng.uniform(-b, b, shape)
        elif kind == "conv2d":                  # kaiming_normal_(fan_out, relu), model/resnet.py:153-155
            v = rng.normal(0, np.sqrt(2.0 / (shape[0] * shape[2] * shape[3])), shape)
        elif kind in ("linear", "linear_q"):    # nn.Linear default: U(+-1/sqrt(fan_in)); to_q x4 so softmax is not flat
            b = 1.0 / np.sqrt(shape[1])
            v = rng.uniform(-b, b, shape) * (4.0 if kind == "linear_q" else 1.0)
        elif kind.startswith("linear_b:"):
            b = 1.0 / np.sqrt(float(kind.split(":")[1]))
            v = rng.uniform(-b, b, shape)
        elif kind in ("bn_w", "bn_v", "ln_w"):
            v = rng.uniform(0.5, 1.5, shape)
        elif kind in ("bn_b", "bn_m", "ln_b"):
            v = rng.normal(0, 0.1, shape)
        else:
            raise ValueError(kind)
        sd[name] = torch.from_numpy(np.asarray(v, dtype=np.float32).reshape(shape))
    return sd


def scaled_state_dict(sd, s_point: float = 1.0, s_image: float = 1.0):
    """A checkpoint whose ACTIVATIONS are s_point (point branch) / s_image (image branch) times those of `sd`, layer by layer: every
    BatchNorm gets gamma, beta times s and -- except the first layer of a branch, whose input is unscaled -- running_mean times s and
    running_var times s^2; the fusion module's output projections (to_out, net.2) are scaled by s_point so that the fused stride-8
    features stay at the branch's magnitude.  Used to test the numeric range of the fp16 hi/lo tier (activations near 1e-6 or 1e4)
    against the fp32 oracle on the same weights."""
    out = OrderedDict((k, v.clone()) for k, v in sd.items())

    def scale_bn(prefix, s, first):
        out[prefix + ".weight"] *= s
        out[prefix + ".bias"] *= s
        if not first:
            out[prefix + ".running_mean"] *= s
            out[prefix + ".running_var"] *= s * s

    for k in sd:
        if not k.endswith(".running_var"):
            continue
        p = k[: -len(".running_var")]
        if p.startswith("img_encoder.backbone."):
            if p.startswith("img_encoder.backbone.layer3") or p.startswith("img_encoder.backbone.layer4"):
                continue                                   # never executed
            scale_bn(p, s_image, first=(p == "img_encoder.backbone.bn1"))
        else:
            scale_bn(p, s_point, first=(p == "norm1.bn"))
    for k in ("attention_fusion.cross_attend_blocks.0.fn.to_out.weight", "attention_fusion.cross_attend_blocks.0.fn.to_out.bias",
              "attention_fusion.cross_attend_blocks.1.fn.net.2.weight", "attention_fusion.cross_attend_blocks.1.fn.net.2.bias"):
        out[k] *= s_point
    return out
