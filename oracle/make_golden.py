"""Generate tests/golden/* by running the reference's UNMODIFIED model files (build container only).

    python oracle/make_golden.py            # needs /root/reference (read-only); writes tests/golden/

What it does
  1. loads /root/reference/model/{resunet,residual_block,common,attention_fusion,Img_Encoder,resnet}.py as-is,
     with `MinkowskiEngine` resolved to oracle/standin (ME 0.5.4 is an un-vendored pip dependency,
     requirements.txt:5) and two import stubs (torchvision.models.utils; pretrained=False);
  2. checks imfnet_b200.synthetic.state_dict_spec against the reference model's state_dict (names+shapes);
  3. runs reference forward vs oracle/imfnet_oracle.py on (a) files/cloud_bin_0.ply @5 cm + 160x120 image
     [BASELINE config 0], (b) a seeded 2-fragment synthetic batch; records max abs differences;
  4. extracts the sparse_quantize golden vector pinned by files/3D_head_map.ply (prefix of the cloud);
  5. runs model/attention_fusion.py alone (stress shape reduced) for the attention golden.
Fixtures are inputs + reference outputs (fp32), small enough to commit; weights are regenerated from the
numpy seed at test time (imfnet_b200.synthetic.make_state_dict).
"""
from __future__ import annotations

import importlib
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def load_reference_model_package():
    """Make `import model` resolve to /root/reference/model with ME -> stand-in."""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "standin"))
    sys.path.insert(0, ROOT)
    stub = types.ModuleType("torchvision.models.utils")
    stub.load_state_dict_from_url = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("offline"))
    sys.modules["torchvision.models.utils"] = stub
    sys.path.insert(0, REF)
    import model.resnet as ref_resnet   # noqa
    orig = ref_resnet._resnet
    ref_resnet._resnet = lambda in_ch, arch, block, layers, pretrained, progress, **kw: orig(
        in_ch, arch, block, layers, False, progress, **kw)       # Img_Encoder.py:13 hard-codes pretrained=True
    import model as ref_model           # noqa
    return ref_model


def read_ply_xyz(path):
    with open(path, "rb") as f:
        header = b""
        while not header.endswith(b"end_header\n"):
            header += f.readline()
        lines = header.decode().split("\n")
        n = int([l for l in lines if l.startswith("element vertex")][0].split()[-1])
        props = [l.split()[1:] for l in lines if l.startswith("property")]
        dt = np.dtype([(name, {"float": "<f4", "double": "<f8", "uchar": "u1"}[t]) for t, name in props])
        data = np.frombuffer(f.read(n * dt.itemsize), dtype=dt, count=n)
    return np.stack([data["x"], data["y"], data["z"]], axis=1)


def read_png_like_matplotlib(path):
    import cv2
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
    return img.astype(np.float32) / 255.0      # matplotlib.image.imread semantics for 8-bit PNG


def rel_err(a, b):
    return float((torch.linalg.norm(a - b, dim=1) / torch.linalg.norm(b, dim=1)).max())


def main():
    os.makedirs(OUT, exist_ok=True)
    ref_model = load_reference_model_package()
    import MinkowskiEngine as ME
    from imfnet_b200 import synthetic
    from oracle import imfnet_oracle

    manifest = {"generator": "oracle/make_golden.py", "torch": torch.__version__, "numpy": np.__version__}
    torch.set_num_threads(8)

    # ---- 2. state_dict contract --------------------------------------------------------------------
    Model = ref_model.load_model("ResUNetBN2C")
    model = Model(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3, config=None)
    ref_sd = model.state_dict()
    spec = synthetic.state_dict_spec("ResUNetBN2C", 1, 32, 5)
    assert {n: tuple(s) for n, s, _ in spec} == {k: tuple(v.shape) for k, v in ref_sd.items()}, "state_dict spec mismatch"
    manifest["state_dict"] = {"entries": len(ref_sd), "elements": int(sum(v.numel() for v in ref_sd.values()))}
    for name in ("ResUNetBN2", "ResUNetBN2B", "ResUNetBN2D", "ResUNetBN2E"):
        m2 = ref_model.load_model(name)(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3)
        s2 = synthetic.state_dict_spec(name, 1, 32, 5)
        assert {n: tuple(s) for n, s, _ in s2} == {k: tuple(v.shape) for k, v in m2.state_dict().items()}, name
    sd = synthetic.make_state_dict(0)
    model.load_state_dict(sd, strict=True)
    model.eval()

    def run_reference(coords, feats, image):
        with torch.no_grad():
            st = ME.SparseTensor(feats, coordinates=coords)
            return model(st, image).F

    # ---- 3a. real fragment, BASELINE config 0: 5 cm voxels + 160x120 image ---------------------------
    import cv2
    xyz = read_ply_xyz(os.path.join(REF, "files", "cloud_bin_0.ply")).astype(np.float64)
    c, inds = ME.utils.sparse_quantize(np.floor(xyz / 0.05), return_index=True)       # util/misc.py:82-83
    coords = ME.utils.batched_coordinates([c])                                          # util/misc.py:86
    feats = torch.ones((len(coords), 1), dtype=torch.float32)
    img = read_png_like_matplotlib(os.path.join(REF, "files", "cloud_bin_0_0.png"))
    img = np.asarray(cv2.resize(img, dsize=(160, 120), interpolation=cv2.INTER_LINEAR), dtype=np.float32)  # util/uio.py:31-40
    image = torch.from_numpy(np.transpose(img, (2, 0, 1))[None].copy())
    d_ref = run_reference(coords, feats, image)
    d_orc, acts = imfnet_oracle.forward(sd, coords, feats, image, return_intermediates=True)
    manifest["c1_real"] = {"voxels": int(len(coords)), "image": [120, 160], "oracle_vs_reference_max_abs": float((d_ref - d_orc).abs().max()),
                           "oracle_vs_reference_rel": rel_err(d_orc, d_ref),
                           "levels": [int(len(acts["levels"][t])) for t in (1, 2, 4, 8)]}
    np.savez_compressed(os.path.join(OUT, "c1_real.npz"), coords=coords.numpy(), image=image.numpy().astype(np.float16),
                        desc=d_ref.numpy(), fused=acts["fused"].numpy()[::4], out_s1=acts["out_s1"].numpy()[::16])
    # the image is stored as fp16 to stay small: tests feed float32(image_fp16) to BOTH sides; regenerate d_ref on that
    image16 = image.half().float()
    d_ref16 = run_reference(coords, feats, image16)
    d_orc16, acts16 = imfnet_oracle.forward(sd, coords, feats, image16, return_intermediates=True)
    np.savez_compressed(os.path.join(OUT, "c1_real.npz"), coords=coords.numpy(), image=image.numpy().astype(np.float16),
                        desc=d_ref16.numpy(), fused=acts16["fused"].numpy(), s8_coords=acts16["levels"][8])
    manifest["c1_real"]["oracle_vs_reference_max_abs_fp16img"] = float((d_ref16 - d_orc16).abs().max())

    # ---- 3b. synthetic batch of two fragments (batched layout, lib/data_loaders.py:68-69) -------------
    ca, _ = synthetic.make_fragment(1500, 0.05, seed=11)
    cb, _ = synthetic.make_fragment(1100, 0.05, seed=12)
    bc, bf = ME.utils.sparse_collate([ca[:, 1:], cb[:, 1:]], [np.ones((len(ca), 1), np.float32), np.ones((len(cb), 1), np.float32)])
    rng = np.random.default_rng(5)
    bf = torch.from_numpy(rng.uniform(0.5, 1.5, (len(bc), 1)).astype(np.float32))    # non-constant features
    bimg = synthetic.make_image(96, 64, seed=3, batch=2)
    d_ref = run_reference(bc, bf, bimg)
    d_orc = imfnet_oracle.forward(sd, bc, bf, bimg)
    manifest["batch2"] = {"voxels": int(len(bc)), "oracle_vs_reference_max_abs": float((d_ref - d_orc).abs().max()),
                          "oracle_vs_reference_rel": rel_err(d_orc, d_ref)}
    np.savez_compressed(os.path.join(OUT, "batch2.npz"), coords=bc.numpy(), feats=bf.numpy(), image=bimg.numpy().astype(np.float16),
                        desc=run_reference(bc, bf, bimg.half().float()).numpy())

    # ---- 4. quantisation golden: files/3D_head_map.ply == xyz[first-occurrence idx] @ 2.5 cm ----------
    head = read_ply_xyz(os.path.join(REF, "files", "3D_head_map.ply"))
    c25, inds25 = ME.utils.sparse_quantize(np.floor(xyz / 0.025), return_index=True)
    assert len(inds25) == len(head) and np.abs(xyz[inds25] - head).max() == 0.0, "3D_head_map.ply does not match unique_first"
    P = 60000          # prefix of the cloud: first-occurrence indices of a prefix are the golden indices < P
    pref_idx = inds25[inds25 < P]
    gold_pts = head[: len(pref_idx)]
    assert np.array_equal(xyz[pref_idx], gold_pts)
    np.savez_compressed(os.path.join(OUT, "quantize_prefix.npz"), xyz=xyz[:P].astype(np.float32),
                        head_map_vertices=gold_pts.astype(np.float32), voxel=np.float64(0.025))
    manifest["quantize"] = {"cloud_points": int(len(xyz)), "voxels_full": int(len(head)), "prefix_points": P,
                            "prefix_voxels": int(len(pref_idx)), "full_cloud_max_abs_diff": 0.0}

    # ---- 5. attention fusion alone (model/attention_fusion.py as-is) ----------------------------------
    af = model.attention_fusion
    rng = np.random.default_rng(9)
    Pq = torch.from_numpy(rng.normal(0, 1, (1, 333, 256)).astype(np.float32))
    Ic = torch.from_numpy(rng.normal(0, 1, (1, 300, 128)).astype(np.float32))
    with torch.no_grad():
        a_ref = af(Ic, queries_encoder=Pq)
    a_orc = imfnet_oracle.attention_fusion(sd, Ic, Pq)
    manifest["attention"] = {"M": 333, "L": 300, "oracle_vs_reference_max_abs": float((a_ref - a_orc).abs().max())}
    np.savez_compressed(os.path.join(OUT, "attention.npz"), queries=Pq.numpy(), data=Ic.numpy(), out=a_ref.numpy())

    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print(json.dumps(manifest, indent=1))


if __name__ == "__main__":
    main()
