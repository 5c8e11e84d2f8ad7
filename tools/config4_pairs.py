"""BASELINE config 3 (the "8 x B200, 64 fragment pairs" line): synthetic 3DMatch-shaped fragment pairs sharded over the ranks,
descriptors + 5000-keypoint mutual-NN matching per pair; prints one JSON line (pairs/s, whole job).

    python tools/config4_pairs.py [--pairs 8]                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/config4_pairs.py --pairs 64
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

import imfnet_b200.me as ME
from imfnet_b200 import load_model, synthetic
from imfnet_b200.pipeline import aggregate_throughput, describe_and_match_pairs, gather_records, shard_indices

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=8)
ap.add_argument("--config", default="C2")
ap.add_argument("--keypoints", type=int, default=5000)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
target, voxel, W, H = synthetic.CONFIGS[args.config]
model = load_model("ResUNetBN2C")(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3, config=None)
model.load_state_dict(synthetic.make_state_dict(0))
model = model.eval().cuda()
mine = shard_indices(args.pairs, rank, world)


def frag(seed):
    c, _ = synthetic.make_fragment(target, voxel, seed)
    return (ME.SparseTensor(torch.ones((len(c), 1)), coordinates=torch.from_numpy(c), device="cuda"), synthetic.make_image(W, H, seed).cuda())


pairs = [(frag(2 * p), frag(2 * p + 1)) for p in mine]
describe_and_match_pairs(model, pairs[:2], args.keypoints)          # warm-up (captures the plans)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
res = describe_and_match_pairs(model, pairs, args.keypoints)
e1.record()
torch.cuda.synchronize()
rec = torch.tensor([rank, len(pairs), e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
allrec = gather_records(rec, world).cpu()
if rank == 0:
    value, ms = aggregate_throughput(allrec)
    print(json.dumps({"workload": f"{args.pairs} synthetic {args.config} fragment pairs, descriptors + {args.keypoints}-keypoint mutual-NN matching",
                      "n_gpus": world, "pairs_per_s": value, "ms_slowest_rank": ms, "voxels_per_s": value * 2 * target,
                      "mutual_matches_first_pair": int(len(res[0]["mutual"])) if res else None}))
if world > 1:
    torch.distributed.destroy_process_group()
