"""Image-sharded data parallelism for the hot path.

Every stage of the path is per image (tf.map_fn over the batch, detector/utils/nms.py:55-60)
or per person (create_pb.py:96-109, detector/prn.py:17-24), so images shard across
ranks with NO collective on the data path: rank r of G runs its own contiguous block of
images through its own Detector, and only the small result blocks are concatenated on
the host in rank order (gloo / object gather).  PRN weights are replicated per GPU.
"""
import numpy as np
import torch.distributed as dist


def shard_range(n_items, rank, world_size):
    """Contiguous block [lo, hi) of `n_items` owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_inputs(inputs, rank, world_size):
    """Slice every batched array of an input dict to this rank's images."""
    B = inputs["class_logits"].shape[0]
    lo, hi = shard_range(B, rank, world_size)
    out = dict(inputs)
    for k in ("class_logits", "encoded_boxes", "heatmap_logits"):
        out[k] = inputs[k][lo:hi]
    if "gt_boxes" in inputs:
        out["gt_boxes"] = inputs["gt_boxes"][lo:hi]
    return out, (lo, hi)


def merge_results(per_rank):
    """Concatenate per-rank result dicts (create_pb.py:53-61 layout) in rank order."""
    per_rank = [r for r in per_rank if r is not None and r["num_boxes"].shape[0] > 0]
    out = {}
    for k in ("boxes", "scores", "num_boxes", "keypoint_scores", "keypoint_positions"):
        out[k] = np.concatenate([r[k] for r in per_rank], axis=0)
    for k in ("keypoint_heatmaps", "segmentation_masks"):
        if all(k in r for r in per_rank):
            out[k] = np.concatenate([r[k] for r in per_rank], axis=0)
    offs = [0]
    for r in per_rank:
        nb = r["num_boxes"].astype(np.int64)
        for n in nb:
            offs.append(offs[-1] + int(n))
    out["person_offsets"] = np.asarray(offs, dtype=np.int32)
    return out


def gather_results(result, dst=0, group=None):
    """Host-side gather of one rank's result dict to `dst`; returns the merged dict there, None elsewhere.
    Uses an object gather (pickled numpy), i.e. no device collective is involved."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bucket = [None] * world if rank == dst else None
    dist.gather_object(result, bucket, dst=dst, group=group)
    if rank != dst:
        return None
    return merge_results(bucket)


_PACKED = ("boxes", "scores", "num_boxes", "keypoint_scores", "keypoint_positions")


def gather_packed(result, dst=0, group=None, merge=True):
    """The same gather for a stream of steps: every rank's detections and keypoints (the padded, fixed-size buffers of one
    Detector call: boxes [B, max_det, 4], scores [B, max_det], num_boxes [B], keypoint_scores / keypoint_positions
    [B * max_det, ...]) travel as ONE fixed-size block per rank in one host-side `gather` (gloo) -- no pickling, no device
    collective.  `result` holds torch CPU tensors or numpy arrays; returns the merged dict (create_pb.py:53-61 layout, persons
    of all ranks concatenated in rank order) on `dst`, None elsewhere."""
    import torch
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    parts = [torch.as_tensor(result[k]) for k in _PACKED]
    flat = torch.cat([p.reshape(-1).view(torch.float32) for p in parts])      # num_boxes: int32 bits carried as float32
    bucket = [torch.empty_like(flat) for _ in range(world)] if rank == dst else None
    dist.gather(flat, bucket, dst=dst, group=group)
    if rank != dst:
        return None
    if not merge:
        return bucket
    per_rank = []
    for blk in bucket:
        r, o = {}, 0
        for k, p in zip(_PACKED, parts):
            n = p.numel()
            a = blk[o:o + n].view(p.dtype).reshape(p.shape).numpy()
            o += n
            r[k] = a
        n_persons = int(r["num_boxes"].sum())
        r["keypoint_scores"] = r["keypoint_scores"][:n_persons]
        r["keypoint_positions"] = r["keypoint_positions"][:n_persons]
        per_rank.append(r)
    return merge_results(per_rank)
