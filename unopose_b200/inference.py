"""Inference driver of the hot path: instance batching, result composition, BOP-style result files.

Mirror of ``inference_and_save_oneref_v1`` (reference core/unopose/engine/oneref_inference_utils_v1.py:13-135):
same function name and arguments, same per-image flow, the same CSV line format

    scene_id,im_id,obj_id,score,R (9 values, space separated),t (3 values, millimetres),time

and the same ``.json`` dump of the detections with ``pred_R`` / ``pred_t`` filled in.

What is different:
* the reference walks the instances of an image in serial chunks of ``instance_batch_size`` on one GPU; here the
  instances of an image can additionally be **partitioned across ranks** (``torch.distributed``, one process per
  GPU): rank r takes a contiguous slice, runs its own chunks, and the per-instance result rows
  ``[R(9) | t(3) | score]`` are all-gathered (the one collective of the path, 52 bytes per instance).  This is
  OPT-IN (pass ``group=`` or ``shard_instances=True``) and requires that every rank iterates the SAME images, i.e. an
  unsharded loader — the reference's ``build_test_loader`` uses ``InferenceSampler``, which shards IMAGES across
  ranks; with that loader leave instance sharding off (each rank then poses its own images, like the reference).
  When it is on, the image identity and instance count are all-gathered per image and a mismatch raises;
* with instance sharding only rank 0 writes the files (the reference lets every rank write the same path);
* tensors are moved with ``.to(device)`` to the model's device, so the driver also runs on CPU tensors with a
  CPU model (the tests do that; the hot-path kernels themselves are CUDA-only).
"""
import json
import logging
import time
from copy import deepcopy
from pathlib import Path

import numpy as np
import torch

from . import dist as D

logger = logging.getLogger(__name__)

# per-instance tensors of a test sample (batch dim 1 in the loader), reference :56-66
_INSTANCE_KEYS = ("pts", "rgb", "rgb_choose", "fps_idx_m", "tem1_rgb", "tem1_choose", "tem1_pts", "fps_idx_o")


def instance_chunks(begin, end, bs):
    """Chunks [(s, e)] of at most `bs` instances covering [begin, end) — the reference's loop (:42-50), per rank."""
    return [(s, min(s + bs, end)) for s in range(begin, end, bs)]


def compose_with_template_pose(pred_R, pred_t, pose_ref_obj):
    """pose_tgt_obj = [R t; 0 1] @ pose_ref_obj (reference :82-89) -> (R (n,3,3), t (n,3))."""
    T = torch.zeros_like(pose_ref_obj)
    T[:, 3, 3] = 1.0
    T[:, :3, :3] = pred_R
    T[:, :3, 3] = pred_t
    T = T @ pose_ref_obj
    return T[:, :3, :3], T[:, :3, 3]


def format_result_line(scene_id, img_id, obj_id, score, R9, t3_mm, image_time):
    """One CSV line, byte-for-byte the reference's formatting (:116-127)."""
    return ",".join((str(scene_id), str(img_id), str(obj_id), str(score), " ".join(str(v) for v in R9),
                     " ".join(str(v) for v in t3_mm), f"{image_time}\n"))


def _model_device(model):
    try:
        return next(model.parameters()).device
    except (StopIteration, AttributeError):
        return torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


def _sync(device):
    if device.type == "cuda":
        torch.cuda.synchronize(device)


def _check_same_image(data, n_instance, device, group):
    """Instance sharding needs every rank to hold the same image: all-gather (scene, image, #instances) and compare."""
    import torch.distributed as dist

    ident = torch.tensor([int(data["scene_id"].item()) if "scene_id" in data else -1,
                          int(data["img_id"].item()) if "img_id" in data else -1, int(n_instance)],
                         dtype=torch.int64, device=device)
    world = dist.get_world_size(group)
    parts = [torch.empty_like(ident) for _ in range(world)]
    dist.all_gather(parts, ident, group=group)
    if any(not torch.equal(p, parts[0]) for p in parts):
        raise RuntimeError("instance sharding needs the same image on every rank (got %s): use an unsharded test loader, "
                           "or leave shard_instances off with the reference's InferenceSampler"
                           % [p.tolist() for p in parts])


def pose_image(model, data, instance_batch_size=16, device=None, group=None, shard_instances=None):
    """All instances of one test image -> (R (n,9), t_mm (n,3), score (n,)) numpy arrays (on every rank when sharded).

    `data` is one sample of the reference's test loader (image batch size 1).  With instance sharding (opt-in:
    `group` given or `shard_instances=True`) the instances are split across the ranks (contiguous balanced slices);
    each rank runs chunks of `instance_batch_size`."""
    device = device or _model_device(model)
    data = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in data.items()}
    n_instance = data["pts"].size(1)
    shard = (group is not None) if shard_instances is None else bool(shard_instances)
    rank, world = D._world(group) if shard else (0, 1)
    if world > 1:
        _check_same_image(data, n_instance, device, group)
    begin, end = D.shard_range(n_instance, rank, world)
    rows = []
    for s, e in instance_chunks(begin, end, instance_batch_size):
        inputs = {k: data[k][0][s:e].contiguous() for k in _INSTANCE_KEYS if k in data}
        with torch.no_grad():
            end_points = model(inputs)
        R, t = end_points["pred_R"], end_points["pred_t"]
        if "tem1_pose" in data:
            R, t = compose_with_template_pose(R, t, data["tem1_pose"][0][s:e].contiguous())
        rows.append(D.pack_results(R.float(), t.float(), end_points["pred_pose_score"].float()))
    dtype_dev = dict(dtype=torch.float32, device=device)
    local = torch.cat(rows, 0) if rows else torch.zeros((0, 13), **dtype_dev)
    if world > 1:
        sizes = [D.shard_range(n_instance, r, world)[1] - D.shard_range(n_instance, r, world)[0] for r in range(world)]
        local = D.all_gather_ragged(local, sizes, dim=0, group=group)
    pred_R = local[:, :9].detach().cpu().numpy()
    pred_t = local[:, 9:12].detach().cpu().numpy() * 1000       # metres -> millimetres (:96)
    score = (local[:, 12] * data["score"][0, :, 0].float()).detach().cpu().numpy()   # x detection score (:97)
    return pred_R, pred_t, score


def inference_and_save_oneref_v1(model, data_loader, save_path, instance_batch_size=16, group=None,
                                 shard_instances=None):
    """Drop-in for the reference driver (same arguments + an optional process group / instance-sharding switch).
    Without instance sharding every rank behaves exactly like the reference (poses the images its loader yields and
    writes `save_path`); with it, the loader must be unsharded and only rank 0 writes."""
    model.eval()
    device = _model_device(model)
    dets = deepcopy(data_loader.dataset.dets)
    shard = (group is not None) if shard_instances is None else bool(shard_instances)
    rank = D._world(group)[0] if shard else 0
    lines = []
    for i, data in enumerate(data_loader):
        _sync(device)
        t0 = time.perf_counter()
        pred_Rs, pred_Ts, pred_scores = pose_image(model, data, instance_batch_size, device, group, shard)
        _sync(device)
        image_time = time.perf_counter() - t0
        scene_id = data["scene_id"].item()
        img_id = data["img_id"].item()
        det_key = f"{scene_id:06d}_{img_id:06d}"
        inst_ids = data["inst_ids"][0].cpu().numpy()
        image_time += data["seg_time"].item()
        for k in range(pred_Rs.shape[0]):
            inst_i = int(inst_ids[k])
            dets[det_key][inst_i]["pred_R"] = pred_Rs[k].tolist()
            dets[det_key][inst_i]["pred_t"] = pred_Ts[k].tolist()
            lines.append(format_result_line(scene_id, img_id, data["obj_id"][0][k].item(), pred_scores[k],
                                            pred_Rs[k], pred_Ts[k], image_time))
    if rank == 0:
        with open(save_path, "w+") as f:
            f.writelines(lines)
        logger.info(f"saved to {save_path}")
        save_json_path = str(save_path).replace(".csv", ".json")
        Path(save_json_path).write_text(json.dumps(dets))
        logger.info(f"json saved to {save_json_path}")
    return lines
