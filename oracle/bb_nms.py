"""Oracle restatement of the best-buddy peak filter.

Test infrastructure (see ``oracle/__init__.py``).  Follows preprocessing_dino_bb/compute_dino_bb_nms.py; the greedy NMS
restates torchvision.ops.batched_nms (third-party, torchvision 0.17 in the reference's requirements; present in this image
as 0.26 and used by the live reference when the golden vectors are generated): boxes of different batch rows never
suppress each other, within a row boxes are visited by descending score and dropped when their IoU with an already kept
box exceeds the threshold.
"""
import torch


def _nms_row(boxes: torch.Tensor, scores: torch.Tensor, thr: float) -> torch.Tensor:
    """Greedy NMS of one row -> bool keep mask.  boxes K x 4 (x1, y1, x2, y2)."""
    order = torch.argsort(scores, descending=True, stable=True)
    area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    keep = torch.zeros(boxes.shape[0], dtype=torch.bool)
    suppressed = torch.zeros(boxes.shape[0], dtype=torch.bool)
    for i in order.tolist():
        if suppressed[i]:
            continue
        keep[i] = True
        xx1 = torch.maximum(boxes[i, 0], boxes[:, 0]); yy1 = torch.maximum(boxes[i, 1], boxes[:, 1])
        xx2 = torch.minimum(boxes[i, 2], boxes[:, 2]); yy2 = torch.minimum(boxes[i, 3], boxes[:, 3])
        inter = (xx2 - xx1).clamp(min=0) * (yy2 - yy1).clamp(min=0)
        iou = inter / (area[i] + area - inter)
        suppressed |= iou > thr
    return keep


def bb_sim_peaks(affs: torch.Tensor, coords: torch.Tensor, box_size=50, iou_thresh=0.5, topk=400):
    """compute_dino_bb_nms.py:12-47 (get_bb_sim_indices).  affs B x N -> (top2_values B x 2, r B)."""
    tk = torch.topk(affs, k=topk, sorted=False, dim=1)
    idx, vals = tk.indices, tk.values
    fc = coords[idx]                                                       # B x topk x 2
    boxes = torch.stack([fc[..., 0] - box_size, fc[..., 1] - box_size, fc[..., 0] + box_size, fc[..., 1] + box_size], dim=-1)
    mask = torch.stack([_nms_row(boxes[b], vals[b], iou_thresh) for b in range(affs.shape[0])]) if affs.shape[0] else \
        torch.zeros_like(vals, dtype=torch.bool)
    peak = vals * mask                                                     # dropped boxes count as 0
    top2 = torch.topk(peak, k=2, dim=1).values
    return top2, top2[:, 1] / top2[:, 0]


def compute_bb_nms(bb_sf_tf: dict, sf: int, tf: int, dino_emb: torch.Tensor, coords: torch.Tensor, stride=7, box_size=50,
                   iou_thresh=0.2, patch=14):
    """compute_dino_bb_nms.py:50-70: similarity maps of the pair's source points against frame tf, then the peak filter."""
    fxy = (bb_sf_tf["source_coords"] - (patch // 2)) / stride              # xy_to_fxy, dino_bb_utils.py:17-19
    target = dino_emb[tf]
    source_f = dino_emb[sf][:, fxy[:, 1].int(), fxy[:, 0].int()]           # C x N
    sim = torch.einsum("cn,chw->nhw", source_f, target)
    sim = sim / torch.clamp(source_f.norm(dim=0)[:, None, None] * target.norm(dim=0)[None], min=1e-08)
    peak, r = bb_sim_peaks(sim.reshape(sim.shape[0], -1), coords, box_size, iou_thresh)
    out = dict(bb_sf_tf)
    out["peak_coords"] = None
    out["peak_affs"] = peak
    out["r"] = r
    return out


def compute_max_r(bb: dict, bb_rev: dict):
    """compute_dino_bb_nms.py:72-82: a pair's r is the larger of the two directions' values."""
    for i in range(bb["target_coords"].shape[0]):
        rev = torch.norm(bb_rev["source_coords"] - bb["target_coords"][i][None], dim=1).argmin(0)
        assert torch.norm(bb_rev["target_coords"][rev] - bb["source_coords"][i]) == 0
        m = max(bb_rev["r"][rev], bb["r"][i])
        bb["r"][i] = m
        bb_rev["r"][rev] = m
    return bb, bb_rev
