"""``SSCMetrics`` -- the scores computed right after the volumetric path:
same class name, buffers (state_dict keys), method names and results as the reference's torchmetrics
module (projects/mmdet3d_plugin/utils/ssc_metric.py:14-168), with the 20-class loop of full-volume
boolean reductions replaced by ONE pass of ``ss_ssc_confusion_fwd`` over the uint8 label volume the
trilinear + argmax kernel already produced.  Per-class tp/fp/fn are row / column sums of the C x C
confusion matrix (400 numbers, host-side torch ops on a tiny tensor).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops

SEMKITTI_CLASS_NAMES = [
    "unlabeled", "car", "bicycle", "motorcycle", "truck", "other-vehicle", "person", "bicyclist", "motorcyclist", "road",
    "parking", "sidewalk", "other-ground", "building", "fence", "vegetation", "trunk", "terrain", "pole", "traffic-sign"]


class SSCMetrics(nn.Module):
    def __init__(self, class_names=None, compute_on_step=False):
        super().__init__()
        self.class_names = list(class_names) if class_names is not None else list(SEMKITTI_CLASS_NAMES)
        self.n_classes = len(self.class_names)
        for name in ("tps", "fps", "fns"):
            self.register_buffer(name, torch.zeros(self.n_classes))
        for name in ("completion_tp", "completion_fp", "completion_fn"):
            self.register_buffer(name, torch.zeros(1))

    # ---- one kernel pass -> (completion[3], tps, fps, fns) as int64 tensors on the inputs' device ----
    def scores(self, y_pred, y_true, nonempty=None, nonsurface=None):
        C = self.n_classes
        counts = ops.ssc_confusion(y_pred, y_true, C, nonempty=nonempty, nonsurface=nonsurface)
        M = counts[:C * C].view(C, C)                     # [target][prediction]
        tps = M.diagonal().clone()
        # predictions outside the class range are misses of their target class and false positives of none (ssc_metric.py:157-163)
        return counts[C * C:C * C + 3], tps, M.sum(0) - tps, M.sum(1) - tps + counts[C * C + 3:]

    def compute_single(self, y_pred, y_true, nonempty=None, nonsurface=None):
        comp, tps, fps, fns = self.scores(y_pred, y_true, nonempty, nonsurface)
        c = comp.cpu().numpy()
        return (c[0], c[1], c[2], tps.float().cpu().numpy(), fps.float().cpu().numpy(), fns.float().cpu().numpy())

    def update(self, y_pred, y_true, nonempty=None, nonsurface=None):
        comp, tps, fps, fns = self.scores(y_pred, y_true, nonempty, nonsurface)
        comp = comp.to(self.completion_tp)
        self.completion_tp += comp[0:1]
        self.completion_fp += comp[1:2]
        self.completion_fn += comp[2:3]
        self.tps += tps.to(self.tps)
        self.fps += fps.to(self.fps)
        self.fns += fns.to(self.fns)

    def get_score_completion(self, predict, target, nonempty=None):
        """Reference signature (ssc_metric.py:109-141): ``nonempty`` is the full selection mask there."""
        comp, _, _, _ = self.scores(predict, target, nonempty, None)
        return comp[0], comp[1], comp[2]

    def get_score_semantic_and_completion(self, predict, target, nonempty=None):
        _, tps, fps, fns = self.scores(predict, target, nonempty, None)
        return tps.float(), fps.float(), fns.float()

    def synced_state(self):
        """The six accumulators summed over the ranks of the default process group (the reference's torchmetrics states use
        dist_reduce_fx='sum', ssc_metric.py:24-35); the local buffers are left untouched, so compute() may be called repeatedly."""
        names = ("completion_tp", "completion_fp", "completion_fn", "tps", "fps", "fns")
        vals = [getattr(self, n) for n in names]
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            flat = torch.cat([v.reshape(-1).double() for v in vals])
            torch.distributed.all_reduce(flat)
            out, off = [], 0
            for v in vals:
                out.append(flat[off:off + v.numel()].to(v.dtype).view_as(v))
                off += v.numel()
            vals = out
        return dict(zip(names, vals))

    def compute(self):
        s = self.synced_state()
        ctp, cfp, cfn, tps, fps, fns = (s[k] for k in ("completion_tp", "completion_fp", "completion_fn", "tps", "fps", "fns"))
        precision = ctp / (ctp + cfp)
        recall = ctp / (ctp + cfn)
        iou = ctp / (ctp + cfp + cfn)
        iou_ssc = tps / (tps + fps + fns + 1e-5)
        return {"precision": precision, "recall": recall, "iou": iou.item(), "iou_ssc": iou_ssc,
                "iou_ssc_mean": iou_ssc[1:].mean().item()}

    def reset(self):
        for name in ("tps", "fps", "fns", "completion_tp", "completion_fp", "completion_fn"):
            getattr(self, name).zero_()
