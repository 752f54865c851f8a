"""``BEVDepthOccupancy`` -- thin re-host of the reference detector's orchestration
(projects/mmdet3d_plugin/occupancy/detectors/bevdepth_occupancy.py:23-297 on top of
bevdepth.py:14-34) around the B200 volumetric modules.

The 2-D image encoder (CustomEfficientNet-B7 + SECONDFPN, stereoscene.py:59-74) is upstream of
the accelerated path and out of scope: the config entries are accepted
and kept, an externally built encoder can be attached with ``set_image_encoder``, and the
volumetric path is entered with backbone features through ``forward_features`` -- the call
``bench.py`` and the parity tests make.
"""
from __future__ import annotations

import collections

import torch
import torch.nn as nn

from .. import ops
from ..ops import Vol
from ..registry import BACKBONES, DETECTORS, HEADS, NECKS, build_backbone, build_head, build_neck


class _ExternalComponent(nn.Module):
    """Placeholder for a config entry whose implementation lives outside the accelerated path."""

    def __init__(self, **cfg):
        super().__init__()
        self.cfg = cfg

    def forward(self, *a, **k):
        raise NotImplementedError(
            f"{type(self).__name__} (2-D image encoder) is out of scope for the volumetric hot path; attach an "
            "implementation with BEVDepthOccupancy.set_image_encoder() or call forward_features() with backbone features")


@BACKBONES.register_module()
class CustomEfficientNet(_ExternalComponent):
    pass


@NECKS.register_module()
class SECONDFPN(_ExternalComponent):
    pass


@DETECTORS.register_module()
class BEVDepthOccupancy(nn.Module):
    def __init__(self, img_view_transformer=None, img_bev_encoder_backbone=None, img_bev_encoder_neck=None,
                 img_backbone=None, img_neck=None, pts_bbox_head=None, loss_cfg=None, use_grid_mask=False,
                 disable_loss_depth=False, train_cfg=None, test_cfg=None, pretrained=None, init_cfg=None, **kwargs):
        super().__init__()
        self.img_backbone = build_backbone(img_backbone) if img_backbone is not None else None
        self.img_neck = build_neck(img_neck) if img_neck is not None else None
        self.img_view_transformer = build_neck(img_view_transformer)
        self.img_bev_encoder_backbone = build_backbone(img_bev_encoder_backbone)
        self.img_bev_encoder_neck = build_neck(img_bev_encoder_neck)
        head = dict(pts_bbox_head)
        head.setdefault("train_cfg", train_cfg.get("pts") if isinstance(train_cfg, dict) else None)
        head.setdefault("test_cfg", test_cfg.get("pts") if isinstance(test_cfg, dict) else None)
        self.pts_bbox_head = build_head(head)
        self.loss_cfg, self.use_grid_mask, self.disable_loss_depth = loss_cfg, use_grid_mask, disable_loss_depth
        self.record_time = False
        self.time_stats = collections.defaultdict(list)
        self._image_encoder = None

    @property
    def with_img_neck(self):
        return self.img_neck is not None

    def set_image_encoder(self, fn):
        """fn(imgs[B*N,3,H,W]) -> features [B*N,C,fH,fW]."""
        self._image_encoder = fn

    def image_encoder(self, img):
        B, N, Cc, H, W = img.shape
        if self._image_encoder is None:
            raise NotImplementedError("no image encoder attached (out of scope); use forward_features()")
        x = self._image_encoder(img.view(B * N, Cc, H, W))
        return x.view(B, N, *x.shape[1:])

    # ---- the volumetric path ---------------------------------------------------------------
    def bev_encoder_vol(self, bev: torch.Tensor) -> Vol:
        with ops.math_scope("voxel.encoder"):
            levels = self.img_bev_encoder_backbone.forward_vol(Vol(bev))
        with ops.math_scope("voxel.neck"):
            neck = self.img_bev_encoder_neck.forward_vol(levels)
        st = self.img_view_transformer.stage_outputs
        if st is not None:          # per-stage capture for the parity tests (logical NCDHW views, neck materialised)
            st.update(bev_feat=bev.permute(0, 4, 1, 2, 3), neck=neck.ncdhw(),
                      **{f"enc{i}": t.permute(0, 4, 1, 2, 3) for i, t in enumerate(levels)})
        return neck

    def forward_features(self, x_left, x_right, left, right, calib, occ_size=None, want_labels=False):
        """Volumetric forward from image-backbone features.
        x_left/x_right: [B,1,Cin,fH,fW]; left/right: calibration dicts (rots, trans, intrins,
        post_rots, post_trans, bda); calib: [B,1].  Returns dict(output_voxels = logical
        [B,classes,*occ_size], depth = depth_prob, labels = uint8 argmax or None)."""
        vt = self.img_view_transformer
        keys = ("rots", "trans", "intrins", "post_rots", "post_trans", "bda")
        # calibration-only vectors: cached per calibration like the splat index
        ml = ops.cached_const("mlp_input", [left[k] for k in keys], lambda: vt.get_mlp_input(*[left[k] for k in keys]).contiguous())
        mr = ops.cached_const("mlp_input", [right[k] for k in keys], lambda: vt.get_mlp_input(*[right[k] for k in keys]).contiguous())
        geo_l = [left[k] for k in keys] + [ml]
        geo_r = [right[k] for k in keys] + [mr]
        bev, depth = vt([x_left] + geo_l + [x_right] + geo_r + [calib, None, None])
        neck = self.bev_encoder_vol(bev.permute(0, 2, 3, 4, 1))
        with ops.math_scope("voxel.head"):
            logits = self.pts_bbox_head.forward_voxel_vol([neck])[0]        # [B,X,Y,Z,classes]
        labels = None
        if occ_size is not None:
            up, labels = ops.trilinear(logits, occ_size, want_labels=want_labels)
        else:
            up = logits
        return {"output_voxels": up.permute(0, 4, 1, 2, 3), "logits_lowres": logits.permute(0, 4, 1, 2, 3),
                "depth": depth, "labels": labels, "output_points": None}

    # ---- reference-signature entry points ----------------------------------------------------
    def extract_img_feat(self, img, img_metas=None):
        """bevdepth_occupancy.py:83-128; needs an attached image encoder."""
        left, right = img[0], img[1]
        B = left[0].shape[0]
        feats = self.image_encoder(torch.cat([left[0], right[0]], 0))
        x, x2 = feats[:B], feats[B:]
        vt = self.img_view_transformer
        ml = vt.get_mlp_input(*left[1:7])
        mr = vt.get_mlp_input(*right[1:7])
        bev, depth = vt([x] + list(left[1:7]) + [ml] + [x2] + list(right[1:7]) + [mr] + [left[-1], left, right])
        neck = self.bev_encoder_vol(bev.permute(0, 2, 3, 4, 1))
        return [neck], depth, x

    def simple_test(self, img_metas, img=None, rescale=False, points_occ=None, gt_occ=None, points_uv=None):
        """bevdepth_occupancy.py:275-297."""
        voxel_feats, depth, img_feats = self.extract_img_feat(img, img_metas)
        with ops.math_scope("voxel.head"):
            logits = self.pts_bbox_head.forward_voxel_vol(voxel_feats)[0]
        up, _ = ops.trilinear(logits, tuple(gt_occ.shape[1:]))
        return {"output_voxels": up.permute(0, 4, 1, 2, 3), "output_points": None, "evaluation_semantic": 0,
                "target_voxels": gt_occ}

    def forward_test(self, img_metas=None, img_inputs=None, **kwargs):
        return self.simple_test(img_metas, img_inputs, **kwargs)

    def forward(self, return_loss=False, **kwargs):
        if return_loss:
            raise NotImplementedError("training (losses, backward) is outside the forward-only hot path")
        return self.forward_test(**kwargs)
