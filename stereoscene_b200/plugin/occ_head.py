"""``OccHead`` -- B200-native replacement of the voxel branch (same registry name, ctor kwargs,
forward contract and state_dict keys as projects/mmdet3d_plugin/occupancy/dense_heads/
occhead.py:28-271).  Forward = Conv3d 384->192 k3 (tensor cores, GroupNorm sums in the epilogue)
-> pending GN(32)+ReLU -> Conv3d 192->20 k1.  The point branch (supervise_points) and the training
losses are outside the accelerated path (stereoscene.py:111 sets supervise_points=False).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..ops import SS_ACT_RELU, Vol
from ..registry import HEADS
from .layers import as_channels_last
from .metrics import SSCMetrics


@HEADS.register_module()
class OccHead(nn.Module):
    def __init__(self, in_channels, out_channel, out_point_channel=None, semantic_kitti=False, supervise_voxel=True,
                 num_level=1, num_img_level=1, in_img_channels=512, sampling_img_feats=False, soft_weights=False,
                 supervise_points=False, loss_weight_cfg=None, semkitti_loss_weight_cfg=None,
                 loss_voxel_prototype="cylinder3d", use_ohem_loss=False, use_sc_ohem_loss=False, ohem_topk=0.25,
                 conv_cfg=dict(type="Conv3d", bias=False), norm_cfg=dict(type="GN", num_groups=32, requires_grad=True),
                 point_cloud_range=(-51.2, -51.2, -5.0, 51.2, 51.2, 3.0), with_cp=False, train_cfg=None, test_cfg=None):
        super().__init__()
        if supervise_points:
            raise NotImplementedError("the point branch is not on the stereoscene.py path (supervise_points=False)")
        if not supervise_voxel:
            raise NotImplementedError("supervise_voxel=False leaves the head without a forward path")
        if norm_cfg.get("type") != "GN" or conv_cfg.get("type") != "Conv3d":
            raise NotImplementedError("OccHead on this path uses Conv3d + GroupNorm")
        if not isinstance(in_channels, (list, tuple)):
            in_channels = [in_channels]
        self.in_channels, self.out_channel, self.num_level = list(in_channels), out_channel, num_level
        self.semantic_kitti = semantic_kitti
        self.supervise_voxel, self.supervise_points = supervise_voxel, supervise_points
        self.semkitti_loss_weight_cfg = semkitti_loss_weight_cfg
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        bias = bool(conv_cfg.get("bias", True))
        self.occ_convs = nn.ModuleList()
        for i in range(num_level):
            mid = self.in_channels[i] // 2
            self.occ_convs.append(nn.Sequential(
                nn.Conv3d(self.in_channels[i], mid, 3, 1, 1, bias=bias),
                nn.GroupNorm(norm_cfg["num_groups"], mid), nn.ReLU(inplace=True),
                nn.Conv3d(mid, out_channel, 1, 1, 0, bias=bias)))
        if semantic_kitti:
            self.ssc_metric = SSCMetrics() if out_channel == 20 else SSCMetrics([str(i) for i in range(out_channel)])

    def forward_voxel_vol(self, voxel_feats):
        """voxel_feats: list[Vol] -> list of plain channels-last logits [B,X,Y,Z,classes]."""
        outs = []
        for v, seq in zip(voxel_feats, self.occ_convs):
            y, st = ops.conv(v, seq[0], want_stats=True)
            h = ops.gn_pending(y, st, seq[1], SS_ACT_RELU)
            logits, _ = ops.conv(h, seq[3])
            outs.append(logits)
        return outs

    def forward(self, voxel_feats, points=None, img_metas=None, img_feats=None, points_uv=None, **kwargs):
        """occhead.py:238-271 (voxel branch): voxel_feats = list of logical [B,C,X,Y,Z] tensors (or Vol)."""
        assert type(voxel_feats) is list and len(voxel_feats) == self.num_level
        if points is not None:
            raise NotImplementedError("query-point outputs are outside the accelerated path")
        dev = voxel_feats[0].data.device
        ops.arena(dev).reset()
        vols = [v if isinstance(v, Vol) else Vol(as_channels_last(v)) for v in voxel_feats]
        logits = self.forward_voxel_vol(vols)
        return {"output_voxels": [t.permute(0, 4, 1, 2, 3) for t in logits], "output_points": None}
