"""``ViewTransformerLiftSplatShootVoxel`` -- B200-native replacement, same registry name, ctor
kwargs, ``forward`` contract and state_dict keys as the reference module
(projects/mmdet3d_plugin/occupancy/image2bev/ViewTransformerLSSVoxel.py:273-526 and its bases in
ViewTransformerLSSBEVDepth.py:74-156, 567-659).

forward(input: list) -> (bev_feat [B,C,X,Y,Z], depth_prob [B*N,D,fH,fW]):
  input[0:8]  = left  (x[B,1,Cin,fH,fW], rots, trans, intrins, post_rots, post_trans, bda, mlp_input)
  input[8:16] = right (same layout), input[16] = calib [B,1] (focal * baseline); trailing items ignored.

Stages and the kernels that run them (all through the C ABI, stereoscene_b200.ops):
  (i)   stereo: reduce conv + pending GN/ReLU/SE gate + 1x1 conv -> ss_gwc_warp_fwd ->
        5 full-res convs + 3 hourglasses (ss_conv3d_fwd with pending affines, ss_affine_join_fwd)
        -> ss_softmax_d_fwd
  (N1)  depth_net: reduce conv + pending GN/ReLU/SE gates, 3 BasicBlocks (pending BatchNorm), ASPP
        (dilated convs into one sliced buffer, pooled branch folded into a per-channel shift), DCN
        (ss_deform_sample_fwd + grouped GEMM), 1x1 head -- all on ss_conv3d_*_fwd
  (iii) MIE: 2 x ss_bri_attn_fwd -> redir1 -> hourglass -> CA3D (gate folded into a pending affine,
        ss_ca3d_gate) -> redir2 -> ss_softmax_d_fwd
  (ii)  lift (x) splat: ss_splat_build_index (calibration-only, cached) + ss_lift_splat_fwd
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..ops import SS_ACT_GELU, SS_ACT_NONE, SS_ACT_RELU, Vol
from ..registry import NECKS
from .layers import HourglassParams, MlpParams, SEParams, conv_gn, conv_gn3d, hourglass


# ------------------------------------------------------------------------------------------
# parameter containers of the stereo branch and the MIE block
# ------------------------------------------------------------------------------------------
_STREAM_OVERLAP = os.environ.get("STEREOSCENE_B200_STREAM_OVERLAP", "1") != "0"      # depth_net on a side stream beside the stereo branch
_side_streams = {}


def _side_stream(device) -> "torch.cuda.Stream":
    key = (device.type, device.index)
    s = _side_streams.get(key)
    if s is None:
        s = _side_streams[key] = torch.cuda.Stream(device)
    return s


class StereoFeatureParams(nn.Module):
    """keys: reduce_conv.{0,1}, depth_mlp, depth_se, depth_conv.0 (ViewTransformerLSSVoxel.py:32-58)."""

    def __init__(self, cin, mid, cout, cam):
        super().__init__()
        self.reduce_conv = nn.Sequential(nn.Conv2d(cin, mid, 3, 1, 1), nn.GroupNorm(2, mid), nn.ReLU())
        self.bn = nn.Identity()
        self.depth_mlp = MlpParams(cam, mid, mid)
        self.depth_se = SEParams(mid)
        self.depth_conv = nn.Sequential(nn.Conv2d(mid, cout, 1, 1, 0))


class StereoVolumeParams(nn.Module):
    """``GwcNet_volume_encoder`` tree (ViewTransformerLSSVoxel.py:158-203)."""

    def __init__(self, maxdisp, out_c=32, cin=640):
        super().__init__()
        self.maxdisp = maxdisp
        self.num_groups = 32
        self.feature_withcam = StereoFeatureParams(cin, 128, 64, 30)
        relu = lambda: nn.ReLU(inplace=True)     # noqa: E731  (index placeholders: keys 0 / 2)
        self.dres0 = nn.Sequential(conv_gn3d(32, 32, 3, 1, 1), relu(), conv_gn3d(32, 32, 3, 1, 1), relu())
        self.dres1 = nn.Sequential(conv_gn3d(32, 32, 3, 1, 1), relu(), conv_gn3d(32, 32, 3, 1, 1))
        self.dres2 = HourglassParams(32)
        self.dres3 = HourglassParams(32)
        self.dres4 = HourglassParams(32)
        self.classif3_1 = nn.Sequential(conv_gn3d(32, out_c, 3, 1, 1), relu())
        self.classif3_2 = nn.Sequential(nn.Conv3d(out_c, 1, 3, padding=1, stride=1, bias=False))
        # the reference re-initialises every Conv2d/Conv3d of this sub-net (VT:189-203)
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Conv3d)):
                n = math.prod(m.kernel_size) * m.out_channels
                nn.init.normal_(m.weight, 0.0, math.sqrt(2.0 / n))
            elif isinstance(m, nn.Linear):
                nn.init.zeros_(m.bias)


class AttentionParams(nn.Module):
    """keys: query_conv, key_conv, value_conv (1x1x1, one channel), gamma (attention.py:45-56)."""

    def __init__(self, in_dim=1):
        super().__init__()
        if in_dim != 1:
            raise NotImplementedError("BRI attention is defined on single-channel depth volumes (in_dim=1)")
        self.query_conv = nn.Conv3d(in_dim, in_dim, 1)
        self.key_conv = nn.Conv3d(in_dim, in_dim, 1)
        self.value_conv = nn.Conv3d(in_dim, in_dim, 1)
        self.gamma = nn.Parameter(torch.zeros(1))

    def packed(self) -> torch.Tensor:
        """device float[7] = wq,bq,wk,bk,wv,bv,gamma for ss_bri_attn_fwd."""
        ps = [self.query_conv.weight, self.query_conv.bias, self.key_conv.weight, self.key_conv.bias,
              self.value_conv.weight, self.value_conv.bias, self.gamma]
        return ops.cached_const("bri_params", ps, lambda: torch.cat([p.detach().view(1) for p in ps]).float().contiguous())


class CA3DParams(nn.Module):
    """keys: conv1.{0,2}, conv2.{0,2}, conv.{0,2} (attention.py:90-112)."""

    def __init__(self, c):
        super().__init__()
        self.conv1 = nn.Sequential(nn.Conv3d(c, c, 3, 1, 1), nn.GELU(), nn.GroupNorm(1, c))
        self.conv2 = nn.Sequential(nn.Conv3d(c, c // 8, 1), nn.GELU(), nn.Conv3d(c // 8, c, 1), nn.GELU())
        self.conv = nn.Sequential(nn.Conv3d(c, c, 3, 1, 1), nn.GELU(), nn.GroupNorm(1, c))


class ResidualParams(nn.Module):
    """keys: alpha, fn.* (ViewTransformerLSSVoxel.py:227-234)."""

    def __init__(self, fn):
        super().__init__()
        self.fn = fn
        self.alpha = nn.Parameter(torch.zeros(1))


class VolumeInteractionParams(nn.Module):
    """MIE block tree (ViewTransformerLSSVoxel.py:236-246)."""

    def __init__(self):
        super().__init__()
        self.redir1 = nn.Conv3d(2, 32, 3, 1, 1)
        self.dres1 = HourglassParams(32)
        self.redir2 = nn.Conv3d(32, 1, 3, 1, 1)
        self.lss2stereo = AttentionParams(1)
        self.stereo2lss = AttentionParams(1)
        self.CA3D = ResidualParams(CA3DParams(32))


# ------------------------------------------------------------------------------------------
# DepthNet (adjacent component, row N1 of SURVEY.md section 8f) on the same kernels as the rest of
# the path: every map is a channels-last [B,1,H,W,C] volume, BatchNorm(eval) / GroupNorm / SE gates
# travel as pending affines, the ASPP concat is a channel-sliced buffer.
# ------------------------------------------------------------------------------------------
def _holder(weight: torch.Tensor) -> nn.Conv3d:
    """A 1x1x1 Conv3d that only carries a derived weight matrix [Cout,Cin] into ops.conv (not a
    parameter of the model: the state_dict stays the reference's)."""
    c = nn.Conv3d(weight.shape[1], weight.shape[0], 1, bias=False).to(weight.device)
    with torch.no_grad():
        c.weight.copy_(weight.reshape(weight.shape[0], weight.shape[1], 1, 1, 1))
    c.weight.requires_grad_(False)
    return c


class BasicBlock2d(nn.Module):
    """mmdet 2.14 BasicBlock (conv3x3-BN-ReLU-conv3x3-BN + identity, ReLU)."""

    def __init__(self, c):
        super().__init__()
        self.conv1 = nn.Conv2d(c, c, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(c)
        self.conv2 = nn.Conv2d(c, c, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(c)

    def forward_vol(self, x: Vol) -> Vol:
        u1, _ = ops.conv(x, self.conv1)
        u2, _ = ops.conv(ops.bn_pending(u1, self.bn1, SS_ACT_RELU), self.conv2)
        return Vol(ops.join(ops.bn_pending(u2, self.bn2), x, out_act=SS_ACT_RELU))


class _ASPPBranch(nn.Module):
    def __init__(self, cin, cout, k, dilation):
        super().__init__()
        self.atrous_conv = nn.Conv2d(cin, cout, k, padding=0 if k == 1 else dilation, dilation=dilation, bias=False)
        self.bn = nn.BatchNorm2d(cout)


class ASPP(nn.Module):
    """ViewTransformerLSSBEVDepth.py:343-414.  The four atrous branches write their raw outputs into
    channel slices of one buffer (their BN+ReLU is the pending affine of the fusing 1x1 conv); the
    global-average-pool branch is constant over the map, so its share of the 1x1 conv is a
    per-(batch,channel) vector folded into the pending shift of bn1."""

    def __init__(self, cin, mid):
        super().__init__()
        self.aspp1 = _ASPPBranch(cin, mid, 1, 1)
        self.aspp2 = _ASPPBranch(cin, mid, 3, 6)
        self.aspp3 = _ASPPBranch(cin, mid, 3, 12)
        self.aspp4 = _ASPPBranch(cin, mid, 3, 18)
        self.global_avg_pool = nn.Sequential(nn.AdaptiveAvgPool2d((1, 1)), nn.Conv2d(cin, mid, 1, bias=False),
                                             nn.GroupNorm(2, mid), nn.ReLU())
        self.conv1 = nn.Conv2d(mid * 5, mid, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(mid)
        self.dropout = nn.Dropout(0.5)
        object.__setattr__(self, "_split", None)
        object.__setattr__(self, "_affine", None)

    def _conv1_split(self):
        """(holder of conv1's columns for the four branches, its columns for the pooled branch)."""
        w = self.conv1.weight
        key = (w.data_ptr(), w._version, w.device)
        if self._split is None or self._split[0] != key:
            mid = w.shape[0]
            w2 = w.detach().flatten(1)
            object.__setattr__(self, "_split", (key, _holder(w2[:, :4 * mid]), w2[:, 4 * mid:].contiguous()))
        return self._split[1], self._split[2]

    def _branch_affine(self, raw: torch.Tensor):
        """BN(eval)+ReLU of the four branches as one [B,4*mid] pending affine (cached)."""
        vs = [ops.bn_pending(raw, br.bn, SS_ACT_RELU) for br in (self.aspp1, self.aspp2, self.aspp3, self.aspp4)]
        key = tuple(id(v.scale) for v in vs)
        if self._affine is None or self._affine[0] != key:
            object.__setattr__(self, "_affine", (key, torch.cat([v.scale for v in vs], 1).contiguous(),
                                                 torch.cat([v.shift for v in vs], 1).contiguous(), vs))
        return self._affine[1], self._affine[2]

    def forward_vol(self, x: Vol) -> Vol:
        B, _, H, W, _ = x.data.shape
        mid = self.conv1.out_channels
        buf = torch.empty((B, 1, H, W, 4 * mid), dtype=torch.float32, device=x.data.device)
        for i, br in enumerate((self.aspp1, self.aspp2, self.aspp3, self.aspp4)):
            ops.conv(x, br.atrous_conv, out=buf[..., i * mid:(i + 1) * mid])
        sc, sh = self._branch_affine(buf)
        # pooled branch on [B,C] vectors: mean -> 1x1 conv -> GroupNorm -> ReLU -> its columns of conv1 (bilinear upsampling of
        # a 1x1 map with align_corners=True is a broadcast), folded into bn1's pending shift by one small kernel
        gp = self.global_avg_pool
        head, w_pool = self._conv1_split()
        y, _ = ops.conv(Vol(buf, sc, sh, SS_ACT_RELU), head)
        bn = ops.bn_pending(y, self.bn1, SS_ACT_RELU)
        shift = ops.aspp_pool_shift(ops.channel_sums(x), H * W, gp[1].weight, gp[2], w_pool, bn.scale, bn.shift)
        return Vol(y, bn.scale, shift, SS_ACT_RELU)


class DCN(nn.Module):
    """mmcv DeformConv2dPack semantics (no bias; offsets from a zero-initialised 3x3 conv),
    ViewTransformerLSSBEVDepth.py:490-498.  Offsets come from the conv kernel in the selected math mode
    (TF32 is also what the reference's own GPU run uses for this conv: cuDNN allows TF32 by default),
    sampling runs in ss_deform_sample_fwd, the grouped GEMM on the tcgen05 conv kernel (one 1x1 conv
    per group over the sampled rows)."""

    def __init__(self, cin, cout, k=3, padding=1, groups=4):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin // groups, k, k))
        nn.init.kaiming_uniform_(self.weight, nonlinearity="relu")
        self.conv_offset = nn.Conv2d(cin, 2 * k * k, k, 1, padding, bias=True)
        nn.init.zeros_(self.conv_offset.weight)
        nn.init.zeros_(self.conv_offset.bias)
        self.padding, self.groups, self.k = padding, groups, k
        object.__setattr__(self, "_gconvs", None)      # per-group 1x1 weight holders (not parameters)
        object.__setattr__(self, "_gkey", None)

    def _group_convs(self):
        w = self.weight
        key = (w.data_ptr(), w._version, w.device)
        if key != self._gkey:
            G, k = self.groups, self.k
            cout_g, cin_g = w.shape[0] // G, w.shape[1]
            # [G][cout_g][cin_g][k*k] -> [G][cout_g][k*k][cin_g]: K index = tap * cin_g + channel
            wm = w.detach().view(G, cout_g, cin_g, k * k).permute(0, 1, 3, 2).reshape(G, cout_g, k * k * cin_g)
            object.__setattr__(self, "_gconvs", [_holder(wm[g]) for g in range(G)])
            object.__setattr__(self, "_gkey", key)
        return self._gconvs

    def forward_vol(self, x: Vol) -> torch.Tensor:
        """x: pending [B,1,H,W,C] volume -> plain [B,1,H,W,Cout]."""
        B, _, H, W, _ = x.data.shape
        off, _ = ops.conv(x, self.conv_offset)                                          # [B,1,H,W,2*k*k]
        off = ops.to_channels_first(off.squeeze(1))                                     # [B,2*k*k,H,W]
        S = ops.deform_sample(x.plain().squeeze(1), off, self.groups, self.k, 1, self.padding, 1)  # [B,H,W,G,k*k,C/G]
        G = self.groups
        S5 = S.view(B, 1, H, W, G, -1)
        cout = self.weight.shape[0]
        out = torch.empty((B, 1, H, W, cout), dtype=torch.float32, device=x.data.device)
        cg = cout // G
        for g, conv in enumerate(self._group_convs()):
            ops.conv(Vol(S5[..., g, :]), conv, out=out[..., g * cg:(g + 1) * cg])
        return out

    def forward(self, x):
        """Reference tensor contract: x [B,C,H,W] -> [B,Cout,H,W] (channels_last memory)."""
        y = self.forward_vol(Vol(ops.to_channels_last(x).unsqueeze(1)))
        return y.squeeze(1).permute(0, 3, 1, 2)


class DepthNet(nn.Module):
    """ViewTransformerLSSBEVDepth.py:457-517.  Output channels: [0:D] depth logits, [D:D+ctx]
    context features."""

    def __init__(self, cin, mid, ctx, depth, cam_channels=27):
        super().__init__()
        self.reduce_conv = nn.Sequential(nn.Conv2d(cin, mid, 3, 1, 1), nn.GroupNorm(2, mid), nn.ReLU(inplace=True))
        self.context_conv = nn.Conv2d(mid, ctx, 1)
        self.bn = nn.GroupNorm(2, cam_channels)
        self.depth_mlp = MlpParams(cam_channels, mid, mid)
        self.depth_se = SEParams(mid)
        self.context_mlp = MlpParams(cam_channels, mid, mid)
        self.context_se = SEParams(mid)
        self.depth_conv = nn.Sequential(BasicBlock2d(mid), BasicBlock2d(mid), BasicBlock2d(mid), ASPP(mid, mid),
                                        DCN(mid, mid, 3, 1, 4), nn.Conv2d(mid, depth, 1))

    def forward_vol(self, x: torch.Tensor, mlp_input: torch.Tensor):
        """x: channels-last [B,1,H,W,Cin] -> (depth logits [B,1,H,W,D], context [B,1,H,W,ctx]), both
        channels-last."""
        # the SE gates depend on the calibration vector and the parameters only: cached per (calibration, checkpoint)
        def gates():
            m = self.bn(mlp_input.reshape(-1, mlp_input.shape[-1]))      # GroupNorm on the [B,cam] calibration vector
            return (self.context_se.gate(self.context_mlp(m)).contiguous(), self.depth_se.gate(self.depth_mlp(m)).contiguous())
        gc, gd = ops.cached_const("depthnet_gates", [mlp_input] + list(self.bn.parameters()) + list(self.context_mlp.parameters()) +
                                  list(self.context_se.parameters()) + list(self.depth_mlp.parameters()) + list(self.depth_se.parameters()), gates)
        with ops.math_scope("depthnet.trunk"):
            y, st = ops.conv(Vol(x), self.reduce_conv[0], want_stats=True)
            # SE gates are > 0: relu(gn(y)) * g == relu(gn(y) * g), so each gate is a rescaled pending affine
            vc, d = ops.gn_pending_gated(y, st, self.reduce_conv[1], SS_ACT_RELU, gc, gd)
            context, _ = ops.conv(vc, self.context_conv)
        with ops.math_scope("depthnet.blocks"):
            for i in range(3):
                d = self.depth_conv[i].forward_vol(d)
        with ops.math_scope("depthnet.aspp"):
            d = self.depth_conv[3].forward_vol(d)
        with ops.math_scope("depthnet.dcn"):
            d = self.depth_conv[4].forward_vol(d)
            depth, _ = ops.conv(Vol(d), self.depth_conv[5])
        return depth, context

    def forward(self, x, mlp_input):
        """Reference tensor contract: x [B,Cin,H,W] -> [B,D+ctx,H,W]."""
        depth, context = self.forward_vol(ops.to_channels_last(x).unsqueeze(1), mlp_input)
        return torch.cat([ops.to_channels_first(depth.squeeze(1)), ops.to_channels_first(context.squeeze(1))], dim=1)


# ------------------------------------------------------------------------------------------
# the view transformer
# ------------------------------------------------------------------------------------------
def gen_dx_bx(xbound, ybound, zbound):
    """ViewTransformerLSSBEVDepth.py:27-31 (fp32 rounding of the stored buffers is part of the
    voxel-index contract)."""
    rows = [xbound, ybound, zbound]
    dx = torch.Tensor([r[2] for r in rows])
    bx = torch.Tensor([r[0] + r[2] / 2.0 for r in rows])
    nx = torch.Tensor([(r[1] - r[0]) / r[2] for r in rows])
    return dx, bx, nx


@NECKS.register_module()
class ViewTransformerLiftSplatShootVoxel(nn.Module):
    def __init__(self, loss_depth_weight, semkitti=False, imgseg=False, imgseg_class=20, lift_with_imgseg=False,
                 point_cloud_range=None, loss_seg_weight=1.0, loss_depth_type="bce", point_xyz_channel=0,
                 point_xyz_mode="cat", cam_channels=27, loss_depth_reg_weight=0.0, use_voxel_net=False,
                 grid_config=None, data_config=None, numC_input=512, numC_Trans=64, downsample=16,
                 accelerate=False, use_bev_pool=True, vp_megvii=False, vp_stero=False, init_cfg=None, **kwargs):
        super().__init__()
        if imgseg or lift_with_imgseg:
            raise NotImplementedError("imgseg auxiliary head is not part of the stereoscene.py path")
        if point_xyz_channel > 0 or point_xyz_mode == "add":
            raise NotImplementedError("point_xyz encoder is dead code for stereoscene.py (point_xyz_channel=0)")
        if use_voxel_net or vp_megvii or accelerate:
            raise NotImplementedError("use_voxel_net / vp_megvii / accelerate are not used by stereoscene.py")
        if grid_config is None:
            grid_config = {"xbound": [-51.2, 51.2, 0.8], "ybound": [-51.2, 51.2, 0.8], "zbound": [-10.0, 10.0, 20.0],
                           "dbound": [1.0, 60.0, 1.0]}
        self.grid_config = grid_config
        dx, bx, nx = gen_dx_bx(grid_config["xbound"], grid_config["ybound"], grid_config["zbound"])
        self.dx = nn.Parameter(dx, requires_grad=False)
        self.bx = nn.Parameter(bx, requires_grad=False)
        self.nx = nn.Parameter(nx, requires_grad=False)
        self.data_config = data_config if data_config is not None else {"input_size": (256, 704)}
        self.downsample = downsample
        self.frustum = self.create_frustum()
        self.D = self.frustum.shape[0]
        self.numC_input, self.numC_Trans = numC_input, numC_Trans
        self.cam_channels = cam_channels
        self.loss_depth_weight = loss_depth_weight
        self.loss_depth_reg_weight = loss_depth_reg_weight
        self.loss_depth_type = loss_depth_type
        self.cam_depth_range = grid_config["dbound"]
        self.semkitti, self.imgseg = semkitti, imgseg
        self.point_xyz_channel, self.point_xyz_mode = point_xyz_channel, point_xyz_mode
        self.depth_net = DepthNet(numC_input, numC_input, numC_Trans, self.D, cam_channels=cam_channels)
        self.stereo_volume_net = StereoVolumeParams(maxdisp=self.D, out_c=32, cin=numC_input)
        self.volume_interaction = VolumeInteractionParams()
        self.forward_dic = {}
        self.cache_splat_index = True
        self._index_cache = None
        self.stage_outputs = None          # set to a dict to capture per-stage tensors (tests)

    # ---- geometry (ViewTransformerLSSBEVDepth.py:107-156, 604-659) --------------------------
    def get_depth_dist(self, x):
        return ops.softmax_d(x)

    def create_frustum(self):
        H, W = self.data_config["input_size"]
        fH, fW = H // self.downsample, W // self.downsample
        ds = torch.arange(*self.grid_config["dbound"], dtype=torch.float).view(-1, 1, 1).expand(-1, fH, fW)
        D = ds.shape[0]
        xs = torch.linspace(0, W - 1, fW, dtype=torch.float).view(1, 1, fW).expand(D, fH, fW)
        ys = torch.linspace(0, H - 1, fH, dtype=torch.float).view(1, fH, 1).expand(D, fH, fW)
        return nn.Parameter(torch.stack((xs, ys, ds), -1), requires_grad=False)

    def get_geometry(self, rots, trans, intrins, post_rots, post_trans, bda):
        """Frustum points in the ego frame, [B,N,D,fH,fW,3] (reference-signature entry, evaluated on the device the
        arguments live on; the product forward evaluates it on the host, see ``splat_index``)."""
        return self._geometry(self.frustum, rots, trans, intrins, post_rots, post_trans, bda)

    @staticmethod
    def _geometry(frustum, rots, trans, intrins, post_rots, post_trans, bda):
        """Small (3 floats per frustum point) and calibration-only: kept in PyTorch ops in the reference's order so
        that it is the reference's arithmetic (VTB:123-156)."""
        B, N, _ = trans.shape
        pts = frustum - post_trans.view(B, N, 1, 1, 1, 3)
        pts = torch.inverse(post_rots).view(B, N, 1, 1, 1, 3, 3).matmul(pts.unsqueeze(-1))
        pts = torch.cat((pts[..., :2, :] * pts[..., 2:3, :], pts[..., 2:3, :]), 5)
        if intrins.shape[3] == 4:          # KITTI: 4x4 with the projection shift in column 3
            pts = pts - intrins[:, :, :3, 3].view(B, N, 1, 1, 1, 3, 1)
            intrins = intrins[:, :, :3, :3]
        comb = rots.matmul(torch.inverse(intrins))
        pts = comb.view(B, N, 1, 1, 1, 3, 3).matmul(pts).squeeze(-1)
        pts = pts + trans.view(B, N, 1, 1, 1, 3)
        if bda.shape[-1] == 4:
            pts = torch.cat((pts, torch.ones_like(pts[..., :1])), dim=-1)
            pts = bda.view(B, 1, 1, 1, 1, 4, 4).matmul(pts.unsqueeze(-1)).squeeze(-1)[..., :3]
        else:
            pts = bda.view(B, 1, 1, 1, 1, 3, 3).matmul(pts.unsqueeze(-1)).squeeze(-1)
        return pts

    def get_mlp_input(self, rot, tran, intrin, post_rot, post_tran, bda=None):
        B, N = rot.shape[:2]
        if bda is None:
            bda = torch.eye(3).to(rot).view(1, 3, 3).repeat(B, 1, 1)
        bda = bda.view(B, 1, *bda.shape[-2:]).repeat(1, N, 1, 1)
        kitti = intrin.shape[-1] == 4
        if kitti:
            items = [intrin[:, :, 0, 0], intrin[:, :, 1, 1], intrin[:, :, 0, 2], intrin[:, :, 1, 2],
                     intrin[:, :, 0, 3], intrin[:, :, 1, 3], intrin[:, :, 2, 3]]
        else:
            items = [intrin[:, :, 0, 0], intrin[:, :, 1, 1], intrin[:, :, 0, 2], intrin[:, :, 1, 2]]
        items += [post_rot[:, :, 0, 0], post_rot[:, :, 0, 1], post_tran[:, :, 0], post_rot[:, :, 1, 0],
                  post_rot[:, :, 1, 1], post_tran[:, :, 1], bda[:, :, 0, 0], bda[:, :, 0, 1], bda[:, :, 1, 0],
                  bda[:, :, 1, 1], bda[:, :, 2, 2]]
        mlp_input = torch.stack(items, dim=-1)
        if kitti and bda.shape[-1] == 4:      # the reference appends the bda translation only in its KITTI branch (VTB:612-636)
            mlp_input = torch.cat((mlp_input, bda[:, :, :3, -1]), dim=2)
        sensor2ego = torch.cat([rot, tran.reshape(B, N, 3, 1)], dim=-1).reshape(B, N, -1)
        return torch.cat([mlp_input, sensor2ego], dim=-1)

    # ---- depth supervision (training helper; ViewTransformerLSSVoxel.py:349-416) -------------
    def get_downsampled_gt_depth(self, gt_depths):
        B, N, H, W = gt_depths.shape
        ds = self.downsample
        g = gt_depths.view(B * N, H // ds, ds, W // ds, ds, 1).permute(0, 1, 3, 5, 2, 4).contiguous().view(-1, ds * ds)
        g = torch.where(g == 0.0, 1e5 * torch.ones_like(g), g).min(dim=-1).values.view(B * N, H // ds, W // ds)
        lo, _, step = self.grid_config["dbound"]
        g = (g - (lo - step / 2)) / step
        vals = g.clone()
        g = torch.where((g < self.D + 1) & (g >= 0.0), g, torch.zeros_like(g))
        return vals, F.one_hot(g.long(), num_classes=self.D + 1).view(-1, self.D + 1)[:, 1:].float()

    def get_depth_loss(self, depth_labels, depth_preds):
        if self.loss_depth_type != "bce":
            raise NotImplementedError("only the 'bce' depth loss is used by stereoscene.py")
        _, labels = self.get_downsampled_gt_depth(depth_labels)
        preds = depth_preds.permute(0, 2, 3, 1).contiguous().view(-1, self.D)
        fg = labels.max(dim=1).values > 0.0
        loss = F.binary_cross_entropy(preds[fg].float(), labels[fg], reduction="none").sum() / max(1.0, float(fg.sum()))
        return self.loss_depth_weight * loss

    # ---- (i) stereo branch ------------------------------------------------------------------
    def stereo_features(self, feat_left, feat_right, mlp_left, mlp_right, pair_cl=None) -> torch.Tensor:
        """stereofeature_net on the batched pair -> channels-last [2B,1,fH,fW,64].  ``pair_cl`` is the
        channels-last [2B,1,fH,fW,Cin] copy of cat(left, right) if the caller already made it."""
        net = self.stereo_volume_net.feature_withcam
        x = pair_cl if pair_cl is not None else self.pair_channels_last(feat_left, feat_right)      # [2B,1,H,W,Cin]
        # SE gate > 0, so relu(gn(y)) * g == relu(gn(y) * g): fold it into the pending affine; it depends on the calibration
        # and the parameters only, so it is cached per (calibration, checkpoint)
        gate = ops.cached_const("stereo_gate", [mlp_left, mlp_right] + list(net.depth_mlp.parameters()) + list(net.depth_se.parameters()),
                                lambda: net.depth_se.gate(net.depth_mlp(torch.cat([mlp_left, mlp_right], 0).reshape(-1, mlp_left.shape[-1]))).contiguous())
        y, st = ops.conv(Vol(x), net.reduce_conv[0], want_stats=True)
        v = ops.gn_pending_gated(y, st, net.reduce_conv[1], SS_ACT_RELU, gate)
        fea, _ = ops.conv(v, net.depth_conv[0])
        return fea

    @staticmethod
    def pair_channels_last(feat_left, feat_right) -> torch.Tensor:
        """[B,Cin,H,W] x 2 -> channels-last [2B,1,H,W,Cin] (left then right) without the intermediate torch.cat copy."""
        B, Cin, H, W = feat_left.shape
        pair = torch.empty((2 * B, 1, H, W, Cin), dtype=torch.float32, device=feat_left.device)
        ops.to_channels_last(feat_left, out=pair[:B])
        ops.to_channels_last(feat_right, out=pair[B:])
        return pair

    def cost_aggregation(self, volume: torch.Tensor) -> torch.Tensor:
        """ViewTransformerLSSVoxel.py:214-222 on a channels-last cost volume -> stereo depth
        distribution [B,D,fH,fW]."""
        net = self.stereo_volume_net
        c = conv_gn(Vol(volume), net.dres0[0], SS_ACT_RELU)
        c = conv_gn(c, net.dres0[2], SS_ACT_RELU)
        r = conv_gn(c, net.dres1[0], SS_ACT_RELU)
        r = conv_gn(r, net.dres1[2], SS_ACT_NONE)
        cost0 = ops.join(r, c)
        o = hourglass(net.dres2, Vol(cost0))
        o = hourglass(net.dres3, Vol(o))
        o = hourglass(net.dres4, Vol(o))
        c31 = conv_gn(Vol(o), net.classif3_1[0], SS_ACT_RELU)
        c3, _ = ops.conv(c31, net.classif3_2[0])                 # [B,D,H,W,1]
        return ops.softmax_d(c3.view(c3.shape[:4]))

    def stereo_volume(self, feat_left, feat_right, mlp_left, mlp_right, calib, pair_cl=None) -> torch.Tensor:
        fea = self.stereo_features(feat_left, feat_right, mlp_left, mlp_right, pair_cl)
        if self.stage_outputs is not None:
            self.stage_outputs["stereo_fea"] = fea
        vol = ops.gwc_warp(fea, calib, self.stereo_volume_net.maxdisp, self.stereo_volume_net.num_groups)
        if self.stage_outputs is not None:
            self.stage_outputs["gwc_warp"] = vol
        return self.cost_aggregation(vol)

    # ---- (iii) MIE ---------------------------------------------------------------------------
    def mutual_interactive_ensemble(self, stereo: torch.Tensor, lss: torch.Tensor) -> torch.Tensor:
        """volume_interaction.forward (ViewTransformerLSSVoxel.py:248-268); stereo, lss: [B,D,H,W]."""
        vi = self.volume_interaction
        B, D, H, W = stereo.shape
        both = torch.empty((B, D, H, W, 2), dtype=torch.float32, device=stereo.device)
        ops.bri_attention(stereo, lss, vi.lss2stereo.packed(), both[..., 0], 2)      # q = stereo, kv = lss
        ops.bri_attention(lss, stereo, vi.stereo2lss.packed(), both[..., 1], 2)      # q = lss, kv = stereo
        with ops.math_scope("mie.redir1"):
            x, _ = ops.conv(Vol(both), vi.redir1, out_act=SS_ACT_RELU)
        with ops.math_scope("mie.hourglass"):
            x = hourglass(vi.dres1, Vol(x))
        fn = vi.CA3D.fn
        with ops.math_scope("mie.ca3d"):
            d, st = ops.conv(Vol(x), fn.conv1[0], out_act=SS_ACT_GELU, want_stats=True)
            dv = ops.ca3d_gate(ops.gn_pending(d, st, fn.conv1[2]), st, fn.conv2[0], fn.conv2[2])
            o, st2 = ops.conv(dv, fn.conv[0], out_act=SS_ACT_GELU, want_stats=True)
            x2 = ops.join(ops.gn_pending(o, st2, fn.conv[2]), Vol(x), alpha=vi.CA3D.alpha)
        with ops.math_scope("mie.redir2"):
            z, _ = ops.conv(Vol(x2), vi.redir2, out_act=SS_ACT_RELU)
        if self.stage_outputs is not None:
            self.stage_outputs.update(bri=both, mie_hourglass=x, mie_ca3d=x2)
        return ops.softmax_d(z.view(B, D, H, W))

    # ---- (ii) lift + splat -------------------------------------------------------------------
    def splat_index(self, rots, trans, intrins, post_rots, post_trans, bda) -> ops.SplatIndex:
        cal = (rots, trans, intrins, post_rots, post_trans, bda)
        key = tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in cal) + (self.frustum.data_ptr(),)
        if self._index_cache is None:
            self._index_cache = {}
        hit = self._index_cache.get(key) if self.cache_splat_index else None
        if hit is not None:
            self._index_cache[key] = self._index_cache.pop(key)          # most recently used last
            return hit[0]
        # Geometry is calibration-only and feeds an INTEGER quantisation, so it is evaluated once per calibration on the
        # host in fp32 -- the reference's own CPU arithmetic (torch.inverse = LAPACK getrf/getri, the same small matmuls),
        # bit-identical to the oracle -- and uploaded; a device evaluation (cuSOLVER / cuBLAS rounding) could move points
        # that sit on a voxel face.  The result is cached below, so the host round trip never recurs in steady state.
        with torch.no_grad():
            frustum = self.frustum.detach().cpu()
            geom = self._geometry(frustum, *[t.detach().cpu() for t in cal]).to(self.frustum.device)
        if self.stage_outputs is not None:
            self.stage_outputs["geom"] = geom
        nx = [int(round(float(v))) for v in self.nx.detach().cpu()]
        idx = ops.splat_build_index(geom, self.dx.detach().cpu().tolist(), self.bx.detach().cpu().tolist(), nx,
                                    want_coords=self.stage_outputs is not None)
        if self.cache_splat_index:
            # holding the calibration tensors keeps their storage from being recycled under the key; a few entries so that
            # sequences (KITTI calibration differs per sequence) / engines alternating on one model do not thrash
            self._index_cache[key] = (idx, cal)
            while len(self._index_cache) > 8:
                self._index_cache.pop(next(iter(self._index_cache)))
        return idx

    def cached_state(self):
        """Tensors of the calibration caches (see ops.cached_state)."""
        keep = []
        for idx, _ in (self._index_cache or {}).values():
            keep += [idx.order, idx.voxel_start, idx.coords]
        return [t for t in keep if t is not None]

    def voxel_pooling(self, geom_feats, x):
        """Reference-signature entry (ViewTransformerLSSVoxel.py:432-476): geom [B,N,D,H,W,3], lifted
        volume x [B,N,D,H,W,C] -> [B,C,X,Y,Z] through the bev_pool-compatible operator."""
        B, N, D, H, W, Cc = x.shape
        nx = [int(round(float(v))) for v in self.nx.detach().cpu()]
        idx = ops.splat_build_index(geom_feats.reshape(B, -1, 3), self.dx.detach().cpu().tolist(),
                                    self.bx.detach().cpu().tolist(), nx, want_coords=True)
        c = idx.coords.long()
        kept = c[:, 3] > 0
        batch_ix = torch.arange(B, device=x.device).repeat_interleave(N * D * H * W)
        coords = torch.stack((c[:, 0], c[:, 1], c[:, 2], batch_ix), 1)[kept]
        out = ops.bev_pool(x.reshape(-1, Cc)[kept], coords, B, nx[2], nx[0], nx[1])
        return out.permute(0, 1, 3, 4, 2)

    # ---- forward -------------------------------------------------------------------------------
    def frustum_forward(self, input, pair_cl=None):
        """Stages (i), (N1), (iii) -- everything up to the MIE boundary: returns (depth_prob [B,D,fH,fW],
        img_feat [B,fH,fW,C] channels-last, depth_logits, lss, stereo).  ``forward`` = this + lift (x) splat; the X-slab
        sharded mode (stereoscene_b200.xshard) all-gathers the first two tensors here and splats slabs."""
        x, rots, trans, intrins, post_rots, post_trans, bda, mlp_input = input[:8]
        feat_left, mlp_left = input[0].squeeze(1), input[7]
        feat_right, mlp_right = input[8].squeeze(1), input[15]
        calib = input[16]
        ops.arena(x.device).reset()

        B, N, Cin, H, W = x.shape
        if N != 1:
            raise NotImplementedError("the stereo path is defined for one camera per side (N=1)")
        # one channels-last copy of the feature pair serves the stereo branch (both maps) and depth_net (left)
        # (``pair_cl``: the image encoder of this package already produces exactly that buffer, plugin/image_encoder.py)
        if pair_cl is None:
            pair_cl = self.pair_channels_last(feat_left, feat_right)                            # [2B,1,H,W,Cin]
        elif tuple(pair_cl.shape) != (2 * B, 1, H, W, Cin) or not pair_cl.is_contiguous():
            raise RuntimeError(f"pair_cl must be a contiguous [2B,1,H,W,Cin] tensor, got {tuple(pair_cl.shape)}")
        def depth_stage():
            with ops.math_scope("depthnet"):
                d_cl, c_cl = self.depth_net.forward_vol(pair_cl[:B], mlp_input)
            logits = ops.to_channels_first(d_cl.squeeze(1))                       # [B,D,H,W]
            return d_cl, c_cl, logits, ops.softmax_d(logits)

        if _STREAM_OVERLAP and x.is_cuda:
            # The stereo branch and depth_net only meet at the MIE block, and depth_net's 2-D layers (7680 pixels: 60-240 CTAs) leave
            # most of the 148 SMs idle: it runs on a side stream (fork / join by events, capturable into the step's CUDA graph)
            # underneath the stereo branch's full-grid kernels: -0.45 ms per pair (STEREOSCENE_B200_STREAM_OVERLAP=0 serialises
            # them again).  Tensors cross streams only at the fork (pair_cl, allocated before it) and after the join.
            main, side = torch.cuda.current_stream(x.device), _side_stream(x.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                ops.arena(x.device).reset()                                      # the GroupNorm-sum arena is per stream
                depth_cl, ctx_cl, depth_logits, lss = depth_stage()
            with ops.math_scope("stereo"):
                stereo = self.stereo_volume(feat_left, feat_right, mlp_left, mlp_right, calib, pair_cl)
            main.wait_stream(side)
        else:
            with ops.math_scope("stereo"):
                stereo = self.stereo_volume(feat_left, feat_right, mlp_left, mlp_right, calib, pair_cl)
            depth_cl, ctx_cl, depth_logits, lss = depth_stage()
        img_feat = ctx_cl.squeeze(1)                                              # [B,H,W,C] channels-last

        with ops.math_scope("mie"):
            depth_prob = self.mutual_interactive_ensemble(stereo, lss)
        return depth_prob, img_feat, depth_logits, lss, stereo

    # ---- the two independent halves of the frustum stage (the X-slab latency mode runs them on two ranks at once) -----
    def stereo_branch(self, feat_left, feat_right, mlp_left, mlp_right, calib) -> torch.Tensor:
        """(i) alone: stereo depth distribution [B,D,fH,fW] from the feature pair [B,Cin,fH,fW] x 2."""
        with ops.math_scope("stereo"):
            return self.stereo_volume(feat_left, feat_right, mlp_left, mlp_right, calib)

    def depth_branch(self, feat_left, mlp_input):
        """(N1) alone: (lss depth distribution [B,D,fH,fW], context features [B,fH,fW,C] channels-last)."""
        left_cl = ops.to_channels_last(feat_left).unsqueeze(1)
        with ops.math_scope("depthnet"):
            depth_cl, ctx_cl = self.depth_net.forward_vol(left_cl, mlp_input)
        return ops.softmax_d(ops.to_channels_first(depth_cl.squeeze(1))), ctx_cl.squeeze(1).contiguous()

    def mie_branch(self, stereo, lss) -> torch.Tensor:
        with ops.math_scope("mie"):
            return self.mutual_interactive_ensemble(stereo, lss)

    def forward(self, input, pair_cl=None):
        rots, trans, intrins, post_rots, post_trans, bda = input[1:7]
        depth_prob, img_feat, depth_logits, lss, stereo = self.frustum_forward(input, pair_cl)
        index = self.splat_index(rots, trans, intrins, post_rots, post_trans, bda)
        bev = ops.lift_splat(depth_prob, img_feat, index)                         # [B,X,Y,Z,C]
        if self.stage_outputs is not None:
            y = torch.cat([depth_logits, ops.to_channels_first(img_feat)], dim=1)
            self.stage_outputs.update(stereo_prob=stereo, depth_net=y, lss_prob=lss, depth_prob=depth_prob,
                                      splat_index=index)
        return bev.permute(0, 4, 1, 2, 3), depth_prob
