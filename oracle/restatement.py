"""TEST INFRASTRUCTURE ONLY -- CPU restatement (PyTorch fp32, functional) of the reference's
volumetric hot path.  Never imported by the product package ``stereoscene_b200``.

Every function states the reference file:line it follows (paths relative to
/root/reference/projects/mmdet3d_plugin/occupancy/ unless noted):
  VT  = image2bev/ViewTransformerLSSVoxel.py      ATT = image2bev/attention.py
  VTB = image2bev/ViewTransformerLSSBEVDepth.py   R3D = backbones/resnet3d.py
  FPN = necks/second_fpn_3d.py                    OCC = dense_heads/occhead.py
  DET = detectors/bevdepth_occupancy.py

It is *pinned*: ``tests/test_oracle_golden.py`` checks it against fixtures under
``tests/golden/`` that ``oracle/make_golden.py`` produced by running the reference's own,
unmodified module code (loaded through ``oracle/ref_loader.py``) on seeded synthetic inputs;
when /root/reference is present the same test also runs the reference live.

The arithmetic the reference takes from third-party code that is absent from the tree is
restated from the published algorithms and named here:
  * ``mmdet3d.ops.bev_pool`` (BEVFusion fork of mmdet3d, unpinned): per-voxel sum -> bev_pool()
  * mmcv-full 1.4.0 ``DCN`` (DeformConv2dPack): torchvision.ops.deform_conv2d
  * mmdet 2.14.0 ``BasicBlock``: conv3x3-BN-ReLU-conv3x3-BN + identity, ReLU
All functions take the model's flat ``state_dict`` (reference key names) and a key prefix.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

EPS = 1e-5


# ------------------------------------------------------------------------------------------
# small building blocks
# ------------------------------------------------------------------------------------------
def _gn(sd, p, x, groups):
    return F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], EPS)


def _bn_eval(sd, p, x):
    """BatchNorm in eval mode = per-channel affine from the running statistics."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                        sd[p + ".bias"], False, 0.0, EPS)


def _conv3(sd, p, x, stride=1, pad=1):
    return F.conv3d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=pad)


def convgn3d(sd, p, x, stride=1, pad=1, groups=2):
    """``convbn_3d`` = Conv3d(no bias) + GroupNorm(2) despite the name (VT:31, 66-69)."""
    return _gn(sd, p + ".1", _conv3(sd, p + ".0", x, stride, pad), groups)


# ------------------------------------------------------------------------------------------
# (i) stereo branch
# ------------------------------------------------------------------------------------------
def mlp2(sd, p, x):
    """``Mlp`` fc1-ReLU-fc2, dropout inactive in eval (VTB:417-439)."""
    return F.linear(F.relu(F.linear(x, sd[p + ".fc1.weight"], sd[p + ".fc1.bias"])),
                    sd[p + ".fc2.weight"], sd[p + ".fc2.bias"])


def se_gate(sd, p, x, x_se):
    """``SELayer``: x * sigmoid(expand(relu(reduce(x_se)))) (VTB:442-454)."""
    g = F.conv2d(x_se, sd[p + ".conv_reduce.weight"], sd[p + ".conv_reduce.bias"])
    g = F.conv2d(F.relu(g), sd[p + ".conv_expand.weight"], sd[p + ".conv_expand.bias"])
    return x * torch.sigmoid(g)


def stereo_feature(sd, p, x, mlp_input):
    """``stereofeature_net.forward`` (VT:59-65): 3x3 conv(+bias) -> GN(2) -> ReLU, camera-aware
    SE gate driven by the 30-float calibration vector, 1x1 conv to 64 channels.  ``bn`` is
    Identity here (VT:48)."""
    m = mlp_input.reshape(-1, mlp_input.shape[-1])
    y = F.conv2d(x, sd[p + ".reduce_conv.0.weight"], sd[p + ".reduce_conv.0.bias"], padding=1)
    y = F.relu(_gn(sd, p + ".reduce_conv.1", y, 2))
    se = mlp2(sd, p + ".depth_mlp", m)[..., None, None]
    y = se_gate(sd, p + ".depth_se", y, se)
    return F.conv2d(y, sd[p + ".depth_conv.0.weight"], sd[p + ".depth_conv.0.bias"])


def gwc_volume(ref, tgt, maxdisp, groups):
    """``build_gwc_volume`` + ``groupwise_correlation`` (VT:97-114):
    vol[b,g,i,h,w] = mean_c ref[b,g*cpg+c,h,w] * tgt[b,g*cpg+c,h,w-i] for w >= i, else 0."""
    B, C, H, W = ref.shape
    cpg = C // groups
    vol = ref.new_zeros(B, groups, maxdisp, H, W)
    for i in range(min(maxdisp, W)):
        prod = ref[:, :, :, i:] * tgt[:, :, :, :W - i]
        vol[:, :, i, :, i:] = prod.view(B, groups, cpg, H, W - i).mean(dim=2)
    return vol


def disparity_sample_positions(calib, n_bins, down=1):
    """Sampling position along the disparity axis for every depth bin (VT:139-150 followed by
    grid_sample's align_corners=True un-normalisation): bin k is treated as depth (k+1) and
    sampled at disparity calib/(4*down*(k+1)) in feature pixels; the value makes the same fp32
    round trip through the [-1,1] normalised coordinate that the reference takes."""
    B = calib.shape[0]
    D = n_bins
    k = torch.arange(1, 1 + n_bins // down, dtype=torch.float32, device=calib.device)
    xx = (calib.reshape(B, -1)[:, :1].float() / (down * 4.0)) / k[None, :]          # [B, D]
    xn = 2.0 * xx / max(D - 1, 1) - 1.0
    return ((xn + 1.0) / 2.0) * (D - 1)


def warp_disparity_to_depth(vol, calib):
    """``warp`` (VT:128-156): resample the disparity axis to depth bins with 1-D linear
    interpolation (grid_sample bilinear whose second coordinate is an exact integer),
    zero padding outside [0, D-1]."""
    B, G, D, H, W = vol.shape
    pos = disparity_sample_positions(calib, D)                       # [B, D]
    i0 = torch.floor(pos)
    w1 = pos - i0
    w0 = (i0 + 1.0) - pos
    i0 = i0.long()
    i1 = i0 + 1
    out = vol.new_zeros(B, G, D, H, W)
    for b in range(B):
        for k in range(D):
            a, c = int(i0[b, k]), int(i1[b, k])
            acc = None
            if 0 <= a <= D - 1:
                acc = vol[b, :, a] * w0[b, k]
            if 0 <= c <= D - 1:
                t = vol[b, :, c] * w1[b, k]
                acc = t if acc is None else acc + t
            if acc is not None:
                out[b, :, k] = acc
    return out


def hourglass(sd, p, x):
    """``hourglass.forward`` (VT:89-96; layers VT:73-88).  conv5/conv6 are ConvTranspose3d
    k3 s2 p1 output_padding 1 followed by BatchNorm3d (eval: running statistics)."""
    c1 = F.relu(convgn3d(sd, p + ".conv1.0", x, 2, 1))
    c2 = F.relu(convgn3d(sd, p + ".conv2.0", c1, 1, 1))
    c3 = F.relu(convgn3d(sd, p + ".conv3.0", c2, 2, 1))
    c4 = F.relu(convgn3d(sd, p + ".conv4.0", c3, 1, 1))
    u5 = F.conv_transpose3d(c4, sd[p + ".conv5.0.weight"], None, stride=2, padding=1, output_padding=1)
    c5 = F.relu(_bn_eval(sd, p + ".conv5.1", u5) + convgn3d(sd, p + ".redir2", c2, 1, 0))
    u6 = F.conv_transpose3d(c5, sd[p + ".conv6.0.weight"], None, stride=2, padding=1, output_padding=1)
    return F.relu(_bn_eval(sd, p + ".conv6.1", u6) + convgn3d(sd, p + ".redir1", x, 1, 0))


def cost_aggregation(sd, p, volume):
    """Body of ``GwcNet_volume_encoder.forward`` after the warp (VT:214-222).  Returns
    (cost3_1 [B,32,D,H,W], pred3 [B,D,H,W])."""
    c = F.relu(convgn3d(sd, p + ".dres0.0", volume))
    c = F.relu(convgn3d(sd, p + ".dres0.2", c))
    r = F.relu(convgn3d(sd, p + ".dres1.0", c))
    c = convgn3d(sd, p + ".dres1.2", r) + c
    o = hourglass(sd, p + ".dres2", c)
    o = hourglass(sd, p + ".dres3", o)
    o = hourglass(sd, p + ".dres4", o)
    c31 = F.relu(convgn3d(sd, p + ".classif3_1.0", o))
    c3 = F.conv3d(c31, sd[p + ".classif3_2.0.weight"], None, padding=1).squeeze(1)
    return c31, F.softmax(c3, dim=1)


def gwc_warp(sd, p, feat_left, feat_right, mlp_left, mlp_right, calib, maxdisp, groups=32):
    """Front half of ``GwcNet_volume_encoder.forward`` (VT:205-213)."""
    B = feat_left.shape[0]
    fea = stereo_feature(sd, p + ".feature_withcam", torch.cat([feat_left, feat_right], 0),
                         torch.cat([mlp_left, mlp_right], 0))
    vol = gwc_volume(fea[:B], fea[B:], maxdisp, groups)
    return warp_disparity_to_depth(vol, calib), fea


def stereo_volume_net(sd, p, feat_left, feat_right, mlp_left, mlp_right, calib, maxdisp):
    """``GwcNet_volume_encoder.forward`` (VT:204-224) -> "single_channel" [B,D,H,W]."""
    vol, _ = gwc_warp(sd, p, feat_left, feat_right, mlp_left, mlp_right, calib, maxdisp)
    return cost_aggregation(sd, p, vol)[1]


# ------------------------------------------------------------------------------------------
# (iii) Mutual Interactive Ensemble
# ------------------------------------------------------------------------------------------
def bri_attention(sd, p, q, kv):
    """``attention.forward`` (ATT:58-86).  q, kv: [B,1,D,H,W].  Tokens are the H*W pixels, the
    feature axis is depth.  conf[j] = max_d softmax_d(q)[d,j] scales *key column* j of the
    attention matrix (ATT:63-65, 76 -- broadcast over the last axis); no 1/sqrt(d) scaling."""
    B, C, D, H, W = kv.shape
    N = H * W
    conf = F.softmax(q, dim=2).max(dim=2)[0].reshape(B, -1, N)                 # [B,1,N]
    wq, bq = sd[p + ".query_conv.weight"].reshape(()), sd[p + ".query_conv.bias"].reshape(())
    wk, bk = sd[p + ".key_conv.weight"].reshape(()), sd[p + ".key_conv.bias"].reshape(())
    wv, bv = sd[p + ".value_conv.weight"].reshape(()), sd[p + ".value_conv.bias"].reshape(())
    Q = (q * wq + bq).reshape(B, -1, N).transpose(1, 2)                       # [B,N,D]
    K = (kv * wk + bk).reshape(B, -1, N)                                      # [B,D,N]
    A = F.softmax(torch.bmm(Q, K), dim=-1) * conf                             # [B,N,N]
    V = (kv * wv + bv).reshape(B, -1, N)
    out = torch.bmm(V, A.transpose(1, 2)).reshape(B, C, D, H, W)
    return sd[p + ".gamma"] * out + kv


def ca3d(sd, p, x):
    """``CA3D.forward`` (ATT:113-120; layers ATT:93-112): GELU comes *before* GroupNorm(1),
    and the squeeze branch ends GELU -> sigmoid."""
    d = F.gelu(F.conv3d(x, sd[p + ".conv1.0.weight"], sd[p + ".conv1.0.bias"], padding=1))
    d = _gn(sd, p + ".conv1.2", d, 1)
    s = d.mean(dim=(2, 3, 4), keepdim=True)
    s = F.gelu(F.conv3d(s, sd[p + ".conv2.0.weight"], sd[p + ".conv2.0.bias"]))
    s = F.gelu(F.conv3d(s, sd[p + ".conv2.2.weight"], sd[p + ".conv2.2.bias"]))
    o = torch.sigmoid(s) * d
    o = F.gelu(F.conv3d(o, sd[p + ".conv.0.weight"], sd[p + ".conv.0.bias"], padding=1))
    return _gn(sd, p + ".conv.2", o, 1)


def volume_interaction(sd, p, stereo, lss):
    """``volume_interaction.forward`` (VT:248-268): BRI both ways, then DVE
    (redir1 -> hourglass -> alpha*CA3D(x)+x (VT:227-234) -> redir2 -> ReLU -> softmax over D)."""
    s, l = stereo.unsqueeze(1), lss.unsqueeze(1)
    a = bri_attention(sd, p + ".lss2stereo", s, l)
    b = bri_attention(sd, p + ".stereo2lss", l, s)
    x = F.relu(F.conv3d(torch.cat((a, b), 1), sd[p + ".redir1.weight"], sd[p + ".redir1.bias"], padding=1))
    x = hourglass(sd, p + ".dres1", x)
    x = sd[p + ".CA3D.alpha"] * ca3d(sd, p + ".CA3D.fn", x) + x
    x = F.relu(F.conv3d(x, sd[p + ".redir2.weight"], sd[p + ".redir2.bias"], padding=1)).squeeze(1)
    return F.softmax(x, dim=1)


# ------------------------------------------------------------------------------------------
# adjacent: DepthNet (row N1)
# ------------------------------------------------------------------------------------------
def _basic_block2d(sd, p, x):
    """mmdet 2.14 BasicBlock (third-party; restated from the published ResNet basic block)."""
    y = F.relu(_bn_eval(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"], None, padding=1)))
    y = _bn_eval(sd, p + ".bn2", F.conv2d(y, sd[p + ".conv2.weight"], None, padding=1))
    return F.relu(y + x)


def _aspp(sd, p, x):
    """``ASPP.forward`` (VTB:390-408; branches VTB:348-388, 312-333); dropout inactive in eval."""
    outs = []
    for name, dil in (("aspp1", 1), ("aspp2", 6), ("aspp3", 12), ("aspp4", 18)):
        w = sd[f"{p}.{name}.atrous_conv.weight"]
        pad = 0 if w.shape[-1] == 1 else dil
        y = F.conv2d(x, w, None, padding=pad, dilation=dil)
        outs.append(F.relu(_bn_eval(sd, f"{p}.{name}.bn", y)))
    g = x.mean(dim=(2, 3), keepdim=True)
    g = F.conv2d(g, sd[p + ".global_avg_pool.1.weight"], None)
    g = F.relu(_gn(sd, p + ".global_avg_pool.2", g, 2))
    outs.append(g.expand(-1, -1, x.shape[2], x.shape[3]))      # bilinear upsample of a 1x1 map
    y = F.conv2d(torch.cat(outs, 1), sd[p + ".conv1.weight"], None)
    return F.relu(_bn_eval(sd, p + ".bn1", y))


def depth_net(sd, p, x, mlp_input):
    """``DepthNet.forward`` (VTB:506-517; ctor VTB:457-504).  ``bn`` = GroupNorm(2, 30) applied
    to the [B*N, 30] calibration vector (VTB:479).  depth_conv = 3 BasicBlocks, ASPP, DCN
    (groups 4, deform_groups 1, no bias), 1x1 conv."""
    from torchvision.ops import deform_conv2d
    m = mlp_input.reshape(-1, mlp_input.shape[-1])
    m = F.group_norm(m, 2, sd[p + ".bn.weight"], sd[p + ".bn.bias"], EPS)
    y = F.conv2d(x, sd[p + ".reduce_conv.0.weight"], sd[p + ".reduce_conv.0.bias"], padding=1)
    y = F.relu(_gn(sd, p + ".reduce_conv.1", y, 2))
    ctx = se_gate(sd, p + ".context_se", y, mlp2(sd, p + ".context_mlp", m)[..., None, None])
    ctx = F.conv2d(ctx, sd[p + ".context_conv.weight"], sd[p + ".context_conv.bias"])
    d = se_gate(sd, p + ".depth_se", y, mlp2(sd, p + ".depth_mlp", m)[..., None, None])
    for i in range(3):
        d = _basic_block2d(sd, f"{p}.depth_conv.{i}", d)
    d = _aspp(sd, p + ".depth_conv.3", d)
    off = F.conv2d(d, sd[p + ".depth_conv.4.conv_offset.weight"], sd[p + ".depth_conv.4.conv_offset.bias"], padding=1)
    d = deform_conv2d(d, off, sd[p + ".depth_conv.4.weight"], None, 1, 1, 1)
    d = F.conv2d(d, sd[p + ".depth_conv.5.weight"], sd[p + ".depth_conv.5.bias"])
    return torch.cat([d, ctx], dim=1)


# ------------------------------------------------------------------------------------------
# (ii) lift + splat
# ------------------------------------------------------------------------------------------
def gen_dx_bx(xbound, ybound, zbound):
    """VTB:27-31 (values are stored as fp32 Parameters; keep the same rounding)."""
    rows = [xbound, ybound, zbound]
    dx = torch.Tensor([r[2] for r in rows])
    bx = torch.Tensor([r[0] + r[2] / 2.0 for r in rows])
    nx = torch.Tensor([(r[1] - r[0]) / r[2] for r in rows])
    return dx, bx, nx


def create_frustum(input_size, downsample, dbound):
    """VTB:110-121 -> [D, fH, fW, 3] (u, v, depth)."""
    H, W = input_size
    fH, fW = H // downsample, W // downsample
    ds = torch.arange(*dbound, dtype=torch.float).view(-1, 1, 1).expand(-1, fH, fW)
    D = ds.shape[0]
    xs = torch.linspace(0, W - 1, fW, dtype=torch.float).view(1, 1, fW).expand(D, fH, fW)
    ys = torch.linspace(0, H - 1, fH, dtype=torch.float).view(1, fH, 1).expand(D, fH, fW)
    return torch.stack((xs, ys, ds), -1)


def get_geometry(frustum, rots, trans, intrins, post_rots, post_trans, bda):
    """VTB:123-156, KITTI branch (4x4 intrinsics carry the P[:,3] shift) and 3x3 ``bda``."""
    B, N, _ = trans.shape
    pts = frustum - post_trans.view(B, N, 1, 1, 1, 3)
    pts = torch.inverse(post_rots).view(B, N, 1, 1, 1, 3, 3).matmul(pts.unsqueeze(-1))
    pts = torch.cat((pts[..., :2, :] * pts[..., 2:3, :], pts[..., 2:3, :]), 5)
    if intrins.shape[3] == 4:
        pts = pts - intrins[:, :, :3, 3].view(B, N, 1, 1, 1, 3, 1)
        intrins = intrins[:, :, :3, :3]
    comb = rots.matmul(torch.inverse(intrins))
    pts = comb.view(B, N, 1, 1, 1, 3, 3).matmul(pts).squeeze(-1)
    pts = pts + trans.view(B, N, 1, 1, 1, 3)
    return bda.view(B, 1, 1, 1, 1, 3, 3).matmul(pts.unsqueeze(-1)).squeeze(-1)


def voxel_indices(geom, dx, bx, nx):
    """VT:441-451: idx = trunc_toward_zero((geom - (bx - dx/2)) / dx) as int64, keep-test on the
    integers.  Returns (coords [N',3] int64, kept [N'] bool) for the flattened points."""
    idx = ((geom - (bx - dx / 2.0)) / dx).long().view(-1, 3)
    kept = ((idx[:, 0] >= 0) & (idx[:, 0] < nx[0]) & (idx[:, 1] >= 0) & (idx[:, 1] < nx[1])
            & (idx[:, 2] >= 0) & (idx[:, 2] < nx[2]))
    return idx, kept


def bev_pool(feats, coords, B, D, H, W):
    """External op ``mmdet3d.ops.bev_pool`` (call VT:473): coords = (x, y, z, b); out[b,c,z,x,y]
    = sum of the feats rows in that voxel, accumulated in point order."""
    B, D, H, W = int(B), int(D), int(H), int(W)
    flat = ((coords[:, 3] * D + coords[:, 2]) * H + coords[:, 0]) * W + coords[:, 1]
    out = feats.new_zeros(B * D * H * W, feats.shape[1]).index_add_(0, flat, feats)
    return out.view(B, D, H, W, -1).permute(0, 4, 1, 2, 3).contiguous()


def lift_splat(depth_prob, img_feat, geom, dx, bx, nx):
    """Lift (VT:517-519) + ``voxel_pooling`` (VT:432-476).  depth_prob [B,D,H,W], img_feat
    [B,C,H,W], geom [B,1,D,H,W,3] -> [B,C,X,Y,Z]."""
    B, D, H, W = depth_prob.shape
    C = img_feat.shape[1]
    vol = (depth_prob.unsqueeze(1) * img_feat.unsqueeze(2)).permute(0, 2, 3, 4, 1).reshape(-1, C)
    idx, kept = voxel_indices(geom, dx, bx, nx)
    batch_ix = torch.arange(B, dtype=torch.long).repeat_interleave(D * H * W).view(-1, 1)
    coords = torch.cat((idx, batch_ix), 1)[kept]
    out = bev_pool(vol[kept], coords, B, nx[2], nx[0], nx[1])
    return out.permute(0, 1, 3, 4, 2)


def get_mlp_input(rot, tran, intrin, post_rot, post_tran, bda):
    """``get_mlp_input`` (VTB:604-659), KITTI branch: 18 calibration scalars + the 3x4
    sensor->ego matrix = 30 floats."""
    B, N = rot.shape[:2]
    bda = bda.view(B, 1, 3, 3).repeat(1, N, 1, 1)
    items = [intrin[:, :, 0, 0], intrin[:, :, 1, 1], intrin[:, :, 0, 2], intrin[:, :, 1, 2],
             intrin[:, :, 0, 3], intrin[:, :, 1, 3], intrin[:, :, 2, 3],
             post_rot[:, :, 0, 0], post_rot[:, :, 0, 1], post_tran[:, :, 0],
             post_rot[:, :, 1, 0], post_rot[:, :, 1, 1], post_tran[:, :, 1],
             bda[:, :, 0, 0], bda[:, :, 0, 1], bda[:, :, 1, 0], bda[:, :, 1, 1], bda[:, :, 2, 2]]
    s2e = torch.cat([rot, tran.reshape(B, N, 3, 1)], dim=-1).reshape(B, N, -1)
    return torch.cat([torch.stack(items, dim=-1), s2e], dim=-1)


def view_transformer(sd, p, xl, xr, left, right, calib, grid_config, input_size, downsample=8,
                     numC_Trans=128, stages=None):
    """``ViewTransformerLiftSplatShootVoxel.forward`` (VT:478-526).  ``left``/``right`` are the
    calibration dicts of synth.kitti_calibration.  Returns (bev_feat [B,C,X,Y,Z], depth_prob);
    if ``stages`` is a dict the per-stage tensors are stored in it."""
    frustum = sd[p + ".frustum"] if (p + ".frustum") in sd else create_frustum(input_size, downsample, grid_config["dbound"])
    D = frustum.shape[0]
    dx, bx, nx = gen_dx_bx(grid_config["xbound"], grid_config["ybound"], grid_config["zbound"])
    ml = get_mlp_input(left["rots"], left["trans"], left["intrins"], left["post_rots"], left["post_trans"], left["bda"])
    mr = get_mlp_input(right["rots"], right["trans"], right["intrins"], right["post_rots"], right["post_trans"], right["bda"])
    fl, fr = xl.squeeze(1), xr.squeeze(1)
    vol, fea = gwc_warp(sd, p + ".stereo_volume_net", fl, fr, ml, mr, calib, D)
    stereo = cost_aggregation(sd, p + ".stereo_volume_net", vol)[1]
    B, N, C, H, W = xl.shape
    y = depth_net(sd, p + ".depth_net", xl.view(B * N, C, H, W), ml)
    lss = F.softmax(y[:, :D], dim=1)                                          # VTB:107-108
    img_feat = y[:, D:D + numC_Trans]
    depth_prob = volume_interaction(sd, p + ".volume_interaction", stereo, lss)
    geom = get_geometry(frustum, left["rots"], left["trans"], left["intrins"], left["post_rots"],
                        left["post_trans"], left["bda"])
    bev = lift_splat(depth_prob, img_feat, geom, dx, bx, nx)
    if stages is not None:
        stages.update(stereo_fea=fea, gwc_warp=vol, stereo_prob=stereo, depth_net=y, lss_prob=lss,
                      depth_prob=depth_prob, geom=geom, bev_feat=bev)
    return bev, depth_prob


# ------------------------------------------------------------------------------------------
# (iv) 3-D encoder, neck, head, upsample
# ------------------------------------------------------------------------------------------
def _basic_block3d(sd, p, x, stride, groups=32):
    """R3D:35-65 with GroupNorm(32) norm layers (config norm_cfg, stereoscene.py:55)."""
    y = F.relu(_gn(sd, p + ".bn1", F.conv3d(x, sd[p + ".conv1.weight"], None, stride=stride, padding=1), groups))
    y = _gn(sd, p + ".bn2", F.conv3d(y, sd[p + ".conv2.weight"], None, padding=1), groups)
    if (p + ".downsample.0.weight") in sd:
        x = _gn(sd, p + ".downsample.1", F.conv3d(x, sd[p + ".downsample.0.weight"], None, stride=stride), groups)
    return F.relu(y + x)


def resnet3d(sd, p, x, strides=(1, 2, 2), blocks=(2, 2, 2), groups=32):
    """``CustomResNet3D.forward`` (R3D:219-246), depth 18, 3 stages."""
    x = F.relu(_gn(sd, p + ".input_proj.1", F.conv3d(x, sd[p + ".input_proj.0.weight"], None), groups))
    res = []
    for i, (s, n) in enumerate(zip(strides, blocks)):
        for j in range(n):
            x = _basic_block3d(sd, f"{p}.layers.{i}.{j}", x, s if j == 0 else 1, groups)
        res.append(x)
    return res


def second_fpn3d(sd, p, xs, strides=(1, 2, 4), groups=32):
    """``SECONDFPN3D.forward`` (FPN:97-117): ConvTranspose3d(k=s, stride=s, no bias) + GN + ReLU
    per level, channel concat."""
    ups = []
    for i, (x, s) in enumerate(zip(xs, strides)):
        y = F.conv_transpose3d(x, sd[f"{p}.deblocks.{i}.0.weight"], None, stride=s)
        ups.append(F.relu(_gn(sd, f"{p}.deblocks.{i}.1", y, groups)))
    return torch.cat(ups, dim=1)


def occ_head(sd, p, x, groups=32):
    """``OccHead.forward_voxel`` (OCC:220-228; layers OCC:96-108)."""
    y = F.conv3d(x, sd[p + ".occ_convs.0.0.weight"], None, padding=1)
    y = F.relu(_gn(sd, p + ".occ_convs.0.1", y, groups))
    return F.conv3d(y, sd[p + ".occ_convs.0.3.weight"], None)


def upsample_logits(x, size):
    """DET:293-294."""
    return F.interpolate(x, size=tuple(size), mode="trilinear", align_corners=False)


def volumetric_forward(sd, xl, xr, left, right, calib, grid_config, input_size, occ_size,
                       downsample=8, numC_Trans=128, stages=None):
    """The whole path named by the north star, as ``BEVDepthOccupancy.simple_test`` runs it
    after the image encoder (DET:109-124, 275-297).  Returns logits [B,20,*occ_size]."""
    bev, depth_prob = view_transformer(sd, "img_view_transformer", xl, xr, left, right, calib, grid_config,
                                       input_size, downsample, numC_Trans, stages)
    levels = resnet3d(sd, "img_bev_encoder_backbone", bev)
    neck = second_fpn3d(sd, "img_bev_encoder_neck", levels)
    logits = occ_head(sd, "pts_bbox_head", neck)
    up = upsample_logits(logits, occ_size)
    if stages is not None:
        stages.update(enc0=levels[0], enc1=levels[1], enc2=levels[2], neck=neck, logits=logits, logits_up=up)
    return up


# ------------------------------------------------------------------------------------------
# next row N4: semantic-scene-completion scores of a label volume
# ------------------------------------------------------------------------------------------
def ssc_scores(y_pred, y_true, nonempty=None, nonsurface=None, n_classes=20, ignore=255):
    """``SSCMetrics.update`` arithmetic (utils/ssc_metric.py:62-85 with ``get_score_completion``
    :109-141 and ``get_score_semantic_and_completion`` :143-168), as int64 counts.

    The reference edits ``y_pred`` / ``y_true`` in place inside the first score function
    (``predict[target == 255] = 0; target[target == 255] = 0``), so by the time ``update`` builds
    the mask of the semantic scores (``y_true != 255``) no ignored voxel is left: ignored voxels are
    counted as (target 0, prediction 0) there, while the completion scores exclude them.
    Returns dict(completion=int64[3] (tp, fp, fn), tps, fps, fns = int64[n_classes])."""
    yp, yt = y_pred.clone().long(), y_true.clone().long()
    sel_c = yt != ignore
    if nonempty is not None:
        sel_c = sel_c & nonempty.bool()
    if nonsurface is not None:
        sel_c = sel_c & nonsurface.bool()
    ign = yt == ignore
    yp[ign] = 0
    yt[ign] = 0
    bp, bt = yp > 0, yt > 0
    completion = torch.stack([(bt & bp & sel_c).sum(), (~bt & bp & sel_c).sum(), (bt & ~bp & sel_c).sum()]).long()
    sel_s = torch.ones_like(yt, dtype=torch.bool) if nonempty is None else nonempty.bool()
    tps = torch.zeros(n_classes, dtype=torch.long)
    fps, fns = tps.clone(), tps.clone()
    for j in range(n_classes):
        tps[j] = ((yt == j) & (yp == j) & sel_s).sum()
        fps[j] = ((yt != j) & (yp == j) & sel_s).sum()
        fns[j] = ((yt == j) & (yp != j) & sel_s).sum()
    return dict(completion=completion, tps=tps, fps=fps, fns=fns)


def ssc_compute(completion, tps, fps, fns):
    """``SSCMetrics.compute`` (utils/ssc_metric.py:87-107)."""
    ctp, cfp, cfn = [completion[i].double() for i in range(3)]
    iou_ssc = tps.double() / (tps.double() + fps.double() + fns.double() + 1e-5)
    return dict(precision=float(ctp / (ctp + cfp)), recall=float(ctp / (ctp + cfn)), iou=float(ctp / (ctp + cfp + cfn)),
                iou_ssc=iou_ssc, iou_ssc_mean=float(iou_ssc[1:].mean()))
