"""TEST INFRASTRUCTURE ONLY -- generates the golden fixtures under tests/golden/ by running the
reference's own, unmodified module code (oracle/ref_loader.py) on seeded synthetic inputs.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
The fixtures are committed; the GPU box (which has no reference tree) only reads them.

Fixtures
  golden_tiny.npz      per-stage outputs of the reference forward for the "tiny" geometry
                       (B=2, 64x128 input -> 8x16 px, 48 depth bins, 16x16x4 LSS grid)
  state_dict_spec.json key -> shape of the reference model's state_dict at the shipped config
                       (the checkpoint-compatibility contract)
Weights are not stored: they are a pure function of (key, shape, seed) -- see
stereoscene_b200/synth.py:randomize_state_dict -- so the same values are rebuilt anywhere.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_loader as R                     # noqa: E402
from stereoscene_b200 import synth                     # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

TINY = dict(
    batch=2, input_size=(64, 128), downsample=8,
    grid_config=dict(xbound=[0.0, 51.2, 3.2], ybound=[-25.6, 25.6, 3.2], zbound=[-2.0, 4.4, 1.6],
                     dbound=[2.0, 26.0, 0.5]),
    occ_size=[32, 32, 8], seed=7, calib_scale=[1.0, 0.37],
)

SHIPPED = dict(
    input_size=(384, 1280), downsample=8,
    grid_config=dict(xbound=[0, 51.2, 0.4], ybound=[-25.6, 25.6, 0.4], zbound=[-2, 4.4, 0.4],
                     dbound=[2.0, 58.0, 0.5]),
    occ_size=[256, 256, 32],
)

GN32 = dict(type="GN", num_groups=32, requires_grad=True)


def build_reference_model(cfg):
    """The four hot-path modules of stereoscene.py:75-123, built from the reference classes."""
    vt = R.vt_module().ViewTransformerLiftSplatShootVoxel(
        loss_depth_weight=1.0, downsample=cfg["downsample"], numC_input=640, cam_channels=30,
        semkitti=False, grid_config=cfg["grid_config"], data_config={"input_size": cfg["input_size"]},
        numC_Trans=128, vp_megvii=False)
    enc = R.resnet3d_module().CustomResNet3D(depth=18, num_stage=3, n_input_channels=128,
                                            block_inplanes=[128, 256, 512], out_indices=(0, 1, 2), norm_cfg=GN32)
    neck = R.neck_module().SECONDFPN3D(norm_cfg=GN32, in_channels=[128, 256, 512], upsample_strides=[1, 2, 4],
                                       out_channels=[128, 128, 128])
    head = R.occhead_module().OccHead(num_level=1, in_channels=[384], out_channel=20, semantic_kitti=True,
                                     point_cloud_range=[0, -25.6, -2, 51.2, 25.6, 4.4], supervise_points=False,
                                     sampling_img_feats=True, in_img_channels=640, soft_weights=True,
                                     semkitti_loss_weight_cfg={"voxel_ce": 1.0, "voxel_sem_scal": 1.0,
                                                               "voxel_geo_scal": 1.0, "voxel_ohem": 0.0,
                                                               "voxel_lovasz": 0.0, "frustum_dist": 0.0},
                                     train_cfg=None, test_cfg=None)
    model = torch.nn.ModuleDict(dict(img_view_transformer=vt, img_bev_encoder_backbone=enc,
                                     img_bev_encoder_neck=neck, pts_bbox_head=head))
    return model.eval()


def synthetic_inputs(cfg):
    B = cfg["batch"]
    xl, xr = synth.stereo_features(B, cfg["input_size"], cfg["downsample"], seed=cfg["seed"])
    left, right, calib = synth.kitti_calibration(B, cfg["input_size"])
    calib = calib * torch.tensor(cfg.get("calib_scale", [1.0] * B)).view(B, 1)
    return xl, xr, left, right, calib


@torch.no_grad()
def run_reference(model, cfg, xl, xr, left, right, calib):
    """Reference forward with per-stage capture (hooks only; no reference code is edited)."""
    vt = model["img_view_transformer"]
    st = {}
    hooks = [
        vt.stereo_volume_net.feature_withcam.register_forward_hook(lambda m, i, o: st.__setitem__("stereo_fea", o)),
        vt.stereo_volume_net.dres0.register_forward_pre_hook(lambda m, i: st.__setitem__("gwc_warp", i[0])),
        vt.stereo_volume_net.register_forward_hook(lambda m, i, o: st.__setitem__("stereo_prob", o["single_channel"])),
        vt.depth_net.register_forward_hook(lambda m, i, o: st.__setitem__("depth_net", o)),
        vt.volume_interaction.register_forward_pre_hook(lambda m, i: st.__setitem__("lss_prob", i[1])),
        vt.volume_interaction.lss2stereo.register_forward_hook(lambda m, i, o: st.__setitem__("bri_lss2stereo", o)),
        vt.volume_interaction.stereo2lss.register_forward_hook(lambda m, i, o: st.__setitem__("bri_stereo2lss", o)),
        vt.volume_interaction.dres1.register_forward_hook(lambda m, i, o: st.__setitem__("mie_hourglass", o)),
        vt.volume_interaction.CA3D.register_forward_hook(lambda m, i, o: st.__setitem__("mie_ca3d", o)),
    ]
    ml = vt.get_mlp_input(left["rots"], left["trans"], left["intrins"], left["post_rots"], left["post_trans"], left["bda"])
    mr = vt.get_mlp_input(right["rots"], right["trans"], right["intrins"], right["post_rots"], right["post_trans"], right["bda"])
    geo_l = [left[k] for k in ("rots", "trans", "intrins", "post_rots", "post_trans", "bda")] + [ml]
    geo_r = [right[k] for k in ("rots", "trans", "intrins", "post_rots", "post_trans", "bda")] + [mr]
    with R.cpu_arange():
        bev, depth_prob = vt([xl] + geo_l + [xr] + geo_r + [calib] + [None, None])
    for h in hooks:
        h.remove()
    st["mlp_input"] = ml
    st["geom"] = vt.get_geometry(*[left[k] for k in ("rots", "trans", "intrins", "post_rots", "post_trans", "bda")])
    st["bev_feat"], st["depth_prob"] = bev, depth_prob
    levels = model["img_bev_encoder_backbone"](bev)
    neck = model["img_bev_encoder_neck"](levels)
    out = model["pts_bbox_head"](voxel_feats=neck)
    logits = out["output_voxels"][0]
    up = torch.nn.functional.interpolate(logits, size=tuple(cfg["occ_size"]), mode="trilinear", align_corners=False)
    st.update(enc0=levels[0], enc1=levels[1], enc2=levels[2], neck=neck[0], logits=logits, logits_up=up)
    return st


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)

    # ---- state_dict contract at the shipped config
    full = build_reference_model(SHIPPED)
    spec = {k: list(v.shape) for k, v in full.state_dict().items()}
    with open(os.path.join(GOLDEN_DIR, "state_dict_spec.json"), "w") as f:
        json.dump(spec, f, indent=0, sort_keys=True)
    del full

    # ---- the model dict of the shipped config, as OUR loader reads the reference's file
    from stereoscene_b200.config import Config
    cfg = Config.fromfile(os.path.join(R.REF_ROOT, "projects/configs/occupancy/semantickitti/stereoscene.py"))
    data_dir = os.path.join(os.path.dirname(GOLDEN_DIR), "..", "stereoscene_b200", "data")
    os.makedirs(data_dir, exist_ok=True)
    with open(os.path.join(data_dir, "stereoscene_model_cfg.json"), "w") as f:
        json.dump(dict(model=cfg.to_dict()["model"], occ_size=cfg.occ_size, point_cloud_range=cfg.point_cloud_range,
                       lss_downsample=cfg.lss_downsample, class_names=cfg.class_names), f, indent=1)

    # ---- tiny geometry: every stage boundary
    model = build_reference_model(TINY)
    synth.randomize_weights_(model, TINY["seed"])
    xl, xr, left, right, calib = synthetic_inputs(TINY)
    st = run_reference(model, TINY, xl, xr, left, right, calib)
    keep = ("stereo_fea", "gwc_warp", "stereo_prob", "lss_prob", "bri_lss2stereo", "bri_stereo2lss",
            "mie_hourglass", "mie_ca3d", "depth_prob", "geom", "bev_feat", "enc0", "enc1", "enc2", "logits",
            "logits_up", "mlp_input")
    arrays = {k: st[k].contiguous().numpy() for k in keep}
    arrays["depth_net_ctx_sample"] = st["depth_net"][:, :, ::2, ::2].contiguous().numpy()
    arrays["neck_sample"] = st["neck"][:, ::8].contiguous().numpy()
    arrays["calib"] = calib.numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "golden_tiny.npz"), **arrays)
    with open(os.path.join(GOLDEN_DIR, "golden_tiny.json"), "w") as f:
        json.dump(TINY, f, indent=1)
    tot = sum(a.nbytes for a in arrays.values())
    print(f"golden_tiny: {len(arrays)} arrays, {tot/1e6:.2f} MB raw")
    for k, a in arrays.items():
        print(f"  {k:20s} {tuple(a.shape)}  mean {a.mean():+.4e}  std {a.std():.4e}")


if __name__ == "__main__":
    main()
