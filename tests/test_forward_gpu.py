"""End-to-end parity of the volumetric path through the registry-built modules and the C ABI:
  * tiny geometry: every stage boundary against the golden fixtures written by the reference's own
    forward (tests/golden/golden_tiny.npz) and against the CPU oracle,
  * shipped geometry (stereoscene.py, 256x256x32): size-independent properties.
"""
import numpy as np
import pytest
import torch

from oracle import restatement as O
from util import build_model, cpu_state_dict, golden_tiny, rel_err, tiny_inputs

pytestmark = pytest.mark.gpu

# rel-to-max tolerances for the *accumulated* error at each stage boundary of the full forward
STAGE_TOL = {
    "precise": dict(stereo_fea=2e-5, gwc_warp=2e-5, stereo_prob=1e-4, lss_prob=1e-3, depth_prob=1e-3, bev_feat=1e-3,
                    enc0=1e-3, enc1=1e-3, enc2=1e-3, logits=1e-3, logits_up=1e-3),
}


def _run_tiny(math_mode):
    from stereoscene_b200 import ops
    cfg, gold = golden_tiny()
    model, mc = build_model("tiny", cfg["seed"], device="cuda")
    xl, xr, left, right, calib = tiny_inputs(cfg, device="cuda")
    ops.set_default_math(math_mode)
    vt = model.img_view_transformer
    vt.stage_outputs = {}
    with torch.no_grad():
        out = model.forward_features(xl, xr, left, right, calib, occ_size=cfg["occ_size"], want_labels=True)
    torch.cuda.synchronize()
    st = dict(vt.stage_outputs)
    vt.stage_outputs = None
    ops.set_default_math(ops.SS_MATH_TF32)
    return cfg, gold, model, out, st


def _stage_tensors(out, st):
    """Bring our stage tensors to the reference's logical layouts."""
    fea = st["stereo_fea"].squeeze(1).permute(0, 3, 1, 2)                       # [2B,64,H,W]
    return {
        "stereo_fea": fea, "gwc_warp": st["gwc_warp"].permute(0, 4, 1, 2, 3), "stereo_prob": st["stereo_prob"],
        "lss_prob": st["lss_prob"], "depth_prob": st["depth_prob"],
        "bri_lss2stereo": st["bri"][..., 0].unsqueeze(1), "bri_stereo2lss": st["bri"][..., 1].unsqueeze(1),
        "mie_hourglass": st["mie_hourglass"].permute(0, 4, 1, 2, 3), "mie_ca3d": st["mie_ca3d"].permute(0, 4, 1, 2, 3),
        "logits": out["logits_lowres"], "logits_up": out["output_voxels"],
    }


def test_tiny_forward_precise_matches_reference_golden():
    from stereoscene_b200 import ops
    cfg, gold, model, out, st = _run_tiny(ops.SS_MATH_3XTF32)
    got = _stage_tensors(out, st)
    report = {k: rel_err(v, gold[k]) for k, v in got.items()}
    print("precise:", {k: f"{v:.2e}" for k, v in report.items()})
    for k, v in report.items():
        assert v < 1e-3, (k, v)
    assert report["stereo_fea"] < 5e-5 and report["gwc_warp"] < 5e-5
    # geometry / index path: bit-exact against the reference's geom -> indices
    dx, bx, nx = O.gen_dx_bx(cfg["grid_config"]["xbound"], cfg["grid_config"]["ybound"], cfg["grid_config"]["zbound"])
    want_idx, want_kept = O.voxel_indices(torch.from_numpy(gold["geom"]), dx, bx, nx)
    idx = ops.splat_build_index(torch.from_numpy(gold["geom"]).cuda(), dx.tolist(), bx.tolist(), nx.tolist(), want_coords=True)
    c = idx.coords.cpu()
    assert torch.equal(c[:, 3].bool(), want_kept) and torch.equal(c[want_kept, :3].long(), want_idx[want_kept])
    # labels = argmax of the upsampled logits
    assert torch.equal(out["labels"].cpu().long(), out["output_voxels"].argmax(dim=1).cpu())


def test_tiny_forward_tf32_within_north_star_tolerance():
    from stereoscene_b200 import ops
    cfg, gold, model, out, st = _run_tiny(ops.SS_MATH_TF32)
    got = _stage_tensors(out, st)
    report = {k: rel_err(v, gold[k]) for k, v in got.items()}
    print("tf32:", {k: f"{v:.2e}" for k, v in report.items()})
    # TF32 multiplies (what the reference's own GPU convs use: cudnn.allow_tf32 defaults to True);
    # error accumulates over ~60 layers, so the end-to-end bound is looser than the per-layer 1e-3
    for k, v in report.items():
        assert v < 2e-2, (k, v)
    assert report["stereo_fea"] < 1e-3 and report["gwc_warp"] < 1e-3


def test_module_level_dropin_signatures():
    """Each registered module called on its own with the reference's tensor contract
    (logical NCDHW tensors in and out), against the oracle."""
    from stereoscene_b200 import ops
    cfg, gold = golden_tiny()
    model, mc = build_model("tiny", cfg["seed"], device="cuda")
    sd = cpu_state_dict(model)
    ops.set_default_math(ops.SS_MATH_3XTF32)
    try:
        bev = torch.from_numpy(gold["bev_feat"])                          # plain NCDHW, as the reference passes it
        with torch.no_grad():
            levels = model.img_bev_encoder_backbone(bev.cuda())
            want_levels = O.resnet3d(sd, "img_bev_encoder_backbone", bev)
            for got, want in zip(levels, want_levels):
                assert got.shape == want.shape and rel_err(got, want) < 2e-4
            neck = model.img_bev_encoder_neck([w.cuda() for w in want_levels])
            want_neck = O.second_fpn3d(sd, "img_bev_encoder_neck", want_levels)
            assert isinstance(neck, list) and rel_err(neck[0], want_neck) < 2e-4
            head = model.pts_bbox_head(voxel_feats=[want_neck.cuda()])
            assert set(head) == {"output_voxels", "output_points"} and head["output_points"] is None
            assert rel_err(head["output_voxels"][0], O.occ_head(sd, "pts_bbox_head", want_neck)) < 2e-4
            # voxel_pooling with the reference's signature (geom, lifted volume)
            vt = model.img_view_transformer
            B, D, H, W = gold["depth_prob"].shape
            dp = torch.from_numpy(gold["depth_prob"])
            feat = torch.randn(B, 128, H, W, generator=torch.Generator().manual_seed(1))
            lifted = (dp.unsqueeze(1) * feat.unsqueeze(2)).view(B, 1, 128, D, H, W).permute(0, 1, 3, 4, 5, 2)
            geom = torch.from_numpy(gold["geom"])
            got = vt.voxel_pooling(geom.cuda(), lifted.contiguous().cuda())
            dx, bx, nx = O.gen_dx_bx(cfg["grid_config"]["xbound"], cfg["grid_config"]["ybound"], cfg["grid_config"]["zbound"])
            want = O.lift_splat(dp, feat, geom, dx, bx, nx)
            assert got.shape == want.shape and rel_err(got, want) < 1e-6
    finally:
        ops.set_default_math(ops.SS_MATH_TF32)


def test_shipped_geometry_properties():
    """stereoscene.py as shipped (384x1280 -> 48x160x112 frustum -> 128x128x16 -> 256x256x32 logits)."""
    from stereoscene_b200 import ops, synth
    model, mc = build_model("config2", 0, device="cuda")
    xl, xr = synth.stereo_features(1, mc["input_size"], 8, seed=0, device="cuda")
    left, right, calib = synth.kitti_calibration(1, mc["input_size"], device="cuda")
    vt = model.img_view_transformer
    vt.stage_outputs = {}
    with torch.no_grad():
        out = model.forward_features(xl, xr, left, right, calib, occ_size=mc["occ_size"], want_labels=True)
        out2 = model.forward_features(xl, xr, left, right, calib, occ_size=mc["occ_size"], want_labels=True)
    torch.cuda.synchronize()
    st = vt.stage_outputs
    vt.stage_outputs = None
    logits = out["output_voxels"]
    assert tuple(logits.shape) == (1, 20, 256, 256, 32) and bool(torch.isfinite(logits).all())
    assert tuple(out["depth"].shape) == (1, 112, 48, 160)
    ones = torch.ones(1, 48, 160, device="cuda")
    assert rel_err(out["depth"].sum(1), ones) < 1e-5 and rel_err(st["stereo_prob"].sum(1), ones) < 1e-5
    # splat conserves mass: sum over voxels == sum over kept frustum points of depth_prob * feature
    idx = st["splat_index"]
    kept = int(idx.voxel_start[-1])
    assert 0.3 < kept / idx.order.numel() < 0.7                    # KITTI-like calibration keeps ~47 %
    # the run is repeatable (fp atomics only touch the GroupNorm sums, in double)
    assert rel_err(out2["output_voxels"], logits) < 1e-4
    assert float((out2["labels"] != out["labels"]).float().mean()) < 1e-3


def test_volumetric_engine_matches_eager_and_keeps_order():
    """Serving runtime (CUDA graph per slot + copy stream): infer() equals the eager forward on the same pair,
    and stream() yields one label volume per pair, in input order, also when pairs differ."""
    from stereoscene_b200 import ops
    from stereoscene_b200.runtime import VolumetricEngine
    cfg, _ = golden_tiny()
    model, mc = build_model("tiny", cfg["seed"], device="cuda")
    xl, xr, left, right, calib = tiny_inputs(cfg, device="cuda")
    ops.set_default_math(ops.SS_MATH_TF32)
    eng = VolumetricEngine(model, left, right, calib, cfg["occ_size"], tuple(xl.shape))
    g = torch.Generator().manual_seed(11)
    pairs = []
    for i in range(5):
        a = (xl.cpu() + 0.25 * i * torch.randn(xl.shape, generator=g)).pin_memory()
        b = (xr.cpu() + 0.25 * i * torch.randn(xr.shape, generator=g)).pin_memory()
        pairs.append((a, b))
    want = []
    with torch.no_grad():
        for a, b in pairs:
            out = model.forward_features(a.cuda(), b.cuda(), left, right, calib, occ_size=cfg["occ_size"], want_labels=True)
            want.append(out["labels"].cpu().clone())
    got_single = eng.infer(*pairs[0]).clone()
    assert got_single.shape == want[0].shape and got_single.dtype == torch.uint8
    assert (got_single != want[0]).float().mean().item() < 1e-3      # atomics order may flip a near-tie argmax
    got = [lab.clone() for lab in eng.stream(pairs)]
    assert len(got) == len(pairs)
    for i, (w, gg) in enumerate(zip(want, got)):
        assert (gg != w).float().mean().item() < 1e-3, i
    # different pairs really give different volumes, so an ordering mix-up would be seen
    assert (want[0] != want[4]).any()


def test_engine_graphs_survive_cache_eviction_by_another_calibration():
    """Two engines with different calibrations on one model (KITTI calibration differs per sequence): the second one
    evicts / replaces the shared cache entries (disparity taps, splat index) the first engine's CUDA graphs point at.
    The first engine must keep replaying the same labels -- it owns references to everything its graphs read."""
    from stereoscene_b200 import ops
    from stereoscene_b200.runtime import VolumetricEngine
    cfg, _ = golden_tiny()
    model, mc = build_model("tiny", cfg["seed"], device="cuda")
    xl, xr, left, right, calib = tiny_inputs(cfg, device="cuda")
    ops.set_default_math(ops.SS_MATH_TF32)
    a, b = xl.cpu().pin_memory(), xr.cpu().pin_memory()
    eng1 = VolumetricEngine(model, left, right, calib, cfg["occ_size"], tuple(xl.shape))
    want = eng1.infer(a, b).clone()
    for i in range(10):                                     # more calibrations than any cache holds entries
        left2 = {k: v.clone() for k, v in left.items()}
        right2 = {k: v.clone() for k, v in right.items()}
        left2["trans"] = left2["trans"] + 0.37 * (i + 1)
        right2["trans"] = right2["trans"] + 0.37 * (i + 1)
        eng2 = VolumetricEngine(model, left2, right2, calib * (0.5 + 0.05 * i), cfg["occ_size"], tuple(xl.shape), warmup=1)
        other = eng2.infer(a, b).clone()
        del eng2
    assert (other != want).any()                            # the other calibration really gives another volume
    torch.cuda.empty_cache()
    junk = [torch.full((1 << 20,), float("nan"), device="cuda") for _ in range(64)]      # recycle freed blocks with garbage
    del junk
    again = eng1.infer(a, b)
    assert (again != want).float().mean().item() < 1e-3


def test_two_stream_frustum_stage_survives_graph_replays():
    """depth_net runs on a side stream beside the stereo branch (plugin/view_transformer.py).  At the shipped size two
    kernels of different streams then really run at once, which once exposed a barrier-phase race of the box kernel
    (conv3d_tc.cu, TcCfg::STAGES must be even): 60 replays of the step's CUDA graph must reproduce the eager output bit for bit,
    and the serialised forward (STEREOSCENE_B200_STREAM_OVERLAP=0 semantics) must give the same result."""
    from stereoscene_b200 import ops, synth
    from stereoscene_b200.plugin import view_transformer as VT
    ops.set_math_policy(None)
    model, mc = build_model("config2", 0, device="cuda")
    xl, xr = synth.stereo_features(1, mc["input_size"], 8, seed=0, device="cuda")
    left, right, calib = synth.kitti_calibration(1, mc["input_size"], device="cuda")
    f = lambda: model.forward_features(xl, xr, left, right, calib, occ_size=mc["occ_size"], want_labels=True)   # noqa: E731
    assert VT._STREAM_OVERLAP
    with torch.no_grad():
        a = f()
        b = f()
        torch.cuda.synchronize()
        assert torch.equal(a["output_voxels"], b["output_voxels"])
        VT._STREAM_OVERLAP = False
        try:
            serial = f()
        finally:
            VT._STREAM_OVERLAP = True
        torch.cuda.synchronize()
        assert torch.equal(serial["output_voxels"], a["output_voxels"]) and torch.equal(serial["labels"], a["labels"])
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = f()
        for _ in range(60):
            g.replay()
        torch.cuda.synchronize()
    assert torch.equal(out["output_voxels"], a["output_voxels"]) and torch.equal(out["labels"], a["labels"])
