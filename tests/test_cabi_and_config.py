"""CPU-side checks: the C-ABI library loads and exports every symbol include/*.h declares, the
ctypes binding covers exactly that set, the config loader reads the reference's config file
unchanged, and the registry builds the model under the reference's type names."""
import ctypes
import json
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CFG = "/root/reference/projects/configs/occupancy/semantickitti/stereoscene.py"


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "stereoscene_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ss_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    from stereoscene_b200 import _build, cabi
    path = _build.library_path()
    assert os.path.exists(path), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(cabi.SIGNATURES) == declared, "ctypes binding and header disagree"
    loaded = cabi.load(path)
    assert loaded.ss_abi_version() == cabi.ABI_VERSION
    assert loaded.ss_launch_count() >= 0


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """The ctypes mirrors of ss_conv3d_desc / ss_conv3d_join have the C compiler's size and field offsets
    (a plain-C translation unit including the public header is compiled with gcc: the header is C, not C++)."""
    import shutil
    import subprocess
    from stereoscene_b200 import cabi
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    fields_desc = [n for n, _ in cabi.ConvDesc._fields_]
    fields_join = [n for n, _ in cabi.ConvJoin._fields_]
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "stereoscene_b200.h"', 'int main(void) {',
           '  printf("%zu %zu %d\\n", sizeof(ss_conv3d_desc), sizeof(ss_conv3d_join), SS_ABI_VERSION);']
    for f in fields_desc:
        src.append(f'  printf("%zu\\n", offsetof(ss_conv3d_desc, {f}));')
    for f in fields_join:
        src.append(f'  printf("%zu\\n", offsetof(ss_conv3d_join, {f}));')
    src.append('  return 0; }')
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert [int(out[0]), int(out[1]), int(out[2])] == [ctypes.sizeof(cabi.ConvDesc), ctypes.sizeof(cabi.ConvJoin), cabi.ABI_VERSION]
    offs = [int(x) for x in out[3:]]
    want = [getattr(cabi.ConvDesc, f).offset for f in fields_desc] + [getattr(cabi.ConvJoin, f).offset for f in fields_join]
    assert offs == want


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from stereoscene_b200 import cabi
    monkeypatch.setattr(cabi, "_lib", None)
    with pytest.raises(cabi.NativeLibraryError):
        cabi.load(str(tmp_path / "nope.so"))
    monkeypatch.setattr(cabi, "_lib", None)


def test_ops_refuse_cpu_tensors():
    from stereoscene_b200 import ops
    with pytest.raises(RuntimeError):
        ops.softmax_d(torch.randn(1, 4, 3, 3))
    with pytest.raises(RuntimeError):
        ops.to_channels_last(torch.randn(1, 4, 3, 3))


def test_config_inheritance_semantics(tmp_path):
    from stereoscene_b200.config import Config
    (tmp_path / "base.py").write_text("a = dict(x=1, y=dict(p=1, q=2), lst=[1, 2, 3])\nb = 5\n")
    (tmp_path / "child.py").write_text(
        "_base_ = ['./base.py']\nimport os\na = dict(y=dict(q=3), lst=[9])\nc = dict(_delete_=True, z=1)\n")
    cfg = Config.fromfile(str(tmp_path / "child.py"))
    assert cfg.a.x == 1 and cfg.a.y.p == 1 and cfg.a.y.q == 3 and cfg.a.lst == [9] and cfg.b == 5
    assert "os" not in cfg and cfg.c.z == 1
    cfg.merge_from_dict({"a.y.q": 7, "new.key": [1]})
    assert cfg.a.y.q == 7 and cfg.new.key == [1]


@pytest.mark.skipif(not os.path.exists(REF_CFG), reason="reference tree not present")
def test_reference_config_loads_unchanged_and_builds():
    """projects/configs/occupancy/semantickitti/stereoscene.py through our loader + registries."""
    import projects.mmdet3d_plugin  # noqa: F401  (plugin_dir = 'projects/mmdet3d_plugin/' resolves to our drop-in)
    from stereoscene_b200 import presets
    from stereoscene_b200.config import Config
    from stereoscene_b200.registry import BACKBONES, DETECTORS, HEADS, NECKS, build_model
    cfg = Config.fromfile(REF_CFG)
    assert cfg.plugin and cfg.plugin_dir == "projects/mmdet3d_plugin/"
    assert cfg.model.type in DETECTORS and cfg.model.img_view_transformer.type in NECKS
    assert cfg.model.img_bev_encoder_backbone.type in BACKBONES and cfg.model.pts_bbox_head.type in HEADS
    # the committed model dict (used on the GPU box, which has no reference tree) is this file's
    shipped = presets.shipped_config()
    assert json.loads(json.dumps(cfg.to_dict()["model"])) == shipped["model"]
    assert list(cfg.occ_size) == shipped["occ_size"]
    model = build_model(cfg.model, train_cfg=cfg.get("train_cfg"), test_cfg=cfg.get("test_cfg"))
    vt = model.img_view_transformer
    assert vt.D == 112 and [int(v) for v in vt.nx] == [128, 128, 16] and tuple(vt.frustum.shape) == (112, 48, 160, 3)
    assert model.pts_bbox_head.occ_convs[0][0].weight.shape == (192, 384, 3, 3, 3)


def test_presets_rederive_geometry():
    from stereoscene_b200 import presets
    mc = presets.model_config("config1")
    g = mc["model"]["img_view_transformer"]["grid_config"]
    assert g["xbound"][2] == pytest.approx(0.8) and mc["occ_size"] == [128, 128, 16]
    m0, mc0 = presets.build("config0")
    assert [int(v) for v in m0.img_view_transformer.nx] == [32, 32, 4]          # BASELINE configs[0]: 32x32x4 LSS grid
    assert tuple(m0.img_view_transformer.frustum.shape) == (112, 16, 32, 3)


def test_unsupported_options_raise():
    from stereoscene_b200.plugin import CustomResNet3D, OccHead, ViewTransformerLiftSplatShootVoxel
    with pytest.raises(NotImplementedError):
        ViewTransformerLiftSplatShootVoxel(loss_depth_weight=1.0, imgseg=True)
    with pytest.raises(NotImplementedError):
        CustomResNet3D(depth=50)
    with pytest.raises(NotImplementedError):
        OccHead(in_channels=[384], out_channel=20, supervise_points=True)


def test_plane_kernel_dispatch_mirror():
    """ops._halo_or_march_layer mirrors the C dispatch (conv3d_halo.cu:try_conv_halo, conv3d_march.cu:try_conv_march32) for
    the layers of stereoscene.py: it decides whether a small pending input is materialised before a per-tap layer."""
    import torch.nn as nn
    from stereoscene_b200 import ops
    f = lambda m, d, h, w: ops._halo_or_march_layer(ops.PackedConv(m), d, h, w, m.in_channels)       # noqa: E731
    assert f(nn.Conv3d(384, 192, 3, 1, 1, bias=False), 128, 128, 16)          # occupancy head: halo kernel
    assert f(nn.Conv3d(128, 128, 3, 1, 1, bias=False), 128, 128, 16)          # encoder stage 0
    assert f(nn.Conv2d(640, 640, 3, 1, 1, bias=False), 1, 48, 160)            # DepthNet 2-D layers (axes swapped)
    assert f(nn.Conv3d(32, 32, 3, 1, 1, bias=False), 112, 48, 160)            # frustum layers: marching kernel
    assert f(nn.Conv3d(64, 64, 3, 1, 1, bias=False), 56, 24, 80)              # hourglass conv2 (swapped: 83 % tile use)
    assert not f(nn.Conv3d(512, 512, 3, 1, 1, bias=False), 32, 32, 4)         # stage 2: per-tap kernel -> materialise
    assert not f(nn.Conv3d(128, 128, 3, 1, 1, bias=False), 28, 12, 40)        # hourglass conv4
    assert not f(nn.Conv3d(128, 256, 3, 2, 1, bias=False), 128, 128, 16)      # strided
    assert not f(nn.Conv2d(640, 640, 3, padding=6, dilation=6, bias=False), 1, 48, 160)   # dilated ASPP branch
    assert not f(nn.ConvTranspose3d(128, 64, 3, 2, 1, 1, bias=False), 28, 12, 40)
