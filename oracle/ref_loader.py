"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference modules from /root/reference.

This file is part of the oracle (the checker).  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import anything under
``oracle/``.  The product package ``stereoscene_b200`` never does.

The reference (Arlo0o/StereoScene) is pure Python on top of mmcv / mmdet / mmdet3d, none of
which are installed in this image.  This loader makes the reference's own hot-path files
importable *unmodified* (they are executed from where they lie under /root/reference; nothing
is copied) by

  1. registering namespace-only packages for ``projects.mmdet3d_plugin.*`` so the package
     ``__init__`` files (which import the whole project) never run,
  2. providing small stand-ins for the handful of mmcv / mmdet symbols the hot-path files use
     (``BaseModule``, ``build_norm_layer``, ``build_conv_layer``, ``build_upsample_layer``,
     registries, mmdet's 2-D ``BasicBlock``, torchmetrics' ``Metric``),
  3. auto-stubbing every other missing third-party import with an inert module,
  4. supplying the two *external compiled ops* the path reaches:
       - ``mmdet3d.ops.bev_pool.bev_pool``  (call site ViewTransformerLSSVoxel.py:473) as an
         ``index_add_`` restatement of the BEVFusion semantics (sum of the rows of ``feats``
         that share a voxel; output [B, C, D=z, H=x, W=y]),
       - mmcv ``DCN`` (ViewTransformerLSSBEVDepth.py:490-498) via torchvision's
         ``deform_conv2d`` (mmcv DeformConv2dPack: no bias, offsets from ``conv_offset``),
  5. patching ``device='cuda'`` out of ``warp`` (ViewTransformerLSSVoxel.py:140,144) so the
     reference runs on CPU.

It only works where /root/reference exists (the build container).  The GPU box has no
reference tree: there the committed fixtures in ``tests/golden/`` (produced by
``oracle/make_golden.py`` with this loader) and the restatement ``oracle/restatement.py``
are used instead.
"""
from __future__ import annotations

import contextlib
import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("STEREOSCENE_REFERENCE", "/root/reference")
_P = os.path.join(REF_ROOT, "projects")
_PLUG = os.path.join(_P, "mmdet3d_plugin")


def available() -> bool:
    return os.path.isfile(os.path.join(_PLUG, "occupancy", "image2bev", "ViewTransformerLSSVoxel.py"))


# --------------------------------------------------------------------------------------
# registries / builders (only what the hot-path files touch)
# --------------------------------------------------------------------------------------
class _Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return deco(module)
        return deco

    def get(self, key):
        return self.module_dict[key]

    def build(self, cfg, **kw):
        cfg = dict(cfg)
        t = cfg.pop("type")
        return self.module_dict[t](**cfg, **kw)


class _BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg

    def init_weights(self):
        pass


def _identity_decorator_factory(*a, **k):
    def deco(f):
        return f
    return deco


def _build_norm_layer(cfg, num_features, postfix=""):
    cfg = dict(cfg)
    t = cfg.pop("type")
    req = cfg.pop("requires_grad", True)
    if t == "GN":
        layer = nn.GroupNorm(num_groups=cfg.pop("num_groups"), num_channels=num_features, **cfg)
        name = "gn"
    elif t in ("BN", "BN2d"):
        layer = nn.BatchNorm2d(num_features, **cfg)
        name = "bn"
    elif t == "BN3d":
        layer = nn.BatchNorm3d(num_features, **cfg)
        name = "bn"
    elif t == "BN1d":
        layer = nn.BatchNorm1d(num_features, **cfg)
        name = "bn"
    else:
        raise KeyError(t)
    for p in layer.parameters():
        p.requires_grad = req
    return name + str(postfix), layer


class _DCN(nn.Module):
    """mmcv DeformConv2dPack semantics: offsets predicted by ``conv_offset`` (zero-init),
    deformable conv without bias.  Backed by torchvision.ops.deform_conv2d."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, deform_groups=1, bias=False, im2col_step=32):
        super().__init__()
        k = kernel_size
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, k, k))
        nn.init.kaiming_uniform_(self.weight, nonlinearity="relu")
        self.conv_offset = nn.Conv2d(in_channels, deform_groups * 2 * k * k, k, stride, padding, dilation, bias=True)
        nn.init.zeros_(self.conv_offset.weight)
        nn.init.zeros_(self.conv_offset.bias)
        self.stride, self.padding, self.dilation = stride, padding, dilation

    def forward(self, x):
        from torchvision.ops import deform_conv2d
        return deform_conv2d(x, self.conv_offset(x), self.weight, None, self.stride, self.padding, self.dilation)


def _build_conv_layer(cfg, *args, **kwargs):
    cfg = dict(cfg) if cfg is not None else dict(type="Conv2d")
    t = cfg.pop("type")
    table = {"Conv1d": nn.Conv1d, "Conv2d": nn.Conv2d, "Conv3d": nn.Conv3d, "Conv": nn.Conv2d, "DCN": _DCN,
             "Conv2dAdaptivePadding": _Conv2dAdaptivePadding}
    return table[t](*args, **kwargs, **cfg)


def _build_upsample_layer(cfg, *args, **kwargs):
    cfg = dict(cfg)
    t = cfg.pop("type")
    table = {"deconv": nn.ConvTranspose2d, "deconv3d": nn.ConvTranspose3d}
    return table[t](*args, **kwargs, **cfg)


# --------------------------------------------------------------------------------------
# the 2-D image encoder's third-party bricks (row N2).  mmcv 1.4.0 / mmdet 2.14.0 / mmdet3d 0.17.1
# are the versions the reference pins (docs/install.md); none is vendored under /root/reference, so
# their PUBLISHED behaviour is restated here, only as far as efficientnet.py and SECONDFPN use it.
# --------------------------------------------------------------------------------------
class _Conv2dAdaptivePadding(nn.Conv2d):
    """mmcv.cnn.bricks.Conv2dAdaptivePadding: TensorFlow 'SAME' padding computed per call from the input
    size -- total = max((ceil(H/s)-1)*s + (k-1)*d + 1 - H, 0), the smaller half in front; the constructor's
    ``padding`` argument is ignored (the layer is built with padding 0)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True):
        super().__init__(in_channels, out_channels, kernel_size, stride, 0, dilation, groups, bias)

    def forward(self, x):
        import math
        import torch.nn.functional as F
        ih, iw = x.shape[-2:]
        kh, kw = self.weight.shape[-2:]
        sh, sw = self.stride
        oh, ow = math.ceil(ih / sh), math.ceil(iw / sw)
        ph = max((oh - 1) * sh + (kh - 1) * self.dilation[0] + 1 - ih, 0)
        pw = max((ow - 1) * sw + (kw - 1) * self.dilation[1] + 1 - iw, 0)
        if ph > 0 or pw > 0:
            x = F.pad(x, [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2])
        return F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)


class _Swish(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(x)


def _build_activation_layer(cfg):
    cfg = dict(cfg)
    t = cfg.pop("type")
    if t == "ReLU":
        cfg.setdefault("inplace", True)
    table = {"ReLU": nn.ReLU, "Swish": _Swish, "Sigmoid": nn.Sigmoid, "GELU": nn.GELU}
    return table[t](**cfg)


class _ConvModule(nn.Module):
    """mmcv.cnn.bricks.ConvModule with the default order (conv, norm, act): bias='auto' means "no bias when a
    norm layer follows"; attribute names ``conv`` / ``bn`` / ``activate`` give the reference's state_dict keys."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias="auto", conv_cfg=None, norm_cfg=None, act_cfg=dict(type="ReLU"), inplace=True, **kw):
        super().__init__()
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == "auto":
            bias = not self.with_norm
        self.conv = _build_conv_layer(conv_cfg, in_channels, out_channels, kernel_size, stride=stride,
                                      padding=padding, dilation=dilation, groups=groups, bias=bias)
        if self.with_norm:
            self.norm_name, norm = _build_norm_layer(norm_cfg, out_channels)
            self.add_module(self.norm_name, norm)
        if self.with_activation:
            self.activate = _build_activation_layer(act_cfg)

    def forward(self, x):
        x = self.conv(x)
        if self.with_norm:
            x = getattr(self, self.norm_name)(x)
        if self.with_activation:
            x = self.activate(x)
        return x


class _DropPath(nn.Module):
    """Stochastic depth: the identity outside training (the path here is inference)."""

    def __init__(self, drop_prob=0.1):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        assert not self.training
        return x


class _Sequential(_BaseModule, nn.Sequential):
    def __init__(self, *args, init_cfg=None):
        _BaseModule.__init__(self, init_cfg)
        nn.Sequential.__init__(self, *args)


class _SELayer(_BaseModule):
    """mmdet.models.utils.SELayer: global average pool -> 1x1 conv (bias) + act[0] -> 1x1 conv (bias) + act[1]
    -> channel gate multiplied onto the input."""

    def __init__(self, channels, ratio=16, conv_cfg=None, act_cfg=(dict(type="ReLU"), dict(type="Sigmoid")),
                 init_cfg=None):
        super().__init__(init_cfg)
        self.global_avgpool = nn.AdaptiveAvgPool2d(1)
        self.conv1 = _ConvModule(channels, int(channels / ratio), 1, stride=1, conv_cfg=conv_cfg, act_cfg=act_cfg[0])
        self.conv2 = _ConvModule(int(channels / ratio), channels, 1, stride=1, conv_cfg=conv_cfg, act_cfg=act_cfg[1])

    def forward(self, x):
        return x * self.conv2(self.conv1(self.global_avgpool(x)))


def _make_divisible(value, divisor, min_value=None, min_ratio=0.9):
    if min_value is None:
        min_value = divisor
    new_value = max(min_value, int(value + divisor / 2) // divisor * divisor)
    if new_value < min_ratio * value:
        new_value += divisor
    return new_value


class SECONDFPN(_BaseModule):
    """mmdet3d.models.necks.SECONDFPN (0.17.1), restated: one deblock per input level -- ConvTranspose2d with
    kernel = stride for upsample_strides >= 1 (a stride of exactly 1 is still a 1x1 transposed conv because
    use_conv_for_no_stride defaults to False), Conv2d with kernel = stride = round(1/s) below 1 -- each without
    bias and followed by BatchNorm2d(eps 1e-3, momentum 0.01) + ReLU; outputs concatenated over channels and
    returned as a one-element list."""

    def __init__(self, in_channels=[128, 128, 256], out_channels=[256, 256, 256], upsample_strides=[1, 2, 4],
                 norm_cfg=dict(type="BN", eps=1e-3, momentum=0.01), upsample_cfg=dict(type="deconv", bias=False),
                 conv_cfg=dict(type="Conv2d", bias=False), use_conv_for_no_stride=False, init_cfg=None):
        super().__init__(init_cfg)
        assert len(out_channels) == len(upsample_strides) == len(in_channels)
        self.in_channels, self.out_channels = in_channels, out_channels
        deblocks = []
        for i, oc in enumerate(out_channels):
            stride = upsample_strides[i]
            if stride > 1 or (stride == 1 and not use_conv_for_no_stride):
                up = _build_upsample_layer(upsample_cfg, in_channels=in_channels[i], out_channels=oc,
                                           kernel_size=upsample_strides[i], stride=upsample_strides[i])
            else:
                stride = int(round(1 / stride))
                up = _build_conv_layer(conv_cfg, in_channels=in_channels[i], out_channels=oc, kernel_size=stride,
                                       stride=stride)
            deblocks.append(nn.Sequential(up, _build_norm_layer(norm_cfg, oc)[1], nn.ReLU(inplace=True)))
        self.deblocks = nn.ModuleList(deblocks)

    def forward(self, x):
        assert len(x) == len(self.in_channels)
        ups = [deblock(x[i]) for i, deblock in enumerate(self.deblocks)]
        out = torch.cat(ups, dim=1) if len(ups) > 1 else ups[0]
        return [out]


class _BasicBlock2d(nn.Module):
    """mmdet 2.14 ``mmdet.models.backbones.resnet.BasicBlock`` (inplanes==planes, stride 1)."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, **kw):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=dilation, dilation=dilation, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        identity = x
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        if self.downsample is not None:
            identity = self.downsample(x)
        return self.relu(out + identity)


class _Metric(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()

    def add_state(self, name, default, dist_reduce_fx=None):
        if isinstance(default, torch.Tensor):
            self.register_buffer(name, default)
        else:
            setattr(self, name, default)


def bev_pool(feats, coords, B, D, H, W):
    """Semantics of the external ``mmdet3d.ops.bev_pool`` (BEVFusion fork): coords columns are
    (x, y, z, b); result [B, C, D(z), H(x), W(y)] holds, per voxel, the sum of the rows of
    ``feats`` that fall in it (empty voxels are 0)."""
    B, D, H, W = int(B), int(D), int(H), int(W)
    flat = ((coords[:, 3] * D + coords[:, 2]) * H + coords[:, 0]) * W + coords[:, 1]
    out = feats.new_zeros(B * D * H * W, feats.shape[1]).index_add_(0, flat, feats)
    return out.view(B, D, H, W, -1).permute(0, 4, 1, 2, 3).contiguous()


class _Dummy:
    """Inert stand-in returned for any attribute of an auto-stubbed module."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return _Dummy()

    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Dummy()

    def __mro_entries__(self, bases):
        return (object,)

    def __iter__(self):
        return iter(())


class _AutoStub(types.ModuleType):
    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Dummy()


_AUTO_ROOTS = {"prettytable", "open3d", "mcubes", "trimesh", "matplotlib", "torchsparse", "nuscenes",
               "pyquaternion", "fvcore", "pycocotools", "IPython", "mmseg", "mmcv", "mmdet", "mmdet3d",
               "torchmetrics", "timm", "spconv", "mayavi", "seaborn"}


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        root = fullname.split(".")[0]
        if root in _AUTO_ROOTS and fullname not in sys.modules:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _AutoStub(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_installed = False
_saved_modules = {}


def _mk(name, **attrs):
    m = _AutoStub(name)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install():
    """Install the stubs (idempotent)."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")
    _installed = True

    # 1. namespace-only packages so package __init__s never execute
    ns = {
        "projects": _P,
        "projects.mmdet3d_plugin": _PLUG,
        "projects.mmdet3d_plugin.utils": _PLUG + "/utils",
        "projects.mmdet3d_plugin.occupancy": _PLUG + "/occupancy",
        "projects.mmdet3d_plugin.occupancy.image2bev": _PLUG + "/occupancy/image2bev",
        "projects.mmdet3d_plugin.occupancy.backbones": _PLUG + "/occupancy/backbones",
        "projects.mmdet3d_plugin.occupancy.necks": _PLUG + "/occupancy/necks",
        "projects.mmdet3d_plugin.occupancy.dense_heads": _PLUG + "/occupancy/dense_heads",
        "projects.mmdet3d_plugin.occupancy.detectors": _PLUG + "/occupancy/detectors",
        "projects.mmdet3d_plugin.core": _PLUG + "/core",
        "projects.mmdet3d_plugin.core.bbox": _PLUG + "/core/bbox",
        "projects.mmdet3d_plugin.models": _PLUG + "/models",
        "projects.mmdet3d_plugin.models.utils": _PLUG + "/models/utils",
    }
    for name, path in ns.items():
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    u = sys.modules["projects.mmdet3d_plugin.utils"]
    for n in ("cm_to_ious", "query_points_from_voxels", "per_class_iu", "fast_hist_crop",
              "SoftDiceLossWithProb", "PositionAwareLoss", "format_results"):
        setattr(u, n, _Dummy())
    mu = sys.modules["projects.mmdet3d_plugin.models.utils"]
    mu.GridMask = _Dummy
    sys.modules["projects.mmdet3d_plugin.core.bbox"].util = _mk(
        "projects.mmdet3d_plugin.core.bbox.util", normalize_bbox=_Dummy())

    # 2. explicit stubs
    necks, backbones, heads, detectors = (_Registry(n) for n in ("neck", "backbone", "head", "detector"))
    _mk("mmcv")
    _mk("mmcv.runner", BaseModule=_BaseModule, force_fp32=_identity_decorator_factory,
        auto_fp16=_identity_decorator_factory, get_dist_info=lambda: (0, 1), Sequential=_Sequential)
    _mk("mmcv.cnn", build_norm_layer=_build_norm_layer, build_conv_layer=_build_conv_layer,
        build_upsample_layer=_build_upsample_layer, ConvModule=_ConvModule)
    _mk("mmcv.cnn.bricks", ConvModule=_ConvModule, DropPath=_DropPath)
    _mk("mmdet")
    _mk("mmdet.models", NECKS=necks, HEADS=heads, DETECTORS=detectors, BACKBONES=backbones)
    _mk("mmdet.models.utils", SELayer=_SELayer, make_divisible=_make_divisible)
    _mk("mmdet.models.backbones")
    _mk("mmdet.models.backbones.resnet", BasicBlock=_BasicBlock2d)
    _mk("mmdet3d")
    _mk("mmdet3d.models")
    _mk("mmdet3d.models.builder", NECKS=necks, BACKBONES=backbones, HEADS=heads, DETECTORS=detectors)
    _mk("mmdet3d.ops")
    _mk("mmdet3d.ops.bev_pool", bev_pool=bev_pool)
    _mk("mmdet3d.ops.voxel_pooling", voxel_pooling=_Dummy())
    _mk("torchmetrics")
    _mk("torchmetrics.metric", Metric=_Metric)

    # 3. everything else: inert
    sys.meta_path.append(_Finder())


_REGISTRIES = None


def _imp(name):
    install()
    return importlib.import_module(name)


def vt_module():
    """The reference ViewTransformerLSSVoxel module (unmodified)."""
    return _imp("projects.mmdet3d_plugin.occupancy.image2bev.ViewTransformerLSSVoxel")


def att_module():
    return _imp("projects.mmdet3d_plugin.occupancy.image2bev.attention")


def vtb_module():
    return _imp("projects.mmdet3d_plugin.occupancy.image2bev.ViewTransformerLSSBEVDepth")


def resnet3d_module():
    return _imp("projects.mmdet3d_plugin.occupancy.backbones.resnet3d")


def neck_module():
    return _imp("projects.mmdet3d_plugin.occupancy.necks.second_fpn_3d")


def efficientnet_module():
    """The reference's CustomEfficientNet file (unmodified), on the restated mmcv / mmdet bricks above."""
    return _imp("projects.mmdet3d_plugin.occupancy.backbones.efficientnet")


def occhead_module():
    return _imp("projects.mmdet3d_plugin.occupancy.dense_heads.occhead")


@contextlib.contextmanager
def cpu_arange():
    """``warp`` hard-codes device='cuda' in torch.arange (VT:140,144); redirect to CPU while the
    reference forward runs."""
    orig = torch.arange

    def patched(*a, **k):
        if str(k.get("device", "")) == "cuda" and not torch.cuda.is_available():
            k = dict(k)
            k["device"] = "cpu"
        return orig(*a, **k)

    torch.arange = patched
    try:
        yield
    finally:
        torch.arange = orig
