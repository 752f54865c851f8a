"""Host-side logic added in round 2, on CPU: per-stage math policy scoping, the calibration-constant cache, the weight packings
of the compensated modes (TF32 hi/lo split, fp16 hi/lo split with power-of-two scale), the CSR slab of the splat index, and the
bench helpers (one workload string for both arms, parity of a forward against the committed reference fixture)."""
import math
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from stereoscene_b200 import ops  # noqa: E402


def test_math_policy_scopes_and_substages():
    try:
        ops.set_math_policy({"depthnet": ops.SS_MATH_TF32X3, "mie": ops.SS_MATH_TF32X3, "mie.ca3d": ops.SS_MATH_TF32})
        assert ops.default_math() == ops.SS_MATH_TF32
        with ops.math_scope("depthnet"):
            assert ops.default_math() == ops.SS_MATH_TF32X3
            with ops.math_scope("depthnet.aspp"):                 # sub-stage without an entry inherits its parent's
                assert ops.default_math() == ops.SS_MATH_TF32X3
        with ops.math_scope("mie"):
            with ops.math_scope("mie.ca3d"):                      # sub-stage entry wins over the parent
                assert ops.default_math() == ops.SS_MATH_TF32
            assert ops.default_math() == ops.SS_MATH_TF32X3
        with ops.math_scope("stereo"):                            # groups the policy does not name: plain TF32
            assert ops.default_math() == ops.SS_MATH_TF32
        assert ops.default_math() == ops.SS_MATH_TF32
        ops.set_default_math(ops.SS_MATH_3XTF32)                  # a uniform mode switches the policy off
        with ops.math_scope("depthnet"):
            assert ops.default_math() == ops.SS_MATH_3XTF32
        ops.set_math_policy(None)                                 # back to the product default
        assert ops.math_policy() == ops.MATH_POLICIES[ops.DEFAULT_POLICY]
        mixed = ops.MATH_POLICIES["mixed"]
        assert mixed["depthnet"] == ops.SS_MATH_TF32X3 and mixed["mie"] == ops.SS_MATH_TF32X3 and mixed["image"] == ops.SS_MATH_TF32X3
        assert mixed["stereo"] == ops.SS_MATH_F16 and mixed["voxel"] == ops.SS_MATH_F16      # fp16 operands = TF32's significand
        assert "stereo" not in ops.MATH_POLICIES["mixed_tf32stereo"] and ops.MATH_POLICIES["mixed16"] == mixed
    finally:
        ops.set_math_policy(None)


def test_cached_const_identity_and_versioning():
    a, b = torch.randn(3), torch.randn(3)
    calls = []
    f = lambda: (calls.append(1), a + b)[1]        # noqa: E731
    v1 = ops.cached_const("t", [a, b], f)
    v2 = ops.cached_const("t", [a, b], f)
    assert v1 is v2 and len(calls) == 1
    a.add_(1.0)                                     # in-place update bumps the version -> recomputed
    v3 = ops.cached_const("t", [a, b], f)
    assert len(calls) == 2 and torch.equal(v3, a + b)
    assert ops.cached_const("other tag", [a, b], f) is not v3 and len(calls) == 3


def _rna_tf32(t):
    return ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def test_tf32_split_weight_packing():
    torch.manual_seed(0)
    m = nn.Conv3d(64, 40, 3, padding=1, bias=False)
    pc = ops.PackedConv(m)
    ws = pc.weights_kmajor_split()
    assert tuple(ws.shape) == (2, 27, 40, 64)
    w = m.weight.detach().permute(2, 3, 4, 0, 1).reshape(27, 40, 64)
    hi, lo = ws[0], ws[1]
    assert torch.equal(hi, _rna_tf32(w)) and torch.equal(hi, pc.weights_kmajor())
    assert torch.equal(_rna_tf32(hi), hi) and torch.equal(_rna_tf32(lo), lo)           # both parts are TF32 numbers
    assert float(((hi.double() + lo.double()) - w.double()).abs().max() / w.abs().max()) < 2.0 ** -21


def test_fp16_split_weight_packing_and_scale():
    torch.manual_seed(1)
    m = nn.ConvTranspose3d(64, 32, 3, stride=2, padding=1, output_padding=1, bias=False)
    with torch.no_grad():
        m.weight.mul_(1e-3)                                       # tiny weights: the power-of-two scale keeps the lo halves normal
    pc = ops.PackedConv(m)
    packed, acc_scale = pc.weights_kmajor_f16()
    assert tuple(packed.shape) == (27, 32, 64) and packed.dtype == torch.float32
    assert math.log2(acc_scale) == int(math.log2(acc_scale))      # a power of two
    halves = packed.view(torch.float16).view(27, 32, 2, 2, 32)    # [tap][cout][32-channel chunk][hi|lo][32]
    hi, lo = halves[:, :, :, 0].reshape(27, 32, 64).double(), halves[:, :, :, 1].reshape(27, 32, 64).double()
    w = m.weight.detach().permute(2, 3, 4, 1, 0).reshape(27, 32, 64).double()
    assert float(hi.abs().max()) < 65504 and float(hi.abs().max()) >= 512      # max|w| / acc_scale ~ 2^10
    assert float(((hi + lo) * acc_scale - w).abs().max() / w.abs().max()) < 2.0 ** -20


def test_splat_index_slab_is_a_csr_range():
    nx, ny, nz = 8, 4, 2
    counts = torch.randint(0, 4, (nx * ny * nz,), generator=torch.Generator().manual_seed(2))
    start = torch.zeros(nx * ny * nz + 1, dtype=torch.int32)
    start[1:] = torch.cumsum(counts, 0)
    order = torch.arange(int(start[-1]) + 5, dtype=torch.int32)
    idx = ops.SplatIndex(order, start, None, nx, ny, nz, 1, order.numel())
    slab = ops.splat_index_slab(idx, 2, 6)
    assert (slab.nx, slab.ny, slab.nz, slab.B) == (4, ny, nz, 1) and slab.voxel_start.numel() == 4 * ny * nz + 1
    assert int(slab.voxel_start[0]) == int(start[2 * ny * nz]) and int(slab.voxel_start[-1]) == int(start[6 * ny * nz])
    assert slab.voxel_start.data_ptr() == start[2 * ny * nz:].data_ptr()           # a view: no copy, no re-sort


def test_bench_helpers_one_workload_string_and_golden_parity():
    import bench
    assert bench.workload_config("config2", 1) == bench.workload_config("config2", 1)
    assert "256x256x32" in bench.workload_config("config2") and "128x128x16" in bench.workload_config("config1")
    # a forward that reproduces the fixture exactly has zero error; a perturbed one reports the perturbation
    import json
    with open(os.path.join(ROOT, "tests", "golden", "golden_config1.json")) as f:
        meta = json.load(f)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "golden_config1.npz"))

    def full(key):
        shape = meta["stats"][key]["shape"]
        t = torch.zeros(shape)
        sl = tuple(slice(*x) for x in meta["samplers"][key])
        t[sl] = torch.from_numpy(gold[key])
        return t
    out = {"output_voxels": full("logits_up"), "logits_lowres": full("logits"), "depth": full("depth_prob")}
    par = bench.golden_parity("config1", out, meta["seed"])
    assert par["logits"]["max_rel"] == 0.0 and par["logits_up"]["rms_rel"] == 0.0 and par["label_agreement"] == 1.0
    out["logits_lowres"] = out["logits_lowres"] + 1e-3 * meta["stats"]["logits"]["absmax"]
    par = bench.golden_parity("config1", out, meta["seed"])
    assert abs(par["logits"]["max_rel"] - 1e-3) < 1e-6
    assert bench.golden_parity("config1", out, meta["seed"] + 1) is None          # other inputs: no fixture
