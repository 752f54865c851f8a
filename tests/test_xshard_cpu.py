"""X-slab sharded mode on CPU (gloo, world_size 2): the orchestration of stereoscene_b200/xshard.py -- slab plan, CSR slab
of the splat index, halo exchange, restricted GroupNorm sums + all-reduce, stride-2 slab arithmetic, replicated-edge
resize -- driven with the torch re-statement of the kernels (tests/cpu_kernels.py) and compared with the UNSHARDED oracle
(oracle/restatement.py) on the same weights and inputs."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))


def _setup(rank, world, port):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _problem():
    """Tiny geometry (16x16x4 LSS grid): seeded model, depth distribution, context features and geometry."""
    import cpu_kernels as K
    from oracle import restatement as O
    from util import build_model, cpu_state_dict, golden_tiny, tiny_inputs
    cfg, gold = golden_tiny()
    model, mc = build_model("tiny", cfg["seed"])
    sd = cpu_state_dict(model)
    dp = torch.from_numpy(gold["depth_prob"])[:1].double()
    B, D, H, W = dp.shape
    feat = torch.randn(1, H, W, 128, generator=torch.Generator().manual_seed(5)).double()
    geom = torch.from_numpy(gold["geom"])[:1]
    gc = cfg["grid_config"]
    dx, bx, nx = O.gen_dx_bx(gc["xbound"], gc["ybound"], gc["zbound"])
    index = K.splat_build_index(geom, dx.tolist(), bx.tolist(), nx.tolist())
    return cfg, model, sd, dp, feat, geom, (dx, bx, nx), index


def _worker(rank, world, port, q):
    _setup(rank, world, port)
    import cpu_kernels as K
    from oracle import restatement as O
    from stereoscene_b200 import xshard
    cfg, model, sd, dp, feat, geom, (dx, bx, nx), index = _problem()
    plan = xshard.SlabPlan(index.nx, world, rank)
    path = xshard.XShardedVoxelPath(model, plan, kernels=K)
    with torch.no_grad():
        out = path.run(dp, feat, index, cfg["occ_size"], want_labels=True)
        # the unsharded oracle on the same inputs (float64 copies of the same weights)
        sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
        bev = O.lift_splat(dp, feat.permute(0, 3, 1, 2), geom, dx, bx, nx)
        levels = O.resnet3d(sd64, "img_bev_encoder_backbone", bev)
        neck = O.second_fpn3d(sd64, "img_bev_encoder_neck", levels)
        logits = O.occ_head(sd64, "pts_bbox_head", neck)
        up = O.upsample_logits(logits, cfg["occ_size"])
    want = up.permute(0, 2, 3, 4, 1)[:, 2 * plan.x0: 2 * plan.x1]
    err = float((out["logits"] - want).abs().max() / want.abs().max())
    low = logits.permute(0, 2, 3, 4, 1)[:, plan.x0: plan.x1]
    err_low = float((out["logits_lowres"] - low).abs().max() / low.abs().max())
    lab_ok = bool(torch.equal(out["labels"].long(), want.argmax(-1)))
    q.put((rank, err, err_low, lab_ok, dict(path.collectives), tuple(out["logits"].shape)))
    dist.barrier()
    dist.destroy_process_group()


def _halo_worker(rank, world, port, q):
    """sharded splat + ONE halo'd 3x3x3 convolution + GroupNorm against the unsharded evaluation (the minimal case)."""
    _setup(rank, world, port)
    import cpu_kernels as K
    from oracle import restatement as O
    from stereoscene_b200 import xshard
    cfg, model, sd, dp, feat, geom, (dx, bx, nx), index = _problem()
    plan = xshard.SlabPlan(index.nx, world, rank)
    path = xshard.XShardedVoxelPath(model, plan, kernels=K)
    blk = model.img_bev_encoder_backbone.layers[0][0]
    xs, Y, Z = plan.xs, index.ny, index.nz
    with torch.no_grad():
        buf = path.halo_buf(feat, xs, Y, Z, 128)
        K.lift_splat(dp, feat, K.splat_index_slab(index, plan.x0, plan.x1), out=buf[:, 1:xs + 1])
        path.exchange(buf, xs)
        y, st = path.conv3(K.Vol(buf), blk.conv1, xs)
        v = path.gn(y, st, blk.bn1, K.SS_ACT_RELU, count=plan.nx * Y * Z)
        got = K._value(v)[:, :, 1:xs + 1]
        bev = O.lift_splat(dp, feat.permute(0, 3, 1, 2), geom, dx, bx, nx)
        full = torch.relu(torch.nn.functional.group_norm(
            torch.nn.functional.conv3d(bev, blk.conv1.weight.double(), None, padding=1), blk.bn1.num_groups,
            blk.bn1.weight.double(), blk.bn1.bias.double(), blk.bn1.eps))
    want = full[:, :, plan.x0:plan.x1]
    splat_err = float((buf[:, 1:xs + 1].permute(0, 4, 1, 2, 3) - bev[:, :, plan.x0:plan.x1]).abs().max())
    q.put((rank, float((got - want).abs().max() / want.abs().max()), splat_err))
    dist.barrier()
    dist.destroy_process_group()


def _spawn(fn, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() * 7 + id(fn)) % 2000
    procs = [ctx.Process(target=fn, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


def test_sharded_splat_conv_groupnorm_two_ranks():
    for rank, err, splat_err in _spawn(_halo_worker):
        assert splat_err < 1e-12, (rank, splat_err)
        assert err < 1e-6, (rank, err)


def test_sharded_voxel_path_two_ranks_equals_unsharded_oracle():
    out = _spawn(_worker)
    for rank, err, err_low, lab_ok, coll, shape in out:
        assert shape == (1, 16, 32, 8, 20), shape             # 2 * xs planes of the 32x32x8 occupancy grid
        assert err_low < 1e-6 and err < 1e-6, (rank, err_low, err)
        assert lab_ok
        # 2 exchanges per BasicBlock (6 blocks) + input_proj + neck + logits; one all-reduce per GroupNorm layer
        assert coll["halo_exchanges"] == 15 and coll["stat_allreduces"] == 1 + 12 + 2 + 3 + 1, coll


def test_slab_plan_rejects_uneven_splits():
    sys.path.insert(0, ROOT)
    from stereoscene_b200 import xshard
    p = xshard.SlabPlan(128, 8, 3)
    assert (p.xs, p.x0, p.x1) == (16, 48, 64)
    with pytest.raises(ValueError):
        xshard.SlabPlan(128, 64, 0)
