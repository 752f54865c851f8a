"""X-slab sharded mode of the volumetric path (BASELINE.json configs[3]): the voxel grid is split on
the scene-X axis across the ranks of one box, with ONE all-gather of ``depth_prob || img_feat`` at the MIE boundary
(after ViewTransformerLSSVoxel.py:508, before the lift (x) splat of :517-523).

What shards how (one process per GPU, torch.distributed over NCCL / NVLink):

  * frustum-space stages (stereo, depth_net, MIE) are not indexed by scene-X: every sample of the global batch is computed
    by ONE rank (sample s on rank s % R), then ``depth_prob || img_feat`` (7.4 MB per sample in fp32) is all-gathered;
  * lift (x) splat: the splat index is sorted by voxel rank with x slowest, so rank r fills voxels
    ``[r X/R, (r+1) X/R)`` from a contiguous range of the CSR offsets (``ops.splat_index_slab``) -- no communication;
  * the 3-D encoder / neck / head run on slabs ``[1, n + 2, Y, Z, C]`` (one halo plane per side).  A 3x3x3 stride-1 layer
    is computed on all n + 2 planes (minus the outer halo at the two ends of the grid, where the kernel's own zero padding
    of the LOGICAL input applies) with the GroupNorm sums restricted to the n interior planes
    (``ss_conv3d_desc.stats_d0/d1``), the sums are all-reduced (2 x C doubles), and the halo planes of every tensor that
    feeds a 3x3x3 layer are exchanged with the two neighbours (one plane each way, send/recv over NVLink).  Stride-2
    layers read the low halo and write n / 2 interior planes; pointwise layers (1x1x1, the neck's k = s up-convolutions)
    work on the interior only;
  * x2 trilinear + argmax: on the logits slab with one halo plane per side (replicated at the two ends of the grid, which is
    what align_corners=False clamping does), keeping the 2n interior output planes.

Results are the unsharded path's: every rank applies the same global GroupNorm statistics, so the per-rank arithmetic is
the single-GPU arithmetic on its planes.  The compute entry points are passed in as ``kernels`` (default:
``stereoscene_b200.ops``, i.e. the CUDA library); the gloo CPU tests drive the same orchestration with a small torch
re-statement of those entry points to check the slab bookkeeping, the halo exchange and the statistics reduction.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist

SS_ACT_NONE, SS_ACT_RELU = 0, 1
# STEREOSCENE_B200_XSHARD_TIMING=1: synchronise around every collective and accumulate its device time per category
# (diagnostic only: it serialises the step)
_TIMING = os.environ.get("STEREOSCENE_B200_XSHARD_TIMING", "0") == "1"
TIMES = {"halo_ms": 0.0, "allreduce_ms": 0.0, "allgather_ms": 0.0, "frustum_ms": 0.0}


class _timed:
    def __init__(self, key):
        self.key = key

    def __enter__(self):
        if _TIMING and torch.cuda.is_available():
            torch.cuda.synchronize()
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _TIMING and torch.cuda.is_available():
            self.e1.record()
            torch.cuda.synchronize()
            TIMES[self.key] += self.e0.elapsed_time(self.e1)
        return False


class SlabPlan:
    """Rank r of R owns LSS-grid planes [x0, x1) of nx; the deepest encoder level (stride 4) must still split evenly."""

    def __init__(self, nx: int, world: int, rank: int, levels: int = 3):
        if nx % (world << (levels - 1)) != 0:
            raise ValueError(f"X extent {nx} does not split into {world} slabs at {levels} encoder levels")
        self.nx, self.world, self.rank = nx, world, rank
        self.xs = nx // world
        self.x0, self.x1 = rank * self.xs, (rank + 1) * self.xs


def exchange_halo(buf: torch.Tensor, n: int, plan: SlabPlan, group=None, edge: str = "zero"):
    """buf [1, n + 2, ...]: planes 1..n are this rank's.  Afterwards plane 0 holds the lower neighbour's plane n and plane
    n + 1 the upper neighbour's plane 1; at the two ends of the grid the halo is zero (= the convolution's padding) or a
    copy of the edge plane (``edge="replicate"``: the clamp of align_corners=False resampling)."""
    ops_ = []
    lo, hi = plan.rank - 1, plan.rank + 1
    if lo >= 0:
        peer = dist.get_global_rank(group, lo) if group is not None else lo
        ops_ += [dist.P2POp(dist.isend, buf[:, 1], peer, group), dist.P2POp(dist.irecv, buf[:, 0], peer, group)]
    elif edge == "replicate":
        buf[:, 0].copy_(buf[:, 1])
    else:
        buf[:, 0].zero_()
    if hi < plan.world:
        peer = dist.get_global_rank(group, hi) if group is not None else hi
        ops_ += [dist.P2POp(dist.isend, buf[:, n], peer, group), dist.P2POp(dist.irecv, buf[:, n + 1], peer, group)]
    elif edge == "replicate":
        buf[:, n + 1].copy_(buf[:, n])
    else:
        buf[:, n + 1].zero_()
    if ops_:
        for req in dist.batch_isend_irecv(ops_):
            req.wait()
    return buf


class XShardedVoxelPath:
    """Lift (x) splat, CustomResNet3D, SECONDFPN3D, OccHead and the x2 resize of ONE sample on this rank's X-slab."""

    def __init__(self, model, plan: SlabPlan, group=None, kernels=None, pool=None):
        """``pool``: a stereoscene_b200.peer.PeerPool -> the halo exchange and the statistics all-reduce run as peer-memory
        kernels over NVLink (no NCCL call per collective); None -> torch.distributed send/recv and all_reduce."""
        if kernels is None:
            from . import ops as kernels          # the CUDA library (fails loudly without it)
        self.K, self.model, self.plan, self.group, self.pool = kernels, model, plan, group, pool
        self.collectives = {"halo_exchanges": 0, "halo_bytes": 0, "stat_allreduces": 0}
        self._pool_off = {}

    # ---- communication ----------------------------------------------------------------------------------------------
    def exchange(self, buf, n, edge="zero"):
        self.collectives["halo_exchanges"] += 1
        self.collectives["halo_bytes"] += 2 * buf[:, 0].numel() * buf.element_size()
        with _timed("halo_ms"):
            if self.pool is not None:
                self.pool.halo_push(buf, self._pool_off[buf.data_ptr()], n, edge == "replicate")
                return buf
            return exchange_halo(buf, n, self.plan, self.group, edge)

    def gn(self, y, stats, gnmod, act, count, scale_out=None, shift_out=None):
        """GroupNorm of a slab from GLOBAL statistics: sum the slab sums over the ranks, then finalise."""
        if self.plan.world > 1:
            with _timed("allreduce_ms"):
                if self.pool is not None:
                    self.pool.allreduce(stats)
                else:
                    dist.all_reduce(stats, group=self.group)
            self.collectives["stat_allreduces"] += 1
        return self.K.gn_pending(y, stats, gnmod, act, scale_out, shift_out, count=count)

    def halo_buf(self, like: torch.Tensor, n: int, Y: int, Z: int, C: int) -> torch.Tensor:
        """[1, n + 2, Y, Z, C] buffer; from the peer pool (same offset on every rank) when one is attached."""
        if self.pool is not None:
            t, off = self.pool.tensor((1, n + 2, Y, Z, C), like.dtype)
            self._pool_off[t.data_ptr()] = off
            return t
        return torch.empty((1, n + 2, Y, Z, C), dtype=like.dtype, device=like.device)

    # ---- 3x3x3 layers on a slab ---------------------------------------------------------------------------------------------
    # A convolution pads the LOGICAL input (after the pending affine / activation) with zeros.  A raw zero in a halo plane is
    # not a logical zero once an affine is pending, so at the two ends of the grid the outer halo plane is left out of the
    # view the kernel sees and the kernel's own padding (TMA out-of-bounds fill, untouched by the fix-up warps) supplies it.
    def _trim(self):
        return (1 if self.plan.rank == 0 else 0), (1 if self.plan.rank == self.plan.world - 1 else 0)

    def conv3(self, x, module, n: int):
        """Stride-1 3x3x3 layer on x = Vol over [1, n + 2, Y, Z, C] with valid inner halos -> (raw output buffer
        [1, n + 2, Y, Z, Cout] whose interior planes are exact, sums over the interior planes)."""
        K = self.K
        lo, hi = self._trim()
        y = self.halo_buf(x.data, n, x.data.shape[2], x.data.shape[3], module.out_channels)
        xin = K.Vol(x.data[:, lo:n + 2 - hi], x.scale, x.shift, x.act)
        _, st = K.conv(xin, module, out=y[:, lo:n + 2 - hi], want_stats=True, stats_planes=(1 - lo, n + 1 - lo))
        return y, st

    def conv3_stride(self, x, module, n: int, s: int):
        """Stride-s (s = 2) 3x3x3 layer: output plane j of the slab reads local planes 2j, 2j+1, 2j+2 of the halo'd input
        (local 0 = the low halo).  Inner ranks: no padding on the slab axis.  Rank 0: the low halo is the grid's padding, so
        the view starts at local plane 1 with the layer's own padding; that yields one surplus output plane, which lands in
        the high halo slot of the result (overwritten by the next exchange) and is excluded from the sums."""
        K = self.K
        if s != 2 or n % 2:
            raise NotImplementedError("sharded strided layers: stride 2 on an even number of planes")
        m = n // 2
        y = self.halo_buf(x.data, m, x.data.shape[2] // s, x.data.shape[3] // s, module.out_channels)
        if self.plan.rank == 0:
            xin = K.Vol(x.data[:, 1:], x.scale, x.shift, x.act)
            _, st = K.conv(xin, module, out=y[:, 1:m + 2], want_stats=True, stats_planes=(0, m))
        else:
            _, st = K.conv(x, module, out=y[:, 1:m + 1], want_stats=True, pad=(0, 1, 1))
        return y, st, m

    # ---- one BasicBlock on a slab (resnet3d.py:35-65) -----------------------------------------------------------------
    def block(self, blk, x, n: int, count_in: int):
        """x: Vol over [1, n + 2, Y, Z, C] with valid halo planes -> (Vol of the block output with valid halos, planes, count)."""
        K = self.K
        if blk.stride == 1:
            m, count = n, count_in
            y1, st = self.conv3(x, blk.conv1, n)
            res = x
        else:
            y1, st, m = self.conv3_stride(x, blk.conv1, n, blk.stride)
            count = count_in // (blk.stride ** 3)
        v1 = self.gn(y1, st, blk.bn1, SS_ACT_RELU, count)
        self.exchange(y1, m)
        y2, st = self.conv3(v1, blk.conv2, m)
        v2 = self.gn(y2, st, blk.bn2, SS_ACT_NONE, count)
        if blk.downsample is not None:
            rbuf = self.halo_buf(x.data, m, y2.shape[2], y2.shape[3], blk.downsample[0].out_channels)
            xin = K.Vol(x.data[:, 1:n + 1], x.scale, x.shift, x.act)          # 1x1x1 stride-s: interior planes only
            _, st = K.conv(xin, blk.downsample[0], out=rbuf[:, 1:m + 1], want_stats=True)
            res = self.gn(rbuf, st, blk.downsample[1], SS_ACT_NONE, count)
        out = self.halo_buf(y2, m, y2.shape[2], y2.shape[3], y2.shape[4])
        K.join(v2, res, out_act=SS_ACT_RELU, out=out)
        self.exchange(out, m)
        return K.Vol(out), m, count

    # ---- the voxel-space path of one sample -----------------------------------------------------------------------------
    def run(self, depth_prob, img_feat, index, occ_size, want_labels: bool = True):
        """depth_prob [1,D,H,W], img_feat [1,H,W,C], index: the full splat index of the sample ->
        dict(logits = [1, 2 xs, Y2, Z2, classes] channels-last slab of the upsampled logits, labels = uint8 slab)."""
        K, plan, model = self.K, self.plan, self.model
        enc, neck, head = model.img_bev_encoder_backbone, model.img_bev_encoder_neck, model.pts_bbox_head
        xs, Y, Z = plan.xs, index.ny, index.nz
        if index.nx != plan.nx:
            raise ValueError("splat index and slab plan disagree on the X extent")
        if tuple(occ_size) != (2 * plan.nx, 2 * Y, 2 * Z):
            raise NotImplementedError("the sharded resize is the x2 case of stereoscene.py (lss_downsample = 2)")
        count = plan.nx * Y * Z
        # (ii) lift (x) splat of this rank's slab: a contiguous range of the CSR offsets
        bev = torch.empty((1, xs, Y, Z, img_feat.shape[-1]), dtype=img_feat.dtype, device=img_feat.device)
        K.lift_splat(depth_prob, img_feat, K.splat_index_slab(index, plan.x0, plan.x1), out=bev)
        # (iv-enc) input_proj is pointwise: interior only
        y0 = self.halo_buf(bev, xs, Y, Z, enc.input_proj[0].out_channels)
        _, st = K.conv(K.Vol(bev), enc.input_proj[0], out=y0[:, 1:xs + 1], want_stats=True)
        v = self.gn(y0, st, enc.input_proj[1], SS_ACT_RELU, count)
        self.exchange(y0, xs)
        n, levels = xs, []
        for i, layer in enumerate(enc.layers):
            for blk in layer:
                v, n, count = self.block(blk, v, n, count)
            if i in enc.out_indices:
                levels.append((v.data, n))
        # neck: k = s up-convolutions are pointwise in the input voxel -> interior planes, written into channel slices
        ctot = sum(neck.out_channels)
        nb = self.halo_buf(bev, xs, Y, Z, ctot)
        ss = torch.empty((2, 1, ctot), dtype=bev.dtype, device=bev.device)
        count0, off = plan.nx * Y * Z, 0
        for (t, ni), blk, co in zip(levels, neck.deblocks, neck.out_channels):
            y, st = K.conv(K.Vol(t[:, 1:ni + 1]), blk[0], out=nb[:, 1:xs + 1, :, :, off:off + co], want_stats=True)
            self.gn(y, st, blk[1], SS_ACT_RELU, count0, ss[0][:, off:off + co], ss[1][:, off:off + co])
            off += co
        self.exchange(nb, xs)
        # head: 3x3x3 on the halo'd slab, classifier on the interior
        seq = head.occ_convs[0]
        y, st = self.conv3(K.Vol(nb, ss[0], ss[1], SS_ACT_RELU), seq[0], xs)
        h = self.gn(y, st, seq[1], SS_ACT_RELU, count0)
        lb = self.halo_buf(bev, xs, Y, Z, seq[3].out_channels)
        K.conv(K.Vol(h.data[:, 1:xs + 1], h.scale, h.shift, h.act), seq[3], out=lb[:, 1:xs + 1])
        self.exchange(lb, xs, edge="replicate")
        up, labels = K.trilinear(lb, (2 * (xs + 2), 2 * Y, 2 * Z), want_labels=want_labels)
        return {"logits": up[:, 2:2 * xs + 2], "labels": labels[:, 2:2 * xs + 2] if labels is not None else None,
                "logits_lowres": lb[:, 1:xs + 1]}


class XShardedPipeline:
    """The whole volumetric forward of a global batch in the sharded layout: frustum stages of sample s on rank s % R,
    one all-gather of ``depth_prob || img_feat`` at the MIE boundary, then every rank runs its X-slab of EVERY sample."""

    def __init__(self, model, world: int, rank: int, group=None, kernels=None, peer_memory: bool = False, pool_bytes: int = 0):
        """``peer_memory``: run the voxel stack's collectives as NVLink peer-memory kernels (stereoscene_b200.peer) instead of
        torch.distributed calls; ``pool_bytes``: size of the per-rank pool (default: sized for one sample's slab buffers)."""
        vt = model.img_view_transformer
        nx = [int(round(float(v))) for v in vt.nx.detach().cpu()]
        self.model, self.world, self.rank, self.group = model, world, rank, group
        self.plan = SlabPlan(nx[0], world, rank)
        pool = None
        if peer_memory:
            from .peer import PeerPool
            if not pool_bytes:      # ~16 slab buffers of the widest level + the 384-channel neck buffer, with head-room
                per_plane = nx[1] * nx[2] * 4
                pool_bytes = int((self.plan.xs + 2) * per_plane * (16 * 128 + 384 + 64) * 1.25) + (64 << 20)
            pool = PeerPool(pool_bytes, world, rank, vt.frustum.device, group)
        self.pool = pool
        self.path = XShardedVoxelPath(model, self.plan, group, kernels, pool)
        self.gathered_bytes = 0
        self.split_frustum = True            # B = 1 on >= 2 ranks: stereo branch and depth_net on two ranks at once
        self.use_graph = False               # peer pool only: replay the voxel-space path as one CUDA graph
        self._graph = None
        self._mlp_cache = {}

    def frustum(self, x_left, x_right, left, right, calib):
        """(depth_prob [b,D,H,W], img_feat [b,H,W,C]) of this rank's own samples."""
        vt = self.model.img_view_transformer
        keys = ("rots", "trans", "intrins", "post_rots", "post_trans", "bda")
        ml, mr = self._mlp(left), self._mlp(right)
        inp = [x_left] + [left[k] for k in keys] + [ml] + [x_right] + [right[k] for k in keys] + [mr] + [calib, None, None]
        dp, feat = vt.frustum_forward(inp)[:2]
        return dp, feat

    def frustum_split(self, x_left, x_right, left, right, calib, owner: int, helper: int):
        """B = 1 latency mode on >= 2 ranks: the stereo branch and depth_net are independent until the MIE block, so the
        owner runs (i) while the helper runs (N1) on the left feature map it receives from the owner, and sends back the lss
        distribution and the context features (7.4 MB).  Returns (depth_prob, img_feat) on the owner, None elsewhere."""
        vt = self.model.img_view_transformer
        keys = ("rots", "trans", "intrins", "post_rots", "post_trans", "bda")
        dev = vt.frustum.device
        fH, fW = vt.frustum.shape[1], vt.frustum.shape[2]
        if self.rank == owner:
            ml, mr = self._mlp(left), self._mlp(right)
            fl = x_left.squeeze(1).contiguous()
            send = dist.isend(fl, self._peer(helper), group=self.group)
            self.path.K.arena(dev).reset()
            stereo = vt.stereo_branch(fl, x_right.squeeze(1), ml, mr, calib)
            lss = torch.empty((1, vt.D, fH, fW), dtype=torch.float32, device=dev)
            feat = torch.empty((1, fH, fW, vt.numC_Trans), dtype=torch.float32, device=dev)
            r1 = dist.irecv(lss, self._peer(helper), group=self.group)
            r2 = dist.irecv(feat, self._peer(helper), group=self.group)
            send.wait(); r1.wait(); r2.wait()
            return vt.mie_branch(stereo, lss), feat
        if self.rank == helper:
            ml = self._mlp(left)
            fl = torch.empty((1, vt.numC_input, fH, fW), dtype=torch.float32, device=dev)
            dist.recv(fl, self._peer(owner), group=self.group)
            self.path.K.arena(dev).reset()
            lss, feat = vt.depth_branch(fl, ml)
            s1 = dist.isend(lss, self._peer(owner), group=self.group)
            s2 = dist.isend(feat, self._peer(owner), group=self.group)
            s1.wait(); s2.wait()
        return None

    def _mlp(self, cal: dict) -> torch.Tensor:
        """Calibration-only MLP input vector, cached per calibration (identity of the dict's tensors)."""
        vt = self.model.img_view_transformer
        keys = ("rots", "trans", "intrins", "post_rots", "post_trans", "bda")
        key = tuple((cal[k].data_ptr(), cal[k]._version) for k in keys)
        hit = self._mlp_cache.get(key)
        if hit is None:
            with torch.no_grad():
                hit = (vt.get_mlp_input(*[cal[k] for k in keys]).contiguous(), [cal[k] for k in keys])
            self._mlp_cache[key] = hit
        return hit[0]

    def _peer(self, r: int) -> int:
        return dist.get_global_rank(self.group, r) if self.group is not None else r

    def gather(self, dp: torch.Tensor, feat: torch.Tensor):
        """The ONE data-path collective: all-gather of depth_prob || img_feat.  Every rank contributes the same number b of
        samples (ranks without work contribute zeros); returns lists of R*b [1,...] tensors, sample g = r*b + i."""
        b = dp.shape[0]
        flat = torch.cat([dp.reshape(b, -1), feat.reshape(b, -1)], dim=1).contiguous()
        if self.world > 1:
            out = torch.empty((self.world,) + tuple(flat.shape), dtype=flat.dtype, device=flat.device)
            with _timed("allgather_ms"):
                dist.all_gather_into_tensor(out, flat, group=self.group)
            self.gathered_bytes += out.numel() * out.element_size()
            flat = out.reshape(self.world * b, -1)
        ndp = dp[0].numel()
        dps = [flat[g:g + 1, :ndp].reshape((1,) + tuple(dp.shape[1:])).contiguous() for g in range(flat.shape[0])]
        fts = [flat[g:g + 1, ndp:].reshape((1,) + tuple(feat.shape[1:])).contiguous() for g in range(flat.shape[0])]
        return dps, fts

    def forward(self, x_left, x_right, left, right, calib, occ_size, counts: Optional[List[int]] = None,
                want_labels=True) -> List[dict]:
        """x_left / x_right: this rank's own samples [b_r,1,C,fH,fW] (None if it owns none); ``counts[r]`` = samples owned
        by rank r (default: one each).  The calibration (hence the splat index) is the sequence's, shared by all samples.
        Returns one dict per GLOBAL sample (rank-major order) holding this rank's X-slab of its logits / labels."""
        vt = self.model.img_view_transformer
        counts = [1] * self.world if counts is None else list(counts)
        b, mine = max(counts), counts[self.rank]
        dev = vt.frustum.device
        fH, fW = vt.frustum.shape[1], vt.frustum.shape[2]
        dp = torch.zeros((b, vt.D, fH, fW), dtype=torch.float32, device=dev)
        feat = torch.zeros((b, fH, fW, vt.numC_Trans), dtype=torch.float32, device=dev)
        split = self.split_frustum and self.world > 1 and sum(counts) == 1
        if split:
            owner = counts.index(1)
            got = self.frustum_split(x_left, x_right, {k: v[:1] for k, v in left.items()}, {k: v[:1] for k, v in right.items()},
                                     calib[:1], owner, (owner + 1) % self.world)
            if got is not None:
                dp[:1].copy_(got[0])
                feat[:1].copy_(got[1])
            elif self.rank != (owner + 1) % self.world and hasattr(self.path.K, "arena"):
                self.path.K.arena(dev).reset()
        elif mine == 0 and hasattr(self.path.K, "arena"):
            self.path.K.arena(dev).reset()               # frustum_forward does this on the ranks that own a sample
        if mine > 0 and not split:
            cut = lambda d: {k: v[:mine] for k, v in d.items()}      # noqa: E731
            with _timed("frustum_ms"):
                d_, f_ = self.frustum(x_left[:mine], x_right[:mine], cut(left), cut(right), calib[:mine])
            dp[:mine].copy_(d_)
            feat[:mine].copy_(f_)
        dps, fts = self.gather(dp, feat)
        one = {k: v[:1] for k, v in left.items()}
        index = vt.splat_index(*[one[k] for k in ("rots", "trans", "intrins", "post_rots", "post_trans", "bda")])
        outs = []
        for r in range(self.world if self.world > 1 else 1):
            for i_ in range(counts[r]):
                g = r * b + i_
                outs.append(self._voxel(dps[g], fts[g], index, occ_size, want_labels))
        return outs

    def _voxel_eager(self, dp, ft, index, occ_size, want_labels):
        K = self.path.K
        with K.math_scope("voxel"):
            if self.pool is not None:
                self.pool.begin()            # one pool generation per sample: rewind the bump allocators, advance the epoch
                self.path._pool_off = {}
            out = self.path.run(dp, ft, index, occ_size, want_labels)
            if hasattr(K, "arena"):
                # leave the GroupNorm sum blocks of this stream zeroed: inside a CUDA graph this IS the reset of the next replay
                # (the capture stream has its own arena, which nothing else rewinds)
                K.arena(dp.device).reset()
            return out

    def _voxel(self, dp, ft, index, occ_size, want_labels):
        """The voxel-space path of one sample: eager, or -- with the peer pool, whose collectives are plain kernels -- as ONE
        CUDA graph replayed from static input buffers (every rank replays its own capture of the same sequence)."""
        if not (self.use_graph and self.pool is not None):
            return self._voxel_eager(dp, ft, index, occ_size, want_labels)
        key = (tuple(occ_size), bool(want_labels))
        if self._graph is None or self._graph[0] != key:
            self._calls = getattr(self, "_calls", 0) + 1
            out = self._voxel_eager(dp, ft, index, occ_size, want_labels)
            if self._calls < 3:                       # warm the caches (packed weights, splat index) before capturing
                return out
            sdp, sft = dp.clone(), ft.clone()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                gout = self._voxel_eager(sdp, sft, index, occ_size, want_labels)
            self._graph = (key, g, sdp, sft, gout, index)
            return out
        _, g, sdp, sft, gout, _ = self._graph
        sdp.copy_(dp)
        sft.copy_(ft)
        g.replay()
        return gout
