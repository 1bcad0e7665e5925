"""Drop-in for the reference's src/utils/Renderer.py (render_batch_ray / render_img / eval_points).

Same constructor and call signatures.  Depth-guided sampling (Renderer.py:163-220) is kept as
the same handful of torch elementwise expressions so that sample points are bit-identical to the
reference's and autograd carries d/d pts into rays_o / rays_d for the tracker; neighbour search,
interpolation, both decoders and compositing run in the hps_b200 kernels.
"""
import torch

from . import ops
from .common import get_rays


def _linspace32(start, end, steps, device):
    """torch.linspace(start, end, steps) in float32 for a 0-d tensor `end`, with the CUDA kernel's arithmetic
    (step = (end - start) / (steps - 1); first half start + step*i, second half end - step*(steps-1-i))."""
    idx = torch.arange(steps, device=device, dtype=torch.float32)
    step = (end - start) / (steps - 1)
    return torch.where(idx < steps // 2, start + step * idx, end - step * (steps - 1 - idx))


class Renderer(object):
    def __init__(self, cfg, args=None, slam=None, points_batch_size=500000, ray_batch_size=3000):
        self.ray_batch_size = ray_batch_size
        self.points_batch_size = points_batch_size
        r = cfg["rendering"]
        self.N_samples, self.N_surface = r["N_samples"], r["N_surface"]
        self.near_end_surface, self.far_end_surface = r["near_end_surface"], r["far_end_surface"]
        self.sample_near_pcl = r["sample_near_pcl"]
        self.skip_zero_depth_pixel = r["skip_zero_depth_pixel"]
        self.near_end = r["near_end"]
        self.fix_interval = cfg["pointcloud"]["fix_interval_when_add_along_ray"]
        self.occupancy = cfg["occupancy"]
        self.use_dynamic_radius = cfg["use_dynamic_radius"]
        self.use_bound = cfg["use_bound"]
        self.crop_edge = 0 if cfg["cam"]["crop_edge"] is None else cfg["cam"]["crop_edge"]
        self.sigmoid_coefficient = r.get("sigmoid_coef_mapper", 0.1)   # owners overwrite it (Mapper.py:51)
        # Mapper / tracker batches are depth-filtered (common.py:248-254): owners may set this to skip the
        # zero-depth branch (a handful of tiny device ops per call).
        self.assume_positive_depth = False
        self.write_back_radius = True        # see render_batch_ray
        self.min_render_chunk = 65536        # see render_img
        if self.N_samples > 0 or self.use_bound or self.fix_interval or not self.occupancy:
            raise NotImplementedError("only the reference's live rendering modes are built: N_samples=0, "
                                      "use_bound False, occupancy True (sample_near_pcl either way)")
        if slam is not None:
            self.H, self.W, self.fx, self.fy, self.cx, self.cy = slam.H, slam.W, slam.fx, slam.fy, slam.cx, slam.cy

    # Renderer.py:38-82
    def eval_points(self, p, decoders, npc, stage=None, device=None, npc_geo_feats=None, npc_col_feats=None,
                    is_tracker=False, cloud_pos=None, normals=None, pts_views_d=None, ray_pts_num=None,
                    dynamic_r_query=None, exposure_feat=None):
        assert torch.is_tensor(p)
        step_pts = (self.points_batch_size // ray_pts_num) * ray_pts_num   # whole rays per chunk
        rets, ray_masks, point_masks = [], [], []
        for s in range(0, p.shape[0], step_pts):
            rq = None
            if dynamic_r_query is not None:
                rq = {k: v[s:s + step_pts] for k, v in dynamic_r_query.items()}
            ret, ray_mask, point_mask = decoders(p[s:s + step_pts], npc, stage, npc_geo_feats, npc_col_feats,
                                                 ray_pts_num, is_tracker, cloud_pos, normals, pts_views_d, rq,
                                                 exposure_feat)
            rets.append(ret)
            ray_masks.append(ray_mask)
            point_masks.append(point_mask)
        if len(rets) == 1:
            return rets[0], ray_masks[0], point_masks[0]
        return torch.cat(rets, 0), torch.cat(ray_masks, 0), torch.cat(point_masks, 0)

    def surface_samples(self, gt_depth, device):
        """z_vals (R,S) and the gt>0 mask: Renderer.py:122-210 for N_samples = 0."""
        S = self.N_surface
        gt = gt_depth.reshape(-1, 1)
        nz = (gt > 0).squeeze(-1)
        t = torch.linspace(0.0, 1.0, steps=S, device=device)
        d = gt[nz].repeat(1, S)
        z = torch.zeros(gt.shape[0], S, device=device)
        z[nz, :] = self.near_end_surface * d * (1. - t) + self.far_end_surface * d * t
        return z, nz

    def _far_end(self, gt_depth):
        """0-d tensor max(far): clamp(min(5 mean(d), max(1.2 d)), 0, max(1.2 d)) when any depth is positive, else the
        unclamped bound (Renderer.py:133-134,148-153); device ops only (no host synchronisation)."""
        gt = gt_depth.reshape(-1).float()
        hi = torch.max(gt * 1.2)
        far_bb = torch.minimum(5 * gt.mean(), hi)
        return torch.where(hi > 0, torch.clamp(far_bb, min=0.0).minimum(hi), far_bb)

    def _z_for_zero_depth(self, gt_depth, device):
        """linspace(near_end, max(far)) for rays without a depth reading (Renderer.py:207-208)."""
        return _linspace32(self.near_end, self._far_end(gt_depth), self.N_surface, device)

    def _z_rows_for_zero_depth(self, gt_depth, chunk, device):
        """(R,S) z rows for render_img's widened chunks: ray r carries linspace(near_end, max(far)) of the REFERENCE's
        chunk r // chunk - the far end of rays without a depth reading is a statistic of the 3000-ray chunk they are
        rendered in (Renderer.py:133-134, 148-153, 207-208), so their colour / uncertainty depend on the chunking.  One
        segmented reduction over the frame instead of one per chunk; device ops only."""
        S = self.N_surface
        gt = gt_depth.reshape(-1).float()
        R = gt.shape[0]
        n = (R + chunk - 1) // chunk
        pad = n * chunk - R
        g2 = (torch.cat([gt, gt.new_zeros(pad)]) if pad else gt).view(n, chunk)      # depths are >= 0: zeros are neutral
        count = torch.full((n,), float(chunk), device=gt.device)
        count[-1] = float(chunk - pad)
        hi = (g2 * 1.2).max(1).values
        far_bb = torch.minimum(5 * (g2.sum(1) / count), hi)
        far = torch.where(hi > 0, torch.clamp(far_bb, min=0.0).minimum(hi), far_bb)
        rows = _linspace32(self.near_end, far[:, None], S, device)                    # (n,S)
        return rows.repeat_interleave(chunk, 0)[:R].contiguous()

    def _z_near_pcl(self, npc, rays_o, rays_d, gt_depth, radius, level):
        """sample_near_pcl mode (Renderer.py:196-206): rays without a depth reading are sampled between the first two
        of 25 probe positions that have neighbours in the cloud.  -> (z rows (R,S) - only the zero-depth rows are
        read by the sampling kernel -, mask_rays_near_pcl (R,) bool)."""
        R, S, dev = rays_o.shape[0], self.N_surface, rays_o.device
        none = torch.nonzero(~(gt_depth > 0)).squeeze(1)              # the mode's one host synchronisation
        z_rows = torch.zeros(R, S, device=dev)
        near_mask = torch.ones(R, dtype=torch.bool, device=dev)
        if none.numel():
            rad = radius[none] if torch.is_tensor(radius) else None
            z0, not_near = npc.sample_near_pcl(rays_o[none].detach(), rays_d[none].detach(), self.near_end,
                                               self._far_end(gt_depth), S, rad, level)
            z_rows[none] = z0
            near_mask[none[not_near]] = False
        return z_rows, near_mask

    # Renderer.py:84-254
    def prepare_batch(self, npc, rays_d, rays_o, stage, gt_depth, dynamic_r_query=None, out=None):
        """The parameter-free head of render_batch_ray for the same arguments (sampling + neighbour search) as an
        ops.RayPrep; pass it back as render_batch_ray(..., prepared=prep).  Lets a caller issue it early - e.g. for the
        next iteration while this iteration's gradients are still on the wire."""
        level = "mid" if stage.endswith("mid") else "fine"
        gt_depth = gt_depth.reshape(-1)
        radius = dynamic_r_query[level].reshape(-1) if self.use_dynamic_radius else npc.get_radius_query()
        index = npc.get_index(level)
        if self.sample_near_pcl and not self.assume_positive_depth:
            raise NotImplementedError("prepare_batch serves depth-filtered batches; sample_near_pcl needs the "
                                      "near-cloud mask of render_batch_ray")
        return ops.prepare_rays(index.grid_for(radius), rays_o, rays_d, gt_depth, radius, self.N_surface,
                                self.near_end_surface, self.far_end_surface,
                                None if self.assume_positive_depth else self._z_for_zero_depth(gt_depth, rays_o.device),
                                out=out)

    def render_batch_ray(self, npc, decoders, rays_d, rays_o, device, stage, gt_depth=None,
                         npc_geo_feats=None, npc_col_feats=None, is_tracker=False, cloud_pos=None,
                         normals=None, dynamic_r_query=None, exposure_feat=None, prepared=None, _loss=None,
                         _z_zero_rows=None, tail_work=None, grad_buffer=None, after_grads=None, after_feature_grads=None):
        """One fused kernel sequence (ops.render_rays): sampling, exact neighbour search, interpolation of both
        feature tables from one neighbour list, both decoders, compositing with the -100 overwrite.
        `prepared` (not in the reference): the ops.RayPrep of prepare_batch() for these very rays."""
        S = self.N_surface
        N_rays = rays_o.shape[0]
        if stage not in ("geometry_mid", "color_mid", "geometry_fine", "color_fine"):
            raise NotImplementedError("stage %r is not live in the reference" % (stage,))
        level = "mid" if stage.endswith("mid") else "fine"
        color_stage = stage.startswith("color")
        no_depth_given = gt_depth is None
        if gt_depth is None or torch.numel(gt_depth) == 0:
            gt_depth = torch.zeros(N_rays, device=device)
        gt_depth = gt_depth.reshape(-1)
        if self.use_dynamic_radius:
            radius = dynamic_r_query[level].reshape(-1)
            if self.write_back_radius:
                # Renderer.py:226-229 repeats every radius vector per sample and writes it back into the caller's
                # dict.  The kernels do not need it (the sampling kernel repeats on the fly): two copy kernels kept
                # only for that visible side effect; set renderer.write_back_radius = False to skip them.
                for key in dynamic_r_query.keys():
                    dynamic_r_query[key] = dynamic_r_query[key].reshape(-1, 1).repeat_interleave(S, dim=0)
        else:
            radius = npc.get_radius_query()
        index = npc.get_index(level)
        near_mask = None
        if self.assume_positive_depth:
            z_zero = None
        elif self.sample_near_pcl:
            z_zero, near_mask = self._z_near_pcl(npc, rays_o, rays_d, gt_depth, radius, level)
        elif _z_zero_rows is not None:      # render_img: rows computed per reference chunk (_z_rows_for_zero_depth)
            z_zero = _z_zero_rows
        elif no_depth_given:                # Renderer.py:135-141: without a depth image the far bound is 10 m
            z_zero = _linspace32(self.near_end, torch.full((), 10.0, device=device), S, device)
        else:
            z_zero = self._z_for_zero_depth(gt_depth, device)
        geo = getattr(decoders, "geo_decoder_" + level)
        col = getattr(decoders, "color_decoder_" + level)
        affine, aff_prm = None, None
        if color_stage and exposure_feat is not None and col.encode_exposure:        # decoder.py:606-612
            mlp = col.mlp_exposure
            if mlp.kernel_fits(exposure_feat):
                # the MLP's own tensors go into the fused call: evaluated beside the neighbour search, its backward
                # beside the feature scatter
                aff_prm = (exposure_feat, mlp.linear1.weight, mlp.linear1.bias, mlp.linear2.weight, mlp.linear2.bias)
            else:
                affine = mlp(exposure_feat)
        spec = ops.RenderSpec(
            grid=index.grid_for(radius), radius=radius,
            z_zero=z_zero, S=S,
            near=self.near_end_surface, far=self.far_end_surface, coef=self.sigmoid_coefficient,
            min_nn=geo.min_nn_num, min_valid=int(S / 2 + 1), color_stage=color_stage, tracker=bool(is_tracker),
            cloud_pos=cloud_pos[level] if is_tracker else None, fill_geo=decoders._fill("geo", rays_o.device),
            fill_col=decoders._fill("col", rays_o.device) if color_stage else None, prep=prepared,
            aff_mlp=aff_prm is not None, tail_work=tail_work, grad_buffer=grad_buffer, after_grads=after_grads,
            after_feature_grads=after_feature_grads)
        loss = None
        if _loss is not None and _loss[0] == "tracker":
            _, gt_color, w_color, reduce_fn = _loss
            spec.loss = ops.TrackerLossSpec(gt_color if color_stage else None, w_color, reduce_fn,
                                            zero_missing=not self.sample_near_pcl and not self.assume_positive_depth)
            loss, depth, uncertainty, color, valid = ops.render_rays(
                spec, rays_o, rays_d, gt_depth, npc_geo_feats[level], npc_col_feats[level] if color_stage else None,
                affine, geo.kernel_params(rays_o.device), col.kernel_params(rays_o.device) if color_stage else None,
                aff_prm)
        elif _loss is not None:
            _, gt_color, w_color, loss_exposure_feat, loss_affine, frame_index = _loss
            if loss_exposure_feat is not None and loss_affine is not None:
                raise ValueError("pass exposure_feat (one frame, MLP evaluated inside) or affine, not both")
            mlp = col.mlp_exposure if (color_stage and loss_exposure_feat is not None) else None
            eprm, mode = None, None
            if mlp is not None:
                mode = "mlp"
                eprm = (loss_exposure_feat, mlp.linear1.weight, mlp.linear1.bias, mlp.linear2.weight, mlp.linear2.bias)
            elif color_stage and loss_affine is not None:
                mode, eprm = "affine", (loss_affine,)
            spec.loss = ops.MapperLossSpec(gt_color if color_stage else None, w_color, mode, frame_index)
            loss, depth, uncertainty, color, valid = ops.render_rays(
                spec, rays_o, rays_d, gt_depth, npc_geo_feats[level], npc_col_feats[level] if color_stage else None,
                affine, geo.kernel_params(rays_o.device), col.kernel_params(rays_o.device) if color_stage else None,
                eprm)
        else:
            depth, uncertainty, color, valid = ops.render_rays(
                spec, rays_o, rays_d, gt_depth, npc_geo_feats[level], npc_col_feats[level] if color_stage else None,
                affine, geo.kernel_params(rays_o.device), col.kernel_params(rays_o.device) if color_stage else None,
                aff_prm)
        if not self.sample_near_pcl and not self.assume_positive_depth:
            depth = torch.where(gt_depth > 0, depth, torch.zeros_like(depth))     # Renderer.py:250-251
        if self.skip_zero_depth_pixel:
            color = torch.where((gt_depth > 0)[:, None], color, torch.zeros_like(color))
        valid = valid.view(torch.bool)
        if near_mask is not None:
            valid = valid & near_mask                                              # Renderer.py:247-248
        if _loss is not None:
            return loss, depth, uncertainty, color, valid
        return depth, uncertainty, color, valid

    def render_batch_ray_loss(self, npc, decoders, rays_d, rays_o, device, stage, gt_depth, gt_color, w_color_loss,
                              npc_geo_feats=None, npc_col_feats=None, cloud_pos=None, dynamic_r_query=None,
                              exposure_feat=None, prepared=None, affine=None, frame_index=None, tail_work=None,
                              grad_buffer=None, after_grads=None, after_feature_grads=None):
        """One mapper iteration's forward as ONE autograd node (not in the reference, which spells it out in
        Mapper.py:783-856): render_batch_ray(..., exposure_feat=None) followed by
            mask = (gt_depth > 0) & valid & ~isnan(depth)
            loss = |gt_depth - depth|[mask].sum()
                   + w_color_loss * |gt_color - sigmoid(color @ A + t)|[mask].sum()        (colour stages)
        with [A | t] = color_decoder_<level>.mlp_exposure(exposure_feat) (no exposure_feat: the colour as rendered).
        A batch that mixes rays of F keyframes (Mapper.py:805-822) passes the affines instead: affine (F,12) =
        stack of mlp_exposure(exposure_feat_f), frame_index (R,) = the keyframe slot of every ray.
        -> (loss, depth, uncertainty, color, valid_ray_mask); only `loss` carries gradient (to the feature tables, the
        decoders, the exposure MLP and exposure_feat).  The compositing and loss kernels are one, the exposure MLP
        runs beside the neighbour search, and loss.backward() feeds the upstream gradient to the kernels as a device
        scalar - ten launches fewer on the critical path than render_batch_ray + mapper_loss.
        tail_work: a callable issued on a side stream beside the tail of loss.backward() (weight gradients + feature
        scatter, HBM-bound) - e.g. `lambda: renderer.prepare_batch(<the next iteration's rays>, out=prep_next)`: the
        next batch's pixels are drawn independently of the parameters (common.py:199-258), so its sampling + neighbour
        search need not wait for this iteration's optimiser step.
        grad_buffer: an ops.GradPool - every gradient of loss.backward() is laid out in its tensor (zeroed first) instead
        of a fresh allocation; with parallel.PeerGradExchange's pool the multi-GPU exchange sums them in place.
        after_grads: a callable issued on the main stream inside loss.backward() once every gradient is complete and before
        tail_work is joined - the place for that exchange (`after_grads=exchange`), which then runs beside the tail work.
        after_feature_grads: a callable issued inside loss.backward() right behind the feature scatter, i.e. when the
        feature-table gradients (grad_buffer.tensor[:grad_buffer.features], 85 % of the exchange's bytes) are complete and
        the decoders' weight-gradient kernels still run (`after_feature_grads=exchange.features`: that part of the exchange
        is then off the critical path and `exchange()` in after_grads only moves the rest)."""
        return self.render_batch_ray(npc, decoders, rays_d, rays_o, device, stage, gt_depth=gt_depth,
                                     npc_geo_feats=npc_geo_feats, npc_col_feats=npc_col_feats, is_tracker=False,
                                     cloud_pos=cloud_pos, dynamic_r_query=dynamic_r_query, exposure_feat=None,
                                     prepared=prepared, tail_work=tail_work, grad_buffer=grad_buffer, after_grads=after_grads,
                                     after_feature_grads=after_feature_grads,
                                     _loss=("mapper", gt_color, w_color_loss, exposure_feat, affine, frame_index))

    def render_batch_ray_tracker_loss(self, npc, decoders, rays_d, rays_o, device, stage, gt_depth, gt_color,
                                      w_color_loss, npc_geo_feats=None, npc_col_feats=None, cloud_pos=None,
                                      dynamic_r_query=None, exposure_feat=None, use_color=True, reduce_fn=None):
        """One tracker iteration's forward as ONE autograd node (the reference spells it out in Tracker.py:199-231):
        render_batch_ray(..., is_tracker=True, exposure_feat=exposure_feat) followed by
            tmp  = |gt_depth - depth| / sqrt(uncertainty + 1e-10)          (uncertainty detached)
            mask = (tmp < 10 * tmp.mean()) & (gt_depth > 0) & finite
            loss = clamp(tmp, 0, 1e3)[mask].sum() + w_color_loss * |gt_color - color|[mask].sum()   (use_color)
        -> (loss, depth, uncertainty, color, valid_ray_mask); only `loss` carries gradient (pose through rays_o / rays_d,
        exposure_feat, and whatever else requires it).  reduce_fn: see ops.tracker_loss (rays sharded over ranks)."""
        return self.render_batch_ray(npc, decoders, rays_d, rays_o, device, stage, gt_depth=gt_depth,
                                     npc_geo_feats=npc_geo_feats, npc_col_feats=npc_col_feats, is_tracker=True,
                                     cloud_pos=cloud_pos, dynamic_r_query=dynamic_r_query, exposure_feat=exposure_feat,
                                     _loss=("tracker", gt_color if use_color else None, w_color_loss, reduce_fn))

    def render_batch_ray_modular(self, npc, decoders, rays_d, rays_o, device, stage, gt_depth=None,
                                 npc_geo_feats=None, npc_col_feats=None, is_tracker=False, cloud_pos=None,
                                 normals=None, dynamic_r_query=None, exposure_feat=None):
        """Same result through the reference's own seams (eval_points -> NICER.forward -> compositing): one autograd
        Function per kernel.  Kept as the cross-check of the fused path and for callers that hook the seams."""
        S = self.N_surface
        N_rays = rays_o.shape[0]
        if gt_depth is None or torch.numel(gt_depth) == 0:
            gt_depth = torch.zeros(N_rays, device=device)
        z_vals, nz = self.surface_samples(gt_depth, device)
        near_mask = None
        if self.sample_near_pcl:
            level = "mid" if stage.endswith("mid") else "fine"
            radius = dynamic_r_query[level].reshape(-1) if self.use_dynamic_radius else None
            z_rows, near_mask = self._z_near_pcl(npc, rays_o, rays_d, gt_depth.reshape(-1), radius, level)
            z_vals = torch.where(nz[:, None], z_vals, z_rows)
        else:
            z_vals = torch.where(nz[:, None], z_vals,
                                 self._z_for_zero_depth(gt_depth, device)[None, :].expand_as(z_vals))
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]      # Renderer.py:218-219
        pointsf = pts.reshape(-1, 3)
        rq = None
        if self.use_dynamic_radius:
            rq = {}
            for key in dynamic_r_query.keys():
                dynamic_r_query[key] = dynamic_r_query[key].reshape(-1, 1).repeat_interleave(S, dim=0)
                rq[key] = dynamic_r_query[key]
        raw, valid_ray_mask, point_mask = self.eval_points(
            pointsf, decoders, npc, stage, device, npc_geo_feats, npc_col_feats, is_tracker, cloud_pos, normals,
            None, ray_pts_num=S, dynamic_r_query=rq, exposure_feat=exposure_feat)
        depth, uncertainty, color, _ = ops.composite(raw.reshape(N_rays, S, 4), z_vals, point_mask,
                                                     self.sigmoid_coefficient)
        if not self.sample_near_pcl:
            depth = torch.where(nz, depth, torch.zeros_like(depth))
        if self.skip_zero_depth_pixel:
            color = torch.where(nz[:, None], color, torch.zeros_like(color))
        if near_mask is not None:
            valid_ray_mask = valid_ray_mask & near_mask
        return depth, uncertainty, color, valid_ray_mask

    # Renderer.py:256-337
    def render_img(self, npc, decoders, c2w, device, stage, gt_depth=None, npc_geo_feats=None,
                   npc_col_feats=None, normals=None, dynamic_r_query=None, cloud_pos=None, exposure_feat=None,
                   row_range=None):
        """row_range (not in the reference): (r0, r1) renders image rows [r0, r1) only and returns (r1 - r0, W) maps -
        the unit parallel.render_img_sharded deals out to the ranks of a multi-GPU job.  Rays are independent, and the
        one frame-level statistic (the far bound of rays without a depth reading, a statistic of the REFERENCE's
        3000-ray chunk) is taken over the whole frame before the rows are cut, so the rows equal the full render's."""
        with torch.no_grad():
            H, W = self.H, self.W
            rays_o, rays_d = get_rays(H, W, self.fx, self.fy, self.cx, self.cy, c2w, device)
            rays_o, rays_d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
            if self.use_dynamic_radius:
                for key in dynamic_r_query.keys():
                    dynamic_r_query[key] = dynamic_r_query[key].reshape(-1, 1)
            gt_depth = gt_depth.reshape(-1) if gt_depth is not None else None
            depths, uncs, colors = [], [], []
            # Rays are independent and the reference only chunks to bound memory (ray_batch_size = 3000 -> 103 chunks
            # per frame, host-bound here): chunks are widened to min_render_chunk rays.  The one thing that depends on
            # the chunking - the sampling range of rays without a depth reading - is computed per reference chunk.
            B = max(self.ray_batch_size, self.min_render_chunk)
            z_rows = None
            if ((B > self.ray_batch_size or row_range is not None) and gt_depth is not None and not self.sample_near_pcl
                    and not self.assume_positive_depth):
                z_rows = self._z_rows_for_zero_depth(gt_depth, self.ray_batch_size, rays_o.device)
            lo, hi = 0, rays_d.shape[0]
            if row_range is not None:
                r0, r1 = int(row_range[0]), int(row_range[1])
                if not 0 <= r0 <= r1 <= H:
                    raise ValueError("row_range must satisfy 0 <= r0 <= r1 <= H")
                if self.sample_near_pcl:
                    raise NotImplementedError("row_range with sample_near_pcl (the probe far bound is a chunk statistic)")
                lo, hi, H = r0 * W, r1 * W, r1 - r0
            for i in range(lo, hi, B):
                j = min(i + B, hi)
                rq = {k: v[i:j] for k, v in dynamic_r_query.items()} if self.use_dynamic_radius else None
                d, u, c, _ = self.render_batch_ray(
                    npc, decoders, rays_d[i:j], rays_o[i:j], device, stage,
                    gt_depth=None if gt_depth is None else gt_depth[i:j], npc_geo_feats=npc_geo_feats,
                    npc_col_feats=npc_col_feats, normals=normals, cloud_pos=cloud_pos, dynamic_r_query=rq,
                    exposure_feat=exposure_feat, _z_zero_rows=None if z_rows is None else z_rows[i:j])
                depths.append(d.double())
                uncs.append(u.double())
                colors.append(c)
            if not depths:
                z = torch.zeros(0, W, dtype=torch.float64, device=device)
                return z, z.clone(), torch.zeros(0, W, 3, device=device)
            depth = torch.cat(depths, 0).reshape(H, W)
            uncertainty = torch.cat(uncs, 0).reshape(H, W)
            color = torch.cat(colors, 0).reshape(H, W, 3)
            return depth, uncertainty, color
