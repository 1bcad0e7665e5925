"""Generate tests/golden/*.pt by running the UNMODIFIED reference modules (imported from
/root/reference under the stubs of oracle/ref_import.py) on seeded synthetic inputs, and check the
oracle restatement (oracle/hps_oracle.py) against them on the spot.

Run in the build container only:   python oracle/make_golden.py
The fixtures it writes are committed; the GPU box (no /root/reference) replays them.

Neutralised reference quirks: the two RNG draws in get_feature_at_pos are pinned to explicit
c_fill vectors; 'cuda:-1' device strings (decoder.py:679, common.py:317) are mapped to cpu.
"""
import contextlib
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
from oracle import hps_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
LEVEL_RADIUS = {"fine": (0.04, 0.16), "mid": (0.24, 0.96)}


class _TorchProxy:
    """torch stand-in for the reference decoder module: maps 'cuda:-1' (CPU tensor's get_device())
    to 'cpu' in torch.zeros (decoder.py:679,686,696)."""

    def __init__(self):
        self._t = torch

    def __getattr__(self, k):
        return getattr(self._t, k)

    def zeros(self, *a, **k):
        if isinstance(k.get("device"), str) and k["device"].startswith("cuda:-"):
            k["device"] = "cpu"
        return torch.zeros(*a, **k)


@contextlib.contextmanager
def pinned_reference_rng(fills):
    """normal_(0, 0.01) -> deterministic: a (c_dim,) draw takes the next vector from `fills`,
    any other shape (the discarded (P,32) draw) becomes zeros.  `.to(-1)` -> identity."""
    orig_normal, orig_to = torch.Tensor.normal_, torch.Tensor.to
    it = iter(fills)

    def normal_(self, mean=0, std=1, *, generator=None):
        if std == 0.01 and mean == 0:
            if self.dim() == 1:
                return self.copy_(next(it))
            return self.zero_()
        return orig_normal(self, mean, std, generator=generator)

    def to(self, *a, **k):
        if len(a) == 1 and isinstance(a[0], int) and a[0] < 0:
            return self
        return orig_to(self, *a, **k)

    torch.Tensor.normal_, torch.Tensor.to = normal_, to
    try:
        yield
    finally:
        torch.Tensor.normal_, torch.Tensor.to = orig_normal, orig_to


def build_reference(ns, cfg, clouds):
    """Reference NeuralPointCloud (exhaustive index), decoders (seed 1219) and Renderer on CPU."""
    ns.decoder.torch = _TorchProxy()
    ns.common.setup_seed(1219)
    decoders = ns.config.get_model(cfg)
    npc = ns.neural_point.NeuralPointCloud(cfg)
    for level, cl in clouds.items():
        npc._cloud_pos[level] = cl["points"].tolist()
        npc._pts_num[level] = cl["points"].shape[0]
        npc.geo_feats[level] = cl["geo_feats"].clone()
        npc.col_feats[level] = cl["col_feats"].clone()
        idx = ref_import.ExhaustiveIndex()
        idx.train(cl["points"])
        idx.add(cl["points"])
        npc.index[level] = idx
    K = O.SCANNET_INTRINSICS
    slam = types.SimpleNamespace(nice=True, **K)
    renderer = ns.renderer.Renderer(cfg, None, slam)
    renderer.sigmoid_coefficient = cfg["rendering"]["sigmoid_coef_mapper"]
    return npc, decoders, renderer


def decoder_params(decoders):
    """Flat dict per decoder incl. the colour embedder's non-registered _B."""
    out = {}
    for name in ("geo_decoder_mid", "geo_decoder_fine", "color_decoder_mid", "color_decoder_fine"):
        m = getattr(decoders, name)
        d = {k: v.detach().clone() for k, v in m.state_dict().items()}
        d["embedder._B"] = m.embedder._B.detach().clone()
        out[name] = d
    return out


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def main():
    os.makedirs(GOLD, exist_ok=True)
    ns = ref_import.load_reference()
    cfg = ref_import.load_cfg(ns)
    torch.manual_seed(0)

    # ---------------- synthetic two-level cloud + rays ----------------
    n_pts = {"fine": 6000, "mid": 1500}
    R = 96
    clouds, rays = {}, None
    for li, level in enumerate(("mid", "fine")):
        sc = O.synthetic_scene(n_pts[level], R, LEVEL_RADIUS[level], seed=1219 + li)
        clouds[level] = {k: sc[k] for k in ("points", "geo_feats", "col_feats")}
        if rays is None:
            rays = {k: sc[k] for k in ("rays_o", "rays_d", "gt_depth", "gt_color")}
        clouds[level]["radius"] = sc["radius"]
    # a few zero-depth rays and a few rays far from the cloud to exercise masks
    rays["gt_depth"][5] = 0.0
    rays["gt_depth"][17] = 0.0
    rays["gt_depth"][40] *= 1.6
    g = torch.Generator().manual_seed(7)
    c_fill_geo = torch.randn(32, generator=g) * 0.01
    c_fill_col = torch.randn(32, generator=g) * 0.01
    exposure_feat = torch.randn(8, generator=g) * 0.5

    npc, decoders, renderer = build_reference(ns, cfg, clouds)
    # make the exposure MLP non-trivial (init std 0.01 gives ~identity-free affine ~0)
    with torch.no_grad():
        for lvl in ("mid", "fine"):
            m = getattr(decoders, "color_decoder_" + lvl).mlp_exposure
            m.linear2.bias.copy_(torch.tensor([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0.0])
                                 + 0.05 * torch.randn(12, generator=g))
            m.linear2.weight.mul_(20.0)
    dec = decoder_params(decoders)

    inputs = dict(clouds=clouds, rays=rays, c_fill_geo=c_fill_geo, c_fill_col=c_fill_col,
                  exposure_feat=exposure_feat, decoders=dec,
                  cfg=dict(near_end_surface=cfg["rendering"]["near_end_surface"],
                           far_end_surface=cfg["rendering"]["far_end_surface"],
                           n_surface=cfg["rendering"]["N_surface"], nn_num=cfg["pointcloud"]["nn_num"],
                           min_nn_num=cfg["pointcloud"]["min_nn_num"], coef=0.1,
                           w_color_mapper=cfg["mapping"]["w_color_loss"],
                           w_color_tracker=cfg["tracking"]["w_color_loss"]))
    torch.save(inputs, os.path.join(GOLD, "scene_small.pt"))
    report = []

    # ---------------- 1. neighbour query (find_neighbors_faiss) ----------------
    knn = {}
    for level in ("mid", "fine"):
        gq = torch.Generator().manual_seed(11)
        Q = 1500
        base = clouds[level]["points"][torch.randint(0, n_pts[level], (Q,), generator=gq)]
        rmin, rmax = LEVEL_RADIUS[level]
        q = (base + torch.randn(Q, 3, generator=gq) * rmax * 0.5).float()
        r = rmin + (rmax - rmin) * torch.rand(Q, 1, generator=gq, dtype=torch.float64)
        r[::7] = rmax                      # exactly-representable-ish repeated radius
        q[3] = clouds[level]["points"][10]  # a query exactly on a point (D = 0)
        q[4] = q[4] + 50.0                  # a query far away from everything
        D, I, n = npc.find_neighbors_faiss(q, step="query", dynamic_radius=r, level=level)
        Ds, Is, ns_ = npc.find_neighbors_faiss(q, step="query", level=level)  # scalar radius_query
        Do, Io, no = O.find_neighbors(clouds[level]["points"], q, r)
        assert torch.equal(D, Do) and torch.equal(I, Io) and torch.equal(n, no), "oracle knn mismatch"
        _, _, nso = O.find_neighbors(clouds[level]["points"], q, cfg["pointcloud"]["radius_query"])
        assert torch.equal(ns_, nso)
        knn[level] = dict(q=q, r=r, D=D, I=I.int(), n=n, n_scalar=ns_,
                          radius_query=cfg["pointcloud"]["radius_query"])
        report.append("knn[%s]: oracle == reference bit-exact (Q=%d, N=%d, mean n=%.2f)"
                      % (level, Q, n_pts[level], n.float().mean()))
    torch.save(knn, os.path.join(GOLD, "knn_small.pt"))

    # ---------------- 2. render_batch_ray + mapper loss + backward ----------------
    cloud_pos = {l: clouds[l]["points"] for l in clouds}
    renders = {}
    for stage in ("geometry_mid", "color_mid", "geometry_fine", "color_fine"):
        level = "mid" if stage.endswith("mid") else "fine"
        gf = {l: clouds[l]["geo_feats"].clone().requires_grad_(True) for l in clouds}
        cf = {l: clouds[l]["col_feats"].clone().requires_grad_(True) for l in clouds}
        decoders.zero_grad()
        rq = {l: clouds[l]["radius"].clone() for l in clouds}
        with pinned_reference_rng([c_fill_geo, c_fill_col]):
            depth, var, color, valid = renderer.render_batch_ray(
                npc, decoders, rays["rays_d"], rays["rays_o"], "cpu", stage,
                gt_depth=rays["gt_depth"], npc_geo_feats=gf, npc_col_feats=cf, is_tracker=False,
                cloud_pos=cloud_pos, dynamic_r_query=rq, exposure_feat=None)
        # Mapper.py:791-856 with one frame's exposure affine (single-frame batch)
        mask = (rays["gt_depth"] > 0) & valid & (~torch.isnan(depth))
        loss = torch.abs(rays["gt_depth"][mask] - depth[mask]).sum()
        affine = None
        if stage.startswith("color"):
            mexp = getattr(decoders, "color_decoder_" + level).mlp_exposure
            affine = mexp(exposure_feat)
            col2 = torch.sigmoid(color @ affine[:9].reshape(3, 3) + affine[-3:])
            loss = loss + cfg["mapping"]["w_color_loss"] * torch.abs(rays["gt_color"][mask] - col2[mask]).sum()
        loss.backward()
        cdec = getattr(decoders, ("color_decoder_" if stage.startswith("color") else "geo_decoder_") + level)
        out = dict(depth=depth.detach(), var=var.detach(), color=color.detach(), valid=valid,
                   loss=loss.detach(), g_geo=gf[level].grad.clone(),
                   g_col=None if cf[level].grad is None else cf[level].grad.clone(),
                   g_dec={k: v.grad.clone() for k, v in cdec.named_parameters() if v.grad is not None},
                   g_geo_dec={k: v.grad.clone() for k, v in
                              getattr(decoders, "geo_decoder_" + level).named_parameters() if v.grad is not None})
        renders[stage] = out
        # oracle check
        gfo = {l: clouds[l]["geo_feats"].clone().requires_grad_(True) for l in clouds}
        cfo = {l: clouds[l]["col_feats"].clone().requires_grad_(True) for l in clouds}
        deco = {k: {kk: vv.clone().requires_grad_(vv.is_floating_point()) for kk, vv in v.items()}
                for k, v in dec.items()}
        d2, v2, c2, m2 = O.render_batch_ray(cloud_pos, deco, rays["rays_d"], rays["rays_o"], stage,
                                            rays["gt_depth"], gfo, cfo,
                                            {l: clouds[l]["radius"] for l in clouds}, c_fill_geo, c_fill_col)
        aff_o = None
        if stage.startswith("color"):
            aff_o = O.exposure_affine(deco["color_decoder_" + level], exposure_feat)
        l2 = O.mapper_loss(d2, c2, m2, rays["gt_depth"], rays["gt_color"], stage,
                           cfg["mapping"]["w_color_loss"], aff_o)
        l2.backward()
        errs = dict(depth=rel_err(d2, depth), var=rel_err(v2, var), color=rel_err(c2, color),
                    loss=rel_err(l2, loss), g_geo=rel_err(gfo[level].grad, out["g_geo"]))
        assert torch.equal(m2, valid)
        if out["g_col"] is not None:
            errs["g_col"] = rel_err(cfo[level].grad, out["g_col"])
        key = ("color_decoder_" if stage.startswith("color") else "geo_decoder_") + level
        errs["g_dec"] = max(rel_err(deco[key][k].grad, g) for k, g in out["g_dec"].items())
        report.append("render[%s]: valid rays %d/%d, oracle-vs-reference max rel err %s"
                      % (stage, int(valid.sum()), R, {k: "%.1e" % v for k, v in errs.items()}))
        assert max(errs.values()) < 2e-5, errs
    torch.save(renders, os.path.join(GOLD, "render_mapper_small.pt"))

    # ---------------- 3. tracker iteration (pose + exposure gradients) ----------------
    track = {}
    K = O.SCANNET_INTRINSICS
    gp = torch.Generator().manual_seed(5)
    cam0 = torch.tensor([1.0, 0.004, -0.003, 0.002, 0.01, -0.008, 0.005])
    R_t = rays["rays_d"].shape[0]
    pi = torch.randint(10, K["W"] - 10, (R_t,), generator=gp).float()
    pj = torch.randint(10, K["H"] - 10, (R_t,), generator=gp).float()
    depth_img = O.synthetic_depth()
    gt_d = depth_img[pj.long(), pi.long()].clone()
    gt_c = torch.rand(R_t, 3, generator=gp)
    for stage in ("color_mid", "color_fine"):
        level = "mid" if stage.endswith("mid") else "fine"
        cam = cam0.clone().requires_grad_(True)
        ef = exposure_feat.clone().requires_grad_(True)
        decoders.zero_grad()
        rq = {l: clouds[l]["radius"].clone() for l in clouds}
        with pinned_reference_rng([c_fill_geo, c_fill_col]):
            c2w = ns.common.get_camera_from_tensor(cam)
            ro, rd = ns.common.get_rays_from_uv(pi, pj, c2w, K["H"], K["W"], K["fx"], K["fy"], K["cx"], K["cy"], "cpu")
            depth, var, color, _ = renderer.render_batch_ray(
                npc, decoders, rd, ro, "cpu", stage, gt_depth=gt_d,
                npc_geo_feats={l: clouds[l]["geo_feats"] for l in clouds},
                npc_col_feats={l: clouds[l]["col_feats"] for l in clouds}, is_tracker=True,
                cloud_pos=cloud_pos, dynamic_r_query=rq, exposure_feat=ef)
        # Tracker.py:213-231
        unc = var.detach()
        nan_mask = (~torch.isnan(depth)) & (~torch.isnan(unc))
        tmp = torch.abs(gt_d - depth) / torch.sqrt(unc + 1e-10)
        mask = (tmp < 10 * tmp.mean()) & (gt_d > 0) & nan_mask
        geo_loss = torch.clamp(tmp, min=0.0, max=1e3)[mask].sum()
        color_loss = torch.abs(gt_c - color)[mask].sum()
        loss = geo_loss + cfg["tracking"]["w_color_loss"] * color_loss
        loss.backward()
        mexp = getattr(decoders, "color_decoder_" + level).mlp_exposure
        out = dict(depth=depth.detach(), var=var.detach(), color=color.detach(), loss=loss.detach(),
                   g_cam=cam.grad.clone(), g_exposure=ef.grad.clone(),
                   g_mlp_exposure={k: v.grad.clone() for k, v in mexp.named_parameters()})
        track[stage] = out
        # oracle
        cam_o = cam0.clone().requires_grad_(True)
        ef_o = exposure_feat.clone().requires_grad_(True)
        deco = {k: {kk: vv.clone().requires_grad_(vv.is_floating_point()) for kk, vv in v.items()}
                for k, v in dec.items()}
        c2w_o = O.camera_from_tensor(cam_o)
        ro_o, rd_o = O.rays_from_uv(pi, pj, c2w_o, K["fx"], K["fy"], K["cx"], K["cy"])
        d2, v2, c2, _ = O.render_batch_ray(cloud_pos, deco, rd_o, ro_o, stage, gt_d,
                                           {l: clouds[l]["geo_feats"] for l in clouds},
                                           {l: clouds[l]["col_feats"] for l in clouds},
                                           {l: clouds[l]["radius"] for l in clouds}, c_fill_geo, c_fill_col,
                                           is_tracker=True, exposure_feat=ef_o)
        l2 = O.tracker_loss(d2, v2, c2, gt_d, gt_c, cfg["tracking"]["w_color_loss"])
        l2.backward()
        errs = dict(depth=rel_err(d2, depth), var=rel_err(v2, var), color=rel_err(c2, color),
                    loss=rel_err(l2, loss), g_cam=rel_err(cam_o.grad, out["g_cam"]),
                    g_exposure=rel_err(ef_o.grad, out["g_exposure"]))
        report.append("tracker[%s]: oracle-vs-reference max rel err %s" % (stage, {k: "%.1e" % v for k, v in errs.items()}))
        assert max(errs.values()) < 2e-4, errs
    track["inputs"] = dict(cam=cam0, i=pi, j=pj, gt_depth=gt_d, gt_color=gt_c)
    torch.save(track, os.path.join(GOLD, "render_tracker_small.pt"))

    with open(os.path.join(GOLD, "REPORT.txt"), "w") as f:
        f.write("generated by oracle/make_golden.py from the imported reference (torch %s)\n" % torch.__version__)
        f.write("\n".join(report) + "\n")
    print("\n".join(report))


if __name__ == "__main__":
    main()
