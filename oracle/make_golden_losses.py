"""TEST INFRASTRUCTURE ONLY.  Generate tests/golden/losses_small.pt by EXECUTING the reference's own loss statements:
the mapper's (src/Mapper.py:793-856: depth mask, L1 depth term, per-keyframe exposure affine + sigmoid + L1 colour term,
for the geometry and both colour stages) and the tracker's (src/Tracker.py:213-231: uncertainty-normalised depth term
with the 10 x mean outlier mask, L1 colour term).  The statements sit in the middle of `Mapper.optimize_map` /
`Tracker.optimize_cam_in_batch`, so they are read from the reference files where they lie (/root/reference, never
copied into the repo), dedented and run on seeded inputs that stand for the renderer's outputs; autograd then yields the
gradients with respect to depth, colour and the keyframes' exposure affines.  `mlp_exposure` is replaced by a look-up of a
leaf (F,12) tensor so that the gradient with respect to every keyframe's affine can be stored.

Run in the build container only:   python oracle/make_golden_losses.py
The fixture is committed; tests/test_oracle_golden.py (oracle + the product's torch expressions, CPU) and the GPU suite
(loss kernels) compare against it.
"""
import os
import textwrap
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("HPS_REFERENCE_ROOT", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")


def ref_lines(name, first, last, starts, ends):
    with open(os.path.join(REF, "src", name)) as f:
        lines = f.readlines()
    src = textwrap.dedent("".join(lines[first - 1:last]))
    assert src.lstrip().startswith(starts), src[:80]
    assert src.rstrip().endswith(ends), src[-80:]
    return compile(src, "%s:%d-%d" % (name, first, last), "exec")


def main():
    mapper_code = ref_lines("Mapper.py", 793, 856, "depth_mask = (batch_gt_depth > 0) & valid_ray_mask", "loss += weighted_color_loss")
    tracker_code = ref_lines("Tracker.py", 213, 231, "uncertainty = uncertainty.detach()", "loss += self.w_color_loss*color_loss")
    g = torch.Generator().manual_seed(1219)
    F, w_map, w_trk = 4, 0.1, 0.5
    sizes = [310, 5, 402, 333]                                  # rays per keyframe before the inside mask
    N0 = sum(sizes)
    inside_mask = torch.rand(N0, generator=g) > 0.07
    inside_mask[sizes[0]:sizes[0] + 2] = True                   # the 5-ray keyframe keeps at least two rays
    R = int(inside_mask.sum())
    frame = torch.cat([torch.full((n,), 10 * i + 3, dtype=torch.long) for i, n in enumerate(sizes)])   # frame ids, not 0..F-1
    depth0 = 1.0 + 2.0 * torch.rand(R, generator=g)
    depth0[5] = float("nan")                                   # a NaN depth is masked out (:794)
    color0 = torch.randn(R, 3, generator=g)                    # raw rgb, before exposure affine and sigmoid
    gt_depth = depth0.nan_to_num(2.0) + 0.1 * torch.randn(R, generator=g)
    gt_depth[::11] = 0.0                                       # missing depth
    gt_color = torch.rand(R, 3, generator=g)
    valid = torch.rand(R, generator=g) > 0.1
    affine0 = torch.cat([torch.eye(3).reshape(1, 9).repeat(F, 1) + 0.2 * torch.randn(F, 9, generator=g),
                         0.1 * torch.randn(F, 3, generator=g)], 1)
    out = dict(F=F, sizes=sizes, inside_mask=inside_mask, frame_ids=frame, depth=depth0, color=color0, gt_depth=gt_depth,
               gt_color=gt_color, valid=valid, affine=affine0, w_color_mapper=w_map, w_color_tracker=w_trk, mapper={},
               tracker={})

    # ---- mapper: every stage, with and without exposure encoding
    for stage in ("geometry_mid", "geometry_fine", "color_mid", "color_fine"):
        for encode in (True,):              # encode_exposure False (colour tail inside the decoder) is not built
            depth = depth0.clone().requires_grad_(True)
            color = color0.clone().requires_grad_(True)
            affine = affine0.clone().requires_grad_(True)
            lookup = types.SimpleNamespace(mlp_exposure=lambda k: affine[k])
            me = types.SimpleNamespace(stage=stage, encode_exposure=encode, w_color_loss=w_map,
                                       decoders=types.SimpleNamespace(color_decoder_mid=lookup, color_decoder_fine=lookup))
            ns = dict(torch=torch, self=me, batch_gt_depth=gt_depth, batch_gt_color=gt_color, valid_ray_mask=valid,
                      depth=depth, color=color * 1.0,          # the statements write into `color` in place (:821)
                      indices_tensor=[frame[sum(sizes[:i]):sum(sizes[:i + 1])] for i in range(F)],
                      inside_mask=inside_mask, exposure_feat_list=list(range(F)))
            exec(mapper_code, ns)
            loss = ns["loss"]
            loss.backward()
            out["mapper"][(stage, encode)] = dict(
                loss=loss.detach().clone(), g_depth=depth.grad.clone(),
                g_color=None if color.grad is None else color.grad.clone(),
                g_affine=None if affine.grad is None else affine.grad.clone())
            print("mapper  %-13s encode_exposure=%-5s loss %.6f" % (stage, encode, float(loss)))

    # ---- tracker (handle_dynamic True = the configured value, and False; with and without the colour term)
    unc0 = 0.01 + 0.2 * torch.rand(R, generator=g)
    unc0[::37] *= 1e-4                                         # confident rays: large normalised residuals (outliers)
    out["uncertainty"] = unc0
    col_t = torch.rand(R, 3, generator=g)                      # tracker colours are already sigmoid outputs
    out["tracker_color"] = col_t
    depth_t = depth0.nan_to_num(1.5)                           # a NaN would poison tmp.mean() and empty the mask (:217-218)
    out["tracker_depth"] = depth_t
    for dyn in (True, False):
        for use_color in (True, False):
            depth = depth_t.clone().requires_grad_(True)
            color = col_t.clone().requires_grad_(True)
            unc = unc0.clone().requires_grad_(True)
            me = types.SimpleNamespace(handle_dynamic=dyn, use_color_in_tracking=use_color, w_color_loss=w_trk)
            ns = dict(torch=torch, self=me, batch_gt_depth=gt_depth, batch_gt_color=gt_color, depth=depth, color=color,
                      uncertainty=unc)
            exec(tracker_code, ns)
            loss = ns["loss"]
            loss.backward()
            assert unc.grad is None                              # detached (:213)
            out["tracker"][(dyn, use_color)] = dict(loss=loss.detach().clone(), g_depth=depth.grad.clone(),
                                                    g_color=None if color.grad is None else color.grad.clone(),
                                                    n_mask=int(ns["mask"].sum()))
            print("tracker handle_dynamic=%-5s use_color=%-5s loss %.6f (%d rays in the mask)"
                  % (dyn, use_color, float(loss), int(ns["mask"].sum())))
    path = os.path.join(GOLD, "losses_small.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
