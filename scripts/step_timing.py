"""GPU: wall-clock and CUDA-event time of the bench step with the TC / SIMT forward (diagnostic)."""
import os, sys, time, importlib, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import __graft_entry__
__graft_entry__.build()
hps = importlib.import_module("hierarchical-point-slam_b200")
dev = "cuda:0"
scene = bench.make_scene()
cfg = hps.scannet_cfg(dev)
npc = hps.NeuralPointCloud(cfg)
for level in ("fine", "mid"):
    s = scene[level]
    npc.set_cloud(level, s["points"].to(dev), s["geo_feats"].to(dev), s["col_feats"].to(dev))
torch.manual_seed(1219)
dec = hps.get_model(cfg).to(dev)
dec.c_fill_geo, dec.c_fill_col = scene["c_fill_geo"].to(dev), scene["c_fill_col"].to(dev)
renderer = hps.Renderer(cfg, None, types.SimpleNamespace(**{k: cfg["cam"][k] for k in ("H", "W", "fx", "fy", "cx", "cy")}))
fine = scene["fine"]
R = bench.RAYS
d = {k: fine[k][:R].to(dev) for k in ("rays_o", "rays_d", "gt_depth", "gt_color", "radius")}
gf = {"fine": npc.get_geo_feats("fine").clone().requires_grad_(True), "mid": npc.get_geo_feats("mid")}
cf = {"fine": npc.get_col_feats("fine").clone().requires_grad_(True), "mid": npc.get_col_feats("mid")}
pos = {l: npc.cloud_pos_tensor(l) for l in ("fine", "mid")}
ef = scene["exposure_feat"].to(dev)
params = list(dec.parameters())
def step():
    for t in [gf["fine"], cf["fine"]] + params:
        t.grad = None
    rq = {"fine": d["radius"], "mid": d["radius"]}
    depth, unc, color, valid = renderer.render_batch_ray(npc, dec, d["rays_d"], d["rays_o"], dev, "color_fine", gt_depth=d["gt_depth"],
        npc_geo_feats=gf, npc_col_feats=cf, is_tracker=False, cloud_pos=pos, dynamic_r_query=rq, exposure_feat=None)
    aff = dec.color_decoder_fine.mlp_exposure(ef)
    loss = hps.mapper_loss(depth, color, valid, d["gt_depth"], d["gt_color"], "color_fine", 0.1, aff)
    loss.backward()
for tc in (False, True, False, True):
    hps.ops.USE_TC = tc
    for _ in range(5): step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): step()
    t_cpu = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print("USE_TC", tc, "cpu issue %.3f ms/step, wall %.3f ms/step" % (t_cpu / 20 * 1e3, t_all / 20 * 1e3))
from torch.profiler import profile, ProfilerActivity
hps.ops.USE_TC = True
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
