"""GPU debug probe: tensor-core decoder forward vs SIMT forward (run with HPS_TC_MMA_MASK=1/9/15)."""
import os, sys, importlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__
__graft_entry__.build()
from tests import parity
hps = importlib.import_module("hierarchical-point-slam_b200")
DEV = "cuda:0"
kind = sys.argv[1] if len(sys.argv) > 1 else "col"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
scene = parity.load("scene_small.pt")
_, _, dec, _ = parity.product(scene, DEV)
m = getattr(dec, "geo_decoder_fine" if kind == "geo" else "color_decoder_fine")
g = torch.Generator().manual_seed(P)
p = (torch.randn(P, 3, generator=g) * 1.5 + torch.tensor([0.0, 0.0, -2.0])).to(DEV)
c = (torch.randn(P, 32, generator=g) * 0.1).to(DEV)
with torch.no_grad():
    hps.ops.USE_TC = False
    ref = m(p, c)
    torch.cuda.synchronize()
    hps.ops.USE_TC = True
    out = m(p, c)
    torch.cuda.synchronize()
print("mask", os.environ.get("HPS_TC_MMA_MASK"), kind, P, "rel err", parity.rel(out, ref), "max ref", float(ref.abs().max()),
      "nan", int(torch.isnan(out).sum()))
print(out[:3].cpu(), ref[:3].cpu())
err = (out - ref).abs().reshape(P, -1).amax(1)
nt = (P + 127) // 128
bad = [t for t in range(nt) if float(err[t * 128:(t + 1) * 128].max()) > 1e-3]
print("tiles", nt, "bad tiles", len(bad), bad[:40], "...", bad[-10:])
if bad:
    t = bad[0]
    e = err[t * 128:(t + 1) * 128]
    print("rows bad in first bad tile:", (e > 1e-3).nonzero().flatten().tolist()[:64])
from oracle import hps_oracle as O
prm = scene["decoders"]["geo_decoder_fine" if kind == "geo" else "color_decoder_fine"]
with torch.no_grad():
    o = O.geo_mlp(p.cpu(), c.cpu(), prm) if kind == "geo" else O.col_mlp(p.cpu(), c.cpu(), prm, None)
def tiles_bad(x):
    e = (x.cpu() - o).abs().reshape(P, -1).amax(1)
    return [t for t in range(nt) if float(e[t * 128:(t + 1) * 128].max()) > 1e-3]
print("vs oracle: SIMT bad tiles", len(tiles_bad(ref)), tiles_bad(ref)[:12], " TC bad tiles", len(tiles_bad(out)), tiles_bad(out)[:12])
if bad and nt > 148 and False:
    t = bad[0]
    a = out[t * 128:(t + 1) * 128].cpu()
    for t2 in (t + 148,):
        if (t2 + 1) * 128 <= P:
            print("out[tile %d] vs ref[tile %d]: %.3e ; vs out[tile %d]: %.3e" % (t, t2, parity.rel(a, ref[t2 * 128:(t2 + 1) * 128]), t2, parity.rel(a, out[t2 * 128:(t2 + 1) * 128])))
    e = (out - ref).abs().reshape(P, -1)[t * 128:(t + 1) * 128]
    print("bad rows count", int((e.amax(1) > 1e-3).sum()), "err per col", e.amax(0).tolist())
    print("row errs (every 8th):", ["%.2g" % x for x in e.amax(1)[::8].tolist()])

e = (out - ref).abs().reshape(P, -1).amax(1)
tm = [float(e[t * 128:(t + 1) * 128].max()) for t in range(nt)]
print("GRID", os.environ.get("HPS_TC_GRID"), "P", P, "tiles", nt, "bad tile ids:", [t for t in range(nt) if tm[t] > 1e-3][:30], "n_bad", sum(1 for x in tm if x > 1e-3))
if bad and nt > 148:
    t, t2 = bad[0], bad[0] + 148
    if (t2 + 1) * 128 <= P:
        sl, sl2 = slice(t * 128, (t + 1) * 128), slice(t2 * 128, (t2 + 1) * 128)
        f = (lambda pp, cc: O.geo_mlp(pp, cc, prm)) if kind == "geo" else (lambda pp, cc: O.col_mlp(pp, cc, prm, None))
        pc, cc = p.cpu(), c.cpu()
        with torch.no_grad():
            combos = {"p1c1": f(pc[sl], cc[sl]), "p1c2": f(pc[sl], cc[sl2]), "p2c1": f(pc[sl2], cc[sl]), "p2c2": f(pc[sl2], cc[sl2])}
        for k, v in combos.items():
            print("out[tile %d] vs oracle(%s): %.3e" % (t, k, parity.rel(out[sl].cpu(), v)))
er = (out - ref).abs().reshape(P, -1).amax(1).cpu()
badrows = (er > 1e-3).nonzero().flatten()
if badrows.numel():
    print("bad rows: count", badrows.numel(), "min", int(badrows.min()), "max", int(badrows.max()), "P-18944 =", P - 18944)
# run TC again (second call) and TC-only fresh process behaviour
with torch.no_grad():
    out2 = m(p, c)
torch.cuda.synchronize()
er2 = (out2 - ref).abs().reshape(P, -1).amax(1).cpu()
print("second TC call bad rows:", int((er2 > 1e-3).sum()))
