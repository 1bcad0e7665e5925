"""Phase timeline of CTA 0 of the tensor-core decoder kernels (hps_tc_trace): run on the GPU box.
usage: python scripts/tc_trace.py [P]"""
import sys

import torch

sys.path.insert(0, ".")
import hps_b200 as hps
from hps_b200 import _lib, ops

P = int(sys.argv[1]) if len(sys.argv) > 1 else 25000
dev = torch.device("cuda:0")
torch.manual_seed(0)
cfg = hps.config.scannet_cfg()
model = hps.get_model(cfg).to(dev)
lib = _lib.load()
buf = torch.zeros(256, dtype=torch.int64, device=dev)


def show(name, a, labels):
    a = [int(x) for x in a]
    t0 = a[0]
    print("--", name)
    prev = t0
    for i, x in enumerate(a):
        if x == 0:
            break
        print("  %-28s t=%8.2f us  (+%6.2f)" % (labels(i), (x - t0) / 1e3, (x - prev) / 1e3))
        prev = x


for kind, dec in ((ops.MLP_COL, model.color_decoder_fine), (ops.MLP_GEO, model.geo_decoder_fine)):
    p = (torch.rand(P, 3, device=dev) * 6 - 3)
    c = (torch.randn(P, 32, device=dev) * 0.3).requires_grad_(True)
    aff = torch.randn(12, device=dev) * 0.1 if kind == ops.MLP_COL else None
    prm = dec.kernel_params(dev)
    for rep in range(3):
        buf.zero_()
        assert lib.hps_tc_trace(_lib.ptr(buf)) == 0
        out = ops.mlp(kind, p, c, prm, aff)
        out.sum().backward()
        torch.cuda.synchronize()
        lib.hps_tc_trace(None)
    h = buf.cpu().tolist()

    def lab_f(i):
        if i == 0: return "tile start"
        if i == 1: return "fourier done"
        if i == 2: return "c done, arrive"
        if i < 13: return ("L%d acc ready" if (i - 3) % 2 == 0 else "L%d epilogue done") % ((i - 3) // 2)
        return "out ready" if i == 13 else ("out written" if i == 14 else "next tile ...")
    show("kind %d forward epilogue warp (tile 0 then tile 1)" % kind, h[0:64], lab_f)
    show("kind %d forward MMA warp (a_ready seen / MMAs issued)" % kind, h[64:128], lambda i: "L%d %s" % ((i // 2) % 6, "operands ready" if i % 2 == 0 else "issued"))

    def lab_b(i):
        if i == 0: return "tile start"
        if i == 1: return "output layer done, arrive"
        return "step %d" % i
    show("kind %d dgrad epilogue warp" % kind, h[128:192], lab_b)
    show("kind %d dgrad MMA warp" % kind, h[192:256], lambda i: "L%d %s" % (4 - (i // 2) % 5, "operands ready" if i % 2 == 0 else "issued"))
