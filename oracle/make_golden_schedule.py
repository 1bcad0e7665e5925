"""TEST INFRASTRUCTURE ONLY.  Generate tests/golden/mapper_schedule.json by EXECUTING the reference's own statements for
the mapper's stage selection and learning-rate rewrite (src/Mapper.py:605-612 iteration counts, :633-666 stage /
learning rates / BA pose learning rate).  Those statements sit in the middle of `Mapper.optimize_map` and cannot be
called on their own, so this script reads the lines from the reference file where it lies (/root/reference, never
copied into the repo), dedents them and runs them against stand-ins for `self`, `cfg` and `optimizer`, using the
reference's own yaml files for the numbers.

Run in the build container only:   python oracle/make_golden_schedule.py
The fixture is committed; tests/test_host_logic.py compares hps_b200.mapper_optim against it.
"""
import json
import os
import textwrap
import types

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("HPS_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")


def ref_lines(first, last):
    with open(os.path.join(REF, "src", "Mapper.py")) as f:
        lines = f.readlines()
    return textwrap.dedent("".join(lines[first - 1:last]))


def merged_mapping_cfg():
    """configs/point_slam.yaml with configs/ScanNet/scannet.yaml on top (config.py update_recursive), mapping section."""
    with open(os.path.join(REF, "configs", "point_slam.yaml")) as f:
        base = yaml.full_load(f)
    with open(os.path.join(REF, "configs", "ScanNet", "scannet.yaml")) as f:
        top = yaml.full_load(f)
    m = dict(base["mapping"])
    m.update(top.get("mapping", {}))
    return m


def main():
    counts_src = ref_lines(605, 612)
    loop_src = ref_lines(633, 666)
    assert counts_src.lstrip().startswith("if self.more_iters_when_adding"), counts_src[:80]
    assert loop_src.lstrip().startswith("if joint_iter <="), loop_src[:80]
    counts_code = compile(counts_src, "Mapper.py:605-612", "exec")
    loop_code = compile(loop_src, "Mapper.py:633-666", "exec")
    mapping = merged_mapping_cfg()
    cfg = {"mapping": mapping}
    cases = []
    grid = [
        # (num_joint_iters, init, BA, idx, n_img, color_refine, more_iters, frame_pts_add, geo_ratio, mid_ratio)
        (mapping["iters"], False, False, 5, 100, False, False, 0, mapping["geo_iter_ratio"], mapping["mid_iter_ratio"]),
        (mapping["iters_first"], True, False, 0, 100, False, False, 0, mapping["geo_iter_ratio"], mapping["mid_iter_ratio"]),
        (mapping["iters"], False, True, 5, 100, False, False, 0, mapping["geo_iter_ratio"], mapping["mid_iter_ratio"]),
        (mapping["iters"], False, False, 99, 100, True, False, 0, mapping["geo_iter_ratio"], mapping["mid_iter_ratio"]),
        (300, False, True, 7, 100, False, True, 450, 0.4, 0.5),
        (300, False, False, 7, 100, False, True, 10, 0.4, 0.35),
        (37, False, True, 3, 100, False, True, 2000, 0.25, 0.6),
    ]
    for n_iters, init, BA, idx, n_img, color_refine, more, pts_add, geo_ratio, mid_ratio in grid:
        self = types.SimpleNamespace(more_iters_when_adding=more, min_iter_ratio=mapping["min_iter_ratio"],
                                     mid_iter_ratio=mid_ratio, geo_iter_ratio=geo_ratio,
                                     geo_iter_first=mapping["geo_iter_first"], n_img=n_img, color_refine=color_refine,
                                     BA=BA, BA_cam_lr=mapping["BA_cam_lr"], stage=None)
        optimizer = types.SimpleNamespace(param_groups=[{"lr": 0} for _ in range(7)])
        ns = dict(self=self, cfg=cfg, optimizer=optimizer, np=np, idx=idx, init=init, color_refine=color_refine,
                  num_joint_iters=n_iters, frame_pts_add=pts_add)
        exec(counts_code, ns)
        total = int(ns["num_joint_iters"])
        rows = []
        for joint_iter in range(total):
            ns["joint_iter"] = joint_iter
            exec(loop_code, ns)
            rows.append([self.stage] + [float(g["lr"]) for g in optimizer.param_groups[:6]])
        runs = []                                   # run-length encoded: [first_iter, n_iters, stage, six lrs]
        for j, r in enumerate(rows):
            if runs and runs[-1][2:] == r:
                runs[-1][1] += 1
            else:
                runs.append([j, 1] + r)
        rows = runs
        cases.append(dict(num_joint_iters_in=n_iters, init=init, BA=BA, idx=idx, n_img=n_img, color_refine=color_refine,
                          more_iters_when_adding=more, frame_pts_add=pts_add, geo_iter_ratio=geo_ratio,
                          mid_iter_ratio=mid_ratio, num_joint_iters=total, num_mid_iters=int(ns["num_mid_iters"]),
                          num_fine_iters=int(ns["num_fine_iters"]), rows=rows))
    out = dict(mapping_cfg={k: mapping[k] for k in ("iters", "iters_first", "geo_iter_ratio", "mid_iter_ratio", "geo_iter_first",
                                                    "min_iter_ratio", "BA_cam_lr", "init", "stage")}, cases=cases)
    path = os.path.join(GOLD, "mapper_schedule.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes;", [c["num_joint_iters"] for c in cases])


if __name__ == "__main__":
    main()
