"""TEST INFRASTRUCTURE ONLY.  Dump the reference's EFFECTIVE ScanNet configuration - src/config.py load_config of
configs/ScanNet/scene0181.yaml over configs/point_slam.yaml, run by the imported reference - for the sections the hot
path reads, to tests/golden/scannet_cfg.json.  tests/test_abi_and_host.py requires hps_b200.scannet_cfg() to carry the
same value for every key it defines (the product ships the numbers as plain data, not the yaml files).

Run in the build container only:   python oracle/make_golden_cfg.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402


def main():
    ref = ref_import.load_reference()
    cfg = ref_import.load_cfg(ref)
    keep = {k: cfg[k] for k in ("pointcloud", "rendering", "model", "mapping", "tracking", "cam", "data")}
    for k in ("occupancy", "use_normals", "use_view_direction", "use_dynamic_radius", "use_bound", "pts_along_ray"):
        keep[k] = cfg.get(k)
    path = os.path.join(ROOT, "tests", "golden", "scannet_cfg.json")
    with open(path, "w") as f:
        json.dump(keep, f, indent=1, sort_keys=True, default=str)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
