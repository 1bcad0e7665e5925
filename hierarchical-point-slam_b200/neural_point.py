"""Drop-in for the reference's src/neural_point.py: NeuralPointCloud with a device-resident store and an
exact hashed-voxel-grid neighbour index instead of faiss IVFFlat.

Kept from the reference: constructor argument (cfg), the per-level dict state
(`geo_feats`, `col_feats`, `index`, `_pts_num`), and the methods other reference modules call:
`find_neighbors_faiss(pos, step, retrain, is_pts_grad, dynamic_radius, level) -> (D f32 (P,8), I i64 (P,8),
neighbor_num i32 (P,))`, `add_neural_points(...)`, `cloud_pos(level)`, `get_geo_feats/get_col_feats`,
`update_geo_feats/update_col_feats`, `index_ntotal(level)`, `get_index(level)` and the index object's
faiss-like surface (`is_trained`, `train`, `add`, `search`, `ntotal`, `nprobe`).

Changed on purpose: positions live in one (N,3) fp32 device tensor per level (the reference keeps a
Python list of lists and re-tensorises it every iteration, neural_point.py:301, Mapper.py:779); the
search is exact (== the reference with nprobe = nlist) and returns -1 / FLT_MAX in slots outside the
query radius, which the reference always weights with 0 (decoder.py:235-236, SURVEY 7 "hard parts").
"""
from collections import namedtuple

import numpy as np
import torch

import os

from . import ops

# voxel edge as a fraction of the level's largest query radius (see NeuralPointCloud._cell_for)
CELL_PER_RADIUS = float(os.environ.get("HPS_CELL_PER_RADIUS", "0.5"))

NeighborQuery = namedtuple("NeighborQuery", "I D count t_up t_dn")


class CloudPositions(torch.Tensor):
    """(N,3) device tensor that still converts like the reference's list: np.array(x) / x.tolist()."""

    def __array__(self, dtype=None, copy=None):
        a = self.detach().cpu().as_subclass(torch.Tensor).numpy()
        return a.astype(dtype) if dtype is not None else a


def _grown(buf, n, extra, width):
    """Capacity-doubling append buffer: (buffer with room for n + extra rows, whether it was re-allocated)."""
    need = n + extra
    if buf is not None and buf.shape[0] >= need:
        return buf, False
    cap = max(need, 2 * (buf.shape[0] if buf is not None else 0), 1024)
    new = torch.empty((cap, width), dtype=buf.dtype if buf is not None else torch.float32,
                      device=buf.device if buf is not None else None)
    if buf is not None and n:
        new[:n] = buf[:n]
    return new, True


class GridIndex:
    """faiss-like facade over ops.VoxelHashGrid for one level.  Positions live in a capacity-doubling (cap,3) device
    buffer (`positions` = its first ntotal rows); add() appends in place and the next query hashes only the new points
    into the built grid (hps_grid_append) - an insertion costs O(new points), like the reference's index.add
    (neural_point.py:327), not a re-hash and re-allocation of the whole cloud."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.is_trained = False
        self.nprobe = 1                       # accepted and ignored: the search is exact
        self._buf = torch.zeros(0, 3, dtype=torch.float32, device=self.device)
        self._n = 0
        self._grid = None
        self._grid_n = -1
        self._cell = None
        self.generation = 0                   # bumped whenever the position buffer is re-allocated (see storage_generation)
        self.appends = self.rebuilds = 0      # bookkeeping for tests / diagnostics
        self.last_append_status = 0
        self.append_failures = []             # status words of appends that had to fall back to a rebuild
        # Voxel edge.  When set (NeuralPointCloud derives it from the level's largest query radius) no
        # host sync is needed per query; when None it is derived from each batch's largest radius.
        self.cell_hint = None

    @property
    def _pos(self):
        return self._buf[:self._n]

    @property
    def ntotal(self):
        return int(self._n)

    def train(self, x=None):
        self.is_trained = True

    def add(self, x):
        x = torch.as_tensor(x, dtype=torch.float32, device=self.device).reshape(-1, 3).detach()
        m = x.shape[0]
        if m == 0:
            return
        self._buf, moved = _grown(self._buf, self._n, m, 3)
        if moved:
            self.generation += 1              # (the grid keeps its own copy of the points: it is still appended to)
        self._buf[self._n:self._n + m] = x
        self._n += m

    def reset(self, pos):
        self._buf = torch.as_tensor(pos, dtype=torch.float32, device=self.device).reshape(-1, 3).detach().clone()
        self._n = self._buf.shape[0]
        self.is_trained = True
        self._grid_n = -1
        self.generation += 1

    @property
    def positions(self):
        return self._pos

    @property
    def storage_generation(self):
        """Changes whenever a device address a captured CUDA graph may hold has changed (position buffer re-allocated,
        grid rebuilt): re-capture graphs of iterations that query this index when it does."""
        return (self.generation, self._grid.generation if self._grid is not None else 0)

    def _ensure_grid(self, t_dn):
        """Bring the grid up to date: append the points added since the last query (the usual case) or (re)build - first
        use, a cell / the arrays outgrew their room, or the cell size is badly matched to the query radius."""
        stale = self._grid is None or self._grid_n != self.ntotal
        if self.cell_hint is not None:
            want = float(self.cell_hint)
            rebuild = self._grid is None or want != self._cell
        else:
            r_max = float(torch.sqrt(t_dn.max()).item()) if t_dn.numel() else 1.0   # one scalar sync
            want = max(r_max * 0.5, 1e-4)
            rebuild = self._grid is None or not (0.5 * self._cell <= want <= 2.0 * self._cell)
            if not rebuild:
                want = self._cell
        if stale and not rebuild and 0 < self._grid_n < self.ntotal:
            self.last_append_status = self._grid.append(self._pos, self._grid_n)
            if self.last_append_status == 0:
                self._grid_n = self.ntotal
                self.appends += 1
                return self._grid
            self.append_failures.append(self.last_append_status)
            rebuild = True                    # out of room somewhere: start over with fresh capacities
        if stale or rebuild:
            if self._grid is None:
                self._grid = ops.VoxelHashGrid()
            self._grid.build(self._pos, want)
            self._grid_n, self._cell = self.ntotal, want
            self.rebuilds += 1
        return self._grid

    def grid_for(self, radius):
        """The built grid for a query with the given radius (scalar or tensor)."""
        if self.cell_hint is None:
            n = radius.numel() if torch.is_tensor(radius) else 1
            _, t_dn = ops.radius_thresholds(radius, n, self.device)
            return self._ensure_grid(t_dn)
        return self._ensure_grid(None)

    def radius_search(self, q, radius, want_cand=False):
        """Exact 8-NN limited to `radius` (float or per-query tensor, float64 semantics)."""
        q = q.detach().to(device=self.device, dtype=torch.float32).reshape(-1, 3)
        t_up, t_dn = ops.radius_thresholds(radius, q.shape[0], self.device)
        grid = self._ensure_grid(t_dn)
        res = grid.query(q, t_up, t_dn, want_cand)
        out = NeighborQuery(res[0], res[1], res[2], t_up, t_dn)
        return (out, res[3]) if want_cand else out

    def search(self, x, k=ops.K_NN):
        """Unbounded k-NN (faiss Index.search semantics) by doubling the radius until every query has k
        neighbours or the radius covers the whole cloud.  A compatibility facade for callers that used the faiss index
        directly: NOT on the hot path (find_neighbors_faiss goes through radius_search), one host synchronisation per
        round."""
        assert k == ops.K_NN, "the kernels are built for nn_num = 8"
        x = x.detach().to(device=self.device, dtype=torch.float32).reshape(-1, 3)
        Q = x.shape[0]
        D = torch.full((Q, k), torch.finfo(torch.float32).max, device=self.device)
        I = torch.full((Q, k), -1, dtype=torch.int64, device=self.device)
        if Q == 0 or self.ntotal == 0:
            return D, I
        lo, hi = self._pos.min(0).values, self._pos.max(0).values
        span = float(torch.maximum((hi - lo).norm() + (x - lo).norm(dim=1).max(), torch.tensor(1e-3)).item())
        todo = torch.arange(Q, device=self.device)
        r = max(span / 64.0, 1e-3)
        coarse = None
        while todo.numel():
            use = self
            if self.cell_hint is not None and r > 4.0 * float(self.cell_hint):
                # a ball much wider than this index's voxels would probe (2 r / cell)^3 hash slots per query (millions for a
                # 20 m scene with 8 cm voxels): such rounds go to a private index whose voxel edge follows the radius
                if coarse is None:
                    coarse = GridIndex(self.device)
                    coarse._buf, coarse._n, coarse.is_trained = self._buf, self._n, True
                use = coarse
            nq = use.radius_search(x[todo], r)
            full = (nq.count >= min(k, self.ntotal)) | (r > span)
            done = todo[full]
            D[done], I[done] = nq.D[full], nq.I[full].long()
            todo = todo[~full]
            r *= 2.0
        return D, I


def remap_bilinear(depth, x, y):
    """cv2.remap(depth, x, y, interpolation=cv2.INTER_LINEAR) with the default constant-0 border, for a float32 image and
    float32 positions, as tensor ops on the positions' device (Mapper.py:231-236).  OpenCV samples on a 1/32-pixel
    lattice: sx = cvRound(32 x), pixel sx >> 5 (saturated to int16), fraction (sx & 31) / 32, four float weights
    (1-fy)(1-fx), (1-fy)fx, fy(1-fx), fy fx, summed left to right in float32; taps outside the image count as 0.  Next to
    missing-depth (zero) pixels this differs from exact bilinear weights by centimetres, so it is restated, not
    approximated; bit-exact against cv2 4.13 on the fixture of oracle/make_golden_frustum.py."""
    H, W = depth.shape
    flat = depth.to(torch.float32).reshape(-1)

    def fixed(t):
        t32 = t.to(torch.float32) * 32.0
        bad = ~torch.isfinite(t32) | (t32.abs() >= 2.0 ** 31)
        r = torch.round(torch.where(bad, torch.zeros_like(t32), t32)).to(torch.int64)      # half to even, like cvRound
        r = torch.where(bad, torch.full_like(r, -2 ** 31), r)
        return (r >> 5).clamp(-32768, 32767), (r & 31).to(torch.float32) * (1.0 / 32.0)
    ix, ax = fixed(x)
    iy, ay = fixed(y)

    def tap(yy, xx):
        inside = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        return torch.where(inside, flat[yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)], torch.zeros_like(ax))
    acc = tap(iy, ix) * ((1.0 - ay) * (1.0 - ax))
    acc = acc + tap(iy, ix + 1) * ((1.0 - ay) * ax)
    acc = acc + tap(iy + 1, ix) * (ay * (1.0 - ax))
    return acc + tap(iy + 1, ix + 1) * (ay * ax)


def frustum_select(pos, c2w, depth, H, W, fx, fy, cx, cy, edge):
    """Mapper.get_mask_from_c2w (Mapper.py:194-250) over a (N,3) position tensor, on its device: -> ascending int64
    indices.  The number types follow the reference so that the selection is the same set, not a similar one: w2c is
    the float32 inverse of the float32 pose (:208-209, computed on the host by the same numpy call), the projection runs in float64 (:213-227), the pixel position is
    cast to float32 (:228), the depth image is sampled like cv2.remap does (`remap_bilinear`), zero samples take the
    maximum (:243-244) and `depths + 0.5` is a float32 sum (:246).  Pinned against the reference's own method run with
    the real cv2 (tests/golden/frustum_small.pt; the same ops run here on CPU tensors and on the GPU)."""
    dev = pos.device
    if pos.shape[0] == 0:
        return torch.zeros(0, dtype=torch.long, device=dev)
    # the pose inverse is the reference's own statement, on the host like there (`np.linalg.inv(c2w.cpu().numpy())`, float32
    # LAPACK): 64 bytes each way per call; a device-side inverse differs from it in the last bit, which moves points
    # across cv2's 1/32-pixel lattice
    c2w = np.asarray(torch.as_tensor(c2w).detach().cpu().numpy(), dtype=np.float32)
    if c2w.shape[0] == 3:
        c2w = np.concatenate([c2w, np.array([[0.0, 0.0, 0.0, 1.0]], dtype=np.float32)], 0)
    w2c = torch.from_numpy(np.linalg.inv(c2w).astype(np.float64)).to(dev)
    depth = torch.as_tensor(depth, device=dev)
    cam = pos.to(torch.float64) @ w2c[:3, :3].T + w2c[:3, 3]
    cam = cam * torch.tensor([-1.0, 1.0, 1.0], dtype=torch.float64, device=dev)        # Mapper.py:225 axis flip
    z = cam[:, 2] + 1e-5
    u = ((fx * cam[:, 0] + cx * cam[:, 2]) / z).to(torch.float32)
    v = ((fy * cam[:, 1] + cy * cam[:, 2]) / z).to(torch.float32)
    dep = remap_bilinear(depth, u, v)
    dep = torch.where(dep == 0, dep.max(), dep)
    far = (dep + 0.5).to(torch.float64)                                                # float32 sum, then widened
    mz = -z
    mask = (u < W - edge) & (u > edge) & (v < H - edge) & (v > edge) & (mz >= 0) & (mz <= far)
    return mask.nonzero().flatten()


class NeuralPointCloud(object):
    def __init__(self, cfg):
        self.cfg = cfg
        self.c_dim = cfg["model"]["c_dim"]
        self.device = cfg["mapping"]["device"]
        self.cuda_id = 0
        self.use_dynamic_radius = cfg["use_dynamic_radius"]
        pc = cfg["pointcloud"]
        self.nn_num = pc["nn_num"]
        if self.nn_num != ops.K_NN or self.c_dim != ops.C_DIM:
            raise NotImplementedError("kernels are built for nn_num=8, c_dim=32 (the reference configuration)")
        self.nlist, self.nprobe = pc["nlist"], pc["nprobe"]
        self.radius_add, self.radius_min = pc["radius_add"], pc["radius_min"]
        self.radius_query, self.radius_mesh = pc["radius_query"], pc["radius_mesh"]
        self.add_along_normals = pc["add_along_normals"]
        self.fix_interval_when_add_along_ray = pc["fix_interval_when_add_along_ray"]
        self.N_add = pc["N_add"]
        self.near_end_surface, self.far_end_surface = pc["near_end_surface"], pc["far_end_surface"]
        self.radius_hierarchy = pc["radius_hierarchy"]
        self.N_samples, self.N_surface = cfg["rendering"]["N_samples"], cfg["rendering"]["N_surface"]
        self._pts_num = {}
        self.geo_feats = {}
        self.col_feats = {}
        self.index = {}
        self._geo_buf, self._col_buf = {}, {}     # capacity buffers behind geo_feats / col_feats (add_neural_points)
        self._storage_generation = 0
        self._cloud_normal = {}
        self._input_pos = []
        self._input_rgb = []
        self._input_normal = []
        self._input_normal_cartesian = []
        self.keyframe_dict = []

    # ---- reference accessors ----------------------------------------------------------------
    def cloud_pos(self, level, index=None):
        pos = self.index[level].positions.as_subclass(CloudPositions)
        return pos if index is None else pos[index]

    def frustum_rows(self, level, c2w, depth, H, W, fx, fy, cx, cy, edge):
        """Frustum feature selection on the device (Mapper.get_mask_from_c2w, Mapper.py:194-250): indices (int64, ascending)
        of the level's points that project inside the frame cropped by `edge` pixels, in front of the camera and not
        further than the sampled depth + 0.5 m.  The reference does this on the host in numpy + cv2.remap for every mapped
        frame and level (four times per frame, :506-509); here it is `frustum_select` over the resident positions.
        Feed the result to ops.LiveRows / LiveRows.pair."""
        return frustum_select(self.index[level].positions, c2w, depth, H, W, fx, fy, cx, cy, edge)

    def storage_generation(self, level):
        """A value that changes whenever a device address of this level may have changed - feature tables or position
        buffer re-allocated (capacity doubling, set_cloud, update_*_feats with a whole table), neighbour grid rebuilt.
        Point insertion normally changes NONE of them (everything is appended in place), so an iteration captured in a
        CUDA graph (graphs.GraphedStep) stays valid across insertions; re-capture when this value differs from the one
        seen at capture time."""
        return (self._storage_generation, self.index[level].storage_generation if level in self.index else None)

    def cloud_pos_tensor(self, level):
        """The (N,3) device tensor itself (what Mapper.py:779 / Tracker.py:257 rebuild every time)."""
        return self.index[level].positions

    def cloud_normal(self, level=None):
        return self._cloud_normal[level]

    def input_pos(self):
        return torch.cat(self._input_pos).tolist() if self._input_pos else []

    def input_rgb(self):
        return torch.cat(self._input_rgb).tolist() if self._input_rgb else []

    def input_normal(self):
        """neural_point.py:93 (checkpoint logger, Logger.py:33).  Stays empty: normals are not built (use_normals False)."""
        return self._input_normal

    def input_normal_cartesian(self):
        return self._input_normal_cartesian

    def pts_num(self):
        return self._pts_num

    def set_pts_num(self, value):
        self._pts_num = value

    def index_is_trained(self, level=None):
        """neural_point.py:124-138 address `self.index` as one faiss index although it is a per-level dict (they would
        raise there); here they take the level, or answer for all levels when it is omitted."""
        levels = [level] if level is not None else list(self.index)
        return bool(levels) and all(self.index[l].is_trained for l in levels)

    def index_train(self, xb, level=None):
        assert torch.is_tensor(xb), "use tensor to train FAISS index"
        for l in ([level] if level is not None else list(self.index)):
            self.index[l].train(xb)
        return self.index_is_trained(level)

    def index_set_nprobe(self, value, level=None):
        """Kept for callers that tune faiss; the grid search is exact whatever nprobe says."""
        self.nprobe = value
        for l in ([level] if level is not None else list(self.index)):
            self.index[l].nprobe = value

    def index_get_nprobe(self, level=None):
        return self.index[level].nprobe if level is not None else self.nprobe

    @staticmethod
    def cart2sph(xyz):
        """neural_point.py:181-189: unit normals (N,3) -> (N,2) polar / azimuth angles."""
        out = torch.zeros(xyz.shape[0], 2, device=xyz.device)
        out[:, 0] = torch.atan2(torch.sqrt(xyz[:, 0] ** 2 + xyz[:, 1] ** 2), xyz[:, 2])
        out[:, 1] = torch.atan2(xyz[:, 1], xyz[:, 0])
        return out

    def merge_points(self, keyframe_list):
        """neural_point.py:419-420: a no-op in the reference as well."""

    def set_keyframe_dict(self, value):
        self.keyframe_dict = value

    def get_keyframe_dict(self):
        """neural_point.py:118-119: clones, so that the tracker's copy does not alias the mapper's tensors."""
        return [{key: value.clone() for key, value in keyframe.items()} for keyframe in self.keyframe_dict]

    def get_index(self, level=None):
        return self.index[level]

    def index_ntotal(self, level):
        return self.index[level].ntotal

    def get_device(self):
        return self.device

    def get_c_dim(self):
        return self.c_dim

    def get_radius_query(self):
        return self.radius_query

    def get_radius_add(self):
        return self.radius_add

    def get_geo_feats(self, level=None):
        return self.geo_feats[level]

    def get_col_feats(self, level=None):
        return self.col_feats[level]

    def _update(self, table, feats, indices, level):
        assert torch.is_tensor(feats), "use tensor to update features"
        if indices is not None:
            table[level][indices] = feats.clone().detach()
        else:
            assert feats.shape[0] == table[level].shape[0], "feature shape[0] mismatch"
            table[level] = feats.clone().detach()
            self._storage_generation += 1

    def update_geo_feats(self, feats, indices=None, level=None):
        self._update(self.geo_feats, feats, indices, level)

    def update_col_feats(self, feats, indices=None, level=None):
        self._update(self.col_feats, feats, indices, level)

    # ---- building ------------------------------------------------------------------------------
    def _init_level(self, level):
        self._cloud_normal.setdefault(level, [])
        self._pts_num.setdefault(level, 0)
        if level not in self.index:
            self.index[level] = GridIndex(self.device)
            self.index[level].nprobe = self.nprobe
            self.index[level].cell_hint = self._cell_for(level)
        self.geo_feats.setdefault(level, None)
        self.col_feats.setdefault(level, None)

    def _cell_for(self, level):
        """Voxel edge = half the largest query radius of the level (fine 0.16 m -> 0.08 m cells, mid
        0.96 m -> 0.48 m): a query ball then overlaps at most 5 cells per axis."""
        if self.use_dynamic_radius and level in self.radius_hierarchy:
            r_max = self.cfg["pointcloud"]["radius_query_ratio"] * max(self.radius_hierarchy[level].values())
        else:
            r_max = self.radius_query
        return CELL_PER_RADIUS * float(r_max)

    def set_cloud(self, level, pos, geo_feats=None, col_feats=None):
        """Install a whole level at once (tests / benchmarks / checkpoint restore)."""
        self._init_level(level)
        self.index[level].reset(pos)
        n = self.index[level].ntotal
        self._pts_num[level] = n
        dev = self.index[level].device
        mk = lambda: torch.zeros(n, self.c_dim, device=dev).normal_(mean=0, std=0.1)
        self.geo_feats[level] = mk() if geo_feats is None else geo_feats.to(dev).float().contiguous()
        self.col_feats[level] = mk() if col_feats is None else col_feats.to(dev).float().contiguous()
        self._geo_buf.pop(level, None)
        self._col_buf.pop(level, None)
        self._storage_generation += 1

    def add_neural_points(self, batch_rays_o, batch_rays_d, batch_gt_depth, batch_gt_color, normals=None,
                          train=False, is_pts_grad=False, dynamic_radius=None, level=None, idx=None):
        """Reference neural_point.py:193-334, entirely on the device: unproject, density test with the
        'add' radius, append N_add points per accepted location, extend the feature tables."""
        if idx == 0 or level not in self.index:
            self._init_level(level)
        if normals is not None:
            raise NotImplementedError("use_normals is off in the ScanNet configuration and not built")
        if self.add_along_normals or self.fix_interval_when_add_along_ray:
            raise NotImplementedError("only the depth-proportional add-along-ray mode is built")
        if not batch_rays_o.shape[0]:
            return 0
        index = self.index[level]
        dev = index.device
        mask = batch_gt_depth > 0
        o, d = batch_rays_o[mask].float(), batch_rays_d[mask].float()
        depth, color = batch_gt_depth[mask].float(), (batch_gt_color * 255)[mask]
        if dynamic_radius is not None:
            dynamic_radius = dynamic_radius.reshape(-1)[mask] if dynamic_radius.numel() == mask.numel() \
                else dynamic_radius
        pts_gt = (o[..., None, :] + d[..., None, :] * depth[..., None, None]).reshape(-1, 3)
        self._input_pos.append(pts_gt.detach())
        self._input_rgb.append(color.detach())
        t = torch.linspace(0.0, 1.0, steps=self.N_add, device=dev)
        gds = depth.unsqueeze(-1).repeat(1, self.N_add)
        z_vals = self.near_end_surface * gds * (1.0 - t) + self.far_end_surface * gds * t \
            if self.N_add > 1 else depth.unsqueeze(-1)
        keep = torch.ones(pts_gt.shape[0], dtype=torch.bool, device=dev)
        if index.is_trained and index.ntotal > 0:
            _, _, n_near = self.find_neighbors_faiss(pts_gt, step="add", is_pts_grad=is_pts_grad,
                                                     dynamic_radius=dynamic_radius, level=level)
            keep = n_near == 0
        pts = (o[..., None, :] + d[..., None, :] * z_vals[..., :, None])[keep].reshape(-1, 3)
        index.train()
        index.add(pts)
        self._pts_num[level] += pts.shape[0]
        # feature rows of the new points, N(0, 0.1^2) (neural_point.py:305-317), appended IN PLACE: the tables live in
        # capacity-doubling buffers (the reference re-allocates both (N,32) tables with torch.cat on every insertion)
        k = pts.shape[0]
        for table, bufs in ((self.geo_feats, self._geo_buf), (self.col_feats, self._col_buf)):
            cur = table[level]
            n_old = 0 if cur is None else cur.shape[0]
            buf = bufs.get(level)
            if cur is not None and (buf is None or cur.data_ptr() != buf.data_ptr() or buf.shape[0] < n_old):
                buf = None                    # the table was replaced from outside (update_*_feats): adopt it
            if buf is None:
                buf = torch.empty((max(2 * (n_old + k), 1024), self.c_dim), dtype=torch.float32, device=dev)
                if n_old:
                    buf[:n_old] = cur
                self._storage_generation += 1
            else:
                buf, moved = _grown(buf, n_old, k, self.c_dim)
                self._storage_generation += int(moved)
            bufs[level] = buf
            buf[n_old:n_old + k].normal_(mean=0, std=0.1)
            table[level] = buf[:n_old + k]
        return torch.sum(keep)

    # ---- neighbour queries -----------------------------------------------------------------------
    def _radius_for(self, step, is_pts_grad):
        assert step in ["add", "query", "mesh"]
        if step == "query":
            return self.radius_query
        if step == "add":
            return self.radius_min if is_pts_grad else self.radius_add
        return self.radius_mesh

    def query_neighbors(self, pos, step="query", is_pts_grad=False, dynamic_radius=None, level=None,
                        want_cand=False):
        """Kernel-facing form: NeighborQuery(I i32 (P,8), D, count, t_up, t_dn)."""
        radius = dynamic_radius if dynamic_radius is not None else self._radius_for(step, is_pts_grad)
        if dynamic_radius is not None:
            assert pos.shape[0] == dynamic_radius.shape[0], "shape mis-match for input points and dynamic radius"
        return self.index[level].radius_search(pos, radius, want_cand)

    def find_neighbors_faiss(self, pos, step="add", retrain=False, is_pts_grad=False, dynamic_radius=None,
                             level=None):
        """Reference signature (neural_point.py:366-417)."""
        q = self.query_neighbors(pos, step, is_pts_grad, dynamic_radius, level)
        return q.D, q.I.long(), q.count

    def sample_near_pcl(self, rays_o, rays_d, near, far, num, dynamic_r_query=None, level=None):
        """For rays without a depth reading: probe 25 positions between near and far, and sample `num` depths between
        the FIRST TWO probes that have a neighbour within the query radius (neural_point.py:443-495; the reference
        takes entries [0] and [1] of the hit list, not the first and the last).  Rays with fewer than two hits keep
        linspace(near, far, num) and are flagged.  -> (z_vals (n_rays, num) float32, invalid (n_rays,) bool).

        One neighbour-count query on the device replaces the reference's 65,000-point faiss batches, numpy
        round trip and per-ray Python list; the float64 np.linspace arithmetic (start + i*step, last = stop) is kept.
        dynamic_r_query: None (the fixed query radius - the only form the reference accepts, its dict form trips the
        shape assert at neural_point.py:411), or a radius per ray / per probe."""
        rays_o, rays_d = rays_o.reshape(-1, 3).float(), rays_d.reshape(-1, 3).float()
        n_rays, probes, dev = rays_d.shape[0], 25, rays_o.device
        far32 = torch.as_tensor(far, dtype=torch.float32, device=dev).reshape(())
        idx = torch.arange(probes, device=dev, dtype=torch.float32)
        step = (far32 - near) / (probes - 1)                            # torch.linspace's float32 arithmetic
        z_probe = torch.where(idx < probes // 2, near + step * idx, far32 - step * (probes - 1 - idx))
        pts = (rays_o[:, None, :] + rays_d[:, None, :] * z_probe[:, None]).reshape(-1, 3)
        radius = dynamic_r_query[level] if isinstance(dynamic_r_query, dict) else dynamic_r_query
        if radius is None:
            radius = self.radius_query
        else:
            radius = radius.reshape(-1)
            if radius.numel() == n_rays:
                radius = radius.repeat_interleave(probes)
            assert radius.numel() == pts.shape[0], "shape mis-match for input points and dynamic radius"
        hit = (self.index[level].radius_search(pts, radius).count > 0).reshape(n_rays, probes)
        invalid = hit.sum(-1) < 2
        rank = hit.cumsum(-1)
        first = ((rank == 1) & hit).int().argmax(-1)
        second = ((rank == 2) & hit).int().argmax(-1)

        def linspace64(start, stop, count):                               # np.linspace, float64, endpoint=True
            i = torch.arange(count, device=dev, dtype=torch.float64)
            out = i * ((stop - start) / (count - 1))[..., None] + start[..., None]
            out[..., -1] = stop
            return out

        near64 = torch.full((), float(near), dtype=torch.float64, device=dev)
        far64 = far32.double()
        section = linspace64(near64, far64, probes)
        z = linspace64(section[first], section[second], num)
        z = torch.where(invalid[:, None], linspace64(near64, far64, num)[None, :], z)
        return z.float(), invalid
