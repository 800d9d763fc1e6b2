"""Drop-in for the Tracker of the reference's predict.py (reference predict.py:127-296): same
constructor, attributes and on_track signature, so the reference's sequence drivers
(predict.py:299-624) and predict_ros.py:59 can call it unchanged -- with the per-frame work
(crop, depth clip, normalise, 17-conv network, pose update) running as one stream of libse3tn
kernels instead of numpy + torch.nn.

Differences, all additive:
  * on_track(..., rgbA=None, depthA=None): the rendered previous view may be passed in.  The
    OpenGL renderers (vispy_renderer.py / offscreen_renderer.py) are out of scope; when they are
    importable the Tracker uses them exactly as the reference does, otherwise rgbA/depthA (or a
    `renderer=` object with render_window(ob2cam)) must be supplied.
  * the second, visualisation-only render + cv2.imshow (predict.py:284-290) only happens with
    show=True.
  * on_track_batch(): N independent tracks of one frame in one batched launch sequence (the
    reference is batch 1, F15).
"""
import os
import numpy as np
import torch

from .engine import Engine
from .se3_tracknet import Se3TrackNet
from .datasets import TrackDataset
from . import Utils as U


class PointCloud:
    """Minimal stand-in for the open3d point cloud the reference keeps in Tracker.object_cloud
    (predict.py:131-133): callers only read `.points`."""
    def __init__(self, points):
        self.points = np.asarray(points, dtype=np.float64)

    def voxel_down_sample(self, voxel_size):
        # open3d semantics: points are bucketed on a grid anchored at (min_bound - voxel/2) and each
        # occupied voxel is replaced by the mean of its points
        pts = self.points
        origin = pts.min(0) - voxel_size * 0.5
        keys = np.floor((pts - origin) / voxel_size).astype(np.int64)
        _, inv = np.unique(keys, axis=0, return_inverse=True)
        inv = inv.reshape(-1)
        sums = np.zeros((inv.max() + 1, 3)); np.add.at(sums, inv, pts)
        cnt = np.bincount(inv).astype(np.float64)[:, None]
        return PointCloud(sums / cnt)


def load_vertices(model_path, merge=True):
    """Vertices of a .ply (ascii / binary little endian) or .obj mesh, as trimesh.load(..).vertices gives them to the
    reference (predict.py:131).  trimesh loads with process=True, which MERGES duplicate vertices (positions equal to its
    tol.merge = 1e-8); duplicates would otherwise weigh twice in the voxel means of voxel_down_sample and move the
    convex-hull diameter -> object_width -> crop window.  merge=True reproduces that (first occurrence kept, file order)."""
    ext = os.path.splitext(model_path)[1].lower()
    if ext == '.obj':
        v = [list(map(float, l.split()[1:4])) for l in open(model_path) if l.startswith('v ')]
        pts = np.asarray(v, dtype=np.float64)
    elif ext == '.ply':
        pts = _ply_vertices(model_path)
    else:
        raise ValueError('unsupported mesh format: ' + model_path)
    if merge and len(pts):
        key = np.round(pts / 1e-8).astype(np.int64)
        _, first = np.unique(key, axis=0, return_index=True)
        pts = pts[np.sort(first)]
    return pts


def _ply_vertices(model_path):
    with open(model_path, 'rb') as f:
        fmt, nvert, props, in_vertex = None, 0, [], False
        while True:
            line = f.readline().decode('ascii', 'replace').strip()
            if line.startswith('format'):
                fmt = line.split()[1]
            elif line.startswith('element'):
                in_vertex = line.split()[1] == 'vertex'
                if in_vertex:
                    nvert = int(line.split()[2])
            elif line.startswith('property') and in_vertex:
                props.append((line.split()[1], line.split()[-1]))
            elif line == 'end_header':
                break
        names = [p[1] for p in props]
        ix = [names.index(a) for a in 'xyz']
        if fmt == 'ascii':
            data = np.loadtxt(f, max_rows=nvert, ndmin=2)
            return data[:, ix].astype(np.float64)
        np_t = {'float': '<f4', 'float32': '<f4', 'double': '<f8', 'float64': '<f8', 'uchar': 'u1', 'uint8': 'u1',
                'char': 'i1', 'int': '<i4', 'int32': '<i4', 'uint': '<u4', 'short': '<i2', 'ushort': '<u2'}
        dt = np.dtype([(n, np_t[t]) for t, n in props])
        data = np.frombuffer(f.read(dt.itemsize * nvert), dtype=dt, count=nvert)
        return np.stack([data['x'], data['y'], data['z']], 1).astype(np.float64)


def compute_obj_max_width(points):
    """Convex-hull diameter in mm (reference Utils.py:101-105, 450-451)."""
    from scipy.spatial import ConvexHull, distance_matrix
    hull = points[ConvexHull(points).vertices]
    return float(np.max(distance_matrix(hull, hull))) * 1000


def _as_numpy_pose(p):
    return np.ascontiguousarray(p, dtype=np.float64)


def crop_windows_union(poses, K, object_width, H, W, margin=2):
    """Bounding rectangle (y0, y1, x0, x1), clipped to the frame, of the crop windows compute_bbox gives these poses (reference
    Utils.py:302-316, same float64 arithmetic and np.round) plus a safety margin; None if a pose is degenerate (then upload everything)."""
    import math
    poses = np.asarray(poses, dtype=np.float64).reshape(-1, 4, 4)
    fx, fy, cx, cy = float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])
    y0, y1, x0, x1 = H, 0, W, 0
    for i in range(len(poses)):                    # plain Python floats: IEEE double like numpy's, round() is half-to-even like np.round
        w = float(object_width if np.ndim(object_width) == 0 else object_width[i])
        ox, oy, oz = float(poses[i, 0, 3]) * 1000.0, float(poses[i, 1, 3]) * 1000.0, float(poses[i, 2, 3]) * 1000.0
        if oz == 0.0 or not all(math.isfinite(v) for v in (ox, oy, oz, w)):
            return None
        us = (round((ox - w / 2) * fx / oz + cx), round((ox + w / 2) * fx / oz + cx))
        vs = (round((oy - w / 2) * fy / oz + cy), round((oy + w / 2) * fy / oz + cy))
        y0 = min(y0, min(vs) - margin); y1 = max(y1, max(vs) + margin)
        x0 = min(x0, min(us) - margin); x1 = max(x1, max(us) + margin)
    y0, y1, x0, x1 = max(y0, 0), min(y1, H), max(x0, 0), min(x1, W)
    if y1 <= y0 or x1 <= x0:
        return (0, 0, 0, 0)                        # every window lies outside the frame: nothing to upload
    return (y0, y1, x0, x1)


class Tracker:
    def __init__(self, dataset_info, images_mean, images_std, ckpt_dir, model_path=None, trans_normalizer=0.03,
                 rot_normalizer=5 * np.pi / 180, engine=None, weight_id=0, renderer=None, precision='bf16x3', max_batch=64):
        self.dataset_info = dataset_info
        self.image_size = (dataset_info['resolution'], dataset_info['resolution'])
        if self.image_size[0] != 176:
            raise NotImplementedError('libse3tn is built for the reference resolution of 176 (dataset_info.yml:15)')
        self.object_cloud = None
        if model_path is not None:
            self.object_cloud = PointCloud(load_vertices(model_path)).voxel_down_sample(voxel_size=0.005)
        if 'object_width' not in dataset_info:
            if self.object_cloud is None:
                raise ValueError("need model_path or dataset_info['object_width']")
            w = compute_obj_max_width(np.asarray(self.object_cloud.points))
            self.object_width = w + dataset_info['boundingbox'] / 100 * w
        else:
            self.object_width = dataset_info['object_width']
        self.mean = images_mean
        self.std = images_std
        cam_cfg = dataset_info['camera']
        self.K = np.array([cam_cfg['focalX'], 0, cam_cfg['centerX'], 0, cam_cfg['focalY'], cam_cfg['centerY'], 0, 0, 1]).reshape(3, 3)

        if isinstance(ckpt_dir, dict):
            checkpoint = ckpt_dir if 'state_dict' in ckpt_dir else {'state_dict': ckpt_dir}
        else:
            checkpoint = torch.load(ckpt_dir, map_location='cpu')
        self.engine = engine if engine is not None else Engine(max_batch=max_batch)
        self.weight_id = weight_id
        self.precision = precision
        self.model = Se3TrackNet(image_size=self.image_size[0], engine=self.engine, weight_id=weight_id, precision=precision)
        self.model.load_state_dict(checkpoint['state_dict'])
        self.model = self.model.cuda()
        self.model.eval()
        self.engine.set_stats(np.asarray(images_mean), np.asarray(images_std), weight_id)
        U.set_engine(self.engine)

        # Input A.  renderer='cuda' (or nothing, with a .ply model that has normals + colours): the CUDA rasteriser
        # (csrc/render.cu) -- no OpenGL, all tracks in one launch.  Otherwise an object with render_window(ob2cam), or the
        # reference's own OpenGL renderers when they are importable.
        # dataset_info['renderer'] == 'pyrenderer' selects the reference's pyrender producer (predict.py:161-164): the rasteriser's
        # unlit full-camera-image mode followed by crop_bbox; anything else its vispy producer.
        pyr = dataset_info.get('renderer') == 'pyrenderer'
        if renderer == 'cuda' or (renderer is None and model_path is not None and str(model_path).lower().endswith(('.ply', '.obj') if pyr else '.ply')):
            from .cuda_renderer import CudaRenderer
            try:
                renderer = CudaRenderer(model_path, self.K, self.engine, self.object_width, mesh_id=weight_id,
                                        mode='pyrender' if pyr else 'vispy', image_hw=(cam_cfg['height'], cam_cfg['width']) if pyr else None)
            except ValueError:
                if renderer == 'cuda':
                    raise
                renderer = None                                    # e.g. a vertices-only ply: fall through to the GL renderers
        self.renderer = renderer if renderer is not None else self._try_reference_renderer(model_path, cam_cfg)
        self._np_bufs = {}
        self._pin_busy = False
        self.prev_rgb = None
        self.prev_depth = None
        self.frame_cnt = 0
        self.errs = []
        self.trans_normalizer = trans_normalizer
        self.rot_normalizer = rot_normalizer
        self.dataset = TrackDataset('', 'eval', images_mean, images_std, None, None, None, dataset_info,
                                    trans_normalizer=trans_normalizer, rot_normalizer=rot_normalizer,
                                    engine=self.engine, weight_id=weight_id, precision=precision)
        self.dataset._stats_set = True

    # ------------------------------------------------------------------ renderer glue (out-of-scope producer)
    def _try_reference_renderer(self, model_path, cam_cfg):
        """The reference picks pyrender or vispy by dataset_info['renderer'] (predict.py:161-182).
        Both are OpenGL stacks outside this package; use them if the user's environment has them."""
        if model_path is None:
            return None
        try:
            if self.dataset_info.get('renderer') == 'pyrenderer':
                from offscreen_renderer import Renderer            # reference module, if on sys.path
            else:
                from vispy_renderer import VispyRenderer           # reference module, if on sys.path
        except ImportError:
            return None                                            # no OpenGL stack in this environment: rgbA/depthA must be passed in
        # the renderer module IS there: a failure to construct it (no GL context, bad mesh) is the user's to see
        if self.dataset_info.get('renderer') == 'pyrenderer':
            return Renderer([model_path], self.K, cam_cfg['height'], cam_cfg['width'])
        return VispyRenderer(model_path, self.K, H=self.dataset_info['resolution'], W=self.dataset_info['resolution'])

    def render_window(self, ob2cam):
        """rgb u8 (176,176,3), depth u16 mm (176,176) of the model at `ob2cam` inside the crop window
        (reference predict.py:193-215)."""
        r = self.renderer
        if r is None:
            raise RuntimeError('no renderer available: pass rgbA/depthA to on_track, or renderer= to Tracker')
        if hasattr(r, 'render_window'):
            return r.render_window(ob2cam)
        glcam_in_cvcam = np.diag([1.0, -1.0, -1.0, 1.0])
        if hasattr(r, 'update_cam_mat'):                           # vispy-style
            bbox = U.compute_bbox(ob2cam, self.K, self.object_width, scale=(1000, -1000, 1000))
            ob2cam_gl = np.linalg.inv(glcam_in_cvcam).dot(ob2cam)
            r.update_cam_mat(self.K, np.min(bbox[:, 1]), np.max(bbox[:, 1]), np.max(bbox[:, 0]), np.min(bbox[:, 0]))
            return r.render_image(ob2cam_gl)
        bbox = U.compute_bbox(ob2cam, self.K, self.object_width, scale=(1000, 1000, 1000))
        rgb, depth = r.render([ob2cam])
        return U.crop_bbox(rgb, (depth * 1000).astype(np.uint16), bbox, self.image_size)

    # ------------------------------------------------------------------ the hot path
    def on_track(self, prev_pose, current_rgb, current_depth, gt_A_in_cam=None, gt_B_in_cam=None, debug=False, samples=1,
                 rgbA=None, depthA=None, show=False):
        """One frame, one object (reference predict.py:217-296) -> new 4x4 float64 pose."""
        A_in_cam = _as_numpy_pose(prev_pose).copy()
        if rgbA is None or depthA is None:
            rgbA, depthA = self.render_window(A_in_cam)
        out = self.on_track_batch(A_in_cam[None], current_rgb, current_depth,
                                  np.ascontiguousarray(rgbA, dtype=np.uint8)[None],
                                  np.ascontiguousarray(depthA).astype(np.uint16)[None])
        final_estimate = out[0]
        self.prev_rgb = current_rgb
        self.prev_depth = current_depth
        if show:
            import cv2
            pred_color, _ = self.render_window(final_estimate)
            cv2.imshow('AB', pred_color[..., ::-1])
            cv2.waitKey(1)
        self.frame_cnt += 1
        return final_estimate

    def on_track_batch(self, prev_poses, current_rgb, current_depth, rgbA=None, depthA=None, weight_ids=None, object_width=None):
        """N independent tracks of ONE frame -> (N,4,4) float64.  rgbA / depthA None: rendered on the device by the CUDA
        rasteriser (needs a CudaRenderer; per-track models follow weight_ids).

        numpy inputs   -> numpy result (synchronous, like the reference's on_track).
        CUDA tensors   -> CUDA tensor, nothing is synchronised.
        CPU tensors    -> CUDA tensor; the host->device copies run on a side stream into double-buffered
                          staging, so the uploads of call k overlap the kernels of call k-1 (pinned memory
                          makes them truly asynchronous).  Nothing is synchronised."""
        dev = self.engine.device
        as_numpy = not torch.is_tensor(prev_poses)
        render = rgbA is None or depthA is None
        if render and not hasattr(self.renderer, 'render_batch'):
            raise RuntimeError('on_track_batch without rgbA/depthA needs the CUDA renderer (Tracker(renderer="cuda", model_path=*.ply))')
        if (not render and all(isinstance(x, np.ndarray) for x in (current_rgb, current_depth, rgbA, depthA)) and not torch.is_tensor(prev_poses)
                and not torch.is_tensor(weight_ids) and not torch.is_tensor(object_width) and os.environ.get('SE3TN_HOST_CALL', '1') != '0'):
            # numpy in, numpy out -- the reference's own calling pattern: ONE library call stages the crop-window rectangle of the
            # frame, the poses and input A through pinned memory, replays the step's graph and hands the poses back
            c = lambda a, dt: a if (a.dtype == dt and a.flags['C_CONTIGUOUS']) else np.ascontiguousarray(a).astype(dt, copy=False)
            poses_h = np.ascontiguousarray(prev_poses, dtype=np.float64).reshape(-1, 4, 4)
            n = len(poses_h)
            ow_h = np.full(n, float(self.object_width)) if object_width is None else np.ascontiguousarray(np.broadcast_to(np.asarray(object_width, dtype=np.float64), (n,)))
            wh = None
            if weight_ids is not None:
                wh = np.ascontiguousarray(weight_ids, dtype=np.int32)
            elif self.weight_id != 0:
                wh = np.full(n, self.weight_id, dtype=np.int32)
            return self.engine.track_host(c(current_rgb, np.uint8), c(current_depth, np.uint16), self.K, poses_h, ow_h, c(rgbA, np.uint8), c(depthA, np.uint16),
                                          self.trans_normalizer, self.rot_normalizer, weight_ids=wh, precision=self.precision)
        staged = not render and all(torch.is_tensor(x) and not x.is_cuda for x in (prev_poses, current_rgb, current_depth, rgbA, depthA))

        def up(x, dt, slot=None):
            if torch.is_tensor(x):
                return x.to(dev, dt).contiguous()
            a = np.ascontiguousarray(x)
            if dt == torch.uint16 and a.dtype != np.uint16:
                a = a.astype(np.uint16)
            src = torch.from_numpy(a)
            if slot is None:
                return src.to(dev).to(dt)
            # numpy inputs land in persistent device buffers (one per argument and shape): the step's CUDA graph is keyed by its
            # device pointers, so stable addresses mean every frame after the first is one graph launch
            key = (slot, tuple(a.shape), dt)
            buf = self._np_bufs.get(key)
            if buf is None:
                buf = self._np_bufs[key] = torch.empty(a.shape, dtype=dt, device=dev)
            buf.copy_(src if src.dtype == dt else src.to(dt))
            return buf

        if staged:
            poses, rgb_d, depth_d, rgbA_d, depthA_d = self._stage_uploads(prev_poses, current_rgb, current_depth, rgbA, depthA)
        else:
            poses = up(prev_poses, torch.float64, 'poses')
            win = None
            if (os.environ.get('SE3TN_WINDOW_UPLOAD', 'pinned') != 'off' and not torch.is_tensor(current_rgb) and not torch.is_tensor(current_depth) and not torch.is_tensor(prev_poses) and object_width is None
                    and len(prev_poses) <= 4 and current_rgb.dtype == np.uint8 and current_depth.dtype == np.uint16
                    and current_rgb.flags['C_CONTIGUOUS'] and current_depth.flags['C_CONTIGUOUS']):
                # a few objects: K0 only reads the frame inside their crop windows -> upload that rectangle, not the whole 1.5 MB frame
                win = crop_windows_union(prev_poses, self.K, self.object_width, current_depth.shape[0], current_depth.shape[1])
            if win is not None and (win[1] - win[0]) * (win[3] - win[2]) * 2 < current_depth.size:
                rk, dk = ('rgb', tuple(current_rgb.shape), torch.uint8), ('depth', tuple(current_depth.shape), torch.uint16)
                for k2, dt in ((rk, torch.uint8), (dk, torch.uint16)):
                    if k2 not in self._np_bufs:
                        self._np_bufs[k2] = torch.zeros(k2[1], dtype=dt, device=dev)
                rgb_d, depth_d = self._np_bufs[rk], self._np_bufs[dk]
                # through pinned staging (same geometry): a 2-D copy from pageable memory is staged row by row by the driver,
                # from pinned memory it is one strided DMA
                pk = ('pin', tuple(current_rgb.shape))
                if os.environ.get('SE3TN_WINDOW_UPLOAD', 'pinned') == 'pageable':
                    self.engine.upload_frame_window(current_rgb, current_depth, rgb_d, depth_d, *win)
                    pk = None
                elif pk not in self._np_bufs:
                    self._np_bufs[pk] = (torch.empty(current_rgb.shape, dtype=torch.uint8).pin_memory().numpy(),
                                         torch.empty(current_depth.shape, dtype=torch.uint16).pin_memory().numpy())
                if pk is not None:
                    pin_rgb, pin_depth = self._np_bufs[pk]
                    y0, y1, x0, x1 = win
                    if self._pin_busy:
                        torch.cuda.current_stream(dev).synchronize()          # the previous call's DMA may still be reading the staging
                    np.copyto(pin_rgb[y0:y1, x0:x1], current_rgb[y0:y1, x0:x1]); np.copyto(pin_depth[y0:y1, x0:x1], current_depth[y0:y1, x0:x1])
                    self.engine.upload_frame_window(pin_rgb, pin_depth, rgb_d, depth_d, *win)
                    self._pin_busy = not as_numpy
            else:
                rgb_d, depth_d = up(current_rgb, torch.uint8, 'rgb'), up(current_depth, torch.uint16, 'depth')
            if not render:
                rgbA_d, depthA_d = up(rgbA, torch.uint8, 'rgbA'), up(depthA, torch.uint16, 'depthA')
        n = poses.shape[0]
        if object_width is None:
            ow = self._np_bufs.get(('ow', n))
            if ow is None:
                ow = self._np_bufs[('ow', n)] = torch.full((n,), float(self.object_width), dtype=torch.float64, device=dev)
        else:
            ow = up(object_width, torch.float64, 'ow_arg')
        if render:
            mids = None
            if weight_ids is not None:
                mids = (weight_ids if torch.is_tensor(weight_ids) else torch.as_tensor(np.asarray(weight_ids))).to(dev, torch.int32)
            rgbA_d, depthA_d = self.renderer.render_batch(poses, ow, mids)
        wh = None
        if weight_ids is not None:
            wh = np.ascontiguousarray(weight_ids.cpu().numpy() if torch.is_tensor(weight_ids) else weight_ids, dtype=np.int32)
        elif self.weight_id != 0:
            wh = np.full(n, self.weight_id, dtype=np.int32)
        outs = {}
        if as_numpy:                                  # results go back to the host: persistent output buffers keep the graph key stable too
            ob = self._np_bufs.get(('out', n))
            if ob is None:
                ob = self._np_bufs[('out', n)] = (torch.empty(n, 4, 4, dtype=torch.float64, device=dev),
                                                  torch.empty(n, 3, dtype=torch.float32, device=dev), torch.empty(n, 3, dtype=torch.float32, device=dev))
            outs = dict(out_poses=ob[0], out_trans=ob[1], out_rot=ob[2])
        wd = None
        if wh is not None:                            # device copy of the ids, cached by value
            wk = ('wids', wh.tobytes())
            wd = self._np_bufs.get(wk)
            if wd is None:
                wd = self._np_bufs[wk] = torch.from_numpy(wh).to(dev)
        out, _, _ = self.engine.track_batch(rgb_d, depth_d, self.K, poses, ow, rgbA_d, depthA_d,
                                            self.trans_normalizer, self.rot_normalizer,
                                            weight_ids_host=wh, weight_ids_dev=wd, precision=self.precision, **outs)
        if staged:
            self._stage_done[self._stage_slot].record(torch.cuda.current_stream(dev))
        return out.cpu().numpy() if as_numpy else out

    # ------------------------------------------------------------------ pipelined uploads
    def _stage_uploads(self, *cpu_tensors):
        """Copy host tensors into one of two device staging sets on a side stream."""
        dev = self.engine.device
        if not hasattr(self, '_copy_stream'):
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage_bufs = [None, None]
            self._stage_done = [torch.cuda.Event(), torch.cuda.Event()]
            self._stage_slot = 0
        self._stage_slot ^= 1
        slot = self._stage_slot
        want = [(torch.float64, cpu_tensors[0]), (torch.uint8, cpu_tensors[1]), (torch.uint16, cpu_tensors[2]),
                (torch.uint8, cpu_tensors[3]), (torch.uint16, cpu_tensors[4])]
        bufs = self._stage_bufs[slot]
        if bufs is None or any(b.shape != t.shape for b, (_, t) in zip(bufs, want)):
            bufs = [torch.empty(t.shape, dtype=dt, device=dev) for dt, t in want]
            self._stage_bufs[slot] = bufs
        cs = self._copy_stream
        cs.wait_event(self._stage_done[slot])            # the kernels that last read this staging set are done
        with torch.cuda.stream(cs):
            for b, (dt, t) in zip(bufs, want):
                b.copy_(t if t.dtype == dt else t.to(dt), non_blocking=True)
            ev = torch.cuda.Event(); ev.record(cs)
        torch.cuda.current_stream(dev).wait_event(ev)
        return bufs


# ====================================================================================================
# Sequence driver + on-disk formats: the reference's predictSequenceYcbInEOAT
# (predict.py:579-623) and __main__ (predict.py:626-672) without the GUI (imshow / waitKey / VideoWriter).
#   <seq>/rgb/*.png, <seq>/depth_filled/*.png (uint16 mm), <seq>/annotated_poses/*.txt (4x4, np.loadtxt)
#   --train_data_path/../dataset_info.yml, --mean_std_path/{mean,std}.npy (train.py:124-125),
#   --ckpt_dir model_best_val.pth.tar = {'epoch', 'state_dict', ...} (problems.py:149-151), --model_path *.ply
#   -> <outdir>/%07d.txt written with np.savetxt (predict.py:611): what eval_ycb.py scores.
# ====================================================================================================
def read_rgb(path):
    """np.array(Image.open(path))[:, :, :3] (predict.py:604)."""
    from PIL import Image
    return np.ascontiguousarray(np.array(Image.open(path))[:, :, :3])


def read_depth(path):
    """cv2.imread(path, IMREAD_UNCHANGED).astype(uint16) (predict.py:606): millimetres."""
    import cv2
    d = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if d is None:
        raise FileNotFoundError(path)
    return d.astype(np.uint16)


def sequence_files(test_data_path):
    """(rgb files, depth files, ground-truth pose files), each sorted (predict.py:584-590)."""
    import glob
    rgb = sorted(glob.glob('{}/rgb/*.png'.format(test_data_path)))
    depth = sorted(glob.glob('{}/depth_filled/*.png'.format(test_data_path)))
    gt = sorted(glob.glob('{}/annotated_poses/*.txt'.format(test_data_path)))
    if not rgb or len(rgb) != len(depth):
        raise FileNotFoundError('need the same number of rgb/*.png and depth_filled/*.png under ' + str(test_data_path))
    if not gt:
        raise FileNotFoundError('need annotated_poses/*.txt (frame 0 initialises the track) under ' + str(test_data_path))
    return rgb, depth, gt


def load_run_config(train_data_path, mean_std_path):
    """dataset_info.yml next to the training data and the channel statistics (predict.py:657-664)."""
    import yaml
    with open(os.path.join(train_data_path, '../dataset_info.yml'), 'r') as ff:
        dataset_info = yaml.safe_load(ff)
    images_mean = np.load(os.path.join(mean_std_path, 'mean.npy'))
    images_std = np.load(os.path.join(mean_std_path, 'std.npy'))
    return dataset_info, images_mean, images_std


def predictSequenceYcbInEOAT(test_data_path, dataset_info, images_mean, images_std, ckpt_dir, model_path, outdir,
                             tracker=None, max_frames=None, **tracker_kwargs):
    """Track one object through a recorded sequence, starting from the first annotated pose, one pose file per frame.
    The reference's normalisers for this data set are 0.03 m / 30 degrees (predict.py:587).  Returns the (N,4,4) poses."""
    rgb_files, depth_files, gt_files = sequence_files(test_data_path)
    if tracker is None:
        tracker = Tracker(dataset_info, images_mean, images_std, ckpt_dir, model_path=model_path, trans_normalizer=0.03,
                          rot_normalizer=30 * np.pi / 180, **tracker_kwargs)
    prev_pose = np.loadtxt(gt_files[0]).copy()
    os.makedirs(outdir, exist_ok=True)
    n = len(rgb_files) if max_frames is None else min(max_frames, len(rgb_files))
    poses = []
    for i in range(n):
        rgb = read_rgb(rgb_files[i])
        depth = read_depth(depth_files[i])
        cur_pose = tracker.on_track(prev_pose.copy(), rgb, depth, gt_A_in_cam=np.eye(4), gt_B_in_cam=np.eye(4), debug=False, samples=1)
        prev_pose = cur_pose.copy()
        np.savetxt(os.path.join(outdir, '%07d.txt' % i), cur_pose)
        poses.append(cur_pose)
    return np.stack(poses)


# ----------------------------------------------------------------------------------------------------
# YCB-Video drivers (reference predict.py:299-575): predictSequenceYcb (one sequence, optional PoseCNN / PoseRBPF
# initialisation and re-initialisation frames, per-sequence ADD-S AUC) and getResultsYcb (every test sequence 0048-0059
# that contains the class; what eval_ycb.py scores).  Headless: no VideoWriter / imshow.  The data-set layout is the
# reference's:  <ycb_dir>/<seq %04d>/{color,depth_filled,seg,pose_gt/<class_id>}/..., <ycb_dir>/image_sets/keyframe.txt,
# <ycb_dir>/YCB_Video_toolbox/results_PoseCNN_RSS2018/%06d.mat (rois, poses_icp), .../PoseRBPF_Results/YCB_results_RGBD/.
# ----------------------------------------------------------------------------------------------------
def quaternion_matrix3(q_wxyz):
    """3x3 rotation of a (w, x, y, z) quaternion -- transformations.quaternion_matrix(q)[:3,:3] (reference predict.py:117)."""
    q = np.array(q_wxyz, dtype=np.float64, copy=True)
    n = np.dot(q, q)
    if n < np.finfo(float).eps * 4.0:
        return np.identity(3)
    q *= np.sqrt(2.0 / n)
    q = np.outer(q, q)
    return np.array([[1.0 - q[2, 2] - q[3, 3], q[1, 2] - q[3, 0], q[1, 3] + q[2, 0]],
                     [q[1, 2] + q[3, 0], 1.0 - q[1, 1] - q[3, 3], q[2, 3] - q[1, 0]],
                     [q[1, 3] - q[2, 0], q[2, 3] + q[1, 0], 1.0 - q[1, 1] - q[2, 2]]])


def read_keyframes(ycb_dir):
    with open('{}/image_sets/keyframe.txt'.format(ycb_dir), 'r') as ff:
        return [l.rstrip() for l in ff.readlines()]


def nearest_keyframe(seq_frames, seq_id, start_frame):
    """The keyframe of `seq_id` closest to `start_frame`, searching outwards (reference predict.py:93-107, 485-496)."""
    neighbor = 0
    while neighbor < 100000:
        for cand in (start_frame + neighbor, start_frame - neighbor):
            tmp = '%04d/%06d' % (seq_id, cand)
            if tmp in seq_frames:
                return tmp, seq_frames.index(tmp), cand
        neighbor += 1
    raise ValueError('sequence %04d has no keyframe' % seq_id)


def posecnn_pose(mat_path, class_id):
    """Pose of `class_id` from a PoseCNN result file: rois[:,1] == class id, poses_icp = (qw,qx,qy,qz,x,y,z) (predict.py:111-122)."""
    import scipy.io
    res = scipy.io.loadmat(mat_path)
    idx = np.where(res['rois'][:, 1] == class_id)
    tmp = res['poses_icp'][idx].reshape(-1)
    if tmp.size < 7:
        raise ValueError('class %d not in %s' % (class_id, mat_path))
    pose = np.eye(4)
    pose[:3, :3] = quaternion_matrix3(tmp[:4])
    pose[:3, 3] = tmp[4:7]
    return pose


def use_posecnn_res(class_id, seq_frame_str, ycb_dir, posecnn_dir=None):
    """PoseCNN's estimate at the keyframe nearest to `seq_frame_str` = '%04d/%06d' (reference predict.py:89-123)."""
    seq_frames = read_keyframes(ycb_dir)
    seq_id, start_frame = int(seq_frame_str.split('/')[0]), int(seq_frame_str.split('/')[1])
    _, index, _ = nearest_keyframe(seq_frames, seq_id, start_frame)
    posecnn_dir = posecnn_dir or '{}/YCB_Video_toolbox/results_PoseCNN_RSS2018/'.format(ycb_dir)
    return posecnn_pose(os.path.join(posecnn_dir, '%06d.mat' % index), class_id)


def poserbpf_pose(ycb_dir, class_id, seq_id, seqs):
    """First pose of PoseRBPF's result file for (class, sequence) (reference predict.py:376-390, 498-513): 'x y z qw qx qy qz' after two tokens."""
    import glob
    res_dir = '{}/YCB_Video_toolbox/PoseRBPF_Results/YCB_results_RGBD/'.format(ycb_dir)
    folders = sorted(os.listdir(res_dir))
    cur = res_dir + folders[class_id - 1] + '/' + 'seq_{}/'.format(seqs.index(seq_id) + 1)
    with open(glob.glob(cur + 'Pose*.txt')[0], 'r') as ff:
        pose = ff.readlines()[0].rstrip().split()[2:]
    out = np.eye(4)
    out[:3, 3] = np.array(pose[:3], dtype=np.float64)
    out[:3, :3] = quaternion_matrix3(np.array(pose[3:7], dtype=np.float64))
    return out


def findClassContainedVideosYcb(class_id, data_dir, testset=True):
    """Sequence ids under `data_dir` whose pose_gt/ has a folder for `class_id` (reference Utils.py:108-123; test set = 0048..0059)."""
    import glob, re
    out = []
    for gt_dir in sorted(glob.glob(os.path.join(data_dir, '**/pose_gt'))):
        video_index = int(re.findall(r'/[0-9]{4}/', gt_dir + '/')[0][1:-1])
        if testset and (video_index < 48 or video_index > 59):
            continue
        if class_id in list(map(int, os.listdir(gt_dir))):
            out.append(video_index)
    return out


def _ycb_sequence_files(seq_dir, class_id):
    import glob
    rgb = sorted(glob.glob(os.path.join(seq_dir, 'color/*')))
    depth = sorted(glob.glob(os.path.join(seq_dir, 'depth_filled/*')))
    gt = sorted(glob.glob(os.path.join(seq_dir, 'pose_gt/{}/*'.format(class_id))))
    if not rgb or len(rgb) != len(depth) or len(gt) < len(rgb):
        raise FileNotFoundError('need matching color/, depth_filled/ and pose_gt/%d/ files under %s' % (class_id, seq_dir))
    return rgb, depth, gt


def predictSequenceYcb(ycb_dir, seq_id, class_id, dataset_info, images_mean, images_std, ckpt_dir, model_path, outdir,
                       init='gt', reinit_frames=None, start_frame=0, tracker=None, max_frames=None, **tracker_kwargs):
    """Track `class_id` through YCB-Video sequence `seq_id` (reference predict.py:446-575).  init: 'gt' | 'posecnn' | 'poserbpf'.
    reinit_frames: '%04d/%06d' strings (1-based frame ids, as the reference compares them): the track restarts there from
    PoseCNN's estimate.  Writes <outdir>/%05d.txt and %05dgt.txt and returns (poses (N,4,4), ADD-S AUC in percent)."""
    test_data_path = '{}/%04d'.format(ycb_dir) % seq_id
    rgb_files, depth_files, gt_files = _ycb_sequence_files(test_data_path, class_id)
    gt_poses = [np.loadtxt(f) for f in gt_files]
    reinit_frames = list(reinit_frames or [])
    if tracker is None:
        tracker = Tracker(dataset_info, images_mean, images_std, ckpt_dir, model_path=model_path, **tracker_kwargs)
    if init == 'gt':
        prev_pose = gt_poses[start_frame].copy()
    elif init == 'posecnn':
        seq_frame_str, _, start_frame = nearest_keyframe(read_keyframes(ycb_dir), seq_id, start_frame)
        prev_pose = use_posecnn_res(class_id, seq_frame_str, ycb_dir)
    elif init == 'poserbpf':
        seqs = sorted(findClassContainedVideosYcb(class_id, ycb_dir, testset=True))
        prev_pose = poserbpf_pose(ycb_dir, class_id, seq_id, seqs)
    else:
        raise ValueError('init must be gt, posecnn or poserbpf')
    pred_poses = [prev_pose]
    os.makedirs(outdir, exist_ok=True)
    n = len(rgb_files) if max_frames is None else min(len(rgb_files), start_frame + 1 + max_frames)
    for i in range(start_frame + 1, n):
        rgb = read_rgb(rgb_files[i])
        depth = read_depth(depth_files[i])
        A_in_cam = prev_pose.copy()
        if '%04d/%06d' % (seq_id, i + 1) in reinit_frames:
            A_in_cam = use_posecnn_res(class_id, '%04d/%06d' % (seq_id, i - 1), ycb_dir)
        cur_pose = tracker.on_track(A_in_cam, rgb, depth, gt_A_in_cam=gt_poses[i - 1], gt_B_in_cam=gt_poses[i], debug=False, samples=1)
        prev_pose = cur_pose.copy()
        pred_poses.append(cur_pose)
    pred_poses = np.array(pred_poses)
    for i in range(len(pred_poses)):
        np.savetxt(os.path.join(outdir, '%05d.txt' % i), pred_poses[i])
        np.savetxt(os.path.join(outdir, '%05dgt.txt' % i), gt_poses[start_frame + i])
    adi_auc = None
    if tracker.object_cloud is not None:                           # per-sequence ADD-S AUC (predict.py:566-575) on the device (csrc/metrics.cu)
        eng = tracker.engine
        pts = torch.from_numpy(np.ascontiguousarray(tracker.object_cloud.points)).to(eng.device)
        gts = torch.from_numpy(np.stack(gt_poses[start_frame:start_frame + len(pred_poses)])).to(eng.device)
        _, adi = eng.add_adi(pts, torch.from_numpy(pred_poses).to(eng.device), gts, want_add=False)
        adi_auc = eng.vocap(adi) * 100
    return pred_poses, adi_auc


def getResultsYcb(ycb_dir, class_id, dataset_info, images_mean, images_std, ckpt_dir, model_path, outdir,
                  initialize_method='gt', tracker=None, max_frames=None, **tracker_kwargs):
    """Every YCB-Video TEST sequence (0048..0059) under <ycb_dir>/data_organized/ that contains `class_id`, tracked from its first
    frame; one <outdir>/seq<id>/%07d.txt per frame -- the files eval_ycb.py globs (reference predict.py:299-443).  Returns {seq_id: poses}."""
    import glob, re
    test_data_dir = '{}/data_organized/'.format(ycb_dir)
    os.makedirs(outdir, exist_ok=True)
    if tracker is None:
        tracker = Tracker(dataset_info, images_mean, images_std, ckpt_dir, model_path=model_path, **tracker_kwargs)
    keyframes_all = read_keyframes(ycb_dir) if initialize_method == 'posecnn' else []
    seqs = sorted(findClassContainedVideosYcb(class_id, test_data_dir, testset=True))
    results = {}
    for gt_dir in sorted(glob.glob(test_data_dir + '**/pose_gt')):
        seq_id = int(re.findall(r'/\d{4}/', gt_dir + '/')[0][1:-1])
        if seq_id not in seqs:
            continue
        seq_dir = os.path.join(gt_dir, '..')
        rgb_files, depth_files, gt_files = _ycb_sequence_files(seq_dir, class_id)
        if initialize_method == 'posecnn':
            seq_frame = '%04d/%06d' % (seq_id, 1)
            prev_pose = posecnn_pose('{}/YCB_Video_toolbox/results_PoseCNN_RSS2018/%06d.mat'.format(ycb_dir) % keyframes_all.index(seq_frame), class_id)
        elif initialize_method == 'poserbpf':
            prev_pose = poserbpf_pose(ycb_dir, class_id, seq_id, seqs)
        elif initialize_method == 'gt':
            prev_pose = np.loadtxt(gt_files[0])
        else:
            raise ValueError('initialize_method must be gt, posecnn or poserbpf')
        pred_poses = [prev_pose]
        n = len(rgb_files) if max_frames is None else min(len(rgb_files), 1 + max_frames)
        for i in range(1, n):
            rgb = read_rgb(rgb_files[i])
            depth = read_depth(depth_files[i])
            cur_pose = tracker.on_track(prev_pose, rgb, depth, gt_A_in_cam=None, gt_B_in_cam=np.loadtxt(gt_files[i]), debug=False, samples=1)
            prev_pose = cur_pose.copy()
            pred_poses.append(cur_pose)
        while len(pred_poses) < len(rgb_files) and max_frames is None:      # predict.py:437-440
            pred_poses.append(pred_poses[-1])
        sdir = os.path.join(outdir, 'seq{}'.format(seq_id))
        os.makedirs(sdir, exist_ok=True)
        for i in range(len(pred_poses)):
            np.savetxt(os.path.join(sdir, '%07d.txt' % i), pred_poses[i])
        results[seq_id] = np.array(pred_poses)
    return results


def main(argv=None):
    import argparse
    parser = argparse.ArgumentParser(description='headless se(3)-TrackNet sequence tracking on libse3tn (flags of the reference predict.py:626-641)')
    parser.add_argument('--mode', default='ycbv', help='ycbv (one YCB-Video sequence) / ycbineoat / anything else: every YCB-Video test sequence of the class')
    parser.add_argument('--seq_id', default=None, type=int)
    parser.add_argument('--ycb_dir', default=None)
    parser.add_argument('--YCBInEOAT_dir', default=None)
    parser.add_argument('--train_data_path', required=True, help='dataset_info.yml is read from <train_data_path>/../')
    parser.add_argument('--class_id', default=-1, type=int, help='class id in YCB Video')
    parser.add_argument('--model_path', type=str, required=True, help='path to mesh (.ply with normals and vertex colours for the CUDA renderer)')
    parser.add_argument('--ckpt_dir', type=str, required=True)
    parser.add_argument('--mean_std_path', type=str, required=True)
    parser.add_argument('--outdir', type=str, required=True)
    parser.add_argument('--reinit_frames', type=str, default=None, help='comma-separated %%04d/%%06d frames to re-initialise from PoseCNN')
    parser.add_argument('--init', default='gt', help='gt / posecnn / poserbpf (the reference hard-codes gt)')
    parser.add_argument('--max_frames', type=int, default=None)
    args = parser.parse_args(argv)
    dataset_info, images_mean, images_std = load_run_config(args.train_data_path, args.mean_std_path)
    if args.mode == 'ycbineoat':
        if not args.YCBInEOAT_dir:
            raise SystemExit('--mode ycbineoat needs --YCBInEOAT_dir')
        poses = predictSequenceYcbInEOAT(args.YCBInEOAT_dir, dataset_info, images_mean, images_std, args.ckpt_dir, args.model_path,
                                         args.outdir, max_frames=args.max_frames)
        print('wrote %d poses to %s' % (len(poses), args.outdir))
        return
    if not args.ycb_dir:
        raise SystemExit('--mode %s needs --ycb_dir' % args.mode)
    if args.mode == 'ycbv':
        if args.seq_id is None:
            raise SystemExit('--mode ycbv needs --seq_id')
        class_id = args.class_id if args.class_id is not None and args.class_id >= 0 else 4      # predict.py:450-452
        reinit = args.reinit_frames.split(',') if args.reinit_frames else None
        poses, auc = predictSequenceYcb(args.ycb_dir, args.seq_id, class_id, dataset_info, images_mean, images_std, args.ckpt_dir,
                                        args.model_path, args.outdir, init=args.init, reinit_frames=reinit, max_frames=args.max_frames)
        print('reinit_frames {}, adi_auc {}'.format(reinit or '', auc))
        return
    res = getResultsYcb(args.ycb_dir, args.class_id, dataset_info, images_mean, images_std, args.ckpt_dir, args.model_path, args.outdir,
                        initialize_method=args.init, max_frames=args.max_frames)
    print('tracked class %d through sequences %s -> %s' % (args.class_id, sorted(res), args.outdir))


if __name__ == '__main__':
    main()
