"""Engine: one libse3tn context bound to one CUDA device, driven with torch tensors.

PyTorch is used here for device-memory ownership, the current stream and (in dist.py)
torch.distributed -- all arithmetic happens inside libse3tn.so.  There is no CPU or eager
fallback: without a CUDA device and the built library every call raises.
"""
import ctypes as C
import numpy as np
import torch

from . import _lib
from .weights import pack_state_dict

PREC = {'tf32': _lib.PREC_TF32, 'fp32': _lib.PREC_FP32, 'bf16x3': _lib.PREC_BF16X3, 'bf16': _lib.PREC_BF16}
IMAGE_SIZE = 176


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Engine:
    def __init__(self, max_batch=64, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError('se3tn Engine needs a CUDA device (sm_100a); there is no CPU fallback')
        self.lib = _lib.load()
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else
                                   (device.index if isinstance(device, torch.device) else int(device)))
        self.max_batch = int(max_batch)
        nbytes = self.lib.se3tn_workspace_bytes(self.max_batch)
        # caller-owned workspace: a torch allocation, so torch's allocator accounts for it
        self._workspace = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
        base = self._workspace.data_ptr()
        aligned = (base + 1023) // 1024 * 1024
        ctx = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.se3tn_create(self.device.index, self.max_batch, C.c_void_p(aligned), C.byref(ctx))
        _lib.check(rc, None)
        self._ctx = ctx
        self._weight_ids = set()

    def close(self):
        if getattr(self, '_ctx', None):
            torch.cuda.synchronize(self.device)
            self.lib.se3tn_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ parameters
    def load_state_dict(self, state_dict, weight_id=0):
        blob = pack_state_dict(state_dict)
        _lib.check(self.lib.se3tn_load_weights(self._ctx, int(weight_id), blob.ctypes.data_as(C.c_void_p), blob.size), self._ctx)
        self._weight_ids.add(int(weight_id))

    def set_stats(self, mean, std, weight_id=0):
        mean = np.ascontiguousarray(mean); std = np.ascontiguousarray(std)
        if mean.shape != (8,) or std.shape != (8,):
            raise ValueError('mean/std must be 8-vectors (A channels then B channels)')
        f64 = (mean.dtype == np.float64) or (std.dtype == np.float64)
        dt = np.float64 if f64 else np.float32
        mean = mean.astype(dt); std = std.astype(dt)
        _lib.check(self.lib.se3tn_set_stats(self._ctx, int(weight_id), mean.ctypes.data_as(C.c_void_p),
                                            std.ctypes.data_as(C.c_void_p), int(f64)), self._ctx)

    # ------------------------------------------------------------------ hot path
    def forward(self, A, B, weight_id=0, precision='bf16x3', want_feature=False):
        """Se3TrackNet.forward on float32 (n,4,176,176) CUDA tensors -> (trans, rot, feature|None)."""
        self._check_img(A); self._check_img(B)
        n = A.shape[0]
        if B.shape[0] != n:
            raise ValueError('A and B batch sizes differ')
        trans = torch.empty(n, 3, dtype=torch.float32, device=self.device)
        rot = torch.empty(n, 3, dtype=torch.float32, device=self.device)
        feat = torch.empty(n, 256, 22, 22, dtype=torch.float32, device=self.device) if want_feature else None
        for i0 in range(0, n, self.max_batch):
            i1 = min(n, i0 + self.max_batch)
            _lib.check(self.lib.se3tn_forward(self._ctx, int(weight_id), _ptr(A[i0:i1]), _ptr(B[i0:i1]), i1 - i0,
                                              _ptr(trans[i0:i1]), _ptr(rot[i0:i1]),
                                              _ptr(feat[i0:i1]) if feat is not None else C.c_void_p(0),
                                              PREC[precision], _stream(self.device)), self._ctx)
        return trans, rot, feat

    def preprocess(self, frame_rgb, frame_depth, K, poses, object_width, rgbA, depthA, weight_ids=None,
                   precision='bf16x3', want_tensors=False, want_crops=False):
        n = poses.shape[0]
        self._check_frame(frame_rgb, frame_depth, rgbA, depthA, poses, object_width, n)
        H, W = frame_depth.shape
        Kh = self._k4(K)
        outA = outB = crop_rgb = crop_depth = None
        if want_tensors:
            outA = torch.empty(n, 4, IMAGE_SIZE, IMAGE_SIZE, dtype=torch.float32, device=self.device)
            outB = torch.empty_like(outA)
        if want_crops:
            crop_rgb = torch.empty(n, IMAGE_SIZE, IMAGE_SIZE, 3, dtype=torch.uint8, device=self.device)
            crop_depth = torch.empty(n, IMAGE_SIZE, IMAGE_SIZE, dtype=torch.uint16, device=self.device)
        _lib.check(self.lib.se3tn_preprocess(self._ctx, _ptr(frame_rgb), _ptr(frame_depth), H, W,
                                             Kh.ctypes.data_as(C.c_void_p), _ptr(poses), _ptr(object_width),
                                             _ptr(rgbA), _ptr(depthA), _ptr(weight_ids), n, PREC[precision],
                                             _ptr(outA), _ptr(outB), _ptr(crop_rgb), _ptr(crop_depth),
                                             _stream(self.device)), self._ctx)
        return outA, outB, crop_rgb, crop_depth

    def normalize(self, rgbA, depthA, rgbB, depthB, poses, weight_ids=None, precision='bf16x3', want_tensors=True):
        """processData's post-transforms on existing 176x176 crops (all CUDA tensors)."""
        n = poses.shape[0]
        outA = outB = None
        if want_tensors:
            outA = torch.empty(n, 4, IMAGE_SIZE, IMAGE_SIZE, dtype=torch.float32, device=self.device)
            outB = torch.empty_like(outA)
        _lib.check(self.lib.se3tn_normalize(self._ctx, _ptr(rgbA), _ptr(depthA), _ptr(rgbB), _ptr(depthB), _ptr(poses),
                                            _ptr(weight_ids), n, PREC[precision], _ptr(outA), _ptr(outB),
                                            _stream(self.device)), self._ctx)
        return outA, outB

    def compute_bbox(self, poses, K, widths, scale=(1000., 1000., 1000.)):
        n = poses.shape[0]
        out = torch.empty(n, 4, 2, dtype=torch.int32, device=self.device)
        Kh = self._k4(K); sc = np.ascontiguousarray(scale, dtype=np.float64)
        _lib.check(self.lib.se3tn_compute_bbox(self._ctx, _ptr(poses), Kh.ctypes.data_as(C.c_void_p), _ptr(widths),
                                               sc.ctypes.data_as(C.c_void_p), _ptr(out), n, _stream(self.device)), self._ctx)
        return out

    def crop_bbox(self, frame_rgb, frame_depth, bbox, out_hw=(IMAGE_SIZE, IMAGE_SIZE)):
        n = bbox.shape[0]
        H, W = frame_depth.shape
        crop_rgb = torch.empty(n, out_hw[0], out_hw[1], 3, dtype=torch.uint8, device=self.device)
        crop_depth = torch.empty(n, out_hw[0], out_hw[1], dtype=torch.uint16, device=self.device)
        _lib.check(self.lib.se3tn_crop_bbox(self._ctx, _ptr(frame_rgb), _ptr(frame_depth), H, W, _ptr(bbox), n,
                                            int(out_hw[0]), int(out_hw[1]), _ptr(crop_rgb), _ptr(crop_depth),
                                            _stream(self.device)), self._ctx)
        return crop_rgb, crop_depth

    def forward_preprocessed(self, n, weight_id=0, first=0, precision='bf16x3', want_feature=False):
        trans = torch.empty(n, 3, dtype=torch.float32, device=self.device)
        rot = torch.empty(n, 3, dtype=torch.float32, device=self.device)
        feat = torch.empty(n, 256, 22, 22, dtype=torch.float32, device=self.device) if want_feature else None
        _lib.check(self.lib.se3tn_forward_preprocessed(self._ctx, int(weight_id), int(first), n, _ptr(trans), _ptr(rot),
                                                       _ptr(feat), PREC[precision], _stream(self.device)), self._ctx)
        return trans, rot, feat

    def pose_update(self, poses, trans, rot, trans_normalizer, rot_normalizer, out=None):
        n = poses.shape[0]
        if poses.dtype != torch.float64 or not poses.is_contiguous() or poses.shape[1:] != (4, 4):
            raise ValueError('poses must be a contiguous float64 (n,4,4) CUDA tensor')
        out = torch.empty_like(poses) if out is None else out
        _lib.check(self.lib.se3tn_pose_update(self._ctx, _ptr(poses), _ptr(trans), _ptr(rot), float(trans_normalizer),
                                              float(rot_normalizer), _ptr(out), n, _stream(self.device)), self._ctx)
        return out

    def so3_log(self, poses_a, poses_b, trans_normalizer, rot_normalizer):
        n = poses_a.shape[0]
        tl = torch.empty(n, 3, dtype=torch.float64, device=self.device)
        rl = torch.empty(n, 3, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.se3tn_so3_log(self._ctx, _ptr(poses_a), _ptr(poses_b), float(trans_normalizer),
                                          float(rot_normalizer), _ptr(tl), _ptr(rl), n, _stream(self.device)), self._ctx)
        return tl, rl

    def track_batch(self, frame_rgb, frame_depth, K, poses, object_width, rgbA, depthA,
                    trans_normalizer, rot_normalizer, weight_ids_host=None, weight_ids_dev=None,
                    precision='bf16x3', out_poses=None, out_trans=None, out_rot=None):
        """n independent tracks of one frame: K0 -> conv stack -> K6, all enqueued on the current stream."""
        n = poses.shape[0]
        self._check_frame(frame_rgb, frame_depth, rgbA, depthA, poses, object_width, n)
        H, W = frame_depth.shape
        Kh = self._k4(K)
        out_poses = torch.empty_like(poses) if out_poses is None else out_poses
        out_trans = torch.empty(n, 3, dtype=torch.float32, device=self.device) if out_trans is None else out_trans
        out_rot = torch.empty(n, 3, dtype=torch.float32, device=self.device) if out_rot is None else out_rot
        wh = None
        if weight_ids_host is not None:
            wh = np.ascontiguousarray(weight_ids_host, dtype=np.int32)
            if weight_ids_dev is None:
                weight_ids_dev = torch.from_numpy(wh).to(self.device)
        _lib.check(self.lib.se3tn_track_batch(self._ctx, _ptr(frame_rgb), _ptr(frame_depth), H, W,
                                              Kh.ctypes.data_as(C.c_void_p), _ptr(poses), _ptr(object_width),
                                              _ptr(rgbA), _ptr(depthA),
                                              wh.ctypes.data_as(C.c_void_p) if wh is not None else C.c_void_p(0),
                                              _ptr(weight_ids_dev), n, float(trans_normalizer), float(rot_normalizer),
                                              PREC[precision], _ptr(out_trans), _ptr(out_rot), _ptr(out_poses),
                                              _stream(self.device)), self._ctx)
        return out_poses, out_trans, out_rot

    def track_host(self, frame_rgb, frame_depth, K, poses, object_width, rgbA, depthA, trans_normalizer, rot_normalizer,
                   weight_ids=None, precision='bf16x3', want_residuals=False):
        """The reference's calling pattern as one library call: numpy arrays in, numpy poses out, synchronous (se3tn_track_host).
        frame_rgb uint8 (H,W,3), frame_depth uint16 (H,W), poses float64 (n,4,4), object_width float64 (n), rgbA uint8
        (n,176,176,3), depthA uint16 (n,176,176), weight_ids int32 (n) or None -- all C-contiguous."""
        n = int(poses.shape[0])
        for name, a, dt, shape in (('frame_rgb', frame_rgb, np.uint8, frame_depth.shape + (3,)), ('frame_depth', frame_depth, np.uint16, frame_depth.shape),
                                   ('poses', poses, np.float64, (n, 4, 4)), ('object_width', object_width, np.float64, (n,)),
                                   ('rgbA', rgbA, np.uint8, (n, IMAGE_SIZE, IMAGE_SIZE, 3)), ('depthA', depthA, np.uint16, (n, IMAGE_SIZE, IMAGE_SIZE))):
            if not (isinstance(a, np.ndarray) and a.dtype == dt and tuple(a.shape) == tuple(shape) and a.flags['C_CONTIGUOUS']):
                raise ValueError('track_host: %s must be a C-contiguous %s array of shape %s' % (name, np.dtype(dt).name, tuple(shape)))
        if frame_depth.ndim != 2:
            raise ValueError('track_host: frame_depth must be (H, W)')
        H, W = frame_depth.shape
        Kh = self._k4(K)
        wid = None
        if weight_ids is not None:
            wid = np.ascontiguousarray(weight_ids, dtype=np.int32)
            if wid.shape != (n,):
                raise ValueError('track_host: weight_ids must have one entry per track')
        out = np.empty((n, 4, 4), dtype=np.float64)
        tr = np.empty((n, 3), dtype=np.float32) if want_residuals else None
        ro = np.empty((n, 3), dtype=np.float32) if want_residuals else None
        vp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else C.c_void_p(0)
        _lib.check(self.lib.se3tn_track_host(self._ctx, vp(frame_rgb), vp(frame_depth), int(H), int(W), vp(Kh), vp(poses), vp(object_width),
                                             vp(rgbA), vp(depthA), vp(wid), n, float(trans_normalizer), float(rot_normalizer), PREC[precision],
                                             vp(out), vp(tr), vp(ro), _stream(self.device)), self._ctx)
        return (out, tr, ro) if want_residuals else out

    def upload_frame_window(self, rgb_host, depth_host, rgb_dev, depth_dev, y0, y1, x0, x1):
        """Copy rows [y0,y1) x columns [x0,x1) of contiguous numpy frames (uint8 (H,W,3), uint16 (H,W)) into full-size device frame
        buffers: K0 only reads a frame inside the tracks' crop windows."""
        H, W = depth_host.shape
        _lib.check(self.lib.se3tn_upload_frame_window(self._ctx, rgb_host.ctypes.data_as(C.c_void_p), depth_host.ctypes.data_as(C.c_void_p), int(H), int(W),
                                                      int(y0), int(y1), int(x0), int(x1), _ptr(rgb_dev), _ptr(depth_dev), _stream(self.device)), self._ctx)

    # ------------------------------------------------------------------ metrics
    def add_adi(self, model_pts, pred, gt, want_add=True, want_adi=True):
        """ADD / ADD-S (reference Utils.py:72-98) of n pose pairs: float64 CUDA tensors model (m,3), pred/gt (n,4,4)."""
        n = pred.shape[0]
        out_add = torch.empty(n, dtype=torch.float64, device=self.device) if want_add else None
        out_adi = torch.empty(n, dtype=torch.float64, device=self.device) if want_adi else None
        _lib.check(self.lib.se3tn_add_adi(self._ctx, _ptr(model_pts), int(model_pts.shape[0]), _ptr(pred), _ptr(gt), n,
                                          _ptr(out_add), _ptr(out_adi), _stream(self.device)), self._ctx)
        return out_add, out_adi

    def vocap(self, errs):
        """VOCap (reference eval_ycb.py:45-64) of a float64 CUDA error vector -> python float in [0,1]."""
        ap = C.c_double(0.0)
        _lib.check(self.lib.se3tn_vocap(self._ctx, _ptr(errs), int(errs.numel()), C.byref(ap), _stream(self.device)), self._ctx)
        return ap.value

    # ------------------------------------------------------------------ input A renderer
    def set_mesh(self, mesh, mesh_id=0):
        """Upload a CAD model (dict pos float32 (nv,3), nrm float32 (nv,3), col uint8 (nv,3), faces int32 (nf,3)) -- the
        vertex / index buffers of the reference's VispyRenderer (vispy_renderer.py:108-129)."""
        pos = np.ascontiguousarray(mesh['pos'], dtype=np.float32); nrm = np.ascontiguousarray(mesh['nrm'], dtype=np.float32)
        col = np.ascontiguousarray(mesh['col'], dtype=np.uint8); faces = np.ascontiguousarray(mesh['faces'], dtype=np.int32)
        if pos.ndim != 2 or pos.shape[1] != 3 or nrm.shape != pos.shape or col.shape != pos.shape or faces.ndim != 2 or faces.shape[1] != 3:
            raise ValueError('mesh arrays must be pos/nrm/col (nv,3) and faces (nf,3)')
        _lib.check(self.lib.se3tn_set_mesh(self._ctx, int(mesh_id), pos.ctypes.data, nrm.ctypes.data, col.ctypes.data, faces.ctypes.data,
                                           int(pos.shape[0]), int(faces.shape[0])), self._ctx)

    def render(self, K, poses, object_width, mesh_ids=None, out_rgb=None, out_depth=None, mode='vispy', image_hw=None):
        """Tracker.render_window for n tracks (reference predict.py:193-215): float64 CUDA poses (n,4,4) and widths (n) ->
        rgbA uint8 (n,176,176,3), depthA uint16 (n,176,176) CUDA tensors.  mode 'vispy' (lit, the crop window is the GL viewport) or
        'pyrender' (unlit render of the whole image_hw = (H, W) camera image, then crop_bbox; dataset_info['renderer'] == 'pyrenderer')."""
        n = int(poses.shape[0])
        rgb = out_rgb if out_rgb is not None else torch.empty((n, 176, 176, 3), dtype=torch.uint8, device=self.device)
        dep = out_depth if out_depth is not None else torch.empty((n, 176, 176), dtype=torch.uint16, device=self.device)
        Kh = self._k4(K)
        if mode not in ('vispy', 'pyrender'):
            raise ValueError("render mode must be 'vispy' or 'pyrender'")
        if mode == 'pyrender' and image_hw is None:
            raise ValueError("render(mode='pyrender') needs image_hw=(H, W), the camera image pyrender draws")
        H, W = (int(image_hw[0]), int(image_hw[1])) if image_hw is not None else (0, 0)
        _lib.check(self.lib.se3tn_render_ex(self._ctx, Kh.ctypes.data_as(C.c_void_p), _ptr(poses), _ptr(object_width), _ptr(mesh_ids), n,
                                            _lib.RENDER_PYRENDER if mode == 'pyrender' else _lib.RENDER_VISPY, H, W, _ptr(rgb), _ptr(dep),
                                            _stream(self.device)), self._ctx)
        return rgb, dep

    # ------------------------------------------------------------------ pose exchange over a raw NCCL communicator
    def allgather_poses_nccl(self, nccl_comm, local_poses, out=None, world_size=None):
        """se3tn_allgather_poses: for hosts that own an ncclComm_t (an int / c_void_p handle).  torch.distributed users call
        dist.all_gather_poses instead.  local_poses (n,4,4) float64 CUDA -> (world*n,4,4), rank-major."""
        n = int(local_poses.shape[0])
        if out is None:
            if world_size is None:
                raise ValueError('need out= or world_size=')
            out = torch.empty((world_size * n, 4, 4), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.se3tn_allgather_poses(self._ctx, C.c_void_p(int(nccl_comm)), _ptr(local_poses), _ptr(out), n, _stream(self.device)), self._ctx)
        return out

    # ------------------------------------------------------------------ live-sensor depth
    def fill_depth(self, depth_mm, max_depth=2.0, want_metres=False, extrapolate=False, blur_type='bilateral'):
        """fill_depth (reference Utils.py:455-514; predict_ros.py:38-41 uses the defaults): uint16 mm CUDA tensor (H,W) ->
        uint16 mm CUDA tensor (and float32 metres with want_metres)."""
        if depth_mm.dtype != torch.uint16 or depth_mm.dim() != 2 or not depth_mm.is_cuda or not depth_mm.is_contiguous():
            raise ValueError('depth_mm must be a contiguous uint16 CUDA tensor (H,W)')
        if blur_type not in ('bilateral', 'gaussian'):
            raise ValueError("blur_type must be 'bilateral' or 'gaussian'")
        H, W = depth_mm.shape
        out = torch.empty_like(depth_mm)
        out_m = torch.empty((H, W), dtype=torch.float32, device=self.device) if want_metres else None
        _lib.check(self.lib.se3tn_fill_depth_ex(self._ctx, _ptr(depth_mm), int(H), int(W), float(max_depth), int(bool(extrapolate)),
                                                1 if blur_type == 'gaussian' else 0, _ptr(out), _ptr(out_m), _stream(self.device)), self._ctx)
        return (out, out_m) if want_metres else out

    # ------------------------------------------------------------------ introspection
    def debug_buffer(self, buf_id, n):
        """A float32 view (n, floats_per_image) of an internal NHWC activation buffer."""
        p = C.c_void_p(); fpi = C.c_size_t()
        _lib.check(self.lib.se3tn_debug_buffer(self._ctx, buf_id, C.byref(p), C.byref(fpi)), self._ctx)
        offb = p.value - self._workspace.data_ptr()
        nbytes = n * fpi.value * 4
        return self._workspace[offb:offb + nbytes].view(torch.float32).view(n, fpi.value)

    def set_profiling(self, enable=True):
        _lib.check(self.lib.se3tn_set_profiling(self._ctx, int(bool(enable))), self._ctx)

    def get_profile(self):
        """Device ms of each kernel of the last call (21 slots, see include/se3tn.h)."""
        ms = (C.c_float * 21)()
        _lib.check(self.lib.se3tn_get_profile(self._ctx, ms), self._ctx)
        return np.array(ms[:], dtype=np.float64)

    def get_trace(self):
        """(14, 256, 8) uint64 globaltimer stamps of the last forward's conv CTAs (needs SE3TN_TRACE=1 at Engine creation)."""
        out = np.zeros((14, 256, 8), dtype=np.uint64)
        _lib.check(self.lib.se3tn_get_trace(self._ctx, out.ctypes.data_as(C.c_void_p)), self._ctx)
        return out

    def last_launch_count(self):
        return self.lib.se3tn_last_launch_count(self._ctx)

    def last_step_was_graph(self):
        """True when the last track_batch call replayed (or just captured and launched) a CUDA graph of the whole step."""
        return bool(self.lib.se3tn_last_step_was_graph(self._ctx))

    # ------------------------------------------------------------------ checks
    def _check_img(self, t):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.dim() == 4 and
                tuple(t.shape[1:]) == (4, IMAGE_SIZE, IMAGE_SIZE)):
            raise ValueError('expected a contiguous float32 CUDA tensor of shape (n,4,176,176), got %s %s' % (t.dtype, tuple(t.shape)))

    def _check_frame(self, rgb, depth, rgbA, depthA, poses, ow, n):
        ok = (rgb.is_cuda and rgb.dtype == torch.uint8 and rgb.is_contiguous() and rgb.dim() == 3 and rgb.shape[2] == 3 and
              depth.is_cuda and depth.dtype == torch.uint16 and depth.is_contiguous() and depth.shape == rgb.shape[:2] and
              rgbA.is_cuda and rgbA.dtype == torch.uint8 and rgbA.is_contiguous() and tuple(rgbA.shape) == (n, IMAGE_SIZE, IMAGE_SIZE, 3) and
              depthA.is_cuda and depthA.dtype == torch.uint16 and depthA.is_contiguous() and tuple(depthA.shape) == (n, IMAGE_SIZE, IMAGE_SIZE) and
              poses.is_cuda and poses.dtype == torch.float64 and poses.is_contiguous() and tuple(poses.shape) == (n, 4, 4) and
              ow.is_cuda and ow.dtype == torch.float64 and ow.is_contiguous() and tuple(ow.shape) == (n,))
        if not ok:
            raise ValueError('bad frame/pose tensors: need uint8 (H,W,3), uint16 (H,W), uint8 (n,176,176,3), uint16 (n,176,176), '
                             'float64 (n,4,4), float64 (n,), all contiguous CUDA tensors')
        if n > self.max_batch:
            raise ValueError('n=%d exceeds max_batch=%d' % (n, self.max_batch))

    @staticmethod
    def _k4(K):
        K = np.asarray(K, dtype=np.float64)
        if K.shape == (3, 3):
            return np.ascontiguousarray([K[0, 0], K[1, 1], K[0, 2], K[1, 2]], dtype=np.float64)
        if K.shape == (4,):
            return np.ascontiguousarray(K)
        raise ValueError('K must be 3x3 or (fx,fy,cx,cy)')
