"""Input A without OpenGL: the reference's VispyRenderer (vispy_renderer.py) + Tracker.render_window
(predict.py:193-215) as one CUDA launch for all tracks (csrc/render.cu).  `CudaRenderer` plugs into
`Tracker(renderer=...)`: it exposes render_window(ob2cam) -> (rgb uint8 (176,176,3), depth uint16 (176,176)),
the contract of the reference method, and render_batch() for device-resident loops."""
import numpy as np
import torch

from .mesh_io import load_ply_mesh


class CudaRenderer:
    def __init__(self, model, K, engine, object_width, mesh_id=0):
        """model: path of a .ply (what VispyRenderer takes, vispy_renderer.py:107-123) or a mesh dict."""
        self.mesh = load_ply_mesh(model) if isinstance(model, str) else model
        self.K = np.asarray(K, dtype=np.float64).copy()
        self.engine = engine
        self.mesh_id = int(mesh_id)
        self.object_width = float(object_width)
        engine.set_mesh(self.mesh, self.mesh_id)

    def render_batch(self, poses, object_width=None, mesh_ids=None, out_rgb=None, out_depth=None):
        """poses (n,4,4) float64 CUDA tensor -> rgbA uint8 (n,176,176,3), depthA uint16 (n,176,176) CUDA tensors (no sync)."""
        dev = self.engine.device
        n = int(poses.shape[0])
        if object_width is None:
            object_width = torch.full((n,), self.object_width, dtype=torch.float64, device=dev)
        if mesh_ids is None and self.mesh_id != 0:
            mesh_ids = torch.full((n,), self.mesh_id, dtype=torch.int32, device=dev)
        return self.engine.render(self.K, poses, object_width, mesh_ids, out_rgb, out_depth)

    def render_window(self, ob2cam):
        """Tracker.render_window's contract (predict.py:193-215): numpy in, numpy out."""
        p = torch.from_numpy(np.ascontiguousarray(ob2cam, dtype=np.float64)[None]).to(self.engine.device)
        rgb, dep = self.render_batch(p)
        return rgb[0].cpu().numpy(), dep[0].cpu().numpy()
