"""Drop-in for the inference half of the reference's datasets.py: TrackDataset.processData and
.processPredict (reference datasets.py:115-175) with the same signatures and return structure,
executed by libse3tn kernels (K0 normalisation, K5 so(3) log, K6 pose update).

The training half (__getitem__, file lists, augmentations; datasets.py:50-112) is out of scope;
pretransforms / augmentations must be None exactly as Tracker passes them (predict.py:191).
"""
import numpy as np
import torch


class TrackDataset:
    def __init__(self, root, mode, images_mean, images_std, pretransforms=None, augmentations=None,
                 posttransforms=None, dataset_info=None, trans_normalizer=0.03, rot_normalizer=5 * np.pi / 180,
                 engine=None, weight_id=0, precision='bf16x3'):
        if pretransforms is not None or augmentations is not None:
            raise NotImplementedError('train-time transforms are out of scope (inference passes None, predict.py:191)')
        self.root = root
        self.mode = mode
        self.images_mean = np.asarray(images_mean)
        self.images_std = np.asarray(images_std)
        self.pretransforms = None
        self.augmentations = None
        # The reference composes OffsetDepth -> NormalizeChannels -> ToTensor here; that chain is
        # what the K0 kernel implements, so the object passed in is only kept for introspection.
        self.posttransforms = posttransforms
        self.dataset_info = dataset_info
        if dataset_info is not None:
            cam = dataset_info['camera']
            self.cam_K = np.array([[cam['focalX'], 0, cam['centerX']], [0, cam['focalY'], cam['centerY']], [0, 0, 1]])
        self.trans_normalizer = trans_normalizer
        self.rot_normalizer = rot_normalizer
        self.weight_id = weight_id
        self.precision = precision
        self._engine = engine
        self._stats_set = False

    def __len__(self):
        return 0

    @property
    def engine(self):
        if self._engine is None:
            from .engine import Engine
            self._engine = Engine(max_batch=1)
        if not self._stats_set:
            self._engine.set_stats(self.images_mean, self.images_std, self.weight_id)
            self._stats_set = True
        return self._engine

    def processData(self, rgbA, depthA, A_in_cam, rgbB, depthB, B_in_cam, maskB=None, original_size=None):
        """-> (sample=[dataA, dataB] float32 CPU tensors (4,H,W), [trans_label, rot_label],
               rgbA_viz, rgbB_viz, maskA, maskB)   -- reference datasets.py:115-156."""
        eng = self.engine
        dev = eng.device
        maskA = (depthA > 100).astype(np.uint8)
        if maskB is None:
            maskB = (depthB > 100).astype(np.uint8)
        A_pose = torch.from_numpy(np.ascontiguousarray(A_in_cam, dtype=np.float64).reshape(1, 4, 4)).to(dev)
        B_pose = torch.from_numpy(np.ascontiguousarray(B_in_cam, dtype=np.float64).reshape(1, 4, 4)).to(dev)
        wid = torch.tensor([self.weight_id], dtype=torch.int32, device=dev)
        tA, tB = eng.normalize(self._u8(rgbA, dev), self._u16(depthA, dev), self._u8(rgbB, dev), self._u16(depthB, dev),
                               A_pose, weight_ids=wid, precision=self.precision, want_tensors=True)
        tl, rl = eng.so3_log(A_pose, B_pose, self.trans_normalizer, self.rot_normalizer)
        sample = [tA[0].cpu(), tB[0].cpu()]
        trans_label, rot_label = tl[0].cpu().numpy(), rl[0].cpu().numpy()
        if self.mode == 'train':
            assert (trans_label <= 1).all() and (trans_label >= -1).all()
            assert (rot_label >= -1).all() and (rot_label <= 1).all()
        return sample, [trans_label, rot_label], rgbA.astype(np.uint8), rgbB.astype(np.uint8), maskA, maskB

    def processPredict(self, A_in_cam, predB, original_size=None):
        """-> 4x4 float64 object pose in the camera frame -- reference datasets.py:159-175."""
        eng = self.engine
        dev = eng.device
        poses = torch.from_numpy(np.ascontiguousarray(A_in_cam, dtype=np.float64).reshape(1, 4, 4)).to(dev)
        trans = torch.as_tensor(np.asarray(predB[0], dtype=np.float32).reshape(1, 3)).to(dev)
        rot = torch.as_tensor(np.asarray(predB[1], dtype=np.float32).reshape(1, 3)).to(dev)
        return eng.pose_update(poses, trans, rot, self.trans_normalizer, self.rot_normalizer)[0].cpu().numpy()

    @staticmethod
    def _u8(a, dev):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint8)[None]).to(dev)

    @staticmethod
    def _u16(a, dev):
        return torch.from_numpy(np.ascontiguousarray(a).astype(np.uint16)[None]).to(dev)
