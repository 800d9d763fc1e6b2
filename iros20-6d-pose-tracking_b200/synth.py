"""Deterministic synthetic weights / frames / poses for tests, smoke() and bench.py.

There is no network for checkpoints or YCB data (reference README.md:117-126), so
every workload in BASELINE.json is driven by the generators here.
Pure numpy/torch, no CUDA and no oracle imports: both the CUDA path and the
oracle consume the SAME arrays produced here.
"""
import math
import numpy as np
import torch

# (key prefix, Cout, Cin, k) in the order Se3TrackNet.__init__ creates them
# (reference se3_tracknet.py:57-78); 'cbr' = ConvBNReLU (conv '.0', bn '.1'),
# 'block' = ResnetBasicBlock (conv1/bn1/conv2/bn2), 'fc' = Linear in a Sequential.
ARCH = [
    ('convA1', 'cbr', 64, 4, 7), ('convA2', 'block', 64, 64, 3),
    ('convB1', 'cbr', 64, 4, 7), ('convB2', 'block', 64, 64, 3), ('convB3', 'block', 64, 64, 3),
    ('convAB1', 'cbr', 256, 128, 3), ('convAB2', 'block', 256, 256, 3),
    ('trans_conv1', 'cbr', 512, 256, 3), ('trans_conv2', 'block', 512, 512, 3), ('trans_out', 'fc', 3, 512, 0),
    ('rot_conv1', 'cbr', 512, 256, 3), ('rot_conv2', 'block', 512, 512, 3), ('rot_out', 'fc', 3, 512, 0),
]

CAMERA_K = np.array([[1066.778, 0.0, 312.9869],       # reference dataset_info.yml:1-7
                     [0.0, 1067.487, 241.3109],
                     [0.0, 0.0, 1.0]])
FRAME_H, FRAME_W = 480, 640
IMAGE_SIZE = 176


def make_state_dict(seed=0):
    """A reference-format state_dict (123 keys) with default-init-like conv/linear
    weights (U(-1/sqrt(fan_in), 1/sqrt(fan_in))) and non-trivial BN statistics
    (gamma~U(.5,1.5), beta~N(0,.1), mean~N(0,.1), var~U(.5,1.5)) so BN folding is
    actually exercised."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(prefix, cout, cin, k):
        bound = 1.0 / math.sqrt(cin * k * k)
        sd[prefix + '.weight'] = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound
        sd[prefix + '.bias'] = (torch.rand(cout, generator=g) * 2 - 1) * bound

    def bn(prefix, c):
        sd[prefix + '.weight'] = torch.rand(c, generator=g) + 0.5
        sd[prefix + '.bias'] = torch.randn(c, generator=g) * 0.1
        sd[prefix + '.running_mean'] = torch.randn(c, generator=g) * 0.1
        sd[prefix + '.running_var'] = torch.rand(c, generator=g) + 0.5
        sd[prefix + '.num_batches_tracked'] = torch.tensor(0, dtype=torch.int64)

    for name, kind, cout, cin, k in ARCH:
        if kind == 'cbr':
            conv(name + '.0', cout, cin, k); bn(name + '.1', cout)
        elif kind == 'block':
            conv(name + '.conv1', cout, cin, k); bn(name + '.bn1', cout)
            conv(name + '.conv2', cout, cout, k); bn(name + '.bn2', cout)
        else:
            bound = 1.0 / math.sqrt(cin)
            sd[name + '.0.weight'] = (torch.rand(cout, cin, generator=g) * 2 - 1) * bound
            sd[name + '.0.bias'] = (torch.rand(cout, generator=g) * 2 - 1) * bound
    assert len(sd) == 123
    return sd


def default_mean_std():
    """float32 8-vectors (A's 4 channels then B's)."""
    mean = np.array([40, 40, 40, 1500, 60, 60, 60, 1200], dtype=np.float32)
    std = np.array([5, 5, 5, 100, 5, 5, 5, 100], dtype=np.float32)
    return mean, std


def depth_from_rgb(rgb):
    """depth(mm,u16) = 700 + (luma-64)//2 where the pixel is
    non-black, else 0; luma = integer mean of the 3 channels."""
    s = rgb.astype(np.int32).sum(-1)
    luma = s // 3
    d = 700 + (luma - 64) // 2
    d[s == 0] = 0
    return d.astype(np.uint16)


def config1_pose():
    p = np.eye(4)
    p[2, 3] = 0.7
    return p


def tensor_pairs(n, seed=0):
    """A~N(0,1), B=A+0.1*N(0,1), float32 (n,4,176,176)."""
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(n, 4, IMAGE_SIZE, IMAGE_SIZE, generator=g)
    B = A + 0.1 * torch.randn(n, 4, IMAGE_SIZE, IMAGE_SIZE, generator=g)
    return A, B


def _random_rotations(rng, n):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.empty((n, 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - z * w); R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w); R[:, 2, 1] = 2 * (y * z + x * w); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def raw_frame(seed=0, h=FRAME_H, w=FRAME_W):
    """rgb U{0..255} u8; depth U{300..1800} u16 with 10 %
    zeros and 2 % > 2000."""
    rng = np.random.default_rng(seed)
    rgb = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    depth = rng.integers(300, 1801, size=(h, w)).astype(np.uint16)
    u = rng.random((h, w))
    depth[u < 0.10] = 0
    depth[(u >= 0.10) & (u < 0.12)] = rng.integers(2001, 4000, size=int(((u >= 0.10) & (u < 0.12)).sum())).astype(np.uint16)
    return rgb, depth


def raw_poses(n, seed=0):
    """t_x,t_y~U(-.15,.15), t_z~U(.4,.9) (reference dataset_info.yml:24-26 ranges,
    tightened in x/y so the crop window mostly overlaps the 480x640 frame), random R."""
    rng = np.random.default_rng(seed + 1)
    poses = np.tile(np.eye(4), (n, 1, 1))
    poses[:, :3, :3] = _random_rotations(rng, n)
    poses[:, 0, 3] = rng.uniform(-0.15, 0.15, n)
    poses[:, 1, 3] = rng.uniform(-0.15, 0.15, n)
    poses[:, 2, 3] = rng.uniform(0.4, 0.9, n)
    return poses


def rendered_views(n, poses, seed=0, size=IMAGE_SIZE):
    """Stand-in for Tracker.render_window (predict.py:193-215, OpenGL, out of scope):
    disk-masked noise, rgbA u8 (n,176,176,3), depthA u16 mm (n,176,176) near the
    object's depth, 0 outside the disk -- the renderer's output contract."""
    rng = np.random.default_rng(seed + 2)
    yy, xx = np.mgrid[0:size, 0:size]
    disk = ((yy - size / 2) ** 2 + (xx - size / 2) ** 2) < (0.4 * size) ** 2
    rgbA = rng.integers(1, 256, size=(n, size, size, 3), dtype=np.uint8)
    rgbA[:, ~disk] = 0
    z_mm = (poses[:, 2, 3] * 1000)[:, None, None]
    depthA = (z_mm + rng.integers(-40, 41, size=(n, size, size))).astype(np.uint16)
    depthA[:, ~disk] = 0
    return rgbA, depthA


def model_points(m=2620, seed=0):
    """A synthetic CAD-model point cloud (metres): m points on a bumpy ellipsoid about 10 x 7 x 5 cm, float64.
    (YCB `points.xyz` files hold 2620 points; reference eval_ycb.py:72-81.)"""
    rng = np.random.default_rng(seed + 3)
    d = rng.normal(size=(m, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = 1.0 + 0.15 * np.sin(7 * d[:, 0]) * np.cos(5 * d[:, 1])
    return d * r[:, None] * np.array([0.05, 0.035, 0.025])


def pose_pairs(n, seed=0, trans_noise=0.01, rot_noise_deg=5.0):
    """n (pred, gt) pairs: gt random in the tracking volume, pred = gt perturbed by a small rigid motion."""
    rng = np.random.default_rng(seed + 4)
    gt = raw_poses(n, seed=seed + 10)
    pred = gt.copy()
    w = rng.normal(size=(n, 3)); w /= np.linalg.norm(w, axis=1, keepdims=True)
    ang = np.deg2rad(rot_noise_deg) * rng.uniform(0, 1, n)
    for i in range(n):
        K = np.array([[0, -w[i, 2], w[i, 1]], [w[i, 2], 0, -w[i, 0]], [-w[i, 1], w[i, 0], 0]])
        R = np.eye(3) + np.sin(ang[i]) * K + (1 - np.cos(ang[i])) * K @ K
        pred[i, :3, :3] = R @ gt[i, :3, :3]
        pred[i, :3, 3] = gt[i, :3, 3] + rng.normal(size=3) * trans_noise
    return pred, gt


def mesh(level=3, seed=0):
    """A synthetic CAD model in the form the reference's renderer consumes (vispy_renderer.py:108-121: a .ply with per-vertex
    position, normal and 8-bit colour, triangular faces): an icosphere subdivided `level` times (20 * 4**level faces, outward
    counter-clockwise), pushed onto the bumpy ellipsoid of `model_points`, smooth procedural colours.
    -> dict(pos float32 (nv,3) metres, nrm float32 (nv,3) unit, col uint8 (nv,3), faces int32 (nf,3))."""
    t = (1.0 + 5 ** 0.5) / 2
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(level):
        cache, nf = {}, []
        def mid(a, b):
            k = (min(a, b), max(a, b))
            if k not in cache:
                m = v[a] + v[b]; v.append(m / np.linalg.norm(m)); cache[k] = len(v) - 1
            return cache[k]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    d = np.array(v); faces = np.array(f, dtype=np.int32)
    r = 1.0 + 0.15 * np.sin(7 * d[:, 0]) * np.cos(5 * d[:, 1])
    pos = (d * r[:, None] * np.array([0.05, 0.035, 0.025])).astype(np.float32)
    fn = np.cross(pos[faces[:, 1]] - pos[faces[:, 0]], pos[faces[:, 2]] - pos[faces[:, 0]]).astype(np.float64)   # area-weighted
    nrm = np.zeros((len(pos), 3))
    for k in range(3): np.add.at(nrm, faces[:, k], fn)
    nrm = nrm.astype(np.float32)
    nrm = nrm / np.linalg.norm(nrm, axis=1).reshape(-1, 1)          # float32, as vispy_renderer.py:121 does to the ply normals
    rng = np.random.default_rng(seed + 5)
    base = 128 + 100 * np.stack((np.sin(9 * d[:, 0] + 1), np.sin(7 * d[:, 1] + 2), np.sin(11 * d[:, 2] + 3)), 1)
    col = np.clip(base + rng.integers(-12, 13, size=base.shape), 0, 255).astype(np.uint8)
    return dict(pos=pos, nrm=nrm.astype(np.float32), col=col, faces=faces)
