#!/usr/bin/env python
"""Generate tests/golden/*.npz by running THE REFERENCE'S OWN CODE in the build
container (the only place /root/reference exists).

  python oracle/make_golden.py [--ref /root/reference]

What is executed from the reference, unmodified:
  se3_tracknet.Se3TrackNet (+ network_modules)        -> model fixtures
  Utils.compute_bbox / crop_bbox / normalize_rotation_matrix
  data_augmentation.OffsetDepth / NormalizeChannels / ToTensor, Utils.Compose
  datasets.TrackDataset.processData / processPredict  -> pre/post fixtures
  Utils.add / Utils.adi, eval_ycb.VOCap                -> metric fixtures
  vispy_renderer.VispyRenderer.update_cam_mat / render_image (numpy parts) -> renderer uniform fixtures
  Utils.fill_depth (as predict_ros.py:38-41 calls it)  -> depth hole-filling fixtures

Accommodations (nothing in the reference tree is edited; it is read-only):
  * `open3d` and `transformations` are not installed here; they are only
    imported, never used, on this path -> empty stub modules in sys.modules.
  * `np.float` was removed in numpy 1.24 (Utils.py:307,330 use it) -> aliased
    to the builtin float, which is what it always was.
  * PYTHONDONTWRITEBYTECODE: the mount is read-only.
Inputs come from the product's deterministic generators (synth.py) so tests can
rebuild them anywhere; small inputs are also stored in the fixture itself.
"""
import argparse, hashlib, importlib, os, sys, types
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
synth = importlib.import_module('iros20-6d-pose-tracking_b200.synth')


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def import_reference(ref):
    for name in ('open3d', 'transformations'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    if not hasattr(np, 'float'):
        np.float = float
    sys.path.insert(0, ref)
    import Utils, data_augmentation, datasets, se3_tracknet   # noqa: reference modules
    return Utils, data_augmentation, datasets, se3_tracknet


def small_frame(seed, h=120, w=160):
    rgb, depth = synth.raw_frame(seed, h, w)
    return rgb, depth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ref', default='/root/reference')
    ap.add_argument('--out', default=os.path.join(ROOT, 'tests', 'golden'))
    args = ap.parse_args()
    U, DA, DS, NET = import_reference(args.ref)
    os.makedirs(args.out, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    # ------------------------------------------------------------------ model
    sd = synth.make_state_dict(0)
    model = NET.Se3TrackNet(image_size=176)
    model.load_state_dict(sd)
    model.eval()
    A, B = synth.tensor_pairs(2, seed=0)
    acts = {}
    hooks = []
    for name in ['convA1', 'poolA1', 'convA2', 'convB1', 'poolB1', 'convB2', 'convB3', 'convAB1',
                 'convAB2', 'trans_conv1', 'trans_conv2', 'rot_conv1', 'rot_conv2']:
        hooks.append(getattr(model, name).register_forward_hook(
            lambda m, i, o, name=name: acts.__setitem__(name, o.detach().clone())))
    with torch.no_grad():
        out = model(A, B)
    for h in hooks:
        h.remove()
    g = dict(trans=out['trans'].numpy(), rot=out['rot'].numpy(),
             feature_sub=out['feature'][:, ::16, ::3, ::3].numpy().copy(),
             feature_sha=sha(out['feature'].numpy()))
    for k, v in acts.items():
        g['act_' + k + '_sub'] = v[:, ::8, ::5, ::5].numpy().copy()
        g['act_' + k + '_absmean'] = np.float64(v.abs().double().mean().item())

    # config 1: the shipped RGB pair + synthesised depth/pose/mean/std
    import cv2
    rgbA = cv2.imread(os.path.join(args.ref, 'media', '0000000rgbA.png'))[..., ::-1].copy()
    rgbB = cv2.imread(os.path.join(args.ref, 'media', '0000000rgbB.png'))[..., ::-1].copy()
    depthA, depthB = synth.depth_from_rgb(rgbA), synth.depth_from_rgb(rgbB)
    mean, std = synth.default_mean_std()
    pose = synth.config1_pose()
    post = U.Compose([DA.OffsetDepth(), DA.NormalizeChannels(mean, std), DA.ToTensor()])
    ds = DS.TrackDataset('', 'eval', mean, std, None, None, post, None,
                         trans_normalizer=0.03, rot_normalizer=5 * np.pi / 180)
    sample = ds.processData(rgbA, depthA, pose, rgbB, depthB, np.eye(4))[0]
    with torch.no_grad():
        o1 = model(sample[0].unsqueeze(0).float(), sample[1].unsqueeze(0).float())
    g['c1_trans'] = o1['trans'].numpy(); g['c1_rot'] = o1['rot'].numpy()
    g['c1_dataA_sha'] = sha(sample[0].numpy()); g['c1_dataB_sha'] = sha(sample[1].numpy())
    g['c1_pose_out'] = ds.processPredict(pose, (o1['trans'][0].numpy(), o1['rot'][0].numpy()))
    np.savez_compressed(os.path.join(args.out, 'golden_model.npz'), **g)
    cv2.imwrite(os.path.join(args.out, 'c1_rgbA.png'), rgbA[..., ::-1])
    cv2.imwrite(os.path.join(args.out, 'c1_rgbB.png'), rgbB[..., ::-1])

    # --------------------------------------------------------- pre-processing
    p = {}
    K_small = synth.CAMERA_K.copy(); K_small[:2] *= 0.25       # 120x160 frame
    rgb_s, depth_s = small_frame(3)
    p['small_rgb'] = rgb_s; p['small_depth'] = depth_s; p['K_small'] = K_small
    # poses chosen to hit: inside frame, clipped left/top, clipped right/bottom,
    # window larger than the frame, far object (tiny window -> upsampling)
    cases = [(0.0, 0.0, 0.7, 200.), (-0.09, -0.06, 0.5, 200.), (0.12, 0.08, 0.6, 230.),
             (0.0, 0.0, 0.3, 400.), (0.02, -0.01, 0.9, 60.), (0.033, 0.021, 0.8123, 187.3)]
    poses = synth.raw_poses(len(cases), seed=5)
    for i, (tx, ty, tz, ow) in enumerate(cases):
        poses[i, :3, 3] = (tx, ty, tz)
    p['poses'] = poses; p['object_width'] = np.array([c[3] for c in cases])
    rgbAs, depthAs = synth.rendered_views(len(cases), poses, seed=7)
    mean64 = mean.astype(np.float64) + 0.123; std64 = std.astype(np.float64) * 1.01
    for i in range(len(cases)):
        bb = U.compute_bbox(poses[i], K_small, p['object_width'][i], scale=(1000, 1000, 1000))
        rB, dB = U.crop_bbox(rgb_s, depth_s, bb, (176, 176))
        p[f'bb_{i}'] = bb; p[f'rgbB_sha_{i}'] = sha(rB); p[f'depthB_sha_{i}'] = sha(dB)
        if i < 2:
            p[f'rgbB_{i}'] = rB; p[f'depthB_{i}'] = dB
        gtB = np.eye(4); gtB[:3, :3] = synth.raw_poses(1, seed=40 + i)[0, :3, :3]; gtB[:3, 3] = poses[i, :3, 3] + 0.01
        for tag, (m_, s_) in {'f32': (mean, std), 'f64': (mean64, std64)}.items():
            post_i = U.Compose([DA.OffsetDepth(), DA.NormalizeChannels(m_, s_), DA.ToTensor()])
            ds_i = DS.TrackDataset('', 'eval', m_, s_, None, None, post_i, None)
            smp, lab = ds_i.processData(rgbAs[i], depthAs[i], poses[i].copy(), rB, dB, gtB.copy())[:2]
            p[f'dataA_sha_{tag}_{i}'] = sha(smp[0].numpy()); p[f'dataB_sha_{tag}_{i}'] = sha(smp[1].numpy())
            p[f'dataA_sub_{tag}_{i}'] = smp[0][:, ::11, ::11].numpy().copy()
            p[f'dataB_sub_{tag}_{i}'] = smp[1][:, ::11, ::11].numpy().copy()
            p[f'label_trans_{i}'] = np.asarray(lab[0]); p[f'label_rot_{i}'] = np.asarray(lab[1])
        p[f'gtB_{i}'] = gtB
    # full-size frame (regenerated from seed in the tests; only hashes stored)
    rgb_f, depth_f = synth.raw_frame(0)
    p['full_rgb_sha'] = sha(rgb_f); p['full_depth_sha'] = sha(depth_f)
    fposes = synth.raw_poses(8, seed=0)
    p['full_poses'] = fposes
    for i in range(8):
        bb = U.compute_bbox(fposes[i], synth.CAMERA_K, 200., scale=(1000, 1000, 1000))
        rB, dB = U.crop_bbox(rgb_f, depth_f, bb, (176, 176))
        p[f'full_bb_{i}'] = bb; p[f'full_rgbB_sha_{i}'] = sha(rB); p[f'full_depthB_sha_{i}'] = sha(dB)
    # render_window's GL-flavoured bbox (predict.py:202): scale=(1000,-1000,1000)
    p['bb_gl_0'] = U.compute_bbox(poses[1], K_small, 200., scale=(1000, -1000, 1000))

    # ------------------------------------------------------------ pose update
    rng = np.random.default_rng(11)
    n = 16
    pp = synth.raw_poses(n, seed=9)
    tr = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    ro = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    ro[0] = 0; tr[0] = 0                      # identity update
    ro[1] = (1e-9, 0, 0)                      # below cv2's small-angle threshold after scaling
    outs = []
    ds5 = DS.TrackDataset('', 'eval', mean, std, None, None, None, None)
    ds30 = DS.TrackDataset('', 'eval', mean, std, None, None, None, None, rot_normalizer=30 * np.pi / 180)
    p['pu_poses'] = pp; p['pu_trans'] = tr; p['pu_rot'] = ro
    p['pu_out_5deg'] = np.stack([ds5.processPredict(pp[i], (tr[i], ro[i])) for i in range(n)])
    p['pu_out_30deg'] = np.stack([ds30.processPredict(pp[i], (tr[i], ro[i])) for i in range(n)])
    Rn = pp[:, :3, :3] * rng.uniform(0.9, 1.1, (n, 1, 3))
    p['nrm_in'] = Rn.copy()
    p['nrm_out'] = np.stack([U.normalize_rotation_matrix(Rn[i].copy()) for i in range(n)])
    np.savez_compressed(os.path.join(args.out, 'golden_pre.npz'), **p)

    # ---------------------------------------------------------------- metrics
    # Utils.add / Utils.adi take an open3d PointCloud; open3d is not installed, so the reference functions run
    # on a stand-in with the three members they use (deepcopy, .transform(T): p -> R p + t, .points).
    # scipy removed cKDTree.query(n_jobs=) (Utils.py:96): `Utils.spatial` is swapped for a wrapper that drops the keyword.
    import copy
    from scipy import spatial

    class Cloud:
        def __init__(self, pts): self.points = np.array(pts, dtype=np.float64)
        def transform(self, T): self.points = self.points @ T[:3, :3].T + T[:3, 3]; return self

    class CompatTree:                          # scipy's cKDTree minus the removed n_jobs keyword
        def __init__(self, pts): self._t = spatial.cKDTree(pts)
        def query(self, x, k=1, n_jobs=None, **kw): return self._t.query(x, k=k, **kw)
    _spatial = U.spatial
    U.spatial = types.SimpleNamespace(cKDTree=CompatTree)
    import eval_ycb as EV                      # reference scorer (VOCap), unmodified
    mt = {}
    model = synth.model_points(2620, seed=0)
    pred, gt = synth.pose_pairs(12, seed=0)
    pred[0] = gt[0]                            # exact pose: ADD = ADD-S = 0
    mt['add'] = np.array([U.add(pred[i], gt[i], Cloud(model)) for i in range(12)])
    mt['adi'] = np.array([U.adi(pred[i], gt[i], Cloud(model)) for i in range(12)])
    rng = np.random.default_rng(21)
    curves = {'mixed': rng.uniform(0, 0.2, 500), 'all_below': rng.uniform(0, 0.09, 300),
              'dups': np.round(rng.uniform(0, 0.15, 400), 2), 'single': np.array([0.03]), 'sorted_add': np.sort(mt['add'])}
    for k, v in curves.items():
        mt['curve_' + k] = v
        mt['vocap_' + k] = np.float64(EV.VOCap(v))
    U.spatial = _spatial
    np.savez_compressed(os.path.join(args.out, 'golden_metrics.npz'), **mt)
    # ------------------------------------------------------------------ renderer uniforms
    # vispy / OpenGL / plyfile are not installed: stub modules let the reference's vispy_renderer.py import; the class's
    # pure-numpy methods (update_cam_mat :135-150, render_image :171-178 up to the draw call) are then run unbound on a
    # stand-in object.  The window (left/right/top/bottom) follows predict.py:202-207 on the reference's compute_bbox.
    for name in ('vispy', 'vispy.app', 'vispy.gloo', 'OpenGL', 'OpenGL.GL', 'plyfile'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['vispy'].app = sys.modules['vispy.app']; sys.modules['vispy'].gloo = sys.modules['vispy.gloo']
    sys.modules['vispy.app'].Canvas = object
    sys.modules['OpenGL'].GL = sys.modules['OpenGL.GL']
    sys.modules['plyfile'].PlyData = sys.modules['plyfile'].PlyElement = object
    import vispy_renderer as VR                                        # noqa: reference module

    class FakeRenderer:
        def __init__(self): self.program = {}; self.rgb = self.depth = None
        def update(self): pass
        def on_draw(self, ev): pass
    rd = {}
    K = synth.CAMERA_K
    poses = synth.raw_poses(6, seed=7)
    glcam_in_cvcam = np.array([[1, 0, 0, 0], [0, -1, 0, 0], [0, 0, -1, 0], [0, 0, 0, 1]])
    wins, projs, views, lights = [], [], [], []
    for ob2cam in poses:
        bbox = U.compute_bbox(ob2cam, K, 200.0, scale=(1000, -1000, 1000))
        ob2cam_gl = np.linalg.inv(glcam_in_cvcam).dot(ob2cam)
        left, right = np.min(bbox[:, 1]), np.max(bbox[:, 1]); top, bottom = np.min(bbox[:, 0]), np.max(bbox[:, 0])
        fr = FakeRenderer()
        VR.VispyRenderer.update_cam_mat(fr, K, left, right, bottom, top)
        VR.VispyRenderer.render_image(fr, ob2cam_gl)
        wins.append([left, right, top, bottom]); projs.append(np.asarray(fr.projection_matrix).T.copy())
        views.append(np.asarray(fr.program['view']).T.copy()); lights.append(np.asarray(fr.program['light_direction']))
    rd = dict(poses=poses, object_width=np.float64(200.0), window=np.array(wins, np.int64), proj64=np.array(projs),
              view=np.array(views), light32=np.array(lights))
    np.savez_compressed(os.path.join(args.out, 'golden_render.npz'), **rd)
    # ------------------------------------------------------------------ depth hole filling
    # the reference's own Utils.fill_depth exactly as predict_ros.py:38-41 drives it
    fd = {}
    for name, seed, hw in (('a', 31, (120, 160)), ('b', 32, (64, 96))):
        _, dmm = synth.raw_frame(seed, *hw)
        dmm = dmm.copy(); dmm[10:30, 20:50] = 0                   # a real hole, not only salt-and-pepper dropouts
        filled = U.fill_depth(dmm / 1e3, max_depth=2.0, extrapolate=False)
        fd['in_' + name] = dmm; fd['out_m_' + name] = filled; fd['out_mm_' + name] = (filled * 1000).astype(np.uint16)
        # the two optional branches (Utils.py:486-497, 506-510), alone and together
        for tag, ex, blur in (('ex', True, 'bilateral'), ('ga', False, 'gaussian'), ('exga', True, 'gaussian')):
            fd['out_m_%s_%s' % (name, tag)] = U.fill_depth(dmm / 1e3, max_depth=2.0, extrapolate=ex, blur_type=blur)
    np.savez_compressed(os.path.join(args.out, 'golden_fill.npz'), **fd)
    for f in sorted(os.listdir(args.out)):
        print(f, os.path.getsize(os.path.join(args.out, f)))


if __name__ == '__main__':
    main()
