"""CAD-model file handling for the CUDA rasteriser (host logic, no GPU)."""
import importlib, os
import numpy as np
import pytest


@pytest.fixture(scope='module')
def mio():
    return importlib.import_module('iros20-6d-pose-tracking_b200.mesh_io')


@pytest.mark.parametrize('binary', [True, False])
def test_ply_round_trip(mio, synth, tmp_path, binary):
    m = synth.mesh(2, seed=3)
    p = str(tmp_path / 'model.ply')
    mio.save_ply_mesh(p, m, binary=binary)
    r = mio.load_ply_mesh(p)
    assert r['pos'].dtype == np.float32 and r['nrm'].dtype == np.float32 and r['col'].dtype == np.uint8 and r['faces'].dtype == np.int32
    assert np.array_equal(r['pos'], m['pos']) and np.array_equal(r['col'], m['col']) and np.array_equal(r['faces'], m['faces'])
    assert np.abs(r['nrm'] - m['nrm']).max() <= 2 ** -22          # the loader re-normalises in float32 (vispy_renderer.py:121)
    assert np.allclose(np.linalg.norm(r['nrm'], axis=1), 1.0, atol=1e-6)
    # the vertices-only reader Tracker uses for object_cloud sees the same points
    pr = importlib.import_module('iros20-6d-pose-tracking_b200.predict')
    assert np.array_equal(pr.load_vertices(p).astype(np.float32), m['pos'])


def test_ply_rejects_what_the_reference_rejects(mio, synth, tmp_path):
    m = synth.mesh(0)
    p = str(tmp_path / 'no_normals.ply')
    with open(p, 'w') as f:                                         # positions + faces only
        f.write('ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n'
                'element face %d\nproperty list uchar int vertex_indices\nend_header\n' % (len(m['pos']), len(m['faces'])))
        for v in m['pos']: f.write('%g %g %g\n' % tuple(v))
        for t in m['faces']: f.write('3 %d %d %d\n' % tuple(t))
    with pytest.raises(ValueError, match='nx'):
        mio.load_ply_mesh(p)
    q = str(tmp_path / 'quads.ply')
    with open(q, 'w') as f:
        f.write('ply\nformat ascii 1.0\nelement vertex 4\nproperty float x\nproperty float y\nproperty float z\nproperty float nx\n'
                'property float ny\nproperty float nz\nproperty uchar red\nproperty uchar green\nproperty uchar blue\n'
                'element face 1\nproperty list uchar int vertex_indices\nend_header\n')
        for i in range(4): f.write('%d %d 0 0 0 1 9 9 9\n' % (i & 1, i >> 1))
        f.write('4 0 1 3 2\n')
    with pytest.raises(ValueError):
        mio.load_ply_mesh(q)
    with pytest.raises(ValueError):
        mio.load_ply_mesh(__file__)


def test_synth_mesh_is_closed_and_outward(synth):
    m = synth.mesh(2)
    f = m['faces']
    edges = np.sort(np.concatenate((f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]])), axis=1)
    _, cnt = np.unique(edges, axis=0, return_counts=True)
    assert (cnt == 2).all() and len(f) == 20 * 4 ** 2                # watertight
    fn = np.cross(m['pos'][f[:, 1]] - m['pos'][f[:, 0]], m['pos'][f[:, 2]] - m['pos'][f[:, 0]])
    assert ((fn * m['pos'][f].mean(1)).sum(1) > 0).all()            # counter-clockwise seen from outside
    assert ((m['nrm'] * m['pos']).sum(1) > 0).all()


def test_sequence_layout_helpers(synth, tmp_path):
    """Host side of the headless sequence driver: file discovery, config loading, image readers (no GPU)."""
    import cv2, yaml
    pr = importlib.import_module('iros20-6d-pose-tracking_b200.predict')
    seq = tmp_path / 'seq'
    for d in ('rgb', 'depth_filled', 'annotated_poses'):
        (seq / d).mkdir(parents=True)
    with pytest.raises(FileNotFoundError):
        pr.sequence_files(str(seq))
    rgb, depth = synth.raw_frame(seed=1, h=48, w=64)
    for i in (1, 0):                                                # written out of order: the driver sorts
        cv2.imwrite(str(seq / 'rgb' / ('%04d.png' % i)), rgb[..., ::-1])
        cv2.imwrite(str(seq / 'depth_filled' / ('%04d.png' % i)), depth)
    pose = synth.raw_poses(1, seed=0)[0]
    np.savetxt(str(seq / 'annotated_poses' / '0000.txt'), pose)
    r, d, g = pr.sequence_files(str(seq))
    assert [os.path.basename(x) for x in r] == ['0000.png', '0001.png'] and len(d) == 2 and len(g) == 1
    assert np.array_equal(pr.read_rgb(r[0]), rgb) and np.array_equal(pr.read_depth(d[0]), depth) and pr.read_depth(d[0]).dtype == np.uint16
    assert np.array_equal(np.loadtxt(g[0]), pose)                   # np.savetxt's %.18e round-trips float64
    (tmp_path / 'train').mkdir()
    info = {'resolution': 176, 'object_width': 200.0, 'boundingbox': 10, 'camera': {'focalX': 1.0, 'focalY': 1.0, 'centerX': 0.0, 'centerY': 0.0, 'height': 48, 'width': 64}}
    yaml.safe_dump(info, open(tmp_path / 'dataset_info.yml', 'w'))
    mean, std = synth.default_mean_std()
    np.save(tmp_path / 'mean.npy', mean); np.save(tmp_path / 'std.npy', std)
    di, m, s = pr.load_run_config(str(tmp_path / 'train'), str(tmp_path))
    assert di == info and np.array_equal(m, mean) and np.array_equal(s, std)


def test_obj_loader_vertex_colours_polygons_and_texture(mio, tmp_path):
    """What the pyrender producer's model files look like (offscreen_renderer.py:57-60 loads them through trimesh):
    v/vt/vn index triples become unique vertices, polygons are fanned, colours come from `v x y z r g b`, else from the
    texture at each vertex's uv (nearest texel, rows flipped: predict.py:167-179), else from the material."""
    import cv2
    # a quad (fanned into two triangles) with per-vertex colours and no normals
    p = str(tmp_path / 'quad.obj')
    with open(p, 'w') as f:
        f.write('# quad\nv 0 0 0 1 0 0\nv 1 0 0 0 1 0\nv 1 1 0 0 0 1\nv 0 1 0 1 1 1\nf 1 2 3 4\n')
    m = mio.load_obj_mesh(p)
    assert m['pos'].shape == (4, 3) and m['pos'].dtype == np.float32 and m['faces'].dtype == np.int32
    assert m['faces'].tolist() == [[0, 1, 2], [0, 2, 3]]
    assert m['col'].tolist() == [[255, 0, 0], [0, 255, 0], [0, 0, 255], [255, 255, 255]]
    assert np.allclose(m['nrm'], [[0, 0, 1]] * 4)                                      # derived: the quad lies in z = 0, counter-clockwise
    # textured: the same position with two different uv's is two vertices; texel lookup with the v axis flipped
    tex = np.zeros((4, 8, 3), np.uint8)
    tex[0, 0] = (10, 20, 30); tex[3, 7] = (200, 100, 50); tex[3, 0] = (1, 2, 3)      # RGB; row 0 is the TOP of the image = v 1
    cv2.imwrite(str(tmp_path / 'tex.png'), tex[..., ::-1])
    with open(str(tmp_path / 'm.mtl'), 'w') as f:
        f.write('newmtl a\nKd 0.2 0.4 0.6\nmap_Kd tex.png\n')
    q = str(tmp_path / 'tri.obj')
    with open(q, 'w') as f:
        f.write('mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 1\nvt 1 0\nvt 0 0\nvn 0 0 1\nusemtl a\nf 1/1/1 2/2/1 3/3/1\nf 1/3/1 2/2/1 3/1/1\n')
    t = mio.load_obj_mesh(q)
    assert t['pos'].shape == (5, 3)                                                   # (v1,vt1) (v2,vt2) (v3,vt3) (v1,vt3) (v3,vt1); (v2,vt2) shared
    assert t['faces'].tolist() == [[0, 1, 2], [3, 1, 4]]
    assert t['col'][0].tolist() == [10, 20, 30] and t['col'][1].tolist() == [200, 100, 50] and t['col'][2].tolist() == [1, 2, 3]
    assert t['col'][3].tolist() == [1, 2, 3] and t['col'][4].tolist() == [10, 20, 30]
    assert np.allclose(t['nrm'], [[0, 0, 1]] * 5)
    # no texture file: the material's diffuse colour; no material at all: mid grey
    os.remove(str(tmp_path / 'tex.png'))
    assert mio.load_obj_mesh(q)['col'][0].tolist() == [51, 102, 153]
    os.remove(str(tmp_path / 'm.mtl'))
    assert mio.load_obj_mesh(q)['col'][0].tolist() == [128, 128, 128]
    assert mio.load_mesh(q)['faces'].shape == (2, 3)
    with pytest.raises(ValueError):
        open(str(tmp_path / 'empty.obj'), 'w').write('# nothing\n')
        mio.load_obj_mesh(str(tmp_path / 'empty.obj'))
