"""CAD-model files for the CUDA rasteriser: the .ply layout the reference's VispyRenderer reads with plyfile
(reference vispy_renderer.py:108-121): vertex x y z nx ny nz red green blue, faces as `list uchar int vertex_indices`,
ascii or binary_little_endian; and the Wavefront .obj the reference's pyrender path takes (offscreen_renderer.py:57-60 through trimesh).
No third-party mesh package is needed (plyfile / trimesh are not in this image)."""
import os
import numpy as np

_NP = {'float': '<f4', 'float32': '<f4', 'double': '<f8', 'float64': '<f8', 'uchar': 'u1', 'uint8': 'u1', 'char': 'i1', 'int8': 'i1',
       'int': '<i4', 'int32': '<i4', 'uint': '<u4', 'uint32': '<u4', 'short': '<i2', 'int16': '<i2', 'ushort': '<u2', 'uint16': '<u2'}


def load_ply_mesh(path):
    """-> dict(pos float32 (nv,3), nrm float32 (nv,3) re-normalised in float32 as vispy_renderer.py:121 does,
    col uint8 (nv,3), faces int32 (nf,3)).  Raises ValueError when normals, colours or triangular faces are missing
    (the reference asserts / fails on the same files)."""
    with open(path, 'rb') as f:
        if f.readline().strip() != b'ply':
            raise ValueError('not a ply file: ' + path)
        fmt, elements = None, []
        while True:
            line = f.readline()
            if not line:
                raise ValueError('ply header without end_header: ' + path)
            tok = line.decode('ascii', 'replace').split()
            if not tok or tok[0] == 'comment':
                continue
            if tok[0] == 'format':
                fmt = tok[1]
            elif tok[0] == 'element':
                elements.append({'name': tok[1], 'count': int(tok[2]), 'props': []})
            elif tok[0] == 'property':
                if tok[1] == 'list':
                    elements[-1]['props'].append(('list', tok[2], tok[3], tok[4]))
                else:
                    elements[-1]['props'].append(('scalar', tok[1], tok[2]))
            elif tok[0] == 'end_header':
                break
        if fmt not in ('ascii', 'binary_little_endian'):
            raise ValueError('unsupported ply format %r' % fmt)
        vert = faces = None
        for el in elements:
            n = el['count']
            scalars = all(p[0] == 'scalar' for p in el['props'])
            if fmt == 'ascii':
                rows = [f.readline().split() for _ in range(n)]
                if el['name'] == 'vertex':
                    names = [p[2] for p in el['props']]
                    arr = np.array(rows, dtype=np.float64).reshape(n, len(names))
                    vert = {nm: arr[:, i] for i, nm in enumerate(names)}
                elif el['name'] == 'face':
                    faces = np.array([[int(v) for v in r[1:1 + int(r[0])]] for r in rows], dtype=np.int64)
            else:
                if scalars:
                    dt = np.dtype([(p[2], _NP[p[1]]) for p in el['props']])
                    data = np.frombuffer(f.read(dt.itemsize * n), dtype=dt, count=n)
                    if el['name'] == 'vertex':
                        vert = {nm: data[nm] for nm in dt.names}
                elif el['name'] == 'face' and len(el['props']) == 1:
                    _, ct, it, _ = el['props'][0]
                    cdt, idt = np.dtype(_NP[ct]), np.dtype(_NP[it])
                    raw = f.read()
                    rec = np.dtype([('n', cdt), ('v', idt, (3,))])
                    if len(raw) < rec.itemsize * n:
                        raise ValueError('truncated face list')
                    fa = np.frombuffer(raw[:rec.itemsize * n], dtype=rec, count=n)
                    if not np.all(fa['n'] == 3):
                        raise ValueError('only triangular faces are supported (vispy_renderer.py:117 asserts the same)')
                    faces = fa['v'].astype(np.int64)
                else:
                    raise ValueError('unsupported ply element layout: ' + el['name'])
    if vert is None or faces is None or faces.ndim != 2 or faces.shape[1] != 3:
        raise ValueError('ply needs a vertex element and triangular faces')
    for need in ('x', 'y', 'z', 'nx', 'ny', 'nz', 'red', 'green', 'blue'):
        if need not in vert:
            raise ValueError('ply vertex property %r missing (the reference reads it, vispy_renderer.py:110-120)' % need)
    pos = np.stack((vert['x'], vert['y'], vert['z']), -1).astype(np.float32)
    nrm = np.stack((vert['nx'], vert['ny'], vert['nz']), -1).astype(np.float32)
    nrm = nrm / np.linalg.norm(nrm, axis=1).reshape(-1, 1)
    col = np.stack((vert['red'], vert['green'], vert['blue']), -1).astype(np.uint8)
    return dict(pos=pos, nrm=nrm.astype(np.float32), col=col, faces=faces.astype(np.int32))


def save_ply_mesh(path, mesh, binary=True):
    """Write the layout load_ply_mesh (and the reference's plyfile code) reads."""
    pos, nrm, col, faces = (np.asarray(mesh[k]) for k in ('pos', 'nrm', 'col', 'faces'))
    nv, nf = len(pos), len(faces)
    hdr = ['ply', 'format %s 1.0' % ('binary_little_endian' if binary else 'ascii'), 'element vertex %d' % nv,
           'property float x', 'property float y', 'property float z', 'property float nx', 'property float ny', 'property float nz',
           'property uchar red', 'property uchar green', 'property uchar blue', 'element face %d' % nf,
           'property list uchar int vertex_indices', 'end_header']
    with open(path, 'wb') as f:
        f.write(('\n'.join(hdr) + '\n').encode('ascii'))
        if binary:
            v = np.empty(nv, dtype=[('p', '<f4', (3,)), ('n', '<f4', (3,)), ('c', 'u1', (3,))])
            v['p'], v['n'], v['c'] = pos, nrm, col
            f.write(v.tobytes())
            fa = np.empty(nf, dtype=[('n', 'u1'), ('v', '<i4', (3,))])
            fa['n'], fa['v'] = 3, faces
            f.write(fa.tobytes())
        else:
            for i in range(nv):
                f.write(('%.9g %.9g %.9g %.9g %.9g %.9g %d %d %d\n' % (*pos[i], *nrm[i], *col[i])).encode('ascii'))
            for t in faces:
                f.write(('3 %d %d %d\n' % tuple(t)).encode('ascii'))


def _vertex_normals(pos, faces):
    """Area-weighted vertex normals (what trimesh / pyrender derive for a mesh that has none)."""
    p = pos.astype(np.float64)
    fn = np.cross(p[faces[:, 1]] - p[faces[:, 0]], p[faces[:, 2]] - p[faces[:, 0]])
    vn = np.zeros_like(p)
    for k in range(3):
        np.add.at(vn, faces[:, k], fn)
    ln = np.linalg.norm(vn, axis=1).reshape(-1, 1)
    return (vn / np.where(ln > 0, ln, 1.0)).astype(np.float32)


def load_obj_mesh(path):
    """Wavefront .obj -> the same dict as load_ply_mesh.  Vertices are unique (v, vt, vn) index triples, as trimesh.load builds
    them; polygons are fanned into triangles.  Colours, in this order: per-vertex `v x y z r g b`; the texture named by the
    .mtl's map_Kd, looked up ONCE per vertex at its uv (nearest texel, rows flipped) -- the conversion the reference itself
    sketches for its vispy path (predict.py:167-179) -- because the rasteriser interpolates vertex colours and does not sample
    textures per fragment; else the material's Kd; else mid grey."""
    base = os.path.dirname(os.path.abspath(path))
    v, vc, vt, vn, corners, faces = [], [], [], [], {}, []
    mtllib = None
    with open(path, 'r', errors='replace') as f:
        for line in f:
            tok = line.split()
            if not tok or tok[0].startswith('#'):
                continue
            if tok[0] == 'v':
                v.append([float(t) for t in tok[1:4]])
                vc.append([float(t) for t in tok[4:7]] if len(tok) >= 7 else None)
            elif tok[0] == 'vt':
                vt.append([float(t) for t in tok[1:3]])
            elif tok[0] == 'vn':
                vn.append([float(t) for t in tok[1:4]])
            elif tok[0] == 'mtllib' and len(tok) > 1:
                mtllib = line.split(None, 1)[1].strip()
            elif tok[0] == 'f':
                ids = []
                for c in tok[1:]:
                    parts = (c.split('/') + ['', ''])[:3]
                    key = tuple((int(q) - 1 if int(q) > 0 else int(q)) if q else None for q in parts)
                    key = (key[0] % len(v), None if key[1] is None else key[1] % len(vt), None if key[2] is None else key[2] % len(vn))
                    if key not in corners:
                        corners[key] = len(corners)
                    ids.append(corners[key])
                for k in range(1, len(ids) - 1):
                    faces.append((ids[0], ids[k], ids[k + 1]))
    if not v or not faces:
        raise ValueError('obj without vertices or faces: ' + path)
    keys = sorted(corners, key=corners.get)
    pos = np.array([v[k[0]] for k in keys], dtype=np.float32)
    faces = np.array(faces, dtype=np.int32)
    if all(k[2] is not None for k in keys):
        nrm = np.array([vn[k[2]] for k in keys], dtype=np.float32)
        ln = np.linalg.norm(nrm, axis=1).reshape(-1, 1)
        nrm = (nrm / np.where(ln > 0, ln, 1.0)).astype(np.float32)
    else:
        nrm = _vertex_normals(pos, faces)
    col = None
    if all(vc[k[0]] is not None for k in keys):
        col = np.rint(np.clip(np.array([vc[k[0]] for k in keys]), 0.0, 1.0) * 255.0).astype(np.uint8)
    kd, tex = None, None
    if col is None and mtllib and os.path.exists(os.path.join(base, mtllib)):
        with open(os.path.join(base, mtllib), 'r', errors='replace') as f:
            for line in f:
                tok = line.split()
                if tok and tok[0] == 'Kd' and len(tok) >= 4 and kd is None:
                    kd = [float(t) for t in tok[1:4]]
                elif tok and tok[0] == 'map_Kd' and tex is None:
                    tex = os.path.join(base, line.split(None, 1)[1].strip().split()[-1])
    if col is None and tex is not None and os.path.exists(tex) and all(k[1] is not None for k in keys):
        import cv2
        img = cv2.imread(tex, cv2.IMREAD_COLOR)
        if img is not None:
            img = img[..., ::-1]                                                        # BGR -> RGB
            th, tw = img.shape[:2]
            uv = np.array([vt[k[1]] for k in keys], dtype=np.float64)
            uv = np.where((uv < 0.0) | (uv > 1.0), uv - np.floor(uv), uv)               # repeat wrapping outside the unit square only
            px = np.rint(uv * np.array([tw - 1, th - 1])).astype(int)
            col = np.ascontiguousarray(img[::-1][px[:, 1], px[:, 0]]).astype(np.uint8)
    if col is None:
        c = np.rint(np.clip(np.array(kd if kd is not None else [0.5, 0.5, 0.5]), 0.0, 1.0) * 255.0).astype(np.uint8)
        col = np.tile(c, (len(pos), 1))
    return dict(pos=pos, nrm=nrm, col=col, faces=faces)


def load_mesh(path):
    """By extension: .ply (load_ply_mesh) or .obj (load_obj_mesh)."""
    return load_obj_mesh(path) if str(path).lower().endswith('.obj') else load_ply_mesh(path)
