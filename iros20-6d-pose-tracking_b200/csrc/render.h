// Input A on the GPU: the object's CAD model rasterised at the previous pose into the 176 x 176 crop window (see render.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
namespace se3tn {
struct MeshDev {            // one CAD model in device memory (what vispy_renderer.py:108-129 uploads as vertex / index buffers)
    const float* pos;       // [nv][3] metres, object frame
    const float* nrm;       // [nv][3] unit normals
    const uint8_t* col;     // [nv][3] 8-bit colours
    const int* faces;       // [nf][3]
    int nv, nf;
};
struct RenderArgs {
    const double* poses;           // [n][16] object in OpenCV camera
    const double* object_width;    // [n] mm
    const int* mesh_ids;           // [n] or null (mesh 0)
    const MeshDev* meshes;         // device table indexed by mesh id
    int n_meshes;
    double fx, fy, cx, cy;
    uint8_t* projected;            // workspace [n][max_nv] projected vertices (render_projected_bytes_per_vertex() each)
    uint8_t* uniforms;             // workspace [n] per-track uniforms (render_uniform_bytes() each)
    int max_nv;                    // vertex count of the largest model
    int mode;                      // 0: vispy-style (lit, the crop window is the viewport); 1: pyrender-style (unlit full camera image vw x vh, then crop_bbox)
    int vw, vh;                    // mode 1: camera image size
    uint8_t* rgb;                  // [n][176][176][3]
    uint16_t* depth;               // [n][176][176] mm, 0 = background
};
size_t render_uniform_bytes();
size_t render_projected_bytes_per_vertex();
cudaError_t launch_render(const RenderArgs& a, int n, cudaStream_t s);
}  // namespace se3tn
