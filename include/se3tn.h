/* libse3tn -- C ABI of the B200-native se(3)-TrackNet inference hot path.
 *
 * The upstream reference (wenbowen123/iros20-6d-pose-tracking @ 18dc5bac) is pure Python and has
 * no FFI of its own; its boundary is the Python class surface.  Each entry
 * point below is what a ctypes binding on the reference side would call INSTEAD of the cited
 * reference code.  All tensor arguments are plain device pointers owned by the caller (PyTorch in
 * the shipped host layer); every launch is enqueued on the caller's cudaStream_t (pass it as a
 * void*); nothing here synchronises the stream unless stated.  No exceptions or aborts cross the
 * boundary: every function returns SE3TN_OK or a negative code, text via se3tn_last_error().
 * A context is single-threaded; distinct contexts are independent.
 */
#ifndef SE3TN_H
#define SE3TN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct se3tn_ctx se3tn_ctx;

enum {
    SE3TN_OK = 0,
    SE3TN_ERR_INVALID = -1,      /* bad argument (null, size, alignment, unknown id) */
    SE3TN_ERR_CUDA = -2,         /* a CUDA runtime / driver call failed               */
    SE3TN_ERR_NOMEM = -3,
    SE3TN_ERR_STATE = -4,        /* e.g. forward before load_weights                  */
    SE3TN_ERR_UNSUPPORTED = -5   /* device is not sm_100                              */
};

/* Arithmetic of the 17 convolutions (accumulation is always fp32). */
enum {
    SE3TN_PREC_TF32 = 0,    /* tcgen05 kind::tf32; operands rounded to tf32 (rna): 10-bit mantissas.  Fastest;
                               meets the 1e-3/1e-4 gate only for well-conditioned weights/inputs              */
    SE3TN_PREC_FP32 = 1,    /* plain FFMA direct convolution, no operand rounding (cross-check mode)         */
    SE3TN_PREC_BF16X3 = 2,  /* tcgen05 kind::f16 on bf16 hi/lo splits, 3 products per MAC: ~2^-16 relative
                               error (fp32-faithful for the gate) at 1.5x the tensor time of TF32            */
    SE3TN_PREC_BF16 = 3     /* tcgen05 kind::f16, bf16 operands, 1 product per MAC (BASELINE configs[2])     */
};

#define SE3TN_IMAGE_SIZE 176           /* reference dataset_info.yml:15 `resolution`          */
#define SE3TN_WEIGHT_BLOB_FLOATS 13528326u  /* see se3tn_load_weights                          */

/* ---- lifetime ------------------------------------------------------------------------------ */

/* Bytes of device workspace a context for batches of up to `max_batch` pairs needs. */
size_t se3tn_workspace_bytes(int max_batch);

/* Create a context on CUDA device `device` for up to `max_batch` RGB-D pairs per call.
 * `workspace` is a caller-owned device buffer of se3tn_workspace_bytes(max_batch) bytes, 1024-byte
 * aligned, that must outlive the context; pass NULL to let the library cudaMalloc its own.
 * Replaces: the `.cuda()` model placement in Tracker.__init__ (reference predict.py:156-158). */
int se3tn_create(int device, int max_batch, void* workspace, se3tn_ctx** out);
void se3tn_destroy(se3tn_ctx* ctx);

/* Message of the last error on `ctx` (or of the last failed se3tn_create when ctx is NULL). */
const char* se3tn_last_error(se3tn_ctx* ctx);

/* ---- per-object parameters --------------------------------------------------------------------
 * One weight set per object class (reference README.md:132: one checkpoint + mean/std per object).
 * `blob` is HOST memory: SE3TN_WEIGHT_BLOB_FLOATS float32 values produced by the packer
 * (weights.py: eval-mode BatchNorm folded into each conv, OIHW -> [Cout][tap*Cin + c] K-major),
 * in launch order:
 *   W[64][224],b[64]            x2   convA1, convB1        (7x7 s2; K = 7 rows x (8 px x 4 ch))
 *   W[64][576],b[64]            x6   convA2.{conv1,conv2}, convB2.{..}, convB3.{..}
 *   W[256][1152],b[256]              convAB1
 *   W[256][2304],b[256]         x2   convAB2.{conv1,conv2}
 *   W[1024][2304],b[1024]            trans_conv1 ++ rot_conv1 (Cout concatenated)
 *   W[1024][4608],b[1024]       x2   {trans,rot}_conv2.conv1, {trans,rot}_conv2.conv2 (2 groups)
 *   W[6][512],b[6]                   trans_out.0 ++ rot_out.0
 * Replaces: Se3TrackNet.load_state_dict (reference predict.py:151-155). */
int se3tn_load_weights(se3tn_ctx* ctx, int weight_id, const float* blob, size_t n_floats);

/* Per-object channel statistics (reference predict.py:657-658 mean.npy/std.npy): 8 values each,
 * A's 4 channels then B's.  `is_f64` selects the arithmetic of the normalisation so that it
 * reproduces numpy's for float32 resp. float64 mean/std arrays (data_augmentation.py:159-163). */
int se3tn_set_stats(se3tn_ctx* ctx, int weight_id, const void* mean8, const void* std8, int is_f64);

/* ---- the hot path ---------------------------------------------------------------------------- */

/* K0.  For each of n tracks: bbox of the previous pose (reference Utils.py:302-316), zero-padded
 * window + nearest-neighbour resize of the observed frame to 176x176 (Utils.py:320-359), depth
 * offset/clip (data_augmentation.py:134-144), channel normalisation (:154-164) and packing
 * (:179-189).  Results land in the context's conv-input buffers; optional outputs (any may be
 * NULL): out_A/out_B float32 (n,4,176,176) exactly as TrackDataset.processData returns them
 * (datasets.py:136-137), crop_rgb uint8 (n,176,176,3) / crop_depth uint16 (n,176,176) exactly
 * as crop_bbox returns them.
 *   frame_rgb  uint8  (H,W,3) device      frame_depth uint16 (H,W) device, millimetres
 *   K          4 doubles HOST: fx, fy, cx, cy
 *   poses      double (n,16) device, row-major 4x4 object-in-camera, metres
 *   object_width double (n) device, millimetres (Tracker.object_width, predict.py:136-142)
 *   rgbA uint8 (n,176,176,3), depthA uint16 (n,176,176) device: render_window's output contract
 *   weight_ids int32 (n) device or NULL (all 0): which mean/std row each track uses */
int se3tn_preprocess(se3tn_ctx* ctx, const uint8_t* frame_rgb, const uint16_t* frame_depth, int H, int W,
                     const double* K, const double* poses, const double* object_width,
                     const uint8_t* rgbA, const uint16_t* depthA, const int32_t* weight_ids, int n,
                     int precision, float* out_A, float* out_B, uint8_t* crop_rgb, uint16_t* crop_depth,
                     void* stream);

/* The post-transform half of TrackDataset.processData (reference datasets.py:136-137 ->
 * data_augmentation.py:124-196) on crops that already exist: rgbA/rgbB uint8 (n,176,176,3),
 * depthA/depthB uint16 (n,176,176), poses double (n,16) (A's pose: both depths are offset by its z).
 * out_A/out_B float32 (n,4,176,176) or both NULL; the conv-input buffers are filled either way. */
int se3tn_normalize(se3tn_ctx* ctx, const uint8_t* rgbA, const uint16_t* depthA,
                    const uint8_t* rgbB, const uint16_t* depthB, const double* poses,
                    const int32_t* weight_ids, int n, int precision, float* out_A, float* out_B, void* stream);

/* compute_bbox (reference Utils.py:302-316): out_bbox int32 (n,4,2), rows (x-,y-),(x-,y+),(x+,y-),(x+,y+),
 * columns (v,u).  K: 4 doubles HOST fx,fy,cx,cy; scale: 3 doubles HOST (the reference passes
 * (1000,1000,1000), or (1000,-1000,1000) for its GL renderer, predict.py:202,232). */
int se3tn_compute_bbox(se3tn_ctx* ctx, const double* poses, const double* K, const double* widths,
                       const double* scale, int32_t* out_bbox, int n, void* stream);

/* crop_bbox (reference Utils.py:320-359): zero-padded window [min v, max v) x [min u, max u) of the
 * frame, cv2.INTER_NEAREST-resized to (out_h, out_w).  bbox int32 (n,4,2) device. */
int se3tn_crop_bbox(se3tn_ctx* ctx, const uint8_t* frame_rgb, const uint16_t* frame_depth, int H, int W,
                    const int32_t* bbox, int n, int out_h, int out_w,
                    uint8_t* crop_rgb, uint16_t* crop_depth, void* stream);

/* Se3TrackNet.forward (reference se3_tracknet.py:81-112).  A, B: float32 (n,4,176,176) contiguous
 * NCHW device tensors.  out_trans/out_rot: float32 (n,3).  out_feature: float32 (n,256,22,22) or
 * NULL.  All n pairs use weight set `weight_id`. */
int se3tn_forward(se3tn_ctx* ctx, int weight_id, const float* A, const float* B, int n,
                  float* out_trans, float* out_rot, float* out_feature, int precision, void* stream);

/* The conv stack on whatever se3tn_preprocess left in the conv-input buffers (tracks
 * [first, first+n) of the last preprocess call). */
int se3tn_forward_preprocessed(se3tn_ctx* ctx, int weight_id, int first, int n,
                               float* out_trans, float* out_rot, float* out_feature, int precision, void* stream);

/* K6.  TrackDataset.processPredict (reference datasets.py:159-175): t' = t + trans*tn,
 * R' = Rodrigues(rot*rn) . R with the reference's float32/float64 dtype chain.
 * poses_in/poses_out double (n,16) device (may alias); trans/rot float32 (n,3) device. */
int se3tn_pose_update(se3tn_ctx* ctx, const double* poses_in, const float* trans, const float* rot,
                      double trans_normalizer, double rot_normalizer, double* poses_out, int n, void* stream);

/* K5.  The label half of TrackDataset.processData (reference datasets.py:141-150, Utils.py:363-367):
 * trans_label = (tB - tA)/tn, rot_label = Rodrigues^-1(normalize_cols(R_B R_A^T))/rn; double (n,3). */
int se3tn_so3_log(se3tn_ctx* ctx, const double* poses_a, const double* poses_b,
                  double trans_normalizer, double rot_normalizer,
                  double* trans_label, double* rot_label, int n, void* stream);

/* Tracker.on_track for n independent tracks of one frame (reference predict.py:217-296 with the
 * renderer's output passed in and the GUI calls dropped): K0 -> conv stack -> K6 on one stream.
 * weight_ids_host: int32 (n) HOST array or NULL (all 0), weight_ids_dev: the same values on the device (or NULL).
 * In the tensor-core modes tracks of ALL object classes share the same 14 conv launches: every work unit
 * takes its weight tensor map / bias from per-set device tables.  Any id order is correct; keeping equal ids
 * contiguous (dist.shard_tracks does) avoids shared-memory weight reloads in the 64-channel layers.  In
 * SE3TN_PREC_FP32 each contiguous run of equal ids is one batched forward.
 * out_trans/out_rot float32 (n,3) device scratch the caller provides (also returned). */
int se3tn_track_batch(se3tn_ctx* ctx, const uint8_t* frame_rgb, const uint16_t* frame_depth, int H, int W,
                      const double* K, const double* poses_in, const double* object_width,
                      const uint8_t* rgbA, const uint16_t* depthA,
                      const int32_t* weight_ids_host, const int32_t* weight_ids_dev, int n,
                      double trans_normalizer, double rot_normalizer, int precision,
                      float* out_trans, float* out_rot, double* poses_out, void* stream);

/* a9, host side of the frame hand-over (reference predict.py:267-268 uploads whole tensors with .cuda()): copy only the
 * rows [y0, y1) x columns [x0, x1) of a HOST frame (rgb uint8 (H,W,3), depth uint16 (H,W); either may be NULL) into the
 * same rectangle of full-size DEVICE frame buffers.  K0 reads a frame only inside the tracks' crop windows, so a caller
 * that tracks a few objects uploads their bounding rectangle (a third of a 480x640 frame for one object) instead of 1.5 MB.
 * Two cudaMemcpy2DAsync on `stream`; pageable host memory makes them synchronous, as any such copy. */
int se3tn_upload_frame_window(se3tn_ctx* ctx, const uint8_t* rgb_host, const uint16_t* depth_host, int H, int W,
                              int y0, int y1, int x0, int x1, uint8_t* rgb_dev, uint16_t* depth_dev, void* stream);

/* The one exchange step of the sharded path: all-gather of the updated poses over an EXISTING NCCL
 * communicator, for hosts that drive libse3tn without torch.distributed (the Python layer uses
 * torch.distributed.all_gather_into_tensor, dist.py).  nccl_comm: the host's ncclComm_t; local_poses double
 * (n_local,16) device; all_poses double (world*n_local,16) device, rank-major; every rank passes the same n_local.
 * NCCL is resolved at run time (dlopen of libnccl.so.2, the copy already loaded in the process if there is one);
 * SE3TN_ERR_UNSUPPORTED if it cannot be found.  The reference has no multi-GPU code to replace. */
int se3tn_allgather_poses(se3tn_ctx* ctx, void* nccl_comm, const double* local_poses, double* all_poses, int n_local,
                          void* stream);

/* ---- pose-error metrics (SURVEY.md 8(f) "next" row 1; not on the per-frame path) ----------------------- */

/* Utils.add / Utils.adi (reference Utils.py:72-98) for n (pred, gt) pose pairs against one model point cloud:
 *   ADD   = mean_i |(R_p x_i + t_p) - (R_g x_i + t_g)|,   ADD-S = mean_i min_j |(R_g x_i + t_g) - (R_p x_j + t_p)|
 * model_pts double (m,3), pred / gt double (n,16), out_add / out_adi double (n), all device; either output may be NULL.
 * float64, exhaustive nearest neighbour (the reference uses scipy's cKDTree: same minimum). */
int se3tn_add_adi(se3tn_ctx* ctx, const double* model_pts, int m, const double* pred, const double* gt, int n,
                  double* out_add, double* out_adi, void* stream);

/* VOCap (reference eval_ycb.py:45-64): errs double (n) device, any order -> *out_ap on the HOST (0..1; the
 * reference multiplies by 100 when printing).  Synchronises the stream.  n == 0 or no error below 0.1 m -> 0
 * (the reference raises IndexError there). */
int se3tn_vocap(se3tn_ctx* ctx, const double* errs, int n, double* out_ap, void* stream);

/* ---- input A: the rendered previous view (SURVEY.md 8(f) "next" row 2) ------------------------------------- */

/* The CAD model the renderer draws: what VispyRenderer.__init__ uploads as vertex / index buffers (reference
 * vispy_renderer.py:108-129).  HOST arrays, copied: pos float32 (nv,3) metres in the object frame, nrm float32 (nv,3)
 * unit normals, col uint8 (nv,3), faces int32 (nf,3).  mesh_id >= 0; a later call with the same id replaces the model. */
int se3tn_set_mesh(se3tn_ctx* ctx, int mesh_id, const float* pos, const float* nrm, const uint8_t* col,
                   const int32_t* faces, int nv, int nf);

/* Tracker.render_window for n tracks (reference predict.py:193-215 -> vispy_renderer.py:135-178): the model at `poses`
 * rasterised into each track's 176x176 window (y-flipped orthographic crop of the pinhole projection, depth test LESS,
 * no culling, Lambert + ambient shading) -> rgbA uint8 (n,176,176,3) and depthA uint16 (n,176,176) mm, 0 = background,
 * both device -- exactly the arrays se3tn_preprocess / se3tn_track_batch take.  K: 4 doubles HOST (fx, fy, cx, cy);
 * poses double (n,16) device; object_width double (n) device; mesh_ids int32 (n) device or NULL (all 0). */
int se3tn_render(se3tn_ctx* ctx, const double* K, const double* poses, const double* object_width,
                 const int32_t* mesh_ids, int n, uint8_t* rgbA, uint16_t* depthA, void* stream);

/* The same for either of the reference's two producers of input A (predict.py:161-182 picks one from dataset_info['renderer']):
 *   SE3TN_RENDER_VISPY     what se3tn_render does (H, W ignored).
 *   SE3TN_RENDER_PYRENDER  offscreen_renderer.py:47-83 + predict.py:210-214: the model is drawn into the WHOLE H x W camera image
 *                          (pinhole K, near 0.1 m, far 2 m, ambient light only: unlit vertex colours), the metric depth becomes
 *                          uint16 mm, and crop_bbox (Utils.py:320-359: window from compute_bbox, zero outside the image, nearest-
 *                          neighbour resize to 176 x 176) cuts the track's window out of both.  Only the camera pixels the resize
 *                          picks are ever shaded; the full image is never materialised.  Per-fragment texture lookups of a
 *                          textured .obj are replaced by per-vertex colours. */
#define SE3TN_RENDER_VISPY 0
#define SE3TN_RENDER_PYRENDER 1
int se3tn_render_ex(se3tn_ctx* ctx, const double* K, const double* poses, const double* object_width,
                    const int32_t* mesh_ids, int n, int mode, int H, int W, uint8_t* rgbA, uint16_t* depthA, void* stream);

/* ---- live-sensor depth (SURVEY.md 8(f) "next" row 4) ------------------------------------------------------ */

/* fill_depth as the reference's ROS node applies it to every depth image before tracking (reference Utils.py:455-514,
 * predict_ros.py:38-41: extrapolate=False, bilateral): invert, 5x5 diamond dilate, 5x5 close, fill empties from a 7x7
 * dilation, 5x5 median, bilateral(5, 1.5, 2.0), invert back.  depth_mm uint16 (H,W) device -> out_mm uint16 (H,W) device
 * (= (fill_depth(depth/1e3) * 1000).astype(uint16)) and/or out_m float32 (H,W) metres; either may be NULL.  Bit-identical
 * to OpenCV up to the median; the bilateral's float32 accumulation order differs (|diff| ~ 5e-7 m). */
int se3tn_fill_depth(se3tn_ctx* ctx, const uint16_t* depth_mm, int H, int W, double max_depth,
                     uint16_t* out_mm, float* out_m, void* stream);

/* The same with the reference's two optional arguments (Utils.py:455: extrapolate=False, blur_type='bilateral'):
 * extrapolate != 0: every column's first valid value is extended to the top row and what is still empty takes a 31x31
 * dilation (Utils.py:486-497); blur_type SE3TN_BLUR_GAUSSIAN: cv2.GaussianBlur(5x5, sigma 0) on the valid pixels instead of
 * the bilateral filter (Utils.py:506-510). */
enum { SE3TN_BLUR_BILATERAL = 0, SE3TN_BLUR_GAUSSIAN = 1 };
int se3tn_fill_depth_ex(se3tn_ctx* ctx, const uint16_t* depth_mm, int H, int W, double max_depth, int extrapolate, int blur_type,
                        uint16_t* out_mm, float* out_m, void* stream);

/* The reference's own calling pattern as ONE call (Tracker.on_track, predict.py:217-296: numpy arrays in, numpy pose out):
 * every pointer is HOST memory.  The frame's crop-window rectangle, the poses, widths, input A and the ids are staged
 * through context-owned pinned memory into context-owned device buffers (stable addresses, so the step's CUDA graph is
 * replayed), se3tn_track_batch runs on them, and the call returns once poses_out (n x 16 doubles; out_trans / out_rot n x 3
 * floats, nullable) hold the result -- it synchronises `stream`.  frame_rgb uint8 (H,W,3), frame_depth uint16 (H,W) mm,
 * K = fx fy cx cy, poses double (n,16), object_width double (n) mm, rgbA uint8 (n,176,176,3), depthA uint16 (n,176,176),
 * weight_ids int32 (n) or NULL (all tracks use set 0).  Errors as se3tn_track_batch. */
int se3tn_track_host(se3tn_ctx* ctx, const uint8_t* frame_rgb, const uint16_t* frame_depth, int H, int W, const double* K,
                     const double* poses, const double* object_width, const uint8_t* rgbA, const uint16_t* depthA,
                     const int32_t* weight_ids, int n, double trans_normalizer, double rot_normalizer, int precision,
                     double* poses_out, float* out_trans, float* out_rot, void* stream);

/* ---- introspection (tests / profiling) -------------------------------------------------------- */

/* Device pointer + per-image float count of an internal NHWC activation buffer.
 * ids: 0 stemA 1 stemB 2 Y1A 3 Y1B 4 P1A 5 P1B 6 T1 7 T2 8 U 9 CAT 10 F1 11 T4 12 F2 13 H1 14 H2 15 H3 */
int se3tn_debug_buffer(se3tn_ctx* ctx, int id, float** ptr, size_t* floats_per_image);

/* Per-kernel device timing for bench.py's roofline line: when enabled, every launch of the hot path
 * is bracketed by CUDA events on the caller's stream.  se3tn_get_profile synchronises those events
 * and writes SE3TN_PROFILE_SLOTS durations (ms) of the LAST call: [0..13] the 14 conv launches in
 * schedule order, [14],[15] the two max-pools, [16] head, [17] preprocess/normalize, [18] pose update,
 * [19] input repack (se3tn_forward only), [20] render.  Slots that did not run read 0. */
#define SE3TN_PROFILE_SLOTS 21
int se3tn_set_profiling(se3tn_ctx* ctx, int enable);
int se3tn_get_profile(se3tn_ctx* ctx, float* ms);

/* Device-side timeline of the conv kernels (contexts created with SE3TN_TRACE=1 in the environment): synchronises the
 * device and copies 14 x 256 x 8 uint64 to the HOST: for conv launch l and CTA b, 8 %globaltimer (ns) stamps of the LAST
 * forward -- 0 entry, 1 setup done, 2 first weights in shared memory, 3 first activation unit, 4 last MMA committed,
 * 5 first accumulator ready, 6 last epilogue done, 7 exit (low 8 bits replaced by the SM id).  Profiling tool only. */
#define SE3TN_TRACE_WORDS (14 * 256 * 8)
int se3tn_get_trace(se3tn_ctx* ctx, unsigned long long* out);

/* Number of kernels the last forward / track_batch call on this context launched (for a replayed CUDA graph: the kernels
 * inside it).  se3tn_track_batch captures each distinct step (same pointers, sizes and precision) into a CUDA graph the
 * first time it sees it and replays it afterwards -- one graph launch per step; SE3TN_GRAPH=0 in the environment, an
 * enabled profiler or SE3TN_PREC_FP32 use plain stream launches.  se3tn_last_step_was_graph: 1 if the last track_batch
 * call was a graph launch. */
int se3tn_last_step_was_graph(se3tn_ctx* ctx);
int se3tn_last_launch_count(se3tn_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SE3TN_H */
