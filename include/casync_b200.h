/*
 * casync_b200 -- C ABI of the B200-native (sm_100a) CASync generator forward pass.
 *
 * The reference (ChrisFourteen/CALipSync) has no FFI / plugin layer: its boundary for this path is the
 * Python class `Model` (module/unet.py:273-345, identical copy image_infer_v1/models/unet.py:241-312)
 * called as `self.net(batch_tensor, hubert_tensor)` at
 * image_infer_v1/tools/frame_synthesizer/infer_api.py:259-260 and step2_train_unet.py:107.
 * `calipsync_b200.unet.Model` keeps that class contract and reaches the kernels through the entry points
 * below (ctypes; see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions: plain pointers and sizes only (no torch types); every device buffer is owned by the
 * caller; calls only ENQUEUE work on `stream` (a cudaStream_t passed as void*) and never synchronise;
 * return 0 on success or a negative CASYNC_E* code, with text in casync_last_error(); nothing aborts or
 * throws across the boundary (the reference's callers wrap the model call in try/except and fall back to
 * the original frames, infer_api.py:352-357, so errors must surface as ordinary return codes).
 */
#ifndef CASYNC_B200_H_
#define CASYNC_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define CASYNC_API __attribute__((visibility("default")))
#else
#define CASYNC_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define CASYNC_OK 0
#define CASYNC_EINVAL (-1)   /* bad argument (shape / null pointer / unknown name) */
#define CASYNC_ECUDA (-2)    /* a CUDA runtime call or launch failed                */
#define CASYNC_EDEVICE (-3)  /* current device is not compute capability 10.x, or is not the plan's device */
#define CASYNC_EUNSUP (-4)   /* requested mode not implemented (e.g. fp32 path)     */

/* casync_forward flags */
#define CASYNC_F_BF16 0u        /* bf16 operands, fp32 accumulate (the only implemented arithmetic)   */
#define CASYNC_F_FP32 1u        /* reserved: fp32 path -> CASYNC_EUNSUP                               */
#define CASYNC_F_OUT_U8_HWC 2u  /* `out` is uint8 [B,160,160,3] = floor(p*255) (infer_api.py:265-266)  */

typedef struct casync_plan casync_plan;

CASYNC_API const char *casync_version(void);
CASYNC_API const char *casync_last_error(void);

/* Packed-weight schema.  The host-side packer (calipsync_b200/packer.py) asks the library which entries
 * it expects, in which order and of what byte size, folds BatchNorm (module/unet.py:18,28,32 ...) into
 * them and writes them back to back into one blob at 256-byte aligned offsets. */
CASYNC_API int casync_weight_entry_count(void);
CASYNC_API int casync_weight_entry(int index, const char **name, size_t *bytes);

/* Replaces: Model.__init__ + load_state_dict + .to(device) (infer_api.py:41-43).  `host_blob` and
 * `dev_blob` hold the same `blob_bytes` bytes (host copy is read only during this call; the device copy
 * must outlive the plan).  `offsets[i]` is the byte offset of schema entry i. */
CASYNC_API int casync_plan_create(const void *host_blob, const void *dev_blob, size_t blob_bytes, const int64_t *offsets,
                       int n_entries, casync_plan **out);
CASYNC_API void casync_plan_destroy(casync_plan *plan);

/* Frames processed per internal pass; workspace is sized for min(batch, chunk). */
CASYNC_API int casync_chunk_frames(const casync_plan *plan);
CASYNC_API size_t casync_workspace_bytes(const casync_plan *plan, int batch);

/* Replaces: Model.forward(x, audio_feat) (module/unet.py:314-345).
 *   x_nchw   fp32 [batch,6,160,160]   (cat([face, masked face]) in [0,1]; not modified)
 *   audio    fp32 [batch,32,32,32]    (HuBERT window; not modified)
 *   out      fp32 [batch,3,160,160] in (0,1), or uint8 [batch,160,160,3] with CASYNC_F_OUT_U8_HWC
 *   workspace  >= casync_workspace_bytes(plan, batch) bytes, 256-byte aligned, device memory
 * Asynchronous: enqueues on `stream` and returns.  Batches of 24 frames or more are additionally spread over
 * plan-owned streams that are forked from and joined to `stream` (stream order is what the caller observes).  The second
 * call with the same (x, audio, out, workspace, batch, flags) is captured into a CUDA graph and later calls with that
 * key replay it; the buffers' CONTENTS may change freely between calls.  One call in flight per plan. */
CASYNC_API int casync_forward(const casync_plan *plan, const float *x_nchw, const float *audio, void *out, void *workspace,
                   int batch, unsigned flags, void *stream);

/* Replaces the caller-side input assembly of FrameSynthesizer.process_batch, on the device:
 *   crops_hwc  uint8 [batch,160,160,3]: `crop_img[4:164, 4:164]` of infer_api.py:238 (channel order as the caller's)
 *   feats      fp32 [n_feat_frames,2,1024]: the clip's HuBERT features, resident on the device
 *   frame_idx  int32 [batch] (device): audio frame index of every output frame
 * ->  x_nchw   fp32 [batch,6,160,160] = cat([crop/255, masked crop/255]) with the mask rectangle rows 5..149 x cols
 *              5..154 zeroed (infer_api.py:239-245);  audio fp32 [batch,32,32,32] = feats[idx-8 : idx+8] reshaped,
 *              zero rows outside the clip (infer_api.py:99-145).  Bit-identical to what the caller builds with numpy. */
CASYNC_API int casync_prepare_inputs(const uint8_t *crops_hwc, const float *feats, int n_feat_frames,
                                     const int32_t *frame_idx, float *x_nchw, float *audio, int batch, void *stream);

/* Paste-back blend of the caller on the device (infer_api.py:333-346):
 *   frames[b][ymin+y][xmin+x][c] = uint8( crop * m + frames * (1 - m) ),  m = face_mask/255 (* soft mask)
 * in float64 with separate multiply / add roundings and truncation -- exactly what numpy computes for
 * `(crop_img * mask) + (img[ymin:ymax, xmin:xmax] * (1.0 - mask))` assigned into the uint8 image.
 *   frames     uint8 [batch, H, W, 3], updated in place
 *   crops      uint8 [batch, ldc, ldc, 3]: the re-sized crop with the prediction pasted in (rows/cols >= the region's
 *              size are ignored).  The caller's cv2.resize stays where it is: its rounding is OpenCV-build dependent.
 *   face_mask  uint8 [batch, ldc, ldc]: the dilated face polygon (0 / 255 after fillPoly + dilate + bitwise_and)
 *   soft_mask  fp32 [batch, ldc, ldc] or NULL: the optional per-frame mask file, already re-sized (infer_api.py:336-343)
 *   rects      int32 [batch, 4] = (ymin, ymax, xmin, xmax) of every frame's region, on the device */
CASYNC_API int casync_blend_paste(uint8_t *frames, int H, int W, const uint8_t *crops, int ldc, const uint8_t *face_mask,
                                  const float *soft_mask, const int32_t *rects, int batch, void *stream);

/* Same as casync_forward, but records one CUDA event after every kernel launch, SYNCHRONISES the stream and
 * returns per-launch device time with the launch's algorithmic FLOPs and activation bytes (weights excluded).
 * Profiling aid for bench.py's roofline line; not for the timed throughput run. */
typedef struct casync_launch_record {
  char name[48];
  float ms;
  double flops;
  double bytes;
} casync_launch_record;
CASYNC_API int casync_forward_profiled(const casync_plan *plan, const float *x_nchw, const float *audio, void *out,
                                       void *workspace, int batch, unsigned flags, void *stream,
                                       casync_launch_record *recs, int max_recs, int *n_recs);

/* Stage activations left in `workspace` by the last casync_forward with batch <= casync_chunk_frames() and below the
 * two-lane split threshold (24 frames; CASYNC_SPLIT=0 lifts it):
 * bf16 row-major [rows, cols] with leading dimension `ld` (elements) at byte `offset` (NHWC: rows =
 * batch*H*W pixels).  Names: x1..x5, audio, tx, ox0..ox3, kx, fuse, up1..up4 (oracle STAGE_NAMES). */
CASYNC_API int casync_stage_view(const casync_plan *plan, int batch, const char *name, size_t *offset, int64_t *rows,
                      int64_t *cols, int64_t *ld);
CASYNC_API int64_t casync_launches_per_forward(const casync_plan *plan, int batch);
/* Forwards of this plan that were replayed from a cached CUDA graph (the second call with the same pointers, batch
 * and flags is captured; CASYNC_GRAPH=0 disables). */
CASYNC_API int64_t casync_graph_replays(const casync_plan *plan);

/* Per-stage entry points (unit tests / ncu).  Activations are NHWC bf16, `scratch` is device memory of
 * at least casync_stage_scratch_bytes(plan, batch).
 * casync_ir_block: one InvertedResidual (module/unet.py:8-40); `ir_index` as listed by casync_ir_info. */
CASYNC_API int casync_ir_count(void);
CASYNC_API int casync_ir_info(int ir_index, const char **name, int *cin, int *cout, int *h_in, int *stride, int *residual);
CASYNC_API size_t casync_stage_scratch_bytes(const casync_plan *plan, int batch);
CASYNC_API int casync_ir_block(const casync_plan *plan, int ir_index, const void *in_nhwc, void *out_nhwc, void *scratch,
                    int batch, void *stream);
/* AudioConvHubert (module/unet.py:147-194): fp32 [B,32,32,32] NCHW -> bf16 [B*100, 512] */
CASYNC_API int casync_audio_cnn(const casync_plan *plan, const float *audio, void *out_nhwc, void *scratch, int batch,
                     void *stream);
/* MLP fusion + bn_tx + 4 attention blocks + bn_kx (module/unet.py:323-336): x5, audio bf16 [B*100,512]
 * -> kx bf16 [B*100,1024] */
CASYNC_API int casync_fusion_attention(const casync_plan *plan, const void *x5, const void *audio, void *kx, void *scratch,
                            int batch, void *stream);
/* Up (module/unet.py:82-97), level 1..4: low [B,h,w,C] + skip [B,2h,2w,C] -> [B,2h,2w,Cout] */
CASYNC_API int casync_up_block(const casync_plan *plan, int level, const void *low, const void *skip, void *out, void *scratch,
                    int batch, void *stream);
/* first InvertedResidual of Up only (up<level>.conv.double_conv.0: bilinear x2 + concat + block), for unit tests / ncu:
   low [B,h,w,C] + skip [B,2h,2w,C] -> [B,2h,2w,Cout0] (module/unet.py:90-96 + :8-40) */
CASYNC_API int casync_up_first(const casync_plan *plan, int level, const void *low, const void *skip, void *out, void *scratch,
                    int batch, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CASYNC_B200_H_ */
