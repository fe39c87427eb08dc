/*
 * pgb200.h -- C-ABI of libpgb200.so: the B200 (sm_100a) implementation of pilotguru's per-frame
 * motion-annotation hot path (ORB extraction, Hamming projection matching, IMU+GPS calibration).
 *
 * Every entry point replaces a call boundary of the reference (file:line relative to waiwnf/pilotguru):
 *   pgb_orb_*       thirdparty/orb-slam2/include/ORBextractor.h:51-85  (ctor, operator(), getters, mvImagePyramid)
 *   pgb_match_*     thirdparty/orb-slam2/include/ORBmatcher.h:41-52    (SearchByProjection(Frame&,const Frame&,th,bMono),
 *                   DescriptorDistance) with Frame.cc:234-249,331-396 grid semantics
 *   pgb_imu_*       include/calibration/velocity.hpp:38-76 (AccelerometerCalibrator ctor, operator(), eval,
 *                   IntegrateTrajectory, ImuTimes) and src/fit_motion.cc:156-293 (sliding-window L-BFGS driver,
 *                   thirdparty/LBFGS/LBFGS.h:79-182)
 *
 * Conventions: opaque handles; caller-owned buffers; plain pointers and sizes; int status return
 * (0 = ok, <0 = error, text via pgb_last_error()); no exceptions cross the boundary.  A handle owns one CUDA
 * stream (or borrows the one given at creation) and is NOT thread-safe; distinct handles are independent.
 * There is no CPU fallback: every compute entry point fails with PGB_ERR_CUDA when no sm_100 device is usable.
 *
 * "device pointer" arguments may come from any allocator in the process (cudaMalloc, torch, ...).
 */
#ifndef PGB200_H_
#define PGB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGB_OK 0
#define PGB_ERR_INVALID (-1)  /* bad argument (mirrors the reference's CHECK/assert failures) */
#define PGB_ERR_CUDA (-2)     /* CUDA runtime/driver failure, or no usable device */
#define PGB_ERR_CAPACITY (-3) /* an internal or caller buffer was too small; nothing was silently truncated */
#define PGB_ERR_NUMERIC (-4)  /* L-BFGS line-search step left [min_step, max_step] (LineSearch.h:103-107) */

/* Same field order and size (28 B) as cv::KeyPoint, the element type of ORBextractor's output vector. */
typedef struct pgb_keypoint {
  float x, y;     /* pt, level-0 pixel coordinates (level coords * scale[octave], ORBextractor.cc:1094-1100) */
  float size;     /* (int)(31 * scale[octave]) */
  float angle;    /* degrees, [0,360) */
  float response; /* FAST score */
  int32_t octave;
  int32_t class_id; /* always -1 */
} pgb_keypoint;

const char* pgb_last_error(void);
/* Library/ABI version: major*10000 + minor*100 + patch. */
int pgb_version(void);
/* Number of CUDA kernels launched by this library in this process so far (bench.py's gpu_launches). */
uint64_t pgb_launch_count(void);

/* ------------------------------------------------------------------ ORB extractor -------------------------- */
typedef struct pgb_orb pgb_orb;

/* ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST) (ORBextractor.cc:410-470)
 * plus the capacity the device buffers are sized for.  stream: a cudaStream_t cast to void*, or NULL to let the
 * handle create its own. */
pgb_orb* pgb_orb_create(int device, int nfeatures, float scale_factor, int nlevels, int ini_th_fast,
                        int min_th_fast, int max_width, int max_height, int max_batch, void* stream);
void pgb_orb_destroy(pgb_orb*);

/* Getters (ORBextractor.h:63-85). Arrays have nlevels entries. */
int pgb_orb_levels(const pgb_orb*);
float pgb_orb_scale_factor(const pgb_orb*);
int pgb_orb_scale_factors(const pgb_orb*, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2);
int pgb_orb_features_per_level(const pgb_orb*, int32_t* n_per_level);
/* Maximum number of keypoints one frame can produce (sum over levels of quota+2): the `cap` callers need. */
int pgb_orb_max_keypoints(const pgb_orb*);
int pgb_orb_level_size(const pgb_orb*, int width, int height, int level, int* w, int* h);

/* ORBextractor::operator() over a batch of n_frames gray images of identical size (ORBextractor.cc:1042-1104).
 * where = 0: host input, host outputs (the reference's own calling convention; synchronous).
 * where = PGB_IN_DEVICE | PGB_OUT_DEVICE: everything device-resident, asynchronous on the handle's stream.
 * where = PGB_OUT_DEVICE: host (ideally pinned) frames are copied in with cudaMemcpy2DAsync, results stay on the
 * device for the matcher; asynchronous.
 * frame i starts at gray + i*frame_stride, rows are `pitch` bytes apart.
 * LIFETIME: device frames that are 16-byte aligned (pointer, pitch, frame_stride) are read IN PLACE as pyramid level 0:
 * the buffer must stay valid and unmodified until the asynchronous work has completed AND for as long as stage /
 * level getters (pgb_orb_run_stage, pgb_orb_get_level, pgb_orb_get_blurred_level, pgb_orb_get_score_map) are used on
 * that batch, i.e. until the next pgb_orb_extract call on the handle.
 * Outputs: kps[n_frames][cap], desc[n_frames][cap][32], counts[n_frames]; level-major then octree list order.
 * width==0 || height==0 => counts zeroed, PGB_OK (the reference returns silently on an empty image, :1045). */
#define PGB_IN_DEVICE 1
#define PGB_OUT_DEVICE 2
int pgb_orb_extract(pgb_orb*, const uint8_t* gray, int where, int n_frames, int width, int height, size_t pitch,
                    size_t frame_stride, pgb_keypoint* kps, uint8_t* desc, int32_t* counts, int cap);

/* mvImagePyramid[level] of frame `frame` of the last extract call, copied to a tight host buffer (w*h bytes). */
int pgb_orb_get_level(pgb_orb*, int frame, int level, uint8_t* out, int* w, int* h);
/* Stage outputs of the last extract call, for parity tests (host buffers):
 * FAST score map of a level (score = max arc threshold; 0 where not a corner at minThFAST or outside the tested
 * region); the per-level candidate list handed to the octree (x,y relative to the 16-px border, response). */
int pgb_orb_get_score_map(pgb_orb*, int frame, int level, uint8_t* out, int* w, int* h);
int pgb_orb_get_candidates(pgb_orb*, int frame, int level, int32_t* xyr /*[cap][3]*/, int cap, int32_t* n);
/* 7x7 sigma=2 fixed-point Gaussian blur of a whole level (ORBextractor.cc:1084-1085); tight host buffer. */
int pgb_orb_get_blurred_level(pgb_orb*, int frame, int level, uint8_t* out, int* w, int* h);
/* Re-run one stage on the frames resident from the last extract call (bench.py's per-stage timing, stage parity
 * tests). which: 0 = pyramid, 1 = FAST-9 score + per-cell NMS + threshold decision (the fused kernel of the hot path),
 * 2 = the unfused pair (score map kernel, then the cell kernel) recomputing the same candidates -- kept for A/B timing
 * and for pgb_orb_get_score_map --, 3 = octree, 4 = orientation+descriptor.  Asynchronous on the handle's stream.
 * With PGB_IN_DEVICE input that was read in place (16-byte aligned frames), the caller's frame buffer must still hold
 * the frames: level 0 is not copied (see pgb_orb_extract). */
int pgb_orb_run_stage(pgb_orb*, int which);
void* pgb_orb_stream(pgb_orb*);
/* Synchronise the handle's stream and report (and clear) device-side capacity flags raised by asynchronous
 * (is_device=1) extract calls. */
int pgb_orb_check(pgb_orb*);

/* Frame feed: cv::flip (src/io/image_sequence_reader.cc:163-175) + cvtColor to gray (Tracking::GrabImageMonocular,
 * thirdparty/orb-slam2/src/Tracking.cc:243-258) of n_frames interleaved 8-bit frames in one device pass.
 * channels 1/3/4; rgb_order = Camera.RGB (1: RGB(A), 0: BGR(A)); formula 0 = OpenCV 2.4 fixed point (the version the
 * reference pins), 1 = OpenCV >= 3 (bit-exact against cv2 4.13).  src/dst may each be host or device (is_device flags);
 * with device buffers the call is asynchronous on `stream`.  The gray output can be handed to pgb_orb_extract
 * (PGB_IN_DEVICE). */
int pgb_frames_to_gray(int device, const uint8_t* src, int src_is_device, int n_frames, int width, int height, int channels,
                       int rgb_order, size_t src_pitch, size_t src_frame_stride, int vertical_flip, int horizontal_flip,
                       int formula, uint8_t* dst_gray, int dst_is_device, size_t dst_pitch, size_t dst_frame_stride,
                       void* stream);

/* The same with the container's `rotate` metadata applied first, as VideoImageSequenceSource::fetchNext does
 * (src/io/image_sequence_reader.cc:113-118 reads it, :186-207 applies it: 90 -> cv::flip(raw.t(), out, 0),
 * 180 -> cv::flip(raw, out, -1), 270 -> cv::flip(raw.t(), out, 1); anything else is fatal there and PGB_ERR_INVALID here).
 * src_width x src_height is the DECODED frame; for 90 / 270 the gray output is src_height wide and src_width high
 * (dst_pitch / dst_frame_stride refer to that).  The flips act on the rotated image, as in the reference's call order. */
int pgb_frames_to_gray_rotated(int device, const uint8_t* src, int src_is_device, int n_frames, int src_width, int src_height,
                               int channels, int rgb_order, size_t src_pitch, size_t src_frame_stride, int rotate_degrees,
                               int vertical_flip, int horizontal_flip, int formula, uint8_t* dst_gray, int dst_is_device,
                               size_t dst_pitch, size_t dst_frame_stride, void* stream);

/* Frame feed, decode part: VideoImageSequenceSource (src/io/image_sequence_reader.cc:74-208) for Motion-JPEG AVI files.
 * pgb_video_open demuxes the container on the host (RIFF 'AVI ' / OpenDML 'AVIX', first video stream as
 * VideoStreamIndexOrDie :63-71 picks it) and indexes the frames; no GPU is needed for open / info.  Any other container or
 * codec is refused (NULL, pgb_last_error names the FOURCC): this image has nvJPEG, not libav / NVDEC.
 * pgb_video_read_rgb decodes frames [first_frame, first_frame + n_frames) with nvJPEG (bound at run time) into interleaved
 * RGB24 DEVICE memory -- the raw_frame_image_ of the reference (:157-170) -- asynchronously on `stream`, and writes the
 * frames' timestamps in seconds (best-effort pts x time_base, :153-155; host array, may be NULL).  Feed the result to
 * pgb_frames_to_gray_rotated (channels 3, rgb_order 1, src_is_device 1) with pgb_video_info's rotate_degrees.
 * A handle is not thread-safe; independent handles on the same file are (frames are independently decodable, which is
 * what lets optical_trajectories --num_gpus shard a file). */
typedef struct pgb_video pgb_video;
pgb_video* pgb_video_open(int device, const char* path);
int pgb_video_info(pgb_video*, int* width, int* height, int64_t* n_frames, double* fps, int* rotate_degrees);
/* byte range of a frame's JPEG image inside the file (demuxer parity tests) */
int pgb_video_frame_span(pgb_video*, int64_t frame, uint64_t* offset, uint32_t* size);
int pgb_video_read_rgb(pgb_video*, int64_t first_frame, int n_frames, uint8_t* rgb_dev, size_t pitch, size_t frame_stride,
                       double* timestamps_sec, void* stream);
void pgb_video_close(pgb_video*);

/* Synthetic frame source for the long BASELINE configs (SURVEY.md 8d's generator rendered on the device: canvas crop at
 * the ping-pong origin of frame first_t + i, plus deterministic per-frame noise).  Stands where a hardware video decoder
 * would: n_frames tight width x height gray frames appear in device memory (out_dev), ready for pgb_orb_extract
 * (PGB_IN_DEVICE).  canvas_dev: canvas_w x canvas_h gray bytes on the device.  Asynchronous on `stream`. */
int pgb_synth_frames(int device, const uint8_t* canvas_dev, int canvas_w, int canvas_h, int first_t, int n_frames,
                     int width, int height, uint8_t* out_dev, void* stream);

/* Device / pinned memory for host programs that keep frames and features on the device between calls and link only this
 * library (its CUDA runtime is private).  kind: 0 host->device, 1 device->host, 2 device->device. */
int pgb_device_count(void);
void* pgb_device_malloc(int device, size_t bytes); /* zero-filled; NULL on failure */
void pgb_device_free(int device, void* p);
void* pgb_host_malloc_pinned(size_t bytes);
void pgb_host_free_pinned(void* p);
int pgb_memcpy_async(int device, void* dst, const void* src, size_t bytes, int kind, void* stream);
int pgb_stream_synchronize(int device, void* stream);
/* Completion markers for callers that keep several batches in flight on one stream: record behind a batch's copies,
 * synchronize when the host needs that batch's results. */
void* pgb_event_create(int device);
int pgb_event_record(int device, void* event, void* stream);
int pgb_event_synchronize(int device, void* event);
void pgb_event_destroy(int device, void* event);

/* ------------------------------------------------------------------ matcher -------------------------------- */
/* ORBmatcher::DescriptorDistance (ORBmatcher.cc:1651-1667) for n pairs of 32-byte descriptors (device or host
 * pointers; host pointers are staged). Mostly a test hook for the popcount primitive. */
int pgb_descriptor_distance(const uint8_t* a, const uint8_t* b, int n, int32_t* dist, int is_device, void* stream);

typedef struct pgb_matcher pgb_matcher;
/* ORBmatcher(nnratio, checkOri) (ORBmatcher.cc:42); max_feats / max_batch size the device scratch. */
pgb_matcher* pgb_matcher_create(int device, float nnratio, int check_orientation, int max_feats, int max_batch,
                                void* stream);
void pgb_matcher_destroy(pgb_matcher*);

/* SearchByProjection(CurrentFrame, LastFrame, th, bMono=true) (ORBmatcher.cc:1332-1474) for n_pairs
 * independent frame pairs.  For pair p:
 *   current frame: cur_kps[p][cap], cur_desc[p][cap][32], cur_counts[p]
 *   queries (the last frame's map points): q_uv[p][cap][2] projected pixel, q_octave, q_angle, q_desc[p][cap][32],
 *   q_valid[p][cap] (0 = no map point / outlier / behind camera: skipped), q_counts[p]
 * image bounds mnMinX..mnMaxY, th, and the extractor scale factors define the search windows.
 * Output: match_of_cur[p][cap] = query index assigned to each current keypoint or -1 (CurrentFrame.mvpMapPoints),
 *         n_matches[p] (the return value).  All buffers device (is_device=1) or host (0). */
int pgb_match_by_projection(pgb_matcher*, int n_pairs, int cap, const pgb_keypoint* cur_kps, const uint8_t* cur_desc,
                            const int32_t* cur_counts, const float* q_uv, const int32_t* q_octave,
                            const float* q_angle, const uint8_t* q_desc, const uint8_t* q_valid,
                            const int32_t* q_counts, float min_x, float max_x, float min_y, float max_y, float th,
                            const float* scale_factors, int nlevels, int32_t* match_of_cur, int32_t* n_matches,
                            int is_device);

/* Convenience used by the synthetic benchmark (SURVEY.md section 8d "match stage definition"): frame t-1's
 * keypoints act as map points projected to (x+flow_x, y+flow_y); retries with 2*th when fewer than 20 matches
 * (Tracking.cc:876-883).  kps/desc/counts are per-frame arrays as produced by pgb_orb_extract; pairs are
 * (prev = first_prev+p, cur = first_prev+p+1) for p in [0,n_pairs). flow[p][2]. Device buffers only. */
int pgb_match_consecutive(pgb_matcher*, int n_pairs, int cap, const pgb_keypoint* kps, const uint8_t* desc,
                          const int32_t* counts, const float* flow, float max_x, float max_y, float th,
                          const float* scale_factors, int nlevels, int32_t* match_of_cur, int32_t* n_matches);

/* What optical_trajectories' flow-tracking loop takes from a matched pair (the stand-in for the pose step of
 * Tracking::TrackWithMotionModel, Tracking.cc:858-919): the median displacement cur - prev of the
 * matched keypoints, per axis (element n/2 of the sorted displacements), and tracked = n_matches >= 20 (the threshold of
 * Tracking.cc:884).  Pairs and arrays as in pgb_match_consecutive (its outputs are this call's inputs); flow_xy[pair][2],
 * tracked[pair]; device buffers, asynchronous on `stream`; cap <= 2048. */
int pgb_match_median_flow(int device, int n_pairs, int cap, const pgb_keypoint* kps, const int32_t* counts,
                          const int32_t* match_of_cur, const int32_t* n_matches, float* flow_xy, int32_t* tracked, void* stream);

/* ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (ORBmatcher.cc:407-522) for
 * n_pairs independent frame pairs: level-0 keypoints of F1 search a +-window_size window around prev_matched_xy in
 * F2 (level 0 only), TH_LOW = 50, the ratio test with the matcher's nnratio, stealing of already matched targets by
 * closer queries, and the rotation-histogram filter.  prev_matched_xy[pair][cap][2] is updated in place for matched
 * keypoints (:517-519); matches12[pair][cap] = index in F2 or -1; n_matches[pair] = the return value.
 * More than 96 candidates in one window -> PGB_ERR_CAPACITY. */
int pgb_match_for_initialization(pgb_matcher*, int n_pairs, int cap, const pgb_keypoint* kps1, const uint8_t* desc1,
                                 const int32_t* counts1, const pgb_keypoint* kps2, const uint8_t* desc2,
                                 const int32_t* counts2, float* prev_matched_xy, int window_size, float min_x, float max_x,
                                 float min_y, float max_y, int32_t* matches12, int32_t* n_matches, int is_device);

/* ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) (ORBmatcher.cc:46-131; Tracking::SearchLocalPoints)
 * for n_frames independent frames.  Per map point (arrays [frame][cap]): projection proj_xy (mTrackProjX/Y), predicted
 * level (mnTrackScaleLevel), viewing cosine (mTrackViewCos -> RadiusByViewingCos, :133-139), descriptor, in_view
 * (mbTrackInView && !isBad()), mp_observed (Observations() > 0).  has_map_point[frame][cap]: the feature already holds
 * a map point with observations (skipped, :85-87).  match_of_feature[frame][cap] = index of the map point this call
 * assigned to the feature, else -1; n_matches[frame] = the return value.  Mono only (mvuRight < 0). */
int pgb_match_map_points(pgb_matcher*, int n_frames, int cap, const pgb_keypoint* kps, const uint8_t* desc,
                         const int32_t* counts, const uint8_t* has_map_point, const float* proj_xy,
                         const int32_t* track_level, const float* view_cos, const uint8_t* mp_desc, const uint8_t* in_view,
                         const uint8_t* mp_observed, const int32_t* mp_counts, float min_x, float max_x, float min_y,
                         float max_y, float th, const float* scale_factors, int nlevels, int32_t* match_of_feature,
                         int32_t* n_matches, int is_device);

/* ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (ORBmatcher.cc:161-290; Tracking::TrackReferenceKeyFrame,
 * Tracking::Relocalization) for n_pairs independent (keyframe, frame) problems.  Per problem p: keyframe descriptors
 * kf_desc[p][cap][32], angles kf_angle[p][cap] (mvKeysUn[i].angle), kf_has_map_point[p][cap] (vpMapPointsKF[i] &&
 * !isBad()); frame descriptors f_desc, angles f_angle (mvKeys[i].angle), f_counts[p] = F.N.  The DBoW2::FeatureVector
 * of each side is a CSR: nodes kf_node_off[p] .. kf_node_off[p+1]-1 belong to problem p, node k has id kf_node_id[k]
 * (strictly increasing inside a problem = std::map order) and lists the features kf_feat_idx[kf_feat_start[k] ..
 * kf_feat_start[k+1]-1] in their vector order; same for f_*.  A frame feature appears in at most one node (what
 * DBoW2's transform produces); anything else -> PGB_ERR_INVALID.  match_of_feature[p][cap] = index of the keyframe
 * feature whose map point the frame feature received (vpMapPointMatches), else -1; n_matches[p] = the return value.
 * TH_LOW = 50, the matcher's nnratio and checkOrientation apply. */
int pgb_match_by_bow(pgb_matcher*, int n_pairs, int cap, const uint8_t* kf_desc, const float* kf_angle,
                     const uint8_t* kf_has_map_point, const int32_t* kf_node_off, const uint32_t* kf_node_id,
                     const int32_t* kf_feat_start, const uint32_t* kf_feat_idx, int kf_nodes_total, int kf_idx_total,
                     const uint8_t* f_desc, const float* f_angle, const int32_t* f_counts, const int32_t* f_node_off,
                     const uint32_t* f_node_id, const int32_t* f_feat_start, const uint32_t* f_feat_idx, int f_nodes_total,
                     int f_idx_total, int32_t* match_of_feature, int32_t* n_matches, int is_device);

/* MapPoint::ComputeDistinctiveDescriptors (thirdparty/orb-slam2/src/MapPoint.cc:259-324) for n_points map points:
 * the observing descriptors of point p are rows offsets[p] .. offsets[p+1]-1 of desc[][32]; best_idx[p] = row (relative
 * to offsets[p]) with the least median Hamming distance to the others, -1 for a point without observations.
 * More than 256 observations of one point -> PGB_ERR_CAPACITY. */
int pgb_distinctive_descriptors(const uint8_t* desc, const int32_t* offsets, int n_points, int32_t* best_idx, int is_device,
                                void* stream);

/* Optimizer::PoseOptimization(Frame*) (thirdparty/orb-slam2/src/Optimizer.cc:239-451; called from
 * Tracking::TrackWithMotionModel / TrackReferenceKeyFrame / TrackLocalMap / Relocalization) for n_frames independent
 * frames, monocular observations only (mvuRight < 0): 4 rounds of <= 10 g2o Levenberg-Marquardt iterations over the
 * 6-DoF pose with Huber(sqrt(5.991)) edges, chi2 > 5.991 classifies outliers after every round, the robust kernel is
 * dropped for the last round, every round restarts from the input pose.  Per frame f (arrays [frame][cap]): Tcw_in[f][16]
 * = pFrame->mTcw (4x4 row-major float), kp_xy = mvKeysUn[i].pt, kp_octave, mp_xyz = mvpMapPoints[i]->GetWorldPos(),
 * has_map_point[i] = mvpMapPoints[i] != NULL, counts[f] = N; inv_level_sigma2[nlevels] = mvInvLevelSigma2; fx, fy, cx, cy.
 * Outputs: Tcw_out[f][16] (what SetPose receives; the input pose when there are fewer than 3 correspondences),
 * outlier[f][cap] = mvbOutlier (0 where there is no map point), n_inliers[f] = the return value
 * (nInitialCorrespondences - nBad).  fp64 inside, like g2o.  An octave outside [0, nlevels) -> PGB_ERR_INVALID. */
int pgb_pose_optimization(int device, int n_frames, int cap, const float* Tcw_in, const float* kp_xy,
                          const int32_t* kp_octave, const float* mp_xyz, const uint8_t* has_map_point,
                          const int32_t* counts, const float* inv_level_sigma2, int nlevels, float fx, float fy, float cx,
                          float cy, float* Tcw_out, uint8_t* outlier, int32_t* n_inliers, int is_device, void* stream);

/* ------------------------------------------------------------------ multi-GPU feature exchange ------------- */
/* Frames shard over the GPUs of one box in contiguous blocks; extraction is independent per
 * frame, and SearchByProjection(frame t, frame t-1) (Tracking.cc:860-883) is the only dependency that crosses a block
 * boundary.  The reference has no multi-GPU code: these entry points are what its frame loop
 * (src/slam/track_image_sequence.cc:63-109 -> System::TrackMonocular) gains when sharded.
 * One communicator rank per GPU; NCCL is loaded at run time (dlopen libnccl.so.2). */
typedef struct pgb_comm pgb_comm;
#define PGB_COMM_ID_BYTES 128
/* ncclGetUniqueId: rank 0 creates the id, the caller ships the bytes to the other ranks (torch.distributed, MPI, a file). */
int pgb_comm_unique_id(uint8_t id[PGB_COMM_ID_BYTES]);
/* One process per GPU (torchrun / bench.py): rank `rank` of `n_ranks` on `device`.  Collective: every rank must call it. */
pgb_comm* pgb_comm_create(int device, int rank, int n_ranks, const uint8_t id[PGB_COMM_ID_BYTES]);
/* One process, one host thread per GPU (optical_trajectories --num_gpus): all ranks at once, out[i] on devices[i]. */
int pgb_comm_create_all(int n, const int* devices, pgb_comm** out);
void pgb_comm_destroy(pgb_comm*);
int pgb_comm_rank(const pgb_comm*);
int pgb_comm_size(const pgb_comm*);
int pgb_comm_nccl_version(void); /* 0 when NCCL cannot be loaded */
/* The path's single collective: ncclAllGather of bytes_per_rank bytes from every rank (device pointers; recv holds
 * size * bytes_per_rank, rank-major).  Callers send either the block-boundary frame record (below) -- all the matcher
 * needs -- or their whole block of per-frame records.  Asynchronous on `stream` (a cudaStream_t). */
int pgb_allgather_feats(pgb_comm*, const void* send, void* recv, size_t bytes_per_rank, void* stream);
/* A frame's features as ONE contiguous record (count | keypoints[cap] | descriptors[cap][32], 16-byte aligned parts):
 * the unit that crosses ranks.  pack: frame `frame` of the per-frame arrays pgb_orb_extract wrote -> record;
 * unpack: record -> frame `frame` of such arrays (e.g. slot 0 = "predecessor of my first frame").  Device pointers,
 * asynchronous on `stream`. */
size_t pgb_frame_record_bytes(int cap);
int pgb_frame_record_pack(const pgb_keypoint* kps, const uint8_t* desc, const int32_t* counts, int frame, int cap,
                          void* record, void* stream);
int pgb_frame_record_unpack(const void* record, pgb_keypoint* kps, uint8_t* desc, int32_t* counts, int frame, int cap,
                            void* stream);

/* ------------------------------------------------------------------ IMU + GPS calibration ------------------ */
typedef struct pgb_imu pgb_imu;

/* AccelerometerCalibrator's sensor series (velocity.cc:29-39): whole-recording gyro + accel, host arrays
 * xyz[n][3] fp64, t[n] int64 usec (strictly increasing, CHECKed as align_time_series.cc:22-26).
 * Merges the two series once (MergeTimeSeries, align_time_series.cc:29-113) and keeps them on the device. */
pgb_imu* pgb_imu_create(int device, const double* gyro_xyz, const int64_t* gyro_t, size_t n_gyro,
                        const double* acc_xyz, const int64_t* acc_t, size_t n_acc, void* stream);
void pgb_imu_destroy(pgb_imu*);
/* ImuTimes(): number of merged events, their effective timestamps and component indices (host out, may be NULL). */
int64_t pgb_imu_merged_count(const pgb_imu*);
int pgb_imu_merged_events(const pgb_imu*, int64_t* t_usec, int64_t* gyro_idx, int64_t* acc_idx);

/* Select the reference (GPS) window: the calibrator a fit_motion window constructs (fit_motion.cc:183-190).
 * Builds the interpolation intervals (align_time_series.cc:155-196) and runs the rotation sweep that reduces the
 * window to per-GPS-interval constants. */
int pgb_imu_set_window(pgb_imu*, const double* gps_v, const int64_t* gps_t, int n_gps);
int64_t pgb_imu_window_intervals(const pgb_imu*);
/* AccelerometerCalibrator::operator()/eval (velocity.cc:41-193): x = (g[3], h[3], v0[3]). */
int pgb_imu_eval(pgb_imu*, const double x[9], double* loss, double grad[9]);
/* LBFGSSolver::minimize on the current window from x (in/out) (LBFGS.h:79-182, fit_motion.cc:167-197). */
int pgb_imu_minimize(pgb_imu*, double x[9], double* fx, int* n_iter, int max_iterations, double epsilon);
/* IntegrateTrajectory (velocity.cc:199-256): one entry per merged event touched by the window, ascending index.
 * Outputs host arrays of capacity cap: merged index, |v|, orientation (w,x,y,z), velocity xyz, duration usec. */
int pgb_imu_integrate(pgb_imu*, const double x[9], int64_t cap, int64_t* merged_idx, double* speed, double* quat_wxyz,
                      double* vel_xyz, int64_t* duration_usec, int64_t* n_out);

/* ComputeAndSaveForwardVelocitiesFromImu's window loop (fit_motion.cc:156-273) for ALL sliding windows at once:
 * windows start at GPS index 0, shift_step, 2*shift_step, ... and hold batch_size samples.  Every window's
 * L-BFGS runs on the device.  first_window/n_windows select a contiguous shard (n_windows<0 = all) so ranks can
 * split the windows; the per-event sums of different shards add up (in window order).
 * Outputs (host): per merged event speed_sum[n_merged], speed_cnt[n_merged] (accumulated in window order);
 * per window of the shard x_out[w][9], fx_out[w], iters_out[w] (any may be NULL). */
int pgb_imu_fit_windows(pgb_imu*, const double* gps_v, const int64_t* gps_t, int n_gps, int batch_size,
                        int shift_step, int max_iterations, double epsilon, int first_window, int n_windows,
                        double* speed_sum, int32_t* speed_cnt, double* x_out, double* fx_out, int32_t* iters_out);
int pgb_imu_num_windows(int n_gps, int shift_step);
/* Device time (CUDA events on the handle's stream) of the three kernels of the last pgb_imu_fit_windows[_fwd] call -- the
 * rotation sweep over every IMU sub-interval (K9), the per-window L-BFGS, the per-sub-interval speeds (K10) -- and the
 * number of IMU sub-intervals the sweep covered (64 algorithmic bytes each).  For bench.py's roofline. */
int pgb_imu_last_kernel_ms(pgb_imu*, float* sweep_ms, float* solve_ms, float* speeds_ms, int64_t* n_intervals);
/* Same, plus the forward-axis evidence of fit_motion.cc:172-173,223-248: the Kahan sum (include/math/math.hpp:8-27) of
 * the device-frame velocities conj(orientation)*velocity over every trajectory point with |v| >= fwd_min_velocity of
 * every window whose largest rotation acos(min |q.w|) reaches fwd_min_rotation_rad.  fwd_sum_xyz[3] is the raw sum
 * (the caller projects out the vertical axis and normalises, fit_motion.cc:281-283); shards add. */
int pgb_imu_fit_windows_fwd(pgb_imu*, const double* gps_v, const int64_t* gps_t, int n_gps, int batch_size,
                            int shift_step, int max_iterations, double epsilon, int first_window, int n_windows,
                            double* speed_sum, int32_t* speed_cnt, double* x_out, double* fx_out, int32_t* iters_out,
                            double fwd_min_velocity, double fwd_min_rotation_rad, double* fwd_sum_xyz,
                            int32_t* fwd_windows_used);

/* GetPrincipalRotationAxes (src/calibration/rotation.cc:16-57): integrates the gyro over intervals of at least
 * integration_interval_usec (device), then cv::PCA of the quaternion vector parts: axes_out = 3x3 row-major
 * eigenvectors, rows by descending eigenvalue, signs as OpenCV's Jacobi solver leaves them.  Fails with
 * PGB_ERR_INVALID when fewer than 3 intervals result (CHECK_GE, rotation.cc:46). */
int pgb_principal_rotation_axes(int device, const double* gyro_xyz, const int64_t* gyro_t, size_t n,
                                int64_t integration_interval_usec, double axes_out[9], int64_t* n_intervals);
/* GetAngularVelocitiesAroundAxisDirect (rotation.cc:103-119): out[i] = gyro[i] . axis / |axis|; |axis| must be
 * within 1e-2 of 1 (CHECK_GT/CHECK_LT). */
int pgb_angular_velocities_around_axis(int device, const double* gyro_xyz, size_t n, const double axis[3], double* out);

/* SmoothTimeSeries (src/slam/smoothing.cc:56-98): host arrays in/out, device compute. */
int pgb_smooth_time_series(int device, const double* values, const double* times, int64_t n,
                           const double* target_times, int64_t n_target, double sigma, double* out);

/* annotate_frames' per-frame labels (src/annotate_frames.cc:59-72): TimeSeries<double>::TimeAveragedValue
 * (include/interpolation/time_series.hpp:129-189) of the series (values, times_usec)[n] over every frame interval
 * (frame_times_usec[i-1], frame_times_usec[i]], i = 1..n_frames-1.  out_values/out_valid have n_frames-1 entries;
 * out_valid[i-1] = 0 where the series does not cover the interval (is_valid = false; the value is NaN).  Host arrays,
 * device compute.  Frame timestamps must be strictly increasing (CHECK_GT(end, start)); an interval ending at or
 * after the last event fails like the reference's CHECK in LinearInterpolate (PGB_ERR_INVALID). */
int pgb_time_averaged_values(int device, const double* values, const int64_t* times_usec, int64_t n,
                             const int64_t* frame_times_usec, int64_t n_frames, double* out_values, uint8_t* out_valid);

#ifdef __cplusplus
}
#endif
#endif /* PGB200_H_ */
