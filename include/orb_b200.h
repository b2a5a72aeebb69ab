/*
 * orb_b200.h -- C ABI of liborb_b200.so: the B200 (sm_100a) implementation of the ORB-SLAM2
 * front-end hot path of serviceberry3/weiner_slamit_v2.
 *
 * This is the drop-in boundary.  The reference has no FFI layer for this path: the seam is
 * the two C++ classes ORB_SLAM2::ORBextractor (I/ORBextractor.h:45-111) and
 * ORB_SLAM2::ORBmatcher (I/ORBmatcher.h:37-102), I/ = oRB_SLAM2_Android/src/main/jni/
 * ORB_SLAM2/include/.  weiner_slamit_v2_b200/shim/ re-implements those two classes with
 * unchanged signatures on top of the functions below (INTEGRATION.md shows the build line).
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a
 * negative ORBB200_E* code (orbb200_last_error() gives the text, per thread); nothing
 * throws; a handle owns all device memory and its CUDA streams (calls are ordered on ONE of them,
 * orbb200_extractor_stream / orbb200_matcher_stream) and is NOT re-entrant (same
 * rule as an ORBextractor instance, S/Frame.cc:93-96 uses two instances from two threads);
 * different handles may be used from different threads / on different GPUs.
 * There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef ORB_B200_H
#define ORB_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORBB200_OK            0
#define ORBB200_EINVAL       -1   /* bad argument (null pointer, size out of range, ...) */
#define ORBB200_ECUDA        -2   /* a CUDA runtime call or kernel failed */
#define ORBB200_ECAPACITY    -3   /* an output buffer is smaller than the result */
#define ORBB200_EGEOMETRY    -4   /* image geometry the reference algorithm is undefined on */
#define ORBB200_ENODEVICE    -5   /* no CUDA device / wrong architecture */

const char *orbb200_last_error(void);
int orbb200_device_count(void);           /* number of visible CUDA devices (0 if none) */
const char *orbb200_version(void);

/* cv::KeyPoint, field for field (28 bytes). */
typedef struct {
    float x, y;       /* pt, in level-0 pixel units */
    float size;       /* 31 * scale[octave], truncated to an integer value */
    float angle;      /* degrees, [0,360] */
    float response;   /* FAST corner score */
    int32_t octave;
    int32_t class_id; /* always -1 */
} orbb200_keypoint;

/* ------------------------------------------------------------------------------------- */
/* ORBextractor                                                                          */
/* ------------------------------------------------------------------------------------- */
typedef struct orbb200_extractor orbb200_extractor;

/* Replaces ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
 * (S/ORBextractor.cc:415-482; 1 < scaleFactor <= 2).  width/height fix the frame geometry of this handle,
 * max_batch the largest batch one call may carry, device the CUDA ordinal.
 * blur_taps: 0 = OpenCV >= 3.x fixed-point Gaussian {18,34,48,56,48,34,18}/256,
 *            1 = OpenCV 2.4.9 {18,34,49,55,49,34,18}/256 (what the Android build linked). */
int orbb200_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST,
                             int width, int height, int max_batch, int device, int blur_taps,
                             orbb200_extractor **out);
void orbb200_extractor_destroy(orbb200_extractor *h);

/* Replaces the getters of I/ORBextractor.h:63-83 and the protected tables. Arrays have
 * nlevels entries (umax: 16). Any pointer may be NULL. */
int orbb200_extractor_tables(const orbb200_extractor *h, float *scale, float *inv_scale, float *sigma2,
                             float *inv_sigma2, int *features_per_level, int *umax16);
int orbb200_extractor_levels(const orbb200_extractor *h);
/* Upper bound on keypoints per frame (sum over levels of quota+3): size output buffers with it. */
int orbb200_extractor_max_keypoints(const orbb200_extractor *h);
int orbb200_extractor_level_size(const orbb200_extractor *h, int level, int *w, int *hgt);

/* Replaces ORBextractor::operator()(image, mask, keypoints, descriptors)
 * (S/ORBextractor.cc:1064-1136) for `batch` frames in HOST memory.  Frame f starts at
 * images + f*frame_stride, rows are `stride` bytes apart.  keypoints: batch x cap entries,
 * descriptors: batch x cap x 32 bytes, counts: batch ints; frame f's results start at index
 * f*cap, in the reference's order (level-major, quadtree list order inside a level).
 * cap must be >= orbb200_extractor_max_keypoints(). */
int orbb200_extract_host(orbb200_extractor *h, const uint8_t *images, int batch, size_t stride, size_t frame_stride,
                         orbb200_keypoint *keypoints, uint8_t *descriptors, int32_t *counts, int cap);

/* The same call split in two, for a caller that streams batches: _async returns once the uploads, kernels
 * and downloads are queued; _wait blocks until the results are in the output arrays and reports device-side
 * failures.  `images` and the three output arrays must stay valid until _wait returns and should be
 * page-locked (cudaHostAlloc / cudaHostRegister), otherwise the copies serialise with the host.  One call is
 * in flight per handle (a second _async first waits for the first).  To overlap the upload of batch k+1 with
 * the kernels of batch k, alternate between two or three handles (measured on B200: three handles reach the
 * device-resident rate; the blocking call, which can only overlap inside one batch, reaches ~60 % of it). */
int orbb200_extract_host_async(orbb200_extractor *h, const uint8_t *images, int batch, size_t stride,
                               size_t frame_stride, orbb200_keypoint *keypoints, uint8_t *descriptors,
                               int32_t *counts, int cap);
int orbb200_extract_host_wait(orbb200_extractor *h);

/* Same, for frames already in DEVICE memory; outputs stay on the device.  Asynchronous on
 * the handle's stream; orbb200_extractor_sync() waits.  d_keypoints/d_descriptors/d_counts
 * may be NULL to use handle-owned buffers (fetch them with orbb200_extractor_outputs). */
int orbb200_extract_device(orbb200_extractor *h, const uint8_t *d_images, int batch, size_t stride, size_t frame_stride,
                           orbb200_keypoint *d_keypoints, uint8_t *d_descriptors, int32_t *d_counts, int cap);
int orbb200_extractor_sync(orbb200_extractor *h);
int orbb200_extractor_outputs(orbb200_extractor *h, orbb200_keypoint **d_keypoints, uint8_t **d_descriptors,
                              int32_t **d_counts, int *cap);
/* CUDA stream the handle launches on (a cudaStream_t), for event timing by the caller. */
void *orbb200_extractor_stream(orbb200_extractor *h);
/* Number of kernel launches the last extract call issued. */
int orbb200_extractor_last_launches(const orbb200_extractor *h);
/* CUDA device ordinal the handle lives on (every entry point makes it current for the calling thread). */
int orbb200_extractor_device(const orbb200_extractor *h);
/* TEST HOOK for the device-side error path: clamps the per-level FAST candidate capacity to `candidates_per_level`
 * (the real capacity is a strict bound, so the overflow bit cannot be provoked otherwise; <= 0 restores it).  A later
 * extract call that finds more candidates on a level drops the excess, raises the device status bit and the
 * synchronising call (orbb200_extract_host, _host_wait, orbb200_extractor_sync) returns ORBB200_ECUDA with
 * orbb200_last_error() naming "candidate overflow". */
int orbb200_extractor_debug_set_capacity(orbb200_extractor *h, int candidates_per_level);

/* Per-stage device time of the most recent call, from CUDA events recorded on the handle's stream
 * between the kernels: ms5 = {pyramid (all resize launches), FAST, quadtree, blur, orientation+descriptor}.
 * In calls of up to 8 frames the blur runs on the handle's internal side stream beside the quadtree (joined
 * before the descriptors; everything stays ordered on the handle's stream): there "blur" is the part of the
 * blur that outlasts the quadtree.
 * Off by default; the bench harness turns it on to compute per-kernel roofline fractions. */
int orbb200_extractor_set_profiling(orbb200_extractor *h, int on);
int orbb200_extractor_stage_ms(orbb200_extractor *h, float *ms5);

/* Stage read-back of the most recent extract call (parity tests; also backs the public
 * member mvImagePyramid of I/ORBextractor.h:85).  All copy into HOST buffers. */
int orbb200_extractor_get_level(orbb200_extractor *h, int frame, int level, int blurred, uint8_t *dst, size_t dst_stride);
/* FAST candidates of one level before the quadtree (S/ORBextractor.cc:805-849): x,y relative
 * to the 16-px border, score; returned sorted into the reference's cell-row-major order. */
int orbb200_extractor_get_candidates(orbb200_extractor *h, int frame, int level, int32_t *x, int32_t *y, int32_t *score,
                                     int cap, int *n);
/* keypoints of one level after DistributeOctTree (level coordinates, border included). */
int orbb200_extractor_get_level_keypoints(orbb200_extractor *h, int frame, int level, int32_t *x, int32_t *y,
                                          int32_t *score, int cap, int *n);

/* ------------------------------------------------------------------------------------- */
/* ORBmatcher                                                                            */
/* ------------------------------------------------------------------------------------- */
typedef struct orbb200_matcher orbb200_matcher;

/* Scratch owner for the matching entry points (the reference's ORBmatcher is stateless,
 * I/ORBmatcher.h:37-102; the handle only holds device buffers + a stream).
 * max_items = largest number of frame pairs / frames per call, max_points = largest number
 * of keypoints (or map points) per side. */
int orbb200_matcher_create(int max_items, int max_points, int device, orbb200_matcher **out);
void orbb200_matcher_destroy(orbb200_matcher *m);
void *orbb200_matcher_stream(orbb200_matcher *m);
int orbb200_matcher_sync(orbb200_matcher *m);
int orbb200_matcher_last_launches(const orbb200_matcher *m);
int orbb200_matcher_device(const orbb200_matcher *m);

/* Replaces ORBmatcher::DescriptorDistance (S/ORBmatcher.cc:1651-1667) for n independent
 * descriptor pairs a[i], b[i] (32 bytes each, HOST memory); dist[i] in 0..256. */
int orbb200_descriptor_distance(orbb200_matcher *m, const uint8_t *a, const uint8_t *b, int n, int32_t *dist);

/* One side of a frame for the search functions: SoA view of Frame::mvKeysUn / mDescriptors
 * (I/Frame.h).  Arrays are `items` x `stride` entries; n[i] keypoints are valid in item i. */
typedef struct {
    const int32_t *n;        /* items */
    const float *x, *y;      /* items x stride */
    const int32_t *octave;   /* items x stride */
    const float *angle;      /* items x stride (may be NULL where unused) */
    const uint8_t *desc;     /* items x stride x 32 */
    int stride;
} orbb200_frame_view;

/* Replaces ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)
 * (S/ORBmatcher.cc:409-524) for `items` independent frame pairs.  bounds = {mnMinX, mnMinY, mnMaxX,
 * mnMaxY} of the Frame class (S/Frame.cc:561-589; {0,0,cols,rows} without lens distortion).
 * prev_matched: items x f1.stride x 2 floats, in/out.  matches12: items x f1.stride ints.
 * nmatches: items ints.  All pointers are HOST pointers unless `on_device` is non-zero, in
 * which case every pointer (including those inside the views) is a device pointer and the
 * call is asynchronous on the matcher's stream. */
int orbb200_search_for_initialization(orbb200_matcher *m, int items, const orbb200_frame_view *f1,
                                      const orbb200_frame_view *f2, const float bounds[4], float nnratio,
                                      int check_orientation, int window_size, float *prev_matched,
                                      int32_t *matches12, int32_t *nmatches, int on_device);

/* Map points flattened from vector<MapPoint*> (the "variables used by the tracking",
 * I/MapPoint.h:96-104, plus GetDescriptor()/isBad()/Observations()). items x stride. */
typedef struct {
    const int32_t *n;
    const uint8_t *in_view, *bad;          /* mbTrackInView, isBad() */
    const float *proj_x, *proj_y, *proj_xr; /* mTrackProjX/Y/XR */
    const int32_t *level;                   /* mnTrackScaleLevel */
    const float *view_cos;                  /* mTrackViewCos */
    const uint8_t *desc;                    /* GetDescriptor(), x32 */
    const int32_t *obs;                     /* Observations() */
    int stride;
} orbb200_mappoint_view;

/* Replaces ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th)
 * (S/ORBmatcher.cc:47-131) for `items` independent frames.  u_right: items x f.stride
 * (Frame::mvuRight; NULL = monocular, all -1).  kp_mp: items x f.stride, in/out: index of
 * the map point each keypoint holds (Frame::mvpMapPoints): -1 none, -2 a map point that is
 * not in `mp` whose Observations() is in kp_mp_obs (may be NULL when no -2 is used).
 * scale_factors: nlevels floats (Frame::mvScaleFactors). nmatches: items ints. */
int orbb200_search_by_projection(orbb200_matcher *m, int items, const orbb200_frame_view *f,
                                 const float *u_right, const orbb200_mappoint_view *mp, int32_t *kp_mp,
                                 const int32_t *kp_mp_obs, const float *scale_factors, int nlevels,
                                 const float bounds[4], float nnratio, float th, int32_t *nmatches, int on_device);

/* The last frame as SearchByProjection(CurrentFrame, LastFrame, ...) reads it: items x stride.  has_mp =
 * LastFrame.mvpMapPoints[i] != NULL, outlier = mvbOutlier[i], world_pos = GetWorldPos() (3 floats), mp_desc =
 * GetDescriptor(), mp_obs = Observations(), octave = mvKeys[i].octave, angle = mvKeysUn[i].angle. */
typedef struct {
    const int32_t *n;
    const uint8_t *has_mp, *outlier;
    const float *world_pos;
    const uint8_t *mp_desc;
    const int32_t *mp_obs;
    const int32_t *octave;
    const float *angle;
    int stride;
} orbb200_lastframe_view;

/* Replaces ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono)
 * (S/ORBmatcher.cc:1332-1474; scope row N2) for `items` independent frame pairs.  Rcw: items x 9 (row major),
 * tcw: items x 3 = CurrentFrame.mTcw; K = {fx, fy, cx, cy}; mbf = CurrentFrame.mbf.  mode: 0 = neither bForward
 * nor bBackward (always the case when bMono), 1 = bForward, 2 = bBackward (:1352-1353; the caller evaluates
 * tlc.z against mb as the reference does).  kp_mp: items x cur.stride in/out = CurrentFrame.mvpMapPoints as an
 * index into the LAST frame's arrays (-1 none, -2 foreign with kp_mp_obs observations). */
int orbb200_search_by_projection_last_frame(orbb200_matcher *m, int items, const orbb200_frame_view *cur,
                                            const float *u_right, const orbb200_lastframe_view *last, const float *Rcw,
                                            const float *tcw, const float K[4], float mbf, int32_t *kp_mp,
                                            const int32_t *kp_mp_obs, const float *scale_factors, int nlevels,
                                            const float bounds[4], float th, int mode, int check_orientation,
                                            int32_t *nmatches, int on_device);

/* The map-point slots of a key frame (pKF->GetMapPointMatches()), items x stride each: valid[i] = the slot holds a
 * map point that is not bad and not in sAlreadyFound; max_distance / min_distance = the raw mfMaxDistance /
 * mfMinDistance (the invariance getters' 1.2f / 0.8f factors, S/MapPoint.cc:379-389, are applied on the device);
 * angle = pKF->mvKeysUn[i].angle (may be NULL when check_orientation is 0). */
typedef struct orbb200_keyframe_view {
    const int32_t *n;
    const uint8_t *valid;
    const float *world_pos;       /* x3 */
    const uint8_t *mp_desc;       /* x32 */
    const float *max_distance, *min_distance;
    const float *angle;
    int stride;
} orbb200_keyframe_view;

/* Replaces ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound,
 * const float th, const int ORBdist) (S/ORBmatcher.cc:1476-1603, the relocalisation search; scope row N2) for
 * `items` independent (frame, key frame) pairs.  Rcw / tcw as above; Ow: items x 3 = -Rcw^T tcw (the caller's
 * cv::Mat expression, :1482); log_scale_factor = CurrentFrame.mfLogScaleFactor.  kp_mp: items x cur.stride in/out =
 * CurrentFrame.mvpMapPoints as an index into the key frame's slots (-1 none; any other value marks a keypoint that
 * already holds a map point and is skipped).  The predicted level (MapPoint::PredictScale, glibc-exact logf) is
 * clamped to [0, nlevels-1]; the reference leaves it unclamped and indexes mvScaleFactors out of range for
 * distances in [0.8 mfMin, mfMin). */
int orbb200_search_by_projection_keyframe(orbb200_matcher *m, int items, const orbb200_frame_view *cur,
                                          const orbb200_keyframe_view *kf, const float *Rcw, const float *tcw,
                                          const float *Ow, const float K[4], int32_t *kp_mp, const float *scale_factors,
                                          int nlevels, float log_scale_factor, const float bounds[4], float th,
                                          int orb_dist, int check_orientation, int32_t *nmatches, int on_device);

/* One side of SearchByBoW for `items` frames: descriptors and keypoint angles as in orbb200_frame_view, plus the
 * DBoW2::FeatureVector (std::map<NodeId, vector<unsigned>>) flattened: node_id ascending (the map's order),
 * node_start = node_stride + 1 offsets per item into feat, feat = the feature indices node after node (every
 * feature index appears at most once).  valid: key-frame side only, 1 = the slot holds a map point that is not
 * bad (NULL = every slot). */
typedef struct orbb200_bow_view {
    const int32_t *n;
    const uint8_t *desc;          /* items x stride x 32 */
    const float *angle;           /* items x stride: pKF->mvKeysUn[].angle / F.mvKeys[].angle */
    const uint8_t *valid;
    const int32_t *n_nodes;       /* items */
    const uint32_t *node_id;      /* items x node_stride */
    const int32_t *node_start;    /* items x (node_stride + 1) */
    const uint32_t *feat;         /* items x stride */
    int stride, node_stride;
} orbb200_bow_view;

/* Replaces ORBmatcher::SearchByBoW(KeyFrame *pKF, Frame &F, vector<MapPoint*> &vpMapPointMatches)
 * (S/ORBmatcher.cc:161-292; scope row N3) for `items` independent (key frame, frame) pairs.
 * matches: items x f->stride out = the key-frame slot whose map point the frame keypoint received, -1 otherwise
 * (vpMapPointMatches as indices); nmatches: items. */
int orbb200_search_by_bow(orbb200_matcher *m, int items, const orbb200_bow_view *kf, const orbb200_bow_view *f,
                          float nnratio, int check_orientation, int32_t *matches, int32_t *nmatches, int on_device);

/* Replaces ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint*> &vpMatches12)
 * (S/ORBmatcher.cc:526-659, loop closing; scope row N3): both views carry `valid`.  matches12: items x kf1->stride
 * out = the slot of key frame 2 whose map point slot idx1 received, -1 otherwise. */
int orbb200_search_by_bow_keyframes(orbb200_matcher *m, int items, const orbb200_bow_view *kf1, const orbb200_bow_view *kf2,
                                    float nnratio, int check_orientation, int32_t *matches12, int32_t *nmatches,
                                    int on_device);

/* Keypoint geometry of one key frame for SearchForTriangulation, items x (the bow view's stride):
 * x, y = mvKeysUn[].pt; octave = mvKeysUn[].octave (read for key frame 2 only, may be NULL for key frame 1);
 * u_right = mvuRight (NULL = monocular: all negative); has_mp[i] = GetMapPoint(i) != NULL. */
typedef struct orbb200_tri_view {
    const float *x, *y;
    const int32_t *octave;
    const float *u_right;
    const uint8_t *has_mp;
} orbb200_tri_view;

/* Replaces ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo)
 * (S/ORBmatcher.cc:661-827; scope row N3) for `items` key-frame pairs (the bow views' `valid` is not read).
 * F12: items x 9 (row major); epipole: items x 2 = {ex, ey} of :668-675 (the caller's cv::Mat expression);
 * scale_factors2 / level_sigma2_2: pKF2->mvScaleFactors / mvLevelSigma2 (nlevels entries).
 * matches12: items x kf1->stride out = index in key frame 2 or -1 (vMatches12; the caller packs vMatchedPairs). */
int orbb200_search_for_triangulation(orbb200_matcher *m, int items, const orbb200_bow_view *kf1, const orbb200_tri_view *g1,
                                     const orbb200_bow_view *kf2, const orbb200_tri_view *g2, const float *F12,
                                     const float *epipole, const float *scale_factors2, const float *level_sigma2_2,
                                     int nlevels, int only_stereo, int check_orientation, int32_t *matches12,
                                     int32_t *nmatches, int on_device);

/* Replaces the selection in MapPoint::ComputeDistinctiveDescriptors (S/MapPoint.cc:248-313; scope row N4) for
 * `items` map points at once: map point p's observed descriptors (those of its non-bad key frames, in the
 * std::map's order) are rows offsets[p] .. offsets[p+1]-1 of `descriptors` (total x 32 bytes, offsets[items] ==
 * total).  best[p] = index, relative to offsets[p], of the descriptor with the least median Hamming distance to
 * all of them (first minimum; -1 for a map point without descriptors); best_median (may be NULL) = that median. */
int orbb200_distinctive_descriptors(orbb200_matcher *m, int items, const int32_t *offsets, const uint8_t *descriptors,
                                    int total, int32_t *best, int32_t *best_median, int on_device);

/* Candidate map points of a Fuse call, items x stride: valid[i] = pMP && !pMP->isBad() && !pMP->IsInKeyFrame(pKF);
 * normal = GetNormal(); max_distance / min_distance = the raw mfMaxDistance / mfMinDistance. */
typedef struct orbb200_fusepoints_view {
    const int32_t *n;
    const uint8_t *valid;
    const float *world_pos;       /* x3 */
    const float *normal;          /* x3 */
    const uint8_t *mp_desc;       /* x32 */
    const float *max_distance, *min_distance;
    int stride;
} orbb200_fusepoints_view;

/* The search of ORBmatcher::Fuse(KeyFrame *pKF, const vector<MapPoint*> &vpMapPoints, const float th)
 * (S/ORBmatcher.cc:829-948; scope row N3) for `items` (key frame, candidate list) pairs: projection, frustum,
 * distance and viewing-angle gates, MapPoint::PredictScale (clamped), KeyFrame::GetFeaturesInArea, level and
 * chi-square gates (5.99 mono / 7.8 stereo), most similar descriptor.  best_idx: items x pts->stride out = the
 * key-frame keypoint to fuse with (distance <= TH_LOW) or -1; best_dist (may be NULL) = the smallest distance seen
 * (256: none).  The replace-or-add surgery of :950-971 mutates the map and stays with the caller (the shim does it
 * in list order, re-checking isBad / IsInKeyFrame as the reference does).  kf: the key frame's mvKeysUn / octaves /
 * descriptors; u_right = mvuRight (NULL = monocular); bounds = the float image bounds of the Frame the key frame
 * was made from (the key frame's own int copies are derived from them); Rcw/tcw/Ow = the key frame's pose.
 * mode 0 = the overload above.
 * mode 1 = the search of Fuse(KeyFrame *pKF, cv::Mat Scw, const vector<MapPoint*> &vpPoints, float th,
 *          vector<MapPoint*> &vpReplacePoint) (:979-1104): Rcw / tcw / Ow = the decomposed Scw (:987-991, evaluated
 *          by the caller), no reprojection-error gates, valid[i] = !isBad && not in pKF->GetMapPoints().
 * mode 2 = one leg of SearchBySim3 (:1106-1330): the camera point Rcw*X + tcw goes through a second similarity
 *          R2 (items x 9 = sR21 or sR12), t2 (items x 3); dist3D is the norm of the result; no viewing-angle gate
 *          (normal, Ow may be NULL); acceptance threshold TH_HIGH.  SearchBySim3 = leg 1 (key frame 1's map points in
 *          key frame 2), leg 2 the other way round, then the agreement test of :1316-1328 (the shim, or
 *          ORBmatcher.search_by_sim3_batch in Python). */
int orbb200_fuse_search(orbb200_matcher *m, int items, const orbb200_frame_view *kf, const float *u_right,
                        const orbb200_fusepoints_view *pts, const float *Rcw, const float *tcw, const float *Ow,
                        const float K[4], float bf, const float *scale_factors, const float *inv_level_sigma2, int nlevels,
                        float log_scale_factor, const float bounds[4], float th, int mode, const float *R2, const float *t2,
                        int32_t *best_idx, int32_t *best_dist, int on_device);

/* Replaces ORBmatcher::SearchByProjection(KeyFrame *pKF, cv::Mat Scw, const vector<MapPoint*> &vpPoints,
 * vector<MapPoint*> &vpMatched, int th) (S/ORBmatcher.cc:294-407, loop closing; scope row N3) for `items` (key frame,
 * candidate list) pairs.  Views and pose as for orbb200_fuse_search mode 1 (Scw decomposed by the caller);
 * valid[i] = !isBad && not already in vpMatched.  matched: items x kf->stride in/out = vpMatched as an index into the
 * candidate list (-1 free; any other value on input = occupied).  Greedy in list order like the reference. */
int orbb200_search_by_projection_sim3(orbb200_matcher *m, int items, const orbb200_frame_view *kf,
                                      const orbb200_fusepoints_view *pts, const float *Rcw, const float *tcw, const float *Ow,
                                      const float K[4], const float *scale_factors, int nlevels, float log_scale_factor,
                                      const float bounds[4], int th, int32_t *matched, int32_t *nmatches, int on_device);

/* A DBoW2 vocabulary (ORBVocabulary = TemplatedVocabulary<cv::Mat, FORB>, TF_IDF weighting, L1 scoring -- what
 * ORBvoc.txt declares) resident on one device, flattened: node 0 is the root; node i's children are
 * children[child_start[i] .. child_start[i+1]) in the order of its `children` vector (for a vocabulary loaded by
 * loadFromTextFile: ascending node id); descriptors n_nodes x 32; word_id and weight are read for leaves only;
 * levels = m_L. */
typedef struct orbb200_vocabulary orbb200_vocabulary;
int orbb200_vocabulary_create(int device, int n_nodes, int levels, const int32_t *child_start, const int32_t *children,
                              const uint8_t *descriptors, const int32_t *word_id, const double *weight,
                              orbb200_vocabulary **out);
void orbb200_vocabulary_destroy(orbb200_vocabulary *v);

/* Replaces mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, levelsup) of Frame::ComputeBoW /
 * KeyFrame::ComputeBoW (S/Frame.cc:520-527; Thirdparty/DBoW2 TemplatedVocabulary.h:1133-1266, BowVector.cpp:34-84;
 * scope row N4) for `items` frames (n: features per frame, desc: items x stride x 32).
 * BowVector: bow_n (items), bow_word / bow_value (items x stride): word ids ascending with their L1-normalised tf-idf
 * values, bit-identical doubles.  FeatureVector: fv_n_nodes (items), fv_node_id (items x stride, ascending),
 * fv_node_start (items x (stride+1)), fv_feat (items x stride) -- the layout orbb200_bow_view takes with
 * node_stride = stride, so on the device the result feeds orbb200_search_by_bow directly.  stride <= 8192. */
int orbb200_bow_transform(orbb200_matcher *m, const orbb200_vocabulary *voc, int items, const int32_t *n, const uint8_t *desc,
                          int stride, int levelsup, int32_t *bow_n, uint32_t *bow_word, double *bow_value,
                          int32_t *fv_n_nodes, uint32_t *fv_node_id, int32_t *fv_node_start, uint32_t *fv_feat, int on_device);

/* ------------------------------------------------------------------------------------- */
/* Frame glue (the "next" row N1 of the scope table): between extractor and matcher       */
/* ------------------------------------------------------------------------------------- */
/* K = {fx, fy, cx, cy}, dist = {k1, k2, p1, p2, k3} (Frame::mK, Frame::mDistCoef as float).
 * All three run on the matcher handle's stream. */

/* Replaces Frame::UndistortKeyPoints (S/Frame.cc:529-559) for a batch that is still on the DEVICE: reads the
 * extractor's cv::KeyPoint records (items x cap, counts per item) and writes the undistorted keypoints as
 * structure-of-arrays (items x cap each) -- exactly the arrays an orbb200_frame_view with stride = cap wants;
 * the descriptors are the extractor's descriptor output as is.  Asynchronous. */
int orbb200_frames_from_keypoints(orbb200_matcher *m, const orbb200_keypoint *d_keypoints, const int32_t *d_counts,
                                  int items, int cap, const float K[4], const float dist[5], float *d_x, float *d_y,
                                  int32_t *d_octave, float *d_angle);
/* cv::undistortPoints(src, dst, K, dist, Mat(), K) for n points in HOST memory (xy interleaved). */
int orbb200_undistort_points(orbb200_matcher *m, const float *xy_in, float *xy_out, int n, const float K[4], const float dist[5]);
/* Replaces Frame::ComputeImageBounds (S/Frame.cc:561-589): bounds = {mnMinX, mnMinY, mnMaxX, mnMaxY}. */
int orbb200_image_bounds(orbb200_matcher *m, int cols, int rows, const float K[4], const float dist[5], float bounds[4]);
/* ---- Frame::ComputeStereoMatches (S/Frame.cc:591-763) -------------------------------------------------------------
 * The image pyramids of `items` frames as the stereo matcher reads them (ORBextractor::mvImagePyramid, I/ORBextractor.h:85,
 * read at S/Frame.cc:596,686,703), WITHOUT the 19-pixel border (ComputeStereoMatches never reaches it for keypoints the
 * extractor produced).  level[l] points at frame 0; frame i of level l starts frame_stride[l] * i bytes later. */
#define ORBB200_MAX_LEVELS 16
typedef struct orbb200_pyramid_view {
    int nlevels;
    const uint8_t *level[ORBB200_MAX_LEVELS];
    size_t frame_stride[ORBB200_MAX_LEVELS];
    int32_t pitch[ORBB200_MAX_LEVELS], width[ORBB200_MAX_LEVELS], height[ORBB200_MAX_LEVELS];
} orbb200_pyramid_view;
/* The pyramids of the handle's last extract call where they lie in DEVICE memory (valid until the next extract call on
 * this handle); level 0 is the caller's own frame buffer when it was used in place. */
int orbb200_extractor_pyramid_view(orbb200_extractor *h, orbb200_pyramid_view *view);
/* `on_device` bits of orbb200_compute_stereo_matches */
#define ORBB200_DEVICE_VIEWS 1      /* frame views, u_right, depth and nmatches are device pointers (asynchronous call) */
#define ORBB200_DEVICE_PYRAMIDS 2   /* the pyramid views point at device memory (e.g. orbb200_extractor_pyramid_view) */
/* Replaces Frame::ComputeStereoMatches for `items` rectified stereo pairs: left / right = mvKeys / mDescriptors and
 * mvKeysRight / mDescriptorsRight (x, y, octave, desc; angle unused), lpyr / rpyr = the two extractors' pyramids (equal
 * level sizes), scale_factors / inv_scale_factors = mvScaleFactors / mvInvScaleFactors (HOST arrays, nlevels entries),
 * mb / mbf = Frame::mb / mbf.  Writes mvuRight and mvDepth (items x left.stride floats, -1 = no stereo match) and the
 * number of left keypoints that end with a depth per item.  A keypoint for which the reference would read outside an
 * image (it throws or is undefined there) gets no depth. */
int orbb200_compute_stereo_matches(orbb200_matcher *m, int items, const orbb200_frame_view *left,
                                   const orbb200_frame_view *right, const orbb200_pyramid_view *lpyr,
                                   const orbb200_pyramid_view *rpyr, const float *scale_factors,
                                   const float *inv_scale_factors, int nlevels, float mb, float mbf,
                                   float *u_right, float *depth, int32_t *nmatches, int on_device);

/* Makes the matcher's stream wait (on the device, no host synchronisation) for everything queued so far on
 * the extractor's stream: extract_device -> frames_from_keypoints -> search_*(on_device) is one pipeline. */
int orbb200_matcher_wait_extractor(orbb200_matcher *m, orbb200_extractor *ex);

#ifdef __cplusplus
}
#endif
#endif
