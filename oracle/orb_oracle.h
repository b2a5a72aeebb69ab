/*
 * orb_oracle.h -- CPU oracle for the ORB front-end hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference algorithm (ORB_SLAM2/src/ORBextractor.cc,
 * ORB_SLAM2/src/ORBmatcher.cc, the grid helpers of ORB_SLAM2/src/Frame.cc) and of the
 * OpenCV primitives those files call (resize INTER_LINEAR, FAST-9/16, GaussianBlur 7x7,
 * copyMakeBorder REFLECT_101, fastAtan2).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library; the product
 * (weiner_slamit_v2_b200/) never does.
 *
 * Pinning (see DESIGN.md "Oracle"): the reference ships no golden vectors for this path
 * (SURVEY.md section 4), so the pins are
 *   (1) every image primitive here == cv2 4.13 bit-for-bit (tests/test_oracle_primitives.py),
 *   (2) the whole extractor / matcher here == the reference's own unmodified .cc files
 *       compiled for x86 into oracle/_ref (tests/test_oracle_vs_ref.py),
 *   (3) committed golden vectors under tests/golden/ produced by (2).
 */
#ifndef ORB_ORACLE_H
#define ORB_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Same 28-byte layout as cv::KeyPoint (pt.x, pt.y, size, angle, response, octave, class_id). */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orc_keypoint;

/* ---- OpenCV primitive restatements ---------------------------------------------------- */
int   orc_round(float v);                 /* cvRound: round half to even */
float orc_fast_atan2(float y, float x);   /* cv::fastAtan2, degrees in [0,360] */
void  orc_sincosf(float a, float *s, float *c); /* glibc sinf/cosf restated (double polynomial) */
void  orc_resize_linear_u8(const uint8_t *src, int sw, int sh, int sstride,
                           uint8_t *dst, int dw, int dh, int dstride);
void  orc_copy_make_border_reflect101(const uint8_t *src, int w, int h, int sstride,
                                      uint8_t *dst, int dstride, int border);
/* cv::FAST(img, kps, threshold, nonmaxSuppression) with TYPE_9_16; returns the number of
 * keypoints found; at most cap are written (row-major order, response = corner score). */
int   orc_fast9_16(const uint8_t *img, int w, int h, int stride, int threshold, int nms,
                   orc_keypoint *out, int cap);
/* cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) for 8UC1.
 * variant 0: OpenCV 4.x fixed-point taps {18,34,48,56,48,34,18} (verifiable against cv2 here)
 * variant 1: OpenCV 2.4.9 cvRound(k*256) taps {18,34,49,55,49,34,18} (SURVEY hard part 5) */
void  orc_gaussian_blur7(const uint8_t *src, int w, int h, int sstride,
                         uint8_t *dst, int dstride, int variant);

/* ---- ORBextractor restatement --------------------------------------------------------- */
typedef struct orc_extractor orc_extractor;
orc_extractor *orc_extractor_create(int nfeatures, float scaleFactor, int nlevels,
                                    int iniThFAST, int minThFAST);
void orc_extractor_destroy(orc_extractor *e);
void orc_extractor_set_blur_variant(orc_extractor *e, int variant);
/* tables computed by the constructor (ORBextractor.cc:415-482) */
const float *orc_scale_factors(const orc_extractor *e);
const float *orc_inv_scale_factors(const orc_extractor *e);
const float *orc_level_sigma2(const orc_extractor *e);
const float *orc_inv_level_sigma2(const orc_extractor *e);
const int   *orc_features_per_level(const orc_extractor *e);
const int   *orc_umax(const orc_extractor *e);         /* 16 entries */
const int8_t *orc_pattern(void);                       /* 512 (x,y) int8 pairs */

/* operator() (ORBextractor.cc:1064-1136).  Returns the number of keypoints (level-major
 * order), or -1 if cap is too small, -2 on a geometry the reference cannot handle
 * (a level narrower than one 30-px cell, or aspect ratio rounding to zero root nodes). */
int orc_extract(orc_extractor *e, const uint8_t *img, int w, int h, int stride,
                orc_keypoint *kps, uint8_t *desc, int cap);

/* stage outputs of the most recent orc_extract() on this handle */
int  orc_level_width(const orc_extractor *e, int level);
int  orc_level_height(const orc_extractor *e, int level);
const uint8_t *orc_level_pixels(const orc_extractor *e, int level);   /* stride == width */
const uint8_t *orc_level_blurred(const orc_extractor *e, int level);  /* stride == width; NULL if level had no keypoints */
int  orc_level_candidates(const orc_extractor *e, int level, const orc_keypoint **p); /* FAST output, cell order, coords relative to the 16-px border */
int  orc_level_keypoints(const orc_extractor *e, int level, const orc_keypoint **p);  /* after quadtree + orientation, level coordinates */

/* DistributeOctTree alone (ORBextractor.cc:552-776); tie_break: 0 = later-created node first
 * (address order under a monotonic allocator; the canonical order), 1 = earlier-created first.
 * Returns number of keypoints written to out (cap >= N + 4*nIni is always enough). */
int orc_distribute_octree(const orc_keypoint *in, int n, int minX, int maxX, int minY, int maxY,
                          int N, int tie_break, orc_keypoint *out, int cap);

/* ---- ORBmatcher restatement (orb_matcher_oracle.c) ------------------------------------ */
int orc_descriptor_distance(const uint8_t *a, const uint8_t *b);   /* ORBmatcher.cc:1651-1667 */

/* Frame grid (Frame.cc:336-357, 447-517): 64x48 cells, CSR. */
typedef struct {
    float min_x, min_y, max_x, max_y;   /* mnMinX.. (image bounds) */
    float inv_w, inv_h;                 /* mfGridElementWidthInv/HeightInv */
    int   n;                            /* keypoints */
    const float *kx, *ky;               /* mvKeysUn pt */
    const int32_t *octave;
    int32_t cell_start[64 * 48 + 1];    /* cell = ix*48+iy */
    int32_t *cell_items;                /* n entries */
} orc_grid;
void orc_grid_bounds(orc_grid *g, const float bounds[4]);         /* {mnMinX,mnMinY,mnMaxX,mnMaxY}: Frame.cc:561-589,317-318 */
void orc_grid_assign(orc_grid *g, int n, const float *kx, const float *ky, const int32_t *octave,
                     int32_t *items_storage);                      /* AssignFeaturesToGrid */
int  orc_features_in_area(const orc_grid *g, float x, float y, float r, int minLevel, int maxLevel,
                          int32_t *out, int cap);                  /* GetFeaturesInArea */

/* SearchForInitialization (ORBmatcher.cc:409-524). prev_matched is in/out (n1 x 2 floats). */
int orc_search_for_initialization(
    int n1, const float *k1x, const float *k1y, const int32_t *k1oct, const float *k1ang, const uint8_t *d1,
    int n2, const float *k2x, const float *k2y, const int32_t *k2oct, const float *k2ang, const uint8_t *d2,
    const float bounds[4], float nnratio, int check_orientation, int window_size,
    float *prev_matched, int32_t *matches12);

/* SearchByProjection(Frame&, vector<MapPoint*>&, th) (ORBmatcher.cc:47-131), flattened:
 * per map point: track_in_view, bad, proj x/y/xr, predicted level, view cos, descriptor, observations;
 * per keypoint: pt, octave, uRight, descriptor, kp_mp (index of the map point already held, -1 none;
 * kp_mp_obs = Observations() of a foreign map point already held, used when kp_mp == -2).
 * On return kp_mp[idx] = iMP for every assignment made. */
int orc_search_by_projection(
    int nmp, const uint8_t *mp_in_view, const uint8_t *mp_bad, const float *mp_x, const float *mp_y,
    const float *mp_xr, const int32_t *mp_level, const float *mp_viewcos, const uint8_t *mp_desc,
    const int32_t *mp_obs,
    int n, const float *kx, const float *ky, const int32_t *koct, const float *kuright, const uint8_t *kdesc,
    int32_t *kp_mp, const int32_t *kp_mp_obs,
    int nlevels, const float *scale_factors, const float bounds[4], float nnratio, float th);

/* SearchByProjection(CurrentFrame, LastFrame, th, bMono) (S/ORBmatcher.cc:1332-1474); see the .c file */
int orc_search_by_projection_last_frame(
    int nlast, const uint8_t *has_mp, const uint8_t *outlier, const float *wpos, const uint8_t *mp_desc,
    const int32_t *mp_obs, const int32_t *last_octave, const float *last_angle,
    const float Rcw[9], const float tcw[3], const float K[4], float mbf,
    int n, const float *kx, const float *ky, const int32_t *koct, const float *kang, const float *kuright,
    const uint8_t *kdesc, int32_t *kp_mp, const int32_t *kp_mp_obs,
    int nlevels, const float *scale_factors, const float bounds[4], float th, int mode, int check_orientation);

/* Frame glue: cv::undistortPoints(src, dst, K, dist, Mat(), K) as Frame::UndistortKeyPoints
 * calls it (S/Frame.cc:529-559), K = {fx, fy, cx, cy}, dist = {k1, k2, p1, p2, k3}; and ComputeImageBounds. */
int orc_search_by_bow(
    int nkf, const uint8_t *kf_valid, const uint8_t *kf_desc, const float *kf_angle,
    int kf_nn, const uint32_t *kf_node, const int32_t *kf_start, const uint32_t *kf_feat,
    int nf, const uint8_t *f_desc, const float *f_angle,
    int f_nn, const uint32_t *f_node, const int32_t *f_start, const uint32_t *f_feat,
    float nnratio, int check_orientation, int32_t *matches);
int orc_search_by_bow_keyframes(
    int n1, const uint8_t *valid1, const uint8_t *desc1, const float *angle1,
    int nn1, const uint32_t *node1, const int32_t *start1, const uint32_t *feat1,
    int n2, const uint8_t *valid2, const uint8_t *desc2, const float *angle2,
    int nn2, const uint32_t *node2, const int32_t *start2, const uint32_t *feat2,
    float nnratio, int check_orientation, int32_t *matches12);
int orc_search_for_triangulation(
    int n1, const uint8_t *has_mp1, const uint8_t *desc1, const float *x1, const float *y1, const float *angle1, const float *uright1,
    int nn1, const uint32_t *node1, const int32_t *start1, const uint32_t *feat1,
    int n2, const uint8_t *has_mp2, const uint8_t *desc2, const float *x2, const float *y2, const int32_t *oct2, const float *angle2,
    const float *uright2, int nn2, const uint32_t *node2, const int32_t *start2, const uint32_t *feat2,
    const float F12[9], const float epipole[2], const float *scale_factors2, const float *level_sigma2_2,
    int only_stereo, int check_orientation, int32_t *matches12);
int orc_distinctive_descriptor(int n, const uint8_t *desc, int *median_out);
void orc_fuse_search(
    int nmp, const uint8_t *valid, const float *wpos, const float *normal, const uint8_t *mp_desc,
    const float *mf_max_distance, const float *mf_min_distance,
    const float Rcw[9], const float tcw[3], const float Ow[3], const float K[4], float bf,
    int n, const float *kx, const float *ky, const int32_t *koct, const float *kuright, const uint8_t *kdesc,
    int nlevels, const float *scale_factors, const float *inv_level_sigma2, float log_scale_factor,
    const float bounds[4], float th, int mode, const float *R2, const float *t2, int32_t *best_idx, int32_t *best_dist);
int orc_sim3_agreement(int n1, const int32_t *match1, int n2, const int32_t *match2, int32_t *matches12);
int orc_search_by_projection_sim3(
    int nmp, const uint8_t *valid, const float *wpos, const float *normal, const uint8_t *mp_desc,
    const float *mf_max_distance, const float *mf_min_distance,
    const float Rcw[9], const float tcw[3], const float Ow[3], const float K[4],
    int n, const float *kx, const float *ky, const int32_t *koct, const uint8_t *kdesc,
    int nlevels, const float *scale_factors, float log_scale_factor, const float bounds[4], int th, int32_t *matched);
void orc_bow_transform(
    int n_nodes, int L, const int32_t *child_start, const int32_t *children, const uint8_t *node_desc,
    const int32_t *word_id, const double *weight,
    int n, const uint8_t *desc, int levelsup,
    int32_t *bow_n, uint32_t *bow_word, double *bow_value,
    int32_t *fv_n, uint32_t *fv_node, int32_t *fv_start, uint32_t *fv_feat);
float orc_logf(float x);
int orc_predict_scale(float mf_max_distance, float dist, float log_scale_factor);
int orc_search_by_projection_keyframe(
    int nkf, const uint8_t *valid, const float *wpos, const uint8_t *mp_desc, const float *mf_max_distance,
    const float *mf_min_distance, const float *kf_angle,
    const float Rcw[9], const float tcw[3], const float Ow[3], const float K[4],
    int n, const float *kx, const float *ky, const int32_t *koct, const float *kang, const uint8_t *kdesc,
    int32_t *kp_mp, int nlevels, const float *scale_factors, float log_scale_factor, const float bounds[4],
    float th, int orb_dist, int check_orientation);
/* Frame::ComputeStereoMatches (Frame.cc:591-763); level images without border, same sizes left and right */
int orc_compute_stereo_matches(
    int n, const float *kx, const float *ky, const int32_t *koct, const uint8_t *desc,
    int nr, const float *rx, const float *ry, const int32_t *roct, const uint8_t *rdesc,
    int nlevels, const float *scale, const float *inv_scale,
    const uint8_t *const *limg, const int32_t *lpitch, const uint8_t *const *rimg, const int32_t *rpitch,
    const int32_t *lw, const int32_t *lh, float mb, float mbf, float *u_right, float *depth, int32_t *skipped);
void orc_undistort_points(int n, const float *xy_in, float *xy_out, const float K[4], const float dist[5]);
void orc_image_bounds(int cols, int rows, const float K[4], const float dist[5], float bounds[4]);

#ifdef __cplusplus
}
#endif
#endif
