/*
 * orb_matcher_oracle.c -- CPU oracle (TEST INFRASTRUCTURE, see orb_oracle.h) for the
 * ORBmatcher Hamming search: DescriptorDistance, SearchForInitialization and
 * SearchByProjection(Frame&, vector<MapPoint*>&, th), plus the Frame grid they query.
 *
 * S/ = /root/reference/oRB_SLAM2_Android/src/main/jni/ORB_SLAM2/src/.  The pointer graph
 * (Frame / MapPoint objects) is flattened to arrays; the loop structure, comparison
 * operators, float types and iteration orders follow the cited lines.
 */
#include "orb_oracle.h"
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "../include/orb_b200_logf.inc"

#define GRID_COLS 64   /* I/Frame.h:41 */
#define GRID_ROWS 48   /* I/Frame.h:40 */
#define TH_HIGH 100    /* S/ORBmatcher.cc:37 */
#define TH_LOW 50      /* S/ORBmatcher.cc:38 */
#define HISTO_LENGTH 30

/* S/ORBmatcher.cc:1651-1667: eight 32-bit words, SWAR population count of the XOR. */
int orc_descriptor_distance(const uint8_t *a, const uint8_t *b)
{
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t wa, wb;
        memcpy(&wa, a + 4 * i, 4);
        memcpy(&wb, b + 4 * i, 4);
        uint32_t v = wa ^ wb;
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int)((((v + (v >> 4)) & 0x0F0F0F0Fu) * 0x01010101u) >> 24);
    }
    return dist;
}

/* Image bounds as Frame::ComputeImageBounds leaves them (S/Frame.cc:561-589; {0,0,cols,rows} when
 * there is no lens distortion) and the inverse cell sizes of S/Frame.cc:317-318. */
void orc_grid_bounds(orc_grid *g, const float bounds[4])
{
    g->min_x = bounds[0]; g->max_x = bounds[2];
    g->min_y = bounds[1]; g->max_y = bounds[3];
    g->inv_w = (float)GRID_COLS / (g->max_x - g->min_x);
    g->inv_h = (float)GRID_ROWS / (g->max_y - g->min_y);
}

/* Frame::PosInGrid (S/Frame.cc:505-517): round-half-away (roundf), reject out of range. */
static int pos_in_grid(const orc_grid *g, float x, float y, int *px, int *py)
{
    *px = (int)roundf((x - g->min_x) * g->inv_w);
    *py = (int)roundf((y - g->min_y) * g->inv_h);
    return !(*px < 0 || *px >= GRID_COLS || *py < 0 || *py >= GRID_ROWS);
}

/* Frame::AssignFeaturesToGrid (S/Frame.cc:336-357): cells keep keypoint indices in
 * increasing index order.  Stored as CSR with cell = ix*GRID_ROWS + iy. */
void orc_grid_assign(orc_grid *g, int n, const float *kx, const float *ky, const int32_t *octave,
                     int32_t *items_storage)
{
    g->n = n; g->kx = kx; g->ky = ky; g->octave = octave; g->cell_items = items_storage;
    int *cnt = (int *)calloc(GRID_COLS * GRID_ROWS, sizeof(int));
    int *cell = (int *)malloc(sizeof(int) * (n + 1));
    for (int i = 0; i < n; i++) {
        int px, py;
        cell[i] = pos_in_grid(g, kx[i], ky[i], &px, &py) ? px * GRID_ROWS + py : -1;
        if (cell[i] >= 0) cnt[cell[i]]++;
    }
    g->cell_start[0] = 0;
    for (int c = 0; c < GRID_COLS * GRID_ROWS; c++) g->cell_start[c + 1] = g->cell_start[c] + cnt[c];
    memset(cnt, 0, sizeof(int) * GRID_COLS * GRID_ROWS);
    for (int i = 0; i < n; i++)
        if (cell[i] >= 0) items_storage[g->cell_start[cell[i]] + cnt[cell[i]]++] = i;
    free(cnt); free(cell);
}

/* Frame::GetFeaturesInArea (S/Frame.cc:447-502). */
int orc_features_in_area(const orc_grid *g, float x, float y, float r, int minLevel, int maxLevel,
                         int32_t *out, int cap)
{
    int n = 0;
    int minCX = (int)floorf((x - g->min_x - r) * g->inv_w);
    if (minCX < 0) minCX = 0;
    if (minCX >= GRID_COLS) return 0;
    int maxCX = (int)ceilf((x - g->min_x + r) * g->inv_w);
    if (maxCX > GRID_COLS - 1) maxCX = GRID_COLS - 1;
    if (maxCX < 0) return 0;
    int minCY = (int)floorf((y - g->min_y - r) * g->inv_h);
    if (minCY < 0) minCY = 0;
    if (minCY >= GRID_ROWS) return 0;
    int maxCY = (int)ceilf((y - g->min_y + r) * g->inv_h);
    if (maxCY > GRID_ROWS - 1) maxCY = GRID_ROWS - 1;
    if (maxCY < 0) return 0;
    const int check = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = minCX; ix <= maxCX; ix++)
        for (int iy = minCY; iy <= maxCY; iy++) {
            int c = ix * GRID_ROWS + iy;
            for (int j = g->cell_start[c]; j < g->cell_start[c + 1]; j++) {
                int idx = g->cell_items[j];
                if (check) {
                    if (g->octave[idx] < minLevel) continue;
                    if (maxLevel >= 0 && g->octave[idx] > maxLevel) continue;
                }
                float dx = g->kx[idx] - x, dy = g->ky[idx] - y;
                if (fabsf(dx) < r && fabsf(dy) < r) {
                    if (n < cap) out[n] = idx;
                    n++;
                }
            }
        }
    return n;
}

/* ORBmatcher::ComputeThreeMaxima (S/ORBmatcher.cc:1605-1646) on bin sizes. */
static void three_maxima(const int *sizes, int L, int *i1, int *i2, int *i3)
{
    int m1 = 0, m2 = 0, m3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = sizes[i];
        if (s > m1) { m3 = m2; m2 = m1; m1 = s; *i3 = *i2; *i2 = *i1; *i1 = i; }
        else if (s > m2) { m3 = m2; m2 = s; *i3 = *i2; *i2 = i; }
        else if (s > m3) { m3 = s; *i3 = i; }
    }
    if (m2 < 0.1f * (float)m1) { *i2 = -1; *i3 = -1; }
    else if (m3 < 0.1f * (float)m1) { *i3 = -1; }
}

/* S/ORBmatcher.cc:409-524 */
int orc_search_for_initialization(
    int n1, const float *k1x, const float *k1y, const int32_t *k1oct, const float *k1ang, const uint8_t *d1,
    int n2, const float *k2x, const float *k2y, const int32_t *k2oct, const float *k2ang, const uint8_t *d2,
    const float bounds[4], float nnratio, int check_orientation, int window_size,
    float *prev_matched, int32_t *matches12)
{
    (void)k1x; (void)k1y;
    int nmatches = 0;
    for (int i = 0; i < n1; i++) matches12[i] = -1;

    orc_grid g;
    int32_t *items = (int32_t *)malloc(sizeof(int32_t) * (n2 + 1));
    orc_grid_bounds(&g, bounds);
    orc_grid_assign(&g, n2, k2x, k2y, k2oct, items);

    int *hist_bin = (int *)malloc(sizeof(int) * (n1 + 1));   /* bin of i1 in insertion order */
    int *hist_i1 = (int *)malloc(sizeof(int) * (n1 + 1));
    int nhist = 0;
    const float factor = 1.0f / HISTO_LENGTH;                /* sic: the reference's bin width (:417) */

    int *matched_dist = (int *)malloc(sizeof(int) * (n2 + 1));
    int *matches21 = (int *)malloc(sizeof(int) * (n2 + 1));
    for (int i = 0; i < n2; i++) { matched_dist[i] = INT_MAX; matches21[i] = -1; }
    int32_t *cand = (int32_t *)malloc(sizeof(int32_t) * (n2 + 1));

    for (int i1 = 0; i1 < n1; i1++) {
        int level1 = k1oct[i1];
        if (level1 > 0) continue;
        int nc = orc_features_in_area(&g, prev_matched[2 * i1], prev_matched[2 * i1 + 1],
                                      (float)window_size, level1, level1, cand, n2);
        if (nc == 0) continue;
        const uint8_t *da = d1 + 32 * (size_t)i1;
        int best = INT_MAX, best2 = INT_MAX, bestIdx2 = -1;
        for (int c = 0; c < nc; c++) {
            int i2 = cand[c];
            int dist = orc_descriptor_distance(da, d2 + 32 * (size_t)i2);
            if (matched_dist[i2] <= dist) continue;
            if (dist < best) { best2 = best; best = dist; bestIdx2 = i2; }
            else if (dist < best2) best2 = dist;
        }
        if (best <= TH_LOW) {
            if ((float)best < (float)best2 * nnratio) {
                if (matches21[bestIdx2] >= 0) { matches12[matches21[bestIdx2]] = -1; nmatches--; }
                matches12[i1] = bestIdx2;
                matches21[bestIdx2] = i1;
                matched_dist[bestIdx2] = best;
                nmatches++;
                if (check_orientation) {
                    float rot = k1ang[i1] - k2ang[bestIdx2];
                    if (rot < 0.0) rot += 360.0f;
                    int bin = (int)roundf(rot * factor);
                    if (bin == HISTO_LENGTH) bin = 0;
                    hist_bin[nhist] = bin; hist_i1[nhist] = i1; nhist++;
                }
            }
        }
    }

    if (check_orientation) {
        int sizes[HISTO_LENGTH] = {0};
        for (int k = 0; k < nhist; k++) sizes[hist_bin[k]]++;
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(sizes, HISTO_LENGTH, &ind1, &ind2, &ind3);
        for (int k = 0; k < nhist; k++) {
            int b = hist_bin[k];
            if (b == ind1 || b == ind2 || b == ind3) continue;
            if (matches12[hist_i1[k]] >= 0) { matches12[hist_i1[k]] = -1; nmatches--; }
        }
    }

    for (int i1 = 0; i1 < n1; i1++)
        if (matches12[i1] >= 0) {
            prev_matched[2 * i1] = k2x[matches12[i1]];
            prev_matched[2 * i1 + 1] = k2y[matches12[i1]];
        }

    free(cand); free(matches21); free(matched_dist); free(hist_i1); free(hist_bin); free(items);
    return nmatches;
}

/* S/ORBmatcher.cc:47-139 */
int orc_search_by_projection(
    int nmp, const uint8_t *mp_in_view, const uint8_t *mp_bad, const float *mp_x, const float *mp_y,
    const float *mp_xr, const int32_t *mp_level, const float *mp_viewcos, const uint8_t *mp_desc,
    const int32_t *mp_obs,
    int n, const float *kx, const float *ky, const int32_t *koct, const float *kuright, const uint8_t *kdesc,
    int32_t *kp_mp, const int32_t *kp_mp_obs,
    int nlevels, const float *scale_factors, const float bounds[4], float nnratio, float th)
{
    (void)nlevels;
    int nmatches = 0;
    const int bFactor = th != 1.0;
    orc_grid g;
    int32_t *items = (int32_t *)malloc(sizeof(int32_t) * (n + 1));
    int32_t *cand = (int32_t *)malloc(sizeof(int32_t) * (n + 1));
    orc_grid_bounds(&g, bounds);
    orc_grid_assign(&g, n, kx, ky, koct, items);

    for (int i = 0; i < nmp; i++) {
        if (!mp_in_view[i]) continue;
        if (mp_bad[i]) continue;
        const int lvl = mp_level[i];
        float r = ((double)mp_viewcos[i] > 0.998) ? 2.5f : 4.0f;     /* RadiusByViewingCos (:133-139) */
        if (bFactor) r *= th;
        const float rs = r * scale_factors[lvl];
        int nc = orc_features_in_area(&g, mp_x[i], mp_y[i], rs, lvl - 1, lvl, cand, n);
        if (nc == 0) continue;
        const uint8_t *dm = mp_desc + 32 * (size_t)i;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int c = 0; c < nc; c++) {
            const int idx = cand[c];
            if (kp_mp[idx] != -1) {
                int obs = kp_mp[idx] >= 0 ? mp_obs[kp_mp[idx]] : kp_mp_obs[idx];
                if (obs > 0) continue;
            }
            if (kuright[idx] > 0) {
                const float er = fabsf(mp_xr[i] - kuright[idx]);
                if (er > rs) continue;
            }
            const int dist = orc_descriptor_distance(dm, kdesc + 32 * (size_t)idx);
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist;
                bestLevel2 = bestLevel; bestLevel = koct[idx];
                bestIdx = idx;
            } else if (dist < bestDist2) {
                bestLevel2 = koct[idx];
                bestDist2 = dist;
            }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && (float)bestDist > nnratio * (float)bestDist2) continue;
            kp_mp[bestIdx] = i;
            nmatches++;
        }
    }
    free(cand); free(items);
    return nmatches;
}

/* ======================================================================================= */
/* Frame glue ("next" row N1): UndistortKeyPoints (S/Frame.cc:529-559) and ComputeImageBounds */
/* (S/Frame.cc:561-589), i.e. cv::undistortPoints(src, dst, K, distCoef, Mat(), K).           */
/* ======================================================================================= */
/* OpenCV's iterative inverse of the plumb-bob model, 5 fixed iterations, everything in double:
 *   x0 = (u - cx)/fx, y0 = (v - cy)/fy;  repeat 5x:  r2 = x*x + y*y,
 *   icdist = (1 + ((k6*r2 + k5)*r2 + k4)*r2) / (1 + ((k3*r2 + k2)*r2 + k1)*r2),
 *   dX = 2*p1*x*y + p2*(r2 + 2*x*x), dY = p1*(r2 + 2*y*y) + 2*p2*x*y,
 *   x = (x0 - dX)*icdist, y = (y0 - dY)*icdist;
 * then re-projected with P = K.  k[] = {k1, k2, p1, p2, k3}; K and k are float matrices in the
 * reference (Tracking.cc:77-112) and are widened to double like OpenCV does.  Pinned bit-exactly
 * against cv2.undistortPoints in tests/test_oracle_primitives.py. */
void orc_undistort_points(int n, const float *xy_in, float *xy_out, const float K[4], const float dist[5])
{
    const double fx = K[0], fy = K[1], cx = K[2], cy = K[3];
    const double ifx = 1. / fx, ify = 1. / fy;
    const double k1 = dist[0], k2 = dist[1], p1 = dist[2], p2 = dist[3], k3 = dist[4];
    for (int i = 0; i < n; i++) {
        double x = xy_in[2 * i], y = xy_in[2 * i + 1];
        const double u = x, v = y;
        x = (x - cx) * ifx;
        y = (y - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; j++) {
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((0 * r2 + 0) * r2 + 0) * r2) / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);
            if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
            const double deltaX = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);
            const double deltaY = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        /* P = K, R = identity: xx = fx*x + 0*y + cx, ww = 1/(0*x + 0*y + 1) */
        const double xx = fx * x + 0 * y + cx, yy = 0 * x + fy * y + cy;
        const double ww = 1. / (0 * x + 0 * y + 1);
        xy_out[2 * i] = (float)(xx * ww);
        xy_out[2 * i + 1] = (float)(yy * ww);
    }
}

/* Frame::ComputeImageBounds (S/Frame.cc:561-589): bounds[4] = {mnMinX, mnMinY, mnMaxX, mnMaxY}. */
void orc_image_bounds(int cols, int rows, const float K[4], const float dist[5], float bounds[4])
{
    if (dist[0] != 0.0f) {
        float c[8] = {0.f, 0.f, (float)cols, 0.f, 0.f, (float)rows, (float)cols, (float)rows}, o[8];
        orc_undistort_points(4, c, o, K, dist);
        bounds[0] = o[0] < o[4] ? o[0] : o[4];     /* min(mat(0,0), mat(2,0)) */
        bounds[2] = o[2] > o[6] ? o[2] : o[6];     /* max(mat(1,0), mat(3,0)) */
        bounds[1] = o[1] < o[3] ? o[1] : o[3];     /* min(mat(0,1), mat(1,1)) */
        bounds[3] = o[5] > o[7] ? o[5] : o[7];     /* max(mat(2,1), mat(3,1)) */
    } else {
        bounds[0] = 0.0f; bounds[2] = (float)cols; bounds[1] = 0.0f; bounds[3] = (float)rows;
    }
}

/* ======================================================================================= */
/* "next" row N2: SearchByProjection(CurrentFrame, LastFrame, th, bMono) (S/ORBmatcher.cc:1332-1474) */
/* ======================================================================================= */
/* Flattened inputs.  Last frame, per keypoint i: has_mp (mvpMapPoints[i] != NULL), outlier (mvbOutlier[i]),
 * wpos (GetWorldPos(), 3 floats), mp_desc (GetDescriptor()), mp_obs (Observations()), octave (mvKeys[i].octave),
 * angle (mvKeysUn[i].angle).  Current frame: Tcw as Rcw (3x3 row-major) and tcw, camera {fx, fy, cx, cy}, mbf,
 * bounds, scale factors, keypoints, uRight, descriptors, kp_mp in/out (index into the LAST frame's arrays;
 * -1 none; -2 foreign map point with kp_mp_obs observations).  mode: 0 = window on octave-1..octave+1 (the
 * monocular case and small baseline motion), 1 = bForward, 2 = bBackward (:1352-1353, decided by the caller
 * from tlc and mb).  The 3x3 float product follows cv::gemm's small-matrix path: float multiplies and adds
 * in source order (pinned against cv2.gemm in tests/test_oracle_primitives.py). */
int orc_search_by_projection_last_frame(
    int nlast, const uint8_t *has_mp, const uint8_t *outlier, const float *wpos, const uint8_t *mp_desc,
    const int32_t *mp_obs, const int32_t *last_octave, const float *last_angle,
    const float Rcw[9], const float tcw[3], const float K[4], float mbf,
    int n, const float *kx, const float *ky, const int32_t *koct, const float *kang, const float *kuright,
    const uint8_t *kdesc, int32_t *kp_mp, const int32_t *kp_mp_obs,
    int nlevels, const float *scale_factors, const float bounds[4], float th, int mode, int check_orientation)
{
    (void)nlevels;
    int nmatches = 0;
    orc_grid g;
    int32_t *items = (int32_t *)malloc(sizeof(int32_t) * (n + 1));
    int32_t *cand = (int32_t *)malloc(sizeof(int32_t) * (n + 1));
    int *hist_bin = (int *)malloc(sizeof(int) * (nlast + 1));
    int *hist_idx = (int *)malloc(sizeof(int) * (nlast + 1));
    int nhist = 0;
    const float factor = 1.0f / HISTO_LENGTH;
    orc_grid_bounds(&g, bounds);
    orc_grid_assign(&g, n, kx, ky, koct, items);
    const float fx = K[0], fy = K[1], cx = K[2], cy = K[3];

    for (int i = 0; i < nlast; i++) {
        if (!has_mp[i] || outlier[i]) continue;
        const float *X = wpos + 3 * (size_t)i;
        float xc3[3];
        for (int r = 0; r < 3; r++) {                       /* x3Dc = Rcw*x3Dw + tcw (:1363) */
            volatile float t = Rcw[3 * r] * X[0];
            volatile float t1 = Rcw[3 * r + 1] * X[1];
            volatile float t2 = Rcw[3 * r + 2] * X[2];
            t = t + t1;
            t = t + t2;
            xc3[r] = t + tcw[r];
        }
        const float xc = xc3[0], yc = xc3[1];
        const float invzc = (float)(1.0 / (double)xc3[2]);
        if (invzc < 0) continue;
        volatile float u = fx * xc; u = u * invzc; u = u + cx;
        volatile float v = fy * yc; v = v * invzc; v = v + cy;
        if (u < bounds[0] || u > bounds[2]) continue;
        if (v < bounds[1] || v > bounds[3]) continue;
        const int oct = last_octave[i];
        const float radius = th * scale_factors[oct];
        int nc;
        if (mode == 1) nc = orc_features_in_area(&g, u, v, radius, oct, -1, cand, n);
        else if (mode == 2) nc = orc_features_in_area(&g, u, v, radius, 0, oct, cand, n);
        else nc = orc_features_in_area(&g, u, v, radius, oct - 1, oct + 1, cand, n);
        if (nc == 0) continue;
        const uint8_t *dm = mp_desc + 32 * (size_t)i;
        int bestDist = 256, bestIdx2 = -1;
        for (int c = 0; c < nc; c++) {
            const int i2 = cand[c];
            if (kp_mp[i2] != -1) {
                const int obs = kp_mp[i2] >= 0 ? mp_obs[kp_mp[i2]] : kp_mp_obs[i2];
                if (obs > 0) continue;
            }
            if (kuright[i2] > 0) {
                volatile float ur = mbf * invzc;
                ur = u - ur;
                const float er = fabsf(ur - kuright[i2]);
                if (er > radius) continue;
            }
            const int dist = orc_descriptor_distance(dm, kdesc + 32 * (size_t)i2);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= TH_HIGH) {
            kp_mp[bestIdx2] = i;
            nmatches++;
            if (check_orientation) {
                float rot = last_angle[i] - kang[bestIdx2];
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)roundf(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                hist_bin[nhist] = bin; hist_idx[nhist] = bestIdx2; nhist++;
            }
        }
    }
    if (check_orientation) {
        int sizes[HISTO_LENGTH] = {0};
        for (int k = 0; k < nhist; k++) sizes[hist_bin[k]]++;
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(sizes, HISTO_LENGTH, &ind1, &ind2, &ind3);
        for (int k = 0; k < nhist; k++) {
            const int b = hist_bin[k];
            if (b != ind1 && b != ind2 && b != ind3) { kp_mp[hist_idx[k]] = -1; nmatches--; }   /* :1460-1465 */
        }
    }
    free(hist_idx); free(hist_bin); free(cand); free(items);
    return nmatches;
}


/* ------------------------------------------------------------------------------------------------
 * glibc >= 2.27 logf (ARM optimized-routines): table of 16 {1/c, log c}, degree-3 polynomial in double, one
 * rounding to float.  MapPoint::PredictScale (S/MapPoint.cc:391-400) calls it through std::log(float), so the
 * predicted pyramid level is only bit-identical with a bit-identical logf.  Identical to this machine's libm on
 * every positive finite float (2,139,095,039 values checked; tests/test_oracle_primitives.py re-checks a sample). */
float orc_logf(float x)
{
    static const struct { double invc, logc; } T[16] = { ORB_B200_LOGF_TABLE };
    uint32_t ix;
    memcpy(&ix, &x, 4);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2 == 0) return -INFINITY;
        if (ix == 0x7f800000u) return x;
        if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return NAN;
        const float xs = x * 0x1p23f;                       /* subnormal: normalise */
        memcpy(&ix, &xs, 4);
        ix -= 23u << 23;
    }
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) % 16);
    const int k = (int32_t)tmp >> 23;
    const uint32_t iz = ix - (tmp & 0xff800000u);
    float zf;
    memcpy(&zf, &iz, 4);
    const double z = (double)zf;
    const double r = z * T[i].invc - 1;
    const double y0 = T[i].logc + (double)k * ORB_B200_LOGF_LN2;
    const double r2 = r * r;
    double y = ORB_B200_LOGF_A1 * r + ORB_B200_LOGF_A2;
    y = ORB_B200_LOGF_A0 * r2 + y;
    y = y * r2 + (y0 + r);
    return (float)y;
}

/* MapPoint::PredictScale (S/MapPoint.cc:391-400): ceil(logf(mfMaxDistance / dist) / logScaleFactor), all float. */
int orc_predict_scale(float mf_max_distance, float dist, float log_scale_factor)
{
    const float ratio = mf_max_distance / dist;
    volatile float q = orc_logf(ratio);
    q = q / log_scale_factor;
    return (int)ceilf(q);
}

/* ------------------------------------------------------------------------------------------------
 * ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound,
 * const float th, const int ORBdist) (S/ORBmatcher.cc:1476-1603), the relocalisation search.
 * Flattened inputs: the key frame's map-point slots i = 0..nkf-1 with valid[i] = (pMP && !pMP->isBad() &&
 * !sAlreadyFound.count(pMP)), world position, descriptor, raw mfMaxDistance / mfMinDistance (the invariance
 * getters multiply by 1.2f / 0.8f, S/MapPoint.cc:379-389) and pKF->mvKeysUn[i].angle; the current frame as
 * for the other searches; Ow = -Rcw^T tcw is an input (the caller's cv::Mat expression).  kp_mp in/out:
 * CurrentFrame.mvpMapPoints as an index into the key frame's slots (-1 none, any other value = occupied).
 * cv::norm of the 3-vector accumulates squares in double.
 * The reference indexes mvScaleFactors with an unclamped predicted level (out of range for distances in
 * [0.8 mfMin, mfMin) -- undefined behaviour, fixed upstream later by clamping); this restatement clamps to
 * [0, nlevels-1] like the upstream fix, and the parity tests stay off that band. */
int orc_search_by_projection_keyframe(
    int nkf, const uint8_t *valid, const float *wpos, const uint8_t *mp_desc, const float *mf_max_distance,
    const float *mf_min_distance, const float *kf_angle,
    const float Rcw[9], const float tcw[3], const float Ow[3], const float K[4],
    int n, const float *kx, const float *ky, const int32_t *koct, const float *kang, const uint8_t *kdesc,
    int32_t *kp_mp, int nlevels, const float *scale_factors, float log_scale_factor, const float bounds[4],
    float th, int orb_dist, int check_orientation)
{
    int nmatches = 0;
    orc_grid g;
    int32_t *items = (int32_t *)malloc(sizeof(int32_t) * (n + 1));
    int32_t *cand = (int32_t *)malloc(sizeof(int32_t) * (n + 1));
    int *hist_bin = (int *)malloc(sizeof(int) * (nkf + 1));
    int *hist_idx = (int *)malloc(sizeof(int) * (nkf + 1));
    int nhist = 0;
    const float factor = 1.0f / HISTO_LENGTH;
    orc_grid_bounds(&g, bounds);
    orc_grid_assign(&g, n, kx, ky, koct, items);
    const float fx = K[0], fy = K[1], cx = K[2], cy = K[3];

    for (int i = 0; i < nkf; i++) {
        if (!valid[i]) continue;
        const float *X = wpos + 3 * (size_t)i;
        float xc3[3];
        for (int r = 0; r < 3; r++) {                       /* x3Dc = Rcw*x3Dw + tcw (:1503) */
            volatile float t = Rcw[3 * r] * X[0];
            volatile float t1 = Rcw[3 * r + 1] * X[1];
            volatile float t2 = Rcw[3 * r + 2] * X[2];
            t = t + t1;
            t = t + t2;
            xc3[r] = t + tcw[r];
        }
        const float invzc = (float)(1.0 / (double)xc3[2]);  /* no positive-depth test in this overload */
        volatile float u = fx * xc3[0]; u = u * invzc; u = u + cx;
        volatile float v = fy * xc3[1]; v = v * invzc; v = v + cy;
        if (u < bounds[0] || u > bounds[2]) continue;
        if (v < bounds[1] || v > bounds[3]) continue;
        double ss = 0.0;                                    /* cv::norm(x3Dw - Ow) (:1518-1519) */
        for (int r = 0; r < 3; r++) { const float po = X[r] - Ow[r]; ss += (double)po * (double)po; }
        const float dist3D = (float)sqrt(ss);
        const float maxDistance = 1.2f * mf_max_distance[i], minDistance = 0.8f * mf_min_distance[i];
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        int level = orc_predict_scale(mf_max_distance[i], dist3D, log_scale_factor);
        if (level < 0) level = 0;
        if (level >= nlevels) level = nlevels - 1;
        const float radius = th * scale_factors[level];
        const int nc = orc_features_in_area(&g, u, v, radius, level - 1, level + 1, cand, n);
        if (nc == 0) continue;
        const uint8_t *dm = mp_desc + 32 * (size_t)i;
        int bestDist = 256, bestIdx2 = -1;
        for (int c = 0; c < nc; c++) {
            const int i2 = cand[c];
            if (kp_mp[i2] != -1) continue;
            const int dist = orc_descriptor_distance(dm, kdesc + 32 * (size_t)i2);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= orb_dist) {
            kp_mp[bestIdx2] = i;
            nmatches++;
            if (check_orientation) {
                float rot = kf_angle[i] - kang[bestIdx2];
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)roundf(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                hist_bin[nhist] = bin; hist_idx[nhist] = bestIdx2; nhist++;
            }
        }
    }
    if (check_orientation) {
        int sizes[HISTO_LENGTH] = {0};
        for (int k = 0; k < nhist; k++) sizes[hist_bin[k]]++;
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(sizes, HISTO_LENGTH, &ind1, &ind2, &ind3);
        for (int k = 0; k < nhist; k++) {
            const int b = hist_bin[k];
            if (b != ind1 && b != ind2 && b != ind3) { kp_mp[hist_idx[k]] = -1; nmatches--; }   /* :1591-1597 */
        }
    }
    free(hist_idx); free(hist_bin); free(cand); free(items);
    return nmatches;
}


/* ------------------------------------------------------------------------------------------------
 * ORBmatcher::SearchByBoW(KeyFrame *pKF, Frame &F, vector<MapPoint*> &vpMapPointMatches)
 * (S/ORBmatcher.cc:161-292): features are compared only inside a shared vocabulary node.
 * A DBoW2::FeatureVector (std::map<NodeId, vector<unsigned>>) is flattened to node ids (ascending, the map's
 * order), node_start (nn+1 offsets) and the feature indices node after node.  kf_valid[i] = the key frame's
 * slot i holds a map point that is not bad; kf_angle = pKF->mvKeysUn[].angle, f_angle = F.mvKeys[].angle.
 * matches[nf] out: the key-frame slot whose map point the frame keypoint received, or -1.  Returns nmatches. */
int orc_search_by_bow(
    int nkf, const uint8_t *kf_valid, const uint8_t *kf_desc, const float *kf_angle,
    int kf_nn, const uint32_t *kf_node, const int32_t *kf_start, const uint32_t *kf_feat,
    int nf, const uint8_t *f_desc, const float *f_angle,
    int f_nn, const uint32_t *f_node, const int32_t *f_start, const uint32_t *f_feat,
    float nnratio, int check_orientation, int32_t *matches)
{
    (void)nkf;
    int nmatches = 0;
    int *hist_bin = (int *)malloc(sizeof(int) * (nf + 1));
    int *hist_idx = (int *)malloc(sizeof(int) * (nf + 1));
    int nhist = 0;
    const float factor = 1.0f / HISTO_LENGTH;
    for (int i = 0; i < nf; i++) matches[i] = -1;
    int a = 0, b = 0;
    while (a < kf_nn && b < f_nn) {
        if (kf_node[a] == f_node[b]) {
            for (int ik = kf_start[a]; ik < kf_start[a + 1]; ik++) {
                const uint32_t realIdxKF = kf_feat[ik];
                if (!kf_valid[realIdxKF]) continue;
                const uint8_t *dKF = kf_desc + 32 * (size_t)realIdxKF;
                int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
                for (int jf = f_start[b]; jf < f_start[b + 1]; jf++) {
                    const uint32_t realIdxF = f_feat[jf];
                    if (matches[realIdxF] != -1) continue;
                    const int dist = orc_descriptor_distance(dKF, f_desc + 32 * (size_t)realIdxF);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = (int)realIdxF; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 <= TH_LOW) {
                    if ((float)bestDist1 < nnratio * (float)bestDist2) {
                        matches[bestIdxF] = (int32_t)realIdxKF;
                        if (check_orientation) {
                            float rot = kf_angle[realIdxKF] - f_angle[bestIdxF];
                            if (rot < 0.0) rot += 360.0f;
                            int bin = (int)roundf(rot * factor);
                            if (bin == HISTO_LENGTH) bin = 0;
                            hist_bin[nhist] = bin; hist_idx[nhist] = bestIdxF; nhist++;
                        }
                        nmatches++;
                    }
                }
            }
            a++; b++;
        } else if (kf_node[a] < f_node[b]) {
            while (a < kf_nn && kf_node[a] < f_node[b]) a++;       /* lower_bound (:270) */
        } else {
            while (b < f_nn && f_node[b] < kf_node[a]) b++;
        }
    }
    if (check_orientation) {
        int sizes[HISTO_LENGTH] = {0};
        for (int k = 0; k < nhist; k++) sizes[hist_bin[k]]++;
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(sizes, HISTO_LENGTH, &ind1, &ind2, &ind3);
        for (int k = 0; k < nhist; k++) {
            const int bn = hist_bin[k];
            if (bn != ind1 && bn != ind2 && bn != ind3) { matches[hist_idx[k]] = -1; nmatches--; }
        }
    }
    free(hist_idx); free(hist_bin);
    return nmatches;
}


/* ------------------------------------------------------------------------------------------------
 * ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint*> &vpMatches12)
 * (S/ORBmatcher.cc:526-659, loop closing): both sides need a good map point, the acceptance test is
 * bestDist1 < TH_LOW (strict), occupancy is vbMatched2, and the result is indexed by key frame 1:
 * matches12[n1] out = the slot of key frame 2 whose map point slot idx1 received, or -1. */
int orc_search_by_bow_keyframes(
    int n1, const uint8_t *valid1, const uint8_t *desc1, const float *angle1,
    int nn1, const uint32_t *node1, const int32_t *start1, const uint32_t *feat1,
    int n2, const uint8_t *valid2, const uint8_t *desc2, const float *angle2,
    int nn2, const uint32_t *node2, const int32_t *start2, const uint32_t *feat2,
    float nnratio, int check_orientation, int32_t *matches12)
{
    int nmatches = 0;
    uint8_t *matched2 = (uint8_t *)calloc((size_t)n2 + 1, 1);
    int *hist_bin = (int *)malloc(sizeof(int) * (n1 + 1));
    int *hist_idx = (int *)malloc(sizeof(int) * (n1 + 1));
    int nhist = 0;
    const float factor = 1.0f / HISTO_LENGTH;
    for (int i = 0; i < n1; i++) matches12[i] = -1;
    int a = 0, b = 0;
    while (a < nn1 && b < nn2) {
        if (node1[a] == node2[b]) {
            for (int i1 = start1[a]; i1 < start1[a + 1]; i1++) {
                const uint32_t idx1 = feat1[i1];
                if (!valid1[idx1]) continue;
                const uint8_t *d1 = desc1 + 32 * (size_t)idx1;
                int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                for (int i2 = start2[b]; i2 < start2[b + 1]; i2++) {
                    const uint32_t idx2 = feat2[i2];
                    if (matched2[idx2] || !valid2[idx2]) continue;
                    const int dist = orc_descriptor_distance(d1, desc2 + 32 * (size_t)idx2);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = (int)idx2; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 < TH_LOW) {                                   /* :601, strict */
                    if ((float)bestDist1 < nnratio * (float)bestDist2) {
                        matches12[idx1] = bestIdx2;
                        matched2[bestIdx2] = 1;
                        if (check_orientation) {
                            float rot = angle1[idx1] - angle2[bestIdx2];
                            if (rot < 0.0) rot += 360.0f;
                            int bin = (int)roundf(rot * factor);
                            if (bin == HISTO_LENGTH) bin = 0;
                            hist_bin[nhist] = bin; hist_idx[nhist] = (int)idx1; nhist++;
                        }
                        nmatches++;
                    }
                }
            }
            a++; b++;
        } else if (node1[a] < node2[b]) {
            while (a < nn1 && node1[a] < node2[b]) a++;
        } else {
            while (b < nn2 && node2[b] < node1[a]) b++;
        }
    }
    if (check_orientation) {
        int sizes[HISTO_LENGTH] = {0};
        for (int k = 0; k < nhist; k++) sizes[hist_bin[k]]++;
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(sizes, HISTO_LENGTH, &ind1, &ind2, &ind3);
        for (int k = 0; k < nhist; k++) {
            const int bn = hist_bin[k];
            if (bn != ind1 && bn != ind2 && bn != ind3) { matches12[hist_idx[k]] = -1; nmatches--; }
        }
    }
    free(hist_idx); free(hist_bin); free(matched2);
    return nmatches;
}


/* ------------------------------------------------------------------------------------------------
 * ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo) (S/ORBmatcher.cc:661-827):
 * for every key-frame-1 feature without a map point, the feature of the same vocabulary node in key frame 2
 * (also without map point) with the smallest descriptor distance <= TH_LOW that is not too close to the epipole
 * and lies near the epipolar line; among equal distances the LAST one in list order wins (`dist > bestDist`
 * rejects, equality replaces).  vbMatched2 is never set in this version, so features are independent.
 * epipole = {ex, ey} (:668-675, the caller's cv::Mat expression).  matches12[n1] out: index in key frame 2 or -1. */
int orc_search_for_triangulation(
    int n1, const uint8_t *has_mp1, const uint8_t *desc1, const float *x1, const float *y1, const float *angle1, const float *uright1,
    int nn1, const uint32_t *node1, const int32_t *start1, const uint32_t *feat1,
    int n2, const uint8_t *has_mp2, const uint8_t *desc2, const float *x2, const float *y2, const int32_t *oct2, const float *angle2,
    const float *uright2, int nn2, const uint32_t *node2, const int32_t *start2, const uint32_t *feat2,
    const float F12[9], const float epipole[2], const float *scale_factors2, const float *level_sigma2_2,
    int only_stereo, int check_orientation, int32_t *matches12)
{
    (void)n2;
    int nmatches = 0;
    int *hist_bin = (int *)malloc(sizeof(int) * (n1 + 1));
    int *hist_idx = (int *)malloc(sizeof(int) * (n1 + 1));
    int nhist = 0;
    const float factor = 1.0f / HISTO_LENGTH;
    const float ex = epipole[0], ey = epipole[1];
    for (int i = 0; i < n1; i++) matches12[i] = -1;
    int a = 0, b = 0;
    while (a < nn1 && b < nn2) {
        if (node1[a] == node2[b]) {
            for (int i1 = start1[a]; i1 < start1[a + 1]; i1++) {
                const uint32_t idx1 = feat1[i1];
                if (has_mp1[idx1]) continue;
                const int stereo1 = uright1[idx1] >= 0;
                if (only_stereo && !stereo1) continue;
                int bestDist = TH_LOW, bestIdx2 = -1;
                for (int i2 = start2[b]; i2 < start2[b + 1]; i2++) {
                    const uint32_t idx2 = feat2[i2];
                    if (has_mp2[idx2]) continue;
                    const int stereo2 = uright2[idx2] >= 0;
                    if (only_stereo && !stereo2) continue;
                    const int dist = orc_descriptor_distance(desc1 + 32 * (size_t)idx1, desc2 + 32 * (size_t)idx2);
                    if (dist > TH_LOW || dist > bestDist) continue;
                    if (!stereo1 && !stereo2) {
                        const float distex = ex - x2[idx2], distey = ey - y2[idx2];
                        volatile float p0 = distex * distex, p1 = distey * distey;
                        const float d2 = p0 + p1;
                        const float lim = 100 * scale_factors2[oct2[idx2]];
                        if (d2 < lim) continue;
                    }
                    /* CheckDistEpipolarLine (:142-159) */
                    volatile float t0, t1;
                    t0 = x1[idx1] * F12[0]; t1 = y1[idx1] * F12[3]; t0 = t0 + t1; const float la = t0 + F12[6];
                    t0 = x1[idx1] * F12[1]; t1 = y1[idx1] * F12[4]; t0 = t0 + t1; const float lb = t0 + F12[7];
                    t0 = x1[idx1] * F12[2]; t1 = y1[idx1] * F12[5]; t0 = t0 + t1; const float lc = t0 + F12[8];
                    t0 = la * x2[idx2]; t1 = lb * y2[idx2]; t0 = t0 + t1; const float num = t0 + lc;
                    t0 = la * la; t1 = lb * lb; const float den = t0 + t1;
                    if (den == 0) continue;
                    t0 = num * num; const float dsqr = t0 / den;
                    if ((double)dsqr < 3.84 * (double)level_sigma2_2[oct2[idx2]]) { bestIdx2 = (int)idx2; bestDist = dist; }
                }
                if (bestIdx2 >= 0) {
                    matches12[idx1] = bestIdx2;
                    nmatches++;
                    if (check_orientation) {
                        float rot = angle1[idx1] - angle2[bestIdx2];
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)roundf(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        hist_bin[nhist] = bin; hist_idx[nhist] = (int)idx1; nhist++;
                    }
                }
            }
            a++; b++;
        } else if (node1[a] < node2[b]) {
            while (a < nn1 && node1[a] < node2[b]) a++;
        } else {
            while (b < nn2 && node2[b] < node1[a]) b++;
        }
    }
    if (check_orientation) {
        int sizes[HISTO_LENGTH] = {0};
        for (int k = 0; k < nhist; k++) sizes[hist_bin[k]]++;
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(sizes, HISTO_LENGTH, &ind1, &ind2, &ind3);
        for (int k = 0; k < nhist; k++) {
            const int bn = hist_bin[k];
            if (bn != ind1 && bn != ind2 && bn != ind3) { matches12[hist_idx[k]] = -1; nmatches--; }
        }
    }
    free(hist_idx); free(hist_bin);
    return nmatches;
}


/* ------------------------------------------------------------------------------------------------
 * The selection of MapPoint::ComputeDistinctiveDescriptors (S/MapPoint.cc:248-313): of the n observed descriptors
 * (those of the map point's non-bad key frames, in std::map order) the one with the least median Hamming distance
 * to all of them (its own 0 included); median = element (int)(0.5*(n-1)) of the sorted row; first minimum wins.
 * Returns the index, -1 for n == 0; *median_out receives the winning median. */
static int cmp_int(const void *a, const void *b) { return *(const int *)a - *(const int *)b; }
int orc_distinctive_descriptor(int n, const uint8_t *desc, int *median_out)
{
    if (n <= 0) return -1;
    int *row = (int *)malloc(sizeof(int) * n);
    int bestMedian = INT_MAX, bestIdx = 0;
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < n; j++) row[j] = i == j ? 0 : orc_descriptor_distance(desc + 32 * (size_t)i, desc + 32 * (size_t)j);
        qsort(row, n, sizeof(int), cmp_int);
        const int median = row[(int)(0.5 * (n - 1))];
        if (median < bestMedian) { bestMedian = median; bestIdx = i; }
    }
    free(row);
    if (median_out) *median_out = bestMedian;
    return bestIdx;
}


/* ------------------------------------------------------------------------------------------------
 * The search inside ORBmatcher::Fuse(KeyFrame *pKF, const vector<MapPoint*> &vpMapPoints, th)
 * (S/ORBmatcher.cc:829-975, up to the "replace or add" surgery, which stays on the host): for every candidate map
 * point the most similar key-frame keypoint around its projection.  valid[i] = pMP && !isBad && !IsInKeyFrame(pKF)
 * at the time the point is visited.  bounds = the Frame's float image bounds: the key frame's grid was assigned
 * with them, while its queries (GetFeaturesInArea, IsInImage) use the int-truncated copies (S/KeyFrame.cc:42,
 * 577-621).  best_idx[i] = keypoint index when bestDist <= TH_LOW, else -1; best_dist[i] = that distance (256: none).
 * Mat::dot accumulates in double; PredictScale is clamped as in the other projection searches. */
void orc_fuse_search(
    int nmp, const uint8_t *valid, const float *wpos, const float *normal, const uint8_t *mp_desc,
    const float *mf_max_distance, const float *mf_min_distance,
    const float Rcw[9], const float tcw[3], const float Ow[3], const float K[4], float bf,
    int n, const float *kx, const float *ky, const int32_t *koct, const float *kuright, const uint8_t *kdesc,
    int nlevels, const float *scale_factors, const float *inv_level_sigma2, float log_scale_factor,
    const float bounds[4], float th, int mode, const float *R2, const float *t2, int32_t *best_idx, int32_t *best_dist)
{
    /* mode 0: Fuse(pKF, vpMapPoints, th); mode 1: Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (:979-1104: no
     * reprojection-error gates); mode 2: one leg of SearchBySim3 (:1106-1330): the camera point goes through a second
     * similarity (R2 = sR21 or sR12, t2), dist3D is the norm of the resulting camera point, there is no viewing-angle
     * gate and the acceptance threshold is TH_HIGH. */
    const int accept = mode == 2 ? TH_HIGH : TH_LOW;
    orc_grid g, q;
    int32_t *items = (int32_t *)malloc(sizeof(int32_t) * (n + 1));
    int32_t *cand = (int32_t *)malloc(sizeof(int32_t) * (n + 1));
    orc_grid_bounds(&g, bounds);
    orc_grid_assign(&g, n, kx, ky, koct, items);
    q = g;                                                  /* the key frame's own copies of the bounds are ints */
    q.min_x = (float)(int)bounds[0]; q.min_y = (float)(int)bounds[1];
    const int maxXi = (int)bounds[2], maxYi = (int)bounds[3];
    const float fx = K[0], fy = K[1], cx = K[2], cy = K[3];
    for (int i = 0; i < nmp; i++) {
        best_idx[i] = -1; if (best_dist) best_dist[i] = 256;
        if (!valid[i]) continue;
        const float *X = wpos + 3 * (size_t)i;
        float c3[3];
        for (int r = 0; r < 3; r++) {
            volatile float t = Rcw[3 * r] * X[0];
            volatile float t1 = Rcw[3 * r + 1] * X[1];
            volatile float t2 = Rcw[3 * r + 2] * X[2];
            t = t + t1; t = t + t2;
            c3[r] = t + tcw[r];
        }
        if (mode == 2) {                                    /* p3Dc2 = sR21*p3Dc1 + t21 (:1157) */
            float d3[3];
            for (int r = 0; r < 3; r++) {
                volatile float t = R2[3 * r] * c3[0];
                volatile float t1 = R2[3 * r + 1] * c3[1];
                volatile float t2v = R2[3 * r + 2] * c3[2];
                t = t + t1; t = t + t2v;
                d3[r] = t + t2[r];
            }
            c3[0] = d3[0]; c3[1] = d3[1]; c3[2] = d3[2];
        }
        if (c3[2] < 0.0f) continue;                         /* :853 */
        const float invz = 1 / c3[2];
        const float x = c3[0] * invz, y = c3[1] * invz;
        volatile float u = fx * x; u = u + cx;
        volatile float v = fy * y; v = v + cy;
        if (!(u >= q.min_x && u < (float)maxXi && v >= q.min_y && v < (float)maxYi)) continue;     /* IsInImage */
        volatile float ur = bf * invz; ur = u - ur;
        double ss = 0.0, dot = 0.0;
        for (int r = 0; r < 3; r++) {
            const float po = mode == 2 ? c3[r] : X[r] - Ow[r];
            ss += (double)po * (double)po;
            if (mode != 2) dot += (double)po * (double)normal[3 * (size_t)i + r];
        }
        const float dist3D = (float)sqrt(ss);
        if (dist3D < 0.8f * mf_min_distance[i] || dist3D > 1.2f * mf_max_distance[i]) continue;
        if (mode != 2 && dot < 0.5 * (double)dist3D) continue;           /* viewing angle (:880) */
        int level = orc_predict_scale(mf_max_distance[i], dist3D, log_scale_factor);
        if (level < 0) level = 0;
        if (level >= nlevels) level = nlevels - 1;
        const float radius = th * scale_factors[level];
        const int nc = orc_features_in_area(&q, u, v, radius, -1, -1, cand, n);
        int bestDist = 256, bestIdx = -1;
        for (int c = 0; c < nc; c++) {
            const int idx = cand[c];
            const int kpLevel = koct[idx];
            if (kpLevel < level - 1 || kpLevel > level) continue;
            const float ex = u - kx[idx], ey = v - ky[idx];
            volatile float e2 = ex * ex, e2b = ey * ey;
            e2 = e2 + e2b;
            if (mode != 0) {
                /* no reprojection-error gate */
            } else if (kuright[idx] >= 0) {
                const float er = ur - kuright[idx];
                volatile float e2c = er * er;
                e2 = e2 + e2c;
                const float chi = e2 * inv_level_sigma2[kpLevel];
                if ((double)chi > 7.8) continue;
            } else {
                const float chi = e2 * inv_level_sigma2[kpLevel];
                if ((double)chi > 5.99) continue;
            }
            const int dist = orc_descriptor_distance(mp_desc + 32 * (size_t)i, kdesc + 32 * (size_t)idx);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        if (best_dist) best_dist[i] = bestDist;
        if (bestDist <= accept) best_idx[i] = bestIdx;
    }
    free(cand); free(items);
}


/* ORBmatcher::SearchBySim3 (S/ORBmatcher.cc:1106-1330) from its two legs: vnMatch1 (map points of key frame 1 searched
 * in key frame 2) and vnMatch2, then the agreement test (:1316-1328).  matches12[i1] out = slot of key frame 2 or -1
 * (the newly found ones only; slots excluded by already1 keep -1).  Returns nFound. */
int orc_sim3_agreement(int n1, const int32_t *match1, int n2, const int32_t *match2, int32_t *matches12)
{
    int found = 0;
    for (int i1 = 0; i1 < n1; i1++) {
        matches12[i1] = -1;
        const int idx2 = match1[i1];
        if (idx2 >= 0 && idx2 < n2 && match2[idx2] == i1) { matches12[i1] = idx2; found++; }
    }
    return found;
}


/* ------------------------------------------------------------------------------------------------
 * ORBmatcher::SearchByProjection(KeyFrame *pKF, cv::Mat Scw, const vector<MapPoint*> &vpPoints,
 * vector<MapPoint*> &vpMatched, int th) (S/ORBmatcher.cc:294-407, loop closing): the Fuse(Scw)-style projection and
 * gates, but keypoints that already hold a match are skipped and an accepted match occupies its keypoint (greedy,
 * in list order).  valid[i] = !isBad && not already in vpMatched.  matched[n] in/out: -1 free, >= 0 index into the
 * candidate list, any other value = occupied by a map point outside the list.  Returns nmatches. */
int orc_search_by_projection_sim3(
    int nmp, const uint8_t *valid, const float *wpos, const float *normal, const uint8_t *mp_desc,
    const float *mf_max_distance, const float *mf_min_distance,
    const float Rcw[9], const float tcw[3], const float Ow[3], const float K[4],
    int n, const float *kx, const float *ky, const int32_t *koct, const uint8_t *kdesc,
    int nlevels, const float *scale_factors, float log_scale_factor, const float bounds[4], int th, int32_t *matched)
{
    orc_grid g, q;
    int nmatches = 0;
    int32_t *items = (int32_t *)malloc(sizeof(int32_t) * (n + 1));
    int32_t *cand = (int32_t *)malloc(sizeof(int32_t) * (n + 1));
    orc_grid_bounds(&g, bounds);
    orc_grid_assign(&g, n, kx, ky, koct, items);
    q = g;
    q.min_x = (float)(int)bounds[0]; q.min_y = (float)(int)bounds[1];
    const int maxXi = (int)bounds[2], maxYi = (int)bounds[3];
    const float fx = K[0], fy = K[1], cx = K[2], cy = K[3];
    for (int i = 0; i < nmp; i++) {
        if (!valid[i]) continue;
        const float *X = wpos + 3 * (size_t)i;
        float c3[3];
        for (int r = 0; r < 3; r++) {
            volatile float t = Rcw[3 * r] * X[0];
            volatile float t1 = Rcw[3 * r + 1] * X[1];
            volatile float t2 = Rcw[3 * r + 2] * X[2];
            t = t + t1; t = t + t2;
            c3[r] = t + tcw[r];
        }
        if (c3[2] < 0.0) continue;
        const float invz = 1 / c3[2];
        const float x = c3[0] * invz, y = c3[1] * invz;
        volatile float u = fx * x; u = u + cx;
        volatile float v = fy * y; v = v + cy;
        if (!(u >= q.min_x && u < (float)maxXi && v >= q.min_y && v < (float)maxYi)) continue;
        double ss = 0.0, dot = 0.0;
        for (int r = 0; r < 3; r++) {
            const float po = X[r] - Ow[r];
            ss += (double)po * (double)po;
            dot += (double)po * (double)normal[3 * (size_t)i + r];
        }
        const float dist = (float)sqrt(ss);
        if (dist < 0.8f * mf_min_distance[i] || dist > 1.2f * mf_max_distance[i]) continue;
        if (dot < 0.5 * (double)dist) continue;
        int level = orc_predict_scale(mf_max_distance[i], dist, log_scale_factor);
        if (level < 0) level = 0;
        if (level >= nlevels) level = nlevels - 1;
        const float radius = th * scale_factors[level];
        const int nc = orc_features_in_area(&q, u, v, radius, -1, -1, cand, n);
        int bestDist = 256, bestIdx = -1;
        for (int c = 0; c < nc; c++) {
            const int idx = cand[c];
            if (matched[idx] != -1) continue;
            if (koct[idx] < level - 1 || koct[idx] > level) continue;
            const int d = orc_descriptor_distance(mp_desc + 32 * (size_t)i, kdesc + 32 * (size_t)idx);
            if (d < bestDist) { bestDist = d; bestIdx = idx; }
        }
        if (bestDist <= TH_LOW) { matched[bestIdx] = i; nmatches++; }
    }
    free(cand); free(items);
    return nmatches;
}


/* ------------------------------------------------------------------------------------------------
 * DBoW2 TemplatedVocabulary<cv::Mat, FORB>::transform(features, BowVector, FeatureVector, levelsup) as called by
 * Frame::ComputeBoW (S/Frame.cc:520-527, levelsup = 4), for TF_IDF weighting and L1 scoring (what ORBvoc.txt
 * declares): Thirdparty/DBoW2/include/DBoW2/TemplatedVocabulary.h:1133-1266, src/BowVector.cpp:34-84.
 * Vocabulary tree flattened: node 0 = root; node i's children are children[child_start[i] .. child_start[i+1])
 * in the order of its `children` vector; leaves have no children, a word id and a weight.
 * Per feature the tree is descended (child with the smallest Hamming distance, first one on ties) down to a leaf;
 * the node passed at level L - levelsup is the feature's FeatureVector node (0 = root if that level is <= 0).
 * A word with weight w > 0 seen c times gets w added c times (addWeight), then the vector is L1-normalised with the
 * norm accumulated in ascending word order.  Outputs: bow_word / bow_value (ascending word id, returns their count in
 * *bow_n), fv_node (ascending) / fv_start / fv_feat (feature indices ascending inside a node), *fv_n nodes. */
void orc_bow_transform(
    int n_nodes, int L, const int32_t *child_start, const int32_t *children, const uint8_t *node_desc,
    const int32_t *word_id, const double *weight,
    int n, const uint8_t *desc, int levelsup,
    int32_t *bow_n, uint32_t *bow_word, double *bow_value,
    int32_t *fv_n, uint32_t *fv_node, int32_t *fv_start, uint32_t *fv_feat)
{
    (void)n_nodes;
    const int nid_level = L - levelsup;
    int32_t *w_of = (int32_t *)malloc(sizeof(int32_t) * (n + 1));      /* word per feature, -1 = stopped */
    int32_t *n_of = (int32_t *)malloc(sizeof(int32_t) * (n + 1));
    double *wt_of = (double *)malloc(sizeof(double) * (n + 1));
    for (int f = 0; f < n; f++) {
        int node = 0, level = 0, nid = 0;
        do {
            ++level;
            const int cs = child_start[node], ce = child_start[node + 1];
            int best = children[cs];
            int best_d = orc_descriptor_distance(desc + 32 * (size_t)f, node_desc + 32 * (size_t)best);
            for (int c = cs + 1; c < ce; c++) {
                const int id = children[c];
                const int d = orc_descriptor_distance(desc + 32 * (size_t)f, node_desc + 32 * (size_t)id);
                if (d < best_d) { best_d = d; best = id; }
            }
            node = best;
            if (level == nid_level) nid = node;
        } while (child_start[node + 1] > child_start[node]);
        const double w = weight[node];
        w_of[f] = w > 0 ? word_id[node] : -1;
        wt_of[f] = w;
        n_of[f] = nid;
    }
    /* BowVector: std::map<WordId, double> with addWeight in feature order; FeatureVector: std::map<NodeId, vector> */
    int nb = 0, nf = 0;
    for (int f = 0; f < n; f++) {
        if (w_of[f] < 0) continue;
        int p = 0;
        while (p < nb && bow_word[p] < (uint32_t)w_of[f]) p++;
        if (p < nb && bow_word[p] == (uint32_t)w_of[f]) bow_value[p] += wt_of[f];
        else {
            memmove(bow_word + p + 1, bow_word + p, sizeof(uint32_t) * (nb - p));
            memmove(bow_value + p + 1, bow_value + p, sizeof(double) * (nb - p));
            bow_word[p] = (uint32_t)w_of[f]; bow_value[p] = wt_of[f]; nb++;
        }
    }
    /* nodes: collect distinct ascending, then fill lists in feature order */
    for (int f = 0; f < n; f++) {
        if (w_of[f] < 0) continue;
        int p = 0;
        while (p < nf && fv_node[p] < (uint32_t)n_of[f]) p++;
        if (!(p < nf && fv_node[p] == (uint32_t)n_of[f])) {
            memmove(fv_node + p + 1, fv_node + p, sizeof(uint32_t) * (nf - p));
            fv_node[p] = (uint32_t)n_of[f]; nf++;
        }
    }
    int pos = 0;
    for (int k = 0; k < nf; k++) {
        fv_start[k] = pos;
        for (int f = 0; f < n; f++) if (w_of[f] >= 0 && (uint32_t)n_of[f] == fv_node[k]) fv_feat[pos++] = (uint32_t)f;
    }
    fv_start[nf] = pos;
    double norm = 0.0;                                            /* BowVector::normalize(L1) */
    for (int p = 0; p < nb; p++) norm += fabs(bow_value[p]);
    if (norm > 0.0) for (int p = 0; p < nb; p++) bow_value[p] /= norm;
    *bow_n = nb; *fv_n = nf;
    free(wt_of); free(n_of); free(w_of);
}

/* ---- Frame::ComputeStereoMatches (S/Frame.cc:591-763) ------------------------------------------------------------
 * Left keypoints against the right keypoints of the same rows (row table of the right image, :598-617), best
 * descriptor (:631-677), then an 11 x 11 sum of absolute differences of centre-subtracted patches slid over +-5
 * columns of the right pyramid level (:680-715), a parabola through the three costs around the minimum (:720-729)
 * and the depth (:731-745); finally the matches whose patch cost reaches 1.5 * 1.4 * median are withdrawn (:749-762).
 * The patches hold small integers in CV_32F, so cv::norm(NORM_L1)'s double sum is exact in any order.
 * level images: pointer / pitch / width / height per pyramid level, WITHOUT border.  Where the reference would leave
 * an image (cv::Mat::rowRange / colRange throw, vRowIndices is indexed out of range) the keypoint is skipped and
 * counted in *skipped.  Returns the number of left keypoints that end with a depth. */
static float orc_roundf_away(float v) { return roundf(v); }

int orc_compute_stereo_matches(
    int n, const float *kx, const float *ky, const int32_t *koct, const uint8_t *desc,
    int nr, const float *rx, const float *ry, const int32_t *roct, const uint8_t *rdesc,
    int nlevels, const float *scale, const float *inv_scale,
    const uint8_t *const *limg, const int32_t *lpitch, const uint8_t *const *rimg, const int32_t *rpitch,
    const int32_t *lw, const int32_t *lh, float mb, float mbf, float *u_right, float *depth, int32_t *skipped)
{
    const int TH_HIGH_ = 100;
    int nskip = 0;
    for (int i = 0; i < n; i++) { u_right[i] = -1.0f; depth[i] = -1.0f; }
    const int nRows = lh[0];                                                      /* :596 */
    /* row table (:598-617): CSR in push_back order */
    int *rowCount = (int *)calloc((size_t)nRows + 1, sizeof(int));
    for (int iR = 0; iR < nr; iR++) {
        const float r = 2.0f * scale[roct[iR]];
        const int maxr = (int)ceilf(ry[iR] + r), minr = (int)floorf(ry[iR] - r);
        for (int yi = minr; yi <= maxr; yi++) { if (yi >= 0 && yi < nRows) rowCount[yi + 1]++; else nskip++; }
    }
    for (int y = 0; y < nRows; y++) rowCount[y + 1] += rowCount[y];
    int *rowItems = (int *)malloc(sizeof(int) * (size_t)(rowCount[nRows] + 1));
    int *fill = (int *)malloc(sizeof(int) * (size_t)(nRows + 1));
    memcpy(fill, rowCount, sizeof(int) * (size_t)(nRows + 1));
    for (int iR = 0; iR < nr; iR++) {
        const float r = 2.0f * scale[roct[iR]];
        const int maxr = (int)ceilf(ry[iR] + r), minr = (int)floorf(ry[iR] - r);
        for (int yi = minr; yi <= maxr; yi++) if (yi >= 0 && yi < nRows) rowItems[fill[yi]++] = iR;
    }
    const float minZ = mb, minD = -3, maxD = mbf / minZ;                          /* :620-622 */
    int *vDist = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    int *vIdx = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    int nd = 0;
    for (int iL = 0; iL < n; iL++) {
        const int levelL = koct[iL];
        const float vL = ky[iL], uL = kx[iL];
        if (!(vL >= 0.0f) || !(vL < (float)nRows)) { nskip++; continue; }            /* vRowIndices[vL] out of range */
        const int row = (int)vL;
        const int cs = rowCount[row], ce = rowCount[row + 1];
        if (cs == ce) continue;                                                   /* :637-638 */
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;                                                   /* :643-644 */
        int bestDist = TH_HIGH_, bestIdxR = 0;
        for (int c = cs; c < ce; c++) {                                           /* :651-672 */
            const int iR = rowItems[c];
            if (roct[iR] < levelL - 1 || roct[iR] > levelL + 1) continue;
            const float uR = rx[iR];
            if (uR >= minU && uR <= maxU) {
                const int dist = orc_descriptor_distance(desc + 32 * (size_t)iL, rdesc + 32 * (size_t)iR);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (!(bestDist < TH_HIGH_)) continue;                                     /* :675 */
        const float uR0 = rx[bestIdxR];
        const float scaleFactor = inv_scale[levelL];
        const float scaleduL = orc_roundf_away(kx[iL] * scaleFactor);
        const float scaledvL = orc_roundf_away(ky[iL] * scaleFactor);
        const float scaleduR0 = orc_roundf_away(uR0 * scaleFactor);
        const int w = 5, L = 5;
        const int W = lw[levelL], H = lh[levelL];
        /* IL (:684): rows scaledvL-w .. scaledvL+w, columns scaleduL-w .. scaleduL+w of the left level */
        if (!(scaledvL - w >= 0 && scaledvL + w + 1 <= H && scaleduL - w >= 0 && scaleduL + w + 1 <= W)) { nskip++; continue; }
        const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;       /* :693-696 */
        if (iniu < 0 || endu >= W) continue;
        if (!(scaleduR0 - L - w >= 0)) { nskip++; continue; }                    /* colRange would throw (:700) */
        const int v0 = (int)scaledvL, u0 = (int)scaleduL, r0 = (int)scaleduR0;
        const uint8_t *Lp = limg[levelL], *Rp = rimg[levelL];
        const int lp = lpitch[levelL], rp = rpitch[levelL];
        const int cL = Lp[(size_t)v0 * lp + u0];
        int best = 2147483647, bestincR = 0;
        float vDists[11];
        for (int incR = -L; incR <= L; incR++) {                                  /* :698-713 */
            const int cR = Rp[(size_t)v0 * rp + r0 + incR];
            int sad = 0;
            for (int dy = -w; dy <= w; dy++)
                for (int dx = -w; dx <= w; dx++) {
                    const int a = Lp[(size_t)(v0 + dy) * lp + u0 + dx] - cL;
                    const int b = Rp[(size_t)(v0 + dy) * rp + r0 + incR + dx] - cR;
                    sad += a > b ? a - b : b - a;
                }
            const float dist = (float)sad;
            if (dist < (float)best) { best = (int)dist; bestincR = incR; }
            vDists[L + incR] = dist;
        }
        if (bestincR == -L || bestincR == L) continue;                            /* :715-716 */
        const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
        const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
        if (deltaR < -1 || deltaR > 1) continue;                                  /* :725-726 */
        float bestuR = scale[levelL] * ((float)scaleduR0 + (float)bestincR + deltaR);
        float disparity = uL - bestuR;
        if (disparity >= 0 && disparity < maxD) {                                 /* :733-744 */
            if (disparity <= 0) { disparity = 0.01; bestuR = uL - 0.01; }
            depth[iL] = mbf / disparity;
            u_right[iL] = bestuR;
            vDist[nd] = best; vIdx[nd] = iL; nd++;
        }
    }
    int kept = nd;
    if (nd > 0) {
        /* sort(vDistIdx) (:749): only the median and "distance >= threshold" matter, both independent of tie order */
        int *sorted = (int *)malloc(sizeof(int) * (size_t)nd);
        memcpy(sorted, vDist, sizeof(int) * (size_t)nd);
        for (int a = 1; a < nd; a++) { const int v = sorted[a]; int b = a - 1; while (b >= 0 && sorted[b] > v) { sorted[b + 1] = sorted[b]; b--; } sorted[b + 1] = v; }
        const float median = (float)sorted[nd / 2];
        const float thDist = 1.5f * 1.4f * median;
        for (int k = 0; k < nd; k++)
            if (!((float)vDist[k] < thDist)) { u_right[vIdx[k]] = -1; depth[vIdx[k]] = -1; kept--; }
        free(sorted);
    }
    if (skipped) *skipped = nskip;
    free(vIdx); free(vDist); free(fill); free(rowItems); free(rowCount);
    return kept;
}
