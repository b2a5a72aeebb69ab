// orb_frame.cu -- device-resident Frame glue: what sits between ORBextractor's output and
// every ORBmatcher call in the reference's Frame constructor (S/Frame.cc:192-334):
//   Frame::UndistortKeyPoints  S/Frame.cc:529-559  (cv::undistortPoints(mat, mat, mK, mDistCoef, Mat(), mK))
//   Frame::ComputeImageBounds  S/Frame.cc:561-589
// plus the array-of-cv::KeyPoint -> structure-of-arrays conversion the matcher kernels want, so that
// extractor -> frame -> matcher runs without leaving the device (AssignFeaturesToGrid is k_build_grid in
// orb_matcher.cu).  The arithmetic is OpenCV's, in double, every operation separately rounded.
#include <cstdint>
#include "common.cuh"

namespace orbb200 {

struct CamModel { double fx, fy, cx, cy, ifx, ify, k1, k2, p1, p2, k3; };

// cv::undistortPoints for one point, P = K, R = I, 5 fixed iterations (TermCriteria(COUNT, 5)).
__device__ __forceinline__ void undistort_point(const CamModel& c, float u_, float v_, float& ox, float& oy)
{
    const double u = (double)u_, v = (double)v_;
    double x = __dmul_rn(__dsub_rn(u, c.cx), c.ifx), y = __dmul_rn(__dsub_rn(v, c.cy), c.ify);
    const double x0 = x, y0 = y;
#pragma unroll 1
    for (int j = 0; j < 5; j++) {
        const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
        // numerator (1 + ((k6*r2 + k5)*r2 + k4)*r2) is exactly 1: the rational terms are not used by the reference
        const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(c.k3, r2), c.k2), r2), c.k1), r2));
        const double icdist = __ddiv_rn(1.0, den);
        if (icdist < 0) { x = __dmul_rn(__dsub_rn(u, c.cx), c.ifx); y = __dmul_rn(__dsub_rn(v, c.cy), c.ify); break; }
        const double dX = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, c.p1), x), y),
                                    __dmul_rn(c.p2, __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x))));
        const double dY = __dadd_rn(__dmul_rn(c.p1, __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y))),
                                    __dmul_rn(__dmul_rn(__dmul_rn(2.0, c.p2), x), y));
        x = __dmul_rn(__dsub_rn(x0, dX), icdist);
        y = __dmul_rn(__dsub_rn(y0, dY), icdist);
    }
    // re-projection with P = K, R = identity, written out like OpenCV's 3x3 product
    const double xx = __dadd_rn(__dadd_rn(__dmul_rn(c.fx, x), __dmul_rn(0.0, y)), c.cx);
    const double yy = __dadd_rn(__dadd_rn(__dmul_rn(0.0, x), __dmul_rn(c.fy, y)), c.cy);
    const double ww = __ddiv_rn(1.0, __dadd_rn(__dadd_rn(__dmul_rn(0.0, x), __dmul_rn(0.0, y)), 1.0));
    ox = (float)__dmul_rn(xx, ww);
    oy = (float)__dmul_rn(yy, ww);
}

// keypoints: items x cap cv::KeyPoint records (the extractor's output) -> SoA, undistorted when dist[0] != 0
__global__ void __launch_bounds__(256) k_frame_prepare(const orbb200_keypoint* kps, const int* counts, int cap, CamModel cam,
                                                       int undist, float* x, float* y, int* octave, float* angle)
{
    const int item = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
    if (i >= min(counts[item], cap)) return;
    const size_t o = (size_t)item * cap + i;
    const orbb200_keypoint k = kps[o];
    float ux = k.x, uy = k.y;
    if (undist) undistort_point(cam, k.x, k.y, ux, uy);                      // S/Frame.cc:531-535: copy when k1 == 0
    x[o] = ux; y[o] = uy; octave[o] = k.octave; angle[o] = k.angle;
}

__global__ void k_undistort_xy(const float* in, float* out, int n, CamModel cam)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    undistort_point(cam, in[2 * i], in[2 * i + 1], out[2 * i], out[2 * i + 1]);
}

static CamModel cam_model(const float* K, const float* dist)
{
    CamModel c;
    c.fx = K[0]; c.fy = K[1]; c.cx = K[2]; c.cy = K[3];
    c.ifx = 1. / c.fx; c.ify = 1. / c.fy;
    c.k1 = dist[0]; c.k2 = dist[1]; c.p1 = dist[2]; c.p2 = dist[3]; c.k3 = dist[4];
    return c;
}

}  // namespace orbb200

using namespace orbb200;

extern "C" int orbb200_frames_from_keypoints(orbb200_matcher* m, const orbb200_keypoint* d_keypoints, const int32_t* d_counts,
                                             int items, int cap, const float* K, const float* dist, float* d_x, float* d_y,
                                             int32_t* d_octave, float* d_angle)
{
    if (!m || !d_keypoints || !d_counts || !K || !dist || !d_x || !d_y || !d_octave || !d_angle || items < 1 || cap < 1) {
        set_error("orbb200_frames_from_keypoints: bad argument");
        return ORBB200_EINVAL;
    }
    ORB_CUDA(cudaSetDevice(orbb200_matcher_device(m)));
    cudaStream_t st = (cudaStream_t)orbb200_matcher_stream(m);
    k_frame_prepare<<<dim3((cap + 255) / 256, items), 256, 0, st>>>(d_keypoints, d_counts, cap, cam_model(K, dist),
                                                                    dist[0] != 0.0f ? 1 : 0, d_x, d_y, d_octave, d_angle);
    ORB_CHECK_LAUNCH("k_frame_prepare");
    return ORBB200_OK;
}

extern "C" int orbb200_undistort_points(orbb200_matcher* m, const float* xy_in, float* xy_out, int n, const float* K, const float* dist)
{
    if (!m || !xy_in || !xy_out || !K || !dist || n < 0) { set_error("orbb200_undistort_points: bad argument"); return ORBB200_EINVAL; }
    if (n == 0) return ORBB200_OK;
    ORB_CUDA(cudaSetDevice(orbb200_matcher_device(m)));
    cudaStream_t st = (cudaStream_t)orbb200_matcher_stream(m);
    float *din = nullptr, *dout = nullptr;
    ORB_CUDA(cudaMallocAsync((void**)&din, sizeof(float) * 2 * n, st));
    ORB_CUDA(cudaMallocAsync((void**)&dout, sizeof(float) * 2 * n, st));
    ORB_CUDA(cudaMemcpyAsync(din, xy_in, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, st));
    k_undistort_xy<<<(n + 255) / 256, 256, 0, st>>>(din, dout, n, cam_model(K, dist));
    ORB_CHECK_LAUNCH("k_undistort_xy");
    ORB_CUDA(cudaMemcpyAsync(xy_out, dout, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaFreeAsync(din, st));
    ORB_CUDA(cudaFreeAsync(dout, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    return ORBB200_OK;
}

extern "C" int orbb200_image_bounds(orbb200_matcher* m, int cols, int rows, const float* K, const float* dist, float* bounds)
{
    if (!m || !K || !dist || !bounds || cols < 1 || rows < 1) { set_error("orbb200_image_bounds: bad argument"); return ORBB200_EINVAL; }
    if (dist[0] != 0.0f) {                                                         // S/Frame.cc:563
        const float c[8] = {0.f, 0.f, (float)cols, 0.f, 0.f, (float)rows, (float)cols, (float)rows};
        float o[8];
        int rc = orbb200_undistort_points(m, c, o, 4, K, dist);
        if (rc != ORBB200_OK) return rc;
        bounds[0] = o[0] < o[4] ? o[0] : o[4];
        bounds[2] = o[2] > o[6] ? o[2] : o[6];
        bounds[1] = o[1] < o[3] ? o[1] : o[3];
        bounds[3] = o[5] > o[7] ? o[5] : o[7];
    } else {
        bounds[0] = 0.0f; bounds[2] = (float)cols; bounds[1] = 0.0f; bounds[3] = (float)rows;
    }
    return ORBB200_OK;
}

extern "C" int orbb200_matcher_wait_extractor(orbb200_matcher* m, orbb200_extractor* ex)
{
    if (!m || !ex) { set_error("orbb200_matcher_wait_extractor: null handle"); return ORBB200_EINVAL; }
    const int exDevice = orbb200_extractor_device(ex);
    if (exDevice != orbb200_matcher_device(m)) { set_error("orbb200_matcher_wait_extractor: extractor on device %d, matcher on device %d", exDevice, orbb200_matcher_device(m)); return ORBB200_EINVAL; }
    ORB_CUDA(cudaSetDevice(orbb200_matcher_device(m)));
    cudaEvent_t ev;
    ORB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    ORB_CUDA(cudaEventRecord(ev, (cudaStream_t)orbb200_extractor_stream(ex)));
    ORB_CUDA(cudaStreamWaitEvent((cudaStream_t)orbb200_matcher_stream(m), ev, 0));
    ORB_CUDA(cudaEventDestroy(ev));                                                // released once the wait has been satisfied
    return ORBB200_OK;
}
