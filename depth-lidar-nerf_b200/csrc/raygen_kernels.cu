// Ray generation on the device: the pinhole-camera rays the reference builds on the
// host with numpy / torch before every training run (whole-image banks, run_nerf.py:1126-1187) and before every
// patch iteration (run_nerf.py:1566).
//
//   gen_rays_grid      get_rays_np            run_nerf_helpers.py:285-300   all pixels of n_poses cameras
//   gen_rays_coord<T>  get_rays_by_coord_np   run_nerf_helpers.py:303-318   fractional (LiDAR / COLMAP) pixel coordinates
//   gen_rays_patch     get_rays_cropped_feature_loss_new  :430-494          a permuted nH x nW crop + its pixel indices
//
// Arithmetic follows the reference operation by operation in the reference's precision (no contraction into FMAs:
// every step is an explicit round-to-nearest op), so the results are bit-identical to numpy / CPU torch:
//   dirs = ((i - W*.5) / focal, -(j - H*.5) / focal, -1),  rays_d[k] = (dirs[0]*R[k][0] + dirs[1]*R[k][1]) + dirs[2]*R[k][2]
// All three are pure streaming kernels (24 B written per ray, nothing but the 48-byte pose read): HBM-write bound.
#include "common.cuh"
#include "../../include/dlnerf_b200.h"

namespace {

template <typename T> struct Rn;
template <> struct Rn<float> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <> struct Rn<double> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};

// one ray through pixel (x, y) of the camera c2w[3][4] (row-major): o = c2w[:, 3], d = R dirs
template <typename T>
__device__ __forceinline__ void pixel_ray(const T* __restrict__ c2w, T x, T y, T half_w, T half_h, T focal,
                                          T* __restrict__ o, T* __restrict__ d) {
  using R = Rn<T>;
  const T d0 = R::div(R::sub(x, half_w), focal);
  const T d1 = -R::div(R::sub(y, half_h), focal);
  const T d2 = (T)-1;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    d[k] = R::add(R::add(R::mul(d0, c2w[4 * k]), R::mul(d1, c2w[4 * k + 1])), R::mul(d2, c2w[4 * k + 2]));
    o[k] = c2w[4 * k + 3];
  }
}

__global__ void gen_rays_grid_kernel(const float* __restrict__ c2w, long long n_rays, int H, int W, float focal,
                                     float* __restrict__ rays_o, float* __restrict__ rays_d, long long stride) {
  const long long hw = (long long)H * W;
  const float half_w = (float)(W * .5), half_h = (float)(H * .5);
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n_rays; t += (long long)gridDim.x * blockDim.x) {
    const long long pose = t / hw;
    const int pix = (int)(t - pose * hw);
    const int y = pix / W, x = pix - y * W;
    float o[3], d[3];
    pixel_ray<float>(c2w + pose * 12, (float)x, (float)y, half_w, half_h, focal, o, d);
#pragma unroll
    for (int k = 0; k < 3; ++k) rays_o[t * stride + k] = o[k], rays_d[t * stride + k] = d[k];
  }
}

template <typename T>
__global__ void gen_rays_coord_kernel(const T* __restrict__ c2w, const T* __restrict__ coords, long long n, int H, int W,
                                      T focal, T* __restrict__ rays_o, T* __restrict__ rays_d, long long stride) {
  const T half_w = (T)(W * 0.5), half_h = (T)(H * 0.5);
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    T o[3], d[3];
    pixel_ray<T>(c2w, coords[2 * t], coords[2 * t + 1], half_w, half_h, focal, o, d);
#pragma unroll
    for (int k = 0; k < 3; ++k) rays_o[t * stride + k] = o[k], rays_d[t * stride + k] = d[k];
  }
}

// element t of the permuted crop: flat crop index perm[t] = row * nW + col (row-major over the nH x nW crop, the order
// of `dirs.reshape(-1, 3)` at run_nerf_helpers.py:459), pixel (start_w + col, start_h + row); points = (row, col)
__global__ void gen_rays_patch_kernel(const float* __restrict__ c2w, int H, int W, float focal, int start_w, int start_h,
                                      int nW, const long long* __restrict__ perm, int n, float* __restrict__ rays_o,
                                      float* __restrict__ rays_d, long long* __restrict__ points) {
  const float half_w = (float)(W * .5), half_h = (float)(H * .5);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const long long idx = perm[t];
    const int row = (int)(idx / nW), col = (int)(idx - (long long)row * nW);
    float o[3], d[3];
    pixel_ray<float>(c2w, (float)(start_w + col), (float)(start_h + row), half_w, half_h, focal, o, d);
#pragma unroll
    for (int k = 0; k < 3; ++k) rays_o[3 * (size_t)t + k] = o[k], rays_d[3 * (size_t)t + k] = d[k];
    points[2 * (size_t)t] = row, points[2 * (size_t)t + 1] = col;
  }
}

int grid_for(long long n, int threads) {
  const long long want = (n + threads - 1) / threads;
  const long long cap = (long long)dln_sm_count() * 8;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace

extern "C" {

int dln_gen_rays(const float* c2w, int n_poses, int H, int W, double focal, float* rays_o, float* rays_d,
                 long long out_stride, void* stream) {
  DLN_CHECK_ARG(c2w && rays_o && rays_d && n_poses > 0 && H > 0 && W > 0 && focal != 0.0 && out_stride >= 3);
  const long long n = (long long)n_poses * H * W;
  gen_rays_grid_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(c2w, n, H, W, (float)focal, rays_o, rays_d,
                                                                            out_stride);
  return dln_launch_status();
}

int dln_gen_rays_by_coord(const void* c2w, const void* coords, long long N, int H, int W, double focal, int is_f64,
                          void* rays_o, void* rays_d, long long out_stride, void* stream) {
  DLN_CHECK_ARG(c2w && rays_o && rays_d && N >= 0 && (N == 0 || coords) && H > 0 && W > 0 && focal != 0.0 && out_stride >= 3);
  if (N == 0) return DLN_OK;
  if (is_f64)
    gen_rays_coord_kernel<double><<<grid_for(N, 256), 256, 0, (cudaStream_t)stream>>>(
        (const double*)c2w, (const double*)coords, N, H, W, focal, (double*)rays_o, (double*)rays_d, out_stride);
  else
    gen_rays_coord_kernel<float><<<grid_for(N, 256), 256, 0, (cudaStream_t)stream>>>(
        (const float*)c2w, (const float*)coords, N, H, W, (float)focal, (float*)rays_o, (float*)rays_d, out_stride);
  return dln_launch_status();
}

int dln_gen_rays_patch(const float* c2w, int H, int W, double focal, int start_w, int start_h, int nH, int nW,
                       const long long* perm, int n, float* rays_o, float* rays_d, long long* points, void* stream) {
  DLN_CHECK_ARG(c2w && perm && rays_o && rays_d && points && H > 0 && W > 0 && focal != 0.0);
  DLN_CHECK_ARG(nH > 0 && nW > 0 && n > 0 && n <= nH * nW && start_w >= 0 && start_h >= 0 && start_w + nW <= W &&
                start_h + nH <= H);
  gen_rays_patch_kernel<<<grid_for(n, 128), 128, 0, (cudaStream_t)stream>>>(c2w, H, W, (float)focal, start_w, start_h, nW,
                                                                            perm, n, rays_o, rays_d, points);
  return dln_launch_status();
}

}  // extern "C"
