// Adam on a flat fp32 parameter buffer: one launch per network instead of the ~10
// element-wise launches per parameter tensor of torch.optim.Adam's default path (48 tensors for the two nets).
// Replaces optimizer.step() of run_nerf.py:440 / :1774 (torch.optim.Adam, betas (0.9, 0.999), eps 1e-8, no weight
// decay, no amsgrad); the learning-rate decay of :1843-1847 stays with the caller (it only rewrites `lr`).
#include "common.cuh"
#include "../../include/dlnerf_b200.h"
#include <math.h>

namespace {

// torch.optim.Adam, single-tensor path, per element:
//   m += (g - m) (1 - b1);  v = v b2 + (1 - b2) g g;  p -= (lr / bc1) m / (sqrt(v) / sqrt(bc2) + eps)
__global__ void __launch_bounds__(256)
    adam_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                long long n4, float* __restrict__ p_tail, const float* __restrict__ g_tail, float* __restrict__ m_tail,
                float* __restrict__ v_tail, int n_tail, float step_size, float one_minus_b1, float b2,
                float one_minus_b2, float inv_sqrt_bc2, float eps, float grad_scale) {
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg *= grad_scale;
    mm = mm + (gg - mm) * one_minus_b1;
    vv = vv * b2 + one_minus_b2 * gg * gg;
    pp = pp - step_size * (mm / (sqrtf(vv) * inv_sqrt_bc2 + eps));
  };
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = p[i], mm = m[i], vv = v[i];
    const float4 gg = __ldg(g + i);
    upd(pp.x, gg.x, mm.x, vv.x), upd(pp.y, gg.y, mm.y, vv.y), upd(pp.z, gg.z, mm.z, vv.z), upd(pp.w, gg.w, mm.w, vv.w);
    p[i] = pp, m[i] = mm, v[i] = vv;
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < n_tail) {
    const int i = threadIdx.x;
    upd(p_tail[i], g_tail[i], m_tail[i], v_tail[i]);
  }
}

}  // namespace

extern "C" int dln_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                             double lr, double beta1, double beta2, double eps, int step, float grad_scale,
                             void* stream) {
  // hyper-parameters arrive as doubles (Python floats) and are combined in double before the single rounding to
  // fp32, as torch does with its Python scalars (1 - 0.999 must be 1e-3, not 1 - 0.999f)
  DLN_CHECK_ARG(n >= 0 && step >= 1 && beta1 >= 0. && beta1 < 1. && beta2 >= 0. && beta2 < 1.);
  if (n == 0) return DLN_OK;
  DLN_CHECK_ARG(params && grads && exp_avg && exp_avg_sq);
  const uintptr_t al = reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) |
                       reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq);
  DLN_CHECK_ARG((al & 15) == 0);
  const double bc1 = 1.0 - pow(beta1, step), bc2 = 1.0 - pow(beta2, step);
  const long long n4 = n / 4;
  const int n_tail = (int)(n - 4 * n4);
  long long blocks = (n4 + 255) / 256;
  blocks = blocks < 1 ? 1 : (blocks > 148 * 8 ? 148 * 8 : blocks);
  adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<float4*>(params), reinterpret_cast<const float4*>(grads), reinterpret_cast<float4*>(exp_avg),
      reinterpret_cast<float4*>(exp_avg_sq), n4, params + 4 * n4, grads + 4 * n4, exp_avg + 4 * n4,
      exp_avg_sq + 4 * n4, n_tail, (float)(lr / bc1), (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
      (float)(1.0 / sqrt(bc2)), (float)eps, grad_scale);
  return dln_launch_status();
}
