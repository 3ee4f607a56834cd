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

// ------------------------------------------------------------------------------------------------
// feature_linear folded into views_linears (plan.py build_plan, fold_feature): W = 256, W/2 = 128.
//   fold   : M[i][j] = sum_k Wv[i][k] Wf[k][j],  b'[i] = sum_k Wv[i][k] bf[k] + bv[i]      (after every weight update)
//   unfold : dWv[i][k] += sum_j dM[i][j] Wf[k][j] + db'[i] bf[k];  dWf[k][j] += sum_i Wv[i][k] dM[i][j];
//            dbf[k] += sum_i Wv[i][k] db'[i];  dbv[i] += db'[i]                               (after every wgrad)
// 8 M multiply-adds each, once per step and network: CUDA cores, fp32.
// ------------------------------------------------------------------------------------------------
struct FoldOffsets {
  long long wv, wf, bf, bv, M, bM;
  int ldv;
};

__global__ void __launch_bounds__(256) fold_kernel(float* __restrict__ flat, FoldOffsets o) {
  __shared__ float wrow[256];
  __shared__ float red[8];
  const int i = blockIdx.x, j = threadIdx.x;
  wrow[j] = flat[o.wv + (long long)i * o.ldv + j];
  __syncthreads();
  const float* wf = flat + o.wf;
  float acc = 0.f;
#pragma unroll 8
  for (int k = 0; k < 256; ++k) acc = fmaf(wrow[k], wf[k * 256 + j], acc);
  flat[o.M + i * 256 + j] = acc;
  float b = dln::warp_sum(wrow[j] * flat[o.bf + j]);
  if ((j & 31) == 0) red[j >> 5] = b;
  __syncthreads();
  if (j == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    flat[o.bM + i] = t + flat[o.bv + i];
  }
}

// 32 x 32 output tiles, 256 threads, 2 x 2 outputs per thread, K in steps of 32 through shared memory (the first
// version gave every CTA a whole row and re-read all of W_f / dM from L2 per row: 64 MB of L2 reads, 23 us).
//   blocks [0, 32)    : dW_v1[i][k] += sum_j dM[i][j] W_f[k][j] + db'_i b_f[k]      (128 x 256, K = 256, "NT")
//   blocks [32, 96)   : dW_f[k][j]  += sum_i W_v1[i][k] dM[i][j]                    (256 x 256, K = 128, "TN")
//   block 96          : db_f[k] += sum_i W_v1[i][k] db'_i ;  db_v[i] += db'_i
__global__ void __launch_bounds__(256) unfold_kernel(const float* __restrict__ flat, float* __restrict__ g, FoldOffsets o) {
  __shared__ float As[32][33];     // [m][k]
  __shared__ float Bs[32][33];     // [n][k]
  const int b = blockIdx.x, t = threadIdx.x;
  if (b == 96) {
    float* sh = &As[0][0];
    if (t < 128) sh[t] = g[o.bM + t];
    __syncthreads();
    float acc = 0.f;
#pragma unroll 16
    for (int i = 0; i < 128; ++i) acc = fmaf(__ldg(flat + o.wv + (long long)i * o.ldv + t), sh[i], acc);
    g[o.bf + t] += acc;
    if (t < 128) g[o.bv + t] += sh[t];
    return;
  }
  const bool nt = b < 32;
  const int tile = nt ? b : b - 32;
  const int m0 = (tile >> 3) * 32, n0 = (tile & 7) * 32;          // 8 tiles across the 256 output columns
  const int K = nt ? 256 : 128;
  const int lr = t >> 5, lc = t & 31;                              // loader: row lr + 8 p, column lc
  const int ty = t >> 4, tx = t & 15;                              // outputs (2 ty + {0,1}, 2 tx + {0,1})
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int r = lr + 8 * p;
      if (nt) {
        As[r][lc] = g[o.M + (m0 + r) * 256 + k0 + lc];                              // dM[i][j]
        Bs[r][lc] = __ldg(flat + o.wf + (n0 + r) * 256 + k0 + lc);                   // W_f[k][j]
      } else {
        As[lc][r] = __ldg(flat + o.wv + (long long)(k0 + r) * o.ldv + m0 + lc);      // W_v1[i][k] -> [k][i]
        Bs[lc][r] = g[o.M + (k0 + r) * 256 + n0 + lc];                               // dM[i][j]   -> [j][i]
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float a0 = As[2 * ty][k], a1 = As[2 * ty + 1][k], b0 = Bs[2 * tx][k], b1 = Bs[2 * tx + 1][k];
      acc[0][0] = fmaf(a0, b0, acc[0][0]), acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]), acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int y = 0; y < 2; ++y)
#pragma unroll
    for (int x = 0; x < 2; ++x) {
      const int m = m0 + 2 * ty + y, n = n0 + 2 * tx + x;
      if (nt)
        g[o.wv + (long long)m * o.ldv + n] += acc[y][x] + g[o.bM + m] * __ldg(flat + o.bf + n);
      else
        g[o.wf + m * 256 + n] += acc[y][x];
    }
}

}  // namespace

extern "C" int dln_mlp_fold(float* params_flat, long long off_views_w, int ld_views, long long off_feature_w,
                            long long off_feature_b, long long off_views_b, long long off_M, long long off_bM,
                            void* stream) {
  DLN_CHECK_ARG(params_flat && ld_views >= 256 && off_views_w >= 0 && off_feature_w >= 0 && off_feature_b >= 0 &&
                off_views_b >= 0 && off_M >= 0 && off_bM >= 0);
  const FoldOffsets o{off_views_w, off_feature_w, off_feature_b, off_views_b, off_M, off_bM, ld_views};
  fold_kernel<<<128, 256, 0, (cudaStream_t)stream>>>(params_flat, o);
  return dln_launch_status();
}

extern "C" int dln_mlp_unfold_grads(const float* params_flat, float* grads_flat, long long off_views_w, int ld_views,
                                    long long off_feature_w, long long off_feature_b, long long off_views_b,
                                    long long off_M, long long off_bM, void* stream) {
  DLN_CHECK_ARG(params_flat && grads_flat && ld_views >= 256 && off_views_w >= 0 && off_feature_w >= 0 &&
                off_feature_b >= 0 && off_views_b >= 0 && off_M >= 0 && off_bM >= 0);
  const FoldOffsets o{off_views_w, off_feature_w, off_feature_b, off_views_b, off_M, off_bM, ld_views};
  unfold_kernel<<<97, 256, 0, (cudaStream_t)stream>>>(params_flat, grads_flat, o);
  return dln_launch_status();
}

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
