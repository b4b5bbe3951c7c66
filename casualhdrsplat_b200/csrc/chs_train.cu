// chs_train.cu — the two streaming steps that sit immediately after the formation path in a trainer
// (SURVEY.md section 8(f) row f4): the photometric loss on the blurred LDR frame, emitting dL/dB in the
// same pass so that the upstream gradient never round-trips through the host framework, and an Adam
// update applied directly to a section of the (all-reduced) flat gradient buffer.
// Both are pure HBM streams: 128-bit loads/stores, grid-stride, one fp64 atomic per block for the loss.
#include "chs_common.cuh"

namespace {

constexpr int kThreads = 256;

template <int KIND>  // 0: L2 (0.5 * scale * sum d^2), 1: L1 (scale * sum |d|)
__global__ void __launch_bounds__(kThreads) loss_kernel(const float* __restrict__ ldr, const float* __restrict__ target, uint64_t n,
                                                        float scale, float* __restrict__ v_ldr, double* __restrict__ loss_acc) {
  const uint64_t n4 = n / 4;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = reinterpret_cast<const float4*>(ldr)[i];
    const float4 b = chs_ldg_stream(reinterpret_cast<const float4*>(target) + i);
    const float d[4] = {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w};
    float g[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (KIND == 0) {
        acc = fmaf(d[k], d[k], acc);
        g[k] = scale * d[k];
      } else {
        acc += fabsf(d[k]);
        g[k] = d[k] > 0.f ? scale : (d[k] < 0.f ? -scale : 0.f);
      }
    }
    reinterpret_cast<float4*>(v_ldr)[i] = make_float4(g[0], g[1], g[2], g[3]);
  }
  if (blockIdx.x == 0)
    for (uint64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
      const float d = ldr[i] - target[i];
      if (KIND == 0) {
        acc = fmaf(d, d, acc);
        v_ldr[i] = scale * d;
      } else {
        acc += fabsf(d);
        v_ldr[i] = d > 0.f ? scale : (d < 0.f ? -scale : 0.f);
      }
    }
  acc = chs_warp_sum(acc);
  __shared__ float s_part[kThreads / 32];
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) s += s_part[w];
    atomicAdd(loss_acc, (double)s * (KIND == 0 ? 0.5 * (double)scale : (double)scale));
  }
}

__global__ void __launch_bounds__(kThreads) adam_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m,
                                                        float* __restrict__ v, uint64_t n, float lr_t, float beta1, float beta2,
                                                        float eps_t, float grad_scale) {
  // lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t), eps_t = eps * sqrt(1 - beta2^t): the bias corrections folded on the host
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float g = grad[i] * grad_scale;
    const float mi = fmaf(beta1, m[i], (1.f - beta1) * g);
    const float vi = fmaf(beta2, v[i], (1.f - beta2) * g * g);
    m[i] = mi;
    v[i] = vi;
    param[i] -= lr_t * mi / (sqrtf(vi) + eps_t);
  }
}

}  // namespace

extern "C" int chs_loss(int32_t kind, const float* ldr, const float* target, uint64_t count, float scale, float* v_ldr,
                        double* loss_acc, void* stream) {
  CHS_REQUIRE(kind == 0 || kind == 1, "chs_loss: kind must be 0 (L2) or 1 (L1)");
  CHS_REQUIRE(ldr && target && v_ldr && loss_acc, "chs_loss: null pointer");
  CHS_REQUIRE((((uintptr_t)ldr | (uintptr_t)target | (uintptr_t)v_ldr) % 16) == 0, "chs_loss: buffers must be 16-byte aligned");
  if (count == 0) return CHS_OK;
  uint64_t want = (count / 4 + kThreads - 1) / kThreads;
  int blocks = (int)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
  if (kind == 0)
    loss_kernel<0><<<blocks, kThreads, 0, (cudaStream_t)stream>>>(ldr, target, count, scale, v_ldr, loss_acc);
  else
    loss_kernel<1><<<blocks, kThreads, 0, (cudaStream_t)stream>>>(ldr, target, count, scale, v_ldr, loss_acc);
  CHS_LAUNCH_CHECK();
  return CHS_OK;
}

extern "C" int chs_adam_step(float* param, const float* grad, float* m, float* v, uint64_t count, float lr, float beta1, float beta2,
                             float eps, int32_t step, float grad_scale, void* stream) {
  CHS_REQUIRE(param && grad && m && v, "chs_adam_step: null pointer");
  CHS_REQUIRE(step >= 1 && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f, "chs_adam_step: bad step / betas");
  if (count == 0) return CHS_OK;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const float lr_t = (float)((double)lr * sqrt(bc2) / bc1), eps_t = (float)((double)eps * sqrt(bc2));
  uint64_t want = (count + kThreads - 1) / kThreads;
  int blocks = (int)(want > 148 * 16 ? 148 * 16 : want);
  adam_kernel<<<blocks, kThreads, 0, (cudaStream_t)stream>>>(param, grad, m, v, count, lr_t, beta1, beta2, eps_t, grad_scale);
  CHS_LAUNCH_CHECK();
  return CHS_OK;
}
