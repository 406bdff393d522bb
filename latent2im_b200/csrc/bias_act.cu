// fused bias + activation (forward, grad=1 pass, fused backward with bias reduction).
// Replaces op/fused_bias_act_kernel.cu of the reference.  HBM-bound: one 16-byte load and one
// 16-byte store per thread per iteration, 64-bit indexing, bias index computed once per vector.
#include "common.cuh"

namespace l2i {

template <typename T> struct Vec16 { static constexpr int N = 16 / sizeof(T); };

template <typename T>
__device__ __forceinline__ float act_apply(float v, float refv, int code, float alpha) {
  // code = act*10+grad, op/fused_bias_act_kernel.cu:36-46
  switch (code) {
    case 30: return v > 0.f ? v : v * alpha;
    case 31: return refv > 0.f ? v : v * alpha;
    case 12:
    case 32: return 0.f;
    default: return v;
  }
}

// vector kernel: requires n % V == 0 and (step_b % V == 0 || (step_b == 1 && size_b % V == 0)) and
// 16-byte aligned pointers.
template <typename T, bool kBiasPerVec>
__global__ void __launch_bounds__(256) bias_act_vec_kernel(T* __restrict__ y, const T* __restrict__ x,
                                                           const T* __restrict__ bias,
                                                           const T* __restrict__ ref, int64_t nvec,
                                                           int64_t step_b, int64_t size_b, int code,
                                                           float alpha, float scale) {
  constexpr int V = Vec16<T>::N;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const int4 xv = __ldcs(reinterpret_cast<const int4*>(x) + i);
    int4 rv = make_int4(0, 0, 0, 0);
    if (ref != nullptr) rv = __ldcs(reinterpret_cast<const int4*>(ref) + i);
    const T* xs = reinterpret_cast<const T*>(&xv);
    const T* rs = reinterpret_cast<const T*>(&rv);
    int4 ov;
    T* os = reinterpret_cast<T*>(&ov);
    float b0 = 0.f;
    int64_t c0 = 0;
    if (bias != nullptr) {
      if (kBiasPerVec) {
        b0 = to_f32<T>(bias[((i * V) / step_b) % size_b]);
      } else {
        c0 = (i * V) % size_b;  // step_b == 1
      }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float v = to_f32<T>(xs[j]);
      if (bias != nullptr) v += kBiasPerVec ? b0 : to_f32<T>(bias[c0 + j]);
      // the reference rounds x + b to scalar_t before the activation; for fp32 this is the same
      v = to_f32<T>(from_f32<T>(v));
      float o = act_apply<T>(v, to_f32<T>(rs[j]), code, alpha) * scale;
      os[j] = from_f32<T>(o);
    }
    __stcs(reinterpret_cast<int4*>(y) + i, ov);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) bias_act_scalar_kernel(T* __restrict__ y, const T* __restrict__ x,
                                                              const T* __restrict__ bias,
                                                              const T* __restrict__ ref, int64_t n,
                                                              int64_t step_b, int64_t size_b, int code,
                                                              float alpha, float scale) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = to_f32<T>(x[i]);
    if (bias != nullptr) v += to_f32<T>(bias[(i / step_b) % size_b]);
    v = to_f32<T>(from_f32<T>(v));
    float r = ref != nullptr ? to_f32<T>(ref[i]) : 0.f;
    y[i] = from_f32<T>(act_apply<T>(v, r, code, alpha) * scale);
  }
}

template <typename T>
static int launch_bias_act(void* y, const void* x, const void* bias, const void* ref, int64_t n,
                           int64_t step_b, int64_t size_b, int code, float alpha, float scale,
                           cudaStream_t st) {
  constexpr int V = Vec16<T>::N;
  const bool aligned = ((uintptr_t)y % 16 == 0) && ((uintptr_t)x % 16 == 0) &&
                       (ref == nullptr || (uintptr_t)ref % 16 == 0) && (n % V == 0);
  const bool has_bias = bias != nullptr && size_b > 0;
  const T* b = has_bias ? (const T*)bias : nullptr;
  if (aligned && (!has_bias || step_b % V == 0)) {
    int64_t nvec = n / V;
    int blocks = (int)std::min<int64_t>(ceil_div64(nvec, 256), (int64_t)kNumSMs * 16);
    bias_act_vec_kernel<T, true><<<blocks, 256, 0, st>>>((T*)y, (const T*)x, b, (const T*)ref, nvec,
                                                         step_b, size_b, code, alpha, scale);
  } else if (aligned && step_b == 1 && size_b % V == 0) {
    int64_t nvec = n / V;
    int blocks = (int)std::min<int64_t>(ceil_div64(nvec, 256), (int64_t)kNumSMs * 16);
    bias_act_vec_kernel<T, false><<<blocks, 256, 0, st>>>((T*)y, (const T*)x, b, (const T*)ref, nvec,
                                                          step_b, size_b, code, alpha, scale);
  } else {
    int blocks = (int)std::min<int64_t>(ceil_div64(n, 256), (int64_t)kNumSMs * 16);
    bias_act_scalar_kernel<T><<<blocks, 256, 0, st>>>((T*)y, (const T*)x, b, (const T*)ref, n, step_b,
                                                      has_bias ? size_b : 1, code, alpha, scale);
  }
  return check_launch("fused_bias_act");
}

// ---- fused backward: grad_in and grad_bias in one pass over (grad_out, out) --------------------
// grid = (size_b, splits); each block reduces its share of the [outer, step_b] plane of channel c.
template <typename T>
__global__ void __launch_bounds__(256) lrelu_bwd_kernel(T* __restrict__ gin, float* __restrict__ gbias,
                                                        const T* __restrict__ gout,
                                                        const T* __restrict__ out, int64_t outer,
                                                        int64_t size_b, int64_t step_b, float alpha,
                                                        float scale) {
  const int64_t c = blockIdx.x;
  const int64_t per_c = outer * step_b;
  float acc = 0.f;
  for (int64_t j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; j < per_c;
       j += (int64_t)gridDim.y * blockDim.x) {
    const int64_t o = j / step_b, s = j - o * step_b;
    const int64_t idx = (o * size_b + c) * step_b + s;
    float g = to_f32<T>(gout[idx]);
    float r = to_f32<T>(out[idx]);
    float gi = (r > 0.f ? g : g * alpha) * scale;
    T gq = from_f32<T>(gi);
    gin[idx] = gq;
    acc += to_f32<T>(gq);
  }
  if (gbias == nullptr) return;
  __shared__ float red[8];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = red[threadIdx.x];
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) v += __shfl_xor_sync(0xffu, v, off);
    if (threadIdx.x == 0) atomicAdd(gbias + c, v);
  }
}

// Vectorised variant (step_b % V == 0, 16-byte aligned): one 16-byte load of grad_out and of out and one 16-byte store per vector,
// one 64-bit division per vector instead of per element, two vectors in flight per thread.
template <typename T>
__global__ void __launch_bounds__(256) lrelu_bwd_vec_kernel(T* __restrict__ gin, float* __restrict__ gbias,
                                                            const T* __restrict__ gout, const T* __restrict__ out,
                                                            int64_t outer, int64_t size_b, int64_t step_v, float alpha, float scale) {
  constexpr int V = Vec16<T>::N;
  const int64_t c = blockIdx.x;
  const int64_t per_c = outer * step_v;          // vectors of channel c
  float acc = 0.f;
  auto one = [&](int64_t j) {
    const int64_t o = j / step_v, sv = j - o * step_v;
    const int64_t iv = (o * size_b + c) * step_v + sv;
    const int4 gv = __ldcs(reinterpret_cast<const int4*>(gout) + iv);
    const int4 rv = __ldcs(reinterpret_cast<const int4*>(out) + iv);
    const T* gs = reinterpret_cast<const T*>(&gv);
    const T* rs = reinterpret_cast<const T*>(&rv);
    int4 ov;
    T* os = reinterpret_cast<T*>(&ov);
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float g = to_f32<T>(gs[e]);
      const T gq = from_f32<T>((to_f32<T>(rs[e]) > 0.f ? g : g * alpha) * scale);
      os[e] = gq;
      acc += to_f32<T>(gq);
    }
    __stcs(reinterpret_cast<int4*>(gin) + iv, ov);
  };
  const int64_t stride = (int64_t)gridDim.y * blockDim.x;
  int64_t j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  for (; j + stride < per_c; j += 2 * stride) { one(j); one(j + stride); }
  if (j < per_c) one(j);
  if (gbias == nullptr) return;
  __shared__ float red[8];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = red[threadIdx.x];
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) v += __shfl_xor_sync(0xffu, v, off);
    if (threadIdx.x == 0) atomicAdd(gbias + c, v);
  }
}

template <typename T>
static int launch_lrelu_bwd(void* gin, float* gbias, const void* gout, const void* out, int64_t outer,
                            int64_t size_b, int64_t step_b, float alpha, float scale, cudaStream_t st) {
  if (gbias != nullptr) L2I_CUDA_TRY(cudaMemsetAsync(gbias, 0, sizeof(float) * size_b, st));
  const int64_t per_c = outer * step_b;
  int splits = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(per_c, 256 * 8),
                                                           ceil_div64((int64_t)kNumSMs * 8, size_b)));
  constexpr int V = Vec16<T>::N;
  if (step_b % V == 0 && (uintptr_t)gin % 16 == 0 && (uintptr_t)gout % 16 == 0 && (uintptr_t)out % 16 == 0) {
    const int64_t per_cv = per_c / V;
    const int vsplits = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(per_cv, 256 * 4), ceil_div64((int64_t)kNumSMs * 8, size_b)));
    lrelu_bwd_vec_kernel<T><<<dim3((unsigned)size_b, (unsigned)vsplits), 256, 0, st>>>((T*)gin, gbias, (const T*)gout, (const T*)out, outer,
                                                                                       size_b, step_b / V, alpha, scale);
    return check_launch("fused_leaky_relu_bwd");
  }
  dim3 grid((unsigned)size_b, (unsigned)splits);
  lrelu_bwd_kernel<T><<<grid, 256, 0, st>>>((T*)gin, gbias, (const T*)gout, (const T*)out, outer, size_b,
                                            step_b, alpha, scale);
  return check_launch("fused_leaky_relu_bwd");
}

}  // namespace l2i

using namespace l2i;

extern "C" int l2i_fused_bias_act(void* y, const void* x, const void* bias, const void* ref, int64_t n,
                                  int64_t step_b, int64_t size_b, int act, int grad, float alpha,
                                  float scale, int dtype, void* stream) {
  L2I_REQUIRE(n >= 0, "fused_bias_act: negative element count");
  if (n == 0) return L2I_OK;
  L2I_REQUIRE(y && x, "fused_bias_act: null tensor");
  L2I_REQUIRE(act == 1 || act == 3, "fused_bias_act: act must be 1 (linear) or 3 (lrelu), got %d", act);
  L2I_REQUIRE(grad >= 0 && grad <= 2, "fused_bias_act: grad must be 0..2, got %d", grad);
  L2I_REQUIRE(!(act == 3 && grad == 1 && ref == nullptr), "fused_bias_act: grad=1 needs ref");
  L2I_REQUIRE(bias == nullptr || (size_b > 0 && step_b > 0), "fused_bias_act: bad bias geometry");
  const int code = act * 10 + grad;
  cudaStream_t st = (cudaStream_t)stream;
  if (size_b <= 0) bias = nullptr;
  switch (dtype) {
    case L2I_F32: return launch_bias_act<float>(y, x, bias, ref, n, step_b, size_b, code, alpha, scale, st);
    case L2I_BF16:
      return launch_bias_act<__nv_bfloat16>(y, x, bias, ref, n, step_b, size_b, code, alpha, scale, st);
    case L2I_F16: return launch_bias_act<__half>(y, x, bias, ref, n, step_b, size_b, code, alpha, scale, st);
    default: set_error("fused_bias_act: unsupported dtype %d", dtype); return L2I_ERR_UNSUPPORTED;
  }
}

extern "C" int l2i_fused_leaky_relu_bwd(void* grad_in, float* grad_bias, const void* grad_out,
                                        const void* out, int64_t outer, int64_t size_b, int64_t step_b,
                                        float alpha, float scale, int dtype, void* stream) {
  L2I_REQUIRE(outer >= 0 && size_b > 0 && step_b > 0, "fused_leaky_relu_bwd: bad geometry");
  if (outer == 0) {
    if (grad_bias) L2I_CUDA_TRY(cudaMemsetAsync(grad_bias, 0, sizeof(float) * size_b, (cudaStream_t)stream));
    return L2I_OK;
  }
  L2I_REQUIRE(grad_in && grad_out && out, "fused_leaky_relu_bwd: null tensor");
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case L2I_F32: return launch_lrelu_bwd<float>(grad_in, grad_bias, grad_out, out, outer, size_b, step_b, alpha, scale, st);
    case L2I_BF16:
      return launch_lrelu_bwd<__nv_bfloat16>(grad_in, grad_bias, grad_out, out, outer, size_b, step_b, alpha, scale, st);
    case L2I_F16: return launch_lrelu_bwd<__half>(grad_in, grad_bias, grad_out, out, outer, size_b, step_b, alpha, scale, st);
    default: set_error("fused_leaky_relu_bwd: unsupported dtype %d", dtype); return L2I_ERR_UNSUPPORTED;
  }
}
