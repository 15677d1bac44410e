// Multi-tensor optimizer steps for the hot-path parameters: the reference builds two optimizers
// with ONE PARAMETER GROUP PER PARAMETER (utils/build.py:497-508), i.e. 54 tiny launches (x3-4 ops) per step for the
// teacher + adapter alone. Here every tensor of an optimizer is updated by one launch: a device table lists the
// tensors, a second one cuts them into equal chunks so that the grid is load-balanced whatever the size mix
// (a 7056x256 matrix next to 64-element biases).
//   SGD   (torch.optim.SGD, momentum, dampening 0, no nesterov):  d = g + wd*p;  buf = first ? d : mu*buf + d;  p -= lr*buf
//   AdamW (torch.optim.AdamW): p *= 1 - lr*wd;  m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;
//                              p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
#include "common.cuh"

namespace lgd {

constexpr int MT_CHUNK = 16384;  // elements per block

__global__ void __launch_bounds__(256)
mt_sgd_kernel(const lgd_mt_tensor_t* __restrict__ tensors, const int2* __restrict__ chunks, float lr, float wd, float mu,
              int first) {
  const int2 ck = chunks[blockIdx.x];
  const lgd_mt_tensor_t t = tensors[ck.x];
  const long long begin = (long long)ck.y * MT_CHUNK;
  const long long end = min(begin + (long long)MT_CHUNK, (long long)t.numel);
  float* __restrict__ p = t.param;
  const float* __restrict__ g = t.grad;
  float* __restrict__ m = t.state0;
  for (long long i = begin + threadIdx.x; i < end; i += blockDim.x) {
    const float pv = p[i];
    float d = g[i];
    if (wd != 0.f) d = fmaf(wd, pv, d);
    if (mu != 0.f) {
      const float b = first ? d : fmaf(mu, m[i], d);
      m[i] = b;
      d = b;
    }
    p[i] = fmaf(-lr, d, pv);
  }
}

__global__ void __launch_bounds__(256)
mt_adamw_kernel(const lgd_mt_tensor_t* __restrict__ tensors, const int2* __restrict__ chunks, float lr, float wd,
                float b1, float b2, float eps, float bc1, float bc2_sqrt) {
  const int2 ck = chunks[blockIdx.x];
  const lgd_mt_tensor_t t = tensors[ck.x];
  const long long begin = (long long)ck.y * MT_CHUNK;
  const long long end = min(begin + (long long)MT_CHUNK, (long long)t.numel);
  float* __restrict__ p = t.param;
  const float* __restrict__ g = t.grad;
  float* __restrict__ m = t.state0;
  float* __restrict__ v = t.state1;
  const float step_size = lr / bc1;
  for (long long i = begin + threadIdx.x; i < end; i += blockDim.x) {
    const float gv = g[i];
    float pv = p[i];
    pv = pv * (1.f - lr * wd);
    const float mv = m[i] + (1.f - b1) * (gv - m[i]);            // lerp, as torch
    const float vv = b2 * v[i] + (1.f - b2) * gv * gv;
    m[i] = mv;
    v[i] = vv;
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    p[i] = pv - step_size * (mv / denom);
  }
}

}  // namespace lgd

using namespace lgd;

extern "C" int lgd_mt_chunk_elems(void) { return MT_CHUNK; }

extern "C" int lgd_mt_sgd(const lgd_mt_tensor_t* tensors_dev, const int32_t* chunks_dev, int num_chunks, float lr,
                          float weight_decay, float momentum, int first_step, void* stream) {
  LGD_CHECK_ARG(tensors_dev && chunks_dev && num_chunks > 0, "lgd_mt_sgd: bad arguments");
  mt_sgd_kernel<<<num_chunks, 256, 0, (cudaStream_t)stream>>>(tensors_dev, reinterpret_cast<const int2*>(chunks_dev), lr,
                                                               weight_decay, momentum, first_step);
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}

extern "C" int lgd_mt_adamw(const lgd_mt_tensor_t* tensors_dev, const int32_t* chunks_dev, int num_chunks, float lr,
                            float weight_decay, float beta1, float beta2, float eps, int step, void* stream) {
  LGD_CHECK_ARG(tensors_dev && chunks_dev && num_chunks > 0 && step >= 1, "lgd_mt_adamw: bad arguments");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  mt_adamw_kernel<<<num_chunks, 256, 0, (cudaStream_t)stream>>>(tensors_dev, reinterpret_cast<const int2*>(chunks_dev), lr,
                                                                 weight_decay, beta1, beta2, eps, (float)bc1,
                                                                 (float)sqrt(bc2));
  LGD_LAUNCH_CHECK();
  return LGD_OK;
}
