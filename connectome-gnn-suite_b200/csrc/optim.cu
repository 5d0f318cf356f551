// optim.cu - the step AFTER backward (reference train.py:51 `optimizer.step()`, demo.py:110,130 Adam with
// weight decay): one kernel over a flat parameter / gradient buffer instead of one launch chain per parameter tensor,
// with the step counter on the device so that the whole training step can be replayed as a CUDA graph.
//
//   cgnn_step_tick   state[0] += 1 (step counter); dropout salt words of this step from (seed, step)
//   cgnn_adam_step   torch.optim.Adam's update (L2 weight decay folded into the gradient, bias-corrected moments),
//                    op for op in fp32 as the single-tensor implementation evaluates it
#include "common.cuh"

namespace cgnn {

// state block (device, 4 x uint64): [0] step counter, [1] seed, [2] the two 32-bit salt words (low, high), [3] spare
__global__ void k_step_tick(unsigned long long* state) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const unsigned long long step = state[0] + 1ull;
    state[0] = step;
    unsigned long long x = state[1] * 0x9E3779B97F4A7C15ull + step * 0xBF58476D1CE4E5B9ull;   // splitmix64 of (seed, step)
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    state[2] = x;
  }
}

struct AdamArgs {
  float* param; const float* grad; float* exp_avg; float* exp_avg_sq; long long n;
  float lr, beta1, beta2, eps, weight_decay;
  const unsigned long long* state;   // state[0] = step AFTER this step's tick
};

__global__ void __launch_bounds__(256) k_adam(AdamArgs p) {
  // bias corrections as torch computes them on the host (doubles), here per thread from the device-side step
  const double step = (double)p.state[0];
  const double bc1 = 1.0 - pow((double)p.beta1, step), bc2 = 1.0 - pow((double)p.beta2, step);
  const float step_size = (float)((double)p.lr / bc1);
  const float inv_bc2_sqrt = 1.0f / (float)sqrt(bc2);     // ATen divides a tensor by a host scalar as a multiplication by its reciprocal
  const float w1 = (float)(1.0 - (double)p.beta1), w2 = (float)(1.0 - (double)p.beta2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (long long)gridDim.x * blockDim.x) {
    const float w = p.param[i];
    float g = p.grad[i];
    if (p.weight_decay != 0.0f) g = fmaf(p.weight_decay, w, g);          // grad.add(param, alpha=weight_decay)
    float m = p.exp_avg[i], v = p.exp_avg_sq[i];
    m = fmaf(w1, g - m, m);                                               // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(__fmul_rn(w2, g), g, __fmul_rn(v, p.beta2));                 // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    const float denom = __fadd_rn(__fmul_rn(sqrtf(v), inv_bc2_sqrt), p.eps);   // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
    p.param[i] = fmaf(-step_size, __fdiv_rn(m, denom), w);                // param.addcdiv_(exp_avg, denom, value=-step_size)
    p.exp_avg[i] = m;
    p.exp_avg_sq[i] = v;
  }
}

}  // namespace cgnn

extern "C" {

int cgnn_step_tick(uint64_t* state, cgnn_stream_t stream_) {
  if (!state) return CGNN_ERR_INVALID_ARG;
  auto kfn = cgnn::k_step_tick;
  CGNN_LAUNCH(kfn, 1, 32, 0, (cudaStream_t)stream_, (unsigned long long*)state);
  CGNN_CHECK_LAUNCH();
  return CGNN_OK;
}

int cgnn_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, const uint64_t* state, cgnn_stream_t stream_) {
  if (n < 0 || !state) return CGNN_ERR_INVALID_ARG;
  if (n == 0) return CGNN_OK;
  if (!param || !grad || !exp_avg || !exp_avg_sq) return CGNN_ERR_INVALID_ARG;
  cgnn::AdamArgs a;
  a.param = param; a.grad = grad; a.exp_avg = exp_avg; a.exp_avg_sq = exp_avg_sq; a.n = n;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  a.state = (const unsigned long long*)state;
  long long blocks = (n + 255) / 256;
  if (blocks > 1184) blocks = 1184;      // 8 x 148
  auto kfn = cgnn::k_adam;
  CGNN_LAUNCH(kfn, (unsigned)blocks, 256, 0, (cudaStream_t)stream_, a);
  CGNN_CHECK_LAUNCH();
  return CGNN_OK;
}

}  // extern "C"
