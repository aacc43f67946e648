// bxg_train.cu -- the two small kernels of the PPO loop that are launch-bound as chains of framework ops
// (SURVEY.md section 8 f-2): generalized advantage estimation and the policy's inference step of the rollout.
//
//   bxg_gae         brax/training/agents/ppo/losses.py:38-101 (compute_gae): one thread per trajectory, the
//                   reverse scan over the T steps in registers; replaces ~40 framework kernels per minibatch.
//   bxg_policy_act  brax/training/acting.py:33-53 with the PPO inference function (agents/ppo/networks.py:66-88):
//                   running-statistics normalisation (acme/running_statistics.py:303-328), the policy MLP
//                   [obs, h1, h2, 2*act] with swish, NormalTanhDistribution (distribution.py:142-172): loc,
//                   scale = softplus(raw) + min_std, raw action = loc + scale * noise, action = tanh(raw action).
//                   One thread per env, the weights staged in shared memory once per CTA.
//
// Both are plain float32 arithmetic (no tensor cores: the matrices are [27..244] x 64).  They are checked against
// the PyTorch statements of the same functions (tests/test_gpu_ppo_kernels.py), which tests/test_ppo_reference.py
// pins to the reference's own source.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/bxg.h"

namespace {

__global__ void gae_kernel(const float* __restrict__ trunc, const float* __restrict__ term, const float* __restrict__ reward,
                           const float* __restrict__ values, const float* __restrict__ bootstrap, int T, int64_t B,
                           float lambda, float discount, float* __restrict__ vs, float* __restrict__ adv) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  // vs_minus_v_xs by the reverse scan (losses.py:78-90), then vs and the advantages (losses.py:91-100)
  float acc = 0.f;
  float v_next = bootstrap[b];          // values_t_plus_1 at the current t
  float vs_next = bootstrap[b];         // vs_t_plus_1
  for (int t = T - 1; t >= 0; --t) {
    const int64_t i = (int64_t)t * B + b;
    const float mask = 1.f - trunc[i], tm = term[i], r = reward[i], v = values[i];
    const float delta = (r + discount * (1.f - tm) * v_next - v) * mask;
    acc = delta + discount * (1.f - tm) * mask * lambda * acc;
    const float vs_t = acc + v;
    adv[i] = (r + discount * (1.f - tm) * vs_next - v) * mask;
    vs[i] = vs_t;
    v_next = v; vs_next = vs_t;
  }
}

__device__ __forceinline__ float swish(float x) { return x / (1.f + expf(-x)); }
__device__ __forceinline__ float softplus(float x) { return x > 20.f ? x : log1pf(expf(x)); }

// dynamic shared memory: W1 [h1][obs], b1 [h1], W2 [h2][h1], b2 [h2], W3 [out][h2], b3 [out], mean [obs], std [obs]
template <int H1, int H2>
__global__ void __launch_bounds__(64) policy_act_kernel(const float* __restrict__ obs, const float* __restrict__ mean, const float* __restrict__ stdv,
                                  float clip, const float* __restrict__ W1, const float* __restrict__ b1,
                                  const float* __restrict__ W2, const float* __restrict__ b2, const float* __restrict__ W3,
                                  const float* __restrict__ b3, const float* __restrict__ noise, int64_t n, int no,
                                  int na, float min_std, float* __restrict__ logits, float* __restrict__ pre,
                                  float* __restrict__ action) {
  constexpr int h1 = H1, h2 = H2;     // compile-time widths: the activations live in registers
  extern __shared__ float sm[];
  float* sW1 = sm; float* sb1 = sW1 + h1 * no; float* sW2 = sb1 + h1; float* sb2 = sW2 + h2 * h1;
  float* sW3 = sb2 + h2; float* sb3 = sW3 + 2 * na * h2; float* smean = sb3 + 2 * na; float* sstd = smean + no;
  for (int i = threadIdx.x; i < h1 * no; i += blockDim.x) sW1[i] = W1[i];
  for (int i = threadIdx.x; i < h1; i += blockDim.x) sb1[i] = b1[i];
  for (int i = threadIdx.x; i < h2 * h1; i += blockDim.x) sW2[i] = W2[i];
  for (int i = threadIdx.x; i < h2; i += blockDim.x) sb2[i] = b2[i];
  for (int i = threadIdx.x; i < 2 * na * h2; i += blockDim.x) sW3[i] = W3[i];
  for (int i = threadIdx.x; i < 2 * na; i += blockDim.x) sb3[i] = b3[i];
  for (int i = threadIdx.x; i < no; i += blockDim.x) { smean[i] = mean[i]; sstd[i] = stdv[i]; }
  __syncthreads();
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float a1[H1], a2[H2];
  // layer 1 over the normalised observation (read once per output would be h1 * no global loads: keep it in a1 first)
  // the observation itself is at most a few hundred floats: stream it, accumulating all h1 outputs
#pragma unroll
  for (int j = 0; j < h1; ++j) a1[j] = sb1[j];
  for (int i = 0; i < no; ++i) {
    float x = (obs[e * no + i] - smean[i]) / sstd[i];
    if (clip > 0.f) x = fminf(fmaxf(x, -clip), clip);
#pragma unroll
    for (int j = 0; j < h1; ++j) a1[j] = fmaf(sW1[j * no + i], x, a1[j]);
  }
#pragma unroll
  for (int j = 0; j < h1; ++j) a1[j] = swish(a1[j]);
#pragma unroll
  for (int j = 0; j < h2; ++j) {
    float acc = sb2[j];
#pragma unroll
    for (int i = 0; i < h1; ++i) acc = fmaf(sW2[j * h1 + i], a1[i], acc);
    a2[j] = swish(acc);
  }
  for (int j = 0; j < na; ++j) {
    float loc = sb3[j], raw = sb3[na + j];
#pragma unroll
    for (int i = 0; i < h2; ++i) { loc = fmaf(sW3[j * h2 + i], a2[i], loc); raw = fmaf(sW3[(na + j) * h2 + i], a2[i], raw); }
    const float scale = softplus(raw) + min_std;
    const float p = fmaf(scale, noise[e * na + j], loc);
    logits[e * 2 * na + j] = loc; logits[e * 2 * na + na + j] = raw;
    pre[e * na + j] = p;
    action[e * na + j] = tanhf(p);
  }
}


// ---- PPO loss head (compute_ppo_loss, agents/ppo/losses.py:143-303, from the network outputs on) ----------------
// pass 1: per trajectory, termination / reward scaling / compute_gae; sum and sum of squares of the advantages
__global__ void ppo_gae_stats_kernel(const float* __restrict__ values, const float* __restrict__ reward, const float* __restrict__ done,
                                     const float* __restrict__ trunc, int T, int64_t B, float reward_scaling, float lambda, float discount,
                                     float* __restrict__ vs, float* __restrict__ adv, double* __restrict__ stats) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double s1 = 0.0, s2 = 0.0;
  if (b < B) {
    float acc = 0.f, v_next = values[(int64_t)T * B + b], vs_next = v_next;
    for (int t = T - 1; t >= 0; --t) {
      const int64_t i = (int64_t)t * B + b;
      const float tr = trunc[i], mask = 1.f - tr, tm = done[i] * (1.f - tr), r = reward[i] * reward_scaling, v = values[i];
      const float delta = (r + discount * (1.f - tm) * v_next - v) * mask;
      acc = delta + discount * (1.f - tm) * mask * lambda * acc;
      const float vs_t = acc + v;
      const float a = (r + discount * (1.f - tm) * vs_next - v) * mask;
      adv[i] = a; vs[i] = vs_t;
      s1 += a; s2 += (double)a * a;
      v_next = v; vs_next = vs_t;
    }
  }
  for (int o = 16; o >= 1; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(stats, s1); atomicAdd(stats + 1, s2); }
}

__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

// pass 2: one thread per (t, b): the three loss terms and their gradients wrt the policy logits and the values
__global__ void ppo_head_kernel(const float* __restrict__ logits, const float* __restrict__ values, const float* __restrict__ beh,
                                const float* __restrict__ pre, const float* __restrict__ noise, const float* __restrict__ vs,
                                const float* __restrict__ adv, const double* __restrict__ stats, int T, int64_t B, int A,
                                float eps_clip, float entropy_cost, int normalize_adv, float min_std, float* __restrict__ loss,
                                float* __restrict__ dlogits, float* __restrict__ dvalues) {
  const int64_t N = (int64_t)T * B;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float lp_loss = 0.f, lv_loss = 0.f, le_loss = 0.f;
  if (i < N) {
    const float inv_n = 1.f / (float)N;
    float a = adv[i];
    if (normalize_adv) {
      const double mean = stats[0] / (double)N;
      double var = stats[1] / (double)N - mean * mean;
      var = var > 0.0 ? var : 0.0;
      a = (float)(((double)a - mean) / (sqrt(var) + 1e-8));
    }
    // log-prob difference (the tanh log-det-jacobian of the raw action is common to both and cancels) and the entropy
    float dl = 0.f, ent = 0.f;
    for (int j = 0; j < A; ++j) {
      const float mu = logits[i * 2 * A + j], raw = logits[i * 2 * A + A + j];
      const float sg = softplus(raw) + min_std;
      const float mub = beh[i * 2 * A + j], sgb = softplus(beh[i * 2 * A + A + j]) + min_std;
      const float p = pre[i * A + j];
      const float z = (p - mu) / sg, zb = (p - mub) / sgb;
      dl += (-0.5f * z * z - logf(sg)) - (-0.5f * zb * zb - logf(sgb));
      const float sm = mu + sg * noise[i * A + j];
      ent += 0.5f + 0.9189385332f + logf(sg) + 2.f * (0.6931471806f - sm - softplus(-2.f * sm));
    }
    const float rho = expf(dl);
    const float lo = 1.f - eps_clip, hi = 1.f + eps_clip;
    const float rc = fminf(fmaxf(rho, lo), hi);
    const float s1 = rho * a, s2 = rc * a;
    lp_loss = -fminf(s1, s2) * inv_n;
    const float in_range = (rho >= lo && rho <= hi) ? 1.f : 0.f;
    // d(-min(s1, s2)) / d rho, torch.minimum's convention (the smaller takes the gradient, a tie splits it)
    float dmin_drho = s1 < s2 ? a : (s1 > s2 ? a * in_range : 0.5f * a + 0.5f * a * in_range);
    const float dlp = -dmin_drho * rho * inv_n;         // d loss / d log-prob
    le_loss = -entropy_cost * ent * inv_n;
    for (int j = 0; j < A; ++j) {
      const float mu = logits[i * 2 * A + j], raw = logits[i * 2 * A + A + j];
      const float sg = softplus(raw) + min_std;
      const float p = pre[i * A + j], nz = noise[i * A + j];
      const float d = p - mu;
      const float th = tanhf(mu + sg * nz);
      const float g_mu = dlp * (d / (sg * sg)) + (-entropy_cost * inv_n) * (-2.f * th);
      const float g_sg = dlp * (d * d / (sg * sg * sg) - 1.f / sg) + (-entropy_cost * inv_n) * (1.f / sg - 2.f * th * nz);
      dlogits[i * 2 * A + j] = g_mu;
      dlogits[i * 2 * A + A + j] = g_sg * sigmoidf(raw);
    }
    const float ve = vs[i] - values[i];
    lv_loss = 0.25f * ve * ve * inv_n;
    dvalues[i] = -0.5f * ve * inv_n;
  } else if (i < N + B) {
    dvalues[i] = 0.f;                                   // the bootstrap value enters through stop_gradient only
  }
  for (int o = 16; o >= 1; o >>= 1) {
    lp_loss += __shfl_xor_sync(0xffffffffu, lp_loss, o); lv_loss += __shfl_xor_sync(0xffffffffu, lv_loss, o); le_loss += __shfl_xor_sync(0xffffffffu, le_loss, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(loss + 1, lp_loss); atomicAdd(loss + 2, lv_loss); atomicAdd(loss + 3, le_loss); atomicAdd(loss, lp_loss + lv_loss + le_loss); }
}

}  // namespace

extern "C" {

int bxg_gae(const float* truncation, const float* termination, const float* reward, const float* values,
            const float* bootstrap, int32_t T, int64_t B, float lambda, float discount, float* vs, float* advantages,
            void* stream) {
  if (!truncation || !termination || !reward || !values || !bootstrap || !vs || !advantages || T < 1 || B < 0) return BXG_E_INVALID;
  if (B == 0) return BXG_OK;
  const int threads = 128;
  gae_kernel<<<(unsigned)((B + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      truncation, termination, reward, values, bootstrap, T, B, lambda, discount, vs, advantages);
  return cudaGetLastError() == cudaSuccess ? BXG_OK : BXG_E_CUDA;
}

int bxg_policy_act(const float* obs, const float* mean, const float* std, float clip, const float* W1, const float* b1,
                   const float* W2, const float* b2, const float* W3, const float* b3, const float* noise, int64_t n,
                   int32_t obs_size, int32_t h1, int32_t h2, int32_t act_size, float min_std, float* logits, float* pre,
                   float* action, void* stream) {
  if (!obs || !mean || !std || !W1 || !b1 || !W2 || !b2 || !W3 || !b3 || !noise || !logits || !pre || !action) return BXG_E_INVALID;
  if (h1 != 64 || h2 != 64 || obs_size < 1 || act_size < 1) return BXG_E_UNSUPPORTED;   // compiled for the [obs, 64, 64, 2 act] policy (notebooks/training_torch.ipynb)
  if (n <= 0) return n == 0 ? BXG_OK : BXG_E_INVALID;
  const size_t smem = sizeof(float) * ((size_t)h1 * obs_size + h1 + (size_t)h2 * h1 + h2 + (size_t)2 * act_size * h2 + 2 * act_size + 2 * obs_size);
  if (smem > 200 * 1024) return BXG_E_UNSUPPORTED;
  static bool attr_set = false;   // (the attribute is per function; setting it again is harmless)
  if (!attr_set) { cudaFuncSetAttribute(policy_act_kernel<64, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr_set = true; }
  const int threads = 64;
  policy_act_kernel<64, 64><<<(unsigned)((n + threads - 1) / threads), threads, smem, (cudaStream_t)stream>>>(
      obs, mean, std, clip, W1, b1, W2, b2, W3, b3, noise, n, obs_size, act_size, min_std, logits, pre, action);
  return cudaGetLastError() == cudaSuccess ? BXG_OK : BXG_E_CUDA;
}

int bxg_ppo_head(const float* logits, const float* values, const float* behaviour_logits, const float* raw_action,
                 const float* reward, const float* done, const float* truncation, const float* entropy_noise, int32_t T, int64_t B,
                 int32_t act_size, float reward_scaling, float lambda, float discount, float clip_epsilon, float entropy_cost,
                 int32_t normalize_advantage, float min_std, float* vs_scratch, float* adv_scratch, double* stats_scratch,
                 float* loss4, float* dlogits, float* dvalues, void* stream) {
  if (!logits || !values || !behaviour_logits || !raw_action || !reward || !done || !truncation || !entropy_noise || !vs_scratch ||
      !adv_scratch || !stats_scratch || !loss4 || !dlogits || !dvalues || T < 1 || B < 1 || act_size < 1) return BXG_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(stats_scratch, 0, 2 * sizeof(double), st) != cudaSuccess || cudaMemsetAsync(loss4, 0, 4 * sizeof(float), st) != cudaSuccess) return BXG_E_CUDA;
  const int threads = 128;
  ppo_gae_stats_kernel<<<(unsigned)((B + threads - 1) / threads), threads, 0, st>>>(values, reward, done, truncation, T, B, reward_scaling, lambda,
                                                                                   discount, vs_scratch, adv_scratch, stats_scratch);
  const int64_t n = (int64_t)T * B + B;
  ppo_head_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, st>>>(logits, values, behaviour_logits, raw_action, entropy_noise, vs_scratch,
                                                                              adv_scratch, stats_scratch, T, B, act_size, clip_epsilon, entropy_cost,
                                                                              normalize_advantage, min_std, loss4, dlogits, dvalues);
  return cudaGetLastError() == cudaSuccess ? BXG_OK : BXG_E_CUDA;
}

}  // extern "C"
