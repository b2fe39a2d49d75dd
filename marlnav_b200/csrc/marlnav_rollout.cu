// marlnav_b200/csrc/marlnav_rollout.cu
//
// The two caller-side pieces SURVEY.md section 8(f) ranks next after the environment step
// (rows f-2 and f-3), as device kernels behind the same C ABI:
//
//   marlnav_actor_sample_f32        Actor.forward + dist.sample() + dist.log_prob()
//                                   /root/reference/marlnav/models.py:27-36, 113-115
//   marlnav_discounted_returns_f64  the backward scan of MAPPO._process_rewards
//                                   /root/reference/marlnav/models.py:131-139
//
// The step kernel stays the product's hot path; these remove the ~10 small launches and the
// per-row Cholesky of MultivariateNormal around it, and the 1000-iteration Python loop after it.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/marlnav_b200.h"

namespace mnr {

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// One thread per (env, agent) row.  The reference's Actor has NO activation after fc1
// (models.py:29-31): h = fc1(x); mu = tanh(fc_mu(h)); "std" = softplus(fc_std(h)), and that
// "std" is used as the diagonal COVARIANCE of MultivariateNormal (models.py:32-34), so the
// standard deviation of the sampled action is sqrt(softplus(.)).
template <int MAX_S, int MAX_H>
__global__ void __launch_bounds__(128)
actor_sample_kernel(const float* __restrict__ obs, long long N, int S, int H,
                    const float* __restrict__ w1, const float* __restrict__ b1,
                    const float* __restrict__ w_mu, const float* __restrict__ b_mu,
                    const float* __restrict__ w_std, const float* __restrict__ b_std,
                    const float* __restrict__ eps, uint64_t seed, uint64_t counter,
                    const unsigned long long* __restrict__ counter_dev,
                    float* __restrict__ actions, float* __restrict__ log_probs,
                    float* __restrict__ mu_out, float* __restrict__ var_out) {
    extern __shared__ float sm[];
    float* s_w1 = sm;                    // (H,S)
    float* s_b1 = s_w1 + H * S;          // (H)
    float* s_wm = s_b1 + H;              // (2,H)
    float* s_ws = s_wm + 2 * H;          // (2,H)
    for (int i = threadIdx.x; i < H * S; i += blockDim.x) s_w1[i] = w1[i];
    for (int i = threadIdx.x; i < H; i += blockDim.x) s_b1[i] = b1[i];
    for (int i = threadIdx.x; i < 2 * H; i += blockDim.x) { s_wm[i] = w_mu[i]; s_ws[i] = w_std[i]; }
    __syncthreads();
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= N) return;

    float x[MAX_S];
#pragma unroll
    for (int k = 0; k < MAX_S; ++k) x[k] = k < S ? obs[row * S + k] : 0.f;
    float m0 = b_mu[0], m1 = b_mu[1], v0 = b_std[0], v1 = b_std[1];
    for (int j = 0; j < H; ++j) {
        float h = s_b1[j];
#pragma unroll
        for (int k = 0; k < MAX_S; ++k)
            if (k < S) h = fmaf(x[k], s_w1[j * S + k], h);
        m0 = fmaf(h, s_wm[j], m0); m1 = fmaf(h, s_wm[H + j], m1);
        v0 = fmaf(h, s_ws[j], v0); v1 = fmaf(h, s_ws[H + j], v1);
    }
    m0 = tanhf(m0); m1 = tanhf(m1);
    v0 = v0 > 20.f ? v0 : log1pf(expf(v0));          // F.softplus, beta 1, threshold 20
    v1 = v1 > 20.f ? v1 : log1pf(expf(v1));

    float e0, e1;
    if (eps) { e0 = eps[row * 2]; e1 = eps[row * 2 + 1]; }
    else {
        if (counter_dev) counter += __ldg(counter_dev);      // host value = offset inside a batch
        const uint4 r = philox4x32_10((uint32_t)row, (uint32_t)((uint64_t)row >> 32), (uint32_t)counter,
                                      0x41435452u /* 'ACTR' */, (uint32_t)seed, (uint32_t)(seed >> 32));
        const float u1 = ((float)(r.x >> 8) + 1.0f) * 5.9604644775390625e-08f;      // (0, 1]
        const float u2 = (float)(r.y >> 8) * 5.9604644775390625e-08f;               // [0, 1)
        const float rad = sqrtf(-2.0f * logf(u1));
        float sn, cs;
        sincospif(2.0f * u2, &sn, &cs);
        e0 = rad * cs; e1 = rad * sn;
    }
    const float a0 = fmaf(sqrtf(v0), e0, m0), a1 = fmaf(sqrtf(v1), e1, m1);
    actions[row * 2] = a0; actions[row * 2 + 1] = a1;
    // MultivariateNormal(mu, diag(v)).log_prob(a), k = 2
    const float d0 = a0 - m0, d1 = a1 - m1;
    const float maha = d0 * d0 / v0 + d1 * d1 / v1;
    log_probs[row] = -0.5f * maha - 0.5f * (logf(v0) + logf(v1)) - 1.8378770664093453f;
    if (mu_out) { mu_out[row * 2] = m0; mu_out[row * 2 + 1] = m1; }
    if (var_out) { var_out[row * 2] = v0; var_out[row * 2 + 1] = v1; }
}

// Critic.forward (models.py:39-56): value = fc2(relu(fc1(flatten(x)))), one thread per env.
template <int MAX_H>
__global__ void __launch_bounds__(128)
critic_value_kernel(const float* __restrict__ obs, long long B, int K, int H, const float* __restrict__ w1,
                    const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                    float* __restrict__ values) {
    extern __shared__ float sm[];
    float* s_w1 = sm;                    // transposed to (K,H) so that a thread walks j contiguously
    for (int i = threadIdx.x; i < H * K; i += blockDim.x) { const int j = i / K, k = i - j * K; s_w1[k * H + j] = w1[i]; }
    __syncthreads();
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B) return;
    float h[MAX_H];
#pragma unroll
    for (int j = 0; j < MAX_H; ++j) h[j] = j < H ? b1[j] : 0.f;
    const float* x = obs + e * K;
    for (int k = 0; k < K; ++k) {
        const float xk = x[k];
#pragma unroll
        for (int j = 0; j < MAX_H; ++j)
            if (j < H) h[j] = fmaf(xk, s_w1[k * H + j], h[j]);
    }
    float v = b2[0];
#pragma unroll
    for (int j = 0; j < MAX_H; ++j)
        if (j < H) v = fmaf(fmaxf(h[j], 0.f), w2[j], v);
    values[e] = v;
}

// models.py:131-139, literally, in float64: curr = done ? 0 : rew + gamma * curr, backwards in t.
// (T,B) row-major: for a fixed t consecutive threads touch consecutive envs.
__global__ void discounted_returns_kernel(const float* __restrict__ rewards, const uint8_t* __restrict__ done,
                                          double gamma, int T, long long B, double* __restrict__ out) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double curr = 0.0;
    int t = T - 1;
    // the loads do not depend on the recurrence: fetch eight steps ahead of it
    for (; t >= 7; t -= 8) {
        float r[8]; uint8_t d[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const long long i = (long long)(t - u) * B + b;
            r[u] = rewards[i]; d[u] = done[i];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            curr = d[u] ? 0.0 : __dadd_rn((double)r[u], __dmul_rn(gamma, curr));
            out[(long long)(t - u) * B + b] = curr;
        }
    }
    for (; t >= 0; --t) {
        const long long i = (long long)t * B + b;
        curr = done[i] ? 0.0 : __dadd_rn((double)rewards[i], __dmul_rn(gamma, curr));
        out[i] = curr;
    }
}

}  // namespace mnr

namespace {
thread_local char g_err2[256] = "";
}

extern "C" {

const char* marlnav_rollout_last_error(void) { return g_err2; }

int marlnav_actor_sample_f32(const float* obs, long long N, int S, int H, const float* w1, const float* b1,
                             const float* w_mu, const float* b_mu, const float* w_std, const float* b_std,
                             const float* eps, uint64_t seed, uint64_t counter, const uint64_t* counter_dev,
                             float* actions, float* log_probs, float* mu_out, float* var_out, void* stream) {
    if (!obs || !w1 || !b1 || !w_mu || !b_mu || !w_std || !b_std || !actions || !log_probs || N < 1) {
        snprintf(g_err2, sizeof g_err2, "marlnav_actor_sample_f32: NULL pointer or empty batch");
        return MARLNAV_ERR_BAD_ARG;
    }
    if (S < 1 || S > 64 || H < 1 || H > 256) {
        snprintf(g_err2, sizeof g_err2, "marlnav_actor_sample_f32: need 1 <= obs_size <= 64 and 1 <= hidden <= 256");
        return MARLNAV_ERR_BAD_SHAPE;
    }
    const int threads = 128;
    const long long grid = (N + threads - 1) / threads;
    const size_t smem = (size_t)(H * S + H + 4 * H) * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    if (S <= 16)
        mnr::actor_sample_kernel<16, 256><<<(unsigned)grid, threads, smem, st>>>(
            obs, N, S, H, w1, b1, w_mu, b_mu, w_std, b_std, eps, seed, counter,
            reinterpret_cast<const unsigned long long*>(counter_dev), actions, log_probs, mu_out, var_out);
    else
        mnr::actor_sample_kernel<64, 256><<<(unsigned)grid, threads, smem, st>>>(
            obs, N, S, H, w1, b1, w_mu, b_mu, w_std, b_std, eps, seed, counter,
            reinterpret_cast<const unsigned long long*>(counter_dev), actions, log_probs, mu_out, var_out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(g_err2, sizeof g_err2, "actor_sample launch: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

int marlnav_critic_value_f32(const float* obs, long long B, int K, int H, const float* w1, const float* b1,
                             const float* w2, const float* b2, float* values, void* stream) {
    if (!obs || !w1 || !b1 || !w2 || !b2 || !values || B < 1) {
        snprintf(g_err2, sizeof g_err2, "marlnav_critic_value_f32: NULL pointer or empty batch");
        return MARLNAV_ERR_BAD_ARG;
    }
    if (K < 1 || H < 1 || H > 64 || (size_t)H * K * sizeof(float) > 200 * 1024) {
        snprintf(g_err2, sizeof g_err2, "marlnav_critic_value_f32: need hidden <= 64 and hidden*inputs*4 <= 200 KiB");
        return MARLNAV_ERR_BAD_SHAPE;
    }
    const int threads = 128;
    const size_t smem = (size_t)H * K * sizeof(float);
    static bool big_smem[64] = {false};
    int dev = 0; cudaGetDevice(&dev);
    if (smem > 48 * 1024 && !big_smem[dev & 63]) {
        cudaFuncSetAttribute(mnr::critic_value_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        big_smem[dev & 63] = true;
    }
    mnr::critic_value_kernel<64><<<(unsigned)((B + threads - 1) / threads), threads, smem, (cudaStream_t)stream>>>(
        obs, B, K, H, w1, b1, w2, b2, values);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(g_err2, sizeof g_err2, "critic_value launch: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

int marlnav_discounted_returns_f64(const float* rewards, const uint8_t* done, double gamma, int T, long long B,
                                   double* out, void* stream) {
    if (!rewards || !done || !out || T < 1 || B < 1) {
        snprintf(g_err2, sizeof g_err2, "marlnav_discounted_returns_f64: NULL pointer or empty buffer");
        return MARLNAV_ERR_BAD_ARG;
    }
    const int threads = 128;
    mnr::discounted_returns_kernel<<<(unsigned)((B + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
        rewards, done, gamma, T, B, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(g_err2, sizeof g_err2, "discounted_returns launch: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

}  // extern "C"
