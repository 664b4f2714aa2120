// env.cu -- K1: vectorised CACC platoon environment (float64 state, float32 observations).
// Compiled with --fmad=false: the reference is NumPy float64 without FMA contraction, and the
// kernels below keep its operation order so h, v, u and the rewards are bit-identical.
//
// One thread owns one environment and sweeps its N vehicles in ascending order (the vehicle
// chain is a serial dependence: headway i needs the old AND new speed of vehicle i-1).  All
// arrays are [agent][env], so a warp's loads/stores are coalesced over envs.  The kernel is
// latency-bound and tiny next to the policy kernels (N*45 B of HBM traffic per env-step).
//
// Restates envs/cacc_env.py: step :191-242, reward :40-52, observation :54-65,
// OVM :360-385, reset :166-189 / :285-318.
#include "common.cuh"

namespace {

struct EnvK {
  nmarl_cacc_cfg c;
  int B;
};

__device__ __forceinline__ double leader_speed(const nmarl_cacc_cfg& c, double v_init, int t) {
  // v0s[t]: catch-up == v*; slow-down == np.linspace(v_init, v*, 300)[t] for t < 300 then v*
  if (c.scenario == NMARL_CATCHUP || t >= 299) return c.v_star;
  const double step = (c.v_star - v_init) / 299.0;
  return (double)t * step + v_init;
}

__device__ __forceinline__ double ovm_vh(const nmarl_cacc_cfg& c, double h) {
  if (h <= c.h_s) return 0.0;
  if (h < c.h_g) return c.v_max / 2 * (1 - cos(3.141592653589793 * (h - c.h_s) / (c.h_g - c.h_s)));
  return c.v_max;
}

__device__ __forceinline__ double clipd(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }

__device__ __forceinline__ void write_obs(const nmarl_cacc_cfg& c, int B, int b, int tcur, const double* hs,
                                          const double* vs, const double* us, const double* v_init, float* obs,
                                          int obs_stride) {
  const int L = c.platoon_len;
  double v_prev = 0.0;
  for (int i = 0; i < c.n_agent; ++i) {
    const int pos = i % L;
    const double v = vs[(size_t)i * B + b], h = hs[(size_t)i * B + b], u = us[(size_t)i * B + b];
    const double lead = pos ? v_prev : leader_speed(c, v_init[(size_t)(i / L) * B + b], tcur);
    float* o = obs + ((size_t)i * B + b) * obs_stride;
    o[0] = (float)((v - c.v_star) / c.v_star);
    o[1] = (float)clipd((lead - v) / 5.0, -2.0, 2.0);
    o[2] = (float)clipd((ovm_vh(c, h) - v) / 5.0, -2.0, 2.0);
    o[3] = (float)((h + (lead - v) * c.dt - c.h_star) / c.h_star);
    o[4] = (float)(u / c.u_max);
    v_prev = v;
  }
}

__global__ void cacc_reset_kernel(const EnvK k, const double* __restrict__ u01, const float* __restrict__ mask,
                                  uint64_t seed, int32_t* episode, double* hs, double* vs, double* us, int32_t* t,
                                  int32_t* collision, double* v_init, float* obs, int obs_stride, float* fp, int n_a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int B = k.B;
  if (b >= B) return;
  if (mask != nullptr && mask[b] == 0.0f) return;
  const nmarl_cacc_cfg& c = k.c;
  const int L = c.platoon_len, P = c.n_agent / L;
  uint32_t ep = 0;
  if (episode != nullptr) { ep = (uint32_t)episode[b]; episode[b] = (int32_t)(ep + 1); }
  for (int p = 0; p < P; ++p) {
    const double u = (u01 != nullptr) ? u01[(size_t)p * B + b] : philox_u01(seed, ((uint64_t)ep << 8) | (uint64_t)p, (uint32_t)b, 0x454e5601u);
    const double scale = 1.5 + u;
    v_init[(size_t)p * B + b] = (c.scenario == NMARL_SLOWDOWN) ? c.v_star * scale : c.v_star;
    for (int pos = 0; pos < L; ++pos) {
      const size_t o = (size_t)(p * L + pos) * B + b;
      hs[o] = (c.scenario == NMARL_CATCHUP && pos == 0) ? c.h_star * scale : c.h_star;
      vs[o] = (c.scenario == NMARL_SLOWDOWN) ? c.v_star * scale : c.v_star;
      us[o] = 0.0;
    }
  }
  t[b] = 0;
  collision[b] = 0;
  write_obs(c, B, b, 0, hs, vs, us, v_init, obs, obs_stride);
  if (fp != nullptr) {
    const float p0 = (float)(1.0 / (double)n_a);
    for (int i = 0; i < c.n_agent; ++i)
      for (int a = 0; a < n_a; ++a) fp[((size_t)i * B + b) * n_a + a] = p0;
  }
}

__global__ void cacc_step_kernel(const EnvK k, int train_mode, const int32_t* __restrict__ action, double* hs,
                                 double* vs, double* us, int32_t* t, int32_t* collision,
                                 const double* __restrict__ v_init, float* obs, int obs_stride, double* reward,
                                 double* greward, float* done) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int B = k.B;
  if (b >= B) return;
  const nmarl_cacc_cfg& c = k.c;
  const int N = c.n_agent, L = c.platoon_len;
  const int tcur = t[b];
  int col = collision[b];
  double gsum;
  if (col) {                       // frozen after a collision: -G for every agent (:193-194)
    for (int i = 0; i < N; ++i)
      if (!c.global_reward) reward[(size_t)i * B + b] = -c.G;
    gsum = -c.G * (double)N;       // sum of N equal values is exact in any order
  } else {
    double v_prev_old = 0.0, v_prev_new = 0.0, hmin = 1e300;
    for (int i = 0; i < N; ++i) {
      const int pos = i % L;
      const size_t o = (size_t)i * B + b;
      const double v = vs[o], h = hs[o];
      double lead, lead_next;
      if (pos) { lead = v_prev_old; lead_next = v_prev_new; }
      else {
        const double vi = v_init[(size_t)(i / L) * B + b];
        lead = leader_speed(c, vi, tcur);
        lead_next = leader_speed(c, vi, tcur + 1);
      }
      const int a = action[o];
      const double al = (a & 1) ? 0.5 : 0.0;          // a_map = [(0,0),(.5,0),(0,.5),(.5,.5)]  (:275)
      const double be = (a & 2) ? 0.5 : 0.0;
      const double u = al * (ovm_vh(c, h) - v) + be * (lead - v);
      double vn = v + clipd(u, c.u_min, c.u_max) * c.dt;
      vn = clipd(vn, 0.0, c.v_max);
      const double uc = (vn - v) / c.dt;
      const double hn = h + 0.5 * c.dt * (lead + lead_next - v - vn);
      hs[o] = hn; vs[o] = vn; us[o] = uc;
      v_prev_old = v; v_prev_new = vn;
      hmin = fmin(hmin, hn);
    }
    if (hmin < c.h_min) {          // collision latch (:42-44)
      col = 1;
      collision[b] = 1;
      for (int i = 0; i < N; ++i)
        if (!c.global_reward) reward[(size_t)i * B + b] = -c.G;
      gsum = -c.G * (double)N;
    } else {
      // np.sum order: sequential for N < 8, otherwise 8 strided accumulators + pairwise tree + tail
      double r8[8];
      double tail = 0.0;
      const int nblk = N - (N % 8);
      for (int i = 0; i < N; ++i) {
        const size_t o = (size_t)i * B + b;
        const double h = hs[o], v = vs[o], u = us[o];
        double r = -((h - c.h_star) * (h - c.h_star));
        r = r + (-c.rew_a * ((v - c.v_star) * (v - c.v_star)));
        r = r + (-c.rew_b * (u * u));
        if (train_mode) {
          const double m = fmin(h - 10.0, 0.0);
          r = r + (-5.0 * (m * m));
        } else {
          r = r + 0.0;
        }
        if (!c.global_reward) reward[o] = r;
        if (N < 8) tail += r;
        else if (i < 8) r8[i] = r;
        else if (i < nblk) r8[i & 7] += r;
        else { if (i == nblk) tail = ((r8[0] + r8[1]) + (r8[2] + r8[3])) + ((r8[4] + r8[5]) + (r8[6] + r8[7])); tail += r; }
      }
      if (N >= 8 && N == nblk) tail = ((r8[0] + r8[1]) + (r8[2] + r8[3])) + ((r8[4] + r8[5]) + (r8[6] + r8[7]));
      gsum = tail;
    }
  }
  const int tn = tcur + 1;
  t[b] = tn;
  greward[b] = gsum;
  if (c.global_reward) reward[b] = gsum;
  const bool d = (col && (tn % c.batch_size == 0)) || (tn == c.T);
  done[b] = d ? 1.0f : 0.0f;
  write_obs(c, B, b, tn, hs, vs, us, v_init, obs, obs_stride);
}

}  // namespace

extern "C" int nmarl_cacc_reset(const nmarl_cacc_cfg* cfg, int B, const double* u01, const float* mask, uint64_t seed,
                                int32_t* episode, double* hs, double* vs, double* us, int32_t* t, int32_t* collision,
                                double* v_init, float* obs, int obs_stride, float* fp, int n_a, void* stream) {
  NMARL_CHECK(cfg && B > 0, "cacc_reset: bad arguments");
  NMARL_CHECK(cfg->platoon_len > 0 && cfg->n_agent % cfg->platoon_len == 0, "cacc_reset: n_agent %% platoon_len != 0");
  NMARL_CHECK(obs_stride >= 5, "cacc_reset: obs_stride < 5");
  EnvK k{*cfg, B};
  const int nt = 64;
  cacc_reset_kernel<<<(B + nt - 1) / nt, nt, 0, (cudaStream_t)stream>>>(k, u01, mask, seed, episode, hs, vs, us, t,
                                                                         collision, v_init, obs, obs_stride, fp, n_a);
  NMARL_LAUNCH_CHECK();
  return 0;
}

extern "C" int nmarl_cacc_step(const nmarl_cacc_cfg* cfg, int B, int train_mode, const int32_t* action, double* hs,
                               double* vs, double* us, int32_t* t, int32_t* collision, const double* v_init, float* obs,
                               int obs_stride, double* reward, double* greward, float* done, void* stream) {
  NMARL_CHECK(cfg && B > 0 && action, "cacc_step: bad arguments");
  NMARL_CHECK(cfg->platoon_len > 0 && cfg->n_agent % cfg->platoon_len == 0, "cacc_step: n_agent %% platoon_len != 0");
  EnvK k{*cfg, B};
  const int nt = 64;
  cacc_step_kernel<<<(B + nt - 1) / nt, nt, 0, (cudaStream_t)stream>>>(k, train_mode, action, hs, vs, us, t, collision,
                                                                        v_init, obs, obs_stride, reward, greward, done);
  NMARL_LAUNCH_CHECK();
  return 0;
}
