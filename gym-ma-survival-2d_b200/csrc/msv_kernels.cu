// msv_kernels.cu -- the sm_100a kernels of libmasurv.so and their launchers.
//
// k_step  : one full MaSurvival.step (env:76-90) for every environment:
//           queue_actions -> pre_step hooks -> b2World::Step x2 -> post_step
//           hooks -> observations -> rewards -> done -> stats (-> auto-reset).
// k_reset : BaseEnv.reset (env:59-74) for every environment.
// k_observe: fetch_observations only (after msv_set_state).
// k_stats : flush_stats reduction.
#include "msv_env.cuh"
#include "msv_launch.h"

// debug phase profile (enabled by msv_debug_profile): sum over threads of the
// clock64() cycles spent in each phase of k_step
__device__ unsigned long long g_prof[16];
#define PROF(k) do { if (C.profile) { long long _t = clock64(); atomicAdd(&g_prof[k], (unsigned long long)(_t - t_last)); t_last = _t; } } while (0)

template <int AC, int BC, int HC>
__global__ void __launch_bounds__(MSV_TPB)
k_step(const __grid_constant__ DevConst C, const __grid_constant__ DevState S,
       const __grid_constant__ DevOut Oc, const uint8_t* __restrict__ actions) {
  extern __shared__ float sm[];
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= C.N) return;
  DevOut O = Oc;
  Env<AC, BC, HC> env(C, S, sm, blockDim.x, threadIdx.x, e);
  long long t_last = C.profile ? clock64() : 0;
  env.load();
  uint8_t act[AC * 6];
  {
    const uint8_t* src = actions + (size_t)e * C.A * 6;
    for (int k = 0; k < AC * 6; ++k) act[k] = k < C.A * 6 ? src[k] : 0;
  }
  PROF(0);
  env.pre_step(act);                       // sim:234-235
  PROF(1);
  for (int sub = 0; sub < 2; ++sub) {      // sim:236-239: b2World::Step x2
    if (sub == 0) env.find_new_contacts(); // newFixture (boxes placed in pre_step)
    PROF(2);
    env.collide();
    PROF(3);
    env.solve(C.dt, env.first_step ? 0.0f : C.dt_ratio1);
    PROF(4);
    env.solve_toi(C.dt);
    env.first_step = 0;
    PROF(5);
  }
  env.post_step_boxes();
  PROF(6);
  env.cameras();
  if (C.lidar_n > 0) env.lidar(O);
  PROF(7);
  env.post_step_rest();                    // sim:241-242
  PROF(8);
  env.observe(O);                          // env:84
  PROF(9);
  bool done = env.rewards_done(O);         // env:85-89
  if (done && C.auto_reset) {              // vector-env extension
    env.st_episodes++;
    env.reset();
    env.cameras();
    if (C.lidar_n > 0) env.lidar(O);
    env.observe(O);
  }
  PROF(10);
  env.store();
  PROF(11);
}

template <int AC, int BC, int HC>
__global__ void __launch_bounds__(MSV_TPB)
k_reset(const __grid_constant__ DevConst C, const __grid_constant__ DevState S,
        const __grid_constant__ DevOut Oc) {
  extern __shared__ float sm[];
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= C.N) return;
  DevOut O = Oc;
  Env<AC, BC, HC> env(C, S, sm, blockDim.x, threadIdx.x, e);
  env.load();
  env.reset();
  env.cameras();
  if (C.lidar_n > 0) env.lidar(O);
  env.observe(O);
  for (int i = 0; i < C.A; ++i) O.rewards[(size_t)e * C.A + i] = 0.0f;
  O.dones[e] = 0;
  env.store();
}

template <int AC, int BC, int HC>
__global__ void __launch_bounds__(MSV_TPB)
k_observe(const __grid_constant__ DevConst C, const __grid_constant__ DevState S,
          const __grid_constant__ DevOut Oc) {
  extern __shared__ float sm[];
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= C.N) return;
  DevOut O = Oc;
  Env<AC, BC, HC> env(C, S, sm, blockDim.x, threadIdx.x, e);
  env.load();
  env.cameras();
  if (C.lidar_n > 0) env.lidar(O);
  env.observe(O);
}

// flush_stats (env:471-480): sum the per-env accumulators, then zero them
__global__ void k_stats(int N, int AC, float* sreward, int* skills, int4* smisc,
                        double* out_reward, unsigned long long* out_kills, unsigned long long* out_misc) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  double r[MSV_MAX_AGENTS]; long long k[MSV_MAX_AGENTS]; long long m[4] = {0, 0, 0, 0};
  for (int i = 0; i < MSV_MAX_AGENTS; ++i) { r[i] = 0.0; k[i] = 0; }
  if (e < N) {
    for (int i = 0; i < AC; ++i) {
      r[i] = sreward[i * N + e]; k[i] = skills[i * N + e];
      sreward[i * N + e] = 0.0f; skills[i * N + e] = 0;
    }
    int4 s = smisc[e]; m[0] = s.x; m[1] = s.y; m[2] = s.z; m[3] = s.w;
    smisc[e] = make_int4(0, 0, 0, 0);
  }
  for (int off = 16; off > 0; off >>= 1) {
    for (int i = 0; i < AC; ++i) { r[i] += __shfl_down_sync(0xffffffffu, r[i], off); k[i] += __shfl_down_sync(0xffffffffu, k[i], off); }
    for (int i = 0; i < 4; ++i) m[i] += __shfl_down_sync(0xffffffffu, m[i], off);
  }
  if ((threadIdx.x & 31) == 0) {
    for (int i = 0; i < AC; ++i) { atomicAdd(&out_reward[i], r[i]); atomicAdd(&out_kills[i], (unsigned long long)k[i]); }
    for (int i = 0; i < 4; ++i) atomicAdd(&out_misc[i], (unsigned long long)m[i]);
  }
}

// ------------------------------------------------------------- launchers ---
template <int AC, int BC, int HC>
static cudaError_t launch_t(int which, const DevConst& C, const DevState& S, const DevOut& O,
                            const uint8_t* actions, cudaStream_t st) {
  int blocks = (C.N + MSV_TPB - 1) / MSV_TPB;
  size_t smem = (size_t)Env<AC, BC, HC>::SM_WORDS * MSV_TPB * sizeof(float);
  if (which == 0) {
    cudaFuncSetAttribute(k_step<AC, BC, HC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_step<AC, BC, HC><<<blocks, MSV_TPB, smem, st>>>(C, S, O, actions);
  } else if (which == 1) {
    cudaFuncSetAttribute(k_reset<AC, BC, HC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_reset<AC, BC, HC><<<blocks, MSV_TPB, smem, st>>>(C, S, O);
  } else {
    cudaFuncSetAttribute(k_observe<AC, BC, HC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_observe<AC, BC, HC><<<blocks, MSV_TPB, smem, st>>>(C, S, O);
  }
  return cudaPeekAtLastError();
}

cudaError_t msv_launch(int cap, int which, const DevConst& C, const DevState& S, const DevOut& O,
                       const uint8_t* actions, cudaStream_t st) {
  switch (cap) {
    case 0: return launch_t<2, 4, 4>(which, C, S, O, actions, st);
    case 1: return launch_t<4, 4, 4>(which, C, S, O, actions, st);
    default: return launch_t<8, 8, 16>(which, C, S, O, actions, st);
  }
}

void msv_capacity(int cap, int* AC, int* BC, int* HC, int* P, int* PW) {
  switch (cap) {
    case 0: *AC = 2; *BC = 4; *HC = 4; *P = PairLayout<2, 4>::P; *PW = PairLayout<2, 4>::PW; break;
    case 1: *AC = 4; *BC = 4; *HC = 4; *P = PairLayout<4, 4>::P; *PW = PairLayout<4, 4>::PW; break;
    default: *AC = 8; *BC = 8; *HC = 16; *P = PairLayout<8, 8>::P; *PW = PairLayout<8, 8>::PW; break;
  }
}

cudaError_t msv_launch_stats(int N, int AC, float* sreward, int* skills, int4* smisc, double* out_reward,
                             unsigned long long* out_kills, unsigned long long* out_misc, cudaStream_t st) {
  k_stats<<<(N + 127) / 128, 128, 0, st>>>(N, AC, sreward, skills, smisc, out_reward, out_kills, out_misc);
  return cudaPeekAtLastError();
}

cudaError_t msv_read_profile(unsigned long long out[16], int reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out, g_prof, sizeof(unsigned long long) * 16);
  if (e != cudaSuccess) return e;
  if (reset) { unsigned long long z[16] = {0}; e = cudaMemcpyToSymbol(g_prof, z, sizeof z); }
  return e;
}
