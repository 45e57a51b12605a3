// msv_launch.h -- host-side interface between the C ABI (msv_abi.cu) and the
// kernel launchers (msv_kernels.cu).
#pragma once
#include "msv_types.cuh"
#ifndef MSV_TPB
#define MSV_TPB 512   // maximum threads per block of k_step / k_reset / k_observe (the actual size is a runtime choice, see plan_blocks)
#endif
// cap: capacity class 0 = <2,4,4>, 1 = <4,4,4>, 2 = <8,8,16>; which: 0 step, 1 reset, 2 observe, 3 one-time kernel attribute setup, 4 reset only the envs whose done flag is set
cudaError_t msv_launch(int cap, int which, const DevConst& C, const DevState& S, const DevOut& O,
                       const uint8_t* actions, cudaStream_t st);
void msv_capacity(int cap, int* AC, int* BC, int* HC, int* P, int* PW, int* sm_words);
cudaError_t msv_launch_stats(int N, int stride, int AC, float* sreward, int* skills, int4* smisc, double* out_reward,
                             unsigned long long* out_kills, unsigned long long* out_misc, cudaStream_t st);
cudaError_t msv_read_profile(unsigned long long out[64], int reset);
cudaError_t msv_read_blocks(unsigned long long* out, int n_words);
cudaError_t msv_read_trace(unsigned long long* out, int n_words);
cudaError_t msv_read_check(unsigned long long out[2]);
cudaError_t msv_launch_spare(const DevConst& C, const DevState& S, const uint8_t* dones, int only_done, cudaStream_t st);
cudaError_t msv_launch_obs(const DevConst& C, const DevState& S, const ObsTable& T, int AC, const uint8_t* only_if, cudaStream_t st);
cudaError_t msv_launch_lidar(const DevConst& C, const DevState& S, const DevOut& O, int BC, int HC, cudaStream_t st);
