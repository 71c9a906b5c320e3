#pragma once
// development aid (-DMJB_STAGE_PROF, device only): cycles per sub-stage, summed and maximised over environments
#if defined(MJB_STAGE_PROF) && defined(__CUDACC__)
static __device__ unsigned long long g_stage_prof[64];
#endif
#if defined(MJB_STAGE_PROF) && defined(__CUDA_ARCH__)
#define MJB_PROF_BEGIN long long pt_ = clock64();
#define MJB_PROF_MARK(id) { const long long t_ = clock64(); if (d.lane == 0) { atomicAdd(&g_stage_prof[2 * (id)], (unsigned long long)(t_ - pt_)); atomicMax(&g_stage_prof[2 * (id) + 1], (unsigned long long)(t_ - pt_)); } pt_ = clock64(); }
#else
#define MJB_PROF_BEGIN
#define MJB_PROF_MARK(id)
#endif

