// bf_kernels.h -- host-visible launchers of the fold kernels (bf_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "bf_params.h"

struct BfBatchDev;

cudaError_t bf_upload_constants();
size_t bf_mfe_slot_ints(int wstride);
size_t bf_pf_slot_doubles(int wstride);
int bf_occupancy_mfe(bool two, int wstride);
int bf_occupancy_pf(bool two, int wstride);

cudaError_t bf_launch_mfe(const BfParams *dP, const BfBatchDev &b, bool two, int *ws, int wstride, int grid, int *work_counter,
                          int *out_mfe, char *out_ss, int ss_stride, cudaStream_t st);
cudaError_t bf_launch_pf(const BfParams *dP, const BfBatchDev &b, bool two, double *ws, int wstride, int grid, int *work_counter,
                         const int *mfe_for_scale, double *out5, cudaStream_t st);
cudaError_t bf_launch_eval(const BfParams *dP, const BfBatchDev &b, const char *targets, int n_targets, int tstride, int *out_e,
                           cudaStream_t st);
