// Tensor-core path of the HEI tower layers (hei_tc.cu), selected by aread_hei_layer_fwd / _bwd (hei.cu) when the
// shape packs into 64-column blocks.  AREAD_HEI_TC=0 keeps the CUDA-core kernels.
#pragma once

#include "common.cuh"

namespace aread {

bool hei_tc_usable(int64_t m, int groups, int k, int n, const float* src, int64_t ld_src);   // forward
bool hei_tc_bwd_usable(int64_t m, int groups, int k, int n, const float* src, int64_t ld_src);
void hei_tc_set_path(int fwd, int bwd);     // 1 / 0 / -1 = environment default
bool hei_tc_shape_ok(int64_t m, int groups, int k, int n);
size_t hei_tc_workspace_floats(int64_t m, int groups, int k, int n);
int hei_tc_fwd(const aread_hei_layer_fwd_args& a, cudaStream_t stream);
int hei_tc_bwd(const aread_hei_layer_bwd_args& a, cudaStream_t stream);

}  // namespace aread
