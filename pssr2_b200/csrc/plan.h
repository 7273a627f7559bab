// Internal plan structures shared by the op implementations (family 2).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <vector>
#include "../../include/pssr_b200.h"

namespace pssr {

// tensor-map table of one convolution: sources, 16-bit weights, 16-bit output, e5m2 weights, spare
static constexpr int kTmapsPerConv = 8;
static constexpr int kTmW = 4, kTmOut = 5, kTmW8 = 6, kTmOutLo = 7;

struct ConvOp {
  CUtensorMap tmaps[kTmapsPerConv];   // host copies ([0..3] sources, [kTmW] weights, [kTmOut] output, [kTmW8] e5m2 weights, [kTmOutLo] second output); uploaded into the plan's device table
  alignas(16) uint8_t kparams[640];
  int grid = 0;
  int smem_bytes = 0;
  int variant = 1;                // 1 = conv_igemm.cu (box per tap), 3 = conv_v3.cu (lean issue)
  int kernel_index = -1;          // variant 3: index of the template instantiation
  int cluster = 1;                // variant 3: CTAs per cluster (2 = cta_group::2 pairs)
};

int conv_prepare(const pssr_conv_desc_t& d, int dtype, ConvOp& op);
int conv_launch(const ConvOp& op, const void* tmaps_dev, cudaStream_t stream);
bool v3_supported(const pssr_conv_desc_t& d);
int v3_prepare(const pssr_conv_desc_t& d, int dtype, ConvOp& op);
int v3_trace_fetch(long long* host, int n);
int v3_launch(const ConvOp& op, const void* tmaps_dev, cudaStream_t stream);

int prep_launch(const pssr_prep_desc_t& d, int dtype, cudaStream_t stream);
int pool_launch(const pssr_pool_desc_t& d, int dtype, cudaStream_t stream);
int tail_launch(const pssr_tail_desc_t& d, int dtype, cudaStream_t stream);
int tailsum_launch(const pssr_tailsum_desc_t& d, cudaStream_t stream);
int stem_launch(const pssr_stem_desc_t& d, int dtype, cudaStream_t stream);
int ln_launch(const pssr_ln_desc_t& d, int dtype, cudaStream_t stream);
int dwln_launch(const pssr_dwln_desc_t& d, int dtype, cudaStream_t stream);
int ese_launch(const pssr_ese_desc_t& d, int dtype, cudaStream_t stream);
int cast8_launch(const pssr_cast8_desc_t& d, int dtype, cudaStream_t stream);
int resample_launch(const pssr_resample_desc_t& d, int dtype, cudaStream_t stream);
int winattn_launch(const pssr_winattn_desc_t& d, int dtype, cudaStream_t stream);

}  // namespace pssr

struct pssr_plan {
  int dtype = 0;
  std::vector<pssr_op_t> ops;
  std::vector<pssr::ConvOp> convs;     // one per PSSR_OP_CONV, in op order
  std::vector<int> conv_index;          // op index -> index into convs (or -1)
  void* tmaps_dev = nullptr;            // device table: 4 CUtensorMap per conv op
};
