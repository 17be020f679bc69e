// tcgen05 / TMEM / TMA GEMMs -- placeholder until the kernels land; every shape falls back to gemm_mma.cu.
#include "gemm.cuh"
int gemm_nt_tc05(const GemmArgs&, int, cudaStream_t) { return TULIP_ERR_UNSUPPORTED; }
int gemm_tn_tc05(const GemmTNArgs&, cudaStream_t) { return TULIP_ERR_UNSUPPORTED; }
