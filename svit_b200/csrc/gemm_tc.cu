// tcgen05 GEMM (bf16) -- placeholder until the TMEM kernel lands.
#include "common.cuh"
#include "../../include/svit_b200.h"
int svit_gemm_tc_supported(const svit_gemm_args* a) { (void)a; return 0; }
int svit_gemm_tc(const svit_gemm_args* a, cudaStream_t st) { (void)a; (void)st; return SVIT_ENOTSUP; }
extern "C" int svit_destroy(void) { return 0; }
