// tcgen05 pooled-attention kernel (bf16) -- placeholder until the TMEM kernel lands.
#include "common.cuh"
#include "../../include/svit_b200.h"
int svit_attn_tc_supported(const svit_attn_args* a) { (void)a; return 0; }
int svit_attn_fwd_tc(const svit_attn_args* a, cudaStream_t st) { (void)a; (void)st; return SVIT_ENOTSUP; }
