// bf16 tensor-core GraspPointCNN path (placeholder until the implicit-GEMM kernels land).
#include "lg_internal.cuh"

int lg_cnn_prepare_bf16(lg_context* c) { (void)c; return LG_OK; }

int lg_run_cnn_bf16(lg_context* c, const float* patches, int n, float* logits, cudaStream_t st) {
    (void)c; (void)patches; (void)n; (void)logits; (void)st;
    lg_set_error("bf16 CNN path is not built");
    return LG_E_ARG;
}
