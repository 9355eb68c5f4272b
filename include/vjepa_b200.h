/* vjepa_b200.h - C ABI of the B200-native V-JEPA pre-training hot path (libvjepa_b200.so).
 *
 * The reference (facebookresearch/jepa) is pure Python: it has no FFI layer, its hot path is the
 * set of torch library calls listed in SURVEY.md section 2.3 (K1..K15).  Every entry point below
 * replaces one of those call sites; the comment on each declaration cites the reference file:line
 * it stands in for.  Conventions:
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise;
 *     the caller owns every buffer; nothing is retained past the call;
 *   - `stream` is a cudaStream_t passed as void*; every call only ENQUEUES work on it;
 *   - return 0 = enqueued, <0 = argument / shape / alignment violation (nothing launched),
 *     >0 = cudaError_t; vj_last_error_string() gives the thread-local detail;
 *   - bf16 activations are row-major [tokens, features]; "ld" is a row stride in ELEMENTS.
 */
#ifndef VJEPA_B200_H_
#define VJEPA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VJ_VERSION 100

/* GEMM epilogues */
#define VJ_EPI_NONE 0  /* D = alpha*acc + bias                                             */
#define VJ_EPI_GELU 1  /* D = gelu_erf(alpha*acc + bias); aux_out (bf16, optional) = pre-activation */
#define VJ_EPI_ADD 2   /* D = alpha*acc + bias + aux[rowmap(r), c]   (residual / pos-embed) */
#define VJ_EPI_DGELU 3 /* D = (alpha*acc + bias) * gelu_erf'(aux[r, c])                     */

const char* vj_last_error_string(void);
int vj_version(void);

/* D[M,N] = epi(alpha * A[M,K] . B[N,K]^T), bf16 operands, fp32 accumulate (tcgen05 / TMEM).
 * a_mn = 0: A stored [M,K] (K contiguous, ld = lda);  a_mn = 1: A stored [K,M] (M contiguous).
 * b_mn = 0: B stored [N,K];                            b_mn = 1: B stored [K,N].
 * Supported (a_mn,b_mn): (0,0) forward / nn.Linear, (0,1) dgrad, (1,1) wgrad.
 * d_f32: D is fp32 (else bf16).  accumulate!=0 or split_k>1 reduce-add into fp32 D.
 * bias: fp32 [N] or NULL.  aux: bf16 or fp32 (aux_f32) tile source for ADD / DGELU; row r reads
 * aux row aux_rowmap[r] if given, else r % aux_period if aux_period > 0, else r.
 * Replaces F.linear / Conv3d-as-GEMM and their backward:
 *   src/models/utils/modules.py:31-34,63,76; src/models/predictor.py:194,237;
 *   src/models/utils/patch_embed.py:54-57. */
int vj_gemm(const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn,
            void* D, long long ldd, int d_f32, int M, int N, int K, const float* bias, float alpha,
            int epi, const void* aux, long long ldaux, int aux_f32, const int* aux_rowmap,
            int aux_period, void* aux_out, long long ldauxout, int split_k, int accumulate,
            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VJEPA_B200_H_ */
