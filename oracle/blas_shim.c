/* TEST INFRASTRUCTURE ONLY.
 * Fallback BLAS for building oracle/_ref/libfaiss_ref.so when no OpenBLAS is found in the image.
 * Only sgemm_ is reached on the hot path (reference call sites utils.cpp:817,869,1132,1346 and
 * ProductQuantizer.cpp:485); the LAPACK entry points referenced by VectorTransform.cpp / utils.cpp
 * (out of scope) are abort-stubs so the library links.
 * Fortran column-major semantics: C(m,n) = alpha * op(A)(m,k) * op(B)(k,n) + beta * C.
 */
#include <stdio.h>
#include <stdlib.h>

static int is_t(const char* s) { return s[0] == 'T' || s[0] == 't' || s[0] == 'C' || s[0] == 'c'; }

int sgemm_(const char* transa, const char* transb, const int* pm, const int* pn, const int* pk, const float* palpha,
           const float* a, const int* plda, const float* b, const int* pldb, const float* pbeta, float* c,
           const int* pldc) {
  const int m = *pm, n = *pn, k = *pk, lda = *plda, ldb = *pldb, ldc = *pldc;
  const float alpha = *palpha, beta = *pbeta;
  const int ta = is_t(transa), tb = is_t(transb);
#pragma omp parallel for schedule(static)
  for (int j = 0; j < n; j++) {
    for (int i = 0; i < m; i++) {
      float acc = 0.f;
      if (ta && !tb) { /* dot of two contiguous columns: the hot-path case */
        const float* ai = a + (size_t)i * lda;
        const float* bj = b + (size_t)j * ldb;
        for (int l = 0; l < k; l++) acc += ai[l] * bj[l];
      } else {
        for (int l = 0; l < k; l++) {
          float av = ta ? a[(size_t)i * lda + l] : a[(size_t)l * lda + i];
          float bv = tb ? b[(size_t)l * ldb + j] : b[(size_t)j * ldb + l];
          acc += av * bv;
        }
      }
      float* cij = c + (size_t)j * ldc + i;
      *cij = (beta == 0.f) ? alpha * acc : alpha * acc + beta * (*cij);
    }
  }
  return 0;
}

#define STUB(name)                                                        \
  int name() {                                                            \
    fprintf(stderr, "blas_shim: " #name " is not implemented (stub)\n"); \
    abort();                                                              \
    return 0;                                                             \
  }
STUB(sgeqrf_)
STUB(sorgqr_)
STUB(ssyrk_)
STUB(ssyev_)
STUB(sgesvd_)
STUB(dgesvd_)
