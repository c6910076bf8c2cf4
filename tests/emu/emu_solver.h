// tests/emu/emu_solver.h -- TEST INFRASTRUCTURE: host stand-ins for the cuSOLVER entry points the library binds at
// run time (cusolverDnXgetrf / Xgetrs, 64-bit API: column-major, 1-based int64 pivots, LAPACK getrf/getrs semantics).
// They follow the documented behaviour of the real routines, so that the library's own calling conventions (row-major
// system handed over as the transpose, op(T) in the solve, leading dimensions, right-hand-side layout) are exercised.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

namespace emu {
inline int solverCreate(void** h) { *h = malloc(8); return 0; }
inline int solverDestroy(void* h) { free(h); return 0; }
inline int solverSetStream(void*, cudaStream_t) { return 0; }
inline int solverCreateParams(void** p) { *p = malloc(8); return 0; }
inline int solverDestroyParams(void* p) { free(p); return 0; }
inline int solverXgetrfBuf(void*, void*, int64_t m, int64_t n, int dtA, const void*, int64_t lda, int ct, size_t* dev, size_t* host) {
    if (dtA != 1 || ct != 1 || lda < m || m < 0 || n < 0) return 3;      // CUSOLVER_STATUS_INVALID_VALUE
    *dev = 256; *host = 0;
    return 0;
}
inline int solverXgetrf(void*, void*, int64_t m, int64_t n, int dtA, void* Av, int64_t lda, int64_t* ipiv, int ct, void* wd, size_t nwd,
                        void*, size_t, int* info) {
    if (dtA != 1 || ct != 1 || lda < m || !wd || nwd < 256) return 3;
    double* A = (double*)Av;
    *info = 0;
    const int64_t k = m < n ? m : n;
    for (int64_t j = 0; j < k; ++j) {
        int64_t p = j;
        for (int64_t i = j + 1; i < m; ++i) if (fabs(A[i + j * lda]) > fabs(A[p + j * lda])) p = i;
        ipiv[j] = p + 1;
        if (A[p + j * lda] == 0.0) { if (*info == 0) *info = (int)(j + 1); continue; }
        if (p != j) for (int64_t c = 0; c < n; ++c) { const double t = A[j + c * lda]; A[j + c * lda] = A[p + c * lda]; A[p + c * lda] = t; }
        const double inv = 1.0 / A[j + j * lda];
        for (int64_t i = j + 1; i < m; ++i) A[i + j * lda] *= inv;
        for (int64_t c = j + 1; c < n; ++c) {
            const double f = A[j + c * lda];
            if (f != 0.0) for (int64_t i = j + 1; i < m; ++i) A[i + c * lda] -= A[i + j * lda] * f;
        }
    }
    return 0;
}
// solves op(A) X = B with A = P L U from solverXgetrf; trans: 0 = N, 1 = T
inline int solverXgetrs(void*, void*, int trans, int64_t n, int64_t nrhs, int dtA, const void* Av, int64_t lda, const int64_t* ipiv, int dtB,
                        void* Bv, int64_t ldb, int* info) {
    if (dtA != 1 || dtB != 1 || lda < n || ldb < n || (trans != 0 && trans != 1)) return 3;
    const double* A = (const double*)Av;
    double* B = (double*)Bv;
    *info = 0;
    for (int64_t r = 0; r < nrhs; ++r) {
        double* b = B + r * ldb;
        if (trans == 0) {                       // A x = b:  x = U^-1 L^-1 P b
            for (int64_t i = 0; i < n; ++i) { const int64_t p = ipiv[i] - 1; if (p != i) { const double t = b[i]; b[i] = b[p]; b[p] = t; } }
            for (int64_t j = 0; j < n; ++j) for (int64_t i = j + 1; i < n; ++i) b[i] -= A[i + j * lda] * b[j];
            for (int64_t j = n - 1; j >= 0; --j) { b[j] /= A[j + j * lda]; for (int64_t i = 0; i < j; ++i) b[i] -= A[i + j * lda] * b[j]; }
        } else {                                // A^T x = b:  U^T L^T P x = b
            for (int64_t j = 0; j < n; ++j) { double s = b[j]; for (int64_t i = 0; i < j; ++i) s -= A[i + j * lda] * b[i]; b[j] = s / A[j + j * lda]; }
            for (int64_t j = n - 1; j >= 0; --j) { double s = b[j]; for (int64_t i = j + 1; i < n; ++i) s -= A[i + j * lda] * b[i]; b[j] = s; }
            for (int64_t i = n - 1; i >= 0; --i) { const int64_t p = ipiv[i] - 1; if (p != i) { const double t = b[i]; b[i] = b[p]; b[p] = t; } }
        }
    }
    return 0;
}
}  // namespace emu
