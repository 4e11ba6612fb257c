/* lapack_stub.c -- stand-in for LAPACK's sgels_ so that the reference's UNMODIFIED epic_aux.cpp links in this image
 * (no LAPACK / BLAS is installed; epic_aux.cpp:11-17, 455-461).  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED at this call: the reference leaves the least-squares solve to whatever LAPACK the host provides.
 * Only the call shape fit_localaffine uses is implemented: TRANS = 'T', M = 6 < N, NRHS = 1, i.e. the overdetermined
 * system A^T x = b with A (column-major, LDA = 6) holding one equation per COLUMN.  Like sgels it returns the minimum
 * residual solution in b[0..M-1]; it is computed by Householder QR of A^T in single precision (what sgels does up to
 * the blocking), so it agrees with any LAPACK to rounding.  A workspace query (LWORK = -1) returns 1. */
#include <math.h>
#include <stdlib.h>

int sgels_(char *trans, int *m, int *n, int *nrhs, float *a, int *lda, float *b, int *ldb, float *work, int *lwork, int *info) {
    (void)ldb;
    *info = 0;
    if (*lwork == -1) {
        work[0] = 1.0f;
        return 0;
    }
    if (!(trans[0] == 'T' || trans[0] == 't') || *nrhs != 1 || *m > *n) {
        *info = -1;
        return 0;
    }
    const int rows = *n, cols = *m, ld = *lda;
    /* Q^T applied in place: column k of A^T is a[k + r*ld], r = 0..rows-1 */
    for (int k = 0; k < cols; k++) {
        float norm = 0.0f;
        for (int r = k; r < rows; r++) norm += a[k + r * ld] * a[k + r * ld];
        norm = sqrtf(norm);
        if (norm == 0.0f) continue;
        const float akk = a[k + k * ld];
        const float alpha = akk > 0.0f ? -norm : norm;
        /* v = x - alpha e1, stored over the column; beta = 2 / (v^T v) */
        a[k + k * ld] = akk - alpha;
        float vtv = 0.0f;
        for (int r = k; r < rows; r++) vtv += a[k + r * ld] * a[k + r * ld];
        const float beta = 2.0f / vtv;
        for (int c = k + 1; c < cols; c++) {
            float s = 0.0f;
            for (int r = k; r < rows; r++) s += a[k + r * ld] * a[c + r * ld];
            s *= beta;
            for (int r = k; r < rows; r++) a[c + r * ld] -= s * a[k + r * ld];
        }
        {
            float s = 0.0f;
            for (int r = k; r < rows; r++) s += a[k + r * ld] * b[r];
            s *= beta;
            for (int r = k; r < rows; r++) b[r] -= s * a[k + r * ld];
        }
        /* R's diagonal entry replaces the head of the reflector; keep v's tail (unused afterwards) */
        a[k + k * ld] = alpha;
        for (int r = k + 1; r < rows; r++) a[k + r * ld] = 0.0f;
    }
    /* back substitution with the upper triangle R (R[r][c] = a[c + r*ld], r <= c) */
    for (int r = cols - 1; r >= 0; r--) {
        float s = b[r];
        for (int c = r + 1; c < cols; c++) s -= a[c + r * ld] * b[c];
        const float d = a[r + r * ld];
        if (d == 0.0f) {
            *info = r + 1;
            return 0;
        }
        b[r] = s / d;
    }
    return 0;
}
