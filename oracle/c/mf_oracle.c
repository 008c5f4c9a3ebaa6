/* ORACLE (test infrastructure, not product code): plain-C restatement of the reference's
 * sequential per-sample MF epoch, src/mf.py:97-108 and :154-216.
 *
 * Order of operations per sample (reference line):
 *   err   = y/ps - sigmoid(P_u.Q_i + bu_u + bi_i + b)      mf.py:99, 165-170, base.py:63-66
 *   P_u  -= lr * (-err*Q_i + reg*P_u)                       mf.py:181-182
 *   Q_i  -= lr * (-err*P_u(new) + reg*Q_i)                  mf.py:193-194
 *   bu_u -= lr * (-err + reg*bu_u)                          mf.py:204-205
 *   bi_i -= lr * (-err + reg*bi_i)                          mf.py:215-216
 * The dot product is accumulated left to right; NumPy's BLAS ddot may use a different
 * association, so agreement with the reference is to rounding (1e-15), not bit-for-bit.
 */
#include <math.h>
#include <stdint.h>

void oracle_mf_epoch(const int64_t *pairs, const double *y, const double *ps, int64_t n, int k,
                     double *P, double *Q, double *bu, double *bi, double b, double lr, double reg)
{
    for (int64_t s = 0; s < n; ++s) {
        int64_t u = pairs[2 * s], i = pairs[2 * s + 1];
        double *p = P + u * k, *q = Q + i * k;
        double z = 0.0;
        for (int f = 0; f < k; ++f) z += p[f] * q[f];
        z += bu[u];
        z += bi[i];
        z += b;
        if (z > 700.0) z = 700.0;
        if (z < -700.0) z = -700.0;
        double err = y[s] / ps[s] - 1.0 / (1.0 + exp(-z));
        for (int f = 0; f < k; ++f) p[f] -= lr * (-err * q[f] + reg * p[f]);
        for (int f = 0; f < k; ++f) q[f] -= lr * (-err * p[f] + reg * q[f]);
        bu[u] -= lr * (-err + reg * bu[u]);
        bi[i] -= lr * (-err + reg * bi[i]);
    }
}
