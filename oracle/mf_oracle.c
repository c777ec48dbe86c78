/* oracle/mf_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See mf_oracle.h.
 *
 * A scalar (one element at a time, SIMD width 1) restatement of the reference
 * algorithms.  Floating-point operations are issued in the same order as the
 * reference's scalar build (tinysimd width 1, `fma` = unfused a += b*c,
 * LibUtilities/SimdLib/scalar.hpp:150-155) and the file is compiled with
 * -ffp-contract=off, so results agree with oracle/_ref/libnekref_scalar.so to the
 * last bit for the sum-factorisation kernels (checked in tests/).
 */
#include "mf_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int g_threads = 1;
int mfo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void mfo_set_threads(int n) { g_threads = n > 0 ? n : 1; }

/* ======================================================================== Polylib */

/* Jacobi polynomial P_n^{alpha,beta} and/or derivative at np points.
 * Three-term recurrence + derivative relation: Polylib.cpp:1034-1119. */
void mfo_jacobfd(int np, const double *z, double *poly_in, double *polyd, int n, double alpha, double beta)
{
    int i, k;
    if (np <= 0) return;
    if (n == 0)
    {
        if (poly_in) for (i = 0; i < np; ++i) poly_in[i] = 1.0;
        if (polyd) for (i = 0; i < np; ++i) polyd[i] = 0.0;
        return;
    }
    if (n == 1)
    {
        if (poly_in) for (i = 0; i < np; ++i) poly_in[i] = 0.5 * (alpha - beta + (alpha + beta + 2.0) * z[i]);
        if (polyd) for (i = 0; i < np; ++i) polyd[i] = 0.5 * (alpha + beta + 2.0);
        return;
    }
    {
        const double apb = alpha + beta;
        double *pn1      = (double *)malloc(3 * (size_t)np * sizeof(double));
        double *pn2      = pn1 + np;
        double *poly     = poly_in ? poly_in : pn2 + np;
        double a1, a2, a3, a4;
        for (i = 0; i < np; ++i)
        {
            pn2[i] = 1.0;
            pn1[i] = 0.5 * (alpha - beta + (alpha + beta + 2.0) * z[i]);
        }
        for (k = 2; k <= n; ++k)
        {
            a1 = 2.0 * k * (k + apb) * (2.0 * k + apb - 2.0);
            a2 = (2.0 * k + apb - 1.0) * (alpha * alpha - beta * beta);
            a3 = (2.0 * k + apb - 2.0) * (2.0 * k + apb - 1.0) * (2.0 * k + apb);
            a4 = 2.0 * (k + alpha - 1.0) * (k + beta - 1.0) * (2.0 * k + apb);
            a2 /= a1;
            a3 /= a1;
            a4 /= a1;
            for (i = 0; i < np; ++i)
            {
                poly[i] = (a2 + a3 * z[i]) * pn1[i] - a4 * pn2[i];
                pn2[i]  = pn1[i];
                pn1[i]  = poly[i];
            }
        }
        if (polyd)
        {
            a1 = n * (alpha - beta);
            a2 = n * (2.0 * n + alpha + beta);
            a3 = 2.0 * (n + alpha) * (n + beta);
            a4 = (2.0 * n + alpha + beta);
            a1 /= a4;
            a2 /= a4;
            a3 /= a4;
            /* pn2 holds P_{n-1} after the loop */
            for (i = 0; i < np; ++i)
            {
                polyd[i] = (a1 - a2 * z[i]) * poly[i] + a3 * pn2[i];
                polyd[i] /= (1.0 - z[i] * z[i]);
            }
        }
        free(pn1);
    }
}

/* d/dz P_n^{a,b} = (a+b+n+1)/2 P_{n-1}^{a+1,b+1}: Polylib.cpp:1134-1147 */
static void jacobd(int np, const double *z, double *polyd, int n, double alpha, double beta)
{
    int i;
    if (n == 0)
    {
        for (i = 0; i < np; ++i) polyd[i] = 0.0;
        return;
    }
    mfo_jacobfd(np, z, polyd, NULL, n - 1, alpha + 1.0, beta + 1.0);
    for (i = 0; i < np; ++i) polyd[i] *= 0.5 * (alpha + beta + (double)n + 1.0);
}

/* Gamma for integers and half-integers: Polylib.cpp:1161-1189 */
static double gammaF(double x)
{
    double g = 1.0;
    if (x == -0.5) return -2.0 * sqrt(M_PI);
    if (x == 0.0) return g;
    if ((x - (int)x) == 0.5)
    {
        int n      = (int)x;
        double tmp = x;
        g          = sqrt(M_PI);
        while (n--)
        {
            tmp -= 1.0;
            g *= tmp;
        }
    }
    else if ((x - (int)x) == 0.0)
    {
        int n      = (int)x;
        double tmp = x;
        while (--n)
        {
            tmp -= 1.0;
            g *= tmp;
        }
    }
    return g;
}

/* Gamma(x+alpha)/Gamma(y+beta), integer alpha/beta branch only (all the points
 * types on this path have integer alpha, beta): Polylib.cpp:1203-1223 */
static double gammaFrac(int x, double alpha, int y, double beta)
{
    double g = 1.0;
    int X = (int)(x + alpha), Y = (int)(y + beta), t;
    if (X > Y)
        for (t = X - 1; t > Y - 1; --t) g *= t;
    else if (Y > X)
    {
        for (t = Y - 1; t > X - 1; --t) g *= t;
        g = 1.0 / g;
    }
    return g;
}

/* zeros of P_n^{alpha,beta}: Newton iteration with deflation, Polylib.cpp:1281-1310 */
static void jacobz(int n, double *z, double alpha, double beta)
{
    int i, j, k;
    double dth, poly, pder, rlast = 0.0, sum, delr, r;
    if (!n) return;
    dth = M_PI / (2.0 * (double)n);
    for (k = 0; k < n; ++k)
    {
        r = -cos((2.0 * (double)k + 1.0) * dth);
        if (k) r = 0.5 * (r + rlast);
        for (j = 1; j < 30; ++j)
        {
            mfo_jacobfd(1, &r, &poly, &pder, n, alpha, beta);
            for (i = 0, sum = 0.0; i < k; ++i) sum += 1.0 / (r - z[i]);
            delr = -poly / (pder - sum * poly);
            r += delr;
            if (fabs(delr) < 100 * DBL_EPSILON) break;
        }
        z[k]  = r;
        rlast = r;
    }
}

/* Gauss-Lobatto-Jacobi: Polylib.cpp:195-230 */
void mfo_zwglj(double *z, double *w, int np, double alpha, double beta)
{
    if (np == 1)
    {
        z[0] = 0.0;
        w[0] = 2.0;
    }
    else if (np == 2)
    {
        z[0] = -1.0;
        z[1] = 1.0;
        w[0] = 1.0;
        w[1] = 1.0;
    }
    else
    {
        int i;
        double fac, apb = alpha + beta;
        z[0]      = -1.0;
        z[np - 1] = 1.0;
        jacobz(np - 2, z + 1, alpha + 1.0, beta + 1.0);
        mfo_jacobfd(np, z, w, NULL, np - 1, alpha, beta);
        fac = pow(2.0, apb + 1) * gammaFrac(np, alpha, np, 0.0) * gammaFrac(np, beta, np + 1, apb);
        fac /= (np - 1);
        for (i = 0; i < np; ++i) w[i] = fac / (w[i] * w[i]);
        w[0] *= (beta + 1.0);
        w[np - 1] *= (alpha + 1.0);
    }
}

/* Gauss-Radau-Jacobi with point at -1: Polylib.cpp:119-145 */
void mfo_zwgrjm(double *z, double *w, int np, double alpha, double beta)
{
    if (np == 1)
    {
        z[0] = 0.0;
        w[0] = 2.0;
    }
    else
    {
        int i;
        double fac, apb = alpha + beta;
        z[0] = -1.0;
        jacobz(np - 1, z + 1, alpha, beta + 1);
        mfo_jacobfd(np, z, w, NULL, np - 1, alpha, beta);
        fac = pow(2.0, apb) * gammaFrac(np, alpha, np, 0.0) * gammaFrac(np, beta, np + 1, apb);
        fac /= (beta + np);
        for (i = 0; i < np; ++i) w[i] = fac * (1 - z[i]) / (w[i] * w[i]);
        w[0] *= (beta + 1.0);
    }
}

/* D[i*np+j] = h_i'(z_j) on Gauss-Lobatto-Jacobi points: Polylib.cpp:690-724 */
void mfo_Dglj(double *D, const double *z, int np, double alpha, double beta)
{
    int i, j;
    double *pd;
    if (np <= 1)
    {
        D[0] = 0.0;
        return;
    }
    pd    = (double *)malloc((size_t)np * sizeof(double));
    pd[0] = 2.0 * pow(-1.0, np) * gammaFrac(np, beta, np - 1, 0.0);
    pd[0] /= gammaF(beta + 2.0);
    jacobd(np - 2, z + 1, pd + 1, np - 2, alpha + 1, beta + 1);
    for (i = 1; i < np - 1; ++i) pd[i] *= (1.0 - z[i] * z[i]);
    pd[np - 1] = -2.0 * gammaFrac(np, alpha, np - 1, 0.0);
    pd[np - 1] /= gammaF(alpha + 2.0);
    for (i = 0; i < np; i++)
        for (j = 0; j < np; j++)
        {
            if (i != j)
                D[i * np + j] = pd[j] / (pd[i] * (z[j] - z[i]));
            else if (j == 0)
                D[i * np + j] = (alpha - (np - 1) * (np + alpha + beta)) / (2.0 * (beta + 2.0));
            else if (j == np - 1)
                D[i * np + j] = -(beta - (np - 1) * (np + alpha + beta)) / (2.0 * (alpha + 2.0));
            else
                D[i * np + j] = (alpha - beta + (alpha + beta) * z[j]) / (2.0 * (1.0 - z[j] * z[j]));
        }
    free(pd);
}

/* Gauss-Radau-Jacobi (-1) derivative matrix: Polylib.cpp:589-624 */
void mfo_Dgrjm(double *D, const double *z, int np, double alpha, double beta)
{
    int i, j;
    double *pd;
    if (np <= 0)
    {
        D[0] = 0.0;
        return;
    }
    pd    = (double *)malloc((size_t)np * sizeof(double));
    pd[0] = pow(-1.0, np - 1) * gammaFrac(np + 1, beta, np, 0.0);
    pd[0] /= gammaF(beta + 2.0);
    jacobd(np - 1, z + 1, pd + 1, np - 1, alpha, beta + 1);
    for (i = 1; i < np; ++i) pd[i] *= (1 + z[i]);
    for (i = 0; i < np; i++)
        for (j = 0; j < np; j++)
        {
            if (i != j)
                D[i * np + j] = pd[j] / (pd[i] * (z[j] - z[i]));
            else if (j == 0)
                D[i * np + j] = -(np + alpha + beta + 1.0) * (np - 1.0) / (2.0 * (beta + 2.0));
            else
                D[i * np + j] = (alpha - beta + 1.0 + (alpha + beta + 1.0) * z[j]) / (2.0 * (1.0 - z[j] * z[j]));
        }
    free(pd);
}

/* ===================================================================== points/basis */

/* Foundations/GaussPoints.cpp:69-147 (points+weights), :154-236 (derivative matrix,
 * Polylib's raw array copied verbatim => D[k*np+i] = h_k'(z_i)) */
void mfo_points(int ptype, int np, double *z, double *w, double *D)
{
    switch (ptype)
    {
        case MFO_GLL:
            mfo_zwglj(z, w, np, 0.0, 0.0);
            if (D) mfo_Dglj(D, z, np, 0.0, 0.0);
            break;
        case MFO_GRJM_A1:
            mfo_zwgrjm(z, w, np, 1.0, 0.0);
            if (D) mfo_Dgrjm(D, z, np, 1.0, 0.0);
            break;
        case MFO_GRJM_A2:
            mfo_zwgrjm(z, w, np, 2.0, 0.0);
            if (D) mfo_Dgrjm(D, z, np, 2.0, 0.0);
            break;
    }
}

int mfo_basis_rows(int btype, int nm)
{
    switch (btype)
    {
        case MFO_MOD_A: return nm;
        case MFO_MOD_B: return nm * (nm + 1) / 2;
        case MFO_MOD_C: return nm * (nm + 1) * (nm + 2) / 6;
        case MFO_MOD_PYR_C: return nm * (nm + 1) * (2 * nm + 1) / 6;
    }
    return 0;
}

/* Modified_A rows: (1-z)/2, (1+z)/2, then (1-z)/2 (1+z)/2 P^{1,1}_{p-2}: Basis.cpp:392-416 */
static void modified_a(int nm, int np, const double *z, double *b)
{
    int i, p;
    for (i = 0; i < np; ++i)
    {
        b[i]      = 0.5 * (1 - z[i]);
        b[np + i] = 0.5 * (1 + z[i]);
    }
    for (p = 2; p < nm; ++p)
    {
        double *mode = b + (size_t)p * np;
        mfo_jacobfd(np, z, mode, NULL, p - 2, 1.0, 1.0);
        for (i = 0; i < np; ++i) mode[i] *= b[i] * b[np + i];
    }
}

/* Modified_B, (p,q) rows with q fastest: Basis.cpp:423-504 */
static void modified_b(int nm, int np, const double *z, double *b)
{
    int i, p, q;
    double *mode;
    const double *one_m_z_pow, *one_p_z;
    for (i = 0; i < np; ++i)
    {
        b[i]      = 0.5 * (1 - z[i]);
        b[np + i] = 0.5 * (1 + z[i]);
    }
    mode = b + 2 * (size_t)np;
    for (q = 2; q < nm; ++q, mode += np)
    {
        mfo_jacobfd(np, z, mode, NULL, q - 2, 1.0, 1.0);
        for (i = 0; i < np; ++i) mode[i] *= b[i] * b[np + i];
    }
    /* second row (p = 1) */
    for (i = 0; i < np; ++i) mode[i] = 0.5 * (1 - z[i]);
    mode += np;
    for (q = 2; q < nm; ++q, mode += np)
    {
        mfo_jacobfd(np, z, mode, NULL, q - 2, 1.0, 1.0);
        for (i = 0; i < np; ++i) mode[i] *= b[i] * b[np + i];
    }
    /* rows p >= 2 */
    one_m_z_pow = b;
    one_p_z     = b + np;
    for (p = 2; p < nm; ++p)
    {
        for (i = 0; i < np; ++i) mode[i] = b[i] * one_m_z_pow[i];
        one_m_z_pow = mode;
        mode += np;
        for (q = 1; q < nm - p; ++q, mode += np)
        {
            mfo_jacobfd(np, z, mode, NULL, q - 1, 2 * p - 1, 1.0);
            for (i = 0; i < np; ++i) mode[i] *= one_m_z_pow[i] * one_p_z[i];
        }
    }
}

void mfo_basis(int btype, int nm, int np, const double *z, const double *D, double *bdata, double *dbdata)
{
    int rows = mfo_basis_rows(btype, nm), m, i, j;
    if (btype == MFO_MOD_A)
        modified_a(nm, np, z, bdata);
    else if (btype == MFO_MOD_B)
        modified_b(nm, np, z, bdata);
    else if (btype == MFO_MOD_PYR_C)
    {
        /* ModifiedPyr_C: Basis.cpp:569-669.  Rows ordered (p, q, r) with r fastest and nm - max(p,q) rows per
         * (p,q); vertex/edge/face rows are copies of Modified_B rows, face-0 rows are powers of (1-z)/2,
         * interior rows [(1-z)/2]^{p+q-2} (1+z)/2 P_{r-1}^{2p+2q-3,1}. */
        int nb       = mfo_basis_rows(MFO_MOD_B, nm), p, q, r;
        double *modb = (double *)malloc((size_t)nb * np * sizeof(double));
        size_t N, boff = 0, off = 0;
        double *mode;
        const double *one_p_z;
        modified_b(nm, np, z, modb);
        N = (size_t)np * nm * (nm + 1) / 2;
        memcpy(bdata, modb, N * sizeof(double));
        off += N;
        boff += (size_t)np * nm;
        N = (size_t)np * (nm - 1);
        memcpy(bdata + off, modb + boff, N * sizeof(double));
        off += N;
        N = (size_t)np * (nm - 1) * nm / 2;
        memcpy(bdata + off, modb + boff, N * sizeof(double));
        off += N;
        boff += (size_t)np * (nm - 1);
        mode = bdata + off;
        for (p = 2; p < nm; ++p)
        {
            N = (size_t)np * (nm - p);
            memcpy(mode, modb + boff, N * sizeof(double));
            mode += N;
            memcpy(mode, modb + boff, N * sizeof(double));
            mode += N;
            boff += N;
            one_p_z = bdata + np;
            for (q = 2; q < nm; ++q)
            {
                double *one_m_z_pow = mode;
                for (i = 0; i < np; ++i) mode[i] = pow(bdata[i], p + q - 2);
                mode += np;
                for (r = 1; r < nm - (p > q ? p : q); ++r)
                {
                    mfo_jacobfd(np, z, mode, NULL, r - 1, 2 * p + 2 * q - 3, 1.0);
                    for (i = 0; i < np; ++i) mode[i] *= one_m_z_pow[i] * one_p_z[i];
                    mode += np;
                }
            }
        }
        free(modb);
    }
    else
    {
        /* Modified_C = re-indexed copy of Modified_B (phi^c_{pqr} = phi^b_{p+q,r}): Basis.cpp:513-558 */
        int nb        = mfo_basis_rows(MFO_MOD_B, nm), p;
        double *modb  = (double *)malloc((size_t)nb * np * sizeof(double));
        size_t boff = 0, off = 0;
        modified_b(nm, np, z, modb);
        for (p = 0; p < nm; ++p)
        {
            size_t N = (size_t)np * (nm - p) * (nm - p + 1) / 2;
            memcpy(bdata + off, modb + boff, N * sizeof(double));
            boff += (size_t)np * (nm - p);
            off += N;
        }
        free(modb);
    }
    /* dbdata = D * bdata as the column-major DGEMM of Basis.cpp:418-420 would do
     * (plain triple loop; BLAS summation order is not part of the reference's contract) */
    for (m = 0; m < rows; ++m)
        for (i = 0; i < np; ++i)
        {
            double s = 0.0;
            for (j = 0; j < np; ++j) s += D[i + (size_t)j * np] * bdata[(size_t)m * np + j];
            dbdata[(size_t)m * np + i] = s;
        }
}

/* ===================================================================== element tables */

struct mfo_elem
{
    int shape, dim, coordim, nm, nmTot, nqTot;
    int nq[3], ptype[3], btype[3], rows[3];
    double *z[3], *w[3], *ws[3], *D[3], *b[3], *db[3];
    double *h0, *h1, *h2, *h3;
};

static int count_modes(int shape, int nm)
{
    switch (shape)
    {
        case MFO_QUAD: return nm * nm;
        case MFO_TRI: return nm * (nm + 1) / 2;
        case MFO_HEX: return nm * nm * nm;
        case MFO_PRISM: return nm * nm * (nm + 1) / 2;
        case MFO_TET: return nm * (nm + 1) * (nm + 2) / 6;
        case MFO_PYR: return nm * (nm + 1) * (2 * nm + 1) / 6;
        case MFO_SEG: return nm;
    }
    return -1;
}

mfo_elem *mfo_create(int shape, int nm, int nq0)
{
    mfo_elem *e = (mfo_elem *)calloc(1, sizeof(mfo_elem));
    int d, i;
    e->shape = shape;
    e->nm    = nm;
    e->dim   = shape == MFO_SEG ? 1 : ((shape == MFO_QUAD || shape == MFO_TRI) ? 2 : 3);
    e->coordim = e->dim;
    for (d = 0; d < 3; ++d)
    {
        e->nq[d]    = nq0;
        e->ptype[d] = MFO_GLL;
        e->btype[d] = MFO_MOD_A;
    }
    /* MeshGraph.cpp:1609-1762 defaults; MatrixFree preconditions Helmholtz.h:329-331,1042-1046,2017-2021 */
    switch (shape)
    {
        case MFO_SEG:
        case MFO_QUAD:
        case MFO_HEX: break;
        case MFO_TRI:
            e->nq[1] = nq0 - 1; e->ptype[1] = MFO_GRJM_A1; e->btype[1] = MFO_MOD_B;
            break;
        case MFO_PRISM:
            e->nq[2] = nq0 - 1; e->ptype[2] = MFO_GRJM_A1; e->btype[2] = MFO_MOD_B;
            break;
        case MFO_TET:
            e->nq[1] = nq0 - 1; e->ptype[1] = MFO_GRJM_A1; e->btype[1] = MFO_MOD_B;
            e->nq[2] = nq0 - 1; e->ptype[2] = MFO_GRJM_A2; e->btype[2] = MFO_MOD_C;
            break;
        case MFO_PYR: /* MatrixFree precondition nq2 = nq0 - 1: Helmholtz.h:1519-1525 */
            e->nq[2] = nq0 - 1; e->ptype[2] = MFO_GRJM_A2; e->btype[2] = MFO_MOD_PYR_C;
            break;
        default: free(e); return NULL;
    }
    e->nmTot = count_modes(shape, nm);
    e->nqTot = 1;
    for (d = 0; d < e->dim; ++d)
    {
        int np = e->nq[d];
        double fac;
        e->nqTot *= np;
        e->rows[d] = mfo_basis_rows(e->btype[d], nm);
        e->z[d]    = (double *)malloc(sizeof(double) * np);
        e->w[d]    = (double *)malloc(sizeof(double) * np);
        e->ws[d]   = (double *)malloc(sizeof(double) * np);
        e->D[d]    = (double *)malloc(sizeof(double) * np * np);
        e->b[d]    = (double *)malloc(sizeof(double) * e->rows[d] * np);
        e->db[d]   = (double *)malloc(sizeof(double) * e->rows[d] * np);
        mfo_points(e->ptype[d], np, e->z[d], e->w[d], e->D[d]);
        mfo_basis(e->btype[d], nm, np, e->z[d], e->D[d], e->b[d], e->db[d]);
        /* Helper<DIM>: collapsed-coordinate Jacobian folded into the weights, Operator.hpp:244-258 */
        fac = e->ptype[d] == MFO_GRJM_A1 ? 0.5 : (e->ptype[d] == MFO_GRJM_A2 ? 0.25 : 1.0);
        for (i = 0; i < np; ++i) e->ws[d][i] = fac * e->w[d][i];
    }
    /* collapsed-coordinate factors: Helmholtz.h:304-315 (Tri), 1017-1028 (Prism), 1985-2004 (Tet) */
    if (shape == MFO_TRI || shape == MFO_PRISM)
    {
        int dl = shape == MFO_TRI ? 1 : 2;
        e->h0  = (double *)malloc(sizeof(double) * e->nq[0]);
        e->h1  = (double *)malloc(sizeof(double) * e->nq[dl]);
        for (i = 0; i < e->nq[0]; ++i) e->h0[i] = 0.5 * (1 + e->z[0][i]);
        for (i = 0; i < e->nq[dl]; ++i) e->h1[i] = 2.0 / (1 - e->z[dl][i]);
    }
    else if (shape == MFO_PYR)
    {
        /* Helmholtz.h:1489-1507 */
        e->h0 = (double *)malloc(sizeof(double) * e->nq[0]);
        e->h1 = (double *)malloc(sizeof(double) * e->nq[1]);
        e->h2 = (double *)malloc(sizeof(double) * e->nq[2]);
        for (i = 0; i < e->nq[0]; ++i) e->h0[i] = 0.5 * (1 + e->z[0][i]);
        for (i = 0; i < e->nq[1]; ++i) e->h1[i] = 0.5 * (1 + e->z[1][i]);
        for (i = 0; i < e->nq[2]; ++i) e->h2[i] = 2.0 / (1 - e->z[2][i]);
    }
    else if (shape == MFO_TET)
    {
        e->h0 = (double *)malloc(sizeof(double) * e->nq[0]);
        e->h1 = (double *)malloc(sizeof(double) * e->nq[1]);
        e->h2 = (double *)malloc(sizeof(double) * e->nq[1]);
        e->h3 = (double *)malloc(sizeof(double) * e->nq[2]);
        for (i = 0; i < e->nq[0]; ++i) e->h0[i] = 0.5 * (1 + e->z[0][i]);
        for (i = 0; i < e->nq[1]; ++i)
        {
            e->h1[i] = 0.5 * (1 + e->z[1][i]);
            e->h2[i] = 2.0 / (1 - e->z[1][i]);
        }
        for (i = 0; i < e->nq[2]; ++i) e->h3[i] = 2.0 / (1 - e->z[2][i]);
    }
    return e;
}

void mfo_destroy(mfo_elem *e)
{
    int d;
    if (!e) return;
    for (d = 0; d < 3; ++d)
    {
        free(e->z[d]); free(e->w[d]); free(e->ws[d]); free(e->D[d]); free(e->b[d]); free(e->db[d]);
    }
    free(e->h0); free(e->h1); free(e->h2); free(e->h3);
    free(e);
}

int mfo_dim(const mfo_elem *e) { return e->dim; }
void mfo_set_coordim(mfo_elem *e, int c) { if (e->shape == MFO_SEG && c >= 1 && c <= 3) e->coordim = c; }
int mfo_nmtot(const mfo_elem *e) { return e->nmTot; }
int mfo_nqtot(const mfo_elem *e) { return e->nqTot; }
int mfo_nq(const mfo_elem *e, int d) { return e->nq[d]; }
int mfo_ptype(const mfo_elem *e, int d) { return e->ptype[d]; }
int mfo_btype(const mfo_elem *e, int d) { return e->btype[d]; }
int mfo_brows(const mfo_elem *e, int d) { return e->rows[d]; }
const double *mfo_table(const mfo_elem *e, int d, int which)
{
    switch (which)
    {
        case 0: return e->b[d];
        case 1: return e->db[d];
        case 2: return e->D[d];
        case 3: return e->z[d];
        case 4: return e->w[d];
    }
    return NULL;
}

/* ============================================================== single-element kernels */

/* IProductKernels.hpp:14-37 */
static inline void scale_append(double *store, double pos, double scale, int SCALE, int APPEND)
{
    if (SCALE && APPEND)
        *store += pos * scale;
    else if (APPEND)
        *store = *store + pos;
    else if (SCALE)
        *store = pos * scale;
    else
        *store = pos;
}

/* ---------------- Quad.  BwdTransKernels.hpp:35-76 */
static void k_bwd_quad(int nm, int nq, const double *in, const double *b0, const double *b1, double *wsp, double *out)
{
    int i, j, p, q;
    for (i = 0; i < nq; ++i)
        for (q = 0; q < nm; ++q)
        {
            double t = in[q * nm] * b0[i];
            for (p = 1; p < nm; ++p) t += in[q * nm + p] * b0[p * nq + i];
            wsp[i * nm + q] = t;
        }
    for (j = 0; j < nq; ++j)
        for (i = 0; i < nq; ++i)
        {
            double t = wsp[i * nm] * b1[j];
            for (q = 1; q < nm; ++q) t += wsp[i * nm + q] * b1[q * nq + j];
            out[j * nq + i] = t;
        }
}

/* IProductKernels.hpp:76-133 */
static void k_ip_quad(int nm, int nq, const double *in, const double *b0, const double *b1, const double *w0,
                      const double *w1, const double *jac, int DEF, double *sums_j, double *out, double scale,
                      int SCALE, int APPEND)
{
    int i, j, p, q;
    for (p = 0; p < nm; ++p)
    {
        for (j = 0; j < nq; ++j)
        {
            double s = 0.0;
            for (i = 0; i < nq; ++i)
            {
                double jv   = DEF ? jac[j * nq + i] : jac[0];
                double prod = in[j * nq + i] * b0[p * nq + i] * jv;
                s += prod * w0[i];
            }
            sums_j[j] = s;
        }
        for (q = 0; q < nm; ++q)
        {
            double s = 0.0;
            for (j = 0; j < nq; ++j)
            {
                double prod = sums_j[j] * b1[q * nq + j];
                s += prod * w1[j];
            }
            scale_append(&out[q * nm + p], s, scale, SCALE, APPEND);
        }
    }
}

/* PhysDerivKernels.hpp:39-90 */
static void k_dtensor2(int nq0, int nq1, const double *in, const double *D0, const double *D1, double *d0, double *d1)
{
    int i, j, k;
    for (i = 0; i < nq0; ++i)
        for (j = 0; j < nq1; ++j)
        {
            double s = 0.0;
            for (k = 0; k < nq0; ++k) s += D0[k * nq0 + i] * in[j * nq0 + k];
            d0[j * nq0 + i] = s;
        }
    for (i = 0; i < nq0; ++i)
        for (j = 0; j < nq1; ++j)
        {
            double s = 0.0;
            for (k = 0; k < nq1; ++k) s += in[k * nq0 + i] * D1[k * nq1 + j];
            d1[j * nq0 + i] = s;
        }
}

/* ---------------- Tri.  BwdTransKernels.hpp:78-126 */
static void k_bwd_tri(int nm, int nq0, int nq1, const double *in, const double *b0, const double *b1, double *ps,
                      double *out)
{
    int e0, e1, p, q, mode, idx = 0;
    for (e1 = 0; e1 < nq1; ++e1)
    {
        for (p = 0, mode = 0; p < nm; ++p)
        {
            double s = 0.0;
            for (q = 0; q < nm - p; ++q, ++mode) s += b1[mode * nq1 + e1] * in[mode];
            ps[p] = s;
        }
        for (e0 = 0; e0 < nq0; ++e0, ++idx)
        {
            double s = 0.0;
            for (p = 0; p < nm; ++p) s += ps[p] * b0[p * nq0 + e0];
            /* CORRECT (eModified_A): singular-vertex term */
            s += (in[1] * b0[nq0 + e0]) * b1[nq1 + e1];
            out[idx] = s;
        }
    }
}

/* IProductKernels.hpp:135-234 */
static void k_ip_tri(int nm, int nq0, int nq1, const double *in, const double *b0, const double *b1,
                     const double *w0, const double *w1, const double *jac, int DEF, double *es, double *out,
                     double scale, int SCALE, int APPEND)
{
    int p, q, e0, e1, mode = 0;
    for (p = 0; p < nm; ++p)
    {
        int idx = 0;
        for (e1 = 0; e1 < nq1; ++e1)
        {
            double s = 0.0;
            for (e0 = 0; e0 < nq0; ++e0, ++idx)
            {
                double jv   = DEF ? jac[e1 * nq0 + e0] : jac[0];
                double prod = in[idx] * b0[p * nq0 + e0] * jv;
                s += prod * w0[e0];
            }
            es[e1] = s;
        }
        for (q = 0; q < nm - p; ++q, ++mode)
        {
            double s = 0.0;
            for (e1 = 0; e1 < nq1; ++e1)
            {
                double prod = es[e1] * b1[mode * nq1 + e1];
                s += prod * w1[e1];
            }
            scale_append(&out[mode], s, scale, SCALE, APPEND);
        }
    }
    {
        int idx  = 0;
        double c = 0.0;
        for (e1 = 0; e1 < nq1; ++e1)
        {
            double pre = DEF ? w1[e1] * b1[nq1 + e1] : w1[e1] * jac[0] * b1[nq1 + e1];
            for (e0 = 0; e0 < nq0; ++e0, ++idx)
            {
                double prod = in[idx] * pre * w0[e0];
                if (DEF) prod = prod * jac[e1 * nq0 + e0];
                c += prod * b0[nq0 + e0];
            }
        }
        scale_append(&out[1], c, scale, SCALE, 1);
    }
}

/* ---------------- Hex.  BwdTransKernels.hpp:302-372 */
static void k_bwd_hex(int nm, int nq, const double *in, const double *b0, const double *b1, const double *b2,
                      double *s_irq, double *s_jir, double *out)
{
    int i, j, k, p, q, r, c;
    for (i = 0, c = 0; i < nq; ++i)
        for (r = 0; r < nm; ++r)
            for (q = 0; q < nm; ++q, ++c)
            {
                const double *u = in + (r * nm + q) * nm;
                double t        = u[0] * b0[i];
                for (p = 1; p < nm; ++p) t += u[p] * b0[p * nq + i];
                s_irq[c] = t;
            }
    for (j = 0, c = 0; j < nq; ++j)
        for (i = 0; i < nq; ++i)
            for (r = 0; r < nm; ++r, ++c)
            {
                const double *u = s_irq + (i * nm + r) * nm;
                double t        = u[0] * b1[j];
                for (q = 1; q < nm; ++q) t += u[q] * b1[q * nq + j];
                s_jir[c] = t;
            }
    for (k = 0, c = 0; k < nq; ++k)
        for (j = 0; j < nq; ++j)
            for (i = 0; i < nq; ++i, ++c)
            {
                const double *u = s_jir + (j * nq + i) * nm;
                double t        = u[0] * b2[k];
                for (r = 1; r < nm; ++r) t += u[r] * b2[r * nq + k];
                out[c] = t;
            }
}

/* IProductKernels.hpp:236-314 */
static void k_ip_hex(int nm, int nq, const double *in, const double *b0, const double *b1, const double *b2,
                     const double *w0, const double *w1, const double *w2, const double *jac, int DEF,
                     double *s_kj, double *s_k, double *out, double scale, int SCALE, int APPEND)
{
    int i, j, k, p, q, r;
    for (p = 0; p < nm; ++p)
    {
        int ckji = 0, ckj = 0;
        for (k = 0; k < nq; ++k)
            for (j = 0; j < nq; ++j, ++ckj)
            {
                double s = 0.0;
                for (i = 0; i < nq; ++i, ++ckji)
                {
                    double jv   = DEF ? jac[nq * nq * k + nq * j + i] : jac[0];
                    double prod = in[ckji] * b0[i + nq * p] * jv;
                    s += prod * w0[i];
                }
                s_kj[ckj] = s;
            }
        for (q = 0; q < nm; ++q)
        {
            ckj = 0;
            for (k = 0; k < nq; ++k)
            {
                double s = 0.0;
                for (j = 0; j < nq; ++j, ++ckj)
                {
                    double prod = s_kj[ckj] * b1[q * nq + j];
                    s += prod * w1[j];
                }
                s_k[k] = s;
            }
            for (r = 0; r < nm; ++r)
            {
                double s = 0.0;
                for (k = 0; k < nq; ++k)
                {
                    double prod = s_k[k] * b2[r * nq + k];
                    s += prod * w2[k];
                }
                scale_append(&out[r * nm * nm + q * nm + p], s, scale, SCALE, APPEND);
            }
        }
    }
}

/* PhysDerivKernels.hpp:219-294 */
static void k_dtensor3(int nq0, int nq1, int nq2, const double *in, const double *D0, const double *D1,
                       const double *D2, double *d0, double *d1, double *d2)
{
    int i, j, k, blk;
    for (i = 0; i < nq0; ++i)
        for (j = 0; j < nq1 * nq2; ++j)
        {
            double s = 0.0;
            for (k = 0; k < nq0; ++k) s += D0[k * nq0 + i] * in[j * nq0 + k];
            d0[j * nq0 + i] = s;
        }
    for (blk = 0; blk < nq2; ++blk)
    {
        int start = blk * nq0 * nq1;
        for (i = 0; i < nq0; ++i)
            for (j = 0; j < nq1; ++j)
            {
                double s = 0.0;
                for (k = 0; k < nq1; ++k) s += in[start + k * nq0 + i] * D1[k * nq1 + j];
                d1[start + j * nq0 + i] = s;
            }
    }
    for (i = 0; i < nq0 * nq1; ++i)
        for (j = 0; j < nq2; ++j)
        {
            double s = 0.0;
            for (k = 0; k < nq2; ++k) s += in[k * nq0 * nq1 + i] * D2[k * nq2 + j];
            d2[j * nq0 * nq1 + i] = s;
        }
}

/* ---------------- Prism.  BwdTransKernels.hpp:224-299 */
static void k_bwd_prism(int nm, int nq0, int nq1, int nq2, const double *in, const double *b0, const double *b1,
                        const double *b2, double *fpq, double *fp, double *out)
{
    int i, j, k, p, q, r, c = 0;
    for (k = 0; k < nq2; ++k)
    {
        int mpqr = 0, mpq = 0, mpr = 0;
        for (p = 0; p < nm; ++p)
        {
            for (q = 0; q < nm; ++q, ++mpq)
            {
                double s = 0.0;
                for (r = 0; r < nm - p; ++r, ++mpqr) s += in[mpqr] * b2[(mpr + r) * nq2 + k];
                fpq[mpq] = s;
            }
            mpr += nm - p;
        }
        for (j = 0; j < nq1; ++j)
        {
            mpq = 0;
            for (p = 0; p < nm; ++p)
            {
                double s = 0.0;
                for (q = 0; q < nm; ++q, ++mpq) s += fpq[mpq] * b1[q * nq1 + j];
                fp[p] = s;
            }
            for (i = 0; i < nq0; ++i, ++c)
            {
                double v = 0.0, ba2 = b2[nq2 + k], ba0 = b0[nq0 + i];
                for (p = 0; p < nm; ++p) v += fp[p] * b0[p * nq0 + i];
                for (q = 0; q < nm; ++q) v += (ba2 * b1[q * nq1 + j]) * (ba0 * in[q * nm + 1]);
                out[c] = v;
            }
        }
    }
}

/* IProductKernels.hpp:316-450 */
static void k_ip_prism(int nm, int nq0, int nq1, int nq2, const double *in, const double *b0, const double *b1,
                       const double *b2, const double *w0, const double *w1, const double *w2, const double *jac,
                       int DEF, double *s_kj, double *s_k, double *corr, double *out, double scale, int SCALE,
                       int APPEND)
{
    int i, j, k, p, q, r, mpr = 0, mpqr = 0;
    for (p = 0; p < nm; ++p)
    {
        int ckji = 0, ckj = 0;
        for (k = 0; k < nq2; ++k)
            for (j = 0; j < nq1; ++j, ++ckj)
            {
                double s = 0.0;
                for (i = 0; i < nq0; ++i, ++ckji)
                {
                    double jv   = DEF ? jac[nq0 * nq1 * k + nq0 * j + i] : jac[0];
                    double prod = b0[nq0 * p + i] * jv * w0[i];
                    s += prod * in[ckji];
                }
                s_kj[ckj] = s;
            }
        for (q = 0; q < nm; ++q)
        {
            ckj = 0;
            for (k = 0; k < nq2; ++k)
            {
                double s = 0.0;
                for (j = 0; j < nq1; ++j, ++ckj) s += (b1[q * nq1 + j] * w1[j]) * s_kj[ckj];
                s_k[k] = s;
            }
            for (r = 0; r < nm - p; ++r, ++mpqr)
            {
                double s = 0.0;
                for (k = 0; k < nq2; ++k) s += (b2[(mpr + r) * nq2 + k] * w2[k]) * s_k[k];
                scale_append(&out[mpqr], s, scale, SCALE, APPEND);
            }
        }
        mpr += nm - p;
    }
    /* CORRECT: singular edge */
    {
        int c = 0;
        for (q = 0; q < nm; ++q) corr[q] = 0.0;
        for (k = 0; k < nq2; ++k)
        {
            double kw = w2[k];
            if (!DEF) kw = kw * jac[0];
            for (j = 0; j < nq1; ++j)
            {
                double kjw = kw * w1[j];
                for (i = 0; i < nq0; ++i, ++c)
                {
                    double kjiw = kjw * w0[i];
                    double prod = kjiw * in[c];
                    double ba2, ba0;
                    if (DEF) prod *= jac[k * nq1 * nq0 + j * nq0 + i];
                    ba2 = b2[nq2 + k];
                    ba0 = b0[nq0 + i];
                    for (q = 0; q < nm; ++q) corr[q] += (ba2 * b1[q * nq1 + j]) * (ba0 * prod);
                }
            }
        }
        for (q = 0; q < nm; ++q) scale_append(&out[nm * q + 1], corr[q], scale, SCALE, 1);
    }
}

/* ---------------- Pyr.  BwdTransKernels.hpp:131-222 (isotropic modes) */
static void k_bwd_pyr(int nm, int nq0, int nq1, int nq2, const double *in, const double *b0, const double *b1,
                      const double *b2, double *fpq, double *fp, double *out)
{
    int i, j, k, p, q, r, c = 0;
    for (k = 0; k < nq2; ++k)
    {
        int mpqr = 0, mpq = 0;
        for (p = 0; p < nm; ++p)
            for (q = 0; q < nm; ++q, ++mpq)
            {
                int len  = nm - (p > q ? p : q);
                double s = 0.0;
                for (r = 0; r < len; ++r, ++mpqr) s += in[mpqr] * b2[mpqr * nq2 + k];
                fpq[mpq] = s;
            }
        for (j = 0; j < nq1; ++j)
        {
            mpq = 0;
            for (p = 0; p < nm; ++p)
            {
                double s = 0.0;
                for (q = 0; q < nm; ++q, ++mpq) s += fpq[mpq] * b1[q * nq1 + j];
                fp[p] = s;
            }
            for (i = 0; i < nq0; ++i, ++c)
            {
                double v = 0.0, t1;
                for (p = 0; p < nm; ++p) v += fp[p] * b0[p * nq0 + i];
                /* CORRECT: top vertex */
                t1 = b0[i] * b1[nq1 + j];
                t1 += b0[nq0 + i] * b1[j];
                t1 += b0[nq0 + i] * b1[nq1 + j];
                t1 = t1 * b2[nq2 + k];
                v += t1 * in[1];
                out[c] = v;
            }
        }
    }
}

/* IProductKernels.hpp:455-598 */
static void k_ip_pyr(int nm, int nq0, int nq1, int nq2, const double *in, const double *b0, const double *b1,
                     const double *b2, const double *w0, const double *w1, const double *w2, const double *jac, int DEF,
                     double *s_kj, double *s_k, double *out, double scale, int SCALE, int APPEND)
{
    int i, j, k, p, q, r, mpqr = 0;
    for (p = 0; p < nm; ++p)
    {
        int ckji = 0, ckj = 0;
        for (k = 0; k < nq2; ++k)
            for (j = 0; j < nq1; ++j, ++ckj)
            {
                double s = 0.0;
                for (i = 0; i < nq0; ++i, ++ckji)
                {
                    double jv   = DEF ? jac[nq0 * nq1 * k + nq0 * j + i] : jac[0];
                    double prod = b0[nq0 * p + i] * jv * w0[i];
                    s += prod * in[ckji];
                }
                s_kj[ckj] = s;
            }
        for (q = 0; q < nm; ++q)
        {
            int len = nm - (p > q ? p : q);
            ckj     = 0;
            for (k = 0; k < nq2; ++k)
            {
                double s = 0.0;
                for (j = 0; j < nq1; ++j, ++ckj) s += (b1[q * nq1 + j] * w1[j]) * s_kj[ckj];
                s_k[k] = s;
            }
            for (r = 0; r < len; ++r, ++mpqr)
            {
                double s = 0.0;
                for (k = 0; k < nq2; ++k) s += (b2[mpqr * nq2 + k] * w2[k]) * s_k[k];
                scale_append(&out[mpqr], s, scale, SCALE, APPEND);
            }
        }
    }
    /* CORRECT: top vertex, accumulated point by point into mode 1 */
    {
        int c = 0;
        for (k = 0; k < nq2; ++k)
        {
            double kw = w2[k];
            if (!DEF) kw = kw * jac[0];
            for (j = 0; j < nq1; ++j)
            {
                double kjw = kw * w1[j];
                for (i = 0; i < nq0; ++i, ++c)
                {
                    double q3 = kjw * w0[i], t;
                    if (DEF) q3 = q3 * jac[k * nq0 * nq1 + j * nq0 + i];
                    t = b0[i] * b1[nq1 + j];
                    t += b0[nq0 + i] * b1[j];
                    t += b0[nq0 + i] * b1[nq1 + j];
                    t = t * b2[nq2 + k];
                    t = t * in[c];
                    scale_append(&out[1], t * q3, scale, SCALE, 1);
                }
            }
        }
    }
}

/* ---------------- Tet.  BwdTransKernels.hpp:374-484 */
static void k_bwd_tet(int nm, int nq0, int nq1, int nq2, const double *in, const double *b0, const double *b1,
                      const double *b2, double *fpq, double *fp, double *out)
{
    int i, j, k, p, q, r, c = 0;
    for (k = 0; k < nq2; ++k)
    {
        int cpq = 0, mode = 0;
        for (p = 0; p < nm; ++p)
            for (q = 0; q < nm - p; ++q, ++cpq)
            {
                double s = in[mode] * b2[k + nq2 * mode];
                ++mode;
                for (r = 1; r < nm - p - q; ++r, ++mode) s += in[mode] * b2[k + nq2 * mode];
                fpq[cpq] = s;
            }
        for (j = 0; j < nq1; ++j)
        {
            mode = cpq = 0;
            for (p = 0; p < nm; ++p)
            {
                double s = fpq[cpq] * b1[mode * nq1 + j];
                ++cpq;
                for (q = 1; q < nm - p; ++q, ++cpq) s += fpq[cpq] * b1[(mode + q) * nq1 + j];
                fp[p] = s;
                mode += nm - p;
            }
            for (i = 0; i < nq0; ++i, ++c)
            {
                double t = b0[i] * fp[0], t1;
                for (p = 1; p < nm; ++p) t += b0[p * nq0 + i] * fp[p];
                /* CORRECT: top vertex */
                t1 = b0[i] * b1[nq1 + j];
                t1 += b0[nq0 + i] * b1[j];
                t1 += b0[nq0 + i] * b1[nq1 + j];
                t1 = t1 * b2[nq2 + k];
                t += t1 * in[1];
                /* bottom vertex */
                t1 = b0[nq0 + i] * b1[nq1 + j];
                t1 = t1 * b2[k];
                t += in[nm] * t1;
                /* singular edge */
                for (r = 1; r < nm - 1; ++r)
                {
                    t1 = b1[nq1 + j] * b0[nq0 + i];
                    t1 = t1 * b2[(r + 1) * nq2 + k];
                    t += in[nm + r] * t1;
                }
                out[c] = t;
            }
        }
    }
}

/* IProductKernels.hpp:600-761 */
static void k_ip_tet(int nm, int nq0, int nq1, int nq2, const double *in, const double *b0, const double *b1,
                     const double *b2, const double *w0, const double *w1, const double *w2, const double *jac,
                     int DEF, double *wsp, double *out, double scale, int SCALE, int APPEND)
{
    double *f = wsp, *fb = wsp + nq1 * nq2;
    int i, j, k, p, q, r, mode = 0, mode2 = 0, cpqr = 0;
    for (p = 0; p < nm; ++p)
    {
        int ckji = 0, ckj = 0;
        for (k = 0; k < nq2; ++k)
            for (j = 0; j < nq1; ++j, ++ckj)
            {
                double jv  = DEF ? jac[nq0 * nq1 * k + nq0 * j] : jac[0];
                double fkj = in[ckji] * b0[nq0 * p] * jv * w0[0];
                ++ckji;
                for (i = 1; i < nq0; ++i, ++ckji)
                {
                    double x;
                    jv = DEF ? jac[nq0 * nq1 * k + nq0 * j + i] : jac[0];
                    x  = in[ckji] * b0[i + nq0 * p] * jv;
                    fkj += x * w0[i];
                }
                f[ckj] = fkj;
            }
        for (q = 0; q < nm - p; ++q, ++mode)
        {
            ckj = 0;
            for (k = 0; k < nq2; ++k)
            {
                double fk = b1[mode * nq1] * f[ckj] * w1[0];
                ++ckj;
                for (j = 1; j < nq1; ++j, ++ckj)
                {
                    double t2 = b1[mode * nq1 + j] * f[ckj];
                    fk += t2 * w1[j];
                }
                fb[k] = fk;
            }
            for (r = 0; r < nm - p - q; ++r, ++mode2, ++cpqr)
            {
                double t = fb[0] * b2[mode2 * nq2] * w2[0];
                for (k = 1; k < nq2; ++k)
                {
                    double t2 = fb[k] * b2[mode2 * nq2 + k];
                    t += t2 * w2[k];
                }
                scale_append(&out[cpqr], t, scale, SCALE, APPEND);
            }
        }
    }
    /* CORRECT */
    {
        int c = 0;
        for (k = 0; k < nq2; ++k)
        {
            double tq2 = w2[k];
            if (!DEF) tq2 = tq2 * jac[0];
            for (j = 0; j < nq1; ++j)
            {
                double tq1 = tq2 * w1[j];
                for (i = 0; i < nq0; ++i, ++c)
                {
                    double tq = tq1 * w0[i], ti = in[c], t, to;
                    if (DEF) tq = tq * jac[k * nq0 * nq1 + j * nq0 + i];
                    t = b0[i] * b1[nq1 + j];
                    t += b0[nq0 + i] * b1[j];
                    t += b0[nq0 + i] * b1[nq1 + j];
                    t  = t * b2[nq2 + k];
                    t  = t * ti;
                    to = t * tq;
                    scale_append(&out[1], to, scale, SCALE, 1);
                    t  = b0[nq0 + i] * b1[nq1 + j] * b2[k] * ti;
                    to = t * tq;
                    scale_append(&out[nm], to, scale, SCALE, 1);
                    for (r = 1; r < nm - 1; ++r)
                    {
                        t  = b2[(r + 1) * nq2 + k] * b1[nq1 + j] * b0[nq0 + i] * ti;
                        to = t * tq;
                        scale_append(&out[nm + r], to, scale, SCALE, 1);
                    }
                }
            }
        }
    }
}

/* =================================================================== per-element work */

typedef struct
{
    double *a, *b, *c, *d, *e, *f, *tmp; /* nqTot-sized scratch (+ tmp) */
} scratch;

static scratch scratch_new(int nqTot, int nmTot)
{
    scratch s;
    size_t n = (size_t)(nqTot > nmTot ? nqTot : nmTot) + 64;
    s.a = (double *)malloc(sizeof(double) * n);
    s.b = (double *)malloc(sizeof(double) * n);
    s.c = (double *)malloc(sizeof(double) * n);
    s.d = (double *)malloc(sizeof(double) * n);
    s.e = (double *)malloc(sizeof(double) * n);
    s.f = (double *)malloc(sizeof(double) * n);
    s.tmp = (double *)malloc(sizeof(double) * n);
    return s;
}
static void scratch_free(scratch *s)
{
    free(s->a); free(s->b); free(s->c); free(s->d); free(s->e); free(s->f); free(s->tmp);
}

static void bwd_one(const mfo_elem *e, const double *in, double *out, scratch *s)
{
    switch (e->shape)
    {
        case MFO_SEG:
        {
            /* BwdTransKernels.hpp:14-33 */
            int i, p, nq = e->nq[0], nm = e->nm;
            for (i = 0; i < nq; ++i)
            {
                double t = in[0] * e->b[0][i];
                for (p = 1; p < nm; ++p) t += in[p] * e->b[0][p * nq + i];
                out[i] = t;
            }
            break;
        }
        case MFO_QUAD: k_bwd_quad(e->nm, e->nq[0], in, e->b[0], e->b[1], s->a, out); break;
        case MFO_TRI: k_bwd_tri(e->nm, e->nq[0], e->nq[1], in, e->b[0], e->b[1], s->a, out); break;
        case MFO_HEX: k_bwd_hex(e->nm, e->nq[0], in, e->b[0], e->b[1], e->b[2], s->a, s->b, out); break;
        case MFO_PRISM:
            k_bwd_prism(e->nm, e->nq[0], e->nq[1], e->nq[2], in, e->b[0], e->b[1], e->b[2], s->a, s->b, out);
            break;
        case MFO_TET:
            k_bwd_tet(e->nm, e->nq[0], e->nq[1], e->nq[2], in, e->b[0], e->b[1], e->b[2], s->a, s->b, out);
            break;
        case MFO_PYR:
            k_bwd_pyr(e->nm, e->nq[0], e->nq[1], e->nq[2], in, e->b[0], e->b[1], e->b[2], s->a, s->b, out);
            break;
    }
}

/* generic inner product with selectable basis arrays (bdata or dbdata per direction) */
static void ip_one(const mfo_elem *e, const double *in, const double *B0, const double *B1, const double *B2,
                   const double *jac, int DEF, double *out, double scale, int SCALE, int APPEND, scratch *s)
{
    switch (e->shape)
    {
        case MFO_SEG:
        {
            /* IProductKernels.hpp:39-75 */
            int i, p, nq = e->nq[0], nm = e->nm;
            for (p = 0; p < nm; ++p)
            {
                double sum = 0.0;
                for (i = 0; i < nq; ++i)
                {
                    double prod = in[i] * B0[p * nq + i] * (DEF ? jac[i] : jac[0]);
                    sum += prod * e->ws[0][i];
                }
                scale_append(&out[p], sum, scale, SCALE, APPEND);
            }
            break;
        }
        case MFO_QUAD:
            k_ip_quad(e->nm, e->nq[0], in, B0, B1, e->ws[0], e->ws[1], jac, DEF, s->a, out, scale, SCALE, APPEND);
            break;
        case MFO_TRI:
            k_ip_tri(e->nm, e->nq[0], e->nq[1], in, B0, B1, e->ws[0], e->ws[1], jac, DEF, s->a, out, scale, SCALE,
                     APPEND);
            break;
        case MFO_HEX:
            k_ip_hex(e->nm, e->nq[0], in, B0, B1, B2, e->ws[0], e->ws[1], e->ws[2], jac, DEF, s->a, s->b, out,
                     scale, SCALE, APPEND);
            break;
        case MFO_PRISM:
            k_ip_prism(e->nm, e->nq[0], e->nq[1], e->nq[2], in, B0, B1, B2, e->ws[0], e->ws[1], e->ws[2], jac,
                       DEF, s->a, s->b, s->tmp, out, scale, SCALE, APPEND);
            break;
        case MFO_TET:
            k_ip_tet(e->nm, e->nq[0], e->nq[1], e->nq[2], in, B0, B1, B2, e->ws[0], e->ws[1], e->ws[2], jac, DEF,
                     s->a, out, scale, SCALE, APPEND);
            break;
        case MFO_PYR:
            k_ip_pyr(e->nm, e->nq[0], e->nq[1], e->nq[2], in, B0, B1, B2, e->ws[0], e->ws[1], e->ws[2], jac, DEF,
                     s->a, s->b, out, scale, SCALE, APPEND);
            break;
    }
}

#define DFV(n) (DEF ? df[(size_t)(n) * dfs + pt] : df[(size_t)(n) * dfs])

/* PhysDeriv for one element; df indexed df[n*dfs + (elmt*nq + pt | elmt)], caller offsets df to the element */
static void pd_one(const mfo_elem *e, const double *in, const double *df, size_t dfs, int DEF, double *o0,
                   double *o1, double *o2, scratch *s)
{
    const int nq0 = e->nq[0], nq1 = e->nq[1], nq2 = e->nq[2];
    int i, j, k, pt;
    if (e->dim == 1)
    {
        /* PhysDerivKernels.hpp:13-38 + PhysDeriv.h:60-250 (one output per space dimension) */
        double *outs[3] = {o0, o1, o2};
        for (i = 0; i < nq0; ++i)
        {
            double d = 0.0;
            int c;
            for (k = 0; k < nq0; ++k) d += e->D[0][k * nq0 + i] * in[k];
            pt = i;
            for (c = e->coordim - 1; c >= 0; --c) outs[c][i] = d * DFV(c);
        }
        return;
    }
    if (e->dim == 2)
    {
        k_dtensor2(nq0, nq1, in, e->D[0], e->D[1], o0, o1);
        for (j = 0, pt = 0; j < nq1; ++j)
        {
            /* PhysDerivKernels.hpp:186 (Tri) */
            double xfrm0 = e->shape == MFO_TRI ? 2.0 / (1.0 - e->z[1][j]) : 0.0;
            for (i = 0; i < nq0; ++i, ++pt)
            {
                double d0 = o0[pt], d1 = o1[pt], r0, r1;
                if (e->shape == MFO_TRI)
                {
                    double xfrm1;
                    d0    = xfrm0 * o0[pt];
                    xfrm1 = 0.5 * (1.0 + e->z[0][i]);
                    d1 += d0 * xfrm1;
                }
                r0 = d0 * DFV(0); r0 += d1 * DFV(1);
                r1 = d0 * DFV(2); r1 += d1 * DFV(3);
                o0[pt] = r0;
                o1[pt] = r1;
            }
        }
        return;
    }
    if (e->shape == MFO_TET)
    {
        /* PhysDerivKernels.hpp:555-696 */
        double *f0 = s->a, *f1 = s->b, *f2 = s->c;
        k_dtensor3(nq0, nq1, nq2, in, e->D[0], e->D[1], e->D[2], f0, f1, f2);
        for (k = 0, pt = 0; k < nq2; ++k)
        {
            double x2 = 2.0 / (1.0 - e->z[2][k]);
            for (j = 0; j < nq1; ++j)
            {
                double x1 = 2.0 / (1.0 - e->z[1][j]);
                double x  = x1 * x2;
                for (i = 0; i < nq0; ++i, ++pt)
                {
                    double d0 = x * f0[pt];
                    o0[pt]    = d0;
                    f0[pt]    = d0;
                }
            }
        }
        for (k = 0, pt = 0; k < nq2; ++k)
        {
            double x2 = 2.0 / (1.0 - e->z[2][k]);
            for (j = 0; j < nq1; ++j)
                for (i = 0; i < nq0; ++i, ++pt)
                {
                    double x0 = 0.5 * (1.0 + e->z[0][i]);
                    double a  = x0 * f0[pt], d1;
                    f0[pt]    = a;
                    d1        = f1[pt];
                    d1        = x2 * d1;
                    o1[pt]    = a + d1;
                    f1[pt]    = d1;
                }
        }
        for (k = 0, pt = 0; k < nq2; ++k)
            for (j = 0; j < nq1; ++j)
            {
                double x1 = 0.5 * (1.0 + e->z[1][j]);
                for (i = 0; i < nq0; ++i, ++pt)
                {
                    double o = f0[pt];
                    o += f1[pt] * x1;
                    o      = o + f2[pt];
                    o2[pt] = o;
                }
            }
    }
    else
    {
        k_dtensor3(nq0, nq1, nq2, in, e->D[0], e->D[1], e->D[2], o0, o1, o2);
    }
    for (k = 0, pt = 0; k < nq2; ++k)
    {
        double xe2 = (e->shape == MFO_PRISM || e->shape == MFO_PYR) ? 2.0 / (1.0 - e->z[2][k]) : 0.0;
        for (j = 0; j < nq1; ++j)
            for (i = 0; i < nq0; ++i, ++pt)
            {
                double d0 = o0[pt], d1 = o1[pt], d2 = o2[pt], r0, r1, r2;
                if (e->shape == MFO_PYR)
                {
                    /* PhysDerivKernels.hpp:505-527 */
                    double xe1 = 0.5 * (1 + e->z[1][j]), xe0;
                    d0  = o0[pt] * xe2;
                    d1  = o1[pt] * xe2;
                    xe0 = 0.5 * (1 + e->z[0][i]);
                    d2 += xe0 * d0;
                    d2 += xe1 * d1;
                }
                else if (e->shape == MFO_PRISM)
                {
                    /* PhysDerivKernels.hpp:417-428 */
                    double xe0;
                    d0  = o0[pt] * xe2;
                    xe0 = 0.5 * (1.0 + e->z[0][i]);
                    d2 += xe0 * d0;
                }
                r0 = d0 * DFV(0); r0 += d1 * DFV(1); r0 += d2 * DFV(2);
                r1 = d0 * DFV(3); r1 += d1 * DFV(4); r1 += d2 * DFV(5);
                r2 = d0 * DFV(6); r2 += d1 * DFV(7); r2 += d2 * DFV(8);
                o0[pt] = r0;
                o1[pt] = r1;
                o2[pt] = r2;
            }
    }
}

/* Helmholtz for one element: Helmholtz.h:138-275 (Quad), 506-635 (Tri), 764-993 (Hex),
 * 1291-1458 (Prism), 2266-2448 (Tet) */
static void helm_one(const mfo_elem *e, const double *in, const double *jac, const double *df, size_t dfs,
                     int DEF, double lambda, double *out, scratch *s)
{
    const int nq0 = e->nq[0], nq1 = e->nq[1], nq2 = e->dim == 3 ? e->nq[2] : 1;
    double *bwd = s->c, *g0 = s->d, *g1 = s->e, *g2 = s->f;
    int i, j, k, pt;
    bwd_one(e, in, bwd, s);
    ip_one(e, bwd, e->b[0], e->b[1], e->b[2], jac, DEF, out, lambda, 1, 0, s);
    if (e->dim == 2)
    {
        k_dtensor2(nq0, nq1, bwd, e->D[0], e->D[1], g0, g1);
        for (j = 0, pt = 0; j < nq1; ++j)
            for (i = 0; i < nq0; ++i, ++pt)
            {
                double df0 = DFV(0), df1 = DFV(1), df2 = DFV(2), df3 = DFV(3);
                double m00, m01, m11, d0 = g0[pt], d1 = g1[pt], t;
                if (e->shape == MFO_QUAD)
                {
                    m00 = df0 * df0; m00 += df2 * df2;
                    m01 = df0 * df1; m01 += df2 * df3;
                    m11 = df1 * df1; m11 += df3 * df3;
                }
                else
                {
                    double h1j = e->h1[j], h0i = e->h0[i];
                    m00 = h1j * (df0 + h0i * df1);
                    m01 = m00 * df1;
                    m00 = m00 * m00;
                    t   = h1j * (df2 + h0i * df3);
                    m01 += t * df3;
                    m00 += t * t;
                    m11 = df1 * df1; m11 += df3 * df3;
                }
                t = m00 * d0; t += m01 * d1; bwd[pt] = t;
                t = m01 * d0; t += m11 * d1; g0[pt] = t;
            }
        ip_one(e, bwd, e->db[0], e->b[1], NULL, jac, DEF, out, 1.0, 0, 1, s);
        ip_one(e, g0, e->b[0], e->db[1], NULL, jac, DEF, out, 1.0, 0, 1, s);
        return;
    }
    k_dtensor3(nq0, nq1, nq2, bwd, e->D[0], e->D[1], e->D[2], g0, g1, g2);
    for (k = 0, pt = 0; k < nq2; ++k)
        for (j = 0; j < nq1; ++j)
            for (i = 0; i < nq0; ++i, ++pt)
            {
                double df0 = DFV(0), df1 = DFV(1), df2 = DFV(2), df3 = DFV(3), df4 = DFV(4), df5 = DFV(5),
                       df6 = DFV(6), df7 = DFV(7), df8 = DFV(8);
                double m00, m01, m02, m11, m12, m22, d0 = g0[pt], d1 = g1[pt], d2 = g2[pt], t;
                if (e->shape == MFO_HEX)
                {
                    m00 = df0 * df0; m00 += df3 * df3; m00 += df6 * df6;
                    m01 = df0 * df1; m01 += df3 * df4; m01 += df6 * df7;
                    m02 = df0 * df2; m02 += df3 * df5; m02 += df6 * df8;
                    m11 = df1 * df1; m11 += df4 * df4; m11 += df7 * df7;
                    m12 = df1 * df2; m12 += df4 * df5; m12 += df7 * df8;
                    m22 = df2 * df2; m22 += df5 * df5; m22 += df8 * df8;
                }
                else if (e->shape == MFO_PRISM)
                {
                    double h1 = e->h1[k], h0 = e->h0[i];
                    double t1 = h1 * (h0 * df2 + df0), t2 = h1 * (h0 * df5 + df3), t3 = h1 * (h0 * df8 + df6);
                    m00 = t1 * t1; m00 += t2 * t2; m00 += t3 * t3;       /* g0 */
                    m01 = df1 * t1; m01 += df4 * t2; m01 += df7 * t3;    /* g3 */
                    m02 = df2 * t1; m02 += df5 * t2; m02 += df8 * t3;    /* g4 */
                    m11 = df1 * df1; m11 += df4 * df4; m11 += df7 * df7; /* g1 */
                    m22 = df2 * df2; m22 += df5 * df5; m22 += df8 * df8; /* g2 */
                    m12 = df1 * df2; m12 += df4 * df5; m12 += df7 * df8; /* g5 */
                }
                else if (e->shape == MFO_PYR)
                {
                    /* Helmholtz.h:1845-1905 */
                    double h2 = e->h2[k], h1 = e->h1[j], h0 = e->h0[i];
                    double h1h2 = h1 * h2, h0h2 = h0 * h2;
                    double t0, t1, t2, t3, t4, t5;
                    t0 = h2 * df0; t0 += h0h2 * df2;
                    t1 = h2 * df3; t1 += h0h2 * df5;
                    t2 = h2 * df6; t2 += h0h2 * df8;
                    t3 = h2 * df1; t3 += h1h2 * df2;
                    t4 = h2 * df4; t4 += h1h2 * df5;
                    t5 = h2 * df7; t5 += h1h2 * df8;
                    m00 = t0 * t0; m00 += t1 * t1; m00 += t2 * t2;       /* g0 */
                    m11 = t3 * t3; m11 += t4 * t4; m11 += t5 * t5;       /* g1 */
                    m22 = df2 * df2; m22 += df5 * df5; m22 += df8 * df8; /* g2 */
                    m01 = t0 * t3; m01 += t1 * t4; m01 += t2 * t5;       /* g3 */
                    m02 = df2 * t0; m02 += df5 * t1; m02 += df8 * t2;    /* g4 */
                    m12 = df2 * t3; m12 += df5 * t4; m12 += df8 * t5;    /* g5 */
                }
                else
                {
                    double h3 = e->h3[k], h1 = e->h1[j], h2 = e->h2[j];
                    double h2h3 = h2 * h3, h1h3 = h1 * h3, h0h2h3 = e->h0[i] * h2h3;
                    double t1, t2, t3, t4, t5, t6;
                    t1 = h0h2h3 * (df1 + df2); t1 += df0 * h2h3;
                    t2 = h0h2h3 * (df4 + df5); t2 += df3 * h2h3;
                    t3 = h0h2h3 * (df7 + df8); t3 += df6 * h2h3;
                    m00 = t1 * t1; m00 += t2 * t2; m00 += t3 * t3;    /* g0 */
                    m02 = df2 * t1; m02 += df5 * t2; m02 += df8 * t3; /* g4 */
                    t4 = df1 * h3; t4 += df2 * h1h3;
                    t5 = df4 * h3; t5 += df5 * h1h3;
                    t6 = df7 * h3; t6 += df8 * h1h3;
                    m01 = t1 * t4; m01 += t2 * t5; m01 += t3 * t6;       /* g3 */
                    m11 = t4 * t4; m11 += t5 * t5; m11 += t6 * t6;       /* g1 */
                    m12 = df2 * t4; m12 += df5 * t5; m12 += df8 * t6;    /* g5 */
                    m22 = df2 * df2; m22 += df5 * df5; m22 += df8 * df8; /* g2 */
                }
                t = m00 * d0; t += m01 * d1; t += m02 * d2; g0[pt] = t;
                t = m01 * d0; t += m11 * d1; t += m12 * d2; g1[pt] = t;
                t = m02 * d0; t += m12 * d1; t += m22 * d2; g2[pt] = t;
            }
    ip_one(e, g0, e->db[0], e->b[1], e->b[2], jac, DEF, out, 1.0, 0, 1, s);
    ip_one(e, g1, e->b[0], e->db[1], e->b[2], jac, DEF, out, 1.0, 0, 1, s);
    ip_one(e, g2, e->b[0], e->b[1], e->db[2], jac, DEF, out, 1.0, 0, 1, s);
}

/* ======================================================================= public ops */

void mfo_bwdtrans(const mfo_elem *e, int nElmt, const double *in, double *out)
{
#pragma omp parallel num_threads(g_threads)
    {
        scratch s = scratch_new(e->nqTot, e->nmTot);
        int el;
#pragma omp for schedule(static)
        for (el = 0; el < nElmt; ++el)
            bwd_one(e, in + (size_t)el * e->nmTot, out + (size_t)el * e->nqTot, &s);
        scratch_free(&s);
    }
}

void mfo_iproduct(const mfo_elem *e, int nElmt, int DEF, const double *jac, const double *in, double *out)
{
#pragma omp parallel num_threads(g_threads)
    {
        scratch s = scratch_new(e->nqTot, e->nmTot);
        int el;
#pragma omp for schedule(static)
        for (el = 0; el < nElmt; ++el)
            ip_one(e, in + (size_t)el * e->nqTot, e->b[0], e->b[1], e->b[2],
                   DEF ? jac + (size_t)el * e->nqTot : jac + el, DEF, out + (size_t)el * e->nmTot, 1.0, 0, 0, &s);
        scratch_free(&s);
    }
}

void mfo_physderiv(const mfo_elem *e, int nElmt, int DEF, const double *df, const double *in, double *o0,
                   double *o1, double *o2)
{
    const size_t dfs = DEF ? (size_t)nElmt * e->nqTot : (size_t)nElmt;
#pragma omp parallel num_threads(g_threads)
    {
        scratch s = scratch_new(e->nqTot, e->nmTot);
        int el;
#pragma omp for schedule(static)
        for (el = 0; el < nElmt; ++el)
        {
            size_t off = (size_t)el * e->nqTot;
            pd_one(e, in + off, DEF ? df + off : df + el, dfs, DEF, o0 + off, o1 ? o1 + off : NULL,
                   o2 ? o2 + off : NULL, &s);
        }
        scratch_free(&s);
    }
}

void mfo_helmholtz(const mfo_elem *e, int nElmt, int DEF, const double *jac, const double *df, double lambda,
                   const double *in, double *out)
{
    const size_t dfs = DEF ? (size_t)nElmt * e->nqTot : (size_t)nElmt;
#pragma omp parallel num_threads(g_threads)
    {
        scratch s = scratch_new(e->nqTot, e->nmTot);
        int el;
#pragma omp for schedule(static)
        for (el = 0; el < nElmt; ++el)
        {
            size_t qoff = (size_t)el * e->nqTot, moff = (size_t)el * e->nmTot;
            helm_one(e, in + moff, DEF ? jac + qoff : jac + el, DEF ? df + qoff : df + el, dfs, DEF, lambda,
                     out + moff, &s);
        }
        scratch_free(&s);
    }
}

/* IProductWRTDerivBase.h:1232-1345 (Hex) and the Quad analogue */
int mfo_iproductwrtderivbase(const mfo_elem *e, int nElmt, int DEF, const double *jac, const double *df,
                             const double *in0, const double *in1, const double *in2, double *out)
{
    const size_t dfs = DEF ? (size_t)nElmt * e->nqTot : (size_t)nElmt;
#pragma omp parallel num_threads(g_threads)
    {
        scratch s = scratch_new(e->nqTot, e->nmTot);
        int el;
#pragma omp for schedule(static)
        for (el = 0; el < nElmt; ++el)
        {
            size_t qoff = (size_t)el * e->nqTot, moff = (size_t)el * e->nmTot;
            const double *dfe = DEF ? df + qoff : df + el, *jc = DEF ? jac + qoff : jac + el;
            double *t0 = s.c, *t1 = s.d, *t2 = s.e;
            int pt;
#define DFE(n) (DEF ? dfe[(size_t)(n) * dfs + pt] : dfe[(size_t)(n) * dfs])
            if (e->dim == 1)
            {
                /* IProductWRTDerivBase.h:182-365: t = sum_c df[c] in_c, then one IProduct with dbdata.  The
                 * reference's regular coordim-3 branch reads df[1] for the third factor (:321-323); it is
                 * restated as written there. */
                for (pt = 0; pt < e->nqTot; ++pt)
                {
                    double v = DFE(0) * in0[qoff + pt];
                    if (e->coordim >= 2) v = v + DFE(1) * in1[qoff + pt];
                    if (e->coordim == 3) v = v + (DEF ? DFE(2) : DFE(1)) * in2[qoff + pt];
                    t0[pt] = v;
                }
                ip_one(e, t0, e->db[0], NULL, NULL, jc, DEF, out + moff, 1.0, 0, 0, &s);
            }
            else if (e->dim == 3)
            {
                for (pt = 0; pt < e->nqTot; ++pt)
                {
                    double a = in0[qoff + pt], b = in1[qoff + pt], c = in2[qoff + pt];
                    double v0 = DFE(0) * a + DFE(3) * b + DFE(6) * c;
                    double v1 = DFE(1) * a + DFE(4) * b + DFE(7) * c;
                    double v2 = DFE(2) * a + DFE(5) * b + DFE(8) * c;
                    if (e->shape == MFO_PRISM)
                    {
                        /* IProductWRTDerivBase.h:1697-1733 */
                        int i = pt % e->nq[0], k = pt / (e->nq[0] * e->nq[1]);
                        double f0 = 2.0 / (1.0 - e->z[2][k]), hf1 = 0.5 * (1.0 + e->z[0][i]), f1t2;
                        v0 *= f0;
                        f1t2 = hf1 * v2;
                        v0 += f1t2 * f0;
                    }
                    else if (e->shape == MFO_PYR)
                    {
                        /* IProductWRTDerivBase.h:2121-2168 */
                        int i = pt % e->nq[0], j = (pt / e->nq[0]) % e->nq[1], k = pt / (e->nq[0] * e->nq[1]);
                        double f0 = 2.0 / (1.0 - e->z[2][k]), hf2 = 0.5 * (1.0 + e->z[1][j]);
                        double hf1 = 0.5 * (1.0 + e->z[0][i]), f1t2;
                        v0 *= f0;
                        f1t2 = hf1 * v2;
                        v0 += f1t2 * f0;
                        v1 *= f0;
                        f1t2 = hf2 * v2;
                        v1 += f1t2 * f0;
                    }
                    else if (e->shape == MFO_TET)
                    {
                        /* IProductWRTDerivBase.h:2551-2603 */
                        int i = pt % e->nq[0], j = (pt / e->nq[0]) % e->nq[1], k = pt / (e->nq[0] * e->nq[1]);
                        double z1 = e->z[1][j];
                        double f2 = 2.0 / (1.0 - e->z[2][k]), f3 = 0.5 * (1.0 + z1), f0 = 2.0 * f2 / (1.0 - z1);
                        double f1 = 0.5 * (1.0 + e->z[0][i]);
                        v0 += (v1 + v2) * f1;
                        v0 *= f0;
                        v1 += v2 * f3;
                        v1 *= f2;
                    }
                    t0[pt] = v0; t1[pt] = v1; t2[pt] = v2;
                }
                ip_one(e, t0, e->db[0], e->b[1], e->b[2], jc, DEF, out + moff, 1.0, 0, 0, &s);
                ip_one(e, t1, e->b[0], e->db[1], e->b[2], jc, DEF, out + moff, 1.0, 0, 1, &s);
                ip_one(e, t2, e->b[0], e->b[1], e->db[2], jc, DEF, out + moff, 1.0, 0, 1, &s);
            }
            else
            {
                for (pt = 0; pt < e->nqTot; ++pt)
                {
                    double a = in0[qoff + pt], b = in1[qoff + pt];
                    double v0 = DFE(0) * a + DFE(2) * b;
                    double v1 = DFE(1) * a + DFE(3) * b;
                    if (e->shape == MFO_TRI)
                    {
                        /* IProductWRTDerivBase.h:1006-1036 */
                        int i = pt % e->nq[0], j = pt / e->nq[0];
                        double f0 = 2.0 / (1.0 - e->z[1][j]), hf1 = 0.5 * (1.0 + e->z[0][i]), c1;
                        v0 *= f0;
                        c1 = hf1 * v1;
                        v0 += c1 * f0;
                    }
                    t0[pt] = v0; t1[pt] = v1;
                }
                ip_one(e, t0, e->db[0], e->b[1], NULL, jc, DEF, out + moff, 1.0, 0, 0, &s);
                ip_one(e, t1, e->b[0], e->db[1], NULL, jc, DEF, out + moff, 1.0, 0, 1, &s);
            }
#undef DFE
        }
        scratch_free(&s);
    }
    return 0;
}

/* ======================================================================= assembly map */

/* AssemblyMapCG.cpp:2853-2876 + Vmath::Gathr (Vmath.hpp:217-230) */
void mfo_global_to_local(int nLocal, const int *map, const double *sign, const double *glob, double *loc)
{
    int i;
    if (sign)
        for (i = 0; i < nLocal; ++i) loc[i] = sign[i] * glob[map[i]];
    else
        for (i = 0; i < nLocal; ++i) loc[i] = glob[map[i]];
}

/* AssemblyMapCG.cpp:2885-2910 + Vmath::Assmb: zero, then sequential scatter-add */
void mfo_assemble(int nLocal, int nGlobal, const int *map, const double *sign, const double *loc, double *glob)
{
    int i;
    for (i = 0; i < nGlobal; ++i) glob[i] = 0.0;
    if (sign)
        for (i = 0; i < nLocal; ++i) glob[map[i]] += sign[i] * loc[i];
    else
        for (i = 0; i < nLocal; ++i) glob[map[i]] += loc[i];
}

/* ================================================================================ CG */

static double dot(int n, const double *a, const double *b)
{
    double s = 0.0;
    int i;
    for (i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

/* NekLinSysIterCG.cpp:104-265 (serial: m_map == 1 everywhere, AllReduce is the identity),
 * mat-vec = GlobalLinSysIterativeFull::v_DoMatrixMultiply (GlobalLinSysIterativeFull.cpp:215-257),
 * preconditioner = diagonal scaling (PreconditionerDiagonal.cpp) or identity (Null). */
int mfo_cg_helmholtz(const mfo_elem *e, int nElmt, int DEF, const double *jac, const double *df, double lambda,
                     int nLocal, int nGlobal, int nDir, const int *map, const double *sign,
                     const double *invdiag, const double *rhs, double *x, double tol, int maxiter,
                     double *final_eps)
{
    const int nNonDir = nGlobal - nDir;
    double *w_A = (double *)calloc(nGlobal, sizeof(double)), *s_A = (double *)calloc(nGlobal, sizeof(double));
    double *p_A = (double *)calloc(nNonDir, sizeof(double)), *r_A = (double *)calloc(nNonDir, sizeof(double));
    double *q_A = (double *)calloc(nNonDir, sizeof(double));
    double *lin = (double *)malloc(sizeof(double) * nLocal), *lout = (double *)malloc(sizeof(double) * nLocal);
    double alpha, beta, rho, rho_new, mu, eps, rhs_mag;
    int k = 0, i, its = 0;
#define PRECON()                                                                                             \
    for (i = 0; i < nNonDir; ++i) w_A[nDir + i] = invdiag ? r_A[i] * invdiag[i] : r_A[i]
#define MATVEC()                                                                                             \
    do                                                                                                       \
    {                                                                                                        \
        mfo_global_to_local(nLocal, map, sign, w_A, lin);                                                    \
        mfo_helmholtz(e, nElmt, DEF, jac, df, lambda, lin, lout);                                            \
        mfo_assemble(nLocal, nGlobal, map, sign, lout, s_A);                                                 \
    } while (0)

    memcpy(r_A, rhs + nDir, sizeof(double) * nNonDir);
    memset(x + nDir, 0, sizeof(double) * nNonDir);
    eps     = dot(nNonDir, r_A, r_A);
    rhs_mag = dot(nGlobal, rhs, rhs); /* Set_Rhs_Magnitude, NekLinSysIter.cpp:128-153 */
    rhs_mag = rhs_mag > 1e-6 ? rhs_mag : 1.0;
    if (eps < tol * tol * rhs_mag) goto done;

    PRECON();
    MATVEC();
    rho   = dot(nNonDir, r_A, w_A + nDir);
    mu    = dot(nNonDir, s_A + nDir, w_A + nDir);
    beta  = 0.0;
    alpha = rho / mu;
    its   = 1;
    for (;;)
    {
        if (k >= maxiter) break;
        for (i = 0; i < nNonDir; ++i) p_A[i] = beta * p_A[i] + w_A[nDir + i];
        for (i = 0; i < nNonDir; ++i) q_A[i] = beta * q_A[i] + s_A[nDir + i];
        for (i = 0; i < nNonDir; ++i) x[nDir + i] = alpha * p_A[i] + x[nDir + i];
        for (i = 0; i < nNonDir; ++i) r_A[i] = -alpha * q_A[i] + r_A[i];
        PRECON();
        MATVEC();
        rho_new = dot(nNonDir, r_A, w_A + nDir);
        mu      = dot(nNonDir, s_A + nDir, w_A + nDir);
        eps     = dot(nNonDir, r_A, r_A);
        its++;
        if (eps < tol * tol * rhs_mag) break;
        beta  = rho_new / rho;
        alpha = rho_new / (mu - rho_new * beta / alpha);
        rho   = rho_new;
        k++;
    }
done:
    if (final_eps) *final_eps = eps;
    free(w_A); free(s_A); free(p_A); free(r_A); free(q_A); free(lin); free(lout);
    return its;
#undef PRECON
#undef MATVEC
}

/* ======================================================================= HelmSolve chain */

/* ContField::v_HelmSolve -> GlobalSolve -> GlobalLinSysIterativeFull::v_Solve -> DoConjugateGradient, then
 * BwdTrans of the result (ContField.cpp:878-945, 516-535; GlobalLinSysIterativeFull.cpp:110-211;
 * NekLinSysIterCG.cpp:104-265), the elemental operators supplied as callbacks so that the same driver runs
 * this file's operators (the checker) and the reference's own kernels from oracle/_ref (the CPU baseline).
 * The global vector loops are OpenMP loops over g_threads threads -- the stand-in for the reference's MPI
 * ranks each updating its own partition; one thread is the reference's sequential Vmath.  Assemble is the
 * gather over a transposed map (ascending local index per global DOF = the order of Vmath::Assmb). */
struct mfo_chain
{
    int nLocal, nGlobal, nDir, nPhys;
    const int *map;
    const double *sign, *invdiag;
    int *rowptr, *col;
    double *w, *s, *p, *r, *q, *lin, *lout, *wsp, *rhs, *glob;
};

mfo_chain *mfo_chain_create(int nLocal, int nGlobal, int nDir, int nPhys, const int *map, const double *sign,
                            const double *invdiag)
{
    mfo_chain *c = (mfo_chain *)calloc(1, sizeof(mfo_chain));
    int i, g, *cur;
    c->nLocal = nLocal; c->nGlobal = nGlobal; c->nDir = nDir; c->nPhys = nPhys;
    c->map = map; c->sign = sign; c->invdiag = invdiag;
    c->rowptr = (int *)calloc((size_t)nGlobal + 1, sizeof(int));
    c->col    = (int *)malloc(sizeof(int) * (size_t)(nLocal ? nLocal : 1));
    for (i = 0; i < nLocal; ++i) c->rowptr[map[i] + 1]++;
    for (g = 0; g < nGlobal; ++g) c->rowptr[g + 1] += c->rowptr[g];
    cur = (int *)malloc(sizeof(int) * (size_t)(nGlobal ? nGlobal : 1));
    memcpy(cur, c->rowptr, sizeof(int) * (size_t)nGlobal);
    for (i = 0; i < nLocal; ++i) c->col[cur[map[i]]++] = i;
    free(cur);
#define CH_ALLOC(f, n) c->f = (double *)calloc((size_t)(n) + 1, sizeof(double))
    CH_ALLOC(w, nGlobal); CH_ALLOC(s, nGlobal); CH_ALLOC(rhs, nGlobal); CH_ALLOC(glob, nGlobal);
    CH_ALLOC(p, nGlobal - nDir); CH_ALLOC(r, nGlobal - nDir); CH_ALLOC(q, nGlobal - nDir);
    CH_ALLOC(lin, nLocal); CH_ALLOC(lout, nLocal); CH_ALLOC(wsp, nLocal);
#undef CH_ALLOC
    return c;
}

void mfo_chain_destroy(mfo_chain *c)
{
    if (!c) return;
    free(c->rowptr); free(c->col); free(c->w); free(c->s); free(c->p); free(c->r); free(c->q);
    free(c->lin); free(c->lout); free(c->wsp); free(c->rhs); free(c->glob);
    free(c);
}

static void ch_g2l(const mfo_chain *c, const double *glob, double *loc)
{
    int i;
#pragma omp parallel for schedule(static) num_threads(g_threads)
    for (i = 0; i < c->nLocal; ++i) loc[i] = c->sign ? c->sign[i] * glob[c->map[i]] : glob[c->map[i]];
}
static void ch_assemble(const mfo_chain *c, const double *loc, double *glob)
{
    int g;
#pragma omp parallel for schedule(static) num_threads(g_threads)
    for (g = 0; g < c->nGlobal; ++g)
    {
        double sum = 0.0;
        int k;
        for (k = c->rowptr[g]; k < c->rowptr[g + 1]; ++k)
        {
            const int i = c->col[k];
            sum += c->sign ? c->sign[i] * loc[i] : loc[i];
        }
        glob[g] = sum;
    }
}
static double ch_dot(int n, const double *a, const double *b)
{
    double sum = 0.0;
    int i;
#pragma omp parallel for schedule(static) reduction(+ : sum) num_threads(g_threads)
    for (i = 0; i < n; ++i) sum += a[i] * b[i];
    return sum;
}

/* forcing [nPhys]; inout [nLocal] (Dirichlet values + initial guess in, solution out); phys_out [nPhys] or NULL.
 * Returns m_totalIterations, or -its when the loop counter reached maxiter (the reference's efatal). */
int mfo_chain_helmsolve(mfo_chain *c, mfo_elop_fn iprod, void *ci, mfo_elop_fn helm, void *ch, mfo_elop_fn bwd, void *cb,
                        const double *forcing, double *inout, double *phys_out, double tol, int maxiter,
                        double *final_eps)
{
    const int nL = c->nLocal, nG = c->nGlobal, nD = c->nDir, nN = nG - nD;
    double *w = c->w, *s = c->s, *p = c->p, *r = c->r, *q = c->q, *x = c->glob;
    double alpha, beta, rho, rho_new, mu, eps, rhs_mag;
    int i, k = 0, its = 0, capped = 0;
    /* ContField.cpp:894-900 */
    iprod(ci, forcing, c->wsp);
    if (nD > 0)
    {
        /* GlobalLinSysIterativeFull.cpp:161-181 */
        helm(ch, inout, c->lout);
#pragma omp parallel for schedule(static) num_threads(g_threads)
        for (i = 0; i < nL; ++i) c->lout[i] = (-c->wsp[i]) - c->lout[i];
    }
    else
    {
#pragma omp parallel for schedule(static) num_threads(g_threads)
        for (i = 0; i < nL; ++i) c->lout[i] = -c->wsp[i];
    }
    ch_assemble(c, c->lout, c->rhs);
    /* DoConjugateGradient */
#pragma omp parallel for schedule(static) num_threads(g_threads)
    for (i = 0; i < nG; ++i) { x[i] = 0.0; w[i] = 0.0; s[i] = 0.0; }
#pragma omp parallel for schedule(static) num_threads(g_threads)
    for (i = 0; i < nN; ++i) { r[i] = c->rhs[nD + i]; p[i] = 0.0; q[i] = 0.0; }
    eps     = ch_dot(nN, r, r);
    rhs_mag = ch_dot(nG, c->rhs, c->rhs);
    rhs_mag = rhs_mag > 1e-6 ? rhs_mag : 1.0;
    if (!(eps < tol * tol * rhs_mag))
    {
#define CH_PRECON()                                                                                          \
    _Pragma("omp parallel for schedule(static) num_threads(g_threads)")                                      \
    for (i = 0; i < nN; ++i) w[nD + i] = c->invdiag ? r[i] * c->invdiag[i] : r[i]
#define CH_MATVEC()                                                                                          \
    do                                                                                                       \
    {                                                                                                        \
        ch_g2l(c, w, c->lin);                                                                                \
        helm(ch, c->lin, c->lout);                                                                           \
        ch_assemble(c, c->lout, s);                                                                          \
    } while (0)
        CH_PRECON();
        CH_MATVEC();
        rho   = ch_dot(nN, r, w + nD);
        mu    = ch_dot(nN, s + nD, w + nD);
        beta  = 0.0;
        alpha = rho / mu;
        its   = 1;
        for (;;)
        {
            if (k >= maxiter) { capped = 1; break; }
#pragma omp parallel for schedule(static) num_threads(g_threads)
            for (i = 0; i < nN; ++i)
            {
                p[i]      = beta * p[i] + w[nD + i];
                q[i]      = beta * q[i] + s[nD + i];
                x[nD + i] = alpha * p[i] + x[nD + i];
                r[i]      = -alpha * q[i] + r[i];
            }
            CH_PRECON();
            CH_MATVEC();
            rho_new = ch_dot(nN, r, w + nD);
            mu      = ch_dot(nN, s + nD, w + nD);
            eps     = ch_dot(nN, r, r);
            its++;
            if (eps < tol * tol * rhs_mag) break;
            beta  = rho_new / rho;
            alpha = rho_new / (mu - rho_new * beta / alpha);
            rho   = rho_new;
            k++;
        }
#undef CH_PRECON
#undef CH_MATVEC
    }
    /* GlobalLinSysIterativeFull.cpp:190-193 / :204 */
    ch_g2l(c, x, c->lin);
    if (nD > 0)
    {
#pragma omp parallel for schedule(static) num_threads(g_threads)
        for (i = 0; i < nL; ++i) inout[i] = c->lin[i] + inout[i];
    }
    else
        memcpy(inout, c->lin, sizeof(double) * (size_t)nL);
    if (phys_out && bwd) bwd(cb, inout, phys_out);
    if (final_eps) *final_eps = eps;
    return capped ? -its : its;
}
