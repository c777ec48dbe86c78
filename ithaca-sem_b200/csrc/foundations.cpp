// foundations.cpp -- host-side source of the 1-D tables the operators consume when they are used
// stand-alone (bench, tests, the C++ Collections mirror): quadrature points / weights /
// derivative matrices and the modified C0 bases.  In a real ITHACA-SEM build these arrays come
// from the reference's own BasisManager / PointsManager (LibUtilities/Foundations/Basis.cpp:
// 217-, GaussPoints.cpp:69-236, Polylib.cpp) and are handed to nekmf_op_create unchanged.
//
// Written from the mathematics, not from Polylib:
//   * Jacobi polynomials by the three-term recurrence;
//   * interior Gauss-Lobatto / Gauss-Radau nodes as zeros of Jacobi polynomials found by Newton
//     iteration with deflation of the already-found roots;
//   * weights from the closed forms for Lobatto-Legendre and Radau-Jacobi rules;
//   * derivative matrices from barycentric weights with the negative-row-sum diagonal.
#include "../../include/nekmf_b200.h"
#include <cmath>
#include <vector>

namespace
{

// P_n^{(a,b)}(x)
double jacobi(int n, double a, double b, double x)
{
    if (n == 0) return 1.0;
    double p0 = 1.0, p1 = 0.5 * (a - b + (a + b + 2.0) * x);
    for (int k = 2; k <= n; ++k)
    {
        const double c  = 2.0 * k + a + b;
        const double a1 = 2.0 * k * (k + a + b) * (c - 2.0);
        const double a2 = (c - 1.0) * (a * a - b * b);
        const double a3 = (c - 2.0) * (c - 1.0) * c;
        const double a4 = 2.0 * (k + a - 1.0) * (k + b - 1.0) * c;
        const double p2 = ((a2 + a3 * x) * p1 - a4 * p0) / a1;
        p0 = p1;
        p1 = p2;
    }
    return p1;
}
// d/dx P_n^{(a,b)}(x) = (n+a+b+1)/2 P_{n-1}^{(a+1,b+1)}(x)
double jacobi_deriv(int n, double a, double b, double x)
{
    return n == 0 ? 0.0 : 0.5 * (n + a + b + 1.0) * jacobi(n - 1, a + 1.0, b + 1.0, x);
}

// the n zeros of P_n^{(a,b)}, ascending
void jacobi_zeros(int n, double a, double b, double *z)
{
    const double pi = 3.14159265358979323846;
    for (int k = 0; k < n; ++k)
    {
        double r = -std::cos((2.0 * k + 1.0) * pi / (2.0 * n));
        if (k > 0) r = 0.5 * (r + z[k - 1]);
        for (int it = 0; it < 50; ++it)
        {
            const double p = jacobi(n, a, b, r), dp = jacobi_deriv(n, a, b, r);
            double defl = 0.0;
            for (int i = 0; i < k; ++i) defl += 1.0 / (r - z[i]);
            const double dr = -p / (dp - defl * p);
            r += dr;
            if (std::fabs(dr) < 1e-16 * (1.0 + std::fabs(r))) break;
        }
        z[k] = r;
    }
}

double gamma_ratio_int(int num, int den) // Gamma(num)/Gamma(den) for positive integers
{
    double g = 1.0;
    if (num > den)
        for (int t = den; t < num; ++t) g *= t;
    else
        for (int t = num; t < den; ++t) g /= t;
    return g;
}

// D[k*np+i] = l_k'(z_i) from barycentric weights
void bary_deriv(int np, const double *z, double *D)
{
    std::vector<double> c(np, 1.0);
    for (int i = 0; i < np; ++i)
        for (int j = 0; j < np; ++j)
            if (i != j) c[i] *= (z[i] - z[j]);
    for (int i = 0; i < np; ++i)
    {
        double diag = 0.0;
        for (int k = 0; k < np; ++k)
        {
            if (k == i) continue;
            const double v = (c[i] / c[k]) / (z[i] - z[k]);
            D[k * np + i]  = v;
            diag -= v;
        }
        D[i * np + i] = diag;
    }
}

void modified_a_rows(int nm, int np, const double *z, double *b)
{
    for (int i = 0; i < np; ++i)
    {
        b[i]      = 0.5 * (1.0 - z[i]);
        b[np + i] = 0.5 * (1.0 + z[i]);
    }
    for (int p = 2; p < nm; ++p)
        for (int i = 0; i < np; ++i) b[p * np + i] = b[i] * b[np + i] * jacobi(p - 2, 1.0, 1.0, z[i]);
}

// Modified_B: rows (p,q), q fastest, q < nm-p.
//   p = 0: Modified_A(q);  p = 1: row q=0 is (1-z)/2, rows q>=1 are Modified_A(q+1);
//   p >= 2: ((1-z)/2)^p for q = 0, ((1-z)/2)^p (1+z)/2 P_{q-1}^{(2p-1,1)} for q >= 1
int modified_b_rows(int nm, int np, const double *z, double *b)
{
    int row = 0;
    std::vector<double> A(nm * np);
    modified_a_rows(nm, np, z, A.data());
    for (int q = 0; q < nm; ++q, ++row)
        for (int i = 0; i < np; ++i) b[row * np + i] = A[q * np + i];
    if (nm > 1)
    {
        for (int i = 0; i < np; ++i) b[row * np + i] = A[i];
        ++row;
        for (int q = 2; q < nm; ++q, ++row)
            for (int i = 0; i < np; ++i) b[row * np + i] = A[q * np + i];
    }
    for (int p = 2; p < nm; ++p)
    {
        for (int q = 0; q < nm - p; ++q, ++row)
            for (int i = 0; i < np; ++i)
            {
                const double om = 0.5 * (1.0 - z[i]), op = 0.5 * (1.0 + z[i]);
                double v = std::pow(om, p);
                if (q > 0) v *= op * jacobi(q - 1, 2.0 * p - 1.0, 1.0, z[i]);
                b[row * np + i] = v;
            }
    }
    return row;
}

} // namespace

extern "C" {

// z[np], w[np] (raw weights), D[np*np] with D[k*np+i] = dh_k/dz(z_i).  D may be NULL.
int nekmf_points(int pointstype, int np, double *z, double *w, double *D)
{
    if (np < 1 || !z || !w) return NEKMF_ERR_ARG;
    if (pointstype == NEKMF_GLL)
    {
        if (np == 1)
        {
            z[0] = 0.0;
            w[0] = 2.0;
        }
        else
        {
            z[0]      = -1.0;
            z[np - 1] = 1.0;
            jacobi_zeros(np - 2, 1.0, 1.0, z + 1);
            const int N = np - 1;
            for (int i = 0; i < np; ++i)
            {
                const double L = jacobi(N, 0.0, 0.0, z[i]);
                w[i]           = 2.0 / (N * (N + 1.0) * L * L);
            }
        }
    }
    else if (pointstype == NEKMF_GRJM_A1B0 || pointstype == NEKMF_GRJM_A2B0)
    {
        const int a = pointstype == NEKMF_GRJM_A1B0 ? 1 : 2;
        if (np == 1)
        {
            z[0] = 0.0;
            w[0] = 2.0;
        }
        else
        {
            z[0] = -1.0;
            jacobi_zeros(np - 1, a, 1.0, z + 1);
            // Gauss-Radau-Jacobi (alpha=a, beta=0), node at -1:
            //   w_i = 2^a Gamma(np+a) / (Gamma(np) np Gamma(np+a+1) / Gamma(np+1)) (1-z_i) / P_{np-1}^{(a,0)}(z_i)^2
            const double fac = std::pow(2.0, a) * gamma_ratio_int(np + a, np) * gamma_ratio_int(np, np + a + 1) / np;
            for (int i = 0; i < np; ++i)
            {
                const double P = jacobi(np - 1, a, 0.0, z[i]);
                w[i]           = fac * (1.0 - z[i]) / (P * P);
            }
            // w[0] carries the (beta+1) = 1 factor
        }
    }
    else
        return NEKMF_ERR_ARG;
    if (D)
    {
        if (np == 1)
            D[0] = 0.0;
        else
            bary_deriv(np, z, D);
    }
    return NEKMF_OK;
}

int nekmf_basis_rows(int basistype, int nm)
{
    switch (basistype)
    {
        case NEKMF_MODIFIED_A: return nm;
        case NEKMF_MODIFIED_B: return nm * (nm + 1) / 2;
        case NEKMF_MODIFIED_C: return nm * (nm + 1) * (nm + 2) / 6;
        case NEKMF_MODIFIEDPYR_C: return nm * (nm + 1) * (2 * nm + 1) / 6;
    }
    return -1;
}

// bdata/dbdata: rows x np, b[m*np+i]; dbdata = derivative of each row evaluated through D
int nekmf_basis(int basistype, int nm, int np, const double *z, const double *D, double *bdata, double *dbdata)
{
    if (nm < 1 || np < 1 || !z || !D || !bdata || !dbdata) return NEKMF_ERR_ARG;
    const int rows = nekmf_basis_rows(basistype, nm);
    if (rows < 0) return NEKMF_ERR_ARG;
    if (basistype == NEKMF_MODIFIED_A)
        modified_a_rows(nm, np, z, bdata);
    else if (basistype == NEKMF_MODIFIED_B)
        modified_b_rows(nm, np, z, bdata);
    else if (basistype == NEKMF_MODIFIEDPYR_C)
    {
        // ModifiedPyr_C(p,q,r), r < nm - max(p,q) (Foundations/Basis.cpp:569-669):
        //   p < 2 or q < 2 (vertices, edges, triangular faces): Modified_B(max(p,q), r)
        //   p,q >= 2: ((1-z)/2)^(p+q-2) for r = 0 (base face), times (1+z)/2 P_{r-1}^{(2p+2q-3,1)} for r >= 1
        std::vector<double> B((size_t)nm * (nm + 1) / 2 * np);
        modified_b_rows(nm, np, z, B.data());
        std::vector<int> boff(nm + 1, 0); // first Modified_B row of block m
        for (int m = 0; m < nm; ++m) boff[m + 1] = boff[m] + (nm - m);
        int row = 0;
        for (int p = 0; p < nm; ++p)
            for (int q = 0; q < nm; ++q)
            {
                const int m = p > q ? p : q;
                for (int r = 0; r < nm - m; ++r, ++row)
                    for (int i = 0; i < np; ++i)
                    {
                        double v;
                        if (p < 2 || q < 2)
                            v = B[(size_t)(boff[m] + r) * np + i];
                        else
                        {
                            const double om = 0.5 * (1.0 - z[i]), op = 0.5 * (1.0 + z[i]);
                            v = std::pow(om, p + q - 2);
                            if (r > 0) v *= op * jacobi(r - 1, 2.0 * p + 2.0 * q - 3.0, 1.0, z[i]);
                        }
                        bdata[(size_t)row * np + i] = v;
                    }
            }
    }
    else
    {
        // Modified_C(p,q,r) = Modified_B(p+q, r): for every p the tail of the B table starting at block p
        std::vector<double> B((size_t)nm * (nm + 1) / 2 * np);
        modified_b_rows(nm, np, z, B.data());
        size_t off = 0;
        int blk    = 0; // first B row of block p
        for (int p = 0; p < nm; ++p)
        {
            const int nrows = (nm - p) * (nm - p + 1) / 2;
            for (size_t t = 0; t < (size_t)nrows * np; ++t) bdata[off + t] = B[(size_t)blk * np + t];
            off += (size_t)nrows * np;
            blk += nm - p;
        }
    }
    for (int m = 0; m < rows; ++m)
        for (int i = 0; i < np; ++i)
        {
            double s = 0.0;
            for (int j = 0; j < np; ++j) s += D[j * np + i] * bdata[m * np + j];
            dbdata[m * np + i] = s;
        }
    return NEKMF_OK;
}

} // extern "C"
