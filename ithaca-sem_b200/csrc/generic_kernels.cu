// generic_kernels.cu -- runtime-sized sum-factorisation kernels for every shape the path
// supports (Quad, Tri, Hex, Prism, Pyr, Tet) and any (nm, nq).  These are the complete-coverage
// kernels: collapsed-coordinate shapes, over-integrated quadrature and any order the
// compile-time specialised hex/quad kernels do not instantiate run here.
//
// One persistent CTA handles one element at a time with everything in shared memory: the 1-D
// tables are loaded once per CTA, the element's data once per element; each sum-factorisation
// stage is distributed over the CTA's threads by OUTPUT entry (one thread = one output value,
// inner runtime loop over the contracted index), stages separated by __syncthreads().
//
// Reference semantics:
//   BwdTrans        MatrixFreeOps/BwdTransKernels.hpp:35-76 (Quad) 78-126 (Tri) 302-372 (Hex)
//                   224-299 (Prism) 374-484 (Tet), incl. the CORRECT terms of eModified_A
//   IProductWRTBase MatrixFreeOps/IProductKernels.hpp:76-133, 135-234, 236-314, 316-450, 600-761
//   PhysDeriv       MatrixFreeOps/PhysDerivKernels.hpp:39-90, 153-217, 219-372, 376-461, 555-696
//   Helmholtz       MatrixFreeOps/Helmholtz.h:138-275, 506-635, 764-993, 1291-1458, 2266-2448
//   IProductWRTDerivBase MatrixFreeOps/IProductWRTDerivBase.h:542- (Quad) 891- (Tri) 1232- (Hex) 1630- (Prism)
//                   2056- (Pyr) 2484- (Tet)
//   Pyramid         BwdTransKernels.hpp:131-222, IProductKernels.hpp:455-598, PhysDerivKernels.hpp:464-553,
//                   Helmholtz.h:1771-1955 (modes (p,q,r), r fastest, nm - max(p,q) per (p,q); the rows of the
//                   eModifiedPyr_C table are indexed by the mode itself)
#include "op_internal.h"

namespace nekmf
{

struct GenArgs
{
    const double *in0, *in1, *in2;
    double *out0, *out1, *out2;
    const double *jac, *df;
    const double *tab; // packed tables, see nekmf_op_s::tab_off
    int tab_len;
    int off[3][5];
    int nm, nq0, nq1, nq2;
    int nmTot, nqTot, nElmt, deformed, optype;
    size_t dfStride; // df row stride (whole collection); jac/df already point at this launch's first element
    int N; // doubles per work buffer
    int groups; // thread groups per CTA (1, 2, 4 or 8); each has 7 work buffers of N doubles
    double lambda;
};

struct GenCtx
{
    int nm, nq0, nq1, nq2, nmTot, nqTot;
    const double *b[3], *db[3], *D[3], *Z[3], *w[3];
    int *offP;  // offP[p] = sum_{p'<p} (nm-p')                       (Tri / Prism / Tet dir-1 rows)
    int *offPQ; // Tet: mode offset of (p,q), running index cpq       (size nm(nm+1)/2 + 1)
                // Pyr: mode offset of (p,q) at [p*nm+q]              (size nm*nm + 1)
    int *cpqP;  // Tet: cpq offset of p
    int tid, tg, bar; // thread within its group, group size, named barrier of the group
};

// The CTA is split into thread groups of c.tg threads; every group works on its own element with its own work
// buffers and synchronises on its own named barrier, so small elements (a quad has 50-150 values per stage) keep
// all 256 threads busy instead of one group's worth.
#define GEN_FOR(idx, n) for (int idx = c.tid; idx < (n); idx += c.tg)
__device__ __forceinline__ void gen_sync(const GenCtx &c)
{
    asm volatile("bar.sync %0, %1;" ::"r"(c.bar), "r"(c.tg) : "memory");
}

// ------------------------------------------------------------------------------ BwdTrans
template <int SHAPE> __device__ void gen_bwd(const GenCtx &c, const double *in, double *out, double *w1, double *w2)
{
    const int nm = c.nm, nq0 = c.nq0, nq1 = c.nq1, nq2 = c.nq2;
    const double *b0 = c.b[0], *b1 = c.b[1], *b2 = c.b[2];
    if (SHAPE == NEKMF_QUAD)
    {
        GEN_FOR(t, nq0 * nm)
        {
            const int i = t / nm, q = t - i * nm;
            double s = 0.0;
            for (int p = 0; p < nm; ++p) s = fma(in[q * nm + p], b0[p * nq0 + i], s);
            w1[t] = s;
        }
        gen_sync(c);
        GEN_FOR(t, nq0 * nq1)
        {
            const int j = t / nq0, i = t - j * nq0;
            double s = 0.0;
            for (int q = 0; q < nm; ++q) s = fma(w1[i * nm + q], b1[q * nq1 + j], s);
            out[t] = s;
        }
        gen_sync(c);
    }
    else if (SHAPE == NEKMF_TRI)
    {
        GEN_FOR(t, nq1 * nm)
        {
            const int e1 = t / nm, p = t - e1 * nm;
            const int m0 = c.offP[p];
            double s = 0.0;
            for (int q = 0; q < nm - p; ++q) s = fma(b1[(m0 + q) * nq1 + e1], in[m0 + q], s);
            w1[t] = s;
        }
        gen_sync(c);
        GEN_FOR(t, nq0 * nq1)
        {
            const int e1 = t / nq0, e0 = t - e1 * nq0;
            double s = 0.0;
            for (int p = 0; p < nm; ++p) s = fma(w1[e1 * nm + p], b0[p * nq0 + e0], s);
            s += (in[1] * b0[nq0 + e0]) * b1[nq1 + e1]; // CORRECT: singular vertex
            out[t] = s;
        }
        gen_sync(c);
    }
    else if (SHAPE == NEKMF_HEX)
    {
        GEN_FOR(t, nq0 * nm * nm)
        {
            const int i = t / (nm * nm), rq = t - i * nm * nm;
            double s = 0.0;
            for (int p = 0; p < nm; ++p) s = fma(in[rq * nm + p], b0[p * nq0 + i], s);
            w1[t] = s; // [i][r][q]
        }
        gen_sync(c);
        GEN_FOR(t, nq1 * nq0 * nm)
        {
            const int j = t / (nq0 * nm), ir = t - j * nq0 * nm;
            double s = 0.0;
            for (int q = 0; q < nm; ++q) s = fma(w1[ir * nm + q], b1[q * nq1 + j], s);
            w2[t] = s; // [j][i][r]
        }
        gen_sync(c);
        GEN_FOR(t, nq0 * nq1 * nq2)
        {
            const int k = t / (nq0 * nq1), ji = t - k * nq0 * nq1;
            double s = 0.0;
            for (int r = 0; r < nm; ++r) s = fma(w2[ji * nm + r], b2[r * nq2 + k], s);
            out[t] = s;
        }
        gen_sync(c);
    }
    else if (SHAPE == NEKMF_PRISM)
    {
        // fpq[k][p][q]
        GEN_FOR(t, nq2 * nm * nm)
        {
            const int k = t / (nm * nm), pq = t - k * nm * nm;
            const int p = pq / nm, q = pq - p * nm;
            const int mpr = c.offP[p], m0 = mpr * nm + q * (nm - p);
            double s = 0.0;
            for (int r = 0; r < nm - p; ++r) s = fma(in[m0 + r], b2[(mpr + r) * nq2 + k], s);
            w1[t] = s;
        }
        gen_sync(c);
        // fp[k][j][p]
        GEN_FOR(t, nq2 * nq1 * nm)
        {
            const int k = t / (nq1 * nm), jp = t - k * nq1 * nm;
            const int j = jp / nm, p = jp - j * nm;
            double s = 0.0;
            for (int q = 0; q < nm; ++q) s = fma(w1[(k * nm + p) * nm + q], b1[q * nq1 + j], s);
            w2[t] = s;
        }
        gen_sync(c);
        GEN_FOR(t, nq0 * nq1 * nq2)
        {
            const int k = t / (nq0 * nq1), ji = t - k * nq0 * nq1;
            const int j = ji / nq0, i = ji - j * nq0;
            double s = 0.0;
            for (int p = 0; p < nm; ++p) s = fma(w2[(k * nq1 + j) * nm + p], b0[p * nq0 + i], s);
            // CORRECT: singular edge, modes (p=0,q,r=1)
            const double ba2 = b2[nq2 + k], ba0 = b0[nq0 + i];
            for (int q = 0; q < nm; ++q) s = fma(ba2 * b1[q * nq1 + j], ba0 * in[q * nm + 1], s);
            out[t] = s;
        }
        gen_sync(c);
    }
    else if (SHAPE == NEKMF_PYR)
    {
        // fpq[k][p][q] = sum_r in[m] b2[m][k]
        GEN_FOR(t, nq2 * nm * nm)
        {
            const int k = t / (nm * nm), pq = t - k * nm * nm;
            const int m0 = c.offPQ[pq], len = c.offPQ[pq + 1] - m0;
            double s = 0.0;
            for (int r = 0; r < len; ++r) s = fma(in[m0 + r], b2[(m0 + r) * nq2 + k], s);
            w1[t] = s;
        }
        gen_sync(c);
        // fp[k][j][p]
        GEN_FOR(t, nq2 * nq1 * nm)
        {
            const int k = t / (nq1 * nm), jp = t - k * nq1 * nm;
            const int j = jp / nm, p = jp - j * nm;
            double s = 0.0;
            for (int q = 0; q < nm; ++q) s = fma(w1[(k * nm + p) * nm + q], b1[q * nq1 + j], s);
            w2[t] = s;
        }
        gen_sync(c);
        GEN_FOR(t, nq0 * nq1 * nq2)
        {
            const int k = t / (nq0 * nq1), ji = t - k * nq0 * nq1;
            const int j = ji / nq0, i = ji - j * nq0;
            double s = 0.0;
            for (int p = 0; p < nm; ++p) s = fma(w2[(k * nq1 + j) * nm + p], b0[p * nq0 + i], s);
            // CORRECT: top vertex (mode 1)
            const double t1 = b0[i] * b1[nq1 + j] + b0[nq0 + i] * b1[j] + b0[nq0 + i] * b1[nq1 + j];
            s = fma(t1 * b2[nq2 + k], in[1], s);
            out[t] = s;
        }
        gen_sync(c);
    }
    else // TET
    {
        const int npq = nm * (nm + 1) / 2;
        // fpq[k][cpq]
        GEN_FOR(t, nq2 * npq)
        {
            const int k = t / npq, cpq = t - k * npq;
            const int m0 = c.offPQ[cpq], len = c.offPQ[cpq + 1] - m0;
            double s = 0.0;
            for (int r = 0; r < len; ++r) s = fma(in[m0 + r], b2[(m0 + r) * nq2 + k], s);
            w1[t] = s;
        }
        gen_sync(c);
        // fp[k][j][p]
        GEN_FOR(t, nq2 * nq1 * nm)
        {
            const int k = t / (nq1 * nm), jp = t - k * nq1 * nm;
            const int j = jp / nm, p = jp - j * nm;
            const int c0 = c.cpqP[p], m1 = c.offP[p];
            double s = 0.0;
            for (int q = 0; q < nm - p; ++q) s = fma(w1[k * npq + c0 + q], b1[(m1 + q) * nq1 + j], s);
            w2[t] = s;
        }
        gen_sync(c);
        GEN_FOR(t, nq0 * nq1 * nq2)
        {
            const int k = t / (nq0 * nq1), ji = t - k * nq0 * nq1;
            const int j = ji / nq0, i = ji - j * nq0;
            double s = 0.0;
            for (int p = 0; p < nm; ++p) s = fma(b0[p * nq0 + i], w2[(k * nq1 + j) * nm + p], s);
            // CORRECT: top vertex (mode 1)
            double t1 = b0[i] * b1[nq1 + j] + b0[nq0 + i] * b1[j] + b0[nq0 + i] * b1[nq1 + j];
            s = fma(t1 * b2[nq2 + k], in[1], s);
            // bottom vertex (mode nm) and singular edge (modes nm+r)
            const double e = b0[nq0 + i] * b1[nq1 + j];
            s = fma(in[nm], e * b2[k], s);
            for (int r = 1; r < nm - 1; ++r) s = fma(in[nm + r], e * b2[(r + 1) * nq2 + k], s);
            out[t] = s;
        }
        gen_sync(c);
    }
}

// --------------------------------------------------------------------------- IProduct
// out (=|+=) scale * sum_pts f(pt) * B0 B1 B2, f already multiplied by jac and the weights.
// B0,B1,B2 are bdata or dbdata per direction.  w1,w2: work buffers.
template <int SHAPE>
__device__ void gen_ip(const GenCtx &c, const double *f, const double *B0, const double *B1, const double *B2,
                       double *out, double scale, bool append, double *w1, double *w2)
{
    const int nm = c.nm, nq0 = c.nq0, nq1 = c.nq1, nq2 = c.nq2;
    if (SHAPE == NEKMF_QUAD || SHAPE == NEKMF_TRI)
    {
        // s1[j][p] = sum_i f[j][i] B0[p][i]
        GEN_FOR(t, nq1 * nm)
        {
            const int j = t / nm, p = t - j * nm;
            double s = 0.0;
            for (int i = 0; i < nq0; ++i) s = fma(f[j * nq0 + i], B0[p * nq0 + i], s);
            w1[t] = s;
        }
        gen_sync(c);
        if (SHAPE == NEKMF_QUAD)
        {
            GEN_FOR(t, nm * nm)
            {
                const int q = t / nm, p = t - q * nm;
                double s = 0.0;
                for (int j = 0; j < nq1; ++j) s = fma(w1[j * nm + p], B1[q * nq1 + j], s);
                out[t] = append ? out[t] + s * scale : s * scale;
            }
        }
        else
        {
            GEN_FOR(t, c.nmTot)
            {
                // mode -> (p,q)
                int p = 0;
                while (p + 1 < nm && c.offP[p + 1] <= t) ++p;
                double s = 0.0;
                for (int j = 0; j < nq1; ++j) s = fma(w1[j * nm + p], B1[t * nq1 + j], s);
                if (t == 1)
                {
                    // CORRECT: singular vertex
                    for (int j = 0; j < nq1; ++j) s = fma(w1[j * nm + 1], B1[nq1 + j], s);
                }
                out[t] = append ? out[t] + s * scale : s * scale;
            }
        }
        gen_sync(c);
        return;
    }
    // 3-D: s1[k][j][p]
    GEN_FOR(t, nq2 * nq1 * nm)
    {
        const int kj = t / nm, p = t - kj * nm;
        double s = 0.0;
        for (int i = 0; i < nq0; ++i) s = fma(f[kj * nq0 + i], B0[p * nq0 + i], s);
        w1[t] = s;
    }
    gen_sync(c);
    if (SHAPE == NEKMF_HEX || SHAPE == NEKMF_PRISM || SHAPE == NEKMF_PYR)
    {
        // s2[k][q][p]
        GEN_FOR(t, nq2 * nm * nm)
        {
            const int k = t / (nm * nm), qp = t - k * nm * nm;
            const int q = qp / nm, p = qp - q * nm;
            double s = 0.0;
            for (int j = 0; j < nq1; ++j) s = fma(w1[(k * nq1 + j) * nm + p], B1[q * nq1 + j], s);
            w2[t] = s;
        }
        gen_sync(c);
        if (SHAPE == NEKMF_HEX)
        {
            GEN_FOR(t, nm * nm * nm)
            {
                const int r = t / (nm * nm), qp = t - r * nm * nm;
                double s = 0.0;
                for (int k = 0; k < nq2; ++k) s = fma(w2[k * nm * nm + qp], B2[r * nq2 + k], s);
                out[t] = append ? out[t] + s * scale : s * scale;
            }
        }
        else if (SHAPE == NEKMF_PYR)
        {
            GEN_FOR(t, c.nmTot)
            {
                // mode -> (p,q): last pq with offPQ[pq] <= t
                int lo = 0, hi = nm * nm - 1;
                while (lo < hi)
                {
                    const int mid = (lo + hi + 1) >> 1;
                    if (c.offPQ[mid] <= t) lo = mid;
                    else hi = mid - 1;
                }
                const int p = lo / nm, q = lo - p * nm;
                double s = 0.0;
                for (int k = 0; k < nq2; ++k) s = fma(w2[(k * nm + q) * nm + p], B2[t * nq2 + k], s);
                if (t == 1)
                {
                    // CORRECT: top vertex collects the (p,q) = (0,1), (1,0), (1,1) lines
                    for (int k = 0; k < nq2; ++k)
                        s = fma(w2[(k * nm + 1) * nm] + w2[k * nm * nm + 1] + w2[(k * nm + 1) * nm + 1], B2[nq2 + k], s);
                }
                out[t] = append ? out[t] + s * scale : s * scale;
            }
        }
        else
        {
            GEN_FOR(t, c.nmTot)
            {
                // mode -> (p,q,r): block p has nm*(nm-p) modes, q-major
                int p = 0;
                while (p + 1 < nm && c.offP[p + 1] * nm <= t) ++p;
                const int rem = t - c.offP[p] * nm;
                const int q = rem / (nm - p), r = rem - q * (nm - p);
                const int row = c.offP[p] + r;
                double s = 0.0;
                for (int k = 0; k < nq2; ++k) s = fma(w2[(k * nm + q) * nm + p], B2[row * nq2 + k], s);
                if (p == 0 && r == 1)
                {
                    // CORRECT: singular edge; corr[q] = sum_k B2[1][k] s2[k][q][p=1]
                    for (int k = 0; k < nq2; ++k) s = fma(w2[(k * nm + q) * nm + 1], B2[nq2 + k], s);
                }
                out[t] = append ? out[t] + s * scale : s * scale;
            }
        }
        gen_sync(c);
        return;
    }
    // TET: s2[k][cpq]
    const int npq = nm * (nm + 1) / 2;
    GEN_FOR(t, nq2 * npq)
    {
        const int k = t / npq, cpq = t - k * npq;
        int p = 0;
        while (p + 1 < nm && c.cpqP[p + 1] <= cpq) ++p;
        const int q = cpq - c.cpqP[p];
        const int row = c.offP[p] + q;
        double s = 0.0;
        for (int j = 0; j < nq1; ++j) s = fma(w1[(k * nq1 + j) * nm + p], B1[row * nq1 + j], s);
        w2[t] = s;
    }
    // correction line sums (use s1 with p = 0, 1):  cc[k] = sum_j a1 B1[1][j],  c0[k] = sum_j (a0 B1[1][j] + a1 B1[0][j])
    double *cc = w2 + nq2 * npq, *c0 = cc + nq2;
    GEN_FOR(k, nq2)
    {
        double s = 0.0, s0 = 0.0;
        for (int j = 0; j < nq1; ++j)
        {
            const double a0 = w1[(k * nq1 + j) * nm], a1 = w1[(k * nq1 + j) * nm + 1];
            s  = fma(a1, B1[nq1 + j], s);
            s0 = fma(a0, B1[nq1 + j], s0);
            s0 = fma(a1, B1[j], s0);
        }
        cc[k] = s;
        c0[k] = s0;
    }
    gen_sync(c);
    GEN_FOR(t, c.nmTot)
    {
        // mode -> cpq
        int lo = 0;
        while (c.offPQ[lo + 1] <= t) ++lo;
        double s = 0.0;
        for (int k = 0; k < nq2; ++k) s = fma(w2[k * npq + lo], B2[t * nq2 + k], s);
        if (t == 1)
        {
            for (int k = 0; k < nq2; ++k) s = fma(B2[nq2 + k], c0[k] + cc[k], s);
        }
        else if (t == nm)
        {
            for (int k = 0; k < nq2; ++k) s = fma(cc[k], B2[k], s);
        }
        else if (t > nm && t < 2 * nm - 1)
        {
            const int r = t - nm;
            for (int k = 0; k < nq2; ++k) s = fma(cc[k], B2[(r + 1) * nq2 + k], s);
        }
        out[t] = append ? out[t] + s * scale : s * scale;
    }
    gen_sync(c);
}

// f[pt] = in[pt] * jac * w0 w1 w2
__device__ void gen_weight(const GenCtx &c, const double *in, double *f, const double *jac, bool deformed, int dim)
{
    const int nq0 = c.nq0, nq1 = c.nq1;
    GEN_FOR(t, c.nqTot)
    {
        const int i = t % nq0, j = (t / nq0) % nq1, k = t / (nq0 * nq1);
        double w = c.w[0][i] * c.w[1][j];
        if (dim == 3) w *= c.w[2][k];
        f[t] = in[t] * ((deformed ? __ldg(jac + t) : __ldg(jac)) * w);
    }
    gen_sync(c);
}

// ------------------------------------------------------------------ tensor derivatives
__device__ void gen_dtensor(const GenCtx &c, int dim, const double *u, double *d0, double *d1, double *d2)
{
    const int nq0 = c.nq0, nq1 = c.nq1, nq2 = dim == 3 ? c.nq2 : 1;
    const double *D0 = c.D[0], *D1 = c.D[1], *D2 = c.D[2];
    GEN_FOR(t, nq0 * nq1 * nq2)
    {
        const int i = t % nq0, j = (t / nq0) % nq1, k = t / (nq0 * nq1);
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        const int row = t - i, col = t - j * nq0, pil = t - k * nq0 * nq1;
        for (int m = 0; m < nq0; ++m) s0 = fma(D0[m * nq0 + i], u[row + m], s0);
        for (int m = 0; m < nq1; ++m) s1 = fma(D1[m * nq1 + j], u[col + m * nq0], s1);
        d0[t] = s0;
        d1[t] = s1;
        if (dim == 3)
        {
            for (int m = 0; m < nq2; ++m) s2 = fma(D2[m * nq2 + k], u[pil + m * nq0 * nq1], s2);
            d2[t] = s2;
        }
    }
    gen_sync(c);
}

#define GDF(n) (deformed ? __ldg(df + (size_t)(n) * dfs + t) : __ldg(df + (size_t)(n) * dfs))

// ---------------------------------------------------------------------------- PhysDeriv
// d0,d1,d2 hold the tensor derivatives on entry and the Cartesian derivatives on exit.
template <int SHAPE>
__device__ void gen_pd_apply(const GenCtx &c, double *d0, double *d1, double *d2, const double *df, size_t dfs,
                             bool deformed)
{
    const int nq0 = c.nq0, nq1 = c.nq1;
    GEN_FOR(t, c.nqTot)
    {
        const int i = t % nq0, j = (t / nq0) % nq1, k = t / (nq0 * nq1);
        if (SHAPE == NEKMF_QUAD || SHAPE == NEKMF_TRI)
        {
            double a = d0[t], b = d1[t];
            if (SHAPE == NEKMF_TRI)
            {
                a = (2.0 / (1.0 - c.Z[1][j])) * a;
                b = fma(a, 0.5 * (1.0 + c.Z[0][i]), b);
            }
            d0[t] = a * GDF(0) + b * GDF(1);
            d1[t] = a * GDF(2) + b * GDF(3);
        }
        else
        {
            double a = d0[t], b = d1[t], g = d2[t];
            if (SHAPE == NEKMF_PRISM)
            {
                a = a * (2.0 / (1.0 - c.Z[2][k]));
                g = fma(0.5 * (1.0 + c.Z[0][i]), a, g);
            }
            else if (SHAPE == NEKMF_PYR)
            {
                // PhysDerivKernels.hpp:505-527
                const double x2 = 2.0 / (1.0 - c.Z[2][k]);
                a = a * x2;
                b = b * x2;
                g = fma(0.5 * (1.0 + c.Z[0][i]), a, g);
                g = fma(0.5 * (1.0 + c.Z[1][j]), b, g);
            }
            else if (SHAPE == NEKMF_TET)
            {
                const double x2 = 2.0 / (1.0 - c.Z[2][k]), x1 = 2.0 / (1.0 - c.Z[1][j]);
                const double x0 = 0.5 * (1.0 + c.Z[0][i]), y1 = 0.5 * (1.0 + c.Z[1][j]);
                const double A = (x1 * x2) * a; // d/d eta0 contribution
                const double f0 = x0 * A;
                const double f1 = x2 * b;
                a = A;
                b = f0 + f1;
                g = f0 + f1 * y1 + g;
            }
            d0[t] = a * GDF(0) + b * GDF(1) + g * GDF(2);
            d1[t] = a * GDF(3) + b * GDF(4) + g * GDF(5);
            d2[t] = a * GDF(6) + b * GDF(7) + g * GDF(8);
        }
    }
    gen_sync(c);
}

// ---------------------------------------------------------------------------- Helmholtz metric
// g0,g1,g2 <- G * (g0,g1,g2) with the Laplacian metric of Helmholtz.h (collapsed-coordinate
// factors folded in); does not include jac*weights (the IProducts add them).
template <int SHAPE>
__device__ void gen_metric(const GenCtx &c, double *g0, double *g1, double *g2, const double *df, size_t dfs,
                           bool deformed)
{
    const int nq0 = c.nq0, nq1 = c.nq1;
    GEN_FOR(t, c.nqTot)
    {
        const int i = t % nq0, j = (t / nq0) % nq1, k = t / (nq0 * nq1);
        if (SHAPE == NEKMF_QUAD || SHAPE == NEKMF_TRI)
        {
            const double df0 = GDF(0), df1 = GDF(1), df2 = GDF(2), df3 = GDF(3);
            double m00, m01, m11;
            if (SHAPE == NEKMF_QUAD)
            {
                m00 = df0 * df0 + df2 * df2;
                m01 = df0 * df1 + df2 * df3;
                m11 = df1 * df1 + df3 * df3;
            }
            else
            {
                const double h1 = 2.0 / (1.0 - c.Z[1][j]), h0 = 0.5 * (1.0 + c.Z[0][i]);
                const double a = h1 * (df0 + h0 * df1), b = h1 * (df2 + h0 * df3);
                m00 = a * a + b * b;
                m01 = a * df1 + b * df3;
                m11 = df1 * df1 + df3 * df3;
            }
            const double d0 = g0[t], d1 = g1[t];
            g0[t] = m00 * d0 + m01 * d1;
            g1[t] = m01 * d0 + m11 * d1;
        }
        else
        {
            const double df0 = GDF(0), df1 = GDF(1), df2 = GDF(2), df3 = GDF(3), df4 = GDF(4), df5 = GDF(5),
                         df6 = GDF(6), df7 = GDF(7), df8 = GDF(8);
            double m00, m01, m02, m11, m12, m22;
            if (SHAPE == NEKMF_HEX)
            {
                m00 = df0 * df0 + df3 * df3 + df6 * df6;
                m01 = df0 * df1 + df3 * df4 + df6 * df7;
                m02 = df0 * df2 + df3 * df5 + df6 * df8;
                m11 = df1 * df1 + df4 * df4 + df7 * df7;
                m12 = df1 * df2 + df4 * df5 + df7 * df8;
                m22 = df2 * df2 + df5 * df5 + df8 * df8;
            }
            else if (SHAPE == NEKMF_PRISM)
            {
                const double h1 = 2.0 / (1.0 - c.Z[2][k]), h0 = 0.5 * (1.0 + c.Z[0][i]);
                const double t1 = h1 * (h0 * df2 + df0), t2 = h1 * (h0 * df5 + df3), t3 = h1 * (h0 * df8 + df6);
                m00 = t1 * t1 + t2 * t2 + t3 * t3;
                m01 = df1 * t1 + df4 * t2 + df7 * t3;
                m02 = df2 * t1 + df5 * t2 + df8 * t3;
                m11 = df1 * df1 + df4 * df4 + df7 * df7;
                m22 = df2 * df2 + df5 * df5 + df8 * df8;
                m12 = df1 * df2 + df4 * df5 + df7 * df8;
            }
            else if (SHAPE == NEKMF_PYR)
            {
                // Helmholtz.h:1845-1905
                const double h0 = 0.5 * (1.0 + c.Z[0][i]), h1 = 0.5 * (1.0 + c.Z[1][j]), h2 = 2.0 / (1.0 - c.Z[2][k]);
                const double h1h2 = h1 * h2, h0h2 = h0 * h2;
                const double t0 = h2 * df0 + h0h2 * df2, t1 = h2 * df3 + h0h2 * df5, t2 = h2 * df6 + h0h2 * df8;
                const double t3 = h2 * df1 + h1h2 * df2, t4 = h2 * df4 + h1h2 * df5, t5 = h2 * df7 + h1h2 * df8;
                m00 = t0 * t0 + t1 * t1 + t2 * t2;
                m11 = t3 * t3 + t4 * t4 + t5 * t5;
                m22 = df2 * df2 + df5 * df5 + df8 * df8;
                m01 = t0 * t3 + t1 * t4 + t2 * t5;
                m02 = df2 * t0 + df5 * t1 + df8 * t2;
                m12 = df2 * t3 + df5 * t4 + df8 * t5;
            }
            else
            {
                const double h0 = 0.5 * (1.0 + c.Z[0][i]), h1 = 0.5 * (1.0 + c.Z[1][j]);
                const double h2 = 2.0 / (1.0 - c.Z[1][j]), h3 = 2.0 / (1.0 - c.Z[2][k]);
                const double h2h3 = h2 * h3, h1h3 = h1 * h3, h0h2h3 = h0 * h2h3;
                const double t1 = h0h2h3 * (df1 + df2) + df0 * h2h3;
                const double t2 = h0h2h3 * (df4 + df5) + df3 * h2h3;
                const double t3 = h0h2h3 * (df7 + df8) + df6 * h2h3;
                const double t4 = df1 * h3 + df2 * h1h3, t5 = df4 * h3 + df5 * h1h3, t6 = df7 * h3 + df8 * h1h3;
                m00 = t1 * t1 + t2 * t2 + t3 * t3;
                m02 = df2 * t1 + df5 * t2 + df8 * t3;
                m01 = t1 * t4 + t2 * t5 + t3 * t6;
                m11 = t4 * t4 + t5 * t5 + t6 * t6;
                m12 = df2 * t4 + df5 * t5 + df8 * t6;
                m22 = df2 * df2 + df5 * df5 + df8 * df8;
            }
            const double d0 = g0[t], d1 = g1[t], d2 = g2[t];
            g0[t] = m00 * d0 + m01 * d1 + m02 * d2;
            g1[t] = m01 * d0 + m11 * d1 + m12 * d2;
            g2[t] = m02 * d0 + m12 * d1 + m22 * d2;
        }
    }
    gen_sync(c);
}

// ---------------------------------------------------------------------------- the kernel
template <int SHAPE> __global__ void __launch_bounds__(256) gen_kernel(const __grid_constant__ GenArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sTab = reinterpret_cast<double *>(smem_raw);
    const int N  = a.N, G = a.groups, TG = 256 / G;
    const int grp = threadIdx.x / TG;
    double *buf  = sTab + a.tab_len + (size_t)grp * 7 * N;
    double *sIn = buf, *sOut = buf + N, *s1 = buf + 2 * N, *s2 = buf + 3 * N, *s3 = buf + 4 * N, *s4 = buf + 5 * N,
           *s5 = buf + 6 * N;
    int *sInt = reinterpret_cast<int *>(sTab + a.tab_len + (size_t)G * 7 * N);
    constexpr int dim = (SHAPE == NEKMF_QUAD || SHAPE == NEKMF_TRI) ? 2 : 3;

    GenCtx c;
    c.tid = threadIdx.x - grp * TG; c.tg = TG; c.bar = 1 + grp;
    c.nm = a.nm; c.nq0 = a.nq0; c.nq1 = a.nq1; c.nq2 = dim == 3 ? a.nq2 : 1;
    c.nmTot = a.nmTot; c.nqTot = a.nqTot;
    for (int d = 0; d < 3; ++d)
    {
        const int dd = d < dim ? d : 0;
        c.b[d] = sTab + a.off[dd][0]; c.db[d] = sTab + a.off[dd][1]; c.D[d] = sTab + a.off[dd][2];
        c.Z[d] = sTab + a.off[dd][3]; c.w[d] = sTab + a.off[dd][4];
    }
    const int nm = a.nm;
    c.offP  = sInt;
    c.cpqP  = sInt + nm + 1;
    c.offPQ = sInt + 2 * (nm + 1);
    for (int i = threadIdx.x; i < a.tab_len; i += blockDim.x) sTab[i] = __ldg(a.tab + i);
    if (threadIdx.x == 0)
    {
        int o = 0, cq = 0, m = 0;
        for (int p = 0; p <= nm; ++p)
        {
            c.offP[p] = o;
            c.cpqP[p] = cq;
            if (p < nm)
            {
                o += nm - p;
                if (SHAPE != NEKMF_PYR)
                    for (int q = 0; q < nm - p; ++q, ++cq)
                    {
                        c.offPQ[cq] = m;
                        m += nm - p - q;
                    }
            }
        }
        if (SHAPE == NEKMF_PYR)
        {
            // mode offset of every (p,q): nm - max(p,q) modes each, r fastest
            for (int pq = 0; pq < nm * nm; ++pq)
            {
                const int p = pq / nm, q = pq - p * nm;
                c.offPQ[pq] = m;
                m += nm - (p > q ? p : q);
            }
            cq = nm * nm;
        }
        c.offPQ[cq] = m;
    }
    __syncthreads();

    const bool deformed = a.deformed != 0;
    const int nmTot = a.nmTot, nqTot = a.nqTot;
    const size_t dfs = a.dfStride;
    const bool coeff_in  = a.optype == NEKMF_BWDTRANS || a.optype == NEKMF_HELMHOLTZ;
    const bool coeff_out = a.optype != NEKMF_BWDTRANS && a.optype != NEKMF_PHYSDERIV;
    const int nin = coeff_in ? nmTot : nqTot, nout = coeff_out ? nmTot : nqTot;

    for (int e = blockIdx.x * G + grp; e < a.nElmt; e += gridDim.x * G)
    {
        const size_t goff  = deformed ? (size_t)e * nqTot : (size_t)e;
        const double *jac = a.jac ? a.jac + goff : nullptr;
        const double *df  = a.df ? a.df + goff : nullptr;
        GEN_FOR(t, nin) sIn[t] = __ldg(a.in0 + (size_t)e * nin + t);
        if (a.optype == NEKMF_IPRODUCTWRTDERIVBASE)
        {
            GEN_FOR(t, nin)
            {
                s3[t] = __ldg(a.in1 + (size_t)e * nin + t);
                if (dim == 3) s4[t] = __ldg(a.in2 + (size_t)e * nin + t);
            }
        }
        gen_sync(c);
        switch (a.optype)
        {
            case NEKMF_BWDTRANS: gen_bwd<SHAPE>(c, sIn, sOut, s1, s2); break;
            case NEKMF_IPRODUCTWRTBASE:
                gen_weight(c, sIn, s3, jac, deformed, dim);
                gen_ip<SHAPE>(c, s3, c.b[0], c.b[1], c.b[2], sOut, 1.0, false, s1, s2);
                break;
            case NEKMF_PHYSDERIV:
                gen_dtensor(c, dim, sIn, sOut, s3, s4);
                gen_pd_apply<SHAPE>(c, sOut, s3, s4, df, dfs, deformed);
                break;
            case NEKMF_HELMHOLTZ:
                gen_bwd<SHAPE>(c, sIn, s5, s1, s2); // u -> s5
                gen_dtensor(c, dim, s5, sIn, s3, s4); // derivatives -> sIn, s3, s4 (sIn is dead)
                gen_weight(c, s5, s5, jac, deformed, dim);
                gen_ip<SHAPE>(c, s5, c.b[0], c.b[1], c.b[2], sOut, a.lambda, false, s1, s2);
                gen_metric<SHAPE>(c, sIn, s3, s4, df, dfs, deformed);
                gen_weight(c, sIn, sIn, jac, deformed, dim);
                gen_ip<SHAPE>(c, sIn, c.db[0], c.b[1], c.b[2], sOut, 1.0, true, s1, s2);
                gen_weight(c, s3, s3, jac, deformed, dim);
                gen_ip<SHAPE>(c, s3, c.b[0], c.db[1], c.b[2], sOut, 1.0, true, s1, s2);
                if (dim == 3)
                {
                    gen_weight(c, s4, s4, jac, deformed, dim);
                    gen_ip<SHAPE>(c, s4, c.b[0], c.b[1], c.db[2], sOut, 1.0, true, s1, s2);
                }
                break;
            case NEKMF_IPRODUCTWRTDERIVBASE:
            {
                // t_d = sum_c df[c*dim+d] in_c   (IProductWRTDerivBase.h:1294-1338)
                GEN_FOR(t, nqTot)
                {
                    if (dim == 3)
                    {
                        const double x = sIn[t], y = s3[t], z = s4[t];
                        double t0 = GDF(0) * x + GDF(3) * y + GDF(6) * z;
                        double t1 = GDF(1) * x + GDF(4) * y + GDF(7) * z;
                        const double t2 = GDF(2) * x + GDF(5) * y + GDF(8) * z;
                        if (SHAPE == NEKMF_PRISM)
                        {
                            // collapsed (xi_0, xi_2): IProductWRTDerivBase.h:1697-1733
                            const int i = t % c.nq0, k = t / (c.nq0 * c.nq1);
                            const double f0 = 2.0 / (1.0 - c.Z[2][k]), hf1 = 0.5 * (1.0 + c.Z[0][i]);
                            t0 = t0 * f0 + (hf1 * t2) * f0;
                        }
                        else if (SHAPE == NEKMF_PYR)
                        {
                            // IProductWRTDerivBase.h:2121-2168
                            const int i = t % c.nq0, j = (t / c.nq0) % c.nq1, k = t / (c.nq0 * c.nq1);
                            const double f0 = 2.0 / (1.0 - c.Z[2][k]);
                            t0 = t0 * f0 + (0.5 * (1.0 + c.Z[0][i]) * t2) * f0;
                            t1 = t1 * f0 + (0.5 * (1.0 + c.Z[1][j]) * t2) * f0;
                        }
                        else if (SHAPE == NEKMF_TET)
                        {
                            // IProductWRTDerivBase.h:2551-2603
                            const int i = t % c.nq0, j = (t / c.nq0) % c.nq1, k = t / (c.nq0 * c.nq1);
                            const double z1 = c.Z[1][j];
                            const double f2 = 2.0 / (1.0 - c.Z[2][k]), f3 = 0.5 * (1.0 + z1), f0 = 2.0 * f2 / (1.0 - z1);
                            const double f1 = 0.5 * (1.0 + c.Z[0][i]);
                            t0 = (t0 + (t1 + t2) * f1) * f0;
                            t1 = (t1 + t2 * f3) * f2;
                        }
                        sIn[t] = t0;
                        s3[t]  = t1;
                        s4[t]  = t2;
                    }
                    else
                    {
                        const double x = sIn[t], y = s3[t];
                        double t0 = GDF(0) * x + GDF(2) * y;
                        const double t1 = GDF(1) * x + GDF(3) * y;
                        if (SHAPE == NEKMF_TRI)
                        {
                            // IProductWRTDerivBase.h:1006-1036
                            const int i = t % c.nq0, j = t / c.nq0;
                            const double f0 = 2.0 / (1.0 - c.Z[1][j]), hf1 = 0.5 * (1.0 + c.Z[0][i]);
                            t0 = t0 * f0 + (hf1 * t1) * f0;
                        }
                        sIn[t] = t0;
                        s3[t]  = t1;
                    }
                }
                gen_sync(c);
                gen_weight(c, sIn, sIn, jac, deformed, dim);
                gen_ip<SHAPE>(c, sIn, c.db[0], c.b[1], c.b[2], sOut, 1.0, false, s1, s2);
                gen_weight(c, s3, s3, jac, deformed, dim);
                gen_ip<SHAPE>(c, s3, c.b[0], c.db[1], c.b[2], sOut, 1.0, true, s1, s2);
                if (dim == 3)
                {
                    gen_weight(c, s4, s4, jac, deformed, dim);
                    gen_ip<SHAPE>(c, s4, c.b[0], c.b[1], c.db[2], sOut, 1.0, true, s1, s2);
                }
                break;
            }
        }
        GEN_FOR(t, nout) a.out0[(size_t)e * nout + t] = sOut[t];
        if (a.optype == NEKMF_PHYSDERIV)
        {
            GEN_FOR(t, nout)
            {
                a.out1[(size_t)e * nout + t] = s3[t];
                if (dim == 3) a.out2[(size_t)e * nout + t] = s4[t];
            }
        }
        gen_sync(c);
    }
}

struct GenState
{
    int N;
    size_t smem;
    int blocks_per_sm;
    int groups;
};

template <int SHAPE> static int gen_launch(nekmf_op_s *op, const double *const in[3], double *const out[3])
{
    GenState *st = static_cast<GenState *>(op->kstate);
    auto kern    = gen_kernel<SHAPE>;
    if (st->blocks_per_sm == 0)
    {
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st->smem));
        int nb = 0;
        NEKMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 256, st->smem));
        if (nb < 1)
        {
            set_error("generic kernel does not fit on an SM (smem %zu)", st->smem);
            return NEKMF_ERR_CUDA;
        }
        st->blocks_per_sm = nb;
    }
    else
    {
        // the attribute is per-function: another operator with a larger footprint may have raised it, never lowers
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st->smem));
    }
    GenArgs a;
    a.in0 = in[0]; a.in1 = in[1]; a.in2 = in[2];
    a.out0 = out[0]; a.out1 = out[1]; a.out2 = out[2];
    const size_t gstep = op->deformed ? (size_t)op->nqTot : 1;
    a.jac = op->d_jac ? op->d_jac + (size_t)op->run_e0 * gstep : nullptr;
    a.df  = op->d_df ? op->d_df + (size_t)op->run_e0 * gstep : nullptr;
    a.dfStride = (size_t)op->nElmt * gstep;
    a.tab = op->d_tab; a.tab_len = op->tab_len;
    for (int d = 0; d < 3; ++d)
        for (int t = 0; t < 5; ++t) a.off[d][t] = op->tab_off[d][t];
    a.nm = op->nm[0]; a.nq0 = op->nq[0]; a.nq1 = op->nq[1]; a.nq2 = op->nq[2];
    a.nmTot = op->nmTot; a.nqTot = op->nqTot; a.nElmt = op->run_ne; a.deformed = op->deformed;
    a.optype = op->optype; a.N = st->N; a.groups = st->groups; a.lambda = op->lambda;
    int grid = st->blocks_per_sm * NUM_SMS;
    const int need = (op->run_ne + st->groups - 1) / st->groups;
    if (grid > need) grid = need;
    if (grid < 1) return NEKMF_OK;
    kern<<<grid, 256, st->smem, op->run_stream>>>(a);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}

bool select_generic(nekmf_op_s *op)
{
    const int nm = op->nm[0], nq0 = op->nq[0], nq1 = op->nq[1], nq2 = op->dim == 3 ? op->nq[2] : 1;
    int N = op->nqTot > op->nmTot ? op->nqTot : op->nmTot;
    const int cands[] = {nq0 * nm * nm, nq1 * nq0 * nm, nq2 * nm * nm, nq2 * nq1 * nm, nq1 * nm,
                         nq2 * nm * (nm + 1) / 2 + 2 * nq2};
    for (int v : cands)
        if (v > N) N = v;
    N = (N + 1) & ~1;
    // thread groups: as many as keep a group's widest stage (N outputs) busy with at least 32 threads per ~N/4
    // outputs and fit three CTAs' worth of shared memory
    int groups = 8;
    while (groups > 1 && (256 / groups < 32 || N > 4 * (256 / groups) ||
                          (size_t)(op->tab_len + groups * 7 * N) * 8 > 72 * 1024))
        groups /= 2;
    const size_t smem = (size_t)(op->tab_len + groups * 7 * N) * 8 + (size_t)(2 * (nm + 1) + nm * nm + 2) * 4 + 16;
    if (smem > 227 * 1024) return false;
    GenState *st    = new GenState{N, smem, 0, groups};
    op->kstate      = st;
    op->kstate_free = [](void *p) { delete static_cast<GenState *>(p); };
    const char *sn[6] = {"Quad", "Tri", "Hex", "Prism", "Pyr", "Tet"};
    char name[96];
    snprintf(name, sizeof(name), "gen_kernel<%s>(op=%d,nm=%d,nq=%d)", sn[op->shape], op->optype, nm, nq0);
    op->kname = name;
    switch (op->shape)
    {
        case NEKMF_QUAD: op->launch = gen_launch<NEKMF_QUAD>; break;
        case NEKMF_TRI: op->launch = gen_launch<NEKMF_TRI>; break;
        case NEKMF_HEX: op->launch = gen_launch<NEKMF_HEX>; break;
        case NEKMF_PRISM: op->launch = gen_launch<NEKMF_PRISM>; break;
        case NEKMF_PYR: op->launch = gen_launch<NEKMF_PYR>; break;
        case NEKMF_TET: op->launch = gen_launch<NEKMF_TET>; break;
        default: return false;
    }
    return true;
}

} // namespace nekmf
