/* oracle/mf_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, scalar, CPU restatement of the reference's matrix-free elemental operator
 * path (ITHACA-SEM / Nektar++ 5.0.0, library/MatrixFreeOps + the Polylib/Basis data
 * it consumes + AssemblyMap gather/assemble + the CG driver).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may load it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every function
 * here against oracle/_ref (the reference's own kernel headers + Polylib.cpp
 * compiled in place from /root/reference) and tests/golden/ holds vectors generated
 * from that reference build (tests/golden/make_golden.py).
 */
#ifndef MF_ORACLE_H
#define MF_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* LibUtilities::ShapeType subset, numbered like the C-ABI (include/nekmf_b200.h) */
enum { MFO_QUAD = 0, MFO_TRI = 1, MFO_HEX = 2, MFO_PRISM = 3, MFO_PYR = 4, MFO_TET = 5, MFO_SEG = 6 };
/* points types: Gauss-Lobatto-Legendre, Gauss-Radau-M (alpha=1|2, beta=0) */
enum { MFO_GLL = 0, MFO_GRJM_A1 = 1, MFO_GRJM_A2 = 2 };
/* basis types */
enum { MFO_MOD_A = 0, MFO_MOD_B = 1, MFO_MOD_C = 2, MFO_MOD_PYR_C = 3 };

/* ---- Polylib restatement (LibUtilities/Polylib/Polylib.cpp) */
void mfo_jacobfd(int np, const double *z, double *poly, double *polyd, int n, double alpha, double beta);
void mfo_zwglj(double *z, double *w, int np, double alpha, double beta);
void mfo_zwgrjm(double *z, double *w, int np, double alpha, double beta);
void mfo_Dglj(double *D, const double *z, int np, double alpha, double beta);
void mfo_Dgrjm(double *D, const double *z, int np, double alpha, double beta);

/* ---- 1-D points + basis (Foundations/GaussPoints.cpp, Basis.cpp) */
/* z,w: np ; D: np*np with D[k*np+i] = dh_k/dz(z_i) */
void mfo_points(int ptype, int np, double *z, double *w, double *D);
/* number of rows of bdata for a basis type */
int mfo_basis_rows(int btype, int nm);
/* bdata/dbdata: rows*np, b[m*np+i] */
void mfo_basis(int btype, int nm, int np, const double *z, const double *D, double *bdata, double *dbdata);

/* ---- element description: tables for one (shape, nm, nq0) with Nektar's default
 *      point/basis choices (SpatialDomains/MeshGraph.cpp:1609-1762) */
typedef struct mfo_elem mfo_elem;
mfo_elem *mfo_create(int shape, int nm, int nq0);
void mfo_destroy(mfo_elem *e);
int mfo_dim(const mfo_elem *e);
/* segments only: number of space dimensions the 1-D element is embedded in (1..3, default 1) */
void mfo_set_coordim(mfo_elem *e, int coordim);
int mfo_nmtot(const mfo_elem *e);
int mfo_nqtot(const mfo_elem *e);
int mfo_nq(const mfo_elem *e, int dir);
int mfo_ptype(const mfo_elem *e, int dir);
int mfo_btype(const mfo_elem *e, int dir);
int mfo_brows(const mfo_elem *e, int dir);
/* which: 0 bdata 1 dbdata 2 D 3 Z 4 w (raw quadrature weights) */
const double *mfo_table(const mfo_elem *e, int dir, int which);

/* ---- operators.  Array layouts are the reference's external ones:
 *   coefficients [elmt][mode], quadrature values [elmt][k][j][i],
 *   jac [elmt] (regular) | [elmt][nqTot] (deformed),
 *   df  [ndf][elmt] | [ndf][elmt*nqTot]  with df[c*dim+d] = d xi_d / d x_c          */
void mfo_bwdtrans(const mfo_elem *e, int nElmt, const double *in, double *out);
void mfo_iproduct(const mfo_elem *e, int nElmt, int deformed, const double *jac, const double *in, double *out);
void mfo_physderiv(const mfo_elem *e, int nElmt, int deformed, const double *df, const double *in,
                   double *out0, double *out1, double *out2);
void mfo_helmholtz(const mfo_elem *e, int nElmt, int deformed, const double *jac, const double *df,
                   double lambda, const double *in, double *out);
/* Quad, Tri, Hex, Prism, Tet (IProductWRTDerivBase.h:542,891,1232,1630,2484) */
int mfo_iproductwrtderivbase(const mfo_elem *e, int nElmt, int deformed, const double *jac, const double *df,
                             const double *in0, const double *in1, const double *in2, double *out);

/* ---- AssemblyMap (MultiRegions/AssemblyMap/AssemblyMapCG.cpp:2853-2923, Vmath.hpp:217-244) */
void mfo_global_to_local(int nLocal, const int *map, const double *sign, const double *glob, double *loc);
void mfo_assemble(int nLocal, int nGlobal, const int *map, const double *sign, const double *loc, double *glob);

/* ---- CG (LibUtilities/LinearAlgebra/NekLinSysIterCG.cpp:104-265) on the assembled Helmholtz
 * operator of one hex/any-shape collection: A = Assemble o Helmholtz o GlobalToLocal.
 * diag: inverse diagonal preconditioner entries for [nDir, nGlobal) or NULL (identity).
 * Returns the number of iterations; x, rhs are global vectors of size nGlobal.          */
int mfo_cg_helmholtz(const mfo_elem *e, int nElmt, int deformed, const double *jac, const double *df,
                     double lambda, int nLocal, int nGlobal, int nDir, const int *map, const double *sign,
                     const double *invdiag, const double *rhs, double *x, double tol, int maxiter,
                     double *final_eps);

int mfo_max_threads(void);
/* ---- ContField::v_HelmSolve chain with the elemental operators as callbacks (see mf_oracle.c) */
typedef void (*mfo_elop_fn)(void *ctx, const double *in, double *out);
typedef struct mfo_chain mfo_chain;
mfo_chain *mfo_chain_create(int nLocal, int nGlobal, int nDir, int nPhys, const int *map, const double *sign,
                       const double *invdiag);
void mfo_chain_destroy(mfo_chain *chain);
int mfo_chain_helmsolve(mfo_chain *chain, mfo_elop_fn iprod, void *ci, mfo_elop_fn helm, void *ch, mfo_elop_fn bwd, void *cb,
                        const double *forcing, double *inout, double *phys_out, double tol, int maxiter,
                        double *final_eps);
void mfo_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
