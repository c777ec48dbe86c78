/* nekmf_b200.h -- C ABI of the B200-native matrix-free elemental operator path.
 *
 * This is the drop-in boundary for ITHACA-SEM / Nektar++ 5.0.0's matrix-free hot path
 * (library/MatrixFreeOps behind library/Collections).  Every entry point below is what a
 * `Collections::Operator` subclass registered as the new ImplementationType `eB200`
 * would bind (see INTEGRATION.md for the adapter class):
 *
 *   nekmf_op_create      <- MatrixFree::GetOperatorFactory().CreateInstance(op_string, basis, nElmt)
 *                           (Collections/Helmholtz.cpp:452-478, MatrixFreeOps/Operator.hpp:205-301:
 *                           Helper<DIM,DEFORMED> copies bdata/dbdata/D/Z/W per direction)
 *   nekmf_op_set_geom    <- Operator::SetJac / SetDF  (MatrixFreeOps/Operator.hpp:38-50) fed from
 *                           CoalescedGeomData::GetJac / GetDerivFactors
 *                           (Collections/CoalescedGeomData.cpp:53-113, 251-313) -- the NON-interleaved
 *                           reference arrays, jac[nElmt(*nq)], df[ndf][nElmt(*nq)]
 *   nekmf_op_set_lambda  <- MatrixFree::Helmholtz::SetLambda (MatrixFreeOps/Operator.hpp:193)
 *   nekmf_op_apply       <- MatrixFree::{BwdTrans,IProduct,PhysDeriv,Helmholtz,IProductWRTDerivBase}::operator()
 *                           (MatrixFreeOps/Operator.hpp:56-202) as called from the Collections wrappers
 *                           (Collections/BwdTrans.cpp:154-175, IProductWRTBase.cpp:179-200,
 *                           PhysDeriv.cpp:281-341, Helmholtz.cpp:400-428, IProductWRTDerivBase.cpp:285-)
 *   nekmf_map_*          <- AssemblyMapCG::v_GlobalToLocal / v_Assemble
 *                           (MultiRegions/AssemblyMap/AssemblyMapCG.cpp:2853-2923; Vmath::Gathr/Assmb,
 *                           LibUtilities/BasicUtils/Vmath.hpp:217-244)
 *   nekmf_exchange_*     <- AssemblyMapCG::v_UniversalAssemble -> Gs::Gather(gs_add)
 *                           (AssemblyMapCG.cpp:2925-2939, LibUtilities/Communication/GsLib.hpp:145-151)
 *   nekmf_cg_*           <- NekLinSysIterCG::DoConjugateGradient (LibUtilities/LinearAlgebra/
 *                           NekLinSysIterCG.cpp:104-265) with GlobalLinSysIterativeFull::v_DoMatrixMultiply
 *                           (MultiRegions/GlobalLinSysIterativeFull.cpp:215-257) as the mat-vec
 *
 * Conventions: every function returns an int status (NEKMF_OK == 0), never throws, keeps no
 * global state besides a thread-local error string.  Array layouts are the reference's external
 * ones: coefficients [elmt][mode], quadrature values [elmt][k][j][i],
 * jac [elmt] | [elmt][nq], df [ndf][elmt] | [ndf][elmt*nq] with df[c*dim+d] = d xi_d / d x_c.
 * All floating point data is FP64.  There is NO CPU fallback: without a CUDA device every
 * compute entry point returns NEKMF_ERR_CUDA.
 */
#ifndef NEKMF_B200_H
#define NEKMF_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NEKMF_ABI_VERSION 2

/* LibUtilities::ShapeType subset (LibUtilities/BasicUtils/ShapeType.hpp); numbering is this ABI's */
enum nekmf_shape
{
    NEKMF_QUAD  = 0,
    NEKMF_TRI   = 1,
    NEKMF_HEX   = 2,
    NEKMF_PRISM = 3,
    NEKMF_PYR   = 4,
    NEKMF_TET   = 5,
    NEKMF_SEG   = 6 /* 1-D element embedded in 1, 2 or 3 space dimensions (coordim) */
};

/* Collections::OperatorType, same order (Collections/Operator.h:65-73) */
enum nekmf_optype
{
    NEKMF_BWDTRANS             = 0,
    NEKMF_HELMHOLTZ            = 1,
    NEKMF_IPRODUCTWRTBASE      = 2,
    NEKMF_IPRODUCTWRTDERIVBASE = 3,
    NEKMF_PHYSDERIV            = 4
};

/* LibUtilities::BasisType / PointsType subset used by the path */
enum nekmf_basistype
{
    NEKMF_MODIFIED_A = 0,
    NEKMF_MODIFIED_B = 1,
    NEKMF_MODIFIED_C = 2,
    NEKMF_MODIFIEDPYR_C = 3 /* eModifiedPyr_C: rows (p,q,r), r fastest, nm - max(p,q) rows per (p,q) */
};
enum nekmf_pointstype
{
    NEKMF_GLL       = 0, /* eGaussLobattoLegendre */
    NEKMF_GRJM_A1B0 = 1, /* eGaussRadauMAlpha1Beta0 */
    NEKMF_GRJM_A2B0 = 2  /* eGaussRadauMAlpha2Beta0 */
};

/* where the caller's arrays live */
enum nekmf_memkind
{
    NEKMF_HOST   = 0, /* pageable or pinned host memory: H2D -> kernel -> D2H pipelined over element
                         chunks inside the call (3 streams), returns when the outputs are complete */
    NEKMF_DEVICE = 1  /* device pointers: no copies, asynchronous on the operator's stream */
};

enum nekmf_status
{
    NEKMF_OK              = 0,
    NEKMF_ERR_ARG         = 1, /* bad argument (the reference would ASSERTL0 / NEKERROR(efatal)) */
    NEKMF_ERR_UNSUPPORTED = 2, /* shape/op/order outside the registered set (reference: NEKERROR "not implemented") */
    NEKMF_ERR_CUDA        = 3, /* CUDA runtime failure or no device */
    NEKMF_ERR_STATE       = 4, /* geometry / lambda not set before apply */
    NEKMF_ERR_COMM        = 5, /* NCCL failure, or a peer-memory wait that timed out */
    NEKMF_ERR_NOCONVERGE  = 6  /* CG reached the iteration cap: the reference's NEKERROR(efatal, "Exceeded maximum
                                  number of iterations") (NekLinSysIterCG.cpp:190-203); x, iterations and final_eps
                                  are still filled in */
};

typedef struct nekmf_op_s *nekmf_op_t;
typedef struct nekmf_map_s *nekmf_map_t;
typedef struct nekmf_exchange_s *nekmf_exchange_t;
typedef struct nekmf_cg_s *nekmf_cg_t;
typedef struct nekmf_helmsolve_s *nekmf_helmsolve_t;
typedef struct nekmf_comm_s *nekmf_comm_t;

/* ---- library ---------------------------------------------------------------------------- */
int nekmf_abi_version(void);
/* message of the last failing call on this thread ("" if none) */
const char *nekmf_last_error(void);
/* number of visible CUDA devices (0 without a GPU; never fails) */
int nekmf_device_count(void);
int nekmf_set_device(int device);
/* number of kernels this library has launched in this process (bench.py "gpu_launches") */
long long nekmf_launch_count(void);

/* ---- device memory helpers (so a C++ host without the CUDA runtime can keep arrays resident) */
int nekmf_malloc_device(void **ptr, size_t bytes);
int nekmf_free_device(void *ptr);
int nekmf_malloc_pinned(void **ptr, size_t bytes);
int nekmf_free_pinned(void *ptr);
/* page-lock an existing host allocation (e.g. the Array<OneD> storage behind ExpList::m_coeffs /
 * m_phys, LibUtilities/BasicUtils/SharedArray.hpp) so NEKMF_HOST applies copy at full PCIe rate and
 * overlap H2D / kernel / D2H; without it the driver stages pageable memory through its own buffers */
int nekmf_host_register(void *ptr, size_t bytes);
int nekmf_host_unregister(void *ptr);
int nekmf_memcpy_h2d(void *dst, const void *src, size_t bytes);
int nekmf_memcpy_d2h(void *dst, const void *src, size_t bytes);
int nekmf_memset_device(void *dst, int value, size_t bytes);
int nekmf_sync(void);

/* ---- 1-D tables (stand-alone use; a Nektar++ build passes its own Basis/Points arrays) ---- */
/* LibUtilities::PointsManager()[key]->GetZW / GetD (Foundations/GaussPoints.cpp:69-236):
 * z[np], w[np] raw weights, D[np*np] with D[k*np+i] = dh_k/dz(z_i) (D may be NULL) */
int nekmf_points(int pointstype, int np, double *z, double *w, double *D);
/* rows of bdata for a basis type: nm | nm(nm+1)/2 | nm(nm+1)(nm+2)/6 */
int nekmf_basis_rows(int basistype, int nm);
/* LibUtilities::Basis::GenBasis (Foundations/Basis.cpp:392-567): bdata/dbdata rows x np */
int nekmf_basis(int basistype, int nm, int np, const double *z, const double *D, double *bdata, double *dbdata);

/* ---- elemental operators ---------------------------------------------------------------- */
/* nm[d], nq[d]: modes / quadrature points per direction (dim entries used; dim = 2 for
 * Quad/Tri, 3 otherwise).  bdata[d]/dbdata[d]: rows(d) x nq[d], b[m*nq+i]; D[d]: nq x nq with
 * D[k*nq+i] = dh_k/dz(z_i); Z[d], W[d]: points and raw quadrature weights (the collapsed-
 * coordinate 0.5 / 0.25 weight scaling of MatrixFreeOps/Operator.hpp:244-258 is applied here).
 * All tables are copied.  deformed: 0 = one jac/df entry per element, 1 = per quadrature point.
 * coordim must equal dim (the reference's 2-D kernels reject 3 outputs, PhysDeriv.h:285-286), except for
 * NEKMF_SEG where coordim = 1, 2 or 3 gives the number of PhysDeriv outputs / IProductWRTDerivBase inputs and of
 * derivative-factor rows (PhysDeriv.h:60-250, IProductWRTDerivBase.h:182-365); Helmholtz has no segment variant. */
int nekmf_op_create(int shape, int optype, const int nm[3], const int nq[3], const int basistype[3],
                    const int pointstype[3], const double *const bdata[3], const double *const dbdata[3],
                    const double *const D[3], const double *const Z[3], const double *const W[3], int nElmt,
                    int deformed, int coordim, nekmf_op_t *op);
/* jac: [nElmt] or [nElmt*nqTot]; df: [ndf][nElmt] or [ndf][nElmt*nqTot] (row n contiguous).
 * Either may be NULL when the operator does not need it (BwdTrans needs neither; IProduct
 * needs jac; PhysDeriv needs df).  memkind says where jac/df live; they are copied. */
int nekmf_op_set_geom(nekmf_op_t op, const double *jac, const double *df, int memkind);
int nekmf_op_set_lambda(nekmf_op_t op, double lambda);
/* in0..2 / out0..2 by operator:
 *   BwdTrans: in0 coeffs -> out0 phys;  IProductWRTBase: in0 phys -> out0 coeffs;
 *   PhysDeriv: in0 phys -> out0,out1(,out2) phys;  Helmholtz: in0 coeffs -> out0 coeffs;
 *   IProductWRTDerivBase: in0,in1(,in2) phys -> out0 coeffs.
 * memkind NEKMF_HOST: synchronous, copies inside.  NEKMF_DEVICE: enqueued on the operator's
 * stream (nekmf_op_set_stream), returns immediately. in and out must not alias. */
int nekmf_op_apply(nekmf_op_t op, const double *in0, const double *in1, const double *in2, double *out0,
                   double *out1, double *out2, int memkind);
/* cudaStream_t as void*; NULL = default stream */
int nekmf_op_set_stream(nekmf_op_t op, void *stream);
int nekmf_op_ncoeff(nekmf_op_t op); /* modes per element (MatrixFree::Operator::Ndof) */
int nekmf_op_nphys(nekmf_op_t op);  /* quadrature points per element */
/* name of the kernel variant apply() launches ("hex_helm_t<5,6,def>" ...) */
const char *nekmf_op_kernel_name(nekmf_op_t op);
/* duration in ms of the last device-side apply, measured with CUDA events on the operator's
 * stream (synchronises); < 0 if timing is disabled */
int nekmf_op_enable_timing(nekmf_op_t op, int on);
int nekmf_op_last_ms(nekmf_op_t op, float *ms);
int nekmf_op_destroy(nekmf_op_t op);

/* ---- AssemblyMap: local <-> global ------------------------------------------------------- */
/* localToGlobal[nLocal] in [0,nGlobal); sign[nLocal] = +-1 or NULL (m_signChange == false). */
int nekmf_map_create(int nLocal, int nGlobal, const int *localToGlobal, const double *sign, nekmf_map_t *map);
/* loc[i] = sign[i] * glob[map[i]]   (AssemblyMapCG::v_GlobalToLocal) */
int nekmf_map_global_to_local(nekmf_map_t map, const double *glob, double *loc, int memkind, void *stream);
/* glob = 0; glob[map[i]] += sign[i] * loc[i]   (AssemblyMapCG::v_Assemble, local part).
 * Deterministic: every global DOF sums its local copies in ascending local index, exactly the
 * order of the reference's sequential Vmath::Assmb. */
int nekmf_map_assemble(nekmf_map_t map, const double *loc, double *glob, int memkind, void *stream);
int nekmf_map_destroy(nekmf_map_t map);

/* ---- communicator (one rank per GPU) ------------------------------------------------------ */
/* Replaces LibUtilities::Comm for this path (AllReduce of the CG dot products, NekLinSysIterCG.cpp:226-235, and
 * the transport under Gs::Gather).  Bootstrapped with NCCL: 128-byte ncclUniqueId produced on rank 0 and
 * distributed by the host (torch.distributed, MPI ...).  COLLECTIVE: every rank calls create.  On an NVLink box
 * create also exports one small reduction window per rank with CUDA IPC and maps all of them: the data path then
 * runs over peer memory (kernels store into / spin on the windows) without NCCL calls.  NEKMF_TRANSPORT=nccl in
 * the environment, or a mapping failure on any rank, keeps every rank on NCCL send/recv/all-reduce instead. */
int nekmf_comm_unique_id(unsigned char id[128]);
int nekmf_comm_create(const unsigned char id[128], int rank, int nranks, nekmf_comm_t *comm);
/* 1 = peer-memory transport (NVLink loads/stores), 0 = NCCL calls */
int nekmf_comm_transport(nekmf_comm_t comm);
int nekmf_comm_destroy(nekmf_comm_t comm);

/* ---- interface-DOF exchange: the Gs::Gather(gs_add) replacement -------------------------- */
/* For each of nNeighbours (<= 255) peer ranks, the (identically ordered on both sides) list of this
 * rank's global indices in [0,nGlobal) shared with that peer: idx[offsets[n] .. offsets[n+1]).  A DOF held by k
 * ranks appears in k-1 lists of each holder (what Gs::Init derives from m_globalToUniversalMap,
 * AssemblyMapCG.cpp:2563-2565).  comm may be NULL when nNeighbours == 0.  COLLECTIVE when comm has more than one
 * rank and uses the peer-memory transport (the receive windows are exported / mapped here). */
int nekmf_exchange_create(nekmf_comm_t comm, int nGlobal, int nNeighbours, const int *peerRanks, const int *offsets,
                          const int *idx, nekmf_exchange_t *ex);
/* Every copy of a shared DOF ends up holding the sum over all copies (gs_add), device memory.  The copies are
 * added in ascending RANK order by one thread per DOF (no atomics): deterministic, and all holders of a DOF
 * compute the bit-identical sum.  Peer-memory transport: deposit kernel (NVLink stores + flags) -> wait/add
 * kernel; NCCL transport: pack -> grouped ncclSend/ncclRecv -> the same add kernel. */
int nekmf_exchange_add(nekmf_exchange_t ex, double *glob, void *stream);
int nekmf_exchange_destroy(nekmf_exchange_t ex);

/* ---- conjugate gradient on A = Assemble o Helmholtz o GlobalToLocal ----------------------- */
/* op: a Helmholtz operator with geometry and lambda set.  nDir: number of leading Dirichlet
 * DOFs (CG runs on [nDir,nGlobal)).  invdiag[nGlobal-nDir]: inverse-diagonal preconditioner or
 * NULL (identity).  ownerMask[nGlobal]: 1.0 where this rank owns the DOF for dot products
 * (Gs::Unique, AssemblyMapCG.cpp:2565-2569) or NULL (all owned).  ex/comm NULL for one rank.
 * Host arrays; copied once. */
int nekmf_cg_create(nekmf_op_t op, nekmf_map_t map, nekmf_exchange_t ex, nekmf_comm_t comm, int nDir,
                    const double *invdiag, const double *ownerMask, nekmf_cg_t *cg);
/* rhs, x: global vectors [nGlobal] (memkind as given).  Follows DoConjugateGradient: x[nDir:]
 * starts at 0, stops when r.r < tol^2 * rhs_magnitude (rhs.rhs, or 1 when that is below 1e-6).  Returns the
 * iteration count (m_totalIterations) and the final r.r.  When the loop counter reaches maxiter the reference
 * raises a fatal error; here the call returns NEKMF_ERR_NOCONVERGE with x / iterations / final_eps filled.
 * alpha, beta and the convergence test live on the device; iterations are replayed from CUDA graphs and the
 * host reads the state one graph behind (cg.cu).  Fails with NEKMF_ERR_STATE when the operator has no geometry
 * or lambda. */
int nekmf_cg_solve(nekmf_cg_t cg, const double *rhs, double *x, int memkind, double tol, int maxiter,
                   int *iterations, double *final_eps);
/* device time (CUDA events on the solver's stream) of the iteration loop of the last solve and the number of
 * iterations launched inside it (set-up, the first mat-vec and the final copies excluded); ms < 0 if none ran */
int nekmf_cg_last_loop(nekmf_cg_t cg, float *ms, int *iterations);
/* one mat-vec s = A w on device global vectors (exposed for tests and the benchmark) */
int nekmf_cg_matvec(nekmf_cg_t cg, const double *w, double *s);
int nekmf_cg_destroy(nekmf_cg_t cg);

/* ---- matrix-free Jacobi preconditioner ---------------------------------------------------- */
/* Replaces PreconditionerDiagonal::DiagonalPreconditionerSum (MultiRegions/PreconditionerDiagonal.cpp:98-162), which
 * sums loc_mat(i,i) of the assembled elemental matrices.  nekmf_op_diagonal: the diagonal of every elemental
 * Helmholtz matrix, diag[nElmt*ncoeff], obtained from the operator itself (ncoeff applies to unit vectors: any
 * shape, regular or deformed, the kernel the mat-vec uses).  nekmf_cg_set_jacobi: that diagonal assembled with the
 * solver's map, summed across the partition interfaces and inverted on the device, installed as the solver's
 * preconditioner (replacing the invdiag given at create, if any).  COLLECTIVE over the solver's communicator. */
int nekmf_op_diagonal(nekmf_op_t op, double *diag, int memkind);
int nekmf_cg_set_jacobi(nekmf_cg_t cg);

/* ---- ContField::v_HelmSolve as one device-resident chain ---------------------------------- */
/* Replaces the string of host-array calls under MultiRegions/ContField.cpp:878-945 (v_HelmSolve) -> GlobalSolve
 * (:516-535) -> GlobalLinSysIterativeFull::v_Solve (GlobalLinSysIterativeFull.cpp:110-211):
 *   wsp = -IProductWRTBase(forcing); tmp1 = wsp - Helmholtz(inout); rhs = Assemble(tmp1) (+ interface exchange);
 *   global = CG(rhs); inout += GlobalToLocal(global)         [no Dirichlet DOF on any rank: rhs = Assemble(wsp),
 *   inout = GlobalToLocal(global)], and, when phys_out != NULL, phys_out = BwdTrans(inout) -- the evaluation the
 * solvers perform next.  cg carries the Helmholtz operator (lambda set), the assembly map, the exchange and the
 * preconditioner; iprod / bwd are the IProductWRTBase / BwdTrans operators of the same collection (bwd may be
 * NULL).  create is COLLECTIVE over cg's communicator (it sums nDir over the ranks). */
int nekmf_helmsolve_create(nekmf_cg_t cg, nekmf_op_t iprod, nekmf_op_t bwd, nekmf_helmsolve_t *hs);
/* forcing: f at the quadrature points [nElmt*nq]; inout: local coefficients [nLocal] holding the Dirichlet values
 * and the initial guess on entry, the solution on return; phys_out: [nElmt*nq] or NULL.  memkind NEKMF_HOST: the
 * three arrays are host arrays, each crosses PCIe once (forcing and inout in, inout and phys_out back);
 * NEKMF_DEVICE: device arrays, nothing is copied.  Synchronous.  Returns what nekmf_cg_solve returns
 * (NEKMF_ERR_NOCONVERGE leaves the capped solve's result in the outputs). */
int nekmf_helmsolve(nekmf_helmsolve_t hs, const double *forcing, double *inout, double *phys_out, int memkind, double tol,
                    int maxiter, int *iterations, double *final_eps);
/* device time of the last call, copies included (CUDA events on the solver's stream); ms < 0 if none ran */
int nekmf_helmsolve_last_ms(nekmf_helmsolve_t hs, float *ms);
/* the same call split into its phases (profiling hook): ms[0] host -> device copies, ms[1] IProductWRTBase + Dirichlet
 * lift + Assemble, ms[2] the CG solve, ms[3] GlobalToLocal + BwdTrans, ms[4] device -> host copies; all < 0 if none ran */
int nekmf_helmsolve_last_phases(nekmf_helmsolve_t hs, float ms[5]);
int nekmf_helmsolve_destroy(nekmf_helmsolve_t hs);

#ifdef __cplusplus
}
#endif
#endif /* NEKMF_B200_H */
