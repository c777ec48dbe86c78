// op_internal.h -- the operator object behind nekmf_op_t and the launcher registry.
#pragma once
#include "../../include/nekmf_b200.h"
#include "common.cuh"
#include <string>
#include <vector>

struct nekmf_op_s
{
    int shape = 0, optype = 0, dim = 0, coordim = 0; // coordim != dim only for segments
    int nm[3] = {1, 1, 1}, nq[3] = {1, 1, 1}, btype[3] = {0, 0, 0}, ptype[3] = {0, 0, 0}, rows[3] = {0, 0, 0};
    int nmTot = 0, nqTot = 0, nElmt = 0, deformed = 0, ndf = 0;
    double lambda       = 0.0;
    bool lambda_set     = false;
    bool has_jac        = false;
    bool has_df         = false;
    // host copies of the 1-D tables; ws = weights with the collapsed-coordinate factor folded in
    std::vector<double> b[3], db[3], D[3], Z[3], W[3], ws[3];
    // device copies for the runtime-sized kernels: packed [b0 db0 D0 Z0 ws0 | b1 ... | b2 ...]
    double *d_tab = nullptr;
    int tab_off[3][5] = {{0}};
    int tab_len       = 0;
    // device geometry.  regular: jac[nElmt], df[ndf][nElmt].  deformed: jac[nElmt][geo_pitch],
    // df[ndf][nElmt][geo_pitch] with geo_pitch >= nqTot chosen by the selected kernel (the TMA-fed
    // kernels need every element block 16-byte aligned; the runtime-sized kernels use nqTot)
    double *d_jac     = nullptr;
    double *d_df      = nullptr;
    int geo_pitch     = 0;
    cudaStream_t stream = nullptr;
    // what one launch covers: elements [run_e0, run_e0 + run_ne) on run_stream.  nekmf_op_apply sets these
    // before every launch (whole collection on `stream` for device arrays; one chunk per pipeline stage for
    // host arrays).  in/out pointers handed to the launcher already point at element run_e0; launchers
    // offset the geometry themselves (the df row stride stays nElmt).
    int run_e0 = 0, run_ne = 0;
    cudaStream_t run_stream = nullptr;
    // staging + copy pipeline for NEKMF_HOST applies
    double *d_stage_in = nullptr, *d_stage_out = nullptr;
    size_t stage_in_sz = 0, stage_out_sz = 0;
    cudaStream_t pipe_stream[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t pipe_done = nullptr;
    std::vector<cudaEvent_t> pipe_ev; // [2*chunk]: copy-in done, kernel done 
    // launcher: device pointers only
    int (*launch)(nekmf_op_s *, const double *const in[3], double *const out[3]) = nullptr;
    std::string kname;
    void *kstate = nullptr; // launcher-private (constant tables etc.)
    void (*kstate_free)(void *) = nullptr;
    // fused AssemblyMap gather (CG mat-vec): when gather_ok, a caller may set gather_map (+ optional sign) and
    // pass the GLOBAL vector as in[0]; the kernel then loads sign[i]*in[map[i]] itself.  Reset to null after.
    bool gather_ok           = false;
    const int *gather_map    = nullptr;
    const double *gather_sign = nullptr;
    // with gather_map: device flag that turns the launch into a no-op when non-zero (iterations a solver enqueued
    // before it knew it had converged)
    const int *gather_skip = nullptr;
    int kron         = 0; // regular Helmholtz: coefficient-space kernel available (1 hex_kron.cu, 2 quad_kron.cu, 3 dense_helm.cu, 4 its prism variant)
    bool timing      = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed_once = false;
};

namespace nekmf
{
// geometry / lambda set? (NEKMF_ERR_STATE otherwise); every entry point that launches the operator calls it
int op_check_ready(nekmf_op_s *op);
// each returns true if it installed a launcher for this operator
bool select_hex_fast(nekmf_op_s *op);
bool select_shape_fast(nekmf_op_s *op);
bool select_generic(nekmf_op_s *op);
bool select_seg(nekmf_op_s *op);
bool select_quad_lane(nekmf_op_s *op);
bool select_tri_lane(nekmf_op_s *op);
// called after set_geom / set_lambda so launchers can precompute (e.g. detect diagonal metrics)
int notify_geom_changed(nekmf_op_s *op); // status of the launchers' geometry pre-passes
void hex_dmma_maybe_wrap(nekmf_op_s *op); // hex_dmma.cu: DMMA BwdTrans / IProductWRTBase at nm = 7
void prism_dmma_maybe_wrap(nekmf_op_s *op); // prism_dmma.cu: DMMA BwdTrans / IProductWRTBase on prisms, nm = 3..7
void pyr_dmma_maybe_wrap(nekmf_op_s *op);   // the same kernel family on pyramids
void tet_dmma_maybe_wrap(nekmf_op_s *op);   // tet_dmma.cu: DMMA + lane-per-mode-pair BwdTrans / IProductWRTBase on tetrahedra, nm = 5..7
bool tet_gemm_maybe_wrap(nekmf_op_s *op);  // tet_gemm.cu: BwdTrans on tetrahedra as DMMA GEMMs over eight elements, nm = 5..7
void kron_maybe_wrap(nekmf_op_s *op);
int kron_geom_changed(nekmf_op_s *op);
void quad_kron_maybe_wrap(nekmf_op_s *op);
int quad_kron_geom_changed(nekmf_op_s *op);
void dense_maybe_wrap(nekmf_op_s *op); // dense_helm.cu: DMMA coefficient-space Helmholtz (regular Tri / Tet / Pyr)
int dense_geom_changed(nekmf_op_s *op);
void prism_maybe_wrap(nekmf_op_s *op); // dense_helm.cu: extruded prisms as nm triangle problems per element
int prism_geom_changed(nekmf_op_s *op);
// prism_helm_dmma.cu: fused quadrature-space Helmholtz for general regular prisms on tensor tiles (nm = 5..7)
void *prism_helm_fused_create(nekmf_op_s *op);
int prism_helm_fused_launch(void *state, nekmf_op_s *op, const double *in, double *out);
void prism_helm_fused_free(void *state, int nm);
} // namespace nekmf
