// hex_dispatch.cu -- picks the compile-time specialised hexahedral kernel for an operator.
#include "op_internal.h"
#include <stdlib.h>

namespace nekmf
{
#define HEX_DECL(n) bool hex_try_nm##n(nekmf_op_s *op);
HEX_DECL(2) HEX_DECL(3) HEX_DECL(4) HEX_DECL(5) HEX_DECL(6) HEX_DECL(7) HEX_DECL(8) HEX_DECL(9) HEX_DECL(10) HEX_DECL(11)
#undef HEX_DECL

bool select_hex_fast(nekmf_op_s *op)
{
    if (op->shape != NEKMF_HEX) return false;
    // isotropic tables are a precondition of the reference too (Helmholtz.h:42-44)
    for (int d = 1; d < 3; ++d)
        if (op->nm[d] != op->nm[0] || op->nq[d] != op->nq[0] || op->b[d] != op->b[0] || op->D[d] != op->D[0] ||
            op->ws[d] != op->ws[0])
            return false;
    bool ok = false;
    switch (op->nm[0])
    {
        case 2: ok = hex_try_nm2(op); break;
        case 3: ok = hex_try_nm3(op); break;
        case 4: ok = hex_try_nm4(op); break;
        case 5: ok = hex_try_nm5(op); break;
        case 6: ok = hex_try_nm6(op); break;
        case 7: ok = hex_try_nm7(op); break;
        case 8: ok = hex_try_nm8(op); break;
        case 9: ok = hex_try_nm9(op); break;
        case 10: ok = hex_try_nm10(op); break;
        case 11: ok = hex_try_nm11(op); break;
    }
    // regular Helmholtz: add the coefficient-space kernel (used when the metric is diagonal)
    if (ok) kron_maybe_wrap(op);
    if (ok) hex_dmma_maybe_wrap(op); // nm = 7 BwdTrans / IProductWRTBase: tensor-core tiles
    return ok;
}

#define SHP_DECL(n) bool shape_try_nm##n(nekmf_op_s *op);
SHP_DECL(2) SHP_DECL(3) SHP_DECL(4) SHP_DECL(5) SHP_DECL(6) SHP_DECL(7) SHP_DECL(8) SHP_DECL(9)
#undef SHP_DECL

// Quad / Tri / Prism / Pyr / Tet with the default quadrature: compile-time sized kernels (shape_kernels.cuh)
bool select_shape_fast(nekmf_op_s *op)
{
    if (op->shape == NEKMF_HEX) return false;
    if (op->shape == NEKMF_PYR && op->optype != NEKMF_PHYSDERIV)
    {
        const char *v = getenv("NEKMF_PYR_SHAPE"); // NEKMF_PYR_SHAPE=0: runtime-sized kernels (A/B comparisons)
        if (v && v[0] == '0') return false;
    }
    if (select_quad_lane(op)) return true; // BwdTrans / IProductWRTBase / regular PhysDeriv on quads: one lane per element
    if (select_tri_lane(op)) return true;  // the same for triangles
    bool ok = false;
    switch (op->nm[0])
    {
        case 2: ok = shape_try_nm2(op); break;
        case 3: ok = shape_try_nm3(op); break;
        case 4: ok = shape_try_nm4(op); break;
        case 5: ok = shape_try_nm5(op); break;
        case 6: ok = shape_try_nm6(op); break;
        case 7: ok = shape_try_nm7(op); break;
        case 8: ok = shape_try_nm8(op); break;
        case 9: ok = shape_try_nm9(op); break;
    }
    // regular quad Helmholtz: add the coefficient-space kernel (used when the metric is diagonal)
    if (ok) quad_kron_maybe_wrap(op);
    if (ok) prism_dmma_maybe_wrap(op); // prisms, BwdTrans / IProductWRTBase: tensor-core tiles
    if (ok && !tet_gemm_maybe_wrap(op)) tet_dmma_maybe_wrap(op); // tetrahedra: GEMM over eight elements (BwdTrans), else tiles + lane per mode pair
    return ok;
}
int notify_geom_changed(nekmf_op_s *op)
{
    int rc = kron_geom_changed(op);
    if (!rc) rc = quad_kron_geom_changed(op);
    if (!rc) rc = dense_geom_changed(op);
    if (!rc) rc = prism_geom_changed(op);
    return rc;
}
} // namespace nekmf
