// hex_dispatch.cu -- picks the compile-time specialised hexahedral kernel for an operator.
#include "op_internal.h"

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
    switch (op->nm[0])
    {
        case 2: return hex_try_nm2(op);
        case 3: return hex_try_nm3(op);
        case 4: return hex_try_nm4(op);
        case 5: return hex_try_nm5(op);
        case 6: return hex_try_nm6(op);
        case 7: return hex_try_nm7(op);
        case 8: return hex_try_nm8(op);
        case 9: return hex_try_nm9(op);
        case 10: return hex_try_nm10(op);
        case 11: return hex_try_nm11(op);
    }
    return false;
}

bool select_quad_fast(nekmf_op_s *) { return false; }
void notify_geom_changed(nekmf_op_s *) {}
} // namespace nekmf
