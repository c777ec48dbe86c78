// dense_helm.cu -- coefficient-space Helmholtz for REGULAR (affine) triangles, tetrahedra and pyramids as a
// batched FP64 tensor-core GEMM (DMMA, mma.sync.m8n8k4.f64).  Hand-written for sm_100a.
//
// Reference semantics: MatrixFreeOps/Helmholtz.h:506-635 (Tri), 1771-1955 (Pyr), 2266-2448 (Tet) -- BwdTrans,
// collapsed-coordinate PhysDeriv, Laplacian metric, three IProducts with the transposed derivative, mass term.
// On a regular element the geometric factors are constant, and the whole chain is linear in the input and in
//   G = df^T df  (G_ab = sum_c df[c*dim+a] df[c*dim+b], the h-factors of the collapsed coordinates only depend on
// the quadrature point), so
//   out_e = J_e ( lambda M + sum_{a<=b} G_ab(e) K_ab ) u_e
// with nT = 1 + dim(dim+1)/2 reference-element matrices shared by EVERY element of the collection (M; K_aa;
// K_ab + K_ba).  The quadrature-space kernels spend 240-280 kflop-equivalents of FP64 pipe time per P=6 tet on
// pencil passes through shared memory; the dense form costs 2 nT n^2 = 99 kflop at n = 84 and is a plain GEMM
//   Out (n x nElmt) = [A_0 | A_1 | ... ] (n x nT n)  *  [c_0(e) u_e ; c_1(e) u_e ; ...] (nT n x nElmt)
// which is FP64-pipe bound -- the one place on this path where the DMMA tiles of BASELINE.json's north_star pay.
//
// The matrices are measured, not re-derived: at the first set_geom the operator's own quadrature-space kernel
// (shape_op_kernel / gen_kernel, the pinned implementation) is applied to unit vectors on probe geometries
// (lambda = 1, df = 0 -> M;  lambda = 0, df = e_a -> K_aa;  df = e_a + e_b -> K_aa + K_bb + K_ab + K_ba) and the
// result is repacked in DMMA fragment order.  Kernel: one CTA = 4 warps = 64 (32) elements; a warp owns 16 (8)
// elements and ALL output rows (MT tiles of 8 rows x NT tiles of 8 elements of accumulators in registers); its
// element coefficients live in a padded shared-memory tile (pitch = 4 mod 8: the B-fragment loads are
// conflict-free), the per-element scale c_t(e) multiplies the B fragment, and the A fragments stream from L2
// through a four-stage cp.async ring shared by the four warps.
#include "common.cuh"
#include "op_internal.h"
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

namespace nekmf
{

constexpr int DW      = 4; // warps per CTA (two or more CTAs per SM: one's tile load / store overlaps another's DMMAs)
constexpr int DSTAGES = 4; // A-fragment ring

template <int MT> struct DenseCfg
{
    static constexpr int NT    = MT <= 18 ? 2 : 1;          // element tiles (of 8) per warp
    static constexpr int NE    = DW * 8 * NT;               // elements per CTA
    static constexpr int KC    = MT <= 4 ? 8 : MT <= 12 ? 24 / MT : MT <= 18 ? 2 : 1; // k-steps (of 4) per ring stage
    static constexpr int CHUNK = KC * MT * 32;              // doubles per ring stage (<= 8 KB: three CTAs per SM at n = 84)
};

struct DenseArgs
{
    const double *in;
    double *out;
    const double *jac, *df, *afrag;
    size_t dfStride;
    int nElmt, n, KS, G, nT, dim, pitch;
    double lambda;
};

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void *s, const void *g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int MT> __global__ void __launch_bounds__(DW * 32, 2) dense_helm_kernel(const __grid_constant__ DenseArgs a)
{
    using Cfg        = DenseCfg<MT>;
    constexpr int NT = Cfg::NT, NE = Cfg::NE, KC = Cfg::KC, CHUNK = Cfg::CHUNK;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sA = reinterpret_cast<double *>(smem_raw); // [DSTAGES][CHUNK]   A fragments
    double *sC = sA + DSTAGES * CHUNK;                 // [7][NE]            c_t(e)
    double *sU = sC + 7 * NE;                          // [NE][pitch]        element coefficients, later the output
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int e0 = blockIdx.x * NE, we = warp * 8 * NT;
    const int n = a.n, pitch = a.pitch, KS = a.KS;
    const int NC = (a.G + KC - 1) / KC;

    auto issue = [&](int c) {
        if (c < NC)
        {
            const int g0 = c * KC, g1 = (g0 + KC < a.G) ? g0 + KC : a.G;
            const int units   = (g1 - g0) * MT * 16; // 16-byte units
            const double *src = a.afrag + (size_t)g0 * MT * 32;
            double *dst       = sA + (c % DSTAGES) * CHUNK;
            for (int i = tid; i < units; i += DW * 32) cp_async16(dst + 2 * i, src + 2 * i);
        }
        cp_async_commit();
    };
    issue(0);
    issue(1);
    issue(2);

    // ---- per warp: c_t(e) of its own elements (J folded in), then the coefficient rows; no CTA barrier needed
    if (lane < 8 * NT)
    {
        const int el = we + lane, e = e0 + el;
        double c[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        if (e < a.nElmt)
        {
            const double J = a.jac[e];
            c[0]           = J * a.lambda;
            if (a.dim == 3)
            {
                double d[9];
#pragma unroll
                for (int q = 0; q < 9; ++q) d[q] = a.df[q * a.dfStride + e];
                c[1] = J * (d[0] * d[0] + d[3] * d[3] + d[6] * d[6]);
                c[2] = J * (d[1] * d[1] + d[4] * d[4] + d[7] * d[7]);
                c[3] = J * (d[2] * d[2] + d[5] * d[5] + d[8] * d[8]);
                c[4] = J * (d[0] * d[1] + d[3] * d[4] + d[6] * d[7]);
                c[5] = J * (d[0] * d[2] + d[3] * d[5] + d[6] * d[8]);
                c[6] = J * (d[1] * d[2] + d[4] * d[5] + d[7] * d[8]);
            }
            else
            {
                double d[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) d[q] = a.df[q * a.dfStride + e];
                c[1] = J * (d[0] * d[0] + d[2] * d[2]);
                c[2] = J * (d[1] * d[1] + d[3] * d[3]);
                c[3] = J * (d[0] * d[1] + d[2] * d[3]);
            }
        }
#pragma unroll
        for (int t = 0; t < 7; ++t) sC[t * NE + el] = c[t];
    }
    // coefficient rows of the warp's own elements: 8 NT independent loads in flight per lane and pass
    for (int k = lane; k < pitch; k += 32)
    {
        double v[8 * NT];
#pragma unroll
        for (int r = 0; r < 8 * NT; ++r)
        {
            const int e = e0 + we + r;
            v[r]        = (e < a.nElmt && k < n) ? __ldg(a.in + (size_t)e * n + k) : 0.0;
        }
#pragma unroll
        for (int r = 0; r < 8 * NT; ++r) sU[(size_t)(we + r) * pitch + k] = v[r];
    }
    __syncwarp();

    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const double *uB[NT], *cB[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j)
    {
        uB[j] = sU + (size_t)(we + j * 8 + (lane >> 2)) * pitch + (lane & 3);
        cB[j] = sC + we + j * 8 + (lane >> 2);
    }
    // One flat loop over the G = nT KS k-steps.  The operands of step g+1 (A fragments, scaled B fragments) are
    // fetched while the DMMAs of step g issue: a[i] is reloaded right after its last use, so no step starts with
    // a load-to-use bubble (the warps leave every barrier in lockstep and would all stall together).  The
    // ring barrier for stage c+1 therefore sits at the top of the LAST step of stage c.
    cp_async_wait<DSTAGES - 2>();
    __syncthreads(); // stage 0 visible
    double av[MT], b[NT];
#pragma unroll
    for (int i = 0; i < MT; ++i) av[i] = sA[i * 32 + lane];
#pragma unroll
    for (int j = 0; j < NT; ++j) b[j] = uB[j][0] * cB[j][0];
    int c = 0, gl = 0, t = 0, kk = 0;
    int cnt = a.G < KC ? a.G : KC;
#pragma unroll 1
    for (int g = 0; g < a.G; ++g)
    {
        const bool last_in_stage = gl == cnt - 1;
        if (last_in_stage && c + 1 < NC)
        {
            cp_async_wait<DSTAGES - 3>(); // this thread's copies of stage c+1 have landed
            __syncthreads();              // everyone's have, and everyone is done with stage c-1
            issue(c + DSTAGES - 1);       // refill the buffer stage c-1 occupied
        }
        int c2 = c, gl2 = gl + 1, kk2 = kk + 1, t2 = t;
        if (last_in_stage) { c2 = c + 1; gl2 = 0; }
        if (kk2 == KS) { kk2 = 0; t2 = t + 1; }
        if (g + 1 == a.G) { c2 = c; gl2 = gl; kk2 = kk; t2 = t; } // nothing follows: re-read this step (unused)
        const double *apn = sA + (c2 % DSTAGES) * CHUNK + gl2 * (MT * 32) + lane;
        double un[NT], cn[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j)
        {
            un[j] = uB[j][kk2 * 4];
            cn[j] = cB[j][t2 * NE];
        }
#pragma unroll
        for (int i = 0; i < MT; ++i)
        {
#pragma unroll
            for (int j = 0; j < NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], av[i], b[j]);
            av[i] = apn[i * 32];
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) b[j] = un[j] * cn[j];
        if (c2 != c) cnt = (a.G - c2 * KC < KC) ? a.G - c2 * KC : KC;
        c = c2; gl = gl2; kk = kk2; t = t2;
    }
    cp_async_wait<0>();

    // ---- epilogue: accumulators -> the warp's own rows of sU -> coalesced element-contiguous stores
    __syncwarp();
#pragma unroll
    for (int i = 0; i < MT; ++i)
    {
        const int m = i * 8 + (lane >> 2);
        if (m < n)
        {
#pragma unroll
            for (int j = 0; j < NT; ++j)
            {
                double *row = sU + (size_t)(we + j * 8 + 2 * (lane & 3)) * pitch + m;
                row[0]      = acc[i][j][0];
                row[pitch]  = acc[i][j][1];
            }
        }
    }
    __syncwarp();
    for (int r = 0; r < 8 * NT; ++r)
    {
        const int e = e0 + we + r;
        if (e >= a.nElmt) break;
        const double *row = sU + (size_t)(we + r) * pitch;
        double *dst       = a.out + (size_t)e * n;
        for (int k = lane; k < n; k += 32) dst[k] = row[k];
    }
}

// P[p][k][m]: response m of probe p to unit vector k.  afrag[t][kk][i][lane] = A_t[i*8 + lane/4][kk*4 + lane%4]
__global__ void dense_pack_kernel(const double *__restrict__ P, double *__restrict__ afrag, int n, int KS, int MT, int nT,
                                  int dim)
{
    const int total = nT * KS * MT * 32;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x)
    {
        const int lane = idx & 31, i = (idx >> 5) % MT, kk = (idx / (32 * MT)) % KS, t = idx / (32 * MT * KS);
        const int m = i * 8 + (lane >> 2), k = kk * 4 + (lane & 3);
        double v = 0.0;
        if (m < n && k < n)
        {
            const size_t o = (size_t)k * n + m, blk = (size_t)n * n;
            if (t <= dim) v = P[t * blk + o];
            else
            {
                const int q = t - dim - 1; // 3-D: (0,1) (0,2) (1,2); 2-D: (0,1)
                const int pa = (dim == 3 && q == 2) ? 1 : 0, pb = (dim == 3) ? (q == 0 ? 1 : 2) : 1;
                v = P[t * blk + o] - P[(1 + pa) * blk + o] - P[(1 + pb) * blk + o];
            }
        }
        afrag[idx] = v;
    }
}

struct DenseState
{
    int (*fallback)(nekmf_op_s *, const double *const in[3], double *const out[3]) = nullptr;
    void *fallback_state          = nullptr;
    void (*fallback_free)(void *) = nullptr;
    std::string fallback_name;
    double *d_afrag = nullptr;
    int MT = 0, KS = 0, G = 0, nT = 0, pitch = 0;
    size_t smem    = 0;
    bool built     = false;
    bool use_dense = false;
    bool attr_set  = false;
};

template <int MT> static int dense_launch_mt(nekmf_op_s *op, DenseState *st, const double *in, double *out)
{
    using Cfg = DenseCfg<MT>;
    auto kern = dense_helm_kernel<MT>;
    if (!st->attr_set)
    {
        // the attribute belongs to the instantiation, not to this operator: size it for the widest pitch an
        // operator with MT row tiles can have, so a later operator never lowers it
        size_t cap = (size_t)(DSTAGES * Cfg::CHUNK + 7 * Cfg::NE + Cfg::NE * (8 * MT + 4)) * 8;
        if (cap > 227 * 1024) cap = 227 * 1024;
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap));
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        st->attr_set = true;
    }
    DenseArgs a;
    a.in = in; a.out = out;
    a.jac = op->d_jac + op->run_e0; a.df = op->d_df + op->run_e0; a.afrag = st->d_afrag;
    a.dfStride = (size_t)op->nElmt;
    a.nElmt = op->run_ne; a.n = op->nmTot; a.KS = st->KS; a.G = st->G; a.nT = st->nT; a.dim = op->dim;
    a.pitch = st->pitch; a.lambda = op->lambda;
    const int grid = (op->run_ne + Cfg::NE - 1) / Cfg::NE;
    if (grid < 1) return NEKMF_OK;
    kern<<<grid, DW * 32, st->smem, op->run_stream>>>(a);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}

template <int MT> static size_t dense_smem(int pitch)
{
    using Cfg = DenseCfg<MT>;
    return (size_t)(DSTAGES * Cfg::CHUNK + 7 * Cfg::NE + Cfg::NE * pitch) * 8;
}

#define DENSE_MTS(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(11) X(12) X(15) X(18) X(21)

static bool dense_mt_supported(int MT)
{
    switch (MT)
    {
#define X(m) case m:
        DENSE_MTS(X)
#undef X
        return true;
        default: return false;
    }
}
static size_t dense_smem_for(int MT, int pitch)
{
    switch (MT)
    {
#define X(m) case m: return dense_smem<m>(pitch);
        DENSE_MTS(X)
#undef X
        default: return 0;
    }
}

static int dense_launch(nekmf_op_s *op, const double *const in[3], double *const out[3])
{
    DenseState *st = static_cast<DenseState *>(op->kstate);
    if (!st->use_dense)
    {
        void *saved  = op->kstate;
        op->kstate   = st->fallback_state;
        const int rc = st->fallback(op, in, out);
        op->kstate   = saved;
        return rc;
    }
    switch (st->MT)
    {
#define X(m) case m: return dense_launch_mt<m>(op, st, in[0], out[0]);
        DENSE_MTS(X)
#undef X
        default: set_error("dense Helmholtz: no instantiation for %d row tiles", st->MT); return NEKMF_ERR_UNSUPPORTED;
    }
}

// apply the operator's own quadrature-space kernel to unit vectors on the probe geometries, repack as fragments
static int dense_build(nekmf_op_s *op, DenseState *st)
{
    const int n = op->nmTot, dim = op->dim, ndf = op->ndf, nT = st->nT, np = nT - 1;
    const size_t nel = (size_t)np * n, blk = (size_t)n * n;
    std::vector<double> h_in(nel * n, 0.0), h_jac(nel, 1.0), h_df((size_t)ndf * nel, 0.0);
    for (size_t e = 0; e < nel; ++e) h_in[e * n + e % n] = 1.0;
    for (int p = 0; p < np; ++p)
    {
        int da = p, db = -1; // df[0*dim + da] (and db) = 1: first row of df = e_a (+ e_b)
        if (p >= dim)
        {
            const int q = p - dim;
            da = (dim == 3 && q == 2) ? 1 : 0;
            db = (dim == 3) ? (q == 0 ? 1 : 2) : 1;
        }
        for (int j = 0; j < n; ++j)
        {
            h_df[(size_t)da * nel + (size_t)p * n + j] = 1.0;
            if (db >= 0) h_df[(size_t)db * nel + (size_t)p * n + j] = 1.0;
        }
    }
    double *d_in = nullptr, *d_P = nullptr, *d_pjac = nullptr, *d_pdf = nullptr, *d_zero = nullptr;
    auto cleanup = [&]() { cudaFree(d_in); cudaFree(d_P); cudaFree(d_pjac); cudaFree(d_pdf); cudaFree(d_zero); };
#define DB_CUDA(call)                                                                                          \
    do                                                                                                         \
    {                                                                                                          \
        cudaError_t _e = (call);                                                                               \
        if (_e != cudaSuccess)                                                                                 \
        {                                                                                                      \
            set_error("dense Helmholtz setup: %s failed: %s", #call, cudaGetErrorString(_e));                  \
            cleanup();                                                                                         \
            return NEKMF_ERR_CUDA;                                                                             \
        }                                                                                                      \
    } while (0)
    DB_CUDA(cudaMalloc(&d_in, h_in.size() * 8));
    DB_CUDA(cudaMalloc(&d_P, (size_t)nT * blk * 8));
    DB_CUDA(cudaMalloc(&d_pjac, nel * 8));
    DB_CUDA(cudaMalloc(&d_pdf, h_df.size() * 8));
    DB_CUDA(cudaMalloc(&d_zero, h_df.size() * 8));
    DB_CUDA(cudaMemcpy(d_in, h_in.data(), h_in.size() * 8, cudaMemcpyHostToDevice));
    DB_CUDA(cudaMemcpy(d_pjac, h_jac.data(), nel * 8, cudaMemcpyHostToDevice));
    DB_CUDA(cudaMemcpy(d_pdf, h_df.data(), h_df.size() * 8, cudaMemcpyHostToDevice));
    DB_CUDA(cudaMemset(d_zero, 0, h_df.size() * 8));
    if (!st->d_afrag) DB_CUDA(cudaMalloc(&st->d_afrag, (size_t)st->G * st->MT * 32 * 8));

    // borrow the operator for the two probe launches
    double *s_jac = op->d_jac, *s_df = op->d_df;
    const int s_nel = op->nElmt, s_e0 = op->run_e0, s_ne = op->run_ne;
    const double s_lambda = op->lambda;
    cudaStream_t s_stream = op->run_stream;
    void *s_state         = op->kstate;
    op->kstate     = st->fallback_state;
    op->nElmt      = (int)nel;
    op->run_e0     = 0;
    op->run_stream = op->stream;
    op->d_jac      = d_pjac;
    const double *in3[3] = {d_in, nullptr, nullptr};
    double *out3[3]      = {d_P, nullptr, nullptr};
    op->lambda = 1.0; op->d_df = d_zero; op->run_ne = n;
    int rc = st->fallback(op, in3, out3);
    if (rc == NEKMF_OK)
    {
        out3[0]    = d_P + blk;
        op->lambda = 0.0; op->d_df = d_pdf; op->run_ne = (int)nel;
        rc         = st->fallback(op, in3, out3);
    }
    op->kstate = s_state; op->nElmt = s_nel; op->run_e0 = s_e0; op->run_ne = s_ne; op->run_stream = s_stream;
    op->d_jac = s_jac; op->d_df = s_df; op->lambda = s_lambda;
    if (rc != NEKMF_OK) { cleanup(); return rc; }
    const int total = nT * st->KS * st->MT * 32;
    dense_pack_kernel<<<(total + 255) / 256, 256, 0, op->stream>>>(d_P, st->d_afrag, n, st->KS, st->MT, nT, dim);
    ++g_launches;
    DB_CUDA(cudaGetLastError());
    DB_CUDA(cudaStreamSynchronize(op->stream));
#undef DB_CUDA
    cleanup();
    st->built = true;
    return NEKMF_OK;
}

// default policy: where the quadrature-space kernels are FP64 / shared-memory bound;
// NEKMF_DENSE=0 disables the kernel, NEKMF_DENSE=1 takes it wherever an instantiation exists
static bool dense_wanted(const nekmf_op_s *op)
{
    const char *env = getenv("NEKMF_DENSE");
    if (env && env[0] == '0') return false;
    if (env && env[0] == '1') return true;
    // measured (profiles/r01_sweep_dense_helm.jsonl): faster than the quadrature-space kernels at every order
    // it is instantiated for (Tet nm 2..9: 2.0-4.7x, Pyr nm 2..7: 5-40x, Tri nm 3..9: 1.3-2.6x)
    switch (op->shape)
    {
        case NEKMF_TET: return true;
        case NEKMF_PYR: return true;
        case NEKMF_TRI: return op->nm[0] >= 3;
        default: return false;
    }
}

// called from nekmf_op_create after a quadrature-space launcher is installed
void dense_maybe_wrap(nekmf_op_s *op)
{
    if (op->optype != NEKMF_HELMHOLTZ || op->deformed || op->kron) return;
    if (op->shape != NEKMF_TET && op->shape != NEKMF_PYR && op->shape != NEKMF_TRI) return;
    if (op->coordim != op->dim) return;
    const int n = op->nmTot, MT = (n + 7) / 8, KS = (n + 3) / 4;
    if (!dense_mt_supported(MT)) return;
    const int K4 = KS * 4, pitch = (K4 % 8 == 4) ? K4 : K4 + 4;
    const size_t smem = dense_smem_for(MT, pitch);
    if (smem == 0 || smem > 227 * 1024) return;
    DenseState *st     = new DenseState;
    st->MT = MT; st->KS = KS; st->nT = 1 + op->dim * (op->dim + 1) / 2; st->G = st->nT * KS; st->pitch = pitch;
    st->smem           = smem;
    st->fallback       = op->launch;
    st->fallback_state = op->kstate;
    st->fallback_free  = op->kstate_free;
    st->fallback_name  = op->kname;
    op->kstate         = st;
    op->kstate_free    = [](void *p) {
        DenseState *s = static_cast<DenseState *>(p);
        if (s->fallback_state && s->fallback_free) s->fallback_free(s->fallback_state);
        cudaFree(s->d_afrag);
        delete s;
    };
    op->launch = dense_launch;
    op->kron   = 3;
}

// called after set_geom
int dense_geom_changed(nekmf_op_s *op)
{
    if (op->kron != 3) return NEKMF_OK;
    DenseState *st = static_cast<DenseState *>(op->kstate);
    st->use_dense  = false;
    op->kname      = st->fallback_name;
    if (!op->has_jac || !op->has_df || op->nElmt == 0 || !dense_wanted(op)) return NEKMF_OK;
    if (!st->built)
    {
        const int rc = dense_build(op, st);
        if (rc != NEKMF_OK) return rc;
    }
    st->use_dense = true;
    const char *sn[6] = {"Quad", "Tri", "Hex", "Prism", "Pyr", "Tet"};
    char name[112];
    snprintf(name, sizeof(name), "dense_helm_kernel<%s,n=%d,MT=%d>(regular,DMMA m8n8k4)", sn[op->shape], op->nmTot, st->MT);
    op->kname = name;
    return NEKMF_OK;
}

} // namespace nekmf
