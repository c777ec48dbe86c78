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
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <utility>
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
// (kk_major: afrag[kk][t][i][lane], the order the prism kernel walks)
__global__ void dense_pack_kernel(const double *__restrict__ P, double *__restrict__ afrag, int n, int KS, int MT, int nT,
                                  int dim, int kk_major = 0)
{
    const int total = nT * KS * MT * 32;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x)
    {
        const int lane = idx & 31, i = (idx >> 5) % MT, g = idx / (32 * MT);
        const int kk = kk_major ? g / nT : g % KS, t = kk_major ? g % nT : g / KS;
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
        // pyramids: 3-40x over the runtime-sized kernel it was measured against; against the compile-time sized pencil
        // kernel (round 2) it wins at nm 2..6 (0.40-1.41 against 1.24-1.65 ms) and loses at nm = 7 (2.30 against 1.61 ms)
        case NEKMF_PYR: return op->nm[0] <= 6;
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


// =====================================================================================================================
// Extruded prisms.  A prism is triangle (xi_0, xi_2) x segment (xi_1): phi_pqr = T_(p,r)(xi_0, xi_2) S_q(xi_1), the
// collapsed-coordinate factors depend on (xi_0, xi_2) only (Helmholtz.h:1382-1430), so on a regular element
//   H = J [ (lambda Mt + G00 Kt00 + G22 Kt22 + G02 (Kt02 + Kt20)) (x) M1  +  G11 Mt (x) K1  +  G01 (..) + G12 (..) ].
// When the segment direction is orthogonal to the triangle plane (G01 = G12 = 0 for every element: extruded meshes,
// prisms cut from boxes), the generalised eigen-decomposition X^T M1 X = I, X^T K1 X = diag(mu) of the SEGMENT pair
// (nm x nm, host, once) decouples the nm segment modes:  with  u~ = (I (x) X^-1) u,
//   v~_(q~) = J [ (lambda + G11 mu_q~) Mt + G00 Kt00 + G22 Kt22 + G02 Kt02s ] u~_(q~),     out = (I (x) X^-T) v~
// i.e. nm independent TRIANGLE Helmholtz problems per element -- the DMMA GEMM above with n = nm(nm+1)/2, nT = 4 and
// one column per (element, q~).  The X^-1 / X^-T transforms are fused into the tile load / store of the kernel.  The
// triangle matrices are measured like the dense ones (probes on the modes with q = 0, divided by M1[0][0]).
struct PrismArgs
{
    const double *in;
    double *out;
    const double *jac, *df, *afrag; // afrag[kk][t][i][lane]
    const double *tab;              // [nm*nm] Tin = X^-1 (row q~, column q) | [nm] mu
    const int *itab;                // [n] offset of mode (i, q=0) in the prism ordering | [n] stride between q for that i
    size_t dfStride;
    int nElmt;
    double lambda;
};

constexpr int PSTAGES = 3; // A-fragment ring of the prism kernel (four CTAs per SM at nm = 7)

template <int NM> struct PrismCfg
{
    static constexpr int N = NM * (NM + 1) / 2, MT = (N + 7) / 8, KS = (N + 3) / 4, K4 = KS * 4;
    static constexpr int PITCH = (K4 % 8 == 4) ? K4 : K4 + 4;
    static constexpr int NP = NM * N, EW = 16 / NM, USED = EW * NM; // elements / live columns per warp
    static constexpr int KPS   = 2;                                  // kk iterations per ring stage
    static constexpr int STEP  = 4 * MT * 32;                        // doubles per kk (four terms)
    static constexpr int CHUNK = KPS * STEP;
    static constexpr int NS    = (KS + KPS - 1) / KPS;
    static constexpr int TABD  = (NM * NM + NM + 1) & ~1;            // Tin | mu
    static constexpr int TABI  = (2 * N + 1) & ~1;                   // off | len (ints)
    static constexpr size_t SMEM = (size_t)(PSTAGES * CHUNK + DW * 16 * PITCH + TABD + TABI / 2 + DW * EW * NP) * 8;
};

template <int NM> __global__ void __launch_bounds__(DW * 32, 2) prism_helm_kernel(const __grid_constant__ PrismArgs a)
{
    using Cfg = PrismCfg<NM>;
    constexpr int N = Cfg::N, MT = Cfg::MT, KS = Cfg::KS, PITCH = Cfg::PITCH, NP = Cfg::NP, EW = Cfg::EW, USED = Cfg::USED;
    constexpr int CHUNK = Cfg::CHUNK, STEP = Cfg::STEP, NS = Cfg::NS, KPS = Cfg::KPS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sA   = reinterpret_cast<double *>(smem_raw); // [PSTAGES][CHUNK]
    double *sU   = sA + PSTAGES * CHUNK;                 // [DW*16][PITCH]
    double *sTin = sU + DW * 16 * PITCH;                 // [NM*NM] | mu[NM]
    int *sOff    = reinterpret_cast<int *>(sTin + Cfg::TABD);
    int *sLen    = sOff + N;
    double *sRaw = sTin + Cfg::TABD + Cfg::TABI / 2;     // [DW][EW*NP] the warp's elements in the reference ordering
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int we  = warp * 16;                           // first column of the warp
    const int eW0 = (blockIdx.x * DW + warp) * EW;       // first real element of the warp
    const double *sMu = sTin + NM * NM;

    auto issue = [&](int st) {
        if (st < NS)
        {
            const int k0 = st * KPS, k1 = (k0 + KPS < KS) ? k0 + KPS : KS;
            const int units   = (k1 - k0) * (STEP / 2);
            const double *src = a.afrag + (size_t)k0 * STEP;
            double *dst       = sA + (st % PSTAGES) * CHUNK;
            for (int i = tid; i < units; i += DW * 32) cp_async16(dst + 2 * i, src + 2 * i);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int st = 0; st < PSTAGES - 1; ++st) issue(st);
    for (int i = tid; i < NM * NM + NM; i += DW * 32) sTin[i] = __ldg(a.tab + i);
    for (int i = tid; i < 2 * N; i += DW * 32) sOff[i] = __ldg(a.itab + i);

    int nLive = a.nElmt - eW0;
    nLive     = nLive < 0 ? 0 : (nLive > EW ? EW : nLive);
    double *raw = sRaw + (size_t)warp * EW * NP;
    {
        const double *src = a.in + (size_t)eW0 * NP;
        for (int idx = lane; idx < nLive * NP; idx += 32) raw[idx] = __ldg(src + idx);
    }
    __syncthreads();

    // per-lane scales of its two B-fragment columns: c_t of (element, q~)
    double cT[4][2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
    {
        const int col = j * 8 + (lane >> 2), el = col / NM, qt = col - el * NM;
        cT[0][j] = cT[1][j] = cT[2][j] = cT[3][j] = 0.0;
        if (col < USED && el < nLive)
        {
            const int e    = eW0 + el;
            const double J = __ldg(a.jac + e);
            double d[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) d[q] = __ldg(a.df + q * a.dfStride + e);
            const double g11 = d[1] * d[1] + d[4] * d[4] + d[7] * d[7];
            cT[0][j] = J * (a.lambda + g11 * sMu[qt]);
            cT[1][j] = J * (d[0] * d[0] + d[3] * d[3] + d[6] * d[6]);
            cT[2][j] = J * (d[2] * d[2] + d[5] * d[5] + d[8] * d[8]);
            cT[3][j] = J * (d[0] * d[2] + d[3] * d[5] + d[6] * d[8]);
        }
    }
    // u~[(e, q~)][i] = sum_q Tin[q~][q] u[e][off_i + q len_i]; lane = i
    for (int k = lane; k < PITCH; k += 32)
    {
        double x[EW][NM];
        const bool kin = k < N;
        const int o = kin ? sOff[k] : 0, l = kin ? sLen[k] : 0;
#pragma unroll
        for (int el = 0; el < EW; ++el)
#pragma unroll
            for (int q = 0; q < NM; ++q) x[el][q] = (kin && el < nLive) ? raw[el * NP + o + q * l] : 0.0;
#pragma unroll
        for (int col = 0; col < 16; ++col)
        {
            const int el = col / NM, qt = col - el * NM;
            double v = 0.0;
            if (col < USED)
            {
#pragma unroll
                for (int q = 0; q < NM; ++q) v = fma(sTin[qt * NM + q], x[el < EW ? el : 0][q], v);
            }
            sU[(size_t)(we + col) * PITCH + k] = v;
        }
    }
    __syncwarp();

    double acc[MT][2][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const double *uB[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) uB[j] = sU + (size_t)(we + j * 8 + (lane >> 2)) * PITCH + (lane & 3);

#pragma unroll
    for (int kk = 0; kk < KS; ++kk)
    {
        if (kk % KPS == 0)
        {
            cp_async_wait<PSTAGES - 2>(); // this thread's copies of the stage have landed
            __syncthreads();              // everyone's have, and everyone is done with the previous stage
            issue(kk / KPS + PSTAGES - 1);
        }
        const double *st = sA + ((kk / KPS) % PSTAGES) * CHUNK + (kk % KPS) * STEP + lane;
        const double u0 = uB[0][kk * 4], u1 = uB[1][kk * 4];
#pragma unroll
        for (int t = 0; t < 4; ++t)
        {
            double av[MT];
#pragma unroll
            for (int i = 0; i < MT; ++i) av[i] = st[(t * MT + i) * 32];
            const double b0 = u0 * cT[t][0], b1 = u1 * cT[t][1];
#pragma unroll
            for (int i = 0; i < MT; ++i)
            {
                dmma884(acc[i][0][0], acc[i][0][1], av[i], b0);
                dmma884(acc[i][1][0], acc[i][1][1], av[i], b1);
            }
        }
    }
    cp_async_wait<0>();

    // v~ -> the warp's rows of sU, then out[e][off_i + q len_i] = sum_q~ Tin[q~][q] v~[(e, q~)][i]; lane = i
    __syncwarp();
#pragma unroll
    for (int i = 0; i < MT; ++i)
    {
        const int m = i * 8 + (lane >> 2);
        if (m < N)
        {
#pragma unroll
            for (int j = 0; j < 2; ++j)
            {
                double *row = sU + (size_t)(we + j * 8 + 2 * (lane & 3)) * PITCH + m;
                row[0]      = acc[i][j][0];
                row[PITCH]  = acc[i][j][1];
            }
        }
    }
    __syncwarp();
    for (int k = lane; k < N; k += 32)
    {
        const int o = sOff[k], l = sLen[k];
#pragma unroll
        for (int el = 0; el < EW; ++el)
        {
            if (el < nLive)
            {
                double y[NM];
#pragma unroll
                for (int qt = 0; qt < NM; ++qt) y[qt] = sU[(size_t)(we + el * NM + qt) * PITCH + k];
#pragma unroll
                for (int q = 0; q < NM; ++q)
                {
                    double v = 0.0;
#pragma unroll
                    for (int qt = 0; qt < NM; ++qt) v = fma(sTin[qt * NM + q], y[qt], v);
                    raw[el * NP + o + q * l] = v;
                }
            }
        }
    }
    __syncwarp();
    {
        double *dst = a.out + (size_t)eW0 * NP;
        for (int idx = lane; idx < nLive * NP; idx += 32) dst[idx] = raw[idx];
    }
}

// Ptri[t][k][m] = Pfull[(t n + k)][off_m] / M1[0][0]
__global__ void prism_extract_kernel(const double *__restrict__ Pfull, double *__restrict__ Ptri, const int *__restrict__ off,
                                     int n, int NP, double inv_m00)
{
    const int total = 4 * n * n;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x)
    {
        const int m = idx % n, tk = idx / n;
        Ptri[idx] = Pfull[(size_t)tk * NP + off[m]] * inv_m00;
    }
}

struct PrismState
{
    int (*fallback)(nekmf_op_s *, const double *const in[3], double *const out[3]) = nullptr;
    void *fallback_state          = nullptr;
    void (*fallback_free)(void *) = nullptr;
    std::string fallback_name;
    double *d_afrag = nullptr, *d_tab = nullptr;
    int *d_itab     = nullptr;
    int n = 0, MT = 0, KS = 0, G = 0, nmq = 0, EW = 0;
    bool built     = false;
    bool use_fast  = false;
    bool attr_set  = false;
    // general (non-extruded) regular prisms: eight-term kernel
    double *d_afrag8 = nullptr, *d_tab8 = nullptr;
    bool built8 = false, use_general = false, attr_set8 = false;
    // general regular prisms at nm = 5..7: fused quadrature-space kernel on tensor tiles (prism_helm_dmma.cu)
    void *fused    = nullptr;
    bool use_fused = false;
};

template <int NM> static int prism_launch_nm(nekmf_op_s *op, PrismState *st, const double *in, double *out)
{
    using Cfg = PrismCfg<NM>;
    auto kern = prism_helm_kernel<NM>;
    if (!st->attr_set)
    {
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        st->attr_set = true;
    }
    PrismArgs a;
    a.in = in; a.out = out;
    a.jac = op->d_jac + op->run_e0; a.df = op->d_df + op->run_e0; a.afrag = st->d_afrag;
    a.tab = st->d_tab; a.itab = st->d_itab;
    a.dfStride = (size_t)op->nElmt;
    a.nElmt = op->run_ne; a.lambda = op->lambda;
    const int perCta = DW * Cfg::EW;
    const int grid   = (op->run_ne + perCta - 1) / perCta;
    if (grid < 1) return NEKMF_OK;
    kern<<<grid, DW * 32, Cfg::SMEM, op->run_stream>>>(a);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// General regular prisms (G01, G12 != 0).  The mixed terms are C_a (x) S^T + C_a^T (x) S with S[q][q'] = int S'_q S_q'
// (segment tables) and C_a the mixed triangle matrices; in the segment eigen-basis they act on two more transformed
// copies of the input, w1 = (I (x) St^T X^-1) u and w2 = (I (x) St X^-1) u with St = X^T S X:
//   v~ = [diagonal part as above] + (G01 C0 + G12 C2) w1 + (G01 C0^T + G12 C2^T) w2
// i.e. eight triangle-matrix terms on three B-operand tiles (tests/test_oracle.py::
// test_prism_general_kronecker_formulation_against_oracle restates exactly this in numpy against the oracle).
template <int NM> struct PrismGenCfg
{
    using C = PrismCfg<NM>;
    static constexpr int N = C::N, MT = C::MT, KS = C::KS, PITCH = C::PITCH, NP = C::NP, EW = C::EW, USED = C::USED;
    static constexpr int NTERM = 8;
    static constexpr int STEP  = NTERM * MT * 32;            // doubles per kk = one ring stage
    static constexpr int TILE  = DW * 16 * PITCH;
    static constexpr int TABD  = (3 * NM * NM + NM + 1) & ~1; // Tin | mu | T1 | T2
    static constexpr int TABI  = C::TABI;
    static constexpr size_t SMEM = (size_t)(PSTAGES * STEP + 3 * TILE + TABD + TABI / 2 + DW * EW * NP) * 8;
};

template <int NM> __global__ void __launch_bounds__(DW * 32, 2) prism_gen_kernel(const __grid_constant__ PrismArgs a)
{
    using Cfg = PrismGenCfg<NM>;
    constexpr int N = Cfg::N, MT = Cfg::MT, KS = Cfg::KS, PITCH = Cfg::PITCH, NP = Cfg::NP, EW = Cfg::EW, USED = Cfg::USED;
    constexpr int STEP = Cfg::STEP, TILE = Cfg::TILE;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sA   = reinterpret_cast<double *>(smem_raw); // [PSTAGES][STEP]
    double *sU0  = sA + PSTAGES * STEP;                  // u~   [DW*16][PITCH]
    double *sU1  = sU0 + TILE;                           // w1
    double *sU2  = sU1 + TILE;                           // w2
    double *sTin = sU2 + TILE;                           // Tin | mu | T1 | T2
    int *sOff    = reinterpret_cast<int *>(sTin + Cfg::TABD);
    int *sLen    = sOff + N;
    double *sRaw = sTin + Cfg::TABD + Cfg::TABI / 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int we  = warp * 16;
    const int eW0 = (blockIdx.x * DW + warp) * EW;
    const double *sMu = sTin + NM * NM, *sT1 = sMu + NM, *sT2 = sT1 + NM * NM;

    auto issue = [&](int st) {
        if (st < KS)
        {
            const double *src = a.afrag + (size_t)st * STEP;
            double *dst       = sA + (st % PSTAGES) * STEP;
            for (int i = tid; i < STEP / 2; i += DW * 32) cp_async16(dst + 2 * i, src + 2 * i);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int st = 0; st < PSTAGES - 1; ++st) issue(st);
    for (int i = tid; i < 3 * NM * NM + NM; i += DW * 32) sTin[i] = __ldg(a.tab + i);
    for (int i = tid; i < 2 * N; i += DW * 32) sOff[i] = __ldg(a.itab + i);

    int nLive = a.nElmt - eW0;
    nLive     = nLive < 0 ? 0 : (nLive > EW ? EW : nLive);
    double *raw = sRaw + (size_t)warp * EW * NP;
    {
        const double *src = a.in + (size_t)eW0 * NP;
        for (int idx = lane; idx < nLive * NP; idx += 32) raw[idx] = __ldg(src + idx);
    }
    __syncthreads();

    double cT[8][2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
    {
        const int col = j * 8 + (lane >> 2), el = col / NM, qt = col - el * NM;
#pragma unroll
        for (int t = 0; t < 8; ++t) cT[t][j] = 0.0;
        if (col < USED && el < nLive)
        {
            const int e    = eW0 + el;
            const double J = __ldg(a.jac + e);
            double d[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) d[q] = __ldg(a.df + q * a.dfStride + e);
            const double g11 = d[1] * d[1] + d[4] * d[4] + d[7] * d[7];
            const double g01 = J * (d[0] * d[1] + d[3] * d[4] + d[6] * d[7]);
            const double g12 = J * (d[1] * d[2] + d[4] * d[5] + d[7] * d[8]);
            cT[0][j] = J * (a.lambda + g11 * sMu[qt]);
            cT[1][j] = J * (d[0] * d[0] + d[3] * d[3] + d[6] * d[6]);
            cT[2][j] = J * (d[2] * d[2] + d[5] * d[5] + d[8] * d[8]);
            cT[3][j] = J * (d[0] * d[2] + d[3] * d[5] + d[6] * d[8]);
            cT[4][j] = g01; cT[5][j] = g12; cT[6][j] = g01; cT[7][j] = g12;
        }
    }
    for (int k = lane; k < PITCH; k += 32)
    {
        double x[EW][NM];
        const bool kin = k < N;
        const int o = kin ? sOff[k] : 0, l = kin ? sLen[k] : 0;
#pragma unroll
        for (int el = 0; el < EW; ++el)
#pragma unroll
            for (int q = 0; q < NM; ++q) x[el][q] = (kin && el < nLive) ? raw[el * NP + o + q * l] : 0.0;
#pragma unroll
        for (int col = 0; col < 16; ++col)
        {
            const int el = col / NM, qt = col - el * NM;
            double v0 = 0.0, v1 = 0.0, v2 = 0.0;
            if (col < USED)
            {
#pragma unroll
                for (int q = 0; q < NM; ++q)
                {
                    const double xv = x[el < EW ? el : 0][q];
                    v0 = fma(sTin[qt * NM + q], xv, v0);
                    v1 = fma(sT1[qt * NM + q], xv, v1);
                    v2 = fma(sT2[qt * NM + q], xv, v2);
                }
            }
            sU0[(size_t)(we + col) * PITCH + k] = v0;
            sU1[(size_t)(we + col) * PITCH + k] = v1;
            sU2[(size_t)(we + col) * PITCH + k] = v2;
        }
    }
    __syncwarp();

    double acc[MT][2][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const size_t fo[2] = {(size_t)(we + (lane >> 2)) * PITCH + (lane & 3), (size_t)(we + 8 + (lane >> 2)) * PITCH + (lane & 3)};

#pragma unroll
    for (int kk = 0; kk < KS; ++kk)
    {
        cp_async_wait<PSTAGES - 2>();
        __syncthreads();
        issue(kk + PSTAGES - 1);
        const double *st = sA + (kk % PSTAGES) * STEP + lane;
        double u[3][2];
#pragma unroll
        for (int j = 0; j < 2; ++j)
        {
            u[0][j] = sU0[fo[j] + kk * 4];
            u[1][j] = sU1[fo[j] + kk * 4];
            u[2][j] = sU2[fo[j] + kk * 4];
        }
#pragma unroll
        for (int t = 0; t < 8; ++t)
        {
            const int src = t < 4 ? 0 : (t < 6 ? 1 : 2);
            double av[MT];
#pragma unroll
            for (int i = 0; i < MT; ++i) av[i] = st[(t * MT + i) * 32];
            const double b0 = u[src][0] * cT[t][0], b1 = u[src][1] * cT[t][1];
#pragma unroll
            for (int i = 0; i < MT; ++i)
            {
                dmma884(acc[i][0][0], acc[i][0][1], av[i], b0);
                dmma884(acc[i][1][0], acc[i][1][1], av[i], b1);
            }
        }
    }
    cp_async_wait<0>();

    __syncwarp();
#pragma unroll
    for (int i = 0; i < MT; ++i)
    {
        const int m = i * 8 + (lane >> 2);
        if (m < N)
        {
#pragma unroll
            for (int j = 0; j < 2; ++j)
            {
                double *row = sU0 + (size_t)(we + j * 8 + 2 * (lane & 3)) * PITCH + m;
                row[0]      = acc[i][j][0];
                row[PITCH]  = acc[i][j][1];
            }
        }
    }
    __syncwarp();
    for (int k = lane; k < N; k += 32)
    {
        const int o = sOff[k], l = sLen[k];
#pragma unroll
        for (int el = 0; el < EW; ++el)
        {
            if (el < nLive)
            {
                double y[NM];
#pragma unroll
                for (int qt = 0; qt < NM; ++qt) y[qt] = sU0[(size_t)(we + el * NM + qt) * PITCH + k];
#pragma unroll
                for (int q = 0; q < NM; ++q)
                {
                    double v = 0.0;
#pragma unroll
                    for (int qt = 0; qt < NM; ++qt) v = fma(sTin[qt * NM + q], y[qt], v);
                    raw[el * NP + o + q * l] = v;
                }
            }
        }
    }
    __syncwarp();
    {
        double *dst = a.out + (size_t)eW0 * NP;
        for (int idx = lane; idx < nLive * NP; idx += 32) dst[idx] = raw[idx];
    }
}

// Pt[t][k][m], t = Mt, Kt00, Kt22, Kt02s, C0, C2, C0^T, C2^T from the probe responses
// Pfull[g][v][NP]: g = M, K00, K22, P02, K11, P01, P12; v = unit vector on mode (i' = v, q' = 0) for v < n, (v - n, 1) above
__global__ void prism_extract8_kernel(const double *__restrict__ Pfull, double *__restrict__ Pt, const int *__restrict__ off,
                                      int n, int NP, double inv_m00, double inv_s00, double inv_s10)
{
    const int total = 8 * n * n;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x)
    {
        const int m = idx % n, k = (idx / n) % n, t = idx / (n * n);
        auto resp = [&](int g, int v, int mo) { return Pfull[((size_t)g * 2 * n + v) * NP + off[mo]]; };
        double val;
        if (t == 0) val = resp(0, k, m) * inv_m00;
        else if (t == 1) val = resp(1, k, m) * inv_m00;
        else if (t == 2) val = resp(2, k, m) * inv_m00;
        else if (t == 3) val = (resp(3, k, m) - resp(1, k, m) - resp(2, k, m)) * inv_m00;
        else
        {
            const int g = (t & 1) ? 6 : 5, gd = (t & 1) ? 2 : 1; // C2: P12 - K11 - K22;  C0: P01 - K00 - K11
            const int kc = t < 6 ? k : m, mr = t < 6 ? m : k;    // transposed matrices swap the roles
            const double a00 = resp(g, kc, mr) - resp(gd, kc, mr) - resp(4, kc, mr);
            const double a01 = resp(g, n + kc, mr) - resp(gd, n + kc, mr) - resp(4, n + kc, mr);
            val = 0.5 * (a00 * inv_s00 + a01 * inv_s10);
        }
        Pt[idx] = val;
    }
}

template <int NM> static int prism_gen_launch_nm(nekmf_op_s *op, PrismState *st, const double *in, double *out)
{
    using Cfg = PrismGenCfg<NM>;
    auto kern = prism_gen_kernel<NM>;
    if (!st->attr_set8)
    {
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        st->attr_set8 = true;
    }
    PrismArgs a;
    a.in = in; a.out = out;
    a.jac = op->d_jac + op->run_e0; a.df = op->d_df + op->run_e0; a.afrag = st->d_afrag8;
    a.tab = st->d_tab8; a.itab = st->d_itab;
    a.dfStride = (size_t)op->nElmt;
    a.nElmt = op->run_ne; a.lambda = op->lambda;
    const int perCta = DW * Cfg::EW;
    const int grid   = (op->run_ne + perCta - 1) / perCta;
    if (grid < 1) return NEKMF_OK;
    kern<<<grid, DW * 32, Cfg::SMEM, op->run_stream>>>(a);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}

static int prism_launch(nekmf_op_s *op, const double *const in[3], double *const out[3])
{
    PrismState *st = static_cast<PrismState *>(op->kstate);
    if (!st->use_fast)
    {
        void *saved  = op->kstate;
        op->kstate   = st->fallback_state;
        const int rc = st->fallback(op, in, out);
        op->kstate   = saved;
        return rc;
    }
    if (st->use_fused) return prism_helm_fused_launch(st->fused, op, in[0], out[0]);
    if (st->use_general)
    {
        switch (st->nmq)
        {
            case 2: return prism_gen_launch_nm<2>(op, st, in[0], out[0]);
            case 3: return prism_gen_launch_nm<3>(op, st, in[0], out[0]);
            case 4: return prism_gen_launch_nm<4>(op, st, in[0], out[0]);
            case 5: return prism_gen_launch_nm<5>(op, st, in[0], out[0]);
            case 6: return prism_gen_launch_nm<6>(op, st, in[0], out[0]);
            case 7: return prism_gen_launch_nm<7>(op, st, in[0], out[0]);
            case 8: return prism_gen_launch_nm<8>(op, st, in[0], out[0]);
            default: set_error("prism Helmholtz: no instantiation for nm = %d", st->nmq); return NEKMF_ERR_UNSUPPORTED;
        }
    }
    switch (st->nmq)
    {
        case 2: return prism_launch_nm<2>(op, st, in[0], out[0]);
        case 3: return prism_launch_nm<3>(op, st, in[0], out[0]);
        case 4: return prism_launch_nm<4>(op, st, in[0], out[0]);
        case 5: return prism_launch_nm<5>(op, st, in[0], out[0]);
        case 6: return prism_launch_nm<6>(op, st, in[0], out[0]);
        case 7: return prism_launch_nm<7>(op, st, in[0], out[0]);
        case 8: return prism_launch_nm<8>(op, st, in[0], out[0]);
        default: set_error("prism Helmholtz: no instantiation for nm = %d", st->nmq); return NEKMF_ERR_UNSUPPORTED;
    }
}

// generalised symmetric eigenproblem K x = mu M x of the segment matrices (nm <= 8): Cholesky + cyclic Jacobi.
// Returns Tin = X^-1 = Q^T L^T (row q~, column q) and mu, with X^T M X = I, X^T K X = diag(mu).
static bool seg_eigen(int nm, const std::vector<double> &M, const std::vector<double> &K, std::vector<double> &Tin,
                      std::vector<double> &mu)
{
    std::vector<double> L(nm * nm, 0.0), C(nm * nm, 0.0), Q(nm * nm, 0.0), Y(nm * nm, 0.0);
    for (int i = 0; i < nm; ++i)
        for (int j = 0; j <= i; ++j)
        {
            double s = M[i * nm + j];
            for (int k = 0; k < j; ++k) s -= L[i * nm + k] * L[j * nm + k];
            if (i == j)
            {
                if (s <= 0.0) return false;
                L[i * nm + i] = sqrt(s);
            }
            else L[i * nm + j] = s / L[j * nm + j];
        }
    // Y = L^-1 K (forward substitution on columns), C = Y L^-T = L^-1 K L^-T
    for (int c = 0; c < nm; ++c)
        for (int i = 0; i < nm; ++i)
        {
            double s = K[i * nm + c];
            for (int k = 0; k < i; ++k) s -= L[i * nm + k] * Y[k * nm + c];
            Y[i * nm + c] = s / L[i * nm + i];
        }
    for (int r = 0; r < nm; ++r)
        for (int i = 0; i < nm; ++i)
        {
            double s = Y[r * nm + i];
            for (int k = 0; k < i; ++k) s -= L[i * nm + k] * C[r * nm + k];
            C[r * nm + i] = s / L[i * nm + i];
        }
    for (int i = 0; i < nm; ++i)
        for (int j = 0; j < i; ++j) C[i * nm + j] = C[j * nm + i] = 0.5 * (C[i * nm + j] + C[j * nm + i]);
    for (int i = 0; i < nm; ++i) Q[i * nm + i] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep)
    {
        double off = 0.0, dia = 0.0;
        for (int i = 0; i < nm; ++i)
            for (int j = 0; j < nm; ++j) (i == j ? dia : off) += C[i * nm + j] * C[i * nm + j];
        if (off <= 1e-32 * dia) break;
        for (int p = 0; p < nm; ++p)
            for (int q = p + 1; q < nm; ++q)
            {
                const double apq = C[p * nm + q];
                if (apq == 0.0) continue;
                const double th = (C[q * nm + q] - C[p * nm + p]) / (2.0 * apq);
                const double tt = (th >= 0.0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
                const double cs = 1.0 / sqrt(tt * tt + 1.0), sn = tt * cs;
                for (int k = 0; k < nm; ++k)
                {
                    const double akp = C[k * nm + p], akq = C[k * nm + q];
                    C[k * nm + p] = cs * akp - sn * akq;
                    C[k * nm + q] = sn * akp + cs * akq;
                }
                for (int k = 0; k < nm; ++k)
                {
                    const double apk = C[p * nm + k], aqk = C[q * nm + k];
                    C[p * nm + k] = cs * apk - sn * aqk;
                    C[q * nm + k] = sn * apk + cs * aqk;
                }
                for (int k = 0; k < nm; ++k)
                {
                    const double qkp = Q[k * nm + p], qkq = Q[k * nm + q];
                    Q[k * nm + p] = cs * qkp - sn * qkq;
                    Q[k * nm + q] = sn * qkp + cs * qkq;
                }
            }
    }
    Tin.assign(nm * nm, 0.0);
    mu.assign(nm, 0.0);
    for (int a = 0; a < nm; ++a)
    {
        mu[a] = C[a * nm + a];
        for (int q = 0; q < nm; ++q)
        {
            double s = 0.0; // (Q^T L^T)[a][q] = sum_k Q[k][a] L[q][k]
            for (int k = 0; k < nm; ++k) s += Q[k * nm + a] * L[q * nm + k];
            Tin[a * nm + q] = s;
        }
    }
    return true;
}

static int prism_build(nekmf_op_s *op, PrismState *st)
{
    const int nm = st->nmq, n = st->n, NP = op->nmTot, nq1 = op->nq[1];
    // segment matrices from the direction-1 tables (Operator.hpp:244-258 weights, no collapsed factor in xi_1)
    std::vector<double> M1(nm * nm), K1(nm * nm), Tin, mu;
    const double *B = op->b[1].data(), *dB = op->db[1].data(), *w = op->ws[1].data();
    for (int a = 0; a < nm; ++a)
        for (int c = 0; c < nm; ++c)
        {
            double m = 0.0, k = 0.0;
            for (int j = 0; j < nq1; ++j)
            {
                m += B[a * nq1 + j] * w[j] * B[c * nq1 + j];
                k += dB[a * nq1 + j] * w[j] * dB[c * nq1 + j];
            }
            M1[a * nm + c] = m;
            K1[a * nm + c] = k;
        }
    if (!seg_eigen(nm, M1, K1, Tin, mu)) { set_error("prism Helmholtz: segment mass matrix not positive definite"); return NEKMF_ERR_ARG; }
    std::vector<double> tab(Tin);
    tab.insert(tab.end(), mu.begin(), mu.end());
    std::vector<int> itab(2 * n);
    for (int p = 0, i = 0; p < nm; ++p)
        for (int r = 0; r < nm - p; ++r, ++i)
        {
            const int tri0 = p * nm - (p * (p - 1)) / 2;
            itab[i]     = nm * tri0 + r; // mode (p, q = 0, r): [p][q][r], nm - p values of r per (p, q)
            itab[n + i] = nm - p;
        }
    // probes: 4 geometries x n unit vectors on the q = 0 modes
    const size_t nel = (size_t)3 * n;
    std::vector<double> h_in((size_t)4 * n * NP, 0.0), h_jac(nel, 1.0), h_df((size_t)9 * nel, 0.0);
    for (int t = 0; t < 4; ++t)
        for (int k = 0; k < n; ++k) h_in[((size_t)t * n + k) * NP + itab[k]] = 1.0;
    for (int k = 0; k < n; ++k)
    {
        h_df[(size_t)0 * nel + 0 * n + k] = 1.0;                                       // G00
        h_df[(size_t)2 * nel + 1 * n + k] = 1.0;                                       // G22
        h_df[(size_t)0 * nel + 2 * n + k] = 1.0; h_df[(size_t)2 * nel + 2 * n + k] = 1.0; // G00 + G22 + 2 G02
    }
    double *d_in = nullptr, *d_P = nullptr, *d_Pt = nullptr, *d_pjac = nullptr, *d_pdf = nullptr, *d_zero = nullptr;
    auto cleanup = [&]() { cudaFree(d_in); cudaFree(d_P); cudaFree(d_Pt); cudaFree(d_pjac); cudaFree(d_pdf); cudaFree(d_zero); };
#define PB_CUDA(call)                                                                                          \
    do                                                                                                         \
    {                                                                                                          \
        cudaError_t _e = (call);                                                                               \
        if (_e != cudaSuccess)                                                                                 \
        {                                                                                                      \
            set_error("prism Helmholtz setup: %s failed: %s", #call, cudaGetErrorString(_e));                  \
            cleanup();                                                                                         \
            return NEKMF_ERR_CUDA;                                                                             \
        }                                                                                                      \
    } while (0)
    PB_CUDA(cudaMalloc(&d_in, h_in.size() * 8));
    PB_CUDA(cudaMalloc(&d_P, h_in.size() * 8));
    PB_CUDA(cudaMalloc(&d_Pt, (size_t)4 * n * n * 8));
    PB_CUDA(cudaMalloc(&d_pjac, nel * 8));
    PB_CUDA(cudaMalloc(&d_pdf, h_df.size() * 8));
    PB_CUDA(cudaMalloc(&d_zero, h_df.size() * 8));
    PB_CUDA(cudaMemcpy(d_in, h_in.data(), h_in.size() * 8, cudaMemcpyHostToDevice));
    PB_CUDA(cudaMemcpy(d_pjac, h_jac.data(), nel * 8, cudaMemcpyHostToDevice));
    PB_CUDA(cudaMemcpy(d_pdf, h_df.data(), h_df.size() * 8, cudaMemcpyHostToDevice));
    PB_CUDA(cudaMemset(d_zero, 0, h_df.size() * 8));
    if (!st->d_afrag) PB_CUDA(cudaMalloc(&st->d_afrag, (size_t)st->G * st->MT * 32 * 8));
    if (!st->d_tab) PB_CUDA(cudaMalloc(&st->d_tab, tab.size() * 8));
    if (!st->d_itab) PB_CUDA(cudaMalloc(&st->d_itab, itab.size() * 4));
    PB_CUDA(cudaMemcpy(st->d_tab, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice));
    PB_CUDA(cudaMemcpy(st->d_itab, itab.data(), itab.size() * 4, cudaMemcpyHostToDevice));

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
        in3[0]     = d_in + (size_t)n * NP;
        out3[0]    = d_P + (size_t)n * NP;
        op->lambda = 0.0; op->d_df = d_pdf; op->run_ne = (int)nel;
        rc         = st->fallback(op, in3, out3);
    }
    op->kstate = s_state; op->nElmt = s_nel; op->run_e0 = s_e0; op->run_ne = s_ne; op->run_stream = s_stream;
    op->d_jac = s_jac; op->d_df = s_df; op->lambda = s_lambda;
    if (rc != NEKMF_OK) { cleanup(); return rc; }
    prism_extract_kernel<<<(4 * n * n + 255) / 256, 256, 0, op->stream>>>(d_P, d_Pt, st->d_itab, n, NP, 1.0 / M1[0]);
    ++g_launches;
    PB_CUDA(cudaGetLastError());
    const int total = 4 * st->KS * st->MT * 32;
    dense_pack_kernel<<<(total + 255) / 256, 256, 0, op->stream>>>(d_Pt, st->d_afrag, n, st->KS, st->MT, 4, 2, 1);
    ++g_launches;
    PB_CUDA(cudaGetLastError());
    PB_CUDA(cudaStreamSynchronize(op->stream));
#undef PB_CUDA
    cleanup();
    st->built = true;
    return NEKMF_OK;
}

static bool small_inverse(int n, std::vector<double> A, std::vector<double> &inv)
{
    inv.assign(n * n, 0.0);
    for (int i = 0; i < n; ++i) inv[i * n + i] = 1.0;
    for (int c = 0; c < n; ++c)
    {
        int piv = c;
        for (int r = c + 1; r < n; ++r)
            if (fabs(A[r * n + c]) > fabs(A[piv * n + c])) piv = r;
        if (A[piv * n + c] == 0.0) return false;
        if (piv != c)
            for (int k = 0; k < n; ++k)
            {
                std::swap(A[piv * n + k], A[c * n + k]);
                std::swap(inv[piv * n + k], inv[c * n + k]);
            }
        const double d = 1.0 / A[c * n + c];
        for (int k = 0; k < n; ++k) { A[c * n + k] *= d; inv[c * n + k] *= d; }
        for (int r = 0; r < n; ++r)
        {
            if (r == c) continue;
            const double f = A[r * n + c];
            if (f == 0.0) continue;
            for (int k = 0; k < n; ++k) { A[r * n + k] -= f * A[c * n + k]; inv[r * n + k] -= f * inv[c * n + k]; }
        }
    }
    return true;
}

// tables and matrices of the eight-term kernel (general regular prisms)
static int prism_build_general(nekmf_op_s *op, PrismState *st)
{
    const int nm = st->nmq, n = st->n, NP = op->nmTot, nq1 = op->nq[1];
    std::vector<double> M1(nm * nm), K1(nm * nm), S(nm * nm), Tin, mu, X;
    const double *B = op->b[1].data(), *dB = op->db[1].data(), *w = op->ws[1].data();
    for (int a = 0; a < nm; ++a)
        for (int c = 0; c < nm; ++c)
        {
            double m = 0.0, k = 0.0, sv = 0.0;
            for (int j = 0; j < nq1; ++j)
            {
                m += B[a * nq1 + j] * w[j] * B[c * nq1 + j];
                k += dB[a * nq1 + j] * w[j] * dB[c * nq1 + j];
                sv += dB[a * nq1 + j] * w[j] * B[c * nq1 + j];
            }
            M1[a * nm + c] = m; K1[a * nm + c] = k; S[a * nm + c] = sv;
        }
    if (!seg_eigen(nm, M1, K1, Tin, mu) || !small_inverse(nm, Tin, X) || S[0] == 0.0 || S[nm] == 0.0)
    {
        set_error("prism Helmholtz: segment eigen-decomposition failed");
        return NEKMF_ERR_ARG;
    }
    // St = X^T S X;  T1 = St^T Tin;  T2 = St Tin
    std::vector<double> SX(nm * nm, 0.0), St(nm * nm, 0.0), T1(nm * nm, 0.0), T2(nm * nm, 0.0);
    for (int i = 0; i < nm; ++i)
        for (int j = 0; j < nm; ++j)
            for (int k = 0; k < nm; ++k) SX[i * nm + j] += S[i * nm + k] * X[k * nm + j];
    for (int i = 0; i < nm; ++i)
        for (int j = 0; j < nm; ++j)
            for (int k = 0; k < nm; ++k) St[i * nm + j] += X[k * nm + i] * SX[k * nm + j];
    for (int i = 0; i < nm; ++i)
        for (int j = 0; j < nm; ++j)
            for (int k = 0; k < nm; ++k)
            {
                T1[i * nm + j] += St[k * nm + i] * Tin[k * nm + j];
                T2[i * nm + j] += St[i * nm + k] * Tin[k * nm + j];
            }
    std::vector<double> tab(Tin);
    tab.insert(tab.end(), mu.begin(), mu.end());
    tab.insert(tab.end(), T1.begin(), T1.end());
    tab.insert(tab.end(), T2.begin(), T2.end());
    std::vector<int> itab(2 * n);
    for (int p = 0, i = 0; p < nm; ++p)
        for (int r = 0; r < nm - p; ++r, ++i)
        {
            itab[i]     = nm * (p * nm - (p * (p - 1)) / 2) + r;
            itab[n + i] = nm - p;
        }
    // probes: 7 geometries (M | K00 K22 P02 K11 P01 P12) x 2n unit vectors (modes (i', 0) then (i', 1))
    const int nv = 2 * n;
    const size_t nel = (size_t)6 * nv;
    std::vector<double> h_in((size_t)7 * nv * NP, 0.0), h_jac(nel, 1.0), h_df((size_t)9 * nel, 0.0);
    for (int g = 0; g < 7; ++g)
        for (int v = 0; v < nv; ++v)
            h_in[((size_t)g * nv + v) * NP + (v < n ? itab[v] : itab[v - n] + itab[n + v - n])] = 1.0;
    const int rows[6][3] = {{1, 0, 0}, {0, 0, 1}, {1, 0, 1}, {0, 1, 0}, {1, 1, 0}, {0, 1, 1}};
    for (int g = 0; g < 6; ++g)
        for (int d = 0; d < 3; ++d)
            if (rows[g][d])
                for (int v = 0; v < nv; ++v) h_df[(size_t)d * nel + (size_t)g * nv + v] = 1.0;
    double *d_in = nullptr, *d_P = nullptr, *d_Pt = nullptr, *d_pjac = nullptr, *d_pdf = nullptr, *d_zero = nullptr;
    auto cleanup = [&]() { cudaFree(d_in); cudaFree(d_P); cudaFree(d_Pt); cudaFree(d_pjac); cudaFree(d_pdf); cudaFree(d_zero); };
#define PG_CUDA(call)                                                                                          \
    do                                                                                                         \
    {                                                                                                          \
        cudaError_t _e = (call);                                                                               \
        if (_e != cudaSuccess)                                                                                 \
        {                                                                                                      \
            set_error("prism Helmholtz setup: %s failed: %s", #call, cudaGetErrorString(_e));                  \
            cleanup();                                                                                         \
            return NEKMF_ERR_CUDA;                                                                             \
        }                                                                                                      \
    } while (0)
    PG_CUDA(cudaMalloc(&d_in, h_in.size() * 8));
    PG_CUDA(cudaMalloc(&d_P, h_in.size() * 8));
    PG_CUDA(cudaMalloc(&d_Pt, (size_t)8 * n * n * 8));
    PG_CUDA(cudaMalloc(&d_pjac, nel * 8));
    PG_CUDA(cudaMalloc(&d_pdf, h_df.size() * 8));
    PG_CUDA(cudaMalloc(&d_zero, h_df.size() * 8));
    PG_CUDA(cudaMemcpy(d_in, h_in.data(), h_in.size() * 8, cudaMemcpyHostToDevice));
    PG_CUDA(cudaMemcpy(d_pjac, h_jac.data(), nel * 8, cudaMemcpyHostToDevice));
    PG_CUDA(cudaMemcpy(d_pdf, h_df.data(), h_df.size() * 8, cudaMemcpyHostToDevice));
    PG_CUDA(cudaMemset(d_zero, 0, h_df.size() * 8));
    if (!st->d_afrag8) PG_CUDA(cudaMalloc(&st->d_afrag8, (size_t)8 * st->KS * st->MT * 32 * 8));
    if (!st->d_tab8) PG_CUDA(cudaMalloc(&st->d_tab8, tab.size() * 8));
    if (!st->d_itab) PG_CUDA(cudaMalloc(&st->d_itab, itab.size() * 4));
    PG_CUDA(cudaMemcpy(st->d_tab8, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice));
    PG_CUDA(cudaMemcpy(st->d_itab, itab.data(), itab.size() * 4, cudaMemcpyHostToDevice));

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
    op->lambda = 1.0; op->d_df = d_zero; op->run_ne = nv;
    int rc = st->fallback(op, in3, out3);
    if (rc == NEKMF_OK)
    {
        in3[0]     = d_in + (size_t)nv * NP;
        out3[0]    = d_P + (size_t)nv * NP;
        op->lambda = 0.0; op->d_df = d_pdf; op->run_ne = (int)nel;
        rc         = st->fallback(op, in3, out3);
    }
    op->kstate = s_state; op->nElmt = s_nel; op->run_e0 = s_e0; op->run_ne = s_ne; op->run_stream = s_stream;
    op->d_jac = s_jac; op->d_df = s_df; op->lambda = s_lambda;
    if (rc != NEKMF_OK) { cleanup(); return rc; }
    prism_extract8_kernel<<<(8 * n * n + 255) / 256, 256, 0, op->stream>>>(d_P, d_Pt, st->d_itab, n, NP, 1.0 / M1[0], 1.0 / S[0],
                                                                           1.0 / S[nm]);
    ++g_launches;
    PG_CUDA(cudaGetLastError());
    const int total = 8 * st->KS * st->MT * 32;
    dense_pack_kernel<<<(total + 255) / 256, 256, 0, op->stream>>>(d_Pt, st->d_afrag8, n, st->KS, st->MT, 8, 99, 1);
    ++g_launches;
    PG_CUDA(cudaGetLastError());
    PG_CUDA(cudaStreamSynchronize(op->stream));
#undef PG_CUDA
    cleanup();
    st->built8 = true;
    return NEKMF_OK;
}

void prism_maybe_wrap(nekmf_op_s *op)
{
    if (op->shape != NEKMF_PRISM || op->optype != NEKMF_HELMHOLTZ || op->deformed || op->kron) return;
    const int nm = op->nm[0];
    if (op->nm[1] != nm || op->nm[2] != nm || nm < 2 || nm > 8) return;
    // default quadrature only (the combinations the GPU parity tests of these kernels cover); other quadratures keep
    // the runtime-sized kernel
    if (op->nq[0] != nm + 1 || op->nq[1] != nm + 1 || op->nq[2] != nm) return;
    const int n = nm * (nm + 1) / 2;
    if (op->nmTot != nm * n || op->coordim != 3) return;
    const int MT = (n + 7) / 8, KS = (n + 3) / 4;
    if (MT > 5) return;
    PrismState *st     = new PrismState;
    st->n = n; st->MT = MT; st->KS = KS; st->G = 4 * KS; st->nmq = nm; st->EW = 16 / nm;
    st->fallback       = op->launch;
    st->fallback_state = op->kstate;
    st->fallback_free  = op->kstate_free;
    st->fallback_name  = op->kname;
    op->kstate         = st;
    op->kstate_free    = [](void *p) {
        PrismState *s = static_cast<PrismState *>(p);
        if (s->fallback_state && s->fallback_free) s->fallback_free(s->fallback_state);
        cudaFree(s->d_afrag);
        cudaFree(s->d_tab);
        cudaFree(s->d_itab);
        cudaFree(s->d_afrag8);
        cudaFree(s->d_tab8);
        if (s->fused) prism_helm_fused_free(s->fused, s->nmq);
        delete s;
    };
    op->launch = prism_launch;
    op->kron   = 4;
}

int prism_geom_changed(nekmf_op_s *op)
{
    if (op->kron != 4) return NEKMF_OK;
    PrismState *st = static_cast<PrismState *>(op->kstate);
    st->use_fast    = false;
    st->use_general = false;
    st->use_fused   = false;
    op->kname       = st->fallback_name;
    const char *env = getenv("NEKMF_DENSE");
    if (env && env[0] == '0') return NEKMF_OK;
    if (!op->has_jac || !op->has_df || op->nElmt == 0) return NEKMF_OK;
    // extruded prisms only: G01 = G12 = 0 for every element
    std::vector<double> df((size_t)9 * op->nElmt);
    NEKMF_CUDA(cudaMemcpy(df.data(), op->d_df, df.size() * 8, cudaMemcpyDeviceToHost));
    const size_t N = op->nElmt;
    for (size_t e = 0; e < N; ++e)
    {
        double g[3][3];
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)
                g[a][b] = df[(0 * 3 + a) * N + e] * df[(0 * 3 + b) * N + e] + df[(1 * 3 + a) * N + e] * df[(1 * 3 + b) * N + e] +
                          df[(2 * 3 + a) * N + e] * df[(2 * 3 + b) * N + e];
        const double tol = 1e-14 * (g[0][0] + g[1][1] + g[2][2]);
        if (fabs(g[0][1]) > tol || fabs(g[1][2]) > tol)
        {
            // general regular prisms: the eight-term kernel where it was measured faster than the quadrature-space
            // kernel (nm 3..5: 1.5-2.2x; equal at nm = 7, 8); NEKMF_PRISM_GENERAL=1 takes it at every order, =0 never
            // nm = 6, 7 (and nm = 5 with NEKMF_PRISM_FUSED=1): the fused quadrature-space kernel on tensor tiles
            // (prism_helm_dmma.cu); NEKMF_PRISM_FUSED=0 never
            // (an explicit NEKMF_PRISM_GENERAL selects between the eight-term and the quadrature-space kernel and leaves this
            // one out unless NEKMF_PRISM_FUSED=1)
            const char *ef   = getenv("NEKMF_PRISM_FUSED");
            const bool fon   = ef && ef[0] == '1';
            const bool fauto = !ef && !getenv("NEKMF_PRISM_GENERAL") && (st->nmq == 6 || st->nmq == 7);
            if ((fon && st->nmq >= 5 && st->nmq <= 7) || fauto)
            {
                if (!st->fused) st->fused = prism_helm_fused_create(op);
                if (st->fused)
                {
                    st->use_fast = st->use_fused = true;
                    char fname[112];
                    snprintf(fname, sizeof(fname), "prism_helm_dmma_kernel<nm=%d>(regular,general,fused quadrature space,DMMA m8n8k4)", st->nmq);
                    op->kname = fname;
                    return NEKMF_OK;
                }
            }
            const char *eg = getenv("NEKMF_PRISM_GENERAL");
            if (eg && eg[0] == '0') return NEKMF_OK;
            if (!(eg && eg[0] == '1') && (st->nmq < 3 || st->nmq > 5)) return NEKMF_OK;
            if (!st->built8)
            {
                const int rc = prism_build_general(op, st);
                if (rc != NEKMF_OK) return rc;
            }
            st->use_fast = st->use_general = true;
            char gname[112];
            snprintf(gname, sizeof(gname), "prism_gen_kernel<nm=%d,MT=%d>(regular,general,DMMA m8n8k4)", st->nmq, st->MT);
            op->kname = gname;
            return NEKMF_OK;
        }
    }
    if (!st->built)
    {
        const int rc = prism_build(op, st);
        if (rc != NEKMF_OK) return rc;
    }
    st->use_fast = true;
    char name[112];
    snprintf(name, sizeof(name), "prism_helm_kernel<nm=%d,MT=%d>(regular,extruded,DMMA m8n8k4)", st->nmq, st->MT);
    op->kname = name;
    return NEKMF_OK;
}

} // namespace nekmf
