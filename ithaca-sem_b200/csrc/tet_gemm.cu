// tet_gemm.cu -- BwdTrans (and, second half of the file, IProductWRTBase) on tetrahedra at nm = 5..7 (default quadrature nq = (nm+1, nm, nm)) with the two collapsed
// contractions PRE-COMBINED into per-p tables and evaluated as FP64 tensor-core GEMMs over eight elements.
//
// Reference semantics: MatrixFreeOps/BwdTransKernels.hpp:374-484 (BwdTransTetKernel, CORRECT = true for the modified
// basis); results differ from shape_kernels.cuh / tet_dmma.cu by summation order only.
//
// phi_pqr = A_p(xi_0) B_pq(xi_1) C_pqr(xi_2).  For a fixed p the xi_1 and xi_2 contractions are ONE linear map from the
// modes (q, r) of that p to the plane values
//     g_p[k][j] = sum_{(q,r)} T_p[(k,j)][(q,r)] c[pqr],      T_p[(k,j)][(q,r)] = C_pqr(k) B_pq(j)
// (the reference forms it in two sum-factorised steps; at nm = 7 the combined table has 49 x 84 entries, 48 KB padded).
// With the columns of the right-hand side = EIGHT ELEMENTS this is a GEMM with exactly the m8n8k4 shape:
//     rows    j (one 8-row tile per quadrature plane k; nq_1 <= 7 rows in use)
//     K       the modes of p, in k-steps of four (28, 21 + 7, 15, 10, 6, 3, 1 at nm = 7: 25 k-steps)
//     columns the eight elements of the warp's batch
// A fragments come from the shared-memory copy of the tables (rows padded to a stride = 4 or 12 mod 16 doubles: two
// wavefronts per load, the minimum for 32 x 8 bytes), B fragments are the coefficients c[e = g][mode = 4s + t], read once
// per batch into registers (the element stride of 84 doubles = 4 mod 16 is conflict-free as it is).  The result fragment
// gives lane (g, t) the values g_p[k][j = g] of elements 2t and 2t + 1 for every p -- the whole p-line of its two points --
// so the last contraction, out[k][j][i] = sum_p A_p(i) g_p[k][j], is plain DFMA in the owning lane and the eight values
// of an i-line leave as 16-byte stores.
// The corrections of the modified basis are linear in the coefficients too and are FOLDED INTO THE TABLES on the host:
// the top-vertex mode (0,0,1) gets an extra term in T_0 and, together with the bottom-vertex and singular-edge modes
// (0,1,r), seven extra columns in T_1 (their coefficients are gathered through the per-p index list like any other mode).
#include "op_internal.h"
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace nekmf
{

template <int NM> struct TetGemmDims
{
    static constexpr int NQ0 = NM + 1, NQ1 = NM, NQ2 = NM, NPAIR = NM * (NM + 1) / 2, NMT = NM * (NM + 1) * (NM + 2) / 6;
    static constexpr int NQT = NQ0 * NQ1 * NQ2;
    __host__ __device__ static constexpr int own(int p) { return (NM - p) * (NM - p + 1) / 2; }      // modes (q, r) of p
    __host__ __device__ static constexpr int kcols(int p) { return own(p) + (p == 1 ? NM : 0); }      // + correction columns of p = 1
    __host__ __device__ static constexpr int kp(int p) { return (kcols(p) + 3) & ~3; }                // padded to whole k-steps
    __host__ __device__ static constexpr int ks(int p) { return kp(p) / 4; }
    __host__ __device__ static constexpr int stride(int p) { return (kp(p) % 16 == 4 || kp(p) % 16 == 12) ? kp(p) : kp(p) + 4; }
    __host__ __device__ static constexpr int toff(int p) { int o = 0; for (int a = 0; a < p; ++a) o += NQ2 * 8 * stride(a); return o; }
    __host__ __device__ static constexpr int ioff(int p) { int o = 0; for (int a = 0; a < p; ++a) o += kp(a); return o; }
    __host__ __device__ static constexpr int soff(int p) { int o = 0; for (int a = 0; a < p; ++a) o += ks(a); return o; } // k-step index
    static constexpr int TOT_T = toff(NM), TOT_I = ioff(NM), TOT_KS = soff(NM);
    static constexpr int EB = 8;                              // elements per batch = tile columns
    static constexpr int BUF = EB * NMT;                      // doubles per input buffer (a multiple of 16 bytes)
    static constexpr int PER_WARP = 2 * BUF + 2;              // double buffer + two mbarriers
    static constexpr int TAB = TOT_T + ((TOT_I + 1) / 2) * 1; // tables: doubles + ints packed behind them
    static constexpr int TABD = (TAB + 1) & ~1;
    static constexpr int W_FIT = (224 * 1024 / 8 - TABD) / PER_WARP;
    static constexpr int WARPS = W_FIT >= 12 ? 12 : (W_FIT >= 8 ? 8 : 4);
    static constexpr int T     = WARPS * 32;
    static constexpr size_t SMEM = (size_t)(TABD + WARPS * PER_WARP) * 8 + 16;
};

template <int NM> struct TetGemmTab
{
    double b0[NM * (NM + 1)]; // A_p(i), [p][i]
};

struct TetGemmArgs
{
    const double *in;
    double *out;
    const double *tabT; // [TOT_T] combined tables, layout T_p[k][j (8)][stride(p)]
    const int *tabI;    // [TOT_I] per-p coefficient index lists
    int nElmt;
    int in_aligned, out_aligned; // 16-byte aligned
};

__device__ __forceinline__ void tg_mma(double &c0, double &c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int I> struct tg_int { static constexpr int value = I; };
template <int I, int N, class F> __device__ __forceinline__ void tg_static_for(F &&f)
{
    if constexpr (I < N)
    {
        f(tg_int<I>{});
        tg_static_for<I + 1, N>(f);
    }
}

template <int NM>
__global__ void __launch_bounds__(TetGemmDims<NM>::T, 1)
    tet_bwd_gemm_kernel(const __grid_constant__ TetGemmTab<NM> tab, const __grid_constant__ TetGemmArgs args)
{
    using Dm = TetGemmDims<NM>;
    constexpr int NQ0 = Dm::NQ0, NQ1 = Dm::NQ1, NQ2 = Dm::NQ2, NMT = Dm::NMT, NQT = Dm::NQT, EB = Dm::EB, BUF = Dm::BUF;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = lane & 3, g = lane >> 2;
    double *sT    = reinterpret_cast<double *>(smem_raw);
    int *sI       = reinterpret_cast<int *>(sT + Dm::TOT_T);
    double *wbase = sT + Dm::TABD + (size_t)warp * Dm::PER_WARP;
    uint64_t *bar = reinterpret_cast<uint64_t *>(wbase + 2 * BUF);

    for (int i = threadIdx.x; i < Dm::TOT_T; i += Dm::T) sT[i] = __ldg(args.tabT + i);
    for (int i = threadIdx.x; i < Dm::TOT_I; i += Dm::T) sI[i] = __ldg(args.tabI + i);
    if (lane == 0)
    {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int nBat = (args.nElmt + EB - 1) / EB;
    const int GW = gridDim.x * Dm::WARPS, gw = blockIdx.x * Dm::WARPS + warp;
    auto bat_ne = [&](int b) { int r = args.nElmt - EB * b; return r < EB ? r : EB; };
    auto tma_ok = [&](int b) { return args.in_aligned && bat_ne(b) == EB; };
    auto issue  = [&](int b, int slot) { // lane 0
        if (!tma_ok(b)) return;
        mbar_expect_tx(bar + slot, (uint32_t)(BUF * 8));
        tma_load_1d(wbase + slot * BUF, args.in + (size_t)b * BUF, (uint32_t)(BUF * 8), bar + slot);
    };

    uint32_t phase[2] = {0, 0};
    int slot          = 0;
    if (lane == 0 && gw < nBat) issue(gw, 0);

    for (int b = gw; b < nBat; b += GW, slot ^= 1)
    {
        const int ne = bat_ne(b);
        double *sIn  = wbase + slot * BUF;
        if (lane == 0 && b + GW < nBat) issue(b + GW, slot ^ 1); // the other buffer was consumed one trip ago
        if (tma_ok(b))
        {
            mbar_wait(bar + slot, phase[slot]);
            phase[slot] ^= 1;
        }
        else
        {
            // ragged last batch or 8-byte aligned caller arrays: plain loads by the warp, missing elements zero
            const double *src = args.in + (size_t)b * BUF;
            for (int i = lane; i < BUF; i += 32) sIn[i] = i < ne * NMT ? __ldg(src + i) : 0.0;
        }
        __syncwarp();

        // B fragments: c[element g][mode 4s + t], every k-step of every p
        double bop[Dm::TOT_KS];
        tg_static_for<0, NM>([&](auto pc) {
            constexpr int p = decltype(pc)::value;
#pragma unroll
            for (int s = 0; s < Dm::ks(p); ++s) bop[Dm::soff(p) + s] = sIn[g * NMT + sI[Dm::ioff(p) + 4 * s + t]];
        });

        const size_t e0 = (size_t)b * EB + 2 * t; // the lane's two result columns = elements e0, e0 + 1
        const bool st0 = g < NQ1 && 2 * t < ne, st1 = g < NQ1 && 2 * t + 1 < ne;
#pragma unroll 1
        for (int k = 0; k < NQ2; ++k)
        {
            // g_p[k][j = g] of elements e0, e0 + 1 for every p
            double C[NM][2];
            tg_static_for<0, NM>([&](auto pc) {
                constexpr int p = decltype(pc)::value;
                C[p][0] = C[p][1] = 0.0;
                const double *tp = sT + Dm::toff(p) + (k * 8 + g) * Dm::stride(p) + t;
#pragma unroll
                for (int s = 0; s < Dm::ks(p); ++s) tg_mma(C[p][0], C[p][1], tp[4 * s], bop[Dm::soff(p) + s]);
            });
            // out[k][j = g][i] = sum_p A_p(i) g_p: the i-lines of the lane's two points
            double o0[NQ0], o1[NQ0];
#pragma unroll
            for (int i = 0; i < NQ0; ++i)
            {
                o0[i] = tab.b0[i] * C[0][0];
                o1[i] = tab.b0[i] * C[0][1];
#pragma unroll
                for (int p = 1; p < NM; ++p)
                {
                    o0[i] = fma(tab.b0[p * NQ0 + i], C[p][0], o0[i]);
                    o1[i] = fma(tab.b0[p * NQ0 + i], C[p][1], o1[i]);
                }
            }
            double *d0 = args.out + e0 * NQT + k * (NQ0 * NQ1) + g * NQ0, *d1 = d0 + NQT;
            if ((NQ0 % 2 == 0) && args.out_aligned)
            {
                if (st0)
                {
#pragma unroll
                    for (int i = 0; i < NQ0; i += 2) *reinterpret_cast<double2 *>(d0 + i) = make_double2(o0[i], o0[i + 1]);
                }
                if (st1)
                {
#pragma unroll
                    for (int i = 0; i < NQ0; i += 2) *reinterpret_cast<double2 *>(d1 + i) = make_double2(o1[i], o1[i + 1]);
                }
            }
            else if ((NQ0 % 2 == 1) && (NQT % 2 == 0) && ((NQ0 * NQ1) % 2 == 0) && args.out_aligned)
            {
                // odd i-lines (nm even): the line starts on a 16-byte boundary for even j and 8 bytes past one for odd j;
                // one scalar store at the odd end, 16-byte stores for the rest
                const int h = g & 1; // 1: o[0] alone, pairs from 1; 0: pairs from 0, o[NQ0-1] alone
                if (st0)
                {
                    d0[h ? 0 : NQ0 - 1] = h ? o0[0] : o0[NQ0 - 1];
#pragma unroll
                    for (int i = 0; i + 1 < NQ0; i += 2)
                        *reinterpret_cast<double2 *>(d0 + i + h) = h ? make_double2(o0[i + 1], o0[i + 2]) : make_double2(o0[i], o0[i + 1]);
                }
                if (st1)
                {
                    d1[h ? 0 : NQ0 - 1] = h ? o1[0] : o1[NQ0 - 1];
#pragma unroll
                    for (int i = 0; i + 1 < NQ0; i += 2)
                        *reinterpret_cast<double2 *>(d1 + i + h) = h ? make_double2(o1[i + 1], o1[i + 2]) : make_double2(o1[i], o1[i + 1]);
                }
            }
            else
            {
                if (st0)
                {
#pragma unroll
                    for (int i = 0; i < NQ0; ++i) d0[i] = o0[i];
                }
                if (st1)
                {
#pragma unroll
                    for (int i = 0; i < NQ0; ++i) d1[i] = o1[i];
                }
            }
            __syncwarp(); // reconverge before the next plane's tiles
        }
        __syncwarp(); // every lane is done with this buffer before lane 0 refills it
    }
}

template <int NM> struct TetGemmState
{
    TetGemmTab<NM> tab;
    double *d_T = nullptr;
    int *d_I    = nullptr;
    void *fallback_state          = nullptr;
    void (*fallback_free)(void *) = nullptr;
    int bps                       = 0;
};

template <int NM> static int tet_gemm_launch(nekmf_op_s *op, const double *const in[3], double *const out[3])
{
    auto *st  = static_cast<TetGemmState<NM> *>(op->kstate);
    using Dm  = TetGemmDims<NM>;
    auto kern = tet_bwd_gemm_kernel<NM>;
    if (st->bps == 0)
    {
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Dm::SMEM));
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int nb = 0;
        NEKMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, Dm::T, Dm::SMEM));
        if (nb < 1) { set_error("tet GEMM kernel <%d> does not fit on an SM", NM); return NEKMF_ERR_CUDA; }
        st->bps = nb;
    }
    TetGemmArgs a;
    a.in = in[0]; a.out = out[0]; a.tabT = st->d_T; a.tabI = st->d_I; a.nElmt = op->run_ne;
    a.in_aligned  = (((uintptr_t)in[0]) & 15) == 0;
    a.out_aligned = (((uintptr_t)out[0]) & 15) == 0;
    const int nBat = (op->run_ne + Dm::EB - 1) / Dm::EB;
    int grid       = st->bps * NUM_SMS;
    const int need = (nBat + Dm::WARPS - 1) / Dm::WARPS;
    if (grid > need) grid = need;
    if (grid < 1) return NEKMF_OK;
    kern<<<grid, Dm::T, Dm::SMEM, op->run_stream>>>(st->tab, a);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}

template <int NM> static bool tet_gemm_wrap(nekmf_op_s *op)
{
    using Dm = TetGemmDims<NM>;
    constexpr int NQ0 = Dm::NQ0, NQ1 = Dm::NQ1, NQ2 = Dm::NQ2, NPAIR = Dm::NPAIR, NMT = Dm::NMT;
    if ((int)op->b[0].size() != NM * NQ0 || (int)op->b[1].size() != NPAIR * NQ1 || (int)op->b[2].size() != NMT * NQ2) return false;
    const double *b1 = op->b[1].data(), *b2 = op->b[2].data();
    std::vector<double> T(Dm::TOT_T, 0.0);
    std::vector<int> I(Dm::TOT_I, 0);
    // own modes of every p, in the reference's order (p outer, q, r inner)
    int mode = 0, pair = 0;
    for (int p = 0; p < NM; ++p)
    {
        int col         = 0;
        const int st    = Dm::stride(p);
        double *tp      = T.data() + Dm::toff(p);
        int *ip         = I.data() + Dm::ioff(p);
        auto add_column = [&](int c, int coef_index, int b2row, auto b1of) {
            ip[c] = coef_index;
            for (int k = 0; k < NQ2; ++k)
                for (int j = 0; j < NQ1; ++j) tp[(k * 8 + j) * st + c] += b2[b2row * NQ2 + k] * b1of(j);
        };
        for (int q = 0; q < NM - p; ++q, ++pair)
            for (int r = 0; r < NM - p - q; ++r, ++mode, ++col)
            {
                const int pr = pair;
                add_column(col, mode, mode, [&](int j) { return b1[pr * NQ1 + j]; });
            }
        if (p == 0 && NM > 1)
        {
            // top vertex (mode 1 = (0,0,1)), p = 0 part: b2 row 1, B_01
            add_column(1, 1, 1, [&](int j) { return b1[NQ1 + j]; });
        }
        if (p == 1)
        {
            // top vertex, p = 1 part: b2 row 1, B_00 + B_01
            add_column(col++, 1, 1, [&](int j) { return b1[j] + b1[NQ1 + j]; });
            // bottom vertex (mode nm = (0,1,0)): b2 row 0, B_01
            add_column(col++, NM, 0, [&](int j) { return b1[NQ1 + j]; });
            // singular edge (modes nm + r = (0,1,r)): b2 row r + 1, B_01
            for (int r = 1; r < NM - 1; ++r) add_column(col++, NM + r, r + 1, [&](int j) { return b1[NQ1 + j]; });
        }
    }
    auto *stt = new TetGemmState<NM>;
    memcpy(stt->tab.b0, op->b[0].data(), sizeof(stt->tab.b0));
    if (cudaMalloc(&stt->d_T, T.size() * 8) != cudaSuccess || cudaMalloc(&stt->d_I, I.size() * 4) != cudaSuccess ||
        cudaMemcpy(stt->d_T, T.data(), T.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(stt->d_I, I.data(), I.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess)
    {
        cudaFree(stt->d_T);
        cudaFree(stt->d_I);
        delete stt;
        return false;
    }
    stt->fallback_state = op->kstate;
    stt->fallback_free  = op->kstate_free;
    op->kstate          = stt;
    op->kstate_free     = [](void *p) {
        auto *s = static_cast<TetGemmState<NM> *>(p);
        if (s->fallback_state && s->fallback_free) s->fallback_free(s->fallback_state);
        cudaFree(s->d_T);
        cudaFree(s->d_I);
        delete s;
    };
    op->launch = tet_gemm_launch<NM>;
    char name[128];
    snprintf(name, sizeof(name), "tet_bwd_gemm_kernel<nm=%d>(DMMA m8n8k4 over 8 elements, collapsed contractions pre-combined)", NM);
    op->kname = name;
    return true;
}

// ------------------------------------------------------------------------------------------------ IProductWRTBase
// The transposed chain (IProductKernels.hpp:600-761): with the weights and the Jacobian applied to the input,
//     s_p[(k,j)][e] = sum_i A_p(i) (w J F)[e][k][j][i]          in the lane that loaded the two i-lines (j = t, 4 + t) of
//                                                                 plane k of ELEMENT g -- straight from global memory,
//                                                                 64 contiguous bytes per line
//     out[m][e]     = sum_{(k,j)} T_p[(k,j)][m] s_p[(k,j)][e]    GEMM: rows = the modes m of p (8-row tiles), K = (k, j)
//                                                                 (two k-steps per plane), columns = eight elements
// A fragments = the transposed tables from shared memory, B fragments = the s values as they are computed, so the plane
// loop needs no staging of the input at all; the accumulators of all 15 row tiles (nm = 7) stay in registers over the
// planes.  The corrections are again table rows: an extra term in row 1 of T_0 and seven extra rows of T_1 whose results
// are ADDED to the modes (0,0,1), (0,1,r) in the staging block before it leaves with coalesced stores.
template <int NM> struct TetIpDims
{
    static constexpr int NQ0 = NM + 1, NQ1 = NM, NQ2 = NM, NPAIR = NM * (NM + 1) / 2, NMT = NM * (NM + 1) * (NM + 2) / 6;
    static constexpr int NQT = NQ0 * NQ1 * NQ2;
    __host__ __device__ static constexpr int own(int p) { return (NM - p) * (NM - p + 1) / 2; }
    __host__ __device__ static constexpr int start(int p) { int o = 0; for (int a = 0; a < p; ++a) o += own(a); return o; }
    __host__ __device__ static constexpr int rows(int p) { return own(p) + (p == 1 ? NM : 0); }
    __host__ __device__ static constexpr int tiles(int p) { return (rows(p) + 7) / 8; }
    __host__ __device__ static constexpr int tioff(int p) { int o = 0; for (int a = 0; a < p; ++a) o += tiles(a); return o; }
    static constexpr int TOT_TILES = tioff(NM);
    static constexpr int KK = NQ2 * 8;                                        // contracted index (k, j padded to 8)
    static constexpr int RS = (KK % 16 == 4 || KK % 16 == 12) ? KK : KK + 4;  // row stride: two wavefronts per A load
    static constexpr int TOT_T = TOT_TILES * 8 * RS;
    static constexpr int EB = 8, STG = EB * NMT;
    static constexpr int TABD = (TOT_T + 1) & ~1;
#ifndef TET_IP_WARPS7
#define TET_IP_WARPS7 8
#endif
    // nm >= 6: up to 30 tile accumulators + the current and the prefetched i-lines per lane; 12 warps (168 registers)
    // spill, 8 do not
    static constexpr int WARPS = NM >= 6 ? TET_IP_WARPS7 : 12, T = WARPS * 32;
    static constexpr size_t SMEM = (size_t)(TABD + WARPS * STG) * 8 + 16;
};

template <int NM> struct TetIpTab
{
    double b0[NM * (NM + 1)]; // A_p(i), [p][i]
    double w0[NM + 1], w1[NM], w2[NM];
};

struct TetIpArgs
{
    const double *in;
    double *out;
    const double *jac;  // [nElmt] | [nElmt][nqTot]
    const double *tabT; // [TOT_T] transposed combined tables, layout [tile][row (8)][RS]
    int nElmt;
    int in_aligned, out_aligned; // 16-byte aligned (in: the input and, deformed, the Jacobian)
};

template <int NM, bool DEF>
__global__ void __launch_bounds__(TetIpDims<NM>::T, 1)
    tet_ip_gemm_kernel(const __grid_constant__ TetIpTab<NM> tab, const __grid_constant__ TetIpArgs args)
{
    using Dm = TetIpDims<NM>;
    constexpr int NQ0 = Dm::NQ0, NQ1 = Dm::NQ1, NQ2 = Dm::NQ2, NMT = Dm::NMT, NQT = Dm::NQT, EB = Dm::EB, RS = Dm::RS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = lane & 3, g = lane >> 2;
    double *sT   = reinterpret_cast<double *>(smem_raw);
    double *sStg = sT + Dm::TABD + (size_t)warp * Dm::STG;

    for (int i = threadIdx.x; i < Dm::TOT_T; i += Dm::T) sT[i] = __ldg(args.tabT + i);
    __syncthreads();

    const int nBat = (args.nElmt + EB - 1) / EB;
    const int GW = gridDim.x * Dm::WARPS, gw = blockIdx.x * Dm::WARPS + warp;
    const int j0 = t, j1 = 4 + t; // the lane's two i-lines of a plane
    const bool vj0 = j0 < NQ1, vj1 = j1 < NQ1;
    const double wj0 = vj0 ? tab.w1[j0] : 0.0, wj1 = vj1 ? tab.w1[j1 < NQ1 ? j1 : 0] : 0.0;
    const bool vec = (NQ0 % 2 == 0) && args.in_aligned;

    for (int b = gw; b < nBat; b += GW)
    {
        const int ne = args.nElmt - EB * b < EB ? args.nElmt - EB * b : EB;
        const bool ve = g < ne;
        const size_t el = (size_t)b * EB + (ve ? g : 0);
        const double *F = args.in + el * NQT;
        const double *J = DEF ? args.jac + el * NQT : nullptr;
        const double jreg = DEF ? 1.0 : __ldg(args.jac + el);
        double C[Dm::TOT_TILES][2];
#pragma unroll
        for (int n = 0; n < Dm::TOT_TILES; ++n) C[n][0] = C[n][1] = 0.0;

        // the lane's two i-lines of plane k, read one plane ahead (registers): the global-load latency hides behind the
        // tiles of the previous plane
        auto load_lines = [&](const double *src, int k, double (&x0)[NQ0], double (&x1)[NQ0]) {
            const int o0 = k * (NQ0 * NQ1) + (vj0 ? j0 : 0) * NQ0, o1 = k * (NQ0 * NQ1) + (vj1 ? j1 : 0) * NQ0;
            if (vec)
            {
#pragma unroll
                for (int i = 0; i < NQ0; i += 2)
                {
                    const double2 a = __ldg(reinterpret_cast<const double2 *>(src + o0 + i)), c = __ldg(reinterpret_cast<const double2 *>(src + o1 + i));
                    x0[i] = a.x; x0[i + 1] = a.y; x1[i] = c.x; x1[i + 1] = c.y;
                }
            }
            else
            {
#pragma unroll
                for (int i = 0; i < NQ0; ++i) { x0[i] = __ldg(src + o0 + i); x1[i] = __ldg(src + o1 + i); }
            }
        };
        // PREFETCH off (deformed, nm = 7): the second set of line registers would spill -- measured 0.47 with, 0.58 without
        constexpr bool PREFETCH = !(DEF && NM >= 7);
        double n0[NQ0], n1[NQ0], m0[DEF ? NQ0 : 1], m1[DEF ? NQ0 : 1]; // next plane: values (and Jacobians), raw
        if constexpr (PREFETCH)
        {
            load_lines(F, 0, n0, n1);
            if constexpr (DEF) load_lines(J, 0, m0, m1);
        }
#pragma unroll 1
        for (int k = 0; k < NQ2; ++k)
        {
            double f0[NQ0], f1[NQ0];
            if constexpr (!PREFETCH)
            {
                load_lines(F, k, n0, n1);
                if constexpr (DEF) load_lines(J, k, m0, m1);
            }
#pragma unroll
            for (int i = 0; i < NQ0; ++i)
            {
                f0[i] = DEF ? n0[i] * m0[DEF ? i : 0] : n0[i];
                f1[i] = DEF ? n1[i] * m1[DEF ? i : 0] : n1[i];
            }
            if (PREFETCH && k + 1 < NQ2)
            {
                load_lines(F, k + 1, n0, n1);
                if constexpr (DEF) load_lines(J, k + 1, m0, m1);
            }
            const double wk = tab.w2[k] * jreg;
            const double wr0 = ve ? wj0 * wk : 0.0, wr1 = ve ? wj1 * wk : 0.0;
#pragma unroll
            for (int i = 0; i < NQ0; ++i)
            {
                f0[i] *= tab.w0[i] * wr0;
                f1[i] *= tab.w0[i] * wr1;
            }
            // s_p of the two lines and the tiles of p
            const double *ta = sT + g * RS + k * 8 + t;
            tg_static_for<0, NM>([&](auto pc) {
                constexpr int p = decltype(pc)::value;
                double s0 = tab.b0[p * NQ0] * f0[0], s1 = tab.b0[p * NQ0] * f1[0];
#pragma unroll
                for (int i = 1; i < NQ0; ++i)
                {
                    s0 = fma(tab.b0[p * NQ0 + i], f0[i], s0);
                    s1 = fma(tab.b0[p * NQ0 + i], f1[i], s1);
                }
#pragma unroll
                for (int tl = 0; tl < Dm::tiles(p); ++tl)
                {
                    const double *tp = ta + (Dm::tioff(p) + tl) * 8 * RS;
                    tg_mma(C[Dm::tioff(p) + tl][0], C[Dm::tioff(p) + tl][1], tp[0], s0); // k-step (k, j = t)
                    tg_mma(C[Dm::tioff(p) + tl][0], C[Dm::tioff(p) + tl][1], tp[4], s1); // k-step (k, j = 4 + t)
                }
            });
        }
        // results: lane (g, t) holds row 8 tile + g of p for the elements 2t, 2t + 1.  Own modes first ...
        double *s0p = sStg + (2 * t) * NMT, *s1p = s0p + NMT;
        tg_static_for<0, NM>([&](auto pc) {
            constexpr int p = decltype(pc)::value;
#pragma unroll
            for (int tl = 0; tl < Dm::tiles(p); ++tl)
            {
                const int row = 8 * tl + g;
                if (row < Dm::own(p))
                {
                    s0p[Dm::start(p) + row] = C[Dm::tioff(p) + tl][0];
                    s1p[Dm::start(p) + row] = C[Dm::tioff(p) + tl][1];
                }
            }
        });
        __syncwarp();
        // ... then the correction rows of p = 1: top vertex -> mode 1, bottom vertex -> mode nm, singular edge -> nm + r
        if (NM > 1)
        {
#pragma unroll
            for (int tl = 0; tl < Dm::tiles(1); ++tl)
            {
                const int cr = 8 * tl + g - Dm::own(1); // correction row index
                if (cr >= 0 && cr < NM)
                {
                    const int target = cr == 0 ? 1 : NM + cr - 1;
                    s0p[target] += C[Dm::tioff(1) + tl][0];
                    s1p[target] += C[Dm::tioff(1) + tl][1];
                }
            }
        }
        __syncwarp();
        {
            double *o   = args.out + (size_t)b * EB * NMT;
            const int n = ne * NMT;
            if ((NMT % 2 == 0) && args.out_aligned)
                for (int i2 = lane; i2 < n / 2; i2 += 32)
                    *reinterpret_cast<double2 *>(o + 2 * i2) = *reinterpret_cast<const double2 *>(sStg + 2 * i2);
            else
                for (int i = lane; i < n; i += 32) o[i] = sStg[i];
        }
        __syncwarp(); // staging block free for the next batch
    }
}

template <int NM> struct TetIpState
{
    TetIpTab<NM> tab;
    double *d_T = nullptr;
    void *fallback_state          = nullptr;
    void (*fallback_free)(void *) = nullptr;
    int bps[2]                    = {0, 0};
};

template <int NM, bool DEF> static int tet_ip_gemm_launch(nekmf_op_s *op, const double *const in[3], double *const out[3])
{
    auto *st  = static_cast<TetIpState<NM> *>(op->kstate);
    using Dm  = TetIpDims<NM>;
    auto kern = tet_ip_gemm_kernel<NM, DEF>;
    int &bps  = st->bps[DEF ? 1 : 0];
    if (bps == 0)
    {
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Dm::SMEM));
        NEKMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int nb = 0;
        NEKMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, Dm::T, Dm::SMEM));
        if (nb < 1) { set_error("tet IProduct GEMM kernel <%d> does not fit on an SM", NM); return NEKMF_ERR_CUDA; }
        bps = nb;
    }
    TetIpArgs a;
    a.in = in[0]; a.out = out[0]; a.tabT = st->d_T; a.nElmt = op->run_ne;
    a.jac = DEF ? op->d_jac + (size_t)op->run_e0 * op->geo_pitch : op->d_jac + op->run_e0;
    a.in_aligned  = ((((uintptr_t)in[0]) | (DEF ? (uintptr_t)a.jac : 0)) & 15) == 0;
    a.out_aligned = (((uintptr_t)out[0]) & 15) == 0;
    const int nBat = (op->run_ne + Dm::EB - 1) / Dm::EB;
    int grid       = bps * NUM_SMS;
    const int need = (nBat + Dm::WARPS - 1) / Dm::WARPS;
    if (grid > need) grid = need;
    if (grid < 1) return NEKMF_OK;
    kern<<<grid, Dm::T, Dm::SMEM, op->run_stream>>>(st->tab, a);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}

template <int NM> static bool tet_ip_gemm_wrap(nekmf_op_s *op)
{
    using Dm = TetIpDims<NM>;
    constexpr int NQ0 = Dm::NQ0, NQ1 = Dm::NQ1, NQ2 = Dm::NQ2, NPAIR = Dm::NPAIR, NMT = Dm::NMT, RS = Dm::RS;
    if ((int)op->b[0].size() != NM * NQ0 || (int)op->b[1].size() != NPAIR * NQ1 || (int)op->b[2].size() != NMT * NQ2) return false;
    if (op->deformed && op->geo_pitch != op->nqTot) return false;
    const double *b1 = op->b[1].data(), *b2 = op->b[2].data();
    std::vector<double> T(Dm::TOT_T, 0.0);
    int mode = 0, pair = 0;
    for (int p = 0; p < NM; ++p)
    {
        double *tp   = T.data() + (size_t)Dm::tioff(p) * 8 * RS; // rows of p are consecutive across its tiles
        auto add_row = [&](int row, int b2row, auto b1of) {
            for (int k = 0; k < NQ2; ++k)
                for (int j = 0; j < NQ1; ++j) tp[row * RS + k * 8 + j] += b2[b2row * NQ2 + k] * b1of(j);
        };
        int row = 0;
        for (int q = 0; q < NM - p; ++q, ++pair)
            for (int r = 0; r < NM - p - q; ++r, ++mode, ++row)
            {
                const int pr = pair;
                add_row(row, mode, [&](int j) { return b1[pr * NQ1 + j]; });
            }
        if (p == 0 && NM > 1) add_row(1, 1, [&](int j) { return b1[NQ1 + j]; }); // top vertex, p = 0 part
        if (p == 1)
        {
            add_row(row++, 1, [&](int j) { return b1[j] + b1[NQ1 + j]; });              // top vertex, p = 1 part -> mode 1
            add_row(row++, 0, [&](int j) { return b1[NQ1 + j]; });                      // bottom vertex -> mode nm
            for (int r = 1; r < NM - 1; ++r) add_row(row++, r + 1, [&](int j) { return b1[NQ1 + j]; }); // edge -> nm + r
        }
    }
    auto *stt = new TetIpState<NM>;
    memcpy(stt->tab.b0, op->b[0].data(), sizeof(stt->tab.b0));
    memcpy(stt->tab.w0, op->ws[0].data(), sizeof(stt->tab.w0));
    memcpy(stt->tab.w1, op->ws[1].data(), sizeof(stt->tab.w1));
    memcpy(stt->tab.w2, op->ws[2].data(), sizeof(stt->tab.w2));
    if (cudaMalloc(&stt->d_T, T.size() * 8) != cudaSuccess || cudaMemcpy(stt->d_T, T.data(), T.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess)
    {
        cudaFree(stt->d_T);
        delete stt;
        return false;
    }
    stt->fallback_state = op->kstate;
    stt->fallback_free  = op->kstate_free;
    op->kstate          = stt;
    op->kstate_free     = [](void *p) {
        auto *s = static_cast<TetIpState<NM> *>(p);
        if (s->fallback_state && s->fallback_free) s->fallback_free(s->fallback_state);
        cudaFree(s->d_T);
        delete s;
    };
    op->launch = op->deformed ? tet_ip_gemm_launch<NM, true> : tet_ip_gemm_launch<NM, false>;
    char name[128];
    snprintf(name, sizeof(name), "tet_ip_gemm_kernel<nm=%d,%s>(DMMA m8n8k4 over 8 elements, collapsed contractions pre-combined)", NM,
             op->deformed ? "deformed" : "regular");
    op->kname = name;
    return true;
}

// called from select_shape_fast before tet_dmma_maybe_wrap: BwdTrans / IProductWRTBase on tetrahedra, default quadrature.
// NEKMF_TET_GEMM=0 leaves the operator to tet_dmma.cu / the pencil kernel (the other arms of the A/B).
bool tet_gemm_maybe_wrap(nekmf_op_s *op)
{
    if (op->shape != NEKMF_TET || (op->optype != NEKMF_BWDTRANS && op->optype != NEKMF_IPRODUCTWRTBASE)) return false;
    const int nm = op->nm[0];
    if (op->nm[1] != nm || op->nm[2] != nm || op->nq[0] != nm + 1 || op->nq[1] != nm || op->nq[2] != nm) return false;
    const char *v = getenv("NEKMF_TET_GEMM");
    if (v && v[0] == '0') return false;
    if (op->optype == NEKMF_IPRODUCTWRTBASE)
    {
        // measured (profiles/r02_sweep_tet_gemm_1.jsonl against r02_sweep_tet_dmma_{all,0}.jsonl, fraction of the HBM peak):
        // regular 0.50 / 0.48 / 0.53 at nm = 5 / 6 / 7 (pencil 0.25 / 0.32 / 0.25); deformed 0.51 / 0.47 / 0.58 (tet_dmma.cu
        // 0.40 / 0.49 / 0.41): deformed nm = 6 (odd i-lines: scalar loads) stays with tet_dmma.cu
        if (op->deformed && nm == 6 && !(v && v[0] == 'a')) return false;
        switch (nm)
        {
            case 5: return tet_ip_gemm_wrap<5>(op);
            case 6: return tet_ip_gemm_wrap<6>(op);
            case 7: return tet_ip_gemm_wrap<7>(op);
            default: return false;
        }
    }
    switch (nm)
    {
        case 5: return tet_gemm_wrap<5>(op);
        case 6: return tet_gemm_wrap<6>(op);
        case 7: return tet_gemm_wrap<7>(op);
        default: return false;
    }
}

} // namespace nekmf
