// assembly.cu -- AssemblyMapCG local <-> global on the device.
//
//   nekmf_map_global_to_local : loc[i] = sign[i] * glob[map[i]]
//       (AssemblyMapCG::v_GlobalToLocal, AssemblyMapCG.cpp:2853-2876 -> Vmath::Gathr, Vmath.hpp:217-230)
//   nekmf_map_assemble        : glob = 0; glob[map[i]] += sign[i] * loc[i]
//       (AssemblyMapCG::v_Assemble, AssemblyMapCG.cpp:2885-2910 -> Vmath::Assmb, Vmath.hpp:232-244)
//
// The reference's scatter-add is a sequential loop.  Here the map is transposed once at creation
// (CSR: for every global DOF the ascending list of local indices that map to it) and Assemble
// becomes a gather: one thread per global DOF sums its copies in ascending local index -- the
// same order as the sequential loop, so the result is bit-identical and deterministic, with no
// atomics and no zero-fill pass.
#include "map_internal.h"
#include <algorithm>

namespace nekmf
{

__global__ void g2l_kernel(const int *__restrict__ map, const double *__restrict__ sign,
                           const double *__restrict__ glob, double *__restrict__ loc, int n, int vec)
{
    // 2 entries per thread, 16-byte stores when loc is 16-byte aligned
    const int i2 = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (!vec)
    {
        for (int i = i2; i < n && i < i2 + 2; ++i)
        {
            double v = __ldg(glob + map[i]);
            if (sign) v *= sign[i];
            loc[i] = v;
        }
        return;
    }
    if (i2 + 1 < n)
    {
        const int2 m = *reinterpret_cast<const int2 *>(map + i2);
        double2 v;
        v.x = __ldg(glob + m.x);
        v.y = __ldg(glob + m.y);
        if (sign)
        {
            const double2 s = *reinterpret_cast<const double2 *>(sign + i2);
            v.x *= s.x;
            v.y *= s.y;
        }
        *reinterpret_cast<double2 *>(loc + i2) = v;
    }
    else if (i2 < n)
    {
        double v = __ldg(glob + map[i2]);
        if (sign) v *= sign[i2];
        loc[i2] = v;
    }
}

__global__ void assemble_kernel(const int *__restrict__ rowptr, const int *__restrict__ col,
                                const double *__restrict__ sign, const double *__restrict__ loc,
                                double *__restrict__ glob, int nGlobal)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nGlobal) return;
    const int b = rowptr[g], e = rowptr[g + 1];
    double s = 0.0;
    for (int k = b; k < e; ++k)
    {
        const int i = col[k];
        s += sign ? sign[i] * __ldg(loc + i) : __ldg(loc + i);
    }
    glob[g] = s;
}

// Assemble fused with the CG dot product mu = sum_{g >= nDir} owned(g) * glob[g] * w[g] (the s.w of
// NekLinSysIterCG.cpp:226-235): fixed grid, grid-stride, one partial sum per block -> deterministic.
// Sharded solves (EX): the first blocks start with the partition-interface entries -- they assemble those DOFs
// and store the values straight into the neighbours' receive windows over NVLink (peer-memory transport) or into
// the NCCL send buffer, and the last of them raises the neighbours' flags; the transfer then overlaps the rest of
// the assemble.  The s.w terms of shared DOFs are left to the unpack kernel (their sums are not complete yet).
template <bool EX>
__global__ void __launch_bounds__(256, 8) // 32 registers: all RED_BLOCKS = 8 x 148 blocks resident, one wave (40 registers
                                           // left 6 per SM: a second, third-full wave)
    assemble_dot_kernel(const int *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ sign,
                        const double *__restrict__ loc, double *__restrict__ glob, int nGlobal,
                        const double *__restrict__ w, const unsigned char *__restrict__ flags, int nDir,
                        double *__restrict__ part, const __grid_constant__ nekmf_exdev ex, int nIfBlocks,
                        const int *__restrict__ skip)
{
    __shared__ double sh[8];
    if (!EX && skip && *skip) return; // single-rank solves: an iteration enqueued after convergence
    auto row_sum = [&](int g) {
        const int b = rowptr[g], e = rowptr[g + 1];
        double s = 0.0;
        for (int k = b; k < e; ++k)
        {
            const int i = col[k];
            s += sign ? sign[i] * __ldg(loc + i) : __ldg(loc + i);
        }
        return s;
    };
    if (EX && (int)blockIdx.x < nIfBlocks)
    {
        const unsigned long long epoch = *ex.epoch + 1ull;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ex.total; i += nIfBlocks * blockDim.x)
            exchange_put(ex, i, epoch, row_sum(ex.idx[i]));
        exchange_signal(ex, epoch, (unsigned int)nIfBlocks);
    }
    // (measured alternative: four rows per thread with interleaved rowptr -> col -> loc chains: 0.73 -> 1.25 ms on
    // 2^20 hex elements -- the four row windows thrash the L2 lines the neighbouring rows share)
    // The row's chain rowptr -> col -> loc is three dependent DRAM latencies (ncu: 45 long-scoreboard stall cycles per
    // issue at full occupancy, 65 % of the HBM peak), so the grid-stride loop is software-pipelined: this trip loads
    // rowptr of the row two trips ahead, the first two column indices of the next row and the values of its own row,
    // all independent of each other.  The sum keeps the order s = ((0 + t0) + t1) + ... of the plain loop.
    auto val = [&](int i) { return sign ? sign[i] * __ldg(loc + i) : __ldg(loc + i); };
    const int stride = gridDim.x * blockDim.x;
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    int b0 = 0, e0 = 0, b1 = 0, e1 = 0, c00 = 0, c01 = 0;
    if (g < nGlobal) { b0 = rowptr[g]; e0 = rowptr[g + 1]; }
    if (g < nGlobal - stride) { b1 = rowptr[g + stride]; e1 = rowptr[g + stride + 1]; }
    if (e0 > b0) c00 = col[b0];
    if (e0 > b0 + 1) c01 = col[b0 + 1];
    double mu = 0.0;
    for (; g < nGlobal; g += stride)
    {
        int b2 = 0, e2 = 0;
        if (g < nGlobal - 2 * stride) { b2 = rowptr[g + 2 * stride]; e2 = rowptr[g + 2 * stride + 1]; }
        double v0 = 0.0, v1 = 0.0, wg = 0.0;
        if (e0 > b0) v0 = val(c00);
        if (e0 > b0 + 1) v1 = val(c01);
        unsigned char f = 0;
        if (g >= nDir)
        {
            f  = flags ? flags[g] : (unsigned char)1;
            wg = w[g];
        }
        int c10 = 0, c11 = 0;
        if (e1 > b1) c10 = col[b1];
        if (e1 > b1 + 1) c11 = col[b1 + 1];
        double s = v0 + v1;
        for (int k = b0 + 2; k < e0; ++k) s += val(col[k]);
        glob[g] = s;
        if ((f & 1) && !(f & 2)) mu = fma(s, wg, mu);
        b0 = b1; e0 = e1; c00 = c10; c01 = c11;
        b1 = b2; e1 = e2;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mu += __shfl_xor_sync(0xffffffffu, mu, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mu;
    __syncthreads();
    if (threadIdx.x < 32)
    {
        double r = threadIdx.x < 8 ? sh[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        if (threadIdx.x == 0) part[blockIdx.x] = r;
    }
}

} // namespace nekmf

using namespace nekmf;

extern "C" {

int nekmf_map_create(int nLocal, int nGlobal, const int *l2g, const double *sign, nekmf_map_t *out)
{
    if (!out) { set_error("nekmf_map_create: null output handle"); return NEKMF_ERR_ARG; }
    *out = nullptr;
    if (nLocal < 0 || nGlobal < 0 || (nLocal > 0 && !l2g))
    {
        set_error("nekmf_map_create: bad sizes or null map");
        return NEKMF_ERR_ARG;
    }
    std::vector<int> rowptr((size_t)nGlobal + 1, 0), col((size_t)nLocal);
    for (int i = 0; i < nLocal; ++i)
    {
        if (l2g[i] < 0 || l2g[i] >= nGlobal)
        {
            set_error("nekmf_map_create: localToGlobal[%d] = %d out of range [0,%d)", i, l2g[i], nGlobal);
            return NEKMF_ERR_ARG;
        }
        ++rowptr[(size_t)l2g[i] + 1];
    }
    for (int g = 0; g < nGlobal; ++g) rowptr[g + 1] += rowptr[g];
    {
        std::vector<int> cur(rowptr.begin(), rowptr.end() - 1);
        for (int i = 0; i < nLocal; ++i) col[cur[l2g[i]]++] = i; // ascending local index per row
    }
    if (nekmf_device_count() < 1)
    {
        set_error("nekmf_map_create: no CUDA device (this library has no CPU fallback)");
        return NEKMF_ERR_CUDA;
    }
    nekmf_map_s *m = new nekmf_map_s;
    m->nLocal      = nLocal;
    m->nGlobal     = nGlobal;
    const size_t nl = nLocal ? nLocal : 1;
    cudaError_t e   = cudaMalloc(&m->d_map, (nl + 1) * 4);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_col, nl * 4);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_rowptr, ((size_t)nGlobal + 1) * 4);
    if (e == cudaSuccess && sign) e = cudaMalloc(&m->d_sign, (nl + 1) * 8);
    if (e == cudaSuccess) e = cudaMemcpy(m->d_map, l2g, (size_t)nLocal * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(m->d_col, col.data(), (size_t)nLocal * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(m->d_rowptr, rowptr.data(), rowptr.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && sign) e = cudaMemcpy(m->d_sign, sign, (size_t)nLocal * 8, cudaMemcpyHostToDevice);
    if (e != cudaSuccess)
    {
        set_error("nekmf_map_create: %s", cudaGetErrorString(e));
        nekmf_map_destroy(m);
        return NEKMF_ERR_CUDA;
    }
    *out = m;
    return NEKMF_OK;
}

int nekmf_map_destroy(nekmf_map_t m)
{
    if (!m) return NEKMF_OK;
    cudaFree(m->d_map);
    cudaFree(m->d_col);
    cudaFree(m->d_rowptr);
    cudaFree(m->d_sign);
    cudaFree(m->d_stage_loc);
    cudaFree(m->d_stage_glob);
    delete m;
    return NEKMF_OK;
}

} // extern "C"

namespace nekmf
{
int map_g2l_device(nekmf_map_s *m, const double *glob, double *loc, cudaStream_t st)
{
    if (m->nLocal == 0) return NEKMF_OK;
    const int threads = 256, pairs = (m->nLocal + 1) / 2;
    g2l_kernel<<<(pairs + threads - 1) / threads, threads, 0, st>>>(m->d_map, m->d_sign, glob, loc, m->nLocal,
                                                                       (((uintptr_t)loc) & 15) == 0);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}
int map_assemble_device(nekmf_map_s *m, const double *loc, double *glob, cudaStream_t st)
{
    if (m->nGlobal == 0) return NEKMF_OK;
    const int threads = 256;
    assemble_kernel<<<(m->nGlobal + threads - 1) / threads, threads, 0, st>>>(m->d_rowptr, m->d_col, m->d_sign, loc,
                                                                               glob, m->nGlobal);
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}
int map_assemble_dot_device(nekmf_map_s *m, const double *loc, double *glob, const double *w,
                            const unsigned char *flags, int nDir, double *part, const nekmf_exdev *ex, cudaStream_t st,
                            const int *skip)
{
    if (ex && ex->total > 0)
    {
        int nIf = (ex->total + 255) / 256;
        if (nIf > RED_BLOCKS) nIf = RED_BLOCKS;
        assemble_dot_kernel<true><<<RED_BLOCKS, 256, 0, st>>>(m->d_rowptr, m->d_col, m->d_sign, loc, glob, m->nGlobal, w,
                                                              flags, nDir, part, *ex, nIf, nullptr);
    }
    else
    {
        nekmf_exdev none;
        assemble_dot_kernel<false><<<RED_BLOCKS, 256, 0, st>>>(m->d_rowptr, m->d_col, m->d_sign, loc, glob, m->nGlobal,
                                                               w, flags, nDir, part, none, 0, skip);
    }
    ++g_launches;
    NEKMF_CUDA(cudaGetLastError());
    return NEKMF_OK;
}
static int map_stage(nekmf_map_s *m)
{
    if (!m->d_stage_loc) NEKMF_CUDA(cudaMalloc(&m->d_stage_loc, ((size_t)m->nLocal + 2) * 8));
    if (!m->d_stage_glob) NEKMF_CUDA(cudaMalloc(&m->d_stage_glob, ((size_t)m->nGlobal + 2) * 8));
    return NEKMF_OK;
}
} // namespace nekmf

extern "C" {

int nekmf_map_global_to_local(nekmf_map_t m, const double *glob, double *loc, int memkind, void *stream)
{
    if (!m || !glob || !loc) { set_error("nekmf_map_global_to_local: null argument"); return NEKMF_ERR_ARG; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (memkind == NEKMF_DEVICE) return map_g2l_device(m, glob, loc, st);
    int rc = map_stage(m);
    if (rc) return rc;
    NEKMF_CUDA(cudaMemcpyAsync(m->d_stage_glob, glob, (size_t)m->nGlobal * 8, cudaMemcpyHostToDevice, st));
    rc = map_g2l_device(m, m->d_stage_glob, m->d_stage_loc, st);
    if (rc) return rc;
    NEKMF_CUDA(cudaMemcpyAsync(loc, m->d_stage_loc, (size_t)m->nLocal * 8, cudaMemcpyDeviceToHost, st));
    NEKMF_CUDA(cudaStreamSynchronize(st));
    return NEKMF_OK;
}

int nekmf_map_assemble(nekmf_map_t m, const double *loc, double *glob, int memkind, void *stream)
{
    if (!m || !glob || !loc) { set_error("nekmf_map_assemble: null argument"); return NEKMF_ERR_ARG; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (memkind == NEKMF_DEVICE) return map_assemble_device(m, loc, glob, st);
    int rc = map_stage(m);
    if (rc) return rc;
    NEKMF_CUDA(cudaMemcpyAsync(m->d_stage_loc, loc, (size_t)m->nLocal * 8, cudaMemcpyHostToDevice, st));
    rc = map_assemble_device(m, m->d_stage_loc, m->d_stage_glob, st);
    if (rc) return rc;
    NEKMF_CUDA(cudaMemcpyAsync(glob, m->d_stage_glob, (size_t)m->nGlobal * 8, cudaMemcpyDeviceToHost, st));
    NEKMF_CUDA(cudaStreamSynchronize(st));
    return NEKMF_OK;
}

} // extern "C"
