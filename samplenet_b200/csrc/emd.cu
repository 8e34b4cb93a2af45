// emd.cu -- approx_match / match_cost / match_cost_grad (EMD).
//
// Reference behaviour restated (not ported): classification/structural_losses/tf_approxmatch_g.cu:1-295.  The reference
// runs ONE CTA of 512 threads per cloud on a fixed grid of 32 CTAs (so at most 32 of the 148 SMs ever work), keeps the
// remain/ratio vectors in global scratch, zeroes `match` and then read-modify-writes it once per level (10 sweeps).
//
// B200 design: a thread-block CLUSTER owns one cloud, so a batch of 50 clouds fills the chip.  The row dimension of every
// phase is split over the cluster's CTAs and, inside a CTA, a row is shared by S lanes that each take the columns
// j == lane (mod S) and merge their partial sums by shuffles; the opposite side (xyz + its per-point weight, as float4) is
// staged through shared memory in tiles.  The four per-point vectors live in a small global scratch that stays in L2;
// phases are separated by cluster barriers (release/acquire).  `match` is written (not accumulated) at the first level,
// which removes the zero-fill sweep and one read sweep.  exp() is evaluated as ex2(level*log2e * d2) with the exact-range
// intrinsic exp2f (not the reference's __expf), so values agree with the CPU oracle to fp32 rounding.
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace snb {

constexpr int kEmdThreads = 512;
constexpr int kEmdTile = 1024;  // opposite-side points per shared-memory tile (float4 each: 16 KB)

struct EmdParams {
    int b, n, m;
    int S;           // lanes per row
    const float *xyz1, *xyz2;
    float *match;    // (b, m, n)
    float *temp;     // (b, 2*(n+m)): remainL[n], remainR[m], ratioL[n], ratioR[m]
};

// Row-parallel reduction: for every row r owned by this CTA group, acc = sum_j f(row r, column j) over all columns.
// RowSide: 0 => rows are xyz1 points (k, n of them), columns xyz2 (l, m of them);  1 => rows xyz2, columns xyz1.
__device__ __forceinline__ float emd_sq(float ax, float ay, float az, float bx, float by, float bz)
{
    const float dx = bx - ax, dy = by - ay, dz = bz - az;
    return dx * dx + dy * dy + dz * dz;
}

__global__ void __launch_bounds__(kEmdThreads) approxmatch_kernel(const __grid_constant__ EmdParams P)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = cluster.block_rank();
    const int csize = cluster.num_blocks();
    const int bi = blockIdx.x / csize;
    const int n = P.n, m = P.m, S = P.S;
    const float *p1 = P.xyz1 + (size_t)bi * n * 3;
    const float *p2 = P.xyz2 + (size_t)bi * m * 3;
    float *match = P.match + (size_t)bi * n * m;
    float *remainL = P.temp + (size_t)bi * (n + m) * 2, *remainR = remainL + n, *ratioL = remainR + m, *ratioR = ratioL + n;

    __shared__ float4 s_o[kEmdTile];

    const int rows_per_pass = (csize * kEmdThreads) / S;           // rows the whole cluster handles at once
    const int my_row_slot = (crank * kEmdThreads + threadIdx.x) / S;  // my row within a pass
    const int l_in = threadIdx.x % S;                               // my lane within the row group

    float multiL, multiR;  // tf_approxmatch_g.cu:4-10 (integer division)
    if (n >= m) { multiL = 1; multiR = (float)(n / m); } else { multiL = (float)(m / n); multiR = 1; }
    for (int j = crank * kEmdThreads + threadIdx.x; j < n; j += csize * kEmdThreads) remainL[j] = multiL;
    for (int j = crank * kEmdThreads + threadIdx.x; j < m; j += csize * kEmdThreads) remainR[j] = multiR;
    cluster.sync();

    for (int lev = 7; lev >= -2; lev--) {
        // level = -4^lev (0 at the last level); exp(level*d) == exp2(level*log2(e)*d)
        const float level = (lev == -2) ? 0.f : -powf(4.0f, (float)lev);
        const float level2 = level * 1.44269504088896340736f;
        const bool first = (lev == 7);

        // ---- phase 1 (:27-60): ratioL[k] = remainL[k] / (1e-9 + sum_l exp(level*d) * remainR[l])
        for (int r0 = 0; r0 < n; r0 += rows_per_pass) {
            const int k = r0 + my_row_slot;
            const bool live = k < n;
            float x1 = 0, y1 = 0, z1 = 0;
            if (live) { x1 = p1[k * 3 + 0]; y1 = p1[k * 3 + 1]; z1 = p1[k * 3 + 2]; }
            float suml = 0.f;
            for (int l0 = 0; l0 < m; l0 += kEmdTile) {
                const int ln = min(kEmdTile, m - l0);
                __syncthreads();
                for (int l = threadIdx.x; l < ln; l += kEmdThreads)
                    s_o[l] = make_float4(p2[(l0 + l) * 3 + 0], p2[(l0 + l) * 3 + 1], p2[(l0 + l) * 3 + 2], remainR[l0 + l]);
                __syncthreads();
                if (live)
                    for (int l = l_in; l < ln; l += S) {
                        const float4 o = s_o[l];
                        suml += exp2f(level2 * emd_sq(x1, y1, z1, o.x, o.y, o.z)) * o.w;
                    }
            }
            for (int o = S >> 1; o > 0; o >>= 1) suml += __shfl_xor_sync(kFullMask, suml, o);
            if (live && l_in == 0) ratioL[k] = remainL[k] / (suml + 1e-9f);
        }
        cluster.sync();

        // ---- phase 2 (:75-111): per xyz2 point l: sumr = remainR[l] * sum_k exp(level*d) * ratioL[k]
        for (int r0 = 0; r0 < m; r0 += rows_per_pass) {
            const int l = r0 + my_row_slot;
            const bool live = l < m;
            float x2 = 0, y2 = 0, z2 = 0;
            if (live) { x2 = p2[l * 3 + 0]; y2 = p2[l * 3 + 1]; z2 = p2[l * 3 + 2]; }
            float sumr = 0.f;
            for (int k0 = 0; k0 < n; k0 += kEmdTile) {
                const int kn = min(kEmdTile, n - k0);
                __syncthreads();
                for (int k = threadIdx.x; k < kn; k += kEmdThreads)
                    s_o[k] = make_float4(p1[(k0 + k) * 3 + 0], p1[(k0 + k) * 3 + 1], p1[(k0 + k) * 3 + 2], ratioL[k0 + k]);
                __syncthreads();
                if (live)
                    for (int k = l_in; k < kn; k += S) {
                        const float4 o = s_o[k];
                        sumr += exp2f(level2 * emd_sq(o.x, o.y, o.z, x2, y2, z2)) * o.w;
                    }
            }
            for (int o = S >> 1; o > 0; o >>= 1) sumr += __shfl_xor_sync(kFullMask, sumr, o);
            if (live && l_in == 0) {
                const float rr = remainR[l];
                sumr *= rr;
                const float consumption = fminf(rr / (sumr + 1e-9f), 1.0f);
                ratioR[l] = consumption * rr;
                remainR[l] = fmaxf(0.0f, rr - sumr);
            }
        }
        cluster.sync();

        // ---- phase 3 (:127-160): w = exp(level*d) * ratioL[k] * ratioR[l]; match[l][k] += w; remainL[k] -= sum_l w
        // Rows are k; the S lanes of a row take different l.  To keep the match stores coalesced (k fastest) S is 1 here:
        // consecutive threads own consecutive k and walk l together.
        for (int r0 = 0; r0 < n; r0 += csize * kEmdThreads) {
            const int k = r0 + crank * kEmdThreads + threadIdx.x;
            const bool live = k < n;
            float x1 = 0, y1 = 0, z1 = 0, rl = 0;
            if (live) { x1 = p1[k * 3 + 0]; y1 = p1[k * 3 + 1]; z1 = p1[k * 3 + 2]; rl = ratioL[k]; }
            float suml = 0.f;
            for (int l0 = 0; l0 < m; l0 += kEmdTile) {
                const int ln = min(kEmdTile, m - l0);
                __syncthreads();
                for (int l = threadIdx.x; l < ln; l += kEmdThreads)
                    s_o[l] = make_float4(p2[(l0 + l) * 3 + 0], p2[(l0 + l) * 3 + 1], p2[(l0 + l) * 3 + 2], ratioR[l0 + l]);
                __syncthreads();
                if (live) {
#pragma unroll 4
                    for (int l = 0; l < ln; l++) {
                        const float4 o = s_o[l];
                        const float w = exp2f(level2 * emd_sq(x1, y1, z1, o.x, o.y, o.z)) * rl * o.w;
                        float *mp = match + (size_t)(l0 + l) * n + k;
                        *mp = first ? w : (*mp + w);
                        suml += w;
                    }
                }
            }
            if (live) remainL[k] = fmaxf(0.0f, remainL[k] - suml);
        }
        cluster.sync();
    }
}

size_t approxmatch_workspace_bytes(int b, int n, int m) { return (size_t)b * (n + m) * 2 * sizeof(float); }

int launch_approxmatch(int b, int n, int m, const float *xyz1, const float *xyz2, float *match, void *workspace, cudaStream_t stream)
{
    EmdParams P;
    P.b = b; P.n = n; P.m = m; P.xyz1 = xyz1; P.xyz2 = xyz2; P.match = match; P.temp = reinterpret_cast<float *>(workspace);
    // cluster size: as many CTAs per cloud as still have >= 1 row per thread in phase 3, bounded by the SM budget
    int csize = 1;
    const int rows = max(n, m);
    while (csize < 8 && (long long)b * csize * 2 <= 2 * kNumSMs && csize * kEmdThreads < rows) csize *= 2;
    // lanes per row for the reductions of phases 1/2: use the idle threads when rows < cluster threads
    int S = 1;
    while (S < 32 && (csize * kEmdThreads) / (S * 2) >= min(n, m)) S *= 2;
    P.S = S;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(b * csize);
    cfg.blockDim = dim3(kEmdThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, approxmatch_kernel, P);
    if (e != cudaSuccess) { set_error("approxmatch: launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return SNB200_ECUDA; }
    return check_launch("approxmatch");
}

// ------------------------------------------------------------------------------------------------------------------
// match_cost (:183-225): cost[b] = sum_{k,l} match[b][l][k] * ||xyz1[k] - xyz2[l]||.
// One cluster-free CTA per (cloud, slab of l); threads run along k so match reads are coalesced; per-CTA partials are
// combined by the LAST CTA of each cloud in slab order (deterministic), using a per-cloud arrival counter in `cost`'s
// shadow... kept simple: slabs write partials to a small static device buffer indexed by (cloud, slab).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kMcThreads = 256;
constexpr int kMcSlabs = 16;

__global__ void __launch_bounds__(kMcThreads) matchcost_partial_kernel(int b, int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                                                                      const float *__restrict__ match, float *__restrict__ partial)
{
    __shared__ float s_red[kMcThreads / 32];
    const int bi = blockIdx.y, slab = blockIdx.x;
    const int l_beg = (int)((long long)m * slab / kMcSlabs), l_end = (int)((long long)m * (slab + 1) / kMcSlabs);
    const float *p1 = xyz1 + (size_t)bi * n * 3, *p2 = xyz2 + (size_t)bi * m * 3;
    const float *mt = match + (size_t)bi * n * m;
    float sub = 0.f;
    for (int k = threadIdx.x; k < n; k += kMcThreads) {
        const float x1 = p1[k * 3 + 0], y1 = p1[k * 3 + 1], z1 = p1[k * 3 + 2];
        for (int l = l_beg; l < l_end; l++) {
            const float x2 = __ldg(p2 + l * 3 + 0), y2 = __ldg(p2 + l * 3 + 1), z2 = __ldg(p2 + l * 3 + 2);
            sub += sqrtf(emd_sq(x1, y1, z1, x2, y2, z2)) * mt[(size_t)l * n + k];
        }
    }
    sub = warp_sum(sub);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = sub;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < kMcThreads / 32; i++) t += s_red[i];
        partial[bi * kMcSlabs + slab] = t;
    }
}
__global__ void matchcost_final_kernel(int b, const float *__restrict__ partial, float *__restrict__ cost)
{
    const int bi = blockIdx.x * blockDim.x + threadIdx.x;
    if (bi < b) {
        float t = 0.f;
        for (int s = 0; s < kMcSlabs; s++) t += partial[bi * kMcSlabs + s];
        cost[bi] = t;
    }
}

// grad1 (:263-291): grad1[k] = sum_l match[l][k] * (x1-x2) / max(|x1-x2|, 1e-10); thread per k, l broadcast from smem.
__global__ void __launch_bounds__(256) matchcostgrad1_kernel(int b, int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                                                             const float *__restrict__ match, float *__restrict__ grad1)
{
    __shared__ float s_o[kEmdTile * 3];
    const int bi = blockIdx.y;
    const int k = blockIdx.x * 256 + threadIdx.x;
    const bool live = k < n;
    const float *p1 = xyz1 + (size_t)bi * n * 3, *p2 = xyz2 + (size_t)bi * m * 3;
    const float *mt = match + (size_t)bi * n * m;
    float x1 = 0, y1 = 0, z1 = 0;
    if (live) { x1 = p1[k * 3 + 0]; y1 = p1[k * 3 + 1]; z1 = p1[k * 3 + 2]; }
    float dx = 0, dy = 0, dz = 0;
    for (int l0 = 0; l0 < m; l0 += kEmdTile) {
        const int ln = min(kEmdTile, m - l0);
        __syncthreads();
        for (int i = threadIdx.x; i < ln * 3; i += 256) s_o[i] = p2[(size_t)l0 * 3 + i];
        __syncthreads();
        if (live)
            for (int l = 0; l < ln; l++) {
                const float ex = x1 - s_o[l * 3 + 0], ey = y1 - s_o[l * 3 + 1], ez = z1 - s_o[l * 3 + 2];
                const float d = mt[(size_t)(l0 + l) * n + k] * rsqrtf(fmaxf(ex * ex + ey * ey + ez * ez, 1e-20f));
                dx += ex * d; dy += ey * d; dz += ez * d;
            }
    }
    if (live) {
        float *g = grad1 + ((size_t)bi * n + k) * 3;
        g[0] = dx; g[1] = dy; g[2] = dz;
    }
}

// grad2 (:229-262): grad2[l] = sum_k match[l][k] * (x2-x1) / max(|x2-x1|, 1e-10); one warp per l, lanes along k (coalesced).
__global__ void __launch_bounds__(256) matchcostgrad2_kernel(int b, int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                                                             const float *__restrict__ match, float *__restrict__ grad2)
{
    const int bi = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int l = blockIdx.x * 8 + warp;
    if (l >= m) return;
    const float *p1 = xyz1 + (size_t)bi * n * 3, *p2 = xyz2 + (size_t)bi * m * 3;
    const float *mt = match + (size_t)bi * n * m + (size_t)l * n;
    const float x2 = p2[l * 3 + 0], y2 = p2[l * 3 + 1], z2 = p2[l * 3 + 2];
    float sx = 0, sy = 0, sz = 0;
    for (int k = lane; k < n; k += 32) {
        const float ex = x2 - __ldg(p1 + k * 3 + 0), ey = y2 - __ldg(p1 + k * 3 + 1), ez = z2 - __ldg(p1 + k * 3 + 2);
        const float d = mt[k] * rsqrtf(fmaxf(ex * ex + ey * ey + ez * ez, 1e-20f));
        sx += ex * d; sy += ey * d; sz += ez * d;
    }
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
    if (lane == 0) {
        float *g = grad2 + ((size_t)bi * m + l) * 3;
        g[0] = sx; g[1] = sy; g[2] = sz;
    }
}

int launch_matchcost(int b, int n, int m, const float *xyz1, const float *xyz2, const float *match, float *cost, float *partial, cudaStream_t stream)
{
    matchcost_partial_kernel<<<dim3(kMcSlabs, b), kMcThreads, 0, stream>>>(b, n, m, xyz1, xyz2, match, partial);
    int rc = check_launch("matchcost(partial)");
    if (rc) return rc;
    matchcost_final_kernel<<<(b + 127) / 128, 128, 0, stream>>>(b, partial, cost);
    return check_launch("matchcost(final)");
}

int launch_matchcostgrad(int b, int n, int m, const float *xyz1, const float *xyz2, const float *match, float *grad1, float *grad2, cudaStream_t stream)
{
    matchcostgrad1_kernel<<<dim3((n + 255) / 256, b), 256, 0, stream>>>(b, n, m, xyz1, xyz2, match, grad1);
    int rc = check_launch("matchcostgrad1");
    if (rc) return rc;
    matchcostgrad2_kernel<<<dim3((m + 7) / 8, b), 256, 0, stream>>>(b, n, m, xyz1, xyz2, match, grad2);
    return check_launch("matchcostgrad2");
}

}  // namespace snb
