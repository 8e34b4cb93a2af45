// conv_stack.cu -- the whole per-point MLP (conv layers 1..L, samplenet.py:90-94) as ONE persistent cooperative kernel whose
// activations never leave the SM.
//
// Per-layer kernels (encoder_tc.cu) pay, per layer, a launch, a 8-17 MB activation write + read through L2 and a cold
// prologue/epilogue; at B=32 that is 65 us for 2.1 GFLOP.  Here every CTA (one per SM) owns up to TWO 128-point tiles for the
// whole stack:
//   * the raw (pre-BatchNorm) output of layer l stays in TENSOR MEMORY (128 lanes x c_out columns per tile; two ping-pong
//     regions per tile slot = 4 x 128 columns = the SM's 512 TMEM columns);
//   * layer l+1's A operand is produced straight from TMEM: tcgen05.ld (thread = point row) -> BatchNorm + ReLU of layer l ->
//     exact hi/lo TF32 split -> canonical K-major SWIZZLE_128B shared-memory tile (ring of two 32-wide K chunks), so operand
//     preparation of chunk k+1 overlaps the tcgen05.mma of chunk k (3 MMAs per K-step: lo*hi, hi*lo, hi*hi);
//   * the layer's weights are split and staged once per CTA per layer (while the previous layer's MMAs and the grid barrier
//     are in flight);
//   * training-mode BatchNorm needs batch statistics of layer l before layer l+1 can start: per-tile column sums are reduced
//     with a halving shuffle network (31 shuffles per 32x32 block), combined per CTA and added to fp64 global accumulators,
//     and a grid-wide barrier (cooperative launch, one atomic counter) separates the layers.  Layer 1 (3 -> 64) is evaluated
//     on the fly from the cloud and its statistics follow analytically from the batch's 9 input moments (phase 0);
//   * the last layer never materialises: only per-tile column max / min leave the SM (the max-pool commutes with the monotone
//     BN+ReLU map).
//   * the max-pool finalise and the FC head (fc1..fc4 with BatchNorm over the batch) run as the tail of the same launch, 8 output
//     channels per CTA; the 32 KB activation matrix of a layer travels between CTAs as self-validating words (a zeroed buffer,
//     producers never store the bit pattern 0, consumers spin on the data itself): no grid barrier in the head.
// Applicable when every CTA's tiles fit its TMEM (tiles <= 2 x CTAs, widths <= 128, K multiples of 32); otherwise the
// per-layer kernels are used.
#include "encoder_internal.cuh"
#include <cooperative_groups.h>
#include <string.h>

namespace snb {
namespace v1 {   // the round-1 kernel, kept selectable (SNB200_GEN_CONV_STACK_V1) until the transposed-GEMM kernel has been validated on hardware

constexpr int kCsThreads = 256;
constexpr int kCsM = 128;
constexpr int kCsMaxLayers = SNB200_MAX_CONV_LAYERS;
constexpr int kCsSlots = 2;           // tiles per CTA
constexpr int kCsRegion = 128;        // TMEM columns per (slot, parity) region
constexpr int kCsProducers = 512;     // 16 producer warps = 2 groups of 8; group (warp >> 3) prepares the K chunks of its parity
constexpr int kCsGroup = 256;
constexpr int kCsThreadsAll = kCsProducers + 32;

struct CsLayer {
    int c_in, c_out;
    const float *weight, *bias;
    // BatchNorm (+ReLU) applied to THIS layer's output when it is consumed by the next layer / the pool
    const float *gamma, *beta, *run_mean, *run_var;
    float eps;
    int has_bn, relu;
    double *stats;                      // [2][c_out] sum, sumsq (training) -- written here, read by the next layer and the head
};

struct CsParams {
    const float *x; int layout;
    int b, n, tiles, tiles_per_cloud;
    int num_layers;                     // including layer 1
    CsLayer L[kCsMaxLayers];
    int training;
    double *mom;                        // [9] input moments (zeroed by the caller)
    unsigned *barrier;                  // grid barrier counter (zeroed by the caller)
    float *tile_max, *tile_min;         // (tiles, c_last)
    int fuse_head;                      // run the pool + FC head as the tail of this launch
    HeadParams H;
    // self-cleaning workspace (SNB200_GEN_WORKSPACE_PRIMED): the caller guarantees moments / barrier word / exit word are zero; the
    // kernel zeroes [clean_ptr, clean_ptr + clean_bytes) itself before its first grid barrier and leaves the three words zero again
    int self_clean;
    char *clean_ptr;
    unsigned clean_bytes;
};

// ---- tcgen05 helpers (same encodings as encoder_tc.cu, validated against fp64 in tests/test_gpu_parity.py::test_tc_gemm_3xtf32)
__device__ __forceinline__ void cs_tmem_alloc(uint32_t *smem_dst, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void cs_tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void cs_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void cs_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void cs_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void cs_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cs_ld16(uint32_t taddr, float *v)
{
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void cs_ld16_issue(uint32_t taddr, uint32_t *r)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void cs_ld8_issue(uint32_t taddr, uint32_t *r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void cs_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void cs_ld32(uint32_t taddr, float *v)
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}
__host__ __device__ constexpr uint32_t cs_idesc(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
constexpr unsigned kCsDescHi = (64u) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t cs_sdesc(uint32_t smem_addr)
{
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t)kCsDescHi << 32);
}
__device__ __forceinline__ uint32_t cs_sw128(int row, int chunk) { return (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4); }
__device__ __forceinline__ void cs_split_store(unsigned char *hi_base, unsigned char *lo_base, uint32_t off, float4 v)
{
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
    *reinterpret_cast<float4 *>(hi_base + off) = h;
    *reinterpret_cast<float4 *>(lo_base + off) = l;
}

// bounded waits: a protocol bug must not hang the GPU box -- trap instead (surfaces as a launch failure in the next API call)
__device__ __forceinline__ void cs_mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    for (unsigned spin = 0; !done; spin++) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (spin > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void cs_grid_arrive(unsigned *counter)
{
    // release is cumulative over everything ordered before it by the preceding CTA barrier (the other threads' statistics atomics)
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
}
__device__ __forceinline__ void cs_grid_wait(unsigned *counter, unsigned target)
{
    unsigned v, spin = 0;
    do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        if (++spin > (1u << 26)) __trap();
    } while (v < target);
}
__device__ __forceinline__ void cs_grid_barrier(unsigned *counter, unsigned target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned v, spin = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if (++spin > (1u << 26)) __trap();
        } while (v < target);
    }
    __syncthreads();
}

// Self-validating exchange words: the exchange buffers are zeroed by the launch's memset and a producer never stores the bit
// pattern 0 (+0.0f travels as -0.0f, which is the same number to every consumer), so "word != 0" means "value present": a
// 4-byte store is atomic, a consumer spins on the data itself, and a value is usable one L2 round trip after it was stored -- no
// fence, no flag word, no grid barrier.  Loads bypass L1 (volatile).
__device__ __forceinline__ void cs_xchg_store(float *p, float v)
{
    unsigned u = __float_as_uint(v);
    if (u == 0u) u = 0x80000000u;
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(u) : "memory");
}
__device__ __forceinline__ unsigned cs_xchg_load1(const float *p)
{
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 cs_xchg_load4(const float *p)   // four consecutive words, 16-byte aligned
{
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

// Column reduction of a 32x32 block held one row per lane: after the 5 halving steps lane i holds the combined value of
// column i (31 shuffles instead of 160).  OP: 0 = sum, 1 = max, 2 = min.
template <int OP>
__device__ __forceinline__ float cs_colreduce(float *s, int lane)
{
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int j = 0; j < half; j++) {
            const float send = up ? s[j] : s[j + half];
            const float keep = up ? s[j + half] : s[j];
            const float recv = __shfl_xor_sync(kFullMask, send, half);
            s[j] = OP == 0 ? keep + recv : (OP == 1 ? fmaxf(keep, recv) : fminf(keep, recv));
        }
    }
    return s[0];
}

// Split + swizzled staging of one layer's whole weight matrix (all K) into shared memory, by the 256 producer threads.
// Only legal once every MMA that reads the previous layer's weights has completed.
__device__ __forceinline__ void cs_stage_weights(const CsLayer &Lc, unsigned char *sWhi, float *sBias, int tid)
{
    const int K = Lc.c_in, N = Lc.c_out;
    const int npad = N <= 64 ? 64 : 128;
    const uint32_t atomB = (uint32_t)npad * 128u;
    unsigned char *sWlo = sWhi + (size_t)(K >> 5) * atomB;
    const int sh4 = (K == 32) ? 3 : (K == 64 ? 4 : 5);   // log2(K / 4); K is 32, 64 or 128 on this path
    const int q4m = (1 << sh4) - 1, total = npad << sh4;
    for (int e0 = tid; e0 < total; e0 += kCsProducers * 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int e = e0 + u * kCsProducers;
            const int nrow = e >> sh4, kq = e & q4m;
            v[u] = (e < total && nrow < N) ? __ldg(reinterpret_cast<const float4 *>(Lc.weight + (size_t)nrow * K) + kq) : make_float4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int e = e0 + u * kCsProducers;
            if (e < total) {
                const int nrow = e >> sh4, kq = e & q4m;
                cs_split_store(sWhi, sWlo, (uint32_t)(kq >> 3) * atomB + cs_sw128(nrow, kq & 7), v[u]);
            }
        }
    }
    for (int c = tid; c < npad; c += kCsProducers) sBias[c] = (c < N && Lc.bias) ? __ldg(Lc.bias + c) : 0.f;
    fence_proxy_async();   // generic-proxy writes -> visible to the tensor core
}

// 32 rows x 16 columns held one row per lane: after the five steps lanes 2c and 2c+1 both hold the combined value of column c.
template <int OP>
__device__ __forceinline__ float cs_colreduce16(float *s, int lane)
{
#pragma unroll
    for (int half = 8; half >= 1; half >>= 1) {
        const bool up = (lane & (half * 2)) != 0;
#pragma unroll
        for (int j = 0; j < half; j++) {
            const float send = up ? s[j] : s[j + half];
            const float keep = up ? s[j + half] : s[j];
            const float recv = __shfl_xor_sync(kFullMask, send, half * 2);
            s[j] = OP == 0 ? keep + recv : (OP == 1 ? fmaxf(keep, recv) : fminf(keep, recv));
        }
    }
    const float o = __shfl_xor_sync(kFullMask, s[0], 1);
    return OP == 0 ? s[0] + o : (OP == 1 ? fmaxf(s[0], o) : fminf(s[0], o));
}

// 8 weight rows (output channels cb..cb+nch-1) of an FC layer into shared memory, row-major as in HBM
__device__ __forceinline__ void cs_head_stage_weights(const HeadLayer &L, int cb, int nch, float *s_wh, int tid, bool producer)
{
    if (!producer) return;
    const int c_in = L.c_in;
    if ((c_in & 3) == 0) {
        const int q4 = c_in >> 2, total = 8 * q4;
        for (int e0 = tid; e0 < total; e0 += kCsProducers * 4) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int e = e0 + u * kCsProducers;
                const int jr = e / q4, kq = e - jr * q4;
                v[u] = (e < total && jr < nch) ? __ldg(reinterpret_cast<const float4 *>(L.weight + (size_t)(cb + jr) * c_in) + kq) : make_float4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int e = e0 + u * kCsProducers;
                if (e < total) *reinterpret_cast<float4 *>(s_wh + (size_t)e * 4) = v[u];
            }
        }
    } else {
        for (int e = tid; e < 8 * c_in; e += kCsProducers) {
            const int jr = e / c_in, k = e - jr * c_in;
            s_wh[e] = (jr < nch) ? __ldg(L.weight + (size_t)(cb + jr) * c_in + k) : 0.f;
        }
    }
}

__device__ long long g_cs_ts[64];
#define CS_TS(i) do { if (blockIdx.x == 0 && threadIdx.x == 0 && (i) < 64) g_cs_ts[(i)] = clock64(); } while (0)

__device__ __forceinline__ void cs_named_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void cs_mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Thread roles: warps 0..7 (256 threads) are PRODUCERS (operand preparation, weight staging, epilogues); warp 8 is the MMA
// ISSUER (lane 0 issues tcgen05.mma / tcgen05.commit, the warp only waits on mbarriers).  Producers and issuer walk the
// same (layer, slot, chunk) sequence and meet through mbarriers, never through __syncthreads inside the main loop:
//   bar_full[rb]  producers -> issuer : ring buffer rb holds a prepared 32-wide K chunk   (256 arrivals)
//   bar_ring[rb]  tensor core -> producers : the MMAs that read ring buffer rb have completed (tcgen05.commit)
//   bar_acc[s]    tensor core -> producers : every MMA of tile slot s of this layer has completed

__global__ void __launch_bounds__(kCsThreadsAll, 1) conv_stack_kernel(const __grid_constant__ CsParams P)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // shared-memory map (dynamic): [A ring: 2 x (hi 16 KB + lo 16 KB)] [W hi | W lo : 2 x c_in x c_out x 4 B]
    unsigned char *sA[2][2];  // [ring][hi/lo]
    sA[0][0] = smem_raw;             sA[0][1] = smem_raw + 16384;
    sA[1][0] = smem_raw + 32768;     sA[1][1] = smem_raw + 49152;
    unsigned char *sWhi = smem_raw + 65536;
    __shared__ __align__(16) float sScale[128];
    __shared__ __align__(16) float sShift[128];
    __shared__ float sBias[128];
    __shared__ float sX[kCsSlots][kCsM * 3];
    __shared__ float sW1[128 * 3], sB1[128];
    __shared__ float sRedA[4][128], sRedB[4][128];
    __shared__ uint64_t bar_full[2], bar_ring[2], bar_acc[kCsSlots];
    __shared__ uint32_t tmem_base_smem;
    __shared__ double sMom[9];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool producer = warp < 16;
    const int q = warp & 3, hsel = (warp >> 2) & 3;     // TMEM lane quarter, column group (epilogue: 4 groups of 16 columns)
    const int grp = (warp >> 3) & 1, hs2 = (warp >> 2) & 1; // main loop: producer group, column half inside the chunk
    const int row = q * 32 + lane;                      // the point row this producer thread owns in every tile
    const int G = gridDim.x;
    int tile_of[kCsSlots], np_of[kCsSlots];
    int nslots = 0;
#pragma unroll
    for (int s = 0; s < kCsSlots; s++) {
        const int t = blockIdx.x + s * G;
        tile_of[s] = t;
        np_of[s] = 0;
        if (t < P.tiles) {
            nslots = s + 1;
            const int p0 = (t % P.tiles_per_cloud) * kCsM;
            np_of[s] = min(kCsM, P.n - p0);
        }
    }

    CS_TS(0);
    if (warp == 16) cs_tmem_alloc(&tmem_base_smem, 512);
    if (tid == 0) {
        mbar_init(&bar_full[0], kCsGroup); mbar_init(&bar_full[1], kCsGroup);
        mbar_init(&bar_ring[0], 1); mbar_init(&bar_ring[1], 1);
        mbar_init(&bar_acc[0], 1); mbar_init(&bar_acc[1], 1);
        fence_mbar_init();
    }
    if (tid < 9) sMom[tid] = 0.0;
    // the cloud tiles of this CTA and layer 1's weights
    const CsLayer &L1 = P.L[0];
    if (producer) {
        for (int s = 0; s < nslots; s++) {
            const int t = tile_of[s], cloud = t / P.tiles_per_cloud, p0 = (t % P.tiles_per_cloud) * kCsM;
            const float *xc = P.x + (size_t)cloud * P.n * 3;
            if (P.layout == SNB200_BNC) {
                const float *src = xc + (size_t)p0 * 3;
                const int nf = np_of[s] * 3;
                for (int e = tid; e < kCsM * 3; e += kCsProducers) sX[s][e] = (e < nf) ? __ldg(src + e) : 0.f;
            } else {
                for (int e = tid; e < kCsM * 3; e += kCsProducers) {
                    const int c = e / kCsM, r = e % kCsM;    // coalesced along points
                    sX[s][r * 3 + c] = (r < np_of[s]) ? __ldg(xc + (size_t)c * P.n + p0 + r) : 0.f;
                }
            }
        }
        for (int e = tid; e < L1.c_out * 3; e += kCsProducers) sW1[e] = __ldg(L1.weight + e);
        for (int e = tid; e < L1.c_out; e += kCsProducers) sB1[e] = L1.bias ? __ldg(L1.bias + e) : 0.f;
    }
    cs_fence_before();
    __syncthreads();
    cs_fence_after();
    const uint32_t tmem0 = tmem_base_smem;
    unsigned barrier_epoch = 0;
    const double cnt = (double)P.b * (double)P.n, inv_cnt = 1.0 / cnt;
    CS_TS(1);
    const bool need_stats = P.training != 0;
    if (P.self_clean) {   // statistics accumulators and FC exchange words: zero before anybody adds to them (ordered by the first grid barrier)
        float4 *z = reinterpret_cast<float4 *>(P.clean_ptr);
        const unsigned n16 = P.clean_bytes >> 4;
        for (unsigned e = blockIdx.x * kCsThreadsAll + tid; e < n16; e += G * kCsThreadsAll) z[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!(need_stats && L1.has_bn)) cs_grid_barrier(P.barrier, ++barrier_epoch * G);   // (no phase-0 barrier on this path)
    }

    // ---- phase 0: input moments (training + BN after layer 1): 9 sums over this CTA's points, fp64 atomics, grid barrier
    if (need_stats && L1.has_bn) {
        if (producer) {
            float a9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            if (tid < kCsM * nslots) {
                const int s = tid / kCsM, r = tid % kCsM;
                if (r < np_of[s]) {
                    const float px = sX[s][r * 3 + 0], py = sX[s][r * 3 + 1], pz = sX[s][r * 3 + 2];
                    a9[0] = px; a9[1] = py; a9[2] = pz;
                    a9[3] = px * px; a9[4] = px * py; a9[5] = px * pz; a9[6] = py * py; a9[7] = py * pz; a9[8] = pz * pz;
                }
            }
#pragma unroll
            for (int j = 0; j < 9; j++) {
                float v = a9[j];
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
                if (lane == 0) atomicAdd(&sMom[j], (double)v);
            }
        }
        __syncthreads();
        if (tid < 9) atomicAdd(P.mom + tid, sMom[tid]);
        __syncthreads();
        if (tid == 0) cs_grid_arrive(P.barrier);
        if (producer) cs_stage_weights(P.L[1], sWhi, sBias, tid);     // overlaps the barrier latency
        if (tid == 0) cs_grid_wait(P.barrier, ++barrier_epoch * G);
        __syncthreads();
        if (tid < 9) sMom[tid] = __ldcg(P.mom + tid);
        __syncthreads();
    } else {
        if (producer) cs_stage_weights(P.L[1], sWhi, sBias, tid);
        __syncthreads();
    }
    CS_TS(2);

    uint32_t g = 0;                     // global chunk counter (same sequence in producers and issuer)
    uint32_t acc_phase[kCsSlots] = {0, 0};
    int parity = 0;                     // TMEM region parity holding the CURRENT layer's input (previous layer's raw output)

    for (int l = 1; l < P.num_layers; l++) {
        const CsLayer &Lp = P.L[l - 1];   // producer of this layer's input (its BN+ReLU is applied on load)
        const CsLayer &Lc = P.L[l];
        const int K = Lc.c_in, N = Lc.c_out;
        const int npad = N <= 64 ? 64 : 128;
        const uint32_t idesc = cs_idesc(kCsM, npad);
        const uint32_t atomB = (uint32_t)npad * 128u;
        const int nchunks = K >> 5;
        unsigned char *sWlo = sWhi + (size_t)nchunks * atomB;
        const bool last = (l == P.num_layers - 1);
        const bool want_stats = need_stats && Lc.has_bn;
        const uint32_t in_region = (uint32_t)(parity * kCsRegion), out_region = (uint32_t)((parity ^ 1) * kCsRegion);

        if (!producer) {
            // =============================== MMA issuer warp ===============================
            for (int s = 0; s < nslots; s++) {
                const uint32_t t_out = tmem0 + (uint32_t)(s * 2 * kCsRegion) + out_region;
                for (int kc = 0; kc < nchunks; kc++, g++) {
                    const int rb = g & 1;
                    cs_mbar_wait(&bar_full[rb], (g >> 1) & 1);
                    cs_fence_after();
                    if (lane == 0) {
                        const uint64_t a_hi = cs_sdesc(smem_u32(sA[rb][0])), a_lo = cs_sdesc(smem_u32(sA[rb][1]));
                        const uint64_t b_hi = cs_sdesc(smem_u32(sWhi) + (uint32_t)kc * atomB), b_lo = cs_sdesc(smem_u32(sWlo) + (uint32_t)kc * atomB);
#pragma unroll
                        for (int ks = 0; ks < 4; ks++) {   // 32 bytes (K = 8 tf32) per step: +2 in the 16-byte address field
                            const uint64_t o = (uint64_t)(ks * 2);
                            cs_umma(t_out, a_lo + o, b_hi + o, idesc, (kc > 0 || ks > 0) ? 1u : 0u);
                            cs_umma(t_out, a_hi + o, b_lo + o, idesc, 1u);
                            cs_umma(t_out, a_hi + o, b_hi + o, idesc, 1u);
                        }
                        cs_commit(&bar_ring[rb]);
                        if (kc == nchunks - 1) cs_commit(&bar_acc[s]);
                    }
                    __syncwarp();
                }
            }
        } else {
            // =============================== producer warps ===============================
            CS_TS(3 + (l - 1) * 8 + 0);
            CS_TS(3 + (l - 1) * 8 + 1);
            // BatchNorm (+ReLU) of the producer layer as a per-channel affine map
            for (int c = tid; c < K; c += kCsProducers) {
                float sc = 1.f, sh = 0.f;
                if (Lp.has_bn) {
                    float mean, var;
                    if (P.training) {
                        double m, v;
                        if (l == 1) {   // analytic statistics of layer 1 from the input moments
                            const double mx = sMom[0] * inv_cnt, my = sMom[1] * inv_cnt, mz = sMom[2] * inv_cnt;
                            const double cxx = sMom[3] * inv_cnt - mx * mx, cxy = sMom[4] * inv_cnt - mx * my, cxz = sMom[5] * inv_cnt - mx * mz;
                            const double cyy = sMom[6] * inv_cnt - my * my, cyz = sMom[7] * inv_cnt - my * mz, czz = sMom[8] * inv_cnt - mz * mz;
                            const double a0 = sW1[c * 3 + 0], a1 = sW1[c * 3 + 1], a2 = sW1[c * 3 + 2];
                            m = a0 * mx + a1 * my + a2 * mz + (double)sB1[c];
                            v = a0 * a0 * cxx + a1 * a1 * cyy + a2 * a2 * czz + 2.0 * (a0 * a1 * cxy + a0 * a2 * cxz + a1 * a2 * cyz);
                            if (v < 0) v = 0;
                            if (blockIdx.x == 0) {   // the (sum, sumsq) form every consumer of the statistics uses
                                Lp.stats[c] = cnt * m;
                                Lp.stats[K + c] = cnt * (v + m * m);
                            }
                        } else {
                            m = __ldcg(Lp.stats + c) * inv_cnt;
                            v = __ldcg(Lp.stats + K + c) * inv_cnt - m * m;
                            if (v < 0) v = 0;
                        }
                        mean = (float)m; var = (float)v;
                    } else {
                        mean = Lp.run_mean[c]; var = Lp.run_var[c];
                    }
                    const float invstd = 1.0f / sqrtf(var + Lp.eps);
                    sc = Lp.gamma[c] * invstd;
                    sh = Lp.beta[c] - mean * sc;
                }
                // tensor memory holds W.a WITHOUT the producer's bias (layer 1 is evaluated with its bias): fold it into the shift
                if (l >= 2 && Lp.bias) sh = fmaf(Lp.bias[c], sc, sh);
                sScale[c] = sc;
                sShift[c] = sh;
            }
            cs_named_sync(1, kCsProducers);
            CS_TS(3 + (l - 1) * 8 + 2);

            // ---- operand preparation: chunk g+1 is prepared while the tensor core works on chunk g; inside a thread the
            //      tensor-memory load of the NEXT chunk is in flight while the current chunk is normalised, split and stored
            {
                // Producer group `grp` prepares the chunks whose global index has its parity (= ring buffer grp): while one group
                // is inside its fence / arrive latency chain the other one is already normalising the next chunk.
                const int total_chunks = nslots * nchunks;
                const uint32_t g0 = g;
                for (int ci = 0; ci < total_chunks; ci++) {
                    const uint32_t gg = g0 + (uint32_t)ci;
                    if ((int)(gg & 1) != grp) continue;
                    const int s = ci / nchunks, kc = ci - s * nchunks;
                    const int np = (s == 0) ? np_of[0] : np_of[1];
                    const int kb = kc * 32 + hs2 * 16;
                    float v[16];
                    if (l == 1) {
                        const float *xr = (s == 0 ? sX[0] : sX[1]) + row * 3;
                        const float px = xr[0], py = xr[1], pz = xr[2];
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            const int c = kb + j;
                            v[j] = fmaf(sW1[c * 3 + 2], pz, fmaf(sW1[c * 3 + 1], py, sW1[c * 3 + 0] * px)) + sB1[c];
                        }
                    } else {
                        cs_ld16(tmem0 + (uint32_t)(s * 2 * kCsRegion) + in_region + ((uint32_t)(q * 32) << 16) + (uint32_t)kb, v);
                    }
                    const bool pv = row < np;
                    float4 tq[4];
#pragma unroll
                    for (int c4 = 0; c4 < 4; c4++) {
                        const float4 sc = *reinterpret_cast<const float4 *>(sScale + kb + c4 * 4);
                        const float4 sh = *reinterpret_cast<const float4 *>(sShift + kb + c4 * 4);
                        float4 t;
                        t.x = fmaf(v[c4 * 4 + 0], sc.x, sh.x); t.y = fmaf(v[c4 * 4 + 1], sc.y, sh.y);
                        t.z = fmaf(v[c4 * 4 + 2], sc.z, sh.z); t.w = fmaf(v[c4 * 4 + 3], sc.w, sh.w);
                        if (Lp.relu) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
                        if (!pv) t = make_float4(0.f, 0.f, 0.f, 0.f);
                        tq[c4] = t;
                    }
                    if (gg >= 2) {   // the MMAs of chunk gg-2 (same ring buffer) must have completed before its operands are overwritten
                        cs_mbar_wait(&bar_ring[grp], ((gg >> 1) - 1) & 1);
                        cs_fence_after();
                    }
#pragma unroll
                    for (int c4 = 0; c4 < 4; c4++) cs_split_store(sA[grp][0], sA[grp][1], cs_sw128(row, hs2 * 4 + c4), tq[c4]);
                    cs_fence_before();
                    fence_proxy_async();
                    cs_mbar_arrive(&bar_full[grp]);
                }
                g = g0 + (uint32_t)total_chunks;
            }
            CS_TS(3 + (l - 1) * 8 + 3);

            // ---- epilogue: wait for the accumulators, then batch statistics and / or per-tile extrema
            for (int s = 0; s < nslots; s++) {
                cs_mbar_wait(&bar_acc[s], acc_phase[s]);
                acc_phase[s] ^= 1;
            }
            cs_fence_after();
            CS_TS(3 + (l - 1) * 8 + 4);
            if (want_stats) {   // sums over BOTH tile slots first (same columns, different rows), one reduction per 16-column block
                for (int cb = hsel * 16; cb < npad; cb += 64) {
                    float v[16], w[16], t[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) { v[j] = 0.f; w[j] = 0.f; }
                    for (int s = 0; s < nslots; s++) {
                        const uint32_t t_out = tmem0 + (uint32_t)(s * 2 * kCsRegion) + out_region;
                        cs_ld16(t_out + ((uint32_t)(q * 32) << 16) + (uint32_t)cb, t);
                        if (row < (s == 0 ? np_of[0] : np_of[1])) {
#pragma unroll
                            for (int j = 0; j < 16; j++) {
                                const float u = t[j] + sBias[cb + j];
                                v[j] += u;
                                w[j] = fmaf(u, u, w[j]);
                            }
                        }
                    }
                    const float sm = cs_colreduce16<0>(v, lane);
                    const float sq = cs_colreduce16<0>(w, lane);
                    if (!(lane & 1)) { sRedA[q][cb + (lane >> 1)] = sm; sRedB[q][cb + (lane >> 1)] = sq; }
                }
                cs_named_sync(1, kCsProducers);
                if (tid < N) {
                    const float sm = (sRedA[0][tid] + sRedA[1][tid]) + (sRedA[2][tid] + sRedA[3][tid]);
                    const float sq = (sRedB[0][tid] + sRedB[1][tid]) + (sRedB[2][tid] + sRedB[3][tid]);
                    atomicAdd(Lc.stats + tid, (double)sm);
                    atomicAdd(Lc.stats + N + tid, (double)sq);
                }
                cs_named_sync(1, kCsProducers);
            }
            if (last) {   // extrema for the max-pool, per tile
                for (int s = 0; s < nslots; s++) {
                    const bool pv = row < (s == 0 ? np_of[0] : np_of[1]);
                    const uint32_t t_out = tmem0 + (uint32_t)(s * 2 * kCsRegion) + out_region;
                    for (int cb = hsel * 16; cb < npad; cb += 64) {
                        float v[16], w[16];
                        cs_ld16(t_out + ((uint32_t)(q * 32) << 16) + (uint32_t)cb, v);
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            const float u = v[j] + sBias[cb + j];
                            v[j] = pv ? u : -INFINITY;
                            w[j] = pv ? u : INFINITY;
                        }
                        const float mx = cs_colreduce16<1>(v, lane);
                        const float mn = cs_colreduce16<2>(w, lane);
                        if (!(lane & 1)) { sRedA[q][cb + (lane >> 1)] = mx; sRedB[q][cb + (lane >> 1)] = mn; }
                    }
                    cs_named_sync(1, kCsProducers);
                    if (tid < N) {
                        const int tile = (s == 0) ? tile_of[0] : tile_of[1];
                        P.tile_max[(size_t)tile * N + tid] = fmaxf(fmaxf(sRedA[0][tid], sRedA[1][tid]), fmaxf(sRedA[2][tid], sRedA[3][tid]));
                        P.tile_min[(size_t)tile * N + tid] = fminf(fminf(sRedB[0][tid], sRedB[1][tid]), fminf(sRedB[2][tid], sRedB[3][tid]));
                    }
                    cs_named_sync(1, kCsProducers);
                }
            }
            CS_TS(3 + (l - 1) * 8 + 5);
        }
        parity ^= 1;
        cs_fence_before();
        __syncthreads();                                         // this CTA's statistics atomics are issued, its MMAs are done
        if (want_stats && tid == 0) cs_grid_arrive(P.barrier);
        if (producer && l + 1 < P.num_layers) cs_stage_weights(P.L[l + 1], sWhi, sBias, tid);   // overlaps the barrier latency
        if (want_stats && tid == 0) cs_grid_wait(P.barrier, (barrier_epoch + 1) * G);   // every tile's statistics are in
        if (want_stats) barrier_epoch++;
        __syncthreads();
        cs_fence_after();
        CS_TS(3 + (l - 1) * 8 + 6);
    }

    // every commit has been observed through bar_acc; release tensor memory
    cs_fence_before();
    __syncthreads();
    if (warp == 16) cs_tmem_dealloc(tmem0, 512);

    // ================================================================================================================
    // Fused tail: max-pool finalise + FC head (samplenet.py:97-104) on the CTAs of the grid.  Each FC layer's output channels
    // are spread over the CTAs, 8 per CTA (BatchNorm over the batch stays inside one warp: lane = batch row); activations go
    // through a few-KB global scratch that lives in L2 as self-validating words (cs_xchg_*), so the layers need no barrier.
    // Versus the 16-CTA cluster kernel this removes a launch and spreads each layer's latency chain over more SMs.
    // ================================================================================================================
    // Programmatic dependent launch: a kernel queued behind this one with the PDL attribute (the fused tail) may be scheduled onto
    // SMs as this grid's CTAs exit; it synchronises on this grid's completion itself (griddepcontrol.wait) before touching our output.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (!P.fuse_head) return;
    const HeadParams &H = P.H;
    // shared memory of the head (the conv stack's buffers are dead): input row group | partial sums | first weight rows of every layer
    float *s_in = reinterpret_cast<float *>(smem_raw);                 // [32 rows][c_in + 1] one row group of the input
    int hcmax = H.c_feat, hcsum = 0;
    for (int l = 0; l < H.num_fc; l++) { hcmax = max(hcmax, H.fc[l].c_in); hcsum += H.fc[l].c_in; }
    float *s_part = reinterpret_cast<float *>(smem_raw) + (size_t)hcmax * 33;   // [8 K slices][8 channels][32 rows]
    float *s_wall = s_part + 8 * 8 * 32;                                       // per layer [8 channels][c_in] weight rows
    __shared__ uint64_t hbar[SNB200_MAX_FC_LAYERS];
    const double inv_cnt_h = 1.0 / H.count;
    const float inv_b = 1.0f / (float)H.b;
    CS_TS(36);
    // ---- weights do not depend on activations: the first 8-channel group of EVERY layer is fetched now, one TMA bulk copy per
    //      layer (the 8 rows are contiguous in HBM), completion on one mbarrier per layer; nobody touches them before the layer's math
    if (tid == 0) {
        for (int l = 0; l < H.num_fc; l++) mbar_init(&hbar[l], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        fence_proxy_async();   // the smem region was written through the generic proxy by the conv stack
        int woff = 0;
        for (int l = 0; l < H.num_fc; l++) {
            const HeadLayer &L = H.fc[l];
            const int cpc = max(8, (((L.c_out + G - 1) / G + 7) / 8) * 8);
            const int lo = blockIdx.x * cpc, hi = min(L.c_out, lo + cpc);
            const bool tma_ok = (L.c_in & 3) == 0 && (hcmax & 3) == 0 && (reinterpret_cast<uintptr_t>(L.weight) & 15) == 0;
            if (lo < hi && tma_ok) {
                const uint32_t bytes = (uint32_t)min(8, hi - lo) * L.c_in * 4u;
                mbar_expect_tx(&hbar[l], bytes);
                tma_load_1d(s_wall + woff, L.weight + (size_t)lo * L.c_in, bytes, &hbar[l]);
            }
            woff += 8 * L.c_in;
        }
    }
    // In training mode the last layer's statistics barrier already ordered every CTA's extrema before this point; in eval mode
    // no grid barrier has been crossed yet.
    if (!(need_stats && P.L[P.num_layers - 1].has_bn)) cs_grid_barrier(P.barrier, ++barrier_epoch * G);
    // From here on CTAs exchange activations point to point through self-validating words (cs_xchg_*): consumers spin on the data
    // itself -- no fence, no flag word, no grid barrier.  The exchange buffers are zeroed by the launch's memset.
    //   stage 0 = the pooled feature, stage l+1 = the output of FC layer l
    // ---- phase P: pooled feature, spread over the grid
    {
        const int total = H.b * H.c_feat;
        const int gt = blockIdx.x * kCsThreadsAll + tid, gn = G * kCsThreadsAll;
        float *ll0 = H.ll[0];
        for (int e = gt; e < total; e += gn) {
            const int bi = e / H.c_feat, c = e % H.c_feat;
            float mx = -INFINITY, mn = INFINITY;
            const float *tm = H.tile_max + (size_t)bi * H.tiles_per_cloud * H.c_feat + c;
            const float *tn = H.tile_min + (size_t)bi * H.tiles_per_cloud * H.c_feat + c;
            // every load of this element is issued before the first use
            const double st0 = (H.last_has_bn && H.training) ? __ldcg(H.last_stats + c) : 0.0;
            const double st1 = (H.last_has_bn && H.training) ? __ldcg(H.last_stats + H.c_feat + c) : 0.0;
            const float lg = H.last_has_bn ? __ldg(H.last_gamma + c) : 1.f, lb = H.last_has_bn ? __ldg(H.last_beta + c) : 0.f;
#pragma unroll 8
            for (int t = 0; t < H.tiles_per_cloud; t++) {
                mx = fmaxf(mx, __ldcg(tm + (size_t)t * H.c_feat));
                mn = fminf(mn, __ldcg(tn + (size_t)t * H.c_feat));
            }
            float v = mx;
            if (H.last_has_bn) {
                float mean, var;
                if (H.training) {
                    const double m = st0 * inv_cnt_h;
                    double vv = st1 * inv_cnt_h - m * m;
                    if (vv < 0) vv = 0;
                    mean = (float)m; var = (float)vv;
                } else {
                    mean = H.last_run_mean[c]; var = H.last_run_var[c];
                }
                const float sc = lg * (1.0f / sqrtf(var + H.last_eps));
                const float sh = lb - mean * sc;
                v = sc >= 0.f ? fmaf(mx, sc, sh) : fmaf(mn, sc, sh);
            }
            if (H.last_relu) v = fmaxf(v, 0.f);
            cs_xchg_store(ll0 + e, v);
            H.feat[e] = v;
        }
    }
    CS_TS(37);
    CS_TS(38);

    int woff = 0;
    for (int l = 0; l < H.num_fc; l++) {
        const HeadLayer &L = H.fc[l];
        const bool lastfc = (l == H.num_fc - 1);
        float *dst = lastfc ? H.out : H.act[l & 1];
        float *lldst = lastfc ? nullptr : H.ll[l + 1];
        const float *llsrc = H.ll[l];
        const int c_in = L.c_in;
        float *s_wh = s_wall + woff;
        woff += 8 * c_in;
        // 8 output channels per pass and per CTA: few enough CTAs read the (shared) input that L2 does not serialise on it
        const int cpc = max(8, (((L.c_out + G - 1) / G + 7) / 8) * 8);
        const int c_lo = blockIdx.x * cpc, c_hi = min(L.c_out, c_lo + cpc);
        const int nrg = (H.b + 31) >> 5;
        const bool w_tma = (c_in & 3) == 0 && (hcmax & 3) == 0 && (reinterpret_cast<uintptr_t>(L.weight) & 15) == 0;
        CS_TS(39 + l * 6 + 0);
        for (int cb = c_lo; cb < c_hi; cb += 8) {                     // one group of 8 channels at a time
            const int nch = min(8, c_hi - cb);
            // per-channel parameters of the channel this warp will finish (warps 0..7): loads start now
            const int cw = cb + (warp & 7);
            const bool cvw = warp < 8 && (warp & 7) < nch;
            const float pbias = (cvw && L.bias) ? __ldg(L.bias + cw) : 0.f;
            const float pgam = (cvw && L.has_bn) ? __ldg(L.gamma + cw) : 1.f;
            const float pbet = (cvw && L.has_bn) ? __ldg(L.beta + cw) : 0.f;
            const float prm = (cvw && L.has_bn && L.run_mean) ? L.run_mean[cw] : 0.f;
            const float prv = (cvw && L.has_bn && L.run_var) ? L.run_var[cw] : 1.f;
            if (cb != c_lo || !w_tma) {   // (the first group of every layer was fetched by TMA at the start of the head)
                __syncthreads();
                cs_head_stage_weights(L, cb, nch, s_wh, tid, producer);
            }
            float yv[8];                                              // finished pre-activation: row group g, lane = row, warp = channel
#pragma unroll
            for (int gq = 0; gq < 8; gq++) yv[gq] = 0.f;
#pragma unroll
            for (int gq = 0; gq < 8; gq++) {
                if (gq < nrg) {
                    const int r0 = gq * 32, rn = min(32, H.b - r0);
                    if (gq > 0 || cb != c_lo || l > 0) __syncthreads();   // the previous user of s_in / s_part is done
                    if (producer) {   // stage rows r0..r0+rn-1 row-major with an odd row stride (conflict-free lane = row reads).  Lanes run
                                      // along k (coalesced 16-byte loads), a thread's loads are requested together and re-requested
                                      // until every word is present.
                        const int ldi = c_in + 1;
                        if ((c_in & 3) == 0) {
                            const int q4 = c_in >> 2, items = 32 * q4;           // item = (row, 4 channels) = one 16-byte load
                            for (int i0 = tid; i0 < items; i0 += kCsProducers * 4) {
                                uint4 v[4];
                                unsigned spin = 0;
                                bool ok;
                                do {
                                    ok = true;
#pragma unroll
                                    for (int u = 0; u < 4; u++) {
                                        const int i = i0 + u * kCsProducers;
                                        const int r = i / q4, kq = i - r * q4;
                                        if (i < items && r < rn) v[u] = cs_xchg_load4(llsrc + (size_t)(r0 + r) * c_in + 4 * kq);
                                        else v[u] = make_uint4(1u, 1u, 1u, 1u);
                                    }
#pragma unroll
                                    for (int u = 0; u < 4; u++) ok = ok && v[u].x != 0u && v[u].y != 0u && v[u].z != 0u && v[u].w != 0u;
                                    if (++spin > (1u << 24)) __trap();
                                } while (!ok);
#pragma unroll
                                for (int u = 0; u < 4; u++) {
                                    const int i = i0 + u * kCsProducers;
                                    if (i < items) {
                                        const int r = i / q4, kq = i - r * q4;
                                        float *d = s_in + r * ldi + 4 * kq;
                                        const bool live = r < rn;
                                        d[0] = live ? __uint_as_float(v[u].x) : 0.f; d[1] = live ? __uint_as_float(v[u].y) : 0.f;
                                        d[2] = live ? __uint_as_float(v[u].z) : 0.f; d[3] = live ? __uint_as_float(v[u].w) : 0.f;
                                    }
                                }
                            }
                        } else {
                            for (int e = tid; e < 32 * c_in; e += kCsProducers) {
                                const int r = e / c_in, k = e - r * c_in;
                                float xv = 0.f;
                                if (r < rn) {
                                    unsigned q, spin = 0;
                                    do {
                                        q = cs_xchg_load1(llsrc + (size_t)(r0 + r) * c_in + k);
                                        if (++spin > (1u << 24)) __trap();
                                    } while (q == 0u);
                                    xv = __uint_as_float(q);
                                }
                                s_in[r * ldi + k] = xv;
                            }
                        }
                    }
                    if (cb == c_lo && gq == 0 && w_tma) mbar_wait(&hbar[l], 0);   // this layer's first weight rows have landed
                    __syncthreads();
                    CS_TS(39 + l * 6 + 1);
                    if (producer) {   // warp -> (channel quad = warp & 1, K eighth = warp >> 1); lane = row
                        const int cq = (warp & 1) * 4, k8 = warp >> 1;
                        const int kr = ((c_in + 31) / 32) * 4;            // K per eighth, multiple of 4
                        const int k_lo = min(c_in, k8 * kr), k_hi = min(c_in, k_lo + kr);
                        const float *wq = s_wh + cq * c_in;
                        float a4[4] = {0.f, 0.f, 0.f, 0.f};
                        int k = k_lo;
                        if ((c_in & 3) == 0) {
                            for (; k + 4 <= k_hi; k += 4) {
                                const float *xr = s_in + lane * (c_in + 1) + k;
                                const float x0 = xr[0], x1 = xr[1], x2 = xr[2], x3 = xr[3];
#pragma unroll
                                for (int j = 0; j < 4; j++) {
                                    const float4 wv = *reinterpret_cast<const float4 *>(wq + j * c_in + k);
                                    a4[j] = fmaf(x3, wv.w, fmaf(x2, wv.z, fmaf(x1, wv.y, fmaf(x0, wv.x, a4[j]))));
                                }
                            }
                        }
                        for (; k < k_hi; k++) {
                            const float xv = s_in[lane * (c_in + 1) + k];
#pragma unroll
                            for (int j = 0; j < 4; j++) a4[j] = fmaf(xv, wq[j * c_in + k], a4[j]);
                        }
#pragma unroll
                        for (int j = 0; j < 4; j++) s_part[(k8 * 8 + cq + j) * 32 + lane] = a4[j];
                    }
                    __syncthreads();
                    CS_TS(39 + l * 6 + 2);
                    if (warp < 8)   // fixed-order combination of the 8 K eighths: warp = channel, lane = row
                    {
                        float t = 0.f;
#pragma unroll
                        for (int e8 = 0; e8 < 8; e8++) t += s_part[(e8 * 8 + warp) * 32 + lane];
                        yv[gq] = t;
                    }
                }
            }
            CS_TS(39 + l * 6 + 3);
            if (cvw) {
                float scale = 1.f, shift = 0.f;
#pragma unroll
                for (int gq = 0; gq < 8; gq++) yv[gq] += pbias;
                float bn_mean = 0.f, bn_q = 0.f;
                if (L.has_bn) {
                    float mean, var;
                    if (H.training) {
                        // batch statistics in one shuffle tree: deviations from a pivot sample (row 0), sum and sum of squares reduced
                        // together; var = (S2 - S1^2/n)/n is well conditioned because the pivot lies inside the data
                        const float pivot = __shfl_sync(kFullMask, yv[0], 0);
                        float s1 = 0.f, s2 = 0.f;
#pragma unroll
                        for (int gq = 0; gq < 8; gq++)
                            if (gq * 32 + lane < H.b) { const float d = yv[gq] - pivot; s1 += d; s2 = fmaf(d, d, s2); }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            s1 += __shfl_xor_sync(kFullMask, s1, o);
                            s2 += __shfl_xor_sync(kFullMask, s2, o);
                        }
                        mean = fmaf(s1, inv_b, pivot);
                        bn_q = fmaxf(fmaf(-s1 * inv_b, s1, s2), 0.f);        // sum of squared deviations from the mean
                        var = bn_q * inv_b;
                        bn_mean = mean;
                    } else {
                        mean = prm; var = prv;
                    }
                    const float invstd = rsqrtf(var + L.eps);
                    scale = pgam * invstd;
                    shift = pbet - mean * scale;
                }
                CS_TS(39 + l * 6 + 4);
#pragma unroll
                for (int gq = 0; gq < 8; gq++) {
                    const int r = gq * 32 + lane;
                    if (r < H.b) {
                        float v = L.has_bn ? fmaf(yv[gq], scale, shift) : yv[gq];
                        if (L.relu) v = fmaxf(v, 0.f);
                        if (lastfc) {
                            const int oc = (H.out_inner > 0) ? (cw % H.out_inner) * (L.c_out / H.out_inner) + cw / H.out_inner : cw;
                            dst[(size_t)r * L.c_out + oc] = v;
                        } else {
                            cs_xchg_store(lldst + (size_t)r * L.c_out + cw, v);   // the next layer's consumers spin on these words
                        }
                    }
                }
                if (L.has_bn && H.training && lane == 0) {   // running statistics: off the critical path
                    const float unb = H.b > 1 ? bn_q / (float)(H.b - 1) : bn_q * inv_b;
                    if (L.run_mean) L.run_mean[cw] = (1.f - L.momentum) * prm + L.momentum * bn_mean;
                    if (L.run_var) L.run_var[cw] = (1.f - L.momentum) * prv + L.momentum * unb;
                }
            }
        }
        CS_TS(39 + l * 6 + 5);
    }
    if (blockIdx.x == G - 1 && tid < H.num_counters) *H.counters[tid] += 1;
    // ---- running statistics of the conv stack: off the critical path, taken by the CTAs from the top of the grid (idle in the
    //      last FC layer); training mode never reads these buffers inside the kernel
    if (H.training) {
        const int gt = (G - 1 - (int)blockIdx.x) * kCsThreadsAll + tid, gn = G * kCsThreadsAll;
        int base = 0;
        for (int l = 0; l < H.ru_num; l++) {
            for (int c = gt - base; c < H.ru_c[l]; c += gn) {
                if (c < 0) continue;
                const double m = __ldcg(H.ru_stats[l] + c) * inv_cnt_h;
                double v = __ldcg(H.ru_stats[l] + H.ru_c[l] + c) * inv_cnt_h - m * m;
                if (v < 0) v = 0;
                const double unb = H.count > 1 ? v * (H.count / (H.count - 1)) : v;
                const float mom = H.ru_momentum[l];
                if (H.ru_mean[l]) H.ru_mean[l][c] = (1.f - mom) * H.ru_mean[l][c] + mom * (float)m;
                if (H.ru_var[l]) H.ru_var[l][c] = (1.f - mom) * H.ru_var[l][c] + mom * (float)unb;
            }
            base = (base + H.ru_c[l]) % gn;
        }
    }
    if (P.self_clean) {   // the last CTA to leave puts the moments, the barrier word and the exit word back to zero for the next launch
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            if (atomicAdd(P.barrier + 1, 1u) == G - 1) {
                for (int j = 0; j < 16; j++) P.mom[j] = 0.0;
                P.barrier[0] = 0u;
                P.barrier[1] = 0u;
            }
        }
    }
}

int debug_conv_stack_timestamps_v1(long long *host_out64)
{
    return cudaMemcpyFromSymbol(host_out64, g_cs_ts, sizeof(long long) * 64) == cudaSuccess ? SNB200_OK : SNB200_ECUDA;
}

// ------------------------------------------------------------------------------------------------------------------
bool conv_stack_supported(int b, int n, int nconv, const snb200_layer *conv)
{
    if (nconv < 2 || nconv > kCsMaxLayers || conv[0].c_in != 3) return false;
    if (conv[0].c_out % 32 != 0 || conv[0].c_out > 128) return false;
    size_t wmax = 0;
    for (int l = 1; l < nconv; l++) {
        if ((conv[l].c_in != 32 && conv[l].c_in != 64 && conv[l].c_in != 128) || conv[l].c_out > 128 || conv[l].c_out < 8) return false;
        const size_t npad = conv[l].c_out <= 64 ? 64 : 128;
        wmax = max(wmax, 2 * (size_t)conv[l].c_in * npad * 4);
    }
    if (65536 + wmax > 200 * 1024) return false;
    const long long tiles = (long long)b * ((n + kCsM - 1) / kCsM);
    return tiles <= (long long)kCsSlots * kNumSMs;
}

int launch_conv_stack(int b, int n, int layout, const float *x, int nconv, const snb200_layer *conv, int training, double *const *stats,
                      double *mom, unsigned *barrier, float *tile_max, float *tile_min, int *tiles_per_cloud_out, const HeadParams *head,
                      char *clean_ptr, size_t clean_bytes, cudaStream_t stream)
{
    CsParams P;
    memset(&P, 0, sizeof(P));
    if (head) { P.fuse_head = 1; P.H = *head; }
    if (head && clean_ptr) { P.self_clean = 1; P.clean_ptr = clean_ptr; P.clean_bytes = (unsigned)clean_bytes; }
    P.x = x; P.layout = layout; P.b = b; P.n = n;
    P.tiles_per_cloud = (n + kCsM - 1) / kCsM;
    P.tiles = b * P.tiles_per_cloud;
    P.num_layers = nconv; P.training = training;
    P.mom = mom; P.barrier = barrier; P.tile_max = tile_max; P.tile_min = tile_min;
    size_t wmax = 0;
    for (int l = 0; l < nconv; l++) {
        CsLayer &D = P.L[l];
        D.c_in = conv[l].c_in; D.c_out = conv[l].c_out; D.weight = conv[l].weight; D.bias = conv[l].bias;
        D.gamma = conv[l].bn_weight; D.beta = conv[l].bn_bias; D.run_mean = conv[l].bn_running_mean; D.run_var = conv[l].bn_running_var;
        D.eps = conv[l].bn_eps; D.has_bn = conv[l].bn_weight != nullptr; D.relu = conv[l].relu; D.stats = stats[l];
        if (l >= 1) wmax = max(wmax, 2 * (size_t)conv[l].c_in * (conv[l].c_out <= 64 ? 64 : 128) * 4);
    }
    if (tiles_per_cloud_out) *tiles_per_cloud_out = P.tiles_per_cloud;
    size_t smem = 65536 + wmax + 1024;
    if (head) {   // the fused tail reuses the same dynamic shared memory: input tile + partial sums + 8 weight rows
        int hcmax = head->c_feat;
        for (int l = 0; l < head->num_fc; l++) hcmax = max(hcmax, head->fc[l].c_in);
        size_t hcsum = 0;
        for (int l = 0; l < head->num_fc; l++) hcsum += head->fc[l].c_in;
        const size_t hs = ((size_t)hcmax * 33 + 2048 + (size_t)8 * hcsum) * sizeof(float) + 1024;
        if (hs > 200 * 1024) { set_error("conv stack: FC width %d too large for the fused head", hcmax); return SNB200_EUNSUPPORTED; }
        smem = max(smem, hs);
    }
    static PerDeviceOnce once;
    if (once.first()) cudaFuncSetAttribute(conv_stack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
    int grid = min(P.tiles, kNumSMs);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kCsThreadsAll); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_stack_kernel, P);
    if (e != cudaSuccess) { set_error("conv stack: cooperative launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return SNB200_ECUDA; }
    return check_launch("conv stack");
}

}  // namespace v1
}  // namespace snb
