// xbench.cu -- measurement hook: latency of the per-step all-gather the GRU recurrence needs (every CTA of a 16-CTA
// group publishes rows x 32 bf16 units and needs all 512 units of every row before its next step), for the exchange
// mechanisms considered in DESIGN.md "recurrence".  Not on the product path; numbers are recorded in profiles/.
//   method 0: L2 "LL" words (bf16x2 + 32-bit tag in one 8-byte st.volatile), consumers poll with ld.volatile.v4
//   method 1: same words with st.relaxed.gpu / ld.relaxed.gpu
//   method 2: thread-block cluster, LL words written straight into every peer's shared memory (st.shared::cluster.v2),
//             consumers poll their OWN shared memory
//   method 3: thread-block cluster, st.async (16-byte payload, no tags) + remote mbarrier complete_tx; consumers wait
//             on their own mbarrier
#include "kernels.h"
#include <cooperative_groups.h>
#include <vector>

namespace {
constexpr int XCL = 16, XUN = 32, XH = 512, XNT = 256;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}

struct XP {
    unsigned long long* gbuf;   // methods 0/1: [group][2][rows][256] LL words
    long long* out;             // [blocks][2]: cycles, checksum
    int rows, iters, method;
    int* smids;                 // [blocks] scratch: grouping by physical SM id instead of block id (method + 16)
    int active_groups;          // groups beyond this one exit at once (method + 32: grid padded to 8 groups = 128 CTAs)
    int skip_blocks;            // the first blocks of the grid exit at once (method + 64: 6 = keep the groups off SMs 142-147)
};

__global__ void __launch_bounds__(XNT, 1) k_xbench(const XP P) {
    extern __shared__ __align__(16) unsigned char sm[];
    // cluster methods: [2 parities][rows][512 units] as LL words (method 2: 8 B per bf16x2) or bf16 (method 3)
    unsigned long long* sll = reinterpret_cast<unsigned long long*>(sm);
    __shared__ __align__(8) unsigned long long bars[2];
    __shared__ uint32_t s_carry;
    const int tid = threadIdx.x;
    int lb = (int)blockIdx.x - P.skip_blocks;
    if (lb < 0) return;
    if (P.smids) {   // logical block id = rank of this CTA's SM id: the 16 CTAs of a group sit on neighbouring SMs
        __shared__ int s_rank;
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if (tid == 0) P.smids[blockIdx.x] = (int)smid * 4096 + blockIdx.x;   // unique key, ordered by SM id
        cooperative_groups::this_grid().sync();
        if (tid == 0) {
            const int mine = P.smids[blockIdx.x];
            int r = 0;
            for (int i = 0; i < (int)gridDim.x; ++i) r += (P.smids[i] < mine);
            s_rank = r;
        }
        __syncthreads();
        lb = s_rank;
    }
    const int grp = lb / XCL, c = lb % XCL;
    if (grp >= P.active_groups) {   // padding CTA: only there to give the launch the 128-CTA placement
        if (tid == 0) { P.out[lb * 2] = 0; P.out[lb * 2 + 1] = 0; }
        return;
    }
    const int rows = P.rows;
    const int wpr = XH / 2;                         // LL words per row
    const size_t par_words = (size_t)rows * wpr;
    unsigned long long* G = P.gbuf + (size_t)grp * 2 * par_words;
    const bool cluster = P.method >= 2;
    if (P.method == 2)
        for (int i = tid; i < 2 * rows * wpr; i += XNT) sll[i] = 0ull;
    if (P.method >= 3 && tid == 0) {
        for (int p = 0; p < 2; ++p)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bars[p])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (cluster) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else if (P.active_groups * XCL == (int)gridDim.x && P.skip_blocks == 0) {
        cooperative_groups::this_grid().sync();
    }
    long long sum = 0;
    uint32_t carry = 1;   // value chain: what is published depends on what was received
    const long long t0 = clock64();
    for (int k = 0; k < P.iters; ++k) {
        const int par = k & 1;
        const uint32_t tag = 0x1000u + (uint32_t)k;
        // ---------------- publish my 32 units of every row
        if (P.method <= 1) {
            for (int i = tid; i < rows * (XUN / 2); i += XNT) {
                const int row = i / (XUN / 2), w = i % (XUN / 2);
                unsigned long long* p = G + par * par_words + (size_t)row * wpr + c * (XUN / 2) + w;
                const uint32_t data = carry + (uint32_t)(row * 7 + w);
                if (P.method == 0) asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(data), "r"(tag) : "memory");
                else asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(data), "r"(tag) : "memory");
            }
        } else if (P.method == 2) {
            for (int i = tid; i < rows * (XUN / 2) * XCL; i += XNT) {
                const int peer = i % XCL, w = (i / XCL) % (XUN / 2), row = i / (XCL * (XUN / 2));
                const uint32_t la = smem_addr(sll + par * par_words + (size_t)row * wpr + c * (XUN / 2) + w);
                const uint32_t data = carry + (uint32_t)(row * 7 + w);
                asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(mapa(la, peer)), "r"(data), "r"(tag) : "memory");
            }
        } else if (P.method >= 4) {
            // one bulk copy per peer: block layout [parity][source CTA][rows][64 B]; the source block is written locally
            // first, then 16 threads (method 4: two lanes of each warp, method 5: 16 lanes of warp 0) push it
            if (tid == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bars[par])), "r"(rows * 64 * XCL) : "memory");
            uint32_t* mine = reinterpret_cast<uint32_t*>(sm + XCL * rows * 64 * 2 + par * rows * 64);   // staging, outside the receive area
            for (int i = tid; i < rows * 16; i += XNT) mine[i] = carry + (uint32_t)((i / 16) * 7 + (i % 16) / 4) + (uint32_t)(i & 3);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            const int lane = tid & 31, wrp = tid >> 5;
            int peer = -1;
            if (P.method == 4) { if (lane < 2) peer = wrp * 2 + lane; }
            else if (wrp == 0 && lane < XCL) peer = lane;
            if (peer >= 0) {
                const uint32_t dst = smem_addr(sm) + (uint32_t)(par * XCL * rows * 64 + c * rows * 64);
                asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(mapa(dst, peer)), "r"(smem_addr(mine)), "r"(rows * 64), "r"(mapa(smem_addr(&bars[par]), peer)) : "memory");
            }
        } else {
            // bf16 payload: row stride 1024 B, my 64 bytes = 4 x 16 B per row and peer
            if (tid == 0) {   // arm my barrier for this iteration: 16 producers x rows x 64 B
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bars[par])), "r"(rows * 64 * XCL) : "memory");
            }
            for (int i = tid; i < rows * 4 * XCL; i += XNT) {
                const int peer = i % XCL, q = (i / XCL) % 4, row = i / (XCL * 4);
                const uint32_t la = smem_addr(sm) + (uint32_t)(par * rows * 1024 + row * 1024 + c * 64 + q * 16);
                const uint32_t data = carry + (uint32_t)(row * 7 + q);
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                             ::"r"(mapa(la, peer)), "r"(data), "r"(data + 1), "r"(data + 2), "r"(data + 3),
                               "r"(mapa(smem_addr(&bars[par]), peer)) : "memory");
            }
        }
        // ---------------- gather: all 512 units of every row
        uint32_t acc = 0;
        if (P.method <= 1) {
            const unsigned long long* src = G + par * par_words;
            for (int base = 0; base < rows * (wpr / 2); base += XNT) {
                const int v = base + tid;
                if (v < rows * (wpr / 2)) {
                    uint4 x;
                    const long long tp = clock64();
                    do {
                        if (P.method == 0) asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "l"(src + 2 * v) : "memory");
                        else asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "l"(src + 2 * v) : "memory");
                        if (clock64() - tp > 2000000000LL) __trap();
                    } while (x.y != tag || x.w != tag);
                    acc += x.x + x.z;
                }
            }
        } else if (P.method == 2) {
            const unsigned long long* src = sll + par * par_words;
            for (int base = 0; base < rows * (wpr / 2); base += XNT) {
                const int v = base + tid;
                if (v < rows * (wpr / 2)) {
                    uint4 x;
                    const long long tp = clock64();
                    do {
                        asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "r"(smem_addr(src + 2 * v)) : "memory");
                        if (clock64() - tp > 2000000000LL) __trap();
                    } while (x.y != tag || x.w != tag);
                    acc += x.x + x.z;
                }
            }
        } else {
            const long long tp = clock64();
            while (!mbar_try(smem_addr(&bars[par]), (uint32_t)((k >> 1) & 1)))
                if (clock64() - tp > 2000000000LL) __trap();
            const uint32_t* src = reinterpret_cast<const uint32_t*>(sm + par * rows * 1024);
            for (int i = tid; i < rows * 256; i += XNT) acc += src[i];
        }
        sum += acc;
        if (tid == 0) s_carry = (acc & 0xffu) + 1u;
        __syncthreads();   // the real step has at least one block barrier between the gather and the next publish
        carry = s_carry;   // block-uniform: every CTA of the group must derive the same value
    }
    const long long t1 = clock64();
    if (cluster) {   // nobody may exit while peers can still write into its shared memory
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (tid == 0) {
        P.out[lb * 2] = t1 - t0;
        P.out[lb * 2 + 1] = sum;
    }
}
}  // namespace

// groups x 16 CTAs; returns mean cycles per exchange round (max over CTAs) and the number of co-resident clusters
// the device reports for the cluster methods (0 for methods 0/1).
int xbench_run(int device, int method, int groups, int rows, int iters, double* cycles_per_iter, int* max_clusters) {
    CUDA_CHECK(cudaSetDevice(device));
    XP P;
    const bool skip6 = method >= 64;     // methods 64+: additionally 6 leading padding blocks (they land on the 6-SM GPC)
    if (skip6) method -= 64;
    const bool padded = method >= 32;    // methods 32 / 33: 0 / 1 in a grid padded to 8 groups (128 CTAs), `groups` of them active
    if (padded) method -= 32;
    const bool by_smid = method >= 16;   // methods 16 / 17: 0 / 1 with groups formed by physical SM id
    if (by_smid) method -= 16;
    P.rows = rows; P.iters = iters; P.method = method;
    P.smids = nullptr;
    const int blocks = (padded ? std::max(groups, 8) : groups) * XCL + (skip6 ? 6 : 0);
    P.active_groups = groups;
    P.skip_blocks = skip6 ? 6 : 0;
    const size_t gwords = (size_t)groups * 2 * rows * (XH / 2);
    CUDA_CHECK(cudaMalloc(&P.gbuf, gwords * 8));
    CUDA_CHECK(cudaMemset(P.gbuf, 0, gwords * 8));
    CUDA_CHECK(cudaMalloc(&P.out, blocks * 2 * sizeof(long long)));
    if (by_smid) {
        if (method >= 2) throw std::runtime_error("xbench: SM-id grouping is for the L2 methods");
        CUDA_CHECK(cudaMalloc(&P.smids, blocks * sizeof(int)));
    }
    const size_t smem = method == 2 ? (size_t)2 * rows * (XH / 2) * 8 : (method == 3 ? (size_t)2 * rows * 1024 : (method >= 4 ? (size_t)2 * rows * 1024 + 2 * rows * 64 : 16));
    CUDA_CHECK(cudaFuncSetAttribute(k_xbench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    *max_clusters = 0;
    if (method >= 2) {
        CUDA_CHECK(cudaFuncSetAttribute(k_xbench, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(XNT); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = XCL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CUDA_CHECK(cudaOccupancyMaxActiveClusters(max_clusters, k_xbench, &cfg));
        if (groups > *max_clusters) throw std::runtime_error("xbench: more clusters than can be co-resident");
        CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_xbench, P));
    } else {
        void* args[] = {&P};
        CUDA_CHECK(cudaLaunchCooperativeKernel((void*)k_xbench, dim3(blocks), dim3(XNT), args, smem, 0));
    }
    CUDA_CHECK(cudaDeviceSynchronize());
    std::vector<long long> h(blocks * 2);
    CUDA_CHECK(cudaMemcpy(h.data(), P.out, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int b = 0; b < groups * XCL; ++b) mx = std::max(mx, h[b * 2]);
    for (int b = 1; b < XCL; ++b)
        if (h[b * 2 + 1] != h[1]) throw std::runtime_error("xbench: CTAs of a group disagree on the gathered data");
    *cycles_per_iter = (double)mx / iters;
    cudaFree(P.gbuf); cudaFree(P.out); cudaFree(P.smids);
    return 0;
}
