// Hardware probes that decide the next kernel designs (run first thing when a GPU is available):
//
//   ingest     per-SM TMA ingest rate vs bytes in flight, L2-hot and HBM-cold, for 1 / 4 / all SMs
//              -> is the chain kernel's layer-1 time (88 GB/s per SM) a latency x ring-depth limit or a port limit?
//   multicast  the same stream, but a 4-CTA cluster where every CTA issues a quarter of each tile with
//              .multicast::cluster -> does multicast raise the per-SM ingest ceiling (4x fewer L2 requests per SM)?
//   dsmem      all-gather of a 16 KB slice between the 4 CTAs of a cluster through st.shared::cluster + one
//              barrier.cluster per round -> cost of the per-layer exchange in a tensor-parallel-cluster MLP kernel
//   cbarrier   barrier.cluster.arrive + wait alone
//   edges      CUDA-graph node-to-node latency: straight line vs fork/join over 8 streams (host + device time)
//
// Build + run (scripts/microbench/run.sh):
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -Icsrc -o /tmp/ssb_probe scripts/microbench/probe.cu
//   /tmp/ssb_probe            # prints one JSON line per measurement
//
// Every spin is bounded (ptx.cuh mbar_wait traps after ~4 s); run under `timeout` anyway.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "kernels/ptx.cuh"

using namespace ssb;

#define CK(expr)                                                                                     \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            printf("{\"error\": \"%s at %s\"}\n", cudaGetErrorString(_e), #expr);                    \
            exit(1);                                                                                 \
        }                                                                                            \
    } while (0)

// ------------------------------------------------------------------------------------------- tensor map
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(const float* base, uint64_t rows, uint32_t box_rows) {
    static EncodeTiledFn enc = nullptr;
    if (!enc) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q));
        enc = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    // [rows, 32] fp32 row-major, contiguous 128-byte rows; box = [box_rows, 32]
    CUtensorMap m;
    cuuint64_t gdim[2] = {32, rows};
    cuuint64_t gstride[1] = {128};
    cuuint32_t box[2] = {32, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        printf("{\"error\": \"cuTensorMapEncodeTiled %d\"}\n", (int)r);
        exit(1);
    }
    return m;
}

// ------------------------------------------------------------------------------------------- cluster helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
                 ::"r"(bar), "r"(cta) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t smem_dst, const void* tmap, uint32_t bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%4, %5}], [%2], %3;"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "h"(mask), "r"(c0), "r"(c1)
        : "memory");
}

// ------------------------------------------------------------------------------------------- ingest probe
// Each CTA streams `tiles` tiles of [128 rows x 128 B] = 16 KB through a ring of `stages` slots.
// Warp 0 produces (TMA), warp 1 consumes (waits for the data, touches nothing, frees the slot).
// MC = true: cluster of 4, CTA r issues rows [32r, 32r+32) of every tile, multicast to all 4.
static constexpr uint32_t kTileRows = 128, kTileBytes = kTileRows * 128;

template <bool MC>
__global__ void __launch_bounds__(64, 1) ingest_kernel(const __grid_constant__ CUtensorMap tm, int tiles, int stages, int tiles_per_stream,
                                                       unsigned long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t bar_base = base + stages * kTileBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = MC ? cluster_ctarank() : 0u;
    // which stream of tiles this CTA (or cluster) reads: disjoint regions so L2 traffic is real
    const int stream = MC ? (int)(blockIdx.x / 4) : (int)blockIdx.x;
    const int row_base = stream * tiles_per_stream * (int)kTileRows;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm);
        for (int s = 0; s < stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), MC ? 4 : 1);       // multicast: all 4 consumers must have released the slot
        }
        fence_barrier_init();
    }
    __syncthreads();
    if (MC) { cluster_arrive(); cluster_wait(); }       // peers' barriers are initialised before anyone multicasts
    const long long t0 = clock64();
    if (warp == 0) {
        for (int t = 0; t < tiles; ++t) {
            const int s = t % stages;
            const uint32_t ph = (t / stages) & 1;
            mbar_wait(empty_bar(s), ph ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(full_bar(s), kTileBytes);   // every CTA receives the whole tile
                const int r0 = row_base + (t % tiles_per_stream) * (int)kTileRows;
                if (MC) tma_load_2d_mc(base + s * kTileBytes + rank * 32u * 128u, &tm, full_bar(s), 0, r0 + (int)rank * 32, (uint16_t)0xF);
                else tma_load_2d(base + s * kTileBytes, &tm, full_bar(s), 0, r0);
            }
            __syncwarp();
        }
    } else {
        for (int t = 0; t < tiles; ++t) {
            const int s = t % stages;
            const uint32_t ph = (t / stages) & 1;
            mbar_wait(full_bar(s), ph);
            if (elect_one()) {
                if (MC) {
                    for (uint32_t c = 0; c < 4; ++c) mbar_arrive_remote(empty_bar(s), c);
                } else {
                    mbar_arrive(empty_bar(s));
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
    if (MC) { cluster_arrive(); cluster_wait(); }       // nobody exits while a peer may still signal its barriers
}

// with MC the box is [32 rows x 128 B]; the swizzle pattern repeats every 8 rows so quarter tiles land consistently
static void run_ingest(const float* buf, size_t buf_rows, int sm_count, float clock_ghz) {
    unsigned long long* cyc = nullptr;
    float* flush = nullptr;
    CK(cudaMalloc(&flush, 256u << 20));                     // > L2: written before every cold measurement
    CK(cudaMalloc(&cyc, 1024 * sizeof(unsigned long long)));
    CK(cudaFuncSetAttribute(ingest_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CK(cudaFuncSetAttribute(ingest_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CUtensorMap tm_full = make_map(buf, buf_rows, kTileRows);
    CUtensorMap tm_quarter = make_map(buf, buf_rows, 32);
    const int grids[3] = {1, 4, sm_count / 4 * 4};
    const int stage_opts[4] = {2, 4, 8, 12};
    for (int mc = 0; mc < 2; ++mc)
        for (int gi = 0; gi < 3; ++gi)
            for (int si = 0; si < 4; ++si)
                for (int cold = 0; cold < 2; ++cold) {
                    const int grid = grids[gi], stages = stage_opts[si];
                    if (mc && grid < 4) continue;
                    const int streams = mc ? grid / 4 : grid;
                    // hot: 32 tiles (0.5 MB) per stream re-read 16 times (L2 resident even with 148 streams); cold: 1024 distinct
                    // tiles (16 MB) per stream
                    const int tiles_per_stream = cold ? 1024 : 32;
                    const int tiles = cold ? 1024 : 512;
                    if ((size_t)streams * tiles_per_stream * kTileRows > buf_rows) continue;
                    const int smem = stages * kTileBytes + 1024 + 16 * stages + 64;
                    cudaEvent_t e0, e1;
                    CK(cudaEventCreate(&e0));
                    CK(cudaEventCreate(&e1));
                    float best_ms = 1e9f;
                    for (int rep = 0; rep < (cold ? 1 : 3); ++rep) {
                        if (cold) CK(cudaMemsetAsync(flush, 0, 256u << 20));   // evict L2
                        CK(cudaEventRecord(e0));
                        if (mc) {
                            cudaLaunchConfig_t cfg = {};
                            cfg.gridDim = dim3(grid);
                            cfg.blockDim = dim3(64);
                            cfg.dynamicSmemBytes = smem;
                            cudaLaunchAttribute at[1];
                            at[0].id = cudaLaunchAttributeClusterDimension;
                            at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                            cfg.attrs = at; cfg.numAttrs = 1;
                            CK(cudaLaunchKernelEx(&cfg, ingest_kernel<true>, tm_quarter, tiles, stages, tiles_per_stream, cyc));
                        } else {
                            ingest_kernel<false><<<grid, 64, smem>>>(tm_full, tiles, stages, tiles_per_stream, cyc);
                        }
                        CK(cudaEventRecord(e1));
                        CK(cudaEventSynchronize(e1));
                        CK(cudaGetLastError());
                        float ms;
                        CK(cudaEventElapsedTime(&ms, e0, e1));
                        if (ms < best_ms) best_ms = ms;
                    }
                    std::vector<unsigned long long> h(grid);
                    CK(cudaMemcpy(h.data(), cyc, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
                    unsigned long long mx = 0;
                    for (auto v : h) mx = v > mx ? v : mx;
                    const double bytes_per_cta = (double)tiles * kTileBytes;
                    const double us = (double)mx / (clock_ghz * 1e3);
                    printf("{\"probe\": \"ingest\", \"multicast\": %d, \"ctas\": %d, \"stages\": %d, \"bytes_in_flight\": %d, \"source\": \"%s\", "
                           "\"GBps_per_sm\": %.1f, \"GBps_total_received\": %.1f, \"us_kernel\": %.2f, \"GBps_total_by_event\": %.1f}\n",
                           mc, grid, stages, stages * (int)kTileBytes, cold ? "hbm" : "l2", bytes_per_cta / us / 1e3,
                           bytes_per_cta * grid / us / 1e3, best_ms * 1e3, bytes_per_cta * grid / (best_ms * 1e6));
                    fflush(stdout);
                    CK(cudaEventDestroy(e0));
                    CK(cudaEventDestroy(e1));
                }
    CK(cudaFree(cyc));
    CK(cudaFree(flush));
}

// ------------------------------------------------------------------------------------------- DSMEM all-gather
// cluster of 4, 128 threads: every round each CTA stores its 16 KB slice into the 3 peers (st.shared::cluster.v4)
// and then the cluster synchronises.  gather = 0 measures the barrier alone.
__global__ void __launch_bounds__(128, 1) dsmem_kernel(int rounds, int gather, int slice_bytes, unsigned long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = smem_u32(smem_raw);
    const uint32_t rank = cluster_ctarank();
    cluster_arrive(); cluster_wait();
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        if (gather) {
            const float v = (float)(r + rank);
            for (uint32_t peer = 1; peer < 4; ++peer) {
                const uint32_t dst = mapa(base + rank * slice_bytes, (rank + peer) & 3u);
                for (int off = threadIdx.x * 16; off < slice_bytes; off += 128 * 16)
                    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(dst + off), "f"(v) : "memory");
            }
        }
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
}

static void run_dsmem(float clock_ghz) {
    unsigned long long* cyc = nullptr;
    CK(cudaMalloc(&cyc, 64 * sizeof(unsigned long long)));
    CK(cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    const int rounds = 2000;
    for (int gather = 0; gather < 2; ++gather)
        for (int slice = 4096; slice <= 16384; slice *= 2) {
            if (!gather && slice != 4096) continue;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(4);
            cfg.blockDim = dim3(128);
            cfg.dynamicSmemBytes = 4 * 16384;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            CK(cudaLaunchKernelEx(&cfg, dsmem_kernel, rounds, gather, slice, cyc));
            CK(cudaDeviceSynchronize());
            unsigned long long h[4];
            CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
            const double ns = (double)h[0] / rounds / clock_ghz;
            printf("{\"probe\": \"%s\", \"slice_bytes\": %d, \"ns_per_round\": %.1f}\n", gather ? "dsmem_allgather4" : "cluster_barrier4",
                   gather ? slice : 0, ns);
            fflush(stdout);
        }
    CK(cudaFree(cyc));
}

// ------------------------------------------------------------------------------------------- graph edge latency
__global__ void tiny_kernel(float* p) {
    if (threadIdx.x == 0 && p != nullptr) p[blockIdx.x] += 1.f;
}

static void run_edges() {
    float* buf = nullptr;
    CK(cudaMalloc(&buf, 4096));
    CK(cudaMemset(buf, 0, 4096));
    cudaStream_t s[9];
    for (auto& st : s) CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaEvent_t fork, join[8], e0, e1;
    CK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
    for (auto& j : join) CK(cudaEventCreateWithFlags(&j, cudaEventDisableTiming));
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int shape = 0; shape < 2; ++shape) {
        // shape 0: 10 kernels in a straight line; shape 1: 1 kernel -> 8 parallel kernels on 8 streams -> 1 kernel
        cudaGraph_t g;
        cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(s[0], cudaStreamCaptureModeThreadLocal));
        if (shape == 0) {
            for (int i = 0; i < 10; ++i) tiny_kernel<<<1, 32, 0, s[0]>>>(buf);
        } else {
            tiny_kernel<<<1, 32, 0, s[0]>>>(buf);
            CK(cudaEventRecord(fork, s[0]));
            for (int i = 0; i < 8; ++i) {
                CK(cudaStreamWaitEvent(s[1 + i], fork, 0));
                tiny_kernel<<<1, 32, 0, s[1 + i]>>>(buf + 32 * (i + 1));
                CK(cudaEventRecord(join[i], s[1 + i]));
                CK(cudaStreamWaitEvent(s[0], join[i], 0));
            }
            tiny_kernel<<<1, 32, 0, s[0]>>>(buf);
        }
        CK(cudaStreamEndCapture(s[0], &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        for (int i = 0; i < 20; ++i) CK(cudaGraphLaunch(ge, s[0]));
        CK(cudaStreamSynchronize(s[0]));
        const int reps = 200;
        CK(cudaEventRecord(e0, s[0]));
        for (int i = 0; i < reps; ++i) CK(cudaGraphLaunch(ge, s[0]));
        CK(cudaEventRecord(e1, s[0]));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("{\"probe\": \"graph_edges\", \"shape\": \"%s\", \"us_per_graph\": %.2f}\n",
               shape == 0 ? "line of 10 kernels" : "1 -> 8 parallel -> 1 (fork/join over 8 streams)", ms * 1e3 / reps);
        fflush(stdout);
        CK(cudaGraphExecDestroy(ge));
        CK(cudaGraphDestroy(g));
    }
    CK(cudaFree(buf));
}

int main(int argc, char** argv) {
    int dev = 0, sms = 0, khz = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    const float ghz = khz / 1e6f;     // clock64 ticks at the SM clock; boost clock is the upper bound
    printf("{\"probe\": \"device\", \"sms\": %d, \"clock_ghz_nominal\": %.3f}\n", sms, ghz);
    const bool only = argc > 1;
    auto want = [&](const char* name) { return !only || std::string(argv[1]) == name; };
    if (want("edges")) run_edges();
    if (want("dsmem")) run_dsmem(ghz);
    if (want("ingest")) {
        const size_t rows = (size_t)(3u << 30) / 128;      // 3 GB of 128-byte rows
        float* buf = nullptr;
        CK(cudaMalloc(&buf, rows * 128));
        CK(cudaMemset(buf, 0, rows * 128));
        run_ingest(buf, rows, sms, ghz);
        CK(cudaFree(buf));
    }
    return 0;
}
