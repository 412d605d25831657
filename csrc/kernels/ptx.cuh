// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05
// (alloc / mma / commit / ld / fences) and system-scope flag ops for peer-memory
// protocols.  Nothing here is generic CUDA C++ - this file only compiles for sm_100a.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ssb {

// ------------------------------------------------------------------ misc
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// A spin that cannot hang the GPU forever: ~4 s at 2 GHz, then trap (the host sees a
// launch failure instead of a dead box).
#ifndef SSB_SPIN_LIMIT_CYCLES
#define SSB_SPIN_LIMIT_CYCLES (8000000000ll)
#endif

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > SSB_SPIN_LIMIT_CYCLES) __trap();
    }
}

// ------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
// 2-D tiled load global -> shared, completion on an mbarrier (complete_tx::bytes).
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* tmap, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// 2-D tiled store shared -> global (bulk group completion).
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1)
                 : "memory");
}
// 2-D tiled reduce-add shared -> global (fp32 add performed in L2).
__device__ __forceinline__ void tma_reduce_add_2d(const void* tmap, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1)
                 : "memory");
}
// contiguous bulk copy shared -> global (the destination may be peer memory mapped over NVLink)
__device__ __forceinline__ void bulk_copy_s2g(void* gdst, uint32_t smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// make generic-proxy smem writes visible to the async proxy (TMA / tcgen05 operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------ tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], TF32 inputs, FP32 accumulate, one CTA.
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread retired
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 3xTF32 keeps TWO accumulators per tile: the hi*hi products go to the main one, the two small cross terms (lo*hi,
// hi*lo; 2^-11 of the result) to a second one `small_off` columns further.  The tensor core's accumulate step truncates
// (measured: error grows linearly with the number of accumulate steps, ~0.7 * n * 2^-24 relative), so keeping the small
// terms out of the main chain cuts the steps that matter by 3x; the two are added here in fp32 (round to nearest).
// small_off == 0: single accumulator (TF32 mode).
// two accumulators, both loads in flight before ONE wait (the common 3xTF32 case: one main accumulator + the small one)
__device__ __forceinline__ void tmem_ld16_pair(uint32_t taddr, uint32_t small_off, float (&v)[16]) {
    uint32_t r[16], w[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
          "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
        : "r"(taddr + small_off)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + __uint_as_float(w[i]);
}
// Long reductions additionally rotate the hi*hi products of successive k-blocks over `n_main` main accumulators
// (`main_stride` columns apart): each chain is n_main times shorter; the partial sums are added here.
__device__ __forceinline__ void tmem_ld16_acc(uint32_t taddr, uint32_t small_off, float (&v)[16], int n_main = 1, uint32_t main_stride = 0u) {
    tmem_ld16(taddr, v);
    for (int r = 1; r < n_main; ++r) {
        float w[16];
        tmem_ld16(taddr + (uint32_t)r * main_stride, w);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += w[i];
    }
    if (small_off != 0u) {
        float w[16];
        tmem_ld16(taddr + small_off, w);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += w[i];
    }
}
// how many main accumulators a reduction over `nkb` k-blocks of a tile with `n_cols` columns uses inside a budget of
// `budget_cols` TMEM columns (one more accumulator of the same width holds the small terms)
__host__ __device__ __forceinline__ int acc_rotation(int nkb, int n_cols, int budget_cols) {
    if (nkb < 8) return 1;
    int r = budget_cols / n_cols - 1;
    return r < 1 ? 1 : (r > 4 ? 4 : r);
}

// ------------------------------------------------------------------ UMMA descriptors
// Shared-memory matrix descriptor, SWIZZLE_128B (layout_type 2), Blackwell version 1.
//   K-major  operand tile [rows x 32 fp32]: rows are 128 B apart, 8-row groups 1024 B apart
//            -> SBO = 1024 B, LBO unused.
//   MN-major operand tile = panels of [32 k-rows x 32 fp32 (128 B)]: 8-k-row groups are
//            1024 B apart (SBO), successive 32-wide MN panels are `panel_bytes` apart (LBO).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;   // descriptor version (Blackwell)
    d |= static_cast<uint64_t>(2) << 61;   // SWIZZLE_128B
    return d;
}
// MN-major 32-bit operands: SWIZZLE_128B_BASE32B (layout type 1): atoms of 4 k-rows x 128 B,
// XOR on 32-byte chunks.  LBO = distance between 32-wide MN panels, SBO = distance between
// 4-k-row atoms (512 B when k-rows are contiguous).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128_32b(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;   // descriptor version (Blackwell)
    d |= static_cast<uint64_t>(1) << 61;   // SWIZZLE_128B_BASE32B
    return d;
}
// Descriptor = (hi << 32) | lo.  Within one k-block only the start address (lo[13:0]) moves, so the
// issue loop builds lo/hi once and then steps lo by a constant - the uniform datapath that feeds
// UTCHMMA is slow (each dependent op ~10+ cycles), so every op removed from the issue loop counts.
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
    return ((smem_addr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint32_t umma_desc_hi(uint32_t sbo_bytes, uint32_t layout_type) {
    return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (layout_type << 29);
}
__device__ __forceinline__ uint64_t umma_desc_pack(uint32_t lo, uint32_t hi) {
    return (static_cast<uint64_t>(hi) << 32) | lo;
}
// Instruction descriptor for kind::tf32, fp32 accumulate, M x N tile.
__device__ __forceinline__ uint32_t umma_idesc_tf32(uint32_t M, uint32_t N, uint32_t a_mn_major, uint32_t b_mn_major) {
    return (1u << 4)                 // D format  = F32
           | (2u << 7)               // A format  = TF32
           | (2u << 10)              // B format  = TF32
           | (a_mn_major << 15) | (b_mn_major << 16)
           | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ------------------------------------------------------------------ 3xTF32 operand splitting
// tcgen05.mma.kind::tf32 TRUNCATES the low 13 mantissa bits of its fp32 operands (measured on
// B200: 1+2^-11+2^-12 -> 1, 1+2^-10-2^-23 -> 1, sign-symmetric).  So the raw fp32 tile is the exact
// "hi" operand trunc(x), and lo = x - trunc(x) (exact in fp32) is the only extra tensor needed for
// an fp32-equivalent product  a*b ~= lo_a*hi_b + hi_a*lo_b + hi_a*hi_b  (error ~2^-20 relative).
__device__ __forceinline__ float tf32_lo(float x) {
    return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
}

// ------------------------------------------------------------------ system-scope flags (peer memory)
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// flag store AFTER an explicit system-scope fence (fence + relaxed store = release pattern; a
// st.release.sys per flag would pay one more ~3 us NVLink-draining fence each)
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_add_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_flag_ge(const uint32_t* p, uint32_t target) {
    if (static_cast<int32_t>(ld_acquire_sys(p) - target) >= 0) return;
    const long long t0 = clock64();
    while (static_cast<int32_t>(ld_acquire_sys(p) - target) < 0) {
        if (clock64() - t0 > SSB_SPIN_LIMIT_CYCLES) __trap();
    }
}
// ---- gpu-scope counters: producer kernel -> concurrently running consumer kernel on the same device
__device__ __forceinline__ void red_add_release_gpu(uint32_t* p, uint32_t v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_counter_ge_gpu(const uint32_t* p, uint32_t target) {
    if (static_cast<int32_t>(ld_acquire_gpu(p) - target) >= 0) return;
    const long long t0 = clock64();
    while (static_cast<int32_t>(ld_acquire_gpu(p) - target) < 0) {
        __nanosleep(64);
        if (clock64() - t0 > SSB_SPIN_LIMIT_CYCLES) __trap();
    }
}
// generic-proxy global writes (of another kernel, observed through an acquire) -> this thread's TMA (async proxy) reads
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

__device__ __forceinline__ float4 ld_nc_f4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}

}  // namespace ssb
