// Layer-chain kernel for narrow MLPs (every layer width <= 128, any first-layer fan-in).
//
// One CTA owns one micro-batch (<= 128 rows) and walks it through the WHOLE stage in a
// single launch:   forward L layers -> loss head -> backward (dgrad) L-1 layers.
//
//   * activations never leave the SM between layers: the epilogue writes the layer output
//     into shared memory directly in the swizzled K-major layout tcgen05.mma wants as the
//     next layer's B operand (and a copy to global for the weight-gradient GEMMs / pipeline);
//   * weights are streamed by a producer thread that runs AHEAD of the math through a deep
//     TMA ring - weight tiles do not depend on activations, so layer l+1's weights are in
//     flight while layer l is being computed;
//   * accumulators live in TMEM; fused epilogues: bias + ReLU (forward), softmax + MSE
//     gradient + softmax Jacobian (loss head, warp shuffles across the class lanes), ReLU
//     mask (backward).
//
// This removes ~14 dependent kernel launches (and their prologues / cold starts) from the
// critical path of the reference workload; the per-layer kernels in tc_gemm.cu remain the
// general path (wide layers) and compute the weight gradients afterwards.
#include "kernels.h"
#include "ptx.cuh"

#include <algorithm>

namespace ssb {

static constexpr int kThreads = 320;                      // 2 control warps + 8 epilogue warps
static constexpr uint32_t kBlockM = 128;
static constexpr uint32_t kBlockK = 32;
static constexpr uint32_t kABytes = kBlockM * 128;
static constexpr uint32_t kPanelBytes = 32 * 128;
static constexpr int kScratchLd = 33;                     // odd pitch: conflict-free row-per-lane access

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// byte offset of element (row n, feature m) inside a K-major SWIZZLE_128B operand tile made of
// 32-feature panels of [n_pad rows x 128 B]
__device__ __forceinline__ uint32_t act_smem_off(int n, int m, uint32_t panel_bytes) {
    const uint32_t c = (uint32_t)(m & 31);
    return (uint32_t)(m >> 5) * panel_bytes + (uint32_t)n * 128u + ((((c >> 2) ^ ((uint32_t)n & 7u))) << 4) + ((c & 3u) << 2);
}

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define DBG(role, idx) do { if (p.dbg != nullptr && blockIdx.x == 0 && (idx) < 256) p.dbg[(role) * 256 + (idx)] = gtime(); } while (0)

// SPLIT (3xTF32) is a compile-time switch: the single-pass TF32 instantiation carries none of the lo-twin code.
// (Round 2 also tried deriving the lo twins of the streamed weight tiles in shared memory instead of loading them - half
// the L2->SM bytes, bit-identical results - but with only 3 ring slots of 40 KB the extra split stage in the slot cycle
// cost more than the bytes saved: 92.8 vs 78.7 us/step.  Removed; see profiles/variants_r2.md.)
// FOLD (pipeline boundaries inside the kernel, see ChainParams) is a compile-time switch as well: merely carrying that code
// - a prologue, two predicated peer stores per element in the epilogues - cost the single-GPU / data-parallel step 8 us
// (same-box bisect, profiles/variants_r2.md), so launches without a folded boundary use the instantiation without it.
// ACC (3xTF32 only): separate + rotating accumulators for long reductions (layer 1, K = 784: error 2.9x cuBLAS fp32 instead
// of 22x, profiles/precision_r2.md).  Compile-time as well, and OFF by default for this kernel: the same one-call, same-box
// comparison of build variants showed that carrying that code costs 6 us per step (85.0 vs 79.2 us) whether or not it runs.
// `SSB_CHAIN_ACC=1` selects the accurate instantiation; the per-layer GEMM kernels (wide layers) always use the scheme.
template <bool SPLIT, bool FOLD, bool ACC>
__global__ void __launch_bounds__(kThreads, 1) mlp_chain_kernel(const ChainParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t smem_base = (raw + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - raw);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int L = p.n_layers;
    const int N = p.n_pad;                                   // UMMA N (rows of this micro-batch, padded to 16)
    const int row0 = (p.mu_base + (int)blockIdx.x) * p.mb_rows;                 // first global row of this micro-batch
    const uint32_t b_bytes = (uint32_t)N * 128u;             // one 32-feature panel of an activation tile
    const uint32_t half_stage = (uint32_t)p.kps * (kABytes + b_bytes);    // kps A tiles, then kps X tiles (layer 1)
    const uint32_t stage_bytes = SPLIT ? 2u * half_stage : half_stage;  // split: lo twins in the second half
    const uint32_t stage_b_off = (uint32_t)p.kps * kABytes;
    const uint32_t abuf_bytes = 4u * b_bytes;
    const uint32_t n_abuf = SPLIT ? 4u : 2u;                            // hi ping-pong (+ lo ping-pong)
    const uint32_t abuf0 = smem_base + p.stages * stage_bytes;
    const uint32_t lo_off = 2u * abuf_bytes;                              // lo tile of buffer b sits lo_off behind its hi tile
    const uint32_t bar_base = abuf0 + n_abuf * abuf_bytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
    const uint32_t tmem_full_bar = bar_base + 8u * (2 * p.stages);
    const uint32_t act_ready_bar = tmem_full_bar + 8u;
    const uint32_t tmem_slot = act_ready_bar + 8u;
    volatile uint32_t* tmem_slot_gen =
        reinterpret_cast<volatile uint32_t*>(smem_gen + p.stages * stage_bytes + n_abuf * abuf_bytes + 8u * (2 * p.stages + 2));
    // 3xTF32: up to 4 rotating main accumulators (long reductions only: layer 1) + one for the small cross terms, see ptx.cuh
    constexpr int kTmemBudget = 512;
    constexpr bool acc_split = SPLIT && ACC;
    const int max_rot = acc_split ? acc_rotation(1 << 20, N, kTmemBudget) : 1;
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)N * (acc_split ? (uint32_t)max_rot + 1u : 1u)) tmem_cols <<= 1;
    const uint32_t small_off = acc_split ? (uint32_t)N * (uint32_t)max_rot : 0u;
    // loss-head transpose scratch [N][kScratchLd] floats, behind the barriers (128 B further)
    const uint32_t scratch_off = p.stages * stage_bytes + n_abuf * abuf_bytes + 8u * (2 * p.stages + 2) + 128u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        mbar_init(act_ready_bar, 8);                         // one arrive per epilogue warp
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    // number of backward GEMMs: layers L..lo (layer 1's dgrad is skipped on the first stage)
    const int bwd_lo = p.first_stage ? 2 : 1;
    const int lo_base = 2 * L + 1;                           // index of the first lo-twin tensor map

    if (warp == 0) {
        // ============================================================ weight (and X) producer
        // The whole warp walks the loop (converged); one elect.sync-chosen lane issues, so ptxas can
        // feed the uniform-datapath instructions (UTMALDG) directly instead of an ELECT/BRA.U.ANY loop.
        {
            int it = 0;
            auto acquire = [&](uint32_t bytes) {
                const int s = it % p.stages;
                mbar_wait(empty_bar(s), ((it / p.stages) & 1) ^ 1);
                return s;
            };
            if (p.do_fwd) {
                for (int l = 1; l <= L; ++l) {
                    const int nkb = (p.layers[l - 1].in + (int)kBlockK - 1) / (int)kBlockK;
                    const bool with_x = (l == 1) && !(FOLD && p.x_from_global);   // folded pipeline input: staged by the epilogue warps
                    for (int kb0 = 0; kb0 < nkb; kb0 += p.kps, ++it) {
                        const int cnt = min(p.kps, nkb - kb0);
                        const int s = acquire(0);
                        const uint32_t a_dst = smem_base + s * stage_bytes;
                        if (elect_one()) {
                            DBG(0, it);
                            mbar_arrive_expect_tx(full_bar(s), (uint32_t)cnt * (kABytes + (with_x ? b_bytes : 0u)) * (SPLIT ? 2u : 1u));
                            for (int j = 0; j < cnt; ++j) {
                                tma_load_2d(a_dst + j * kABytes, p.maps + 2 * (l - 1), full_bar(s), (kb0 + j) * kBlockK, 0);
                                if (with_x)
                                    tma_load_2d(a_dst + stage_b_off + j * b_bytes, p.maps + 2 * L, full_bar(s), (kb0 + j) * kBlockK, row0);
                                if (SPLIT) {
                                    tma_load_2d(a_dst + half_stage + j * kABytes, p.maps + lo_base + 2 * (l - 1), full_bar(s), (kb0 + j) * kBlockK, 0);
                                    if (with_x)
                                        tma_load_2d(a_dst + half_stage + stage_b_off + j * b_bytes, p.maps + lo_base + 2 * L, full_bar(s),
                                                    (kb0 + j) * kBlockK, row0);
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
            }
            if (p.do_bwd) {
                for (int l = L; l >= bwd_lo; --l) {
                    const int nkb = (p.layers[l - 1].out + (int)kBlockK - 1) / (int)kBlockK;
                    for (int kb0 = 0; kb0 < nkb; kb0 += p.kps, ++it) {
                        const int cnt = min(p.kps, nkb - kb0);
                        const int s = acquire(0);
                        const uint32_t a_dst = smem_base + s * stage_bytes;
                        if (elect_one()) {
                            DBG(0, it);
                            mbar_arrive_expect_tx(full_bar(s), (uint32_t)cnt * kABytes * (SPLIT ? 2u : 1u));
                            for (int j = 0; j < cnt; ++j) {
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    tma_load_2d(a_dst + j * kABytes + i * kPanelBytes, p.maps + 2 * (l - 1) + 1, full_bar(s), 32 * i,
                                                (kb0 + j) * kBlockK);
                                    if (SPLIT)
                                        tma_load_2d(a_dst + half_stage + j * kABytes + i * kPanelBytes, p.maps + lo_base + 2 * (l - 1) + 1,
                                                    full_bar(s), 32 * i, (kb0 + j) * kBlockK);
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ============================================================ MMA issuer (converged warp, elected lane issues)
        {
            int it = 0, act_waits = 0, gemm_i = 0;
            const uint32_t idesc_f = umma_idesc_tf32(kBlockM, N, 0u, 0u);
            const uint32_t idesc_b = umma_idesc_tf32(kBlockM, N, 1u, 0u);
            const uint32_t k_hi = umma_desc_hi(1024u, 2u), mn_hi = umma_desc_hi(512u, 1u);
            auto run_gemm = [&](int nkb, bool a_mn, bool b_from_stage, uint32_t bsrc) {
                if (!b_from_stage) {                          // operand tile written by the epilogue warps; their
                    mbar_wait(act_ready_bar, (act_waits & 1));   // arrive also means the accumulator was drained
                    ++act_waits;
                    tc_fence_after();
                }
                if (lane == 0) DBG(1, 3 * gemm_i);
                // only LONG reductions (layer 1 of the first stage: K = 784) use the extra accumulators; the 128-wide layers
                // keep one accumulator and the lean epilogue (their error is ~4x cuBLAS fp32 already)
                const bool acc_l = acc_split && nkb >= 8;
                const int rot = acc_l ? acc_rotation(nkb, N, kTmemBudget) : 1;
                uint32_t rot_i = 0u;                          // main accumulator of the current k-block (round robin)
                for (int kb0 = 0; kb0 < nkb; kb0 += p.kps, ++it) {
                    const int cnt = min(p.kps, nkb - kb0);
                    const int s = it % p.stages;
                    mbar_wait(full_bar(s), (it / p.stages) & 1);
                    tc_fence_after();
                    if (kb0 + cnt >= nkb && lane == 0) DBG(1, 3 * gemm_i + 1);
                    const uint32_t a_src = smem_base + s * stage_bytes;
                    if (elect_one()) {
                        uint32_t rot_e = rot_i;               // the elected lane's running copy; every lane advances rot_i below
                        for (int j = 0; j < cnt; ++j) {
                            const uint32_t a_j = a_src + j * kABytes;
                            const uint32_t b_j = b_from_stage ? a_src + stage_b_off + j * b_bytes : bsrc + (kb0 + j) * b_bytes;
                            const uint32_t a_lo = a_mn ? umma_desc_lo(a_j, kPanelBytes) : umma_desc_lo(a_j, 16u);
                            const uint32_t b_lo = umma_desc_lo(b_j, 16u);
                            const uint32_t a_step = a_mn ? 64u : 2u;
                            const uint32_t ah = a_mn ? mn_hi : k_hi, id = a_mn ? idesc_b : idesc_f;
                            if (SPLIT) {                           // lo*hi + hi*lo + hi*hi (raw tiles are the hi parts)
                                const uint32_t al_lo = a_lo + (half_stage >> 4);
                                const uint32_t bl_lo = b_lo + ((b_from_stage ? half_stage : lo_off) >> 4);
                                if (!acc_l) {
                                    // short reduction: one accumulator, nothing but the three MMAs in the issue path (this warp sits
                                    // on the ring's critical cycle: every extra instruction per k-block shows up 73 times per step)
#pragma unroll
                                    for (int k4 = 0; k4 < 4; ++k4) {
                                        umma_tf32(tmem_base, umma_desc_pack(al_lo + k4 * a_step, ah), umma_desc_pack(b_lo + k4 * 2u, k_hi), id,
                                                  ((kb0 + j) | k4) != 0 ? 1u : 0u);
                                        umma_tf32(tmem_base, umma_desc_pack(a_lo + k4 * a_step, ah), umma_desc_pack(bl_lo + k4 * 2u, k_hi), id, 1u);
                                        umma_tf32(tmem_base, umma_desc_pack(a_lo + k4 * a_step, ah), umma_desc_pack(b_lo + k4 * 2u, k_hi), id, 1u);
                                    }
                                } else {
                                    // long reduction: small cross terms in their own accumulator, hi*hi rotating over `rot` accumulators
                                    const bool first_kb = (kb0 + j) == 0;
                                    const uint32_t small_acc = tmem_base + small_off;
                                    const uint32_t main_acc = tmem_base + rot_e * (uint32_t)N;
                                    const bool main_fresh = (kb0 + j) < rot;   // first product into this main accumulator
                                    rot_e = (rot_e + 1u == (uint32_t)rot) ? 0u : rot_e + 1u;
#pragma unroll
                                    for (int k4 = 0; k4 < 4; ++k4) {
                                        umma_tf32(small_acc, umma_desc_pack(al_lo + k4 * a_step, ah), umma_desc_pack(b_lo + k4 * 2u, k_hi), id,
                                                  (first_kb && k4 == 0) ? 0u : 1u);
                                        umma_tf32(small_acc, umma_desc_pack(a_lo + k4 * a_step, ah), umma_desc_pack(bl_lo + k4 * 2u, k_hi), id, 1u);
                                        umma_tf32(main_acc, umma_desc_pack(a_lo + k4 * a_step, ah), umma_desc_pack(b_lo + k4 * 2u, k_hi), id,
                                                  (main_fresh && k4 == 0) ? 0u : 1u);
                                    }
                                }
                            } else
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4)
                                umma_tf32(tmem_base, umma_desc_pack(a_lo + k4 * a_step, ah), umma_desc_pack(b_lo + k4 * 2u, k_hi), id,
                                          ((kb0 + j) | k4) != 0 ? 1u : 0u);
                        }
                        umma_commit(empty_bar(s));
                        if (kb0 + cnt >= nkb) umma_commit(tmem_full_bar);
                    }
                    if (acc_l)
                        for (int j = 0; j < cnt; ++j) rot_i = (rot_i + 1u == (uint32_t)rot) ? 0u : rot_i + 1u;   // all lanes, off the issue path
                    __syncwarp();
                }
                if (lane == 0) DBG(1, 3 * gemm_i + 2);
                ++gemm_i;
            };
            int buf = 0;                                     // activation buffer the NEXT gemm reads
            if (p.do_fwd) {
                for (int l = 1; l <= L; ++l) {
                    const int nkb = (p.layers[l - 1].in + (int)kBlockK - 1) / (int)kBlockK;
                    // layer 1 reads X from its ring slots (TMA), or - folded pipeline input - from abuf 1 (staged by the epilogue warps)
                    const bool x_glob = FOLD && (l == 1) && p.x_from_global;
                    run_gemm(nkb, false, l == 1 && !x_glob, abuf0 + (x_glob ? 1 : buf) * abuf_bytes);
                    if (l > 1) buf ^= 1;                     // layer l read buf, wrote buf^1
                    else buf = 0;                            // layer 1 wrote abuf 0
                }
            }
            if (p.do_bwd) {
                // the loss head / boundary loader wrote dZ_L into the buffer after the last forward output
                for (int l = L; l >= bwd_lo; --l) {
                    const int nkb = (p.layers[l - 1].out + (int)kBlockK - 1) / (int)kBlockK;
                    run_gemm(nkb, true, false, abuf0 + buf * abuf_bytes);
                    buf ^= 1;
                }
            }
        }
    } else {
        // ============================================================ epilogue warps
        const int q = warp & 3;                              // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;                    // two warps per quarter: each takes half of the columns
        const int m = q * 32 + lane;                         // feature handled by this thread (TMEM lane)
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        // column (= row of the micro-batch) range of this warp, in chunks of 16
        const int n_chunks = N / 16;
        const int c_lo = 16 * ((n_chunks * half) / 2), c_hi = 16 * ((n_chunks * (half + 1)) / 2);
        int tmem_waits = 0;
        int wbuf = 0;                                        // activation buffer the NEXT epilogue writes
        auto st_tile = [&](uint32_t dst, int n, int feat, float x) {   // operand tile(s) for the next GEMM
            const uint32_t off = (dst - smem_base) + act_smem_off(n, feat, b_bytes);
            *reinterpret_cast<float*>(smem_gen + off) = x;
            if (SPLIT) *reinterpret_cast<float*>(smem_gen + off + lo_off) = tf32_lo(x);
        };
        // dz[l] (and every global store this warp issued before) is complete: one release-add per warp on ready[l]
        // (consumers: the gated weight-gradient / data-parallel kernels running next to this one)
        auto signal_ready = [&](int l) {
            if (p.ready == nullptr) return;
            __syncwarp();
            if (lane == 0) red_add_release_gpu(p.ready + l, 1u);
        };
        auto publish = [&]() {                               // smem tile complete -> MMA warp may read it
            if (threadIdx.x == 64) DBG(2, 2 * (tmem_waits - 1) + 1);
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(act_ready_bar);
        };
        // ---------------- folded pipeline boundaries (see ChainParams): credit of the slots we will write, arrival of
        // the tile we consume.  One thread spins (bounded), a named barrier among the 8 epilogue warps publishes the result.
        uint32_t pp_epoch = 0u;
        if constexpr (FOLD) {
        if (p.in_flag != nullptr || p.out_peer != nullptr) {
            pp_epoch = *reinterpret_cast<const volatile uint32_t*>(p.pp_epoch);
            if (threadIdx.x == 64) {
                if (p.out_peer != nullptr) wait_flag_ge(p.out_credit, pp_epoch - 1u);
                if (p.in_flag != nullptr) wait_flag_ge(p.in_flag, pp_epoch);
            }
            asm volatile("bar.sync 2, 256;" ::: "memory");
        }
        if (p.do_fwd && p.x_from_global) {
            // stage the input tile [rows x in] (written into act[0] by the previous stage over NVLink) as layer 1's B operand
            // in abuf 1, and its lo twin for the weight-gradient GEMM of layer 1
            const int in0 = p.layers[0].in;
            const uint32_t dstx = abuf0 + 1u * abuf_bytes;
            const float* __restrict__ xin = p.act[0] + (int64_t)row0 * p.act_ld[0];
            for (int n = c_lo; n < c_hi; ++n) {
                float x = 0.f;
                if (m < in0 && n < p.mb_rows) {
                    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(x) : "l"(xin + (int64_t)n * p.act_ld[0] + m));
                    if (SPLIT) p.act_lo[0][(int64_t)(row0 + n) * p.act_ld[0] + m] = tf32_lo(x);
                }
                st_tile(dstx, n, m, x);
            }
            publish();
        }
        }   // FOLD
        if (p.do_fwd) {
            for (int l = 1; l <= L; ++l) {
                const ChainLayer& ly = p.layers[l - 1];
                const bool m_ok = m < ly.out;
                const float bias = m_ok ? __ldg(p.W + ly.w_off + (int64_t)m * ly.ldw + ly.in) : 0.f;
                const bool is_logits = (l == L) && p.do_loss;
                const int nkb_f = (ly.in + (int)kBlockK - 1) / (int)kBlockK;
                const bool acc_f = acc_split && nkb_f >= 8;   // as the MMA warp chose: extra accumulators for long reductions only
                mbar_wait(tmem_full_bar, tmem_waits & 1);
                ++tmem_waits;
                tc_fence_after();
                if (p.sync_debug) asm volatile("bar.sync 2, 256;" ::: "memory");   // racecheck aid, see kernels.h
                if (threadIdx.x == 64) DBG(2, 2 * (tmem_waits - 1));
                const uint32_t dst = abuf0 + wbuf * abuf_bytes;
                float* __restrict__ gout = p.act[l] + (int64_t)row0 * p.act_ld[l];
                if (!is_logits) {
                    // one 16-row chunk per warp when N <= 32: write the operand tile for the next GEMM first,
                    // publish it, and only then drain the global copy (needed later by wgrad / the pipeline)
                    const bool single = (c_hi - c_lo == 16);
                    float keep[16];
                    for (int c = c_lo; c < c_hi; c += 16) {
                        float v[16];
                        if (acc_f) tmem_ld16_acc(taddr + c, small_off, v, acc_rotation(nkb_f, N, kTmemBudget), (uint32_t)N); else tmem_ld16(taddr + c, v);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float x = v[j] + bias;
                            if (ly.relu) x = fmaxf(x, 0.f);
                            x = m_ok ? x : 0.f;
                            const int n = c + j;
                            st_tile(dst, n, m, x);
                            if (single) keep[j] = x;
                            else if (m_ok && n < p.mb_rows) {
                                gout[(int64_t)n * p.act_ld[l] + m] = x;
                                if (SPLIT) p.act_lo[l][(int64_t)(row0 + n) * p.act_ld[l] + m] = tf32_lo(x);
                                if constexpr (FOLD) { if (l == L && p.out_peer != nullptr) p.out_peer[(int64_t)n * p.act_ld[l] + m] = x; }
                            }
                        }
                    }
                    publish();
                    if (single && m_ok) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c_lo + j < p.mb_rows) {
                                gout[(int64_t)(c_lo + j) * p.act_ld[l] + m] = keep[j];
                                if (SPLIT) p.act_lo[l][(int64_t)(row0 + c_lo + j) * p.act_ld[l] + m] = tf32_lo(keep[j]);
                                if constexpr (FOLD) { if (l == L && p.out_peer != nullptr) p.out_peer[(int64_t)(c_lo + j) * p.act_ld[l] + m] = keep[j]; }
                            }
                    }
                    wbuf ^= 1;
                    if (l == 1) wbuf = 1;                    // layer 1 wrote abuf 0; layer 2 reads 0, writes 1
                } else {
                    // ---------------- loss head on the logits tile (class = TMEM lane, row = TMEM column)
                    // softmax contract of the reference: shift by the GLOBAL max of the micro-batch, +1e-7.
                    // Step 1: warp q == 0 (the class lanes) transposes logits (+bias) into a [row][class] scratch
                    // in smem.  Step 2: each lane owns ROWS, loops over the <= 32 classes in registers - no
                    // cross-lane reductions except one max and one sum for the whole micro-batch.
                    const int C = ly.out;
                    float* scratch = reinterpret_cast<float*>(smem_gen + scratch_off);   // [N][kScratchLd]
                    float loss = 0.f;
                    if (q == 0 && half == 0) {
                        for (int c = 0; c < N; c += 16) {
                            float v[16];
                            if (acc_f) tmem_ld16_acc(taddr + c, small_off, v, acc_rotation(nkb_f, N, kTmemBudget), (uint32_t)N); else tmem_ld16(taddr + c, v);
                            if (m < C) {
#pragma unroll
                                for (int j = 0; j < 16; ++j) scratch[(c + j) * kScratchLd + m] = v[j] + bias;
                            }
                        }
                        __syncwarp();
                        float gmax = -3.0e38f;
                        for (int n = lane; n < p.mb_rows; n += 32)
                            for (int k = 0; k < C; ++k) gmax = fmaxf(gmax, scratch[n * kScratchLd + k]);
                        gmax = warp_max_f(gmax);
                        for (int n = lane; n < N; n += 32) {
                            const bool row_ok = n < p.mb_rows;
                            const float* __restrict__ tg = p.target + (int64_t)(row0 + n) * p.ldt;
                            const float* __restrict__ zr = scratch + n * kScratchLd;
                            if (C <= 16) {
                                // fast path (the usual 10-class head): everything in registers, loops fully
                                // unrolled to a constant bound so the 16 target loads / exps are independent
                                float tgv[16], ev[16];
                                float ssum = 0.f;
#pragma unroll
                                for (int k = 0; k < 16; ++k) {
                                    tgv[k] = (k < C && row_ok) ? __ldg(tg + k) : 0.f;
                                    ev[k] = (k < C) ? expf(zr[k] - gmax) : 0.f;
                                    ssum += ev[k];
                                }
                                const float inv = 1.f / (ssum + 1e-7f);
                                float gs = 0.f;
#pragma unroll
                                for (int k = 0; k < 16; ++k) {
                                    const float pr = ev[k] * inv;
                                    const float d = tgv[k] - pr;
                                    if (k < C && row_ok) loss += d * d;
                                    const float g = pr * (-2.f * d * p.inv_batch);
                                    gs += (k < C) ? g : 0.f;
                                    ev[k] = pr;
                                    tgv[k] = g;
                                }
#pragma unroll
                                for (int k = 0; k < 16; ++k) {
                                    if (k < C) {
                                        const float dzv = row_ok ? (tgv[k] - ev[k] * gs) : 0.f;
                                        if (row_ok) {
                                            gout[(int64_t)n * p.act_ld[l] + k] = zr[k];
                                            p.probs[(int64_t)(row0 + n) * p.ldp + k] = ev[k];
                                            if (p.dz[l] != nullptr) p.dz[l][(int64_t)(row0 + n) * p.act_ld[l] + k] = dzv;
                                            if (SPLIT && p.dz_lo[l] != nullptr) p.dz_lo[l][(int64_t)(row0 + n) * p.act_ld[l] + k] = tf32_lo(dzv);
                                        }
                                        if (p.do_bwd) st_tile(dst, n, k, dzv);
                                    }
                                }
                            } else {
                                float ssum = 0.f;
                                for (int k = 0; k < C; ++k) ssum += expf(zr[k] - gmax);
                                const float inv = 1.f / (ssum + 1e-7f);
                                float gs = 0.f;
                                for (int k = 0; k < C; ++k) {
                                    const float pr = expf(zr[k] - gmax) * inv;
                                    const float d = (row_ok ? __ldg(tg + k) : pr) - pr;
                                    if (row_ok) loss += d * d;
                                    gs += pr * (-2.f * d * p.inv_batch);
                                }
                                for (int k = 0; k < C; ++k) {
                                    const float pr = expf(zr[k] - gmax) * inv;
                                    const float d = (row_ok ? __ldg(tg + k) : pr) - pr;
                                    const float dzv = row_ok ? (pr * (-2.f * d * p.inv_batch) - pr * gs) : 0.f;
                                    if (row_ok) {
                                        gout[(int64_t)n * p.act_ld[l] + k] = zr[k];
                                        p.probs[(int64_t)(row0 + n) * p.ldp + k] = pr;
                                        if (p.dz[l] != nullptr) p.dz[l][(int64_t)(row0 + n) * p.act_ld[l] + k] = dzv;
                                        if (SPLIT && p.dz_lo[l] != nullptr) p.dz_lo[l][(int64_t)(row0 + n) * p.act_ld[l] + k] = tf32_lo(dzv);
                                    }
                                    if (p.do_bwd) st_tile(dst, n, k, dzv);
                                }
                            }
                            if (p.do_bwd)                               // features [C, 32) of the k-block must be finite zeros
                                for (int k = C; k < 32; ++k) st_tile(dst, n, k, 0.f);
                        }
                        loss = warp_sum_f(loss);
                        if (lane == 0 && p.loss != nullptr) p.loss[p.mu_base + blockIdx.x] = loss * p.inv_batch;
                    }
                    if (p.do_bwd) {
                        publish();
                        wbuf ^= 1;
                    }
                    signal_ready(l);                          // dz[L] written by the class-lane warp; the others arrive empty
                }
            }
        }
        if (FOLD && p.do_fwd && !p.do_bwd && p.out_peer != nullptr) {
            // the stage's output tile is in the next stage's receive slot: one fence, then its arrival flag
            asm volatile("bar.sync 2, 256;" ::: "memory");
            if (threadIdx.x == 64) {
                __threadfence_system();
                st_relaxed_sys(p.out_flag, pp_epoch);
            }
        }
        if (p.do_bwd && !(p.do_fwd && p.do_loss)) {
            // backward-only launch (pipeline stage): stage dZ_L = gout (.) relu'(act_L) into smem
            const ChainLayer& ly = p.layers[L - 1];
            const bool m_ok = m < ly.out;
            const uint32_t dst = abuf0 + wbuf * abuf_bytes;
            float* __restrict__ g = p.dz[L] + (int64_t)row0 * p.act_ld[L];
            const float* __restrict__ y = p.act[L] + (int64_t)row0 * p.act_ld[L];
            for (int n = c_lo; n < c_hi; ++n) {
                float x = 0.f;
                if (m_ok && n < p.mb_rows) {
                    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(x) : "l"(g + (int64_t)n * p.act_ld[L] + m));   // may be peer-written
                    if (ly.relu && !(y[(int64_t)n * p.act_ld[L] + m] > 0.f)) x = 0.f;
                    g[(int64_t)n * p.act_ld[L] + m] = x;     // the wgrad GEMM reads the masked gradient
                    if (SPLIT) p.dz_lo[L][(int64_t)(row0 + n) * p.act_ld[L] + m] = tf32_lo(x);
                }
                st_tile(dst, n, m, x);
            }
            publish();
            wbuf ^= 1;
        }
        if (p.do_bwd) {
            for (int l = L; l >= bwd_lo; --l) {
                const ChainLayer& ly = p.layers[l - 1];
                const bool m_ok = m < ly.in;                 // output feature of dgrad = input feature of layer l
                const bool mask_on = (l >= 2) && p.layers[l - 2].relu;
                const float* __restrict__ yprev = p.act[l - 1] + (int64_t)row0 * p.act_ld[l - 1];
                float* __restrict__ gprev = p.dz[l - 1] + (int64_t)row0 * p.act_ld[l - 1];
                // the masks do not depend on the MMA: fetch them while it runs (N <= 32: all of them)
                float mk0[16];
                const bool pre = (c_hi - c_lo == 16);        // N <= 32: one chunk per warp, prefetch its masks
                if (pre) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        mk0[j] = (mask_on && m_ok && c_lo + j < p.mb_rows) ? yprev[(int64_t)(c_lo + j) * p.act_ld[l - 1] + m] : 1.f;
                }
                mbar_wait(tmem_full_bar, tmem_waits & 1);
                ++tmem_waits;
                tc_fence_after();
                if (p.sync_debug) asm volatile("bar.sync 2, 256;" ::: "memory");   // racecheck aid, see kernels.h
                if (threadIdx.x == 64) DBG(2, 2 * (tmem_waits - 1));
                const uint32_t dst = abuf0 + wbuf * abuf_bytes;
                const bool more = (l > bwd_lo);              // another dgrad consumes this tile
                const bool single = (c_hi - c_lo == 16);
                float keep[16];
                for (int c = c_lo; c < c_hi; c += 16) {
                    float mk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int n = c + j;
                        if (pre) mk[j] = mk0[j];
                        else mk[j] = (mask_on && m_ok && n < p.mb_rows) ? yprev[(int64_t)n * p.act_ld[l - 1] + m] : 1.f;
                    }
                    float v[16];
                    tmem_ld16(taddr + c, v);   // backward reductions run over <= 128 output features: one accumulator
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int n = c + j;
                        float x = (mk[j] > 0.f) ? v[j] : 0.f;
                        x = m_ok ? x : 0.f;
                        if (more) st_tile(dst, n, m, x);
                        if (single) keep[j] = x;
                        else if (m_ok && n < p.mb_rows) {
                            gprev[(int64_t)n * p.act_ld[l - 1] + m] = x;
                            if (SPLIT) p.dz_lo[l - 1][(int64_t)(row0 + n) * p.act_ld[l - 1] + m] = tf32_lo(x);
                            if constexpr (FOLD) { if (l == 1 && p.out_peer != nullptr) p.out_peer[(int64_t)n * p.act_ld[0] + m] = x; }
                        }
                    }
                }
                if (more) {
                    publish();
                    wbuf ^= 1;
                }
                if (single && m_ok) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c_lo + j < p.mb_rows) {
                            gprev[(int64_t)(c_lo + j) * p.act_ld[l - 1] + m] = keep[j];
                            if (SPLIT) p.dz_lo[l - 1][(int64_t)(row0 + c_lo + j) * p.act_ld[l - 1] + m] = tf32_lo(keep[j]);
                            if constexpr (FOLD) { if (l == 1 && p.out_peer != nullptr) p.out_peer[(int64_t)(c_lo + j) * p.act_ld[0] + m] = keep[j]; }
                        }
                }
                signal_ready(l - 1);
            }
            if (FOLD && p.out_peer != nullptr && !p.first_stage) {
                // dz[0] is in the previous stage's receive slot: one fence, then its arrival flag
                asm volatile("bar.sync 2, 256;" ::: "memory");
                if (threadIdx.x == 64) {
                    __threadfence_system();
                    st_relaxed_sys(p.out_flag, pp_epoch);
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// =========================================================================== host side
const char* make_tmap_mn(CUtensorMap* map, const float* base, int inner, int outer, int ld);                 // tc_gemm.cu
const char* make_tmap_k(CUtensorMap* map, const float* base, int inner, int outer, int ld, int box_outer);   // tc_gemm.cu

// Shared-memory plan of the chain kernel for one micro-batch size: k-panels per ring stage, ring depth, total bytes.
// Pure host arithmetic (no CUDA calls) so it can be unit-tested without a GPU (tests/test_planning.py).
bool chain_budget(int mb_rows, bool split, int* kps_out, int* stages_out, int* smem_bytes_out) {
    const int n_pad = (mb_rows + 15) / 16 * 16;
    const int abuf_bytes = 4 * n_pad * 128;
    const int scratch_bytes = n_pad * kScratchLd * 4 + 256;
    const int budget = 222 * 1024 - (split ? 4 : 2) * abuf_bytes - scratch_bytes;
    int kps = 4, stage_bytes = 0, stages = 0;
    for (; kps >= 1; kps >>= 1) {            // biggest stage (fewest waits/commits) that still double-buffers
        stage_bytes = kps * ((int)kABytes + n_pad * 128) * (split ? 2 : 1);
        stages = budget / stage_bytes;
        if (stages >= 2) break;
    }
    if (kps < 1 || stages < 2) return false;
    if (stages > 8) stages = 8;
    *kps_out = kps;
    *stages_out = stages;
    *smem_bytes_out = stages * stage_bytes + (split ? 4 : 2) * abuf_bytes + 1024 + 8 * (2 * stages + 3) + 16 + scratch_bytes;
    return true;
}

static bool chain_smem_fits(int mb_rows, bool split) {
    int kps, stages, smem;
    return chain_budget(mb_rows, split, &kps, &stages, &smem);
}

bool chain_eligible(const ChainLayer* layers, int n_layers, int mb_rows, int out_dim, bool has_loss, bool split) {
    if (!chain_smem_fits(mb_rows, split)) return false;
    if (n_layers < 1 || n_layers > kChainMaxLayers || mb_rows < 1 || mb_rows > 128) return false;
    for (int l = 0; l < n_layers; ++l) {
        if (layers[l].out > 128) return false;
        if (l > 0 && layers[l].in > 128) return false;
    }
    if (has_loss && (out_dim > 32 || layers[n_layers - 1].out != out_dim)) return false;
    return true;
}

const char* chain_plan(ChainPlan* plan, const ChainParams& params, const float* x, int ldx, int total_rows, int n_mubatches,
                       const float* W_lo, const float* x_lo) {
    *plan = ChainPlan{};
    ChainParams& p = plan->p;
    p = params;
    const int L = p.n_layers;
    p.n_pad = (p.mb_rows + 15) / 16 * 16;
    p.split = (W_lo != nullptr) ? 1 : 0;
    const int nmaps = 2 * L + 1;
    const int halves = p.split ? 2 : 1;
    std::vector<CUtensorMap> host(nmaps * halves);
    for (int half = 0; half < halves; ++half) {
        const float* Wb = half ? W_lo : p.W;
        const float* xb = half ? x_lo : x;
        for (int l = 0; l < L; ++l) {
            const ChainLayer& ly = p.layers[l];
            if (const char* e = make_tmap_k(&host[half * nmaps + 2 * l], Wb + ly.w_off, ly.in, ly.out, ly.ldw, kBlockM)) return e;
            if (const char* e = make_tmap_mn(&host[half * nmaps + 2 * l + 1], Wb + ly.w_off, ly.in, ly.out, ly.ldw)) return e;
        }
        if (xb != nullptr) {
            if (const char* e = make_tmap_k(&host[half * nmaps + 2 * L], xb, p.layers[0].in, total_rows, ldx, p.n_pad)) return e;
        } else {
            host[half * nmaps + 2 * L] = host[half * nmaps];
        }
    }
    CUtensorMap* dev = nullptr;
    if (cudaMalloc(&dev, host.size() * sizeof(CUtensorMap)) != cudaSuccess) return "chain_plan: cudaMalloc failed";
    if (cudaMemcpy(dev, host.data(), host.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice) != cudaSuccess)
        return "chain_plan: tensor map upload failed";
    plan->maps_dev = dev;
    p.maps = dev;
    int kps = 0, stages = 0, smem = 0;
    if (!chain_budget(p.mb_rows, p.split != 0, &kps, &stages, &smem)) return "chain_plan: shared memory budget too small";
    p.kps = kps;
    p.stages = stages;
    plan->smem_bytes = smem;
    plan->grid = n_mubatches;
    return nullptr;
}

void chain_plan_free(ChainPlan* plan) {
    if (plan->maps_dev) cudaFree(plan->maps_dev);
    plan->maps_dev = nullptr;
}

template <bool SPLIT, bool FOLD, bool ACC>
static cudaError_t chain_configure_one() {
    return cudaFuncSetAttribute(mlp_chain_kernel<SPLIT, FOLD, ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

cudaError_t chain_configure() {
    cudaError_t e;
    if ((e = chain_configure_one<false, false, false>()) != cudaSuccess) return e;
    if ((e = chain_configure_one<false, true, false>()) != cudaSuccess) return e;
    if ((e = chain_configure_one<true, false, false>()) != cudaSuccess) return e;
    if ((e = chain_configure_one<true, true, false>()) != cudaSuccess) return e;
    if ((e = chain_configure_one<true, false, true>()) != cudaSuccess) return e;
    return chain_configure_one<true, true, true>();
}

cudaError_t chain_launch(const ChainPlan& plan, cudaStream_t stream) {
    const ChainParams& p = plan.p;
    const bool fold = p.in_flag != nullptr || p.out_peer != nullptr || p.x_from_global != 0;
    const bool acc = p.split && p.acc_split != 0;
#define SSB_CHAIN_LAUNCH(S, F, A) mlp_chain_kernel<S, F, A><<<plan.grid, kThreads, plan.smem_bytes, stream>>>(p)
    if (!p.split) {
        if (fold) SSB_CHAIN_LAUNCH(false, true, false); else SSB_CHAIN_LAUNCH(false, false, false);
    } else if (!acc) {
        if (fold) SSB_CHAIN_LAUNCH(true, true, false); else SSB_CHAIN_LAUNCH(true, false, false);
    } else {
        if (fold) SSB_CHAIN_LAUNCH(true, true, true); else SSB_CHAIN_LAUNCH(true, false, true);
    }
#undef SSB_CHAIN_LAUNCH
    return cudaGetLastError();
}

}  // namespace ssb
