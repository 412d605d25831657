// Host-side launch API of the hand-written sm_100a kernels (no torch dependency, so the
// .cu files compile stand-alone with nvcc and the C++ runtime can call them directly).
#pragma once
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

namespace ssb {

enum GemmMode : int { GEMM_FWD = 0, GEMM_DGRAD = 1, GEMM_WGRAD = 2 };

// One description for the three Linear GEMMs.  Every GEMM is issued "swapped": the
// FEATURE dimension sits on the tcgen05 M axis (128 TMEM lanes) and the skinny dimension
// (micro-batch rows: 4..256) on the N axis, so a 4-row micro-batch costs an N=16 MMA
// instead of a 128-row tile that is 97 % padding.
//
//   FWD   D[m=out, n=row] = sum_k W[m,k]  * X[n,k]      A=W   (K-major)  B=X  (K-major)
//   DGRAD D[m=in,  n=row] = sum_k W[k,m]  * dZ[n,k]     A=W   (MN-major) B=dZ (K-major)
//   WGRAD D[m=out, n=in ] = sum_k dZ[k,m] * X[k,n]      A=dZ  (MN-major) B=X  (MN-major)
struct GemmParams {
    int m_total, n_total, k_total;
    int block_n;        // UMMA N: multiple of 16 (32 when B is MN-major), <= 256
    int stages;         // smem pipeline depth
    // FWD / DGRAD epilogue: out[n * ldo + m]
    float* out;
    int ldo;
    const float* bias;  // FWD: + bias[m * bias_stride] (nullable)
    int bias_stride;
    int relu;           // FWD: max(.,0)
    const float* mask;  // DGRAD: zero where mask[n * ldmask + m] <= 0 (nullable) - the ReLU
    int ldmask;         //        backward of the PREVIOUS layer fused into this epilogue
    // WGRAD epilogue: G[m * ldg + n] (+)= D ; db[m * db_stride] (+)= sum_k dZ[k, m]
    float* G;
    int ldg;
    int accumulate;
    float* db;          // nullable
    int db_stride;
    // WGRAD with the SGD update fused (single replica, not combinable with accumulate):
    // the -lr-scaled tile is TMA-reduce-added into W; G is never touched
    float* W;
    int ldw;
    float lr;
    int fuse_sgd;
    int acc_split;      // 3xTF32: small cross terms in their own accumulator + rotating main accumulators (ptx.cuh); 0 = one accumulator
    // fp32-equivalent mode (3xTF32): both operands come with a `lo` twin (x - trunc_tf32(x)); FWD/DGRAD
    // also emit the lo twin of their output so the next GEMM can consume it
    int split;
    float* out_lo;
    // split-K (FWD / DGRAD, see gemm_plan_enable_splitk): blockIdx.z owns a k-range, partial tiles go through
    // `partial` [tile][split][block_n][128], the last arriver at tile_counter[tile] reduces + runs the epilogue
    int k_splits;                 // 0 / 1 = off
    float* partial;
    unsigned int* tile_counter;   // zero before the first launch; re-armed by the kernel
};

struct GemmPlan {          // a fully prepared launch (tensor maps are 128 B each)
    CUtensorMap tmA, tmB, tmC;   // C: WGRAD output tile (G, or W when the SGD update is fused)
    CUtensorMap tmAlo, tmBlo;    // lo twins of the operands (split mode)
    GemmParams p;
    int mode;
    dim3 grid;
    int smem_bytes;
};

// Build tensor maps + launch geometry.  All matrices are fp32 row-major with a leading
// dimension that is a multiple of 4 floats and a 16-byte aligned base.
//   FWD:   W[out, in] (ldw), X[rows, in] (ldx)            -> Y[rows, out] (ldy)
//   DGRAD: W[out, in] (ldw), dZ[rows, out] (lddz)         -> dX[rows, in] (lddx)
//   WGRAD: dZ[rows, out] (lddz), X[rows, in] (ldx)        -> G[out, in] (ldg)
// *_lo pointers: nullptr = plain TF32; non-null = 3xTF32 (twins share the layout of their tensor)
struct GemmLo {
    const float* A = nullptr;   // lo twin of the first operand  (W for fwd/dgrad, dZ for wgrad)
    const float* B = nullptr;   // lo twin of the second operand (X for fwd, dZ for dgrad, X for wgrad)
    float* out = nullptr;       // fwd/dgrad: where to write the lo twin of the output
};
const char* gemm_plan_fwd(GemmPlan* plan, const float* W, int ldw, const float* X, int ldx, float* Y, int ldy,
                          int rows, int in, int out, const float* bias, int bias_stride, int relu, GemmLo lo = GemmLo());
const char* gemm_plan_dgrad(GemmPlan* plan, const float* W, int ldw, const float* dZ, int lddz, float* dX, int lddx,
                            int rows, int in, int out, const float* mask, int ldmask, GemmLo lo = GemmLo());
const char* gemm_plan_wgrad(GemmPlan* plan, const float* dZ, int lddz, const float* X, int ldx, float* G, int ldg,
                            int rows, int in, int out, int accumulate, float* db, int db_stride, float* W, int ldw,
                            float lr, int fuse_sgd, GemmLo lo = GemmLo());
// Split-K for FWD / DGRAD plans whose output has fewer tiles than the chip has SMs (wide, weight-bound layers):
// choice() returns the number of k-splits (1 = leave the plan alone); enable() rewires the plan (grid.z, ring
// depth for two CTAs per SM).  workspace: gemm_splitk_workspace_floats() floats; counters: one zeroed uint per tile.
int gemm_splitk_choice(const GemmPlan& plan, int num_sms);
size_t gemm_splitk_workspace_floats(const GemmPlan& plan, int k_splits);
const char* gemm_plan_enable_splitk(GemmPlan* plan, int k_splits, float* workspace, unsigned int* counters);
cudaError_t gemm_launch(const GemmPlan& plan, cudaStream_t stream);
// Grouped launch of several WGRAD plans (different layers) as ONE grid: a device table holds one entry per CTA.
struct alignas(128) GemmGroupEntry {
    CUtensorMap tmA, tmB, tmC, tmAlo, tmBlo;
    GemmParams p;
    int bx, by;                    // tile coordinates of this CTA inside its GEMM
};
struct GemmGroupPlan {
    GemmGroupEntry* entries_dev;
    int n, smem_bytes, n_gemms;
};
const char* gemm_group_plan(GemmGroupPlan* out, const GemmPlan* plans, int n_plans);
cudaError_t gemm_group_launch(const GemmGroupPlan& plan, cudaStream_t stream);
void gemm_group_free(GemmGroupPlan* plan);
cudaError_t gemm_configure();   // opt every instantiation into > 48 KB dynamic smem (call outside graph capture)
int gemm_kernel_count();   // number of launches issued so far by this module (bench accounting)

// ---- fused data-parallel path (csrc/kernels/fused_dp.cu) ---------------------------------
static constexpr int kMaxDp = 8;
struct DpPeers {                 // symmetric-memory pointer tables (index = DP rank)
    float* W[kMaxDp];            // weight arenas (owners publish updated tiles into every replica)
    float* stage[kMaxDp];        // partial-gradient staging: [src][layer slots]
    uint32_t* arrive[kMaxDp];    // arrive[owner][src * slots_per_src + slot] = epoch when src's partial landed
    uint32_t* done[kMaxDp];      // done[rank][tile] = epoch when the tile's new weights landed on rank
};
struct DpLayerParams {
    int m_total, n_total, k_total;   // out, in, micro-batch rows
    int block_n, stages;
    int n_tiles_m, n_tiles_n;
    int dp, rank;
    int64_t w_offset;                // float offset of the layer's [out, ld] block in the arena
    int ldw;
    int64_t stage_offset;            // float offset of the layer's slots inside one src region
    int64_t stage_src_stride;        // floats per src region
    int tile_flag_base, slot_flag_base, slots_per_src;
    float lr;
    const uint32_t* epoch_ptr;       // device step counter (bumped once per step, same value on all ranks)
    const float* G;                  // dp_reduce_sgd: accumulated gradient block to reduce
    int ldg;
    // one_shot: small (latency-bound) layers - every replica pushes its partial tile to ALL replicas
    // and every replica reduces every tile itself in the same fixed rank order (bit-identical results,
    // one NVLink hop instead of two).  Staging is double-buffered by epoch parity (stage_parity_stride).
    int one_shot;
    int64_t stage_parity_stride;
    int split;                       // 3xTF32: lo twins of dZ / X are loaded too
    int helpers;                     // CTAs per tile: CTA 0 computes + pushes, all of them share the reduce rows
    int bulk_push;                   // 1: push tiles with cp.async.bulk (TMA engine), 0: coalesced st.global
    unsigned long long* dbg;         // optional: 8 globaltimer stamps of CTA 0 (phase timeline)
};
struct FusedDpPlan {
    CUtensorMap tmA, tmB, tmAlo, tmBlo;
    DpLayerParams p;
    DpPeers peers;
    int grid;
    int smem_bytes;
};
void dp_layer_geometry(int in, int out, int dp, int* block_n, int* n_tiles_m, int* n_tiles_n, int64_t* slots,
                       int64_t* slot_floats, int one_shot);
// dZ == nullptr: plan for dp_reduce_sgd (no GEMM)
const char* fused_dp_plan(FusedDpPlan* plan, const float* dZ, int lddz, const float* X, int ldx, int rows,
                          const DpLayerParams& lp, const DpPeers& peers, int max_ctas, const float* dZ_lo = nullptr,
                          const float* X_lo = nullptr);
cudaError_t fused_dp_configure();
cudaError_t launch_fused_wgrad_dp(const FusedDpPlan& plan, cudaStream_t stream);
cudaError_t launch_dp_reduce_sgd(const FusedDpPlan& plan, cudaStream_t stream);
cudaError_t launch_bump_epoch(uint32_t* epoch, cudaStream_t stream);

// ---- LL two-shot fused DP path for narrow layers (csrc/kernels/dp_ll.cu): all layers of a stage in ONE launch
struct DpLLEntry {               // one CTA = one [128 x 32] tile of one layer's weight gradient
    CUtensorMap tmA, tmB, tmAlo, tmBlo;   // dZ (MN-major A), X (MN-major B) and their lo twins
    int m0, n0;                  // tile origin in (out, in)
    int m_total, n_total, k_total;
    int split, has_bias;
    int tile;                    // index in the landing zones
    int64_t w_offset;            // float offset of the layer's [out, ld] block in the arena
    int ldw;
    const uint32_t* gate_flag;   // chain-kernel counter the GEMM waits for: dz[l] final (nullptr: ordered by the stream instead)
    const uint32_t* gate_w;      // counter the in-place weight update waits for: dgrad of layer l consumed W_l (nullptr: none)
    uint32_t gate_mult;
};
struct DpLLParams {
    int dp, rank;
    float lr;
    const uint32_t* epoch_ptr;   // DP step counter (same value on every replica)
    const uint32_t* gate_step;   // steps of the launching engine (gate target = gate_mult * *gate_step)
    int n_tiles, stages;
    float* W;                    // local weight arena
    float* W_lo;                 // optional: lo twins of the weights (3xTF32), refreshed by whoever stores a weight
    uint4* llA[kMaxDp];          // reduce-scatter landing zones  [parity][src][tile][128 / dp rows][17 lines]
    uint4* llC[kMaxDp];          // all-gather landing zones      [parity][tile][128 rows][17 lines]
    unsigned long long* dbg;     // optional phase timeline: 8 globaltimer stamps per tile (first 28 tiles)
};
struct DpLLLayer {
    const float *dZ, *X, *dZ_lo, *X_lo;
    int lddz, ldx, in, out, ldw;
    int64_t w_offset;
    const uint32_t* gate_flag;
    const uint32_t* gate_w;
    uint32_t gate_mult;
};
struct DpLLPlan {
    DpLLParams p;
    DpLLEntry* entries_dev;
    int grid, smem_bytes;
};
int acc_split_default();   // tc_gemm.cu: SSB_ACC_SPLIT (default 1)
int dp_ll_tiles(int in, int out);
size_t dp_ll_zone_lines(int dp, int n_tiles);
const char* dp_ll_plan(DpLLPlan* plan, const DpLLLayer* layers, int n_layers, int rows, const DpLLParams& base);
void dp_ll_free(DpLLPlan* plan);
cudaError_t dp_ll_configure();
cudaError_t launch_dp_ll(const DpLLPlan& plan, cudaStream_t stream);

// ---- layer-chain kernel (csrc/kernels/mlp_chain.cu) --------------------------------------
static constexpr int kChainMaxLayers = 16;
struct ChainLayer {
    int in, out, relu, ldw;
    int64_t w_off;                   // float offset of the [out, ld] block in the weight arena
};
struct ChainParams {
    int n_layers;
    ChainLayer layers[kChainMaxLayers];
    const CUtensorMap* maps;         // device array: [2l] W_l K-major (fwd), [2l+1] W_l MN-major (dgrad), [2L] X;
                                     // split mode: the same 2L+1 maps for the lo twins follow at [2L+1 ..]
    int split;                       // 3xTF32 (fp32-equivalent) products
    float* act_lo[kChainMaxLayers + 1];   // lo twins written next to act / dz (consumed by the wgrad GEMMs)
    float* dz_lo[kChainMaxLayers + 1];
    const float* W;                  // weight arena (bias reads)
    float* act[kChainMaxLayers + 1]; // act[0] = stage input, act[l] = output of layer l  ([all rows, ld])
    float* dz[kChainMaxLayers + 1];  // dz[l] = gradient w.r.t. the pre-activation of layer l (dz[0]: stage input grad)
    int act_ld[kChainMaxLayers + 1];
    const float* target;             // [all rows, ldt] one-hot targets (last stage)
    int ldt;
    float* probs;                    // [all rows, ldp]
    int ldp;
    float* loss;                     // [n_mubatches]
    int mb_rows, n_pad, stages;
    int kps;                         // 32-wide k-blocks per pipeline stage (one wait / one commit per stage)
    int mu_base;                     // first micro-batch handled by CTA 0 (per-micro-batch launches)
    float inv_batch;
    int do_fwd, do_loss, do_bwd, first_stage;
    // Pipeline boundaries folded into the kernel (peer-memory transport, one launch = one micro-batch of a stage):
    //  * in_flag  != nullptr: the tile this launch consumes (activations of the previous stage for a forward launch, output
    //    gradients of the next stage for a backward launch) is final in local memory once *in_flag >= *pp_epoch; the
    //    epilogue warps wait for it and stage the tile into shared memory - weights stream in meanwhile;
    //  * x_from_global: forward launch of a stage > 0: the input tile is read from act[0] by the epilogue warps (after the
    //    flag) instead of by TMA next to layer 1's weights, its lo twin is written to act_lo[0] on the way;
    //  * out_peer != nullptr: the tile this launch produces for the neighbour (last forward output / dz[0]) is ALSO stored
    //    into the neighbour's receive slot, then one system fence + *out_flag = epoch; *out_credit >= epoch - 1 (the
    //    neighbour released its slots of the previous step) is awaited first.
    const uint32_t* in_flag;
    const uint32_t* pp_epoch;
    int x_from_global;
    float* out_peer;
    uint32_t* out_flag;
    const uint32_t* out_credit;
    int acc_split;                   // 3xTF32: launch the instantiation with separate + rotating accumulators for long reductions (SSB_CHAIN_ACC=1)
    int sync_debug;                  // SSB_RACECHECK=1: an explicit named barrier among the epilogue warps per layer, so that
                                     // compute-sanitizer's racecheck (which cannot see tcgen05.commit -> mbarrier ordering) can
                                     // verify the reuse of the activation ping-pong tiles
    uint32_t* ready;                 // optional [n_layers + 1] device counters: every epilogue warp of every CTA adds 1 to
                                     // ready[l] once its part of dz[l] (and everything it wrote before) is globally visible
    unsigned long long* dbg;         // optional timeline buffer (3 roles x 256 globaltimer stamps), CTA 0 only
};
struct ChainPlan {
    ChainParams p;
    CUtensorMap* maps_dev;
    int grid, smem_bytes;
};
bool chain_eligible(const ChainLayer* layers, int n_layers, int mb_rows, int out_dim, bool has_loss, bool split = false);
const char* chain_plan(ChainPlan* plan, const ChainParams& params, const float* x, int ldx, int total_rows, int n_mubatches,
                       const float* W_lo = nullptr, const float* x_lo = nullptr);
bool chain_budget(int mb_rows, bool split, int* kps, int* stages, int* smem_bytes);   // host arithmetic only
void chain_plan_free(ChainPlan* plan);
cudaError_t chain_configure();
cudaError_t chain_launch(const ChainPlan& plan, cudaStream_t stream);

// ---- small fused kernels --------------------------------------------------------------
// logits[rows, cols] (ld) -> probs (nullable) ; training: dlogits (softmax-Jacobian x MSE
// grad, 1/batch_size inside) and loss_out[0] = sum((t-p)^2)/batch_size.
cudaError_t launch_loss_head(const float* logits, int ldl, const float* target, int ldt, float* probs, int ldp,
                             float* dlogits, int ldd, float* loss_out, int rows, int cols, float inv_batch,
                             cudaStream_t stream, int rows_per_mubatch = 0, float* dlogits_lo = nullptr);   // >0: one CTA per micro-batch, loss_out[mu]
// generic softmax backward for the functional API: dz = p*up - p*sum(p*up)
cudaError_t launch_softmax_grad(const float* logits, int ldl, const float* upstream, int ldu, float* dlogits, int ldd,
                                int rows, int cols, cudaStream_t stream);
// g[r, c] = y[r, c] > 0 ? g[r, c] : 0   (stage-boundary ReLU backward)
cudaError_t launch_relu_mask(float* g, int ldg, const float* y, int ldy, int rows, int cols, cudaStream_t stream,
                             float* g_lo = nullptr);   // g_lo: recomputed lo twin of the masked gradient
cudaError_t launch_relu_fwd(const float* x, float* y, long n, cudaStream_t stream);
// y = a * x + b * t  (mse grad: a=2/B, b=-2/B)
cudaError_t launch_axpby(const float* x, const float* t, float* y, float a, float b, long n, cudaStream_t stream);
// lo[i] = x[i] - trunc_tf32(x[i]) over a flat buffer (weights after the update, staged inputs)
cudaError_t launch_split_lo(const float* x, float* lo, long n, cudaStream_t stream);
// w -= lr * g over a flat arena
cudaError_t launch_sgd(float* w, const float* g, float lr, long n, cudaStream_t stream);
// correct[0] += #rows with argmax(pred) == argmax(target)
cudaError_t launch_argmax_correct(const float* pred, int ldp, const float* target, int ldt, int rows, int cols,
                                  int* correct, cudaStream_t stream);

}  // namespace ssb
