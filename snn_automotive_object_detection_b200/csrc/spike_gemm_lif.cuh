// spike_gemm_lif: the contraction of a spiking layer with its LIF recurrence fused
// into the epilogue, for sm_100a (tcgen05 + TMEM + TMA).
//
// Replaces, per layer, the reference's per-timestep
//     cur = conv/linear(z_t) ; spk_t, state = LIFCell(cur, state)
// (rpn.py:105-106, faster_rcnn.py:498-501) by ONE launch:
//
//   D[c, (t, u)] = sum_k  Wsplit[c, k] * Z[t, u, k]         (tcgen05.mma, bf16|fp16 x {0,1} -> fp32 in TMEM)
//   for every neuron (u, c):  run the LIF recurrence over t in registers, emit its spike train
//
//  * A operand  = weights [M rows = output channels][K] 16-bit, K-major, TMA 2-D tiles 128 x 64.
//    "fp32-exact" mode keeps 2 or 3 bf16 pieces of every weight (hi/mid/lo); because the other
//    operand is exactly {0,1} every product is exact and the pieces are simply extra k-steps.
//    fp16 modes keep 1 or 2 fp16 pieces of the power-of-two row-scaled weight (11 bits each) and
//    the epilogue multiplies the accumulator row by the inverse scale (exact).
//  * B operand  = input spikes.  They live in HBM ONLY as time-packed spike-train words (one word
//    per input neuron, bit t = spike at step t: the encoder's output, or the previous layer's
//    epilogue output) -- 1-2 bytes per neuron instead of one 16-bit plane per timestep.  The
//    producer warps expand the words of a k-block into the 128-byte-swizzled K-major tile the
//    tensor core reads (rows ordered (t, unit): ALL timesteps of a unit sit in the same
//    accumulator tile, time folded into the MMA N dimension, N = T_box * J <= 256):
//      fc   : words [R][K]            TMA box (64 words, Jh rows) per k-block; one spike tile per k-block.
//      conv : words [N][H][W][C]      ONE TMA box (64 words, 8+2, TH/kCG+2, 1) per (tile, 64-channel block): the
//             CTA's pixels plus a one-pixel halo (image border = TMA out-of-bounds zero fill).  It is expanded
//             ONCE into a halo'd spike tile with rows ordered (halo row, t, halo column); the 9 taps of the
//             3x3 conv are 9 MMA descriptors into that same tile: start row (dy * T_box * 10 + dx), 8-row
//             groups 10 rows (1280 B) apart.  This relies on the 128-byte swizzle of tcgen05.mma following
//             ABSOLUTE shared-memory address bits (scratch/swizzle_experiment.py: any 128-B row start and
//             group stride works with base_offset 0), so the producers swizzle by the absolute row address.
//    The word tile of a stage is 1-3 KB per CTA, so the L2->SM feed of the kernel is the weight tiles alone,
//    and the expansion work of the conv is 40/16 halo overhead x 1/9 = 0.28 of expanding every tap.
//  * accumulators: 2 x 256 TMEM columns.  conv / short fc: double buffered -> the MMA of tile i+1 overlaps
//    the LIF epilogue of tile i.  Long fc (kDual): one tile = 2J units, both buffers fed from every weight
//    tile (half the L2 -> SM weight stream and half the cross-CTA hand-offs per FLOP).
//    The LIF state (v, i) never leaves registers; nothing of size T x state is ever written to HBM.
//    Output per neuron: one time-packed spike-train word (bit t = spike at step t; popc = spike count),
//    which is also the next layer's B operand.
//  * More than 16 live steps: the host cuts the time axis into passes (one launch each) and the neuron
//    state is carried between them as one float4 per neuron (GemmLifParams::state).
//  * kCG = 2 pairs two SMs (cta_group::2, UMMA M = 256): each CTA owns 128 output channels and
//    produces half of the unit tile.
//  * Profiling: GemmLifParams::role_cycles (snn_set_role_timers) collects, per CTA pair, where the
//    MMA-issuing thread and one epilogue warp spend their cycles and the kernel's entry-to-exit time in
//    SM cycles and nanoseconds.
#pragma once
#include <cuda_bf16.h>

#include <type_traits>

#include "ptx.cuh"

namespace snn {

constexpr int kMaxLevels = 8;
constexpr int kStagesA = 8;                 // weight ring: up to 8 stages of 16 KB (p.stages_a)
constexpr int kMaxStagesB = 8;              // ring of p.stages_b slots of p.slot_b bytes; weight + spike rings share 176 KB
constexpr int kMaxStagesW = 8;              // ring of p.stages_w slots of p.slot_w bytes, 16 KB in total
constexpr int kTileBytesA = 128 * 128;      // 128 rows x 64 16-bit
constexpr int kRingBytesAB = 193 * 1024;        // the weight ring (stages_a x 16 KB) and the spike-tile ring share it; the CTA uses all 227 KB
constexpr int kRingBytesW = 16 * 1024;
constexpr int kBarBytes = 512;
constexpr int kEpiGroups = 1;               // LIF epilogue warp groups (4 warps each): warps 4-7 (+ 16-19)
constexpr int kGemmThreads = 512 + (kEpiGroups - 1) * 128;   // warps 0-3 control, 4-7 epilogue, 8-15 spike-tile producers
constexpr int kProducerWarps = 8;           // split into p.n_pg groups (1, 2 or 4); group g expands the k-blocks i = g (mod n_pg)
constexpr int kMaxPairs = 4;                // (unit, 8-channel chunk) pairs a producer thread expands per k-block
constexpr int kMaxUnitsPerCta = kMaxPairs * (kProducerWarps * 32 / 4) / 8;   // Jh <= 32: Jh * 8 pairs <= 4 x 64 threads
constexpr int kRoMaxOut = 16;               // fused readout: objectness + box deltas per pixel (5 * A <= 16)
constexpr int kRoWStride = 132;             // floats per readout-weight row in smem (128 + pad, 16-B aligned)
constexpr int kRoSmemBytes = kRoMaxOut * kRoWStride * 4 + kEpiGroups * 2 * 8 * kRoWStride * 4;   // weights + per epilogue group 2 x [8 px][128 ch] sums
constexpr size_t kGemmSmemBytes =
    1024 /*align slack*/ + kRingBytesAB + kRingBytesW + kBarBytes + kRoSmemBytes;

struct LevelDesc {
    int H, W, tiles_w, tiles_h;
    int tile_begin;           // first unit-tile index of this level
    int pad_;
    void* trains;             // [N][H][W][m_total] spike-train words (nullable when the readout is fused)
    float* logits;            // fused readout: [N][A][H][W]   (zero-initialised; 2 CTAs add their channel halves)
    float* bbox;              // fused readout: [N][4A][H][W]
    unsigned long long* counts;   // nullable: [N] spikes per image of this level
    float4* state;                // multi-pass only: [N][H][W][m_total] (v, i, kappa-weighted spike sum, train word bits)
};

struct GemmLifParams {
    CUtensorMap tmA;
    CUtensorMap tmA_half;          // kMC: the same weight tensor with a box of 64 rows (each CTA fetches half of its tile and multicasts it)
    CUtensorMap tmW[kMaxLevels];   // input spike-train words as byte tensors: conv [N][H][W][k_in*in_wb], fc [rows][k_in*in_wb]
    LevelDesc lv[kMaxLevels];
    int n_levels, conv, n_images;
    int m_total, m_tiles, nsplit;
    int kblocks, cblocks;
    int k_in;                 // input neurons per unit (conv: channels; fc: K)
    int T_total, t0, T_live, T_box;
    int J, Jh, TWh, THh, TW, TH, sub_dw, sub_dh;
    int rows;                 // fc: number of units (RoIs)
    int total_tiles, unit_tiles;
    int train_bytes;          // output word size: 1, 2 or 4
    int in_wb, in_bit0;       // input word size; bit of the input word that is step t0 of this layer
    int stages_a;             // weight ring stages (<= kStagesA)
    int stages_b, slot_b;     // B ring geometry
    int stages_w, slot_w;     // word ring geometry (slot = Jh units x 64 words)
    int hrows;                // conv: halo rows per CTA = TH / kCG + 2 (halo columns = TWh + 2 = 10)
    int n_pg;                 // producer groups (each owns every n_pg-th k-block): min(4, stages_b, stages_w) rounded to 1/2/4
    int n_mma;                // T_box * J
    int dual;                 // fc only: one tile = 2J units, BOTH accumulator buffers fed from every weight tile
    uint32_t idesc;
    void* trains;             // fc: [rows][m_total]
    uint32_t spike_one;       // 1.0 as bf16 (0x3F80) or fp16 (0x3C00)
    const float* w_scale;     // [m_total] power of two each accumulator row is multiplied with (1 for bf16 pieces)
    // More than 16 live steps do not fit one accumulator tile with a useful number of units: the time axis is then cut
    // into passes of <= 16 steps, one launch each, and the neuron state (v, i, kappa-weighted spike sum, train word)
    // is carried between the launches through `state` (fc: [rows][m_total]; conv: per level).  A pass with
    // state_store set stops after its live steps and writes the state; only the last pass drains and emits.
    int state_load, state_store;
    float4* state;
    float* dump;              // debug (fc only): raw currents [T_live][dump_rows][m_total]
    int dump_rows;
    // profiling only (nullable): per CTA pair 8 counters of the MMA-issuing thread, in SM clock cycles:
    // [0] whole role, [1] waiting for a free accumulator, [2] for this CTA's spike-tile half, [3] for the peer's
    // half, [4] for weight tiles, [5] tiles; [6] the epilogue role (warp 4 of the leader CTA), [7] of which waiting
    // for a full accumulator
    unsigned long long* role_cycles;
    // measurement only (nullable): [pairs][2]; every launch ADDS the leader CTA's kernel entry-to-exit time in SM
    // cycles ([0]) and in globaltimer nanoseconds ([1]) -> the SM clock the kernel really ran at, taken on the very
    // launches a bench times with CUDA events (two timer reads per CTA pair: no effect on the measurement)
    unsigned long long* clock_probe;
    int wait_backoff_ns;      // conv producers: sleep between polls of the spike-tile stage they wait for (0 = none)
    int prod_word_reuse;      // fc producers: keep a pair's words in registers across its steps (1; 0 = reload per item, for A/B runs)
    int dbg_shift, dbg_sbo, dbg_boff;   // swizzle experiment (scratch/swizzle_experiment.py, fc only): row shift, group stride, base-offset field
    // fused leaky-integrator readout (conv, cta_group 2, m_total == 256): mem_{T-1} = W . sum_t kappa_{T-1-t} spk_t
    int fuse_readout, A;
    const float* w_cls;       // [A][m_total]
    const float* w_bbox;      // [4A][m_total]
    float kappa[32];          // kappa[t] = 0.9^{T-t} - 0.8^{T-t}: weight of a spike at step t in the last membrane
};

// One LIF step of Norse's lif_feed_forward_step, op for op (no FMA contraction):
//   v_dec = v + 0.1f*((0 - v) + i);  i_dec = i + (-0.2f)*i;  z = (v_dec - 0.1f > 0);
//   v' = (1-z)*v_dec + z*0;  i' = i_dec + cur
// The threshold test is evaluated as v_dec > 0.1f, the same bit as fl(v_dec - 0.1f) > 0 (see encode_train).
__device__ __forceinline__ bool lif_update(float& v, float& i, float cur) {
    const float dv = __fmul_rn(0.1f, __fsub_rn(i, v));
    const float v_dec = __fadd_rn(v, dv);
    const float i_dec = __fadd_rn(i, __fmul_rn(-0.2f, i));
    const bool z = v_dec > 0.1f;
    v = z ? 0.0f : v_dec;
    i = __fadd_rn(i_dec, cur);
    return z;
}
// The same with the input current given as (raw accumulator, power-of-two row scale): acc * scale is exact, so the one
// rounding of fma(acc, scale, i_dec) is the rounding of i_dec + cur -- bit-identical, one instruction less per step.
__device__ __forceinline__ bool lif_update_scaled(float& v, float& i, float acc, float scale) {
    const float dv = __fmul_rn(0.1f, __fsub_rn(i, v));
    const float v_dec = __fadd_rn(v, dv);
    const float i_dec = __fadd_rn(i, __fmul_rn(-0.2f, i));
    const bool z = v_dec > 0.1f;
    v = z ? 0.0f : v_dec;
    i = __fmaf_rn(acc, scale, i_dec);
    return z;
}
// named barriers of the fused readout (conv): the epilogue warps hand the kappa-weighted spike sums of an 8-pixel chunk
// to the readout warp through two shared-memory buffers
constexpr int kBarRoFull = 2;               // + buffer: chunk written (128 epilogue threads arrive, the readout warp syncs)
constexpr int kBarRoFree = 4;               // + buffer: chunk consumed (the readout warp arrives, the epilogue threads sync)
constexpr int kRoBarThreads = 128 + 32;
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Decoded position of a unit tile.
struct TilePos { int lvl, n, h0, w0; };

__device__ __forceinline__ TilePos decode_tile(const GemmLifParams& p, int ut) {
    TilePos tp{0, 0, 0, 0};
    if (p.conv) {
        while (tp.lvl + 1 < p.n_levels && ut >= p.lv[tp.lvl + 1].tile_begin) ++tp.lvl;
        const LevelDesc& L = p.lv[tp.lvl];
        int local = ut - L.tile_begin;
        const int per_img = L.tiles_w * L.tiles_h;
        tp.n = local / per_img; local -= tp.n * per_img;
        const int ty = local / L.tiles_w;
        tp.h0 = ty * p.TH; tp.w0 = (local - ty * L.tiles_w) * p.TW;
    }
    return tp;
}

// kDual (fc only): a tile covers 2J units; the CTA of rank r holds the units [2J*ut + 2Jh*r, +2Jh), the first Jh of
// them accumulate in TMEM buffer 0, the next Jh in buffer 1, and every weight tile feeds both MMAs -- half the
// L2 -> SM weight stream and half the producer -> relay -> MMA hand-offs per FLOP; the (short) fc epilogue of a tile
// then no longer overlaps the next tile's main loop.
// kMC (conv, cta_group 2): clusters of TWO CTA pairs.  Both pairs walk the same weight sequence (every tile does), so each
// weight tile is read from L2 once per cluster: every CTA fetches half of its 128-row tile and multicasts it to its
// counterpart in the other pair (`.multicast::cluster`), a ring stage is refilled when BOTH pairs' MMAs have retired it
// (a_empty counts two commits, each multicast to the four CTAs), and the pair with one tile less runs a dummy tile with
// its outputs suppressed so that the two stay in lock-step.  Everything else (spike tiles, accumulators, epilogue,
// readout) stays per pair.
template <int kCG, int CW, bool kConv, bool kDual = false, bool kMC = false>
__global__ void __launch_bounds__(kGemmThreads, 1) spike_gemm_lif_kernel(const __grid_constant__ GemmLifParams p) {
    static_assert(!kMC || (kCG == 2 && kConv && !kDual), "weight multicast: conv tiles on CTA pairs only");
    extern __shared__ uint8_t smem_raw[];
    const long long k_begin = clock64();          // profiling only (role_cycles)
    unsigned long long k_begin_ns = 0;
    if ((p.role_cycles != nullptr || p.clock_probe != nullptr) && threadIdx.x == 64)
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(k_begin_ns));
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_ring = smem;
    const int stages_a = p.stages_a;
    uint8_t* b_ring = smem + stages_a * kTileBytesA;
    uint8_t* w_ring = smem + kRingBytesAB;
    uint64_t* bars = reinterpret_cast<uint64_t*>(w_ring + kRingBytesW);
    uint64_t* a_full = bars;                         // [kStagesA]   weight tile landed (TMA tx, leader CTA)
    uint64_t* a_empty = a_full + kStagesA;           // [kStagesA]   MMAs reading it retired
    uint64_t* b_ready = a_empty + kStagesA;          // [kMaxStagesB] this CTA's half of the spike tile written
    uint64_t* b_peer = b_ready + kMaxStagesB;        // [kMaxStagesB] leader only: the peer CTA's half written
    uint64_t* b_empty = b_peer + kMaxStagesB;        // [kMaxStagesB] MMAs reading the spike tile retired
    uint64_t* w_full = b_empty + kMaxStagesB;        // [kMaxStagesW] input words landed (TMA tx)
    uint64_t* w_empty = w_full + kMaxStagesW;        // [kMaxStagesW] producers have read them
    uint64_t* acc_full = w_empty + kMaxStagesW;      // [2]
    uint64_t* acc_empty = acc_full + 2;              // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* ro_w = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kBarBytes);   // [kRoMaxOut][kRoWStride]
    float* ro_s = ro_w + kRoMaxOut * kRoWStride;                                              // [2][CW][kRoWStride]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t crank = (kCG == 2) ? cluster_ctarank() : 0u;      // rank in the cluster (0..3 with kMC)
    const uint32_t rank = crank & 1u;                                // rank inside the CTA pair
    const uint32_t leader_cta = crank & ~1u;                         // cluster rank of this pair's leader
    const uint16_t pair_mask = static_cast<uint16_t>(0b11u << leader_cta);
    const int n_groups = gridDim.x / kCG;
    const int group = blockIdx.x / kCG;
    const int stages_b = p.stages_b;
    const int stages_w = p.stages_w;

    if (warp == 0 && elect_one()) { tma_prefetch_desc(&p.tmA); if constexpr (kMC) tma_prefetch_desc(&p.tmA_half); }
    if (warp == 3 && elect_one())
        for (int l = 0; l < p.n_levels; ++l) tma_prefetch_desc(&p.tmW[l]);
    if (warp == 1 && elect_one()) {
        for (int s = 0; s < kStagesA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], kMC ? 2 : 1); }
        const uint32_t wpg = static_cast<uint32_t>(kProducerWarps / p.n_pg);      // warps per producer group
        for (int s = 0; s < kMaxStagesB; ++s) { mbar_init(&b_ready[s], wpg); mbar_init(&b_peer[s], 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < kMaxStagesW; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], wpg); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4 * kEpiGroups * kCG); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<kCG>(tmem_slot, 512);
    tcgen05_fence_before();
    if constexpr (kCG == 2) cluster_sync_all(); else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // barriers, TMEM and descriptor prefetch are set up: let the next kernel of the stream be scheduled as soon as SMs
    // free up, and wait for the preceding kernel's results (the spike-train words / the carried state) from here on
    griddep_launch_dependents();
    griddep_wait();

    const int n_half = p.n_mma / kCG;                 // B rows (= accumulator columns) produced per CTA
    // spike-tile ring stages per tile, and MMA k-blocks per stage: fc one k-block per stage; conv one stage per
    // 64-channel block, read by the 9 taps
    const int n_outer = kConv ? p.cblocks : p.kblocks;
    const int n_inner = kConv ? 9 : 1;
    // tiles of this CTA pair: group, group + n_groups, ...; with kMC every pair runs the same number of iterations
    // (the last one may be a dummy: tile >= total_tiles, computed on the last real tile's inputs, outputs suppressed)
    const int my_iters = kMC ? (p.total_tiles + n_groups - 1) / n_groups
                             : ((group < p.total_tiles) ? (p.total_tiles - group + n_groups - 1) / n_groups : 0);

    if (warp == 0) {
        // ===================================================== TMA producer (weight tiles)
        if (elect_one()) {
            uint32_t sa = 0, pa = 0;
            for (int k_it = 0; k_it < my_iters; ++k_it) {
                const int tile = min(group + k_it * n_groups, p.total_tiles - 1);      // a dummy iteration repeats the last tile
                const int ut = tile / p.m_tiles, mt = tile - ut * p.m_tiles;
                const int m0 = mt * 128 * kCG + static_cast<int>(rank) * 128;
                // k order: fc kb = 0..K/64; conv (64-channel block outer, tap inner) -- the order the MMA issuer uses
                for (int ko = 0; ko < n_outer; ++ko) {
                    for (int ki = 0; ki < n_inner; ++ki) {
                        const int kcol = kConv ? (ki * p.k_in + ko * 64) : ko * 64;
                        for (int s = 0; s < p.nsplit; ++s) {
                            mbar_wait_parked(&a_empty[sa], pa ^ 1u);
                            if (rank == 0) mbar_expect_tx(&a_full[sa], kTileBytesA * kCG);
                            uint8_t* adst = a_ring + sa * kTileBytesA;
                            if constexpr (kCG == 1) tma_load_2d(adst, &p.tmA, &a_full[sa], kcol, s * p.m_total + m0);
                            else if constexpr (!kMC) tma_load_2d_2sm(adst, &p.tmA, &a_full[sa], kcol, s * p.m_total + m0);
                            else {      // this CTA's half (64 rows, 8 KB) of the tile, to itself and its counterpart in the other pair
                                const int half = static_cast<int>(crank >> 1);
                                tma_load_2d_2sm_mcast(adst + half * (kTileBytesA / 2), &p.tmA_half, &a_full[sa], kcol,
                                                      s * p.m_total + m0 + half * 64,
                                                      static_cast<uint16_t>((1u << crank) | (1u << (crank ^ 2u))));
                            }
                            if (++sa == static_cast<uint32_t>(stages_a)) { sa = 0; pa ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ======================================================= MMA issuer
        if (rank == 0 && elect_one()) {
            uint32_t sa = 0, pa = 0, sb = 0, pb = 0, it = 0;
            const bool timed = p.role_cycles != nullptr;
            long long c_acc = 0, c_b = 0, c_peer = 0, c_a = 0, c0 = 0;
            const long long c_begin = timed ? clock64() : 0;
            for (int k_it = 0; k_it < my_iters; ++k_it, ++it) {
                const uint32_t buf = kDual ? 0u : (it & 1u);
                if (timed) c0 = clock64();
                if constexpr (kDual) {
                    mbar_wait(&acc_empty[0], (it & 1u) ^ 1u);
                    mbar_wait(&acc_empty[1], (it & 1u) ^ 1u);
                } else {
                    mbar_wait(&acc_empty[buf], ((it >> 1) & 1u) ^ 1u);
                }
                if (timed) c_acc += clock64() - c0;
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + buf * 256u;
                for (int ko = 0; ko < n_outer; ++ko) {
                    if (timed) c0 = clock64();
                    mbar_wait(&b_ready[sb], pb);
                    if (timed) { const long long c1 = clock64(); c_b += c1 - c0; c0 = c1; }
                    if constexpr (kCG == 2) mbar_wait_cluster(&b_peer[sb], pb);
                    if (timed) c_peer += clock64() - c0;
                    tcgen05_fence_after();
                    const uint32_t b_slot = smem_u32(b_ring + sb * p.slot_b);
                    for (int ki = 0; ki < n_inner; ++ki) {
                        uint64_t b_desc;
                        if constexpr (kConv) {         // tap (dy, dx): shifted window of the halo'd tile
                            const int dy = ki / 3, dx = ki - dy * 3;
                            b_desc = umma_desc_sw128_strided(b_slot + static_cast<uint32_t>(dy * p.T_box * 10 + dx) * 128u, 1280u);
                        } else {
                            b_desc = umma_desc_sw128(b_slot);
                            if (p.dbg_sbo != 0)        // swizzle experiment: shifted start, custom group stride, base offset
                                b_desc = umma_desc_sw128_strided(b_slot + static_cast<uint32_t>(p.dbg_shift) * 128u,
                                                                 static_cast<uint32_t>(p.dbg_sbo)) |
                                         (static_cast<uint64_t>(p.dbg_boff & 7) << 49);
                        }
                        for (int s = 0; s < p.nsplit; ++s) {
                            if (timed) c0 = clock64();
                            mbar_wait(&a_full[sa], pa);
                            if (timed) c_a += clock64() - c0;
                            tcgen05_fence_after();
                            const uint64_t a_desc = umma_desc_sw128(smem_u32(a_ring + sa * kTileBytesA));
#pragma unroll
                            for (int k = 0; k < 4; ++k)     // 4 x (K = 16 x 16-bit = 32 B) inside the 128-B swizzle span
                                umma_f16<kCG>(d_tmem, a_desc + 2u * k, b_desc + 2u * k, p.idesc,
                                              (ko | ki | s | k) != 0 ? 1u : 0u);
                            if constexpr (kDual) {          // the second half of the slot -> accumulator buffer 1
                                const uint64_t b_desc1 = umma_desc_sw128(b_slot + static_cast<uint32_t>(n_half) * 128u);
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma_f16<kCG>(d_tmem + 256u, a_desc + 2u * k, b_desc1 + 2u * k, p.idesc,
                                                  (ko | s | k) != 0 ? 1u : 0u);
                            }
                            if constexpr (kCG == 1) umma_commit<1>(&a_empty[sa]);
                            else umma_commit_2sm_mcast(&a_empty[sa], kMC ? static_cast<uint16_t>(0b1111) : pair_mask);
                            if (++sa == static_cast<uint32_t>(stages_a)) { sa = 0; pa ^= 1u; }
                        }
                    }
                    if constexpr (kCG == 1) umma_commit<1>(&b_empty[sb]);
                    else umma_commit_2sm_mcast(&b_empty[sb], pair_mask);
                    if (++sb == static_cast<uint32_t>(stages_b)) { sb = 0; pb ^= 1u; }
                }
                if constexpr (kCG == 1) umma_commit<1>(&acc_full[buf]);
                else umma_commit_2sm_mcast(&acc_full[buf], pair_mask);
                if constexpr (kDual) {
                    if constexpr (kCG == 1) umma_commit<1>(&acc_full[1]);
                    else umma_commit_2sm_mcast(&acc_full[1], pair_mask);
                }
            }
            if (timed) {
                unsigned long long* out = p.role_cycles + static_cast<size_t>(group) * 12;
                out[0] = static_cast<unsigned long long>(clock64() - c_begin);
                out[1] = static_cast<unsigned long long>(c_acc); out[2] = static_cast<unsigned long long>(c_b);
                out[3] = static_cast<unsigned long long>(c_peer); out[4] = static_cast<unsigned long long>(c_a);
                out[5] = it;
                out[8] = static_cast<unsigned long long>(c_begin - k_begin);      // kernel entry -> first tile
            }
        } else if (kCG == 2 && rank == 1 && lane < stages_b) {
            // relay (one lane per spike-tile ring stage): tell the leader's MMA thread that this CTA's half
            // of the stage is written.  The cluster-scope release costs about a microsecond; one lane per
            // stage keeps stages_b of them in flight, and none of them sits in a producer warp.
            const long long total_kb = static_cast<long long>(my_iters) * n_outer;
            uint32_t pb = 0;
            for (long long i = lane; i < total_kb; i += stages_b, pb ^= 1u) {
                mbar_wait_parked(&b_ready[lane], pb);
                mbar_arrive_remote(&b_peer[lane], leader_cta);
            }
        }
    } else if (warp == 3) {
        // ===================================================== TMA producer (input spike-train words)
        if (elect_one()) {
            uint32_t sw = 0, pw = 0;
            const uint32_t w_bytes = (kConv ? static_cast<uint32_t>(p.hrows) * 10u : static_cast<uint32_t>(p.Jh * (kDual ? 2 : 1))) *
                                     64u * static_cast<uint32_t>(p.in_wb);
            for (int k_it = 0; k_it < my_iters; ++k_it) {
                const int tile = min(group + k_it * n_groups, p.total_tiles - 1);
                const int ut = tile / p.m_tiles;
                const TilePos tp = decode_tile(p, ut);
                const int h0 = tp.h0 + static_cast<int>(rank) * p.sub_dh, w0 = tp.w0 + static_cast<int>(rank) * p.sub_dw;
                const int r0 = (ut * p.J + static_cast<int>(rank) * p.Jh) * (kDual ? 2 : 1);
                for (int ko = 0; ko < n_outer; ++ko) {
                    mbar_wait_parked(&w_empty[sw], pw ^ 1u);
                    mbar_expect_tx(&w_full[sw], w_bytes);
                    uint8_t* wdst = w_ring + sw * p.slot_w;
                    if constexpr (kConv)      // the CTA's pixels + one-pixel halo of 64-channel block ko
                        tma_load_4d(wdst, &p.tmW[tp.lvl], &w_full[sw], ko * 64 * p.in_wb, w0 - 1, h0 - 1, tp.n);
                    else
                        tma_load_2d(wdst, &p.tmW[0], &w_full[sw], ko * 64 * p.in_wb, r0);
                    if (++sw == static_cast<uint32_t>(stages_w)) { sw = 0; pw ^= 1u; }
                }
            }
        }
    } else if (warp == 2) {
        // ===================================================== fused LI readout (conv): out[o][px] += sum_c W[o][c] * sk[c][px]
        // The epilogue warps stage the kappa-weighted spike sums of an 8-pixel chunk ([8 px][128 ch of this CTA]) in one
        // of two shared-memory buffers; this warp (idle after the TMEM allocation) runs the 16 outputs x 8 pixels dot
        // products and adds them to the zeroed outputs (two commutative adds per output, one per CTA of the pair:
        // deterministic).  ncu r02c (bf16, where the epilogue bounds the conv): the readout was 27 % of the epilogue
        // warps' samples plus a barrier per chunk; here it overlaps the LIF recurrence of the next chunk.
        if constexpr (kConv && CW == 8) {
            if (p.fuse_readout != 0) {
                const int c0 = static_cast<int>(rank) * 128;
                const int n_out = 5 * p.A;
                for (int i = lane; i < kRoMaxOut * 128; i += 32) {
                    const int o = i >> 7, cc = i & 127;
                    float wv = 0.f;
                    if (o < p.A) wv = p.w_cls[o * p.m_total + c0 + cc];
                    else if (o < n_out) wv = p.w_bbox[(o - p.A) * p.m_total + c0 + cc];
                    ro_w[o * kRoWStride + cc] = wv;
                }
                __syncwarp();
                const int u = lane & 7, og0 = (lane >> 3) * 4;          // lane = (pixel u, outputs og0 .. og0 + 3)
                const int chunks_per_sub = p.Jh / CW;
                const uint32_t total_chunks = static_cast<uint32_t>(my_iters) * kCG * chunks_per_sub;
                const uint32_t w_a = smem_u32(ro_w) + static_cast<uint32_t>(og0 * kRoWStride) * 4u;
                uint32_t ctr = 0;
                for (int k_it = 0; k_it < my_iters; ++k_it) {
                    const bool tile_valid = group + k_it * n_groups < p.total_tiles;        // false: dummy iteration (kMC)
                    const int tile = min(group + k_it * n_groups, p.total_tiles - 1);
                    const TilePos tp = decode_tile(p, tile / p.m_tiles);
                    const LevelDesc& L = p.lv[tp.lvl];
                    const size_t hw = static_cast<size_t>(L.H) * L.W;
                    for (int ch = 0; ch < kCG * chunks_per_sub; ++ch, ++ctr) {
                        const int sub = ch / chunks_per_sub, j0 = (ch - sub * chunks_per_sub) * CW;
                        const int hh = tp.h0 + sub * p.sub_dh + (j0 >> 3), ww = tp.w0 + sub * p.sub_dw + (j0 & 7);
                        const uint32_t buf = ctr & 1u;
                        named_bar_sync(kBarRoFull + static_cast<int>(buf), kRoBarThreads);
                        const uint32_t s_a = smem_u32(ro_s) + (buf * 8u * kRoWStride + static_cast<uint32_t>(u) * kRoWStride) * 4u;
                        float a[4][2];
#pragma unroll
                        for (int j = 0; j < 4; ++j) a[j][0] = a[j][1] = 0.f;
#pragma unroll 4
                        for (int cc = 0; cc < 32; ++cc) {
                            const uint4 sv = lds_v4(s_a + 16u * cc);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint4 wv = lds_v4(w_a + static_cast<uint32_t>(j * kRoWStride) * 4u + 16u * cc);
                                a[j][0] = fmaf(__uint_as_float(wv.x), __uint_as_float(sv.x), a[j][0]);
                                a[j][1] = fmaf(__uint_as_float(wv.y), __uint_as_float(sv.y), a[j][1]);
                                a[j][0] = fmaf(__uint_as_float(wv.z), __uint_as_float(sv.z), a[j][0]);
                                a[j][1] = fmaf(__uint_as_float(wv.w), __uint_as_float(sv.w), a[j][1]);
                            }
                        }
                        // the buffer is free again (the epilogue only waits for it from its third chunk on, so the last two
                        // arrivals have no partner and are not made)
                        if (ctr + 2u < total_chunks) named_bar_arrive(kBarRoFree + static_cast<int>(buf), kRoBarThreads);
                        if (tile_valid && hh < L.H && ww + u < L.W) {
                            const size_t pix = static_cast<size_t>(hh) * L.W + ww + u;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int og = og0 + j;
                                const float r = a[j][0] + a[j][1];
                                if (og < p.A) atomicAdd(&L.logits[(static_cast<size_t>(tp.n) * p.A + og) * hw + pix], r);
                                else if (og < n_out) atomicAdd(&L.bbox[(static_cast<size_t>(tp.n) * 4 * p.A + (og - p.A)) * hw + pix], r);
                            }
                        }
                    }
                }
            }
        }
    } else if (warp >= 8 && warp < 8 + kProducerWarps) {
        // ============================== spike-tile producers (n_pg groups): words -> swizzled {0,1} tile
        // The k-blocks of this CTA's tile sequence are numbered i = 0, 1, 2, ...; producer group g expands
        // the k-blocks i = g (mod n_pg): word-ring stage i % stages_w -> spike-tile ring stage i % stages_b,
        // so n_pg stages are in production at any time.  A thread owns up to kMaxPairs (unit j, 16-byte
        // chunk q) pairs of the CTA's half tile; per k-block it reads a pair's 8 input words and writes T_box
        // 16-byte chunks: row r = t * Jh + j, chunk q ^ (r & 7) -- the 128-byte swizzle TMA would have
        // produced for a K-major 16-bit tile.
        const int n_pg = p.n_pg;
        const int tpg = kProducerWarps * 32 / n_pg;    // threads per group
        const int ptid = static_cast<int>(threadIdx.x) - 256;
        const int grp = ptid / tpg;
        const int pt = ptid - grp * tpg;
        // conv: every pixel of the halo'd region; fc: the CTA's units
        const int n_pairs = (kConv ? p.hrows * 10 : p.Jh) * 8;
        const int wb = p.in_wb;
        const uint32_t tmask = (p.T_live >= 32) ? 0xFFFFFFFFu : ((1u << p.T_live) - 1u);
        const uint32_t pmask = tmask | (tmask << 16);
        const uint32_t one = p.spike_one;
        const bool packed = p.T_box <= 16;             // both neurons of a 32-bit output fit one register
        const long long total_kb = static_cast<long long>(my_iters) * n_outer;
        const uint32_t b_base = smem_u32(b_ring), w_base = smem_u32(w_ring);
        // row of (unit j, step t): fc t * Jh + j; conv halo pixel (hh, ww): (hh * T_box + t) * 10 + ww
        const uint32_t row_step = (kConv ? 10u : static_cast<uint32_t>(p.Jh)) * 128u;

        // ---- fc fast path: the spike tile of a k-block is produced by ALL producer threads, one 16-byte chunk
        // (row t * Jh + j, chunk q) per item, at most 8 items per thread (4 with cta_group 2).  The item geometry does not depend on the
        // k-block, so each thread keeps its items' word offset, swizzled store offset and step in registers; a stage
        // is then ~14 instructions per item with every item independent (the per-pair loop below walks the T_box
        // steps of a pair serially: ~5x the latency per stage, which the fc pipeline -- one stage per k-block,
        // handed across the CTA pair -- cannot hide).
        if (!kConv && packed && p.dbg_sbo == 0 && n_pg == 1) {
            constexpr int kMaxItems = 8;                     // 256 rows x 8 chunks / 256 threads (cta_group 1, or cta_group 2 dual); 4 with cta_group 2
            const int n_items = n_pairs * p.T_box * (kDual ? 2 : 1);   // <= 256 rows * 8 chunks
            // expansion of one 32-bit output (two neurons' bits of step t -> two 16-bit ones):
            //   ((P >> pre) & mask) * mul  with  mask = 0x00010001 << low, mul = one >> low, low = min(shift, ctz(one)),
            //   pre = shift - low: the bit is multiplied where it stands, so for shift <= ctz(one) (10 for fp16 1.0,
            //   7 for bf16 1.0) there is no shift instruction at all
            uint32_t it_src[kMaxItems], it_dst[kMaxItems], it_pre[kMaxItems], it_mask[kMaxItems], it_mul[kMaxItems];
            const uint32_t one_ctz = static_cast<uint32_t>(__ffs(static_cast<int>(one)) - 1);
            int n_my = 0;
            // items of one thread are kProducerWarps * 32 apart: when that is a multiple of n_pairs (fc6: 64 pairs) they are
            // the T_box steps of ONE pair and read the same 8 words, so the words are loaded once per run of equal sources
            // (bit i of `reload`: item i's source differs from item i-1's) -- 2 instead of 6 LDS.128 per thread and stage
            // in the dual fc6, where the producers' 32 KB of loads + 32 KB of stores per stage were most of the shared-memory
            // wavefronts of a 768-cycle bf16 stage (r02w/r02x A/B on one box: bf16 fc6 0.946 -> 0.926 ms, fp16x2 unchanged)
            uint32_t reload = 1u;
#pragma unroll
            for (int i = 0; i < kMaxItems; ++i) {
                const int item = ptid + i * (kProducerWarps * 32);
                it_src[i] = it_dst[i] = 0u; it_pre[i] = it_mask[i] = it_mul[i] = 0u;
                if (item < n_items) {
                    const int tb = item / n_pairs, pr = item - tb * n_pairs;
                    const int b = kDual ? tb / p.T_box : 0, t = tb - b * p.T_box;     // b: accumulator buffer (dual)
                    const uint32_t j = pr >> 3, q = pr & 7, r = static_cast<uint32_t>(b * n_half + t * p.Jh) + j;
                    it_src[i] = static_cast<uint32_t>(b * n_pairs + pr) * 8u * wb;
                    it_dst[i] = r * 128u + ((q ^ (r & 7u)) << 4);      // slots are 1024-B aligned
                    if (t < p.T_live) {                                            // else: padding step, zero row (mask 0)
                        const uint32_t shift = (wb == 4 ? 0u : static_cast<uint32_t>(p.in_bit0)) + static_cast<uint32_t>(t);
                        const uint32_t low = shift < one_ctz ? shift : one_ctz;
                        it_pre[i] = shift - low; it_mask[i] = 0x00010001u << low; it_mul[i] = one >> low;
                    }
                    n_my = i + 1;
                    if (i > 0 && (it_src[i] != it_src[i - 1] || !p.prod_word_reuse)) reload |= 1u << i;
                }
            }
            // The stage loop, specialised on the input word size (no per-item branch on it) and with a trip count that is
            // the same for every thread (n_max = the items of the busiest thread; only its last item is predicated per
            // thread), branch-free inside: ncu r02n counted 31 warp instructions per item in the bf16 fc6, where the
            // producers bound the kernel (60 % of all its instructions), against ~17 here.
            const int n_max = (n_items + kProducerWarps * 32 - 1) / (kProducerWarps * 32);
            auto run = [&](auto wb_tag) {
                constexpr int WB = decltype(wb_tag)::value;
                uint32_t sb = 0u, pb = 0u, sw = 0u, pw = 0u;
                for (long long i_kb = 0; i_kb < total_kb; ++i_kb) {
                    mbar_wait_parked(&w_full[sw], pw);
                    mbar_wait_parked(&b_empty[sb], pb ^ 1u);   // (a sleep between polls, as in the conv, costs the bf16 fc6 3 % and gives fp16x2 nothing: r02x)
                    const uint32_t wslot = w_base + sw * p.slot_w;
                    const uint32_t slot = b_base + sb * p.slot_b;
                    uint32_t P[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                    for (int i = 0; i < kMaxItems; ++i) {
                        if (i < n_max) {                       // uniform over the block
                            const uint32_t src = wslot + it_src[i];
                            if (!((reload >> i) & 1u)) {
                                // same words as the previous item: P is still in registers
                            } else if constexpr (WB == 1) {
                                const uint2 v = lds_v2(src);
                                P[0] = __byte_perm(v.x, 0u, 0x4140); P[1] = __byte_perm(v.x, 0u, 0x4342);
                                P[2] = __byte_perm(v.y, 0u, 0x4140); P[3] = __byte_perm(v.y, 0u, 0x4342);
                            } else if constexpr (WB == 2) {
                                const uint4 v = lds_v4(src);
                                P[0] = v.x; P[1] = v.y; P[2] = v.z; P[3] = v.w;
                            } else {
                                const uint4 v = lds_v4(src), c = lds_v4(src + 16u);
                                P[0] = ((v.x >> p.in_bit0) & tmask) | (((v.y >> p.in_bit0) & tmask) << 16);
                                P[1] = ((v.z >> p.in_bit0) & tmask) | (((v.w >> p.in_bit0) & tmask) << 16);
                                P[2] = ((c.x >> p.in_bit0) & tmask) | (((c.y >> p.in_bit0) & tmask) << 16);
                                P[3] = ((c.z >> p.in_bit0) & tmask) | (((c.w >> p.in_bit0) & tmask) << 16);
                            }
                            const uint32_t pre = it_pre[i], m = it_mask[i], mu = it_mul[i];
                            uint4 o;
                            o.x = ((P[0] >> pre) & m) * mu; o.y = ((P[1] >> pre) & m) * mu;
                            o.z = ((P[2] >> pre) & m) * mu; o.w = ((P[3] >> pre) & m) * mu;
                            if (i < n_my) sts_v4(slot + it_dst[i], o);
                        }
                    }
                    fence_proxy_async_smem();          // generic-proxy smem writes -> visible to the tensor core
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(&b_ready[sb]); mbar_arrive(&w_empty[sw]); }
                    if (++sb == static_cast<uint32_t>(stages_b)) { sb = 0u; pb ^= 1u; }
                    if (++sw == static_cast<uint32_t>(stages_w)) { sw = 0u; pw ^= 1u; }
                }
            };
            if (wb == 2) run(std::integral_constant<int, 2>{});
            else if (wb == 1) run(std::integral_constant<int, 1>{});
            else run(std::integral_constant<int, 4>{});
        } else {
        // ---- general path (conv halo tiles; T_box > 16)
        // group g starts on stage g of both rings (n_pg <= stages, so its first phase parity is 0)
        uint32_t sb = static_cast<uint32_t>(grp), pb = 0u, sw = static_cast<uint32_t>(grp), pw = 0u;
        for (long long i_kb = grp; i_kb < total_kb; i_kb += n_pg) {
            mbar_wait_parked(&w_full[sw], pw);
            // the long wait of a conv producer: a halo'd spike tile serves 9 taps x the pieces (microseconds of MMAs)
            if (kConv && p.wait_backoff_ns > 0) mbar_wait_backoff(&b_empty[sb], pb ^ 1u, static_cast<uint32_t>(p.wait_backoff_ns));
            else mbar_wait_parked(&b_empty[sb], pb ^ 1u);
            const uint32_t wslot = w_base + sw * p.slot_w;
            const uint32_t slot = b_base + sb * p.slot_b;
#pragma unroll
            for (int i = 0; i < kMaxPairs; ++i) {
                const int pr = pt + i * tpg;
                if (pr >= n_pairs) break;
                const uint32_t j = pr >> 3, q = pr & 7;
                const uint32_t src = wslot + static_cast<uint32_t>(pr) * 8u * wb;
                uint32_t r0 = j;
                if constexpr (kConv) { const uint32_t hh = j / 10u; r0 = hh * static_cast<uint32_t>(p.T_box) * 10u + (j - hh * 10u); }
                uint32_t addr = slot + r0 * 128u;      // slot is 1024-B aligned: (addr >> 7) & 7 == row & 7
                if (packed) {
                    uint32_t P[4];
                    if (wb == 1) {
                        const uint2 v = lds_v2(src);
                        P[0] = __byte_perm(v.x, 0u, 0x4140); P[1] = __byte_perm(v.x, 0u, 0x4342);
                        P[2] = __byte_perm(v.y, 0u, 0x4140); P[3] = __byte_perm(v.y, 0u, 0x4342);
                    } else if (wb == 2) {
                        const uint4 v = lds_v4(src);
                        P[0] = v.x; P[1] = v.y; P[2] = v.z; P[3] = v.w;
                    } else {
                        const uint4 a = lds_v4(src), c = lds_v4(src + 16u);
                        P[0] = ((a.x >> p.in_bit0) & tmask) | (((a.y >> p.in_bit0) & tmask) << 16);
                        P[1] = ((a.z >> p.in_bit0) & tmask) | (((a.w >> p.in_bit0) & tmask) << 16);
                        P[2] = ((c.x >> p.in_bit0) & tmask) | (((c.y >> p.in_bit0) & tmask) << 16);
                        P[3] = ((c.z >> p.in_bit0) & tmask) | (((c.w >> p.in_bit0) & tmask) << 16);
                    }
                    if (wb != 4) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) P[e] = (P[e] >> p.in_bit0) & pmask;
                    }
#pragma unroll 4
                    for (int t = 0; t < p.T_box; ++t, addr += row_step) {
                        uint4 o;
                        o.x = ((P[0] >> t) & 0x00010001u) * one; o.y = ((P[1] >> t) & 0x00010001u) * one;
                        o.z = ((P[2] >> t) & 0x00010001u) * one; o.w = ((P[3] >> t) & 0x00010001u) * one;
                        if (!kConv && p.dbg_sbo != 0) {  // swizzle experiment: rows in groups of 8 at stride dbg_sbo, shifted
                            const uint32_t r = r0 + static_cast<uint32_t>(t * p.Jh);
                            const uint32_t ra = slot + static_cast<uint32_t>(p.dbg_shift) * 128u + (r >> 3) * p.dbg_sbo + (r & 7u) * 128u;
                            sts_v4(ra + ((q ^ ((ra >> 7) & 7u)) << 4), o);
                        } else {
                            sts_v4(addr + ((q ^ ((addr >> 7) & 7u)) << 4), o);     // swizzle by the absolute row address
                        }
                    }
                } else {                           // T_box > 16: 32-bit words, one neuron per register
                    const uint4 a = lds_v4(src), c = lds_v4(src + 16u);
                    uint32_t wv[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
                    for (int e = 0; e < 8; ++e) wv[e] = (wv[e] >> p.in_bit0) & tmask;
                    for (int t = 0; t < p.T_box; ++t, addr += row_step) {
                        uint4 o;
                        o.x = (((wv[0] >> t) & 1u) | (((wv[1] >> t) & 1u) << 16)) * one;
                        o.y = (((wv[2] >> t) & 1u) | (((wv[3] >> t) & 1u) << 16)) * one;
                        o.z = (((wv[4] >> t) & 1u) | (((wv[5] >> t) & 1u) << 16)) * one;
                        o.w = (((wv[6] >> t) & 1u) | (((wv[7] >> t) & 1u) << 16)) * one;
                        sts_v4(addr + ((q ^ ((addr >> 7) & 7u)) << 4), o);
                    }
                }
            }
            fence_proxy_async_smem();          // generic-proxy smem writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) { mbar_arrive(&b_ready[sb]); mbar_arrive(&w_empty[sw]); }
            // this group's next k-block is n_pg ring stages further
            sb += n_pg; if (sb >= static_cast<uint32_t>(stages_b)) { sb -= stages_b; pb ^= 1u; }
            sw += n_pg; if (sw >= static_cast<uint32_t>(stages_w)) { sw -= stages_w; pw ^= 1u; }
        }
        }
    } else if (warp >= 4) {
        // ============================================ LIF epilogue (2 groups x 4 warps)
        // Group eg takes every other unit chunk of a tile; a warp reads the TMEM lane quadrant warp % 4.
        const int eg = warp >= 16 ? 1 : 0;
        const int q = warp & 3;                        // TMEM lane quadrant of this warp
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const int te = q * 32 + lane;                  // 0..127: channel of this thread inside the CTA's 128
        const bool fused = kConv && CW == 8 && p.fuse_readout != 0;
        float* ro_sg = ro_s + eg * (2 * 8 * kRoWStride);
        uint32_t chunk_ctr = 0;
        uint32_t it = 0;
        const bool e_timed = p.role_cycles != nullptr && rank == 0 && warp == 4 && lane == 0;
        uint32_t e_wait = 0, e_t0 = 0;                 // 32-bit cycle counters: the role is short of registers
        const uint32_t e_begin = e_timed ? static_cast<uint32_t>(clock()) : 0u;
        for (int k_it = 0; k_it < my_iters; ++k_it, ++it) {
            const bool tile_valid = group + k_it * n_groups < p.total_tiles;                // false: dummy iteration (kMC)
            const int tile = min(group + k_it * n_groups, p.total_tiles - 1);
            const int ut = tile / p.m_tiles, mt = tile - ut * p.m_tiles;
            const int c = mt * 128 * kCG + static_cast<int>(rank) * 128 + te;
            int H = 1, W = 1, n = 0, h0 = 0, w0 = 0, lvl = 0;
            uint8_t* trains = reinterpret_cast<uint8_t*>(p.trains);
            unsigned int tile_spikes = 0;
            if constexpr (kConv) {
                const TilePos tp = decode_tile(p, ut);
                lvl = tp.lvl; n = tp.n; h0 = tp.h0; w0 = tp.w0;
                H = p.lv[lvl].H; W = p.lv[lvl].W;
                trains = reinterpret_cast<uint8_t*>(p.lv[lvl].trains);
            }
            const float wscale = __ldg(&p.w_scale[c]);
            for (int bi = 0; bi < (kDual ? 2 : 1); ++bi) {      // dual: both buffers belong to this tile
            const uint32_t buf = kDual ? static_cast<uint32_t>(bi) : (it & 1u);
            if (e_timed) e_t0 = static_cast<uint32_t>(clock());
            mbar_wait(&acc_full[buf], kDual ? (it & 1u) : ((it >> 1) & 1u));   // (a sleep between these polls changes nothing, sustained or burst: r02aa)
            if (e_timed) e_wait += static_cast<uint32_t>(clock()) - e_t0;
            tcgen05_fence_after();
            const uint32_t acc = tmem_base + lane_addr + buf * 256u;
            // first unit of CTA half `sub` of this accumulator (fc)
            const int unit0 = (ut * p.J) * (kDual ? 2 : 1) + bi * p.Jh;
            const int sub_units = p.Jh * (kDual ? 2 : 1);

            const int chunks_per_sub = p.Jh / CW;
            for (int ch = eg; ch < kCG * chunks_per_sub; ch += kEpiGroups) {
                const int sub = ch / chunks_per_sub;
                const int j0 = (ch - sub * chunks_per_sub) * CW;
                {
                    float v[CW], ii[CW], sk[CW];
                    uint32_t tr[CW];
#pragma unroll
                    for (int u = 0; u < CW; ++u) { v[u] = 0.f; ii[u] = 0.f; tr[u] = 0u; sk[u] = 0.f; }
                    // position of the chunk's units
                    int hh = 0, ww = 0;
                    bool row_ok;
                    size_t r0;
                    int lim;                                   // units u < lim are inside the image / row range
                    if constexpr (kConv) {                     // a chunk lies inside one tile row (TWh == 8, CW <= 8)
                        hh = h0 + sub * p.sub_dh + (j0 >> 3);
                        ww = w0 + sub * p.sub_dw + (j0 & 7);
                        row_ok = tile_valid && hh < H;
                        lim = W - ww;
                        r0 = (static_cast<size_t>(n) * H + hh) * W + ww;
                    } else {
                        const int rr = unit0 + sub * sub_units + j0;
                        row_ok = tile_valid;
                        lim = p.rows - rr;
                        r0 = static_cast<size_t>(rr);
                    }
                    float4* state = nullptr;
                    if (p.state_load | p.state_store) {
                        if constexpr (kConv) state = p.lv[lvl].state; else state = p.state;
                        state += r0 * p.m_total + c;
                        if (p.state_load && row_ok) {
#pragma unroll
                            for (int u = 0; u < CW; ++u) {
                                if (u >= lim) break;
                                const float4 sv = state[static_cast<size_t>(u) * p.m_total];
                                v[u] = sv.x; ii[u] = sv.y; sk[u] = sv.z; tr[u] = __float_as_uint(sv.w);
                            }
                        }
                    }
                    // ---- steps that receive an input current (the accumulator columns of this unit chunk)
                    // accumulator column of (unit j, step t): fc t * Jh + j; conv (tile row, t, tile column)
                    uint32_t col = acc + static_cast<uint32_t>(sub * n_half + (kConv ? (j0 >> 3) * p.T_box * 8 + (j0 & 7) : j0));
                    const uint32_t col_step = kConv ? 8u : static_cast<uint32_t>(p.Jh);
                    // fc: the loads of G steps are issued before one wait (the fc epilogue of a dual tile is not hidden
                    // behind the next tile's main loop)
                    constexpr int G = kConv ? (CW == 8 ? 2 : 4) : 16 / CW;
                    for (int tl0 = 0; tl0 < p.T_live; tl0 += G) {
                        float cu[G][CW];
#pragma unroll
                        for (int g = 0; g < G; ++g)
                            if (tl0 + g < p.T_live) tmem_ld<CW>(col + g * col_step, reinterpret_cast<uint32_t*>(cu[g]));
                        tmem_ld_wait();
                        col += G * col_step;
#pragma unroll
                        for (int g = 0; g < G; ++g) {
                            const int tl = tl0 + g;
                            if (tl >= p.T_live) break;
                            const int t = p.t0 + tl;
                            const float kap = p.kappa[t];
                            const uint32_t bit = 1u << t;
#pragma unroll
                            for (int u = 0; u < CW; ++u)
                                if (lif_update_scaled(v[u], ii[u], cu[g][u], wscale)) { tr[u] |= bit; sk[u] = __fadd_rn(sk[u], kap); }
                            if constexpr (!kConv) {
                                if (p.dump != nullptr) {
#pragma unroll
                                    for (int u = 0; u < CW; ++u) {
                                        const int r = unit0 + sub * sub_units + j0 + u;
                                        if (r < p.rows)
                                            p.dump[(static_cast<size_t>(tl) * p.dump_rows + r) * p.m_total + c] = __fmul_rn(cu[g][u], wscale);
                                    }
                                }
                            }
                        }
                    }
                    if (p.state_store) {           // not the last pass over the time axis: carry the state, emit nothing
                        if (row_ok) {
#pragma unroll
                            for (int u = 0; u < CW; ++u) {
                                if (u >= lim) break;
                                state[static_cast<size_t>(u) * p.m_total] = make_float4(v[u], ii[u], sk[u], __uint_as_float(tr[u]));
                            }
                        }
                        continue;
                    }
                    // ---- remaining steps: the synapse only drains (no new input reaches an output later)
                    for (int t = p.t0 + p.T_live; t < p.T_total; ++t) {
                        const float kap = p.kappa[t];
                        const uint32_t bit = 1u << t;
#pragma unroll
                        for (int u = 0; u < CW; ++u)
                            if (lif_update(v[u], ii[u], 0.0f)) { tr[u] |= bit; sk[u] = __fadd_rn(sk[u], kap); }
                    }
                    // ---- emit the spike-train words of the chunk
                    if (row_ok) {
#pragma unroll
                        for (int u = 0; u < CW; ++u)
                            if (u < lim) tile_spikes += __popc(tr[u]);
                        if (trains != nullptr) {
                            uint8_t* dst = trains + (r0 * p.m_total + c) * p.train_bytes;
                            const size_t step = static_cast<size_t>(p.m_total) * p.train_bytes;
#pragma unroll
                            for (int u = 0; u < CW; ++u, dst += step) {
                                if (u >= lim) break;
                                if (p.train_bytes == 1) *dst = static_cast<uint8_t>(tr[u]);
                                else if (p.train_bytes == 2) *reinterpret_cast<uint16_t*>(dst) = static_cast<uint16_t>(tr[u]);
                                else *reinterpret_cast<uint32_t*>(dst) = tr[u];
                            }
                        }
                    }
                    // ---- fused LI readout: hand the chunk's kappa-weighted spike sums to the readout warp (warp 2)
                    if constexpr (kConv) {
                        if (fused) {
                            const uint32_t rb = chunk_ctr & 1u;
                            // the readout warp must have consumed what this buffer held two chunks ago
                            if (chunk_ctr >= 2u) named_bar_sync(kBarRoFree + static_cast<int>(rb), kRoBarThreads);
                            ++chunk_ctr;
                            // explicit shared-space stores (the generic-pointer form compiled to ST.E): S[px u][channel te]
                            const uint32_t s_dst = smem_u32(ro_sg) + (rb * 8u * kRoWStride + static_cast<uint32_t>(te)) * 4u;
#pragma unroll
                            for (int u = 0; u < CW; ++u)
                                asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_dst + static_cast<uint32_t>(u * kRoWStride) * 4u), "f"(sk[u]) : "memory");
                            named_bar_arrive(kBarRoFull + static_cast<int>(rb), kRoBarThreads);
                        }
                    }
                }
            }
            if constexpr (kConv) {
                if (p.lv[lvl].counts != nullptr) {
                    for (int o = 16; o > 0; o >>= 1) tile_spikes += __shfl_xor_sync(0xffffffffu, tile_spikes, o);
                    if (lane == 0 && tile_spikes) atomicAdd(&p.lv[lvl].counts[n], static_cast<unsigned long long>(tile_spikes));
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                // the TMEM reads of this warp are complete (tcgen05.wait::ld) and fenced (fence::before_thread_sync); the
                // arrive only carries the signal to the MMA thread, which fences after its wait -- CTA-scope release as in
                // CUTLASS's ClusterBarrier::arrive(cta_id) (the cluster-scope release compiled to an ERRBAR that was 6 %
                // of the epilogue warps' samples, ncu r02c)
                if constexpr (kCG == 1) mbar_arrive(&acc_empty[buf]);
                else mbar_arrive_remote(&acc_empty[buf], leader_cta);
            }
            }   // bi
        }
        if (e_timed) {      // [6] the epilogue role of warp 4, [7] of which waiting for a full accumulator
            p.role_cycles[static_cast<size_t>(group) * 12 + 6] = static_cast<uint32_t>(clock()) - e_begin;
            p.role_cycles[static_cast<size_t>(group) * 12 + 7] = e_wait;
        }
    }

    tcgen05_fence_before();
    if constexpr (kCG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 2) tmem_dealloc<kCG>(tmem_base, 512);
    if ((p.role_cycles != nullptr || p.clock_probe != nullptr) && rank == 0 && threadIdx.x == 64) {
        unsigned long long k_end_ns;                       // kernel entry -> exit of the leader CTA
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(k_end_ns));
        const unsigned long long cyc = static_cast<unsigned long long>(clock64() - k_begin);
        if (p.role_cycles != nullptr) {                    // [9] in SM cycles, [10] in nanoseconds
            p.role_cycles[static_cast<size_t>(group) * 12 + 9] = cyc;
            p.role_cycles[static_cast<size_t>(group) * 12 + 10] = k_end_ns - k_begin_ns;
        }
        if (p.clock_probe != nullptr) {
            atomicAdd(&p.clock_probe[static_cast<size_t>(group) * 2], cyc);
            atomicAdd(&p.clock_probe[static_cast<size_t>(group) * 2 + 1], k_end_ns - k_begin_ns);
        }
    }
}

}  // namespace snn
