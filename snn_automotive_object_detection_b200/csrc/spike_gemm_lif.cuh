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
//  * A operand  = weights [M rows = output channels][K] bf16, K-major, TMA 2-D tiles 128 x 64.
//    "fp32-exact" mode keeps 2 or 3 bf16 pieces of every weight (hi/mid/lo); because the other
//    operand is exactly {0,1} every product is exact and the pieces are simply extra k-steps.
//    fp16 modes keep 1 or 2 fp16 pieces of the power-of-two row-scaled weight (11 bits each) and
//    the epilogue multiplies the accumulator row by the inverse scale (exact).
//  * B operand  = input spikes, rows ordered (t, unit) so that ALL timesteps of a unit sit in
//    the same accumulator tile (time folded into the MMA N dimension, N = T_box * J <= 256):
//      fc   : 3-D tensor map  [T][R][K]          box (64, Jh, T_box)
//      conv : 5-D tensor map  [T][N][H][W][C]    box (64, TWh, THh, 1, T_box), one box per
//             (tap, 64-channel block); the 3x3 halo and image border are TMA out-of-bounds zero fill.
//  * accumulators: 2 x 256 TMEM columns (double buffered) -> the MMA of tile i+1 overlaps the
//    LIF epilogue of tile i.  The LIF state (v, i) never leaves registers; nothing of size
//    T x state is ever written to HBM.  Output per neuron: one time-packed spike-train word
//    (bit t = spike at step t; popc = spike count) and, optionally, bf16 {0,1} planes that feed
//    the next layer's contraction.
//  * kCG = 2 pairs two SMs (cta_group::2, UMMA M = 256): each CTA owns 128 output channels and
//    loads half of the unit tile, halving B traffic per SM.
#pragma once
#include <cuda_bf16.h>

#include "ptx.cuh"

namespace snn {

constexpr int kMaxLevels = 8;
constexpr int kStagesA = 6;                 // 16 KB each
constexpr int kStagesB = 3;                 // up to 32 KB each
constexpr int kTileBytesA = 128 * 128;      // 128 rows x 64 bf16
constexpr int kSlotBytesB = 256 * 128;      // up to 256 rows x 64 bf16
constexpr int kGemmThreads = 256;
constexpr int kRoMaxOut = 16;               // fused readout: objectness + box deltas per pixel (5 * A <= 16)
constexpr int kRoWStride = 132;             // floats per readout-weight row in smem (128 + pad, 16-B aligned)
constexpr int kRoSStride = 20;              // floats per channel row of the kappa-weighted spike sums (16 + pad)
constexpr int kRoSmemBytes = kRoMaxOut * kRoWStride * 4 + 2 * 128 * kRoSStride * 4;
constexpr size_t kGemmSmemBytes =
    1024 /*align slack*/ + kStagesA * kTileBytesA + kStagesB * kSlotBytesB + 256 + kRoSmemBytes;

struct LevelDesc {
    int H, W, tiles_w, tiles_h;
    int tile_begin;           // first unit-tile index of this level
    int pad_;
    void* trains;             // [N][H][W][m_total] spike-train words (nullable when the readout is fused)
    float* logits;            // fused readout: [N][A][H][W]   (zero-initialised; 2 CTAs add their channel halves)
    float* bbox;              // fused readout: [N][4A][H][W]
    unsigned long long* counts;   // nullable: [N] spikes per image of this level
};

struct GemmLifParams {
    CUtensorMap tmA;
    CUtensorMap tmB[kMaxLevels];
    LevelDesc lv[kMaxLevels];
    int n_levels, conv, n_images;
    int m_total, m_tiles, nsplit;
    int kblocks, cblocks;
    int T_total, t0, T_live, T_box;
    int J, Jh, TWh, THh, TW, TH, sub_dw, sub_dh;
    int rows;                 // fc: number of units (RoIs)
    int total_tiles, unit_tiles;
    int train_bytes;          // 1, 2 or 4
    int n_mma;                // T_box * J
    uint32_t idesc;
    void* trains;             // fc: [rows][m_total]
    uint16_t* spikes_out;     // optional: [spk_t_hi - spk_t_lo][rows][m_total] 16-bit {0,1} (pattern spike_one)
    int spk_t_lo, spk_t_hi;
    uint32_t spike_one;       // 1.0 as bf16 (0x3F80) or fp16 (0x3C00)
    const float* w_scale;     // [m_total] power of two each accumulator row is multiplied with (1 for bf16 pieces)
    float* dump;              // debug (fc only): raw currents [T_live][rows][m_total]
    // fused leaky-integrator readout (conv, cta_group 2, m_total == 256): mem_{T-1} = W . sum_t kappa_{T-1-t} spk_t
    int fuse_readout, A;
    const float* w_cls;       // [A][m_total]
    const float* w_bbox;      // [4A][m_total]
    float kappa[32];          // kappa[t] = 0.9^{T-t} - 0.8^{T-t}: weight of a spike at step t in the last membrane
};

// One LIF step of Norse's lif_feed_forward_step, op for op (no FMA contraction):
//   v_dec = v + 0.1f*((0 - v) + i);  i_dec = i + (-0.2f)*i;  z = (v_dec - 0.1f > 0);
//   v' = (1-z)*v_dec + z*0;  i' = i_dec + cur
__device__ __forceinline__ uint32_t lif_update(float& v, float& i, float cur) {
    const float dv = __fmul_rn(0.1f, __fsub_rn(i, v));
    const float v_dec = __fadd_rn(v, dv);
    const float i_dec = __fadd_rn(i, __fmul_rn(-0.2f, i));
    const bool z = __fsub_rn(v_dec, 0.1f) > 0.0f;
    v = z ? 0.0f : v_dec;
    i = __fadd_rn(i_dec, cur);
    return z ? 1u : 0u;
}

template <int kCG, int CW>
__global__ void __launch_bounds__(kGemmThreads, 1) spike_gemm_lif_kernel(const __grid_constant__ GemmLifParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_ring = smem;
    uint8_t* b_ring = smem + kStagesA * kTileBytesA;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + kStagesB * kSlotBytesB);
    uint64_t* a_full = bars;                         // [kStagesA]
    uint64_t* a_empty = a_full + kStagesA;           // [kStagesA]
    uint64_t* b_full = a_empty + kStagesA;           // [kStagesB]
    uint64_t* b_empty = b_full + kStagesB;           // [kStagesB]
    uint64_t* acc_full = b_empty + kStagesB;         // [2]
    uint64_t* acc_empty = acc_full + 2;              // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* ro_w = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);     // [kRoMaxOut][kRoWStride]
    float* ro_s = ro_w + kRoMaxOut * kRoWStride;                                          // [2][128][kRoSStride]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = (kCG == 2) ? cluster_ctarank() : 0u;
    const int n_groups = gridDim.x / kCG;
    const int group = blockIdx.x / kCG;

    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&p.tmA);
        for (int l = 0; l < p.n_levels; ++l) tma_prefetch_desc(&p.tmB[l]);
    }
    if (warp == 1 && elect_one()) {
        for (int s = 0; s < kStagesA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < kStagesB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4 * kCG); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<kCG>(tmem_slot, 512);
    tcgen05_fence_before();
    if constexpr (kCG == 2) cluster_sync_all(); else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_half = p.n_mma / kCG;                 // B rows (= accumulator columns) loaded per CTA
    const uint32_t b_bytes = static_cast<uint32_t>(p.n_mma) * 128u;   // per k-block, both CTAs together

    if (warp == 0) {
        // ===================================================== TMA producer
        if (elect_one()) {
            uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
            for (int tile = group; tile < p.total_tiles; tile += n_groups) {
                const int ut = tile / p.m_tiles, mt = tile - ut * p.m_tiles;
                const int m0 = mt * 128 * kCG + static_cast<int>(rank) * 128;
                int lvl = 0, n = 0, h0 = 0, w0 = 0;
                if (p.conv) {
                    while (lvl + 1 < p.n_levels && ut >= p.lv[lvl + 1].tile_begin) ++lvl;
                    const LevelDesc& L = p.lv[lvl];
                    int local = ut - L.tile_begin;
                    const int per_img = L.tiles_w * L.tiles_h;
                    n = local / per_img; local -= n * per_img;
                    const int ty = local / L.tiles_w;
                    h0 = ty * p.TH + static_cast<int>(rank) * p.sub_dh;
                    w0 = (local - ty * L.tiles_w) * p.TW + static_cast<int>(rank) * p.sub_dw;
                }
                const int r0 = ut * p.J + static_cast<int>(rank) * p.Jh;
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    mbar_wait(&b_empty[sb], pb ^ 1u);
                    if (rank == 0) mbar_expect_tx(&b_full[sb], b_bytes);
                    uint8_t* bdst = b_ring + sb * kSlotBytesB;
                    if (p.conv) {
                        const int tap = kb / p.cblocks, cb = kb - tap * p.cblocks;
                        const int dy = tap / 3, dx = tap - dy * 3;
                        if constexpr (kCG == 1)
                            tma_load_5d(bdst, &p.tmB[lvl], &b_full[sb], cb * 64, w0 + dx - 1, h0 + dy - 1, n, 0);
                        else
                            tma_load_5d_2sm(bdst, &p.tmB[lvl], &b_full[sb], cb * 64, w0 + dx - 1, h0 + dy - 1, n, 0);
                    } else {
                        if constexpr (kCG == 1) tma_load_3d(bdst, &p.tmB[0], &b_full[sb], kb * 64, r0, 0);
                        else tma_load_3d_2sm(bdst, &p.tmB[0], &b_full[sb], kb * 64, r0, 0);
                    }
                    if (++sb == kStagesB) { sb = 0; pb ^= 1u; }
                    for (int s = 0; s < p.nsplit; ++s) {
                        mbar_wait(&a_empty[sa], pa ^ 1u);
                        if (rank == 0) mbar_expect_tx(&a_full[sa], kTileBytesA * kCG);
                        uint8_t* adst = a_ring + sa * kTileBytesA;
                        if constexpr (kCG == 1) tma_load_2d(adst, &p.tmA, &a_full[sa], kb * 64, s * p.m_total + m0);
                        else tma_load_2d_2sm(adst, &p.tmA, &a_full[sa], kb * 64, s * p.m_total + m0);
                        if (++sa == kStagesA) { sa = 0; pa ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ======================================================= MMA issuer
        if (rank == 0 && elect_one()) {
            uint32_t sa = 0, pa = 0, sb = 0, pb = 0, it = 0;
            for (int tile = group; tile < p.total_tiles; tile += n_groups, ++it) {
                const uint32_t buf = it & 1u;
                mbar_wait(&acc_empty[buf], ((it >> 1) & 1u) ^ 1u);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + buf * 256u;
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    mbar_wait(&b_full[sb], pb);
                    tcgen05_fence_after();
                    const uint64_t b_desc = umma_desc_sw128(smem_u32(b_ring + sb * kSlotBytesB));
                    for (int s = 0; s < p.nsplit; ++s) {
                        mbar_wait(&a_full[sa], pa);
                        tcgen05_fence_after();
                        const uint64_t a_desc = umma_desc_sw128(smem_u32(a_ring + sa * kTileBytesA));
#pragma unroll
                        for (int k = 0; k < 4; ++k)     // 4 x (K = 16 bf16 = 32 B) inside the 128-B swizzle span
                            umma_f16<kCG>(d_tmem, a_desc + 2u * k, b_desc + 2u * k, p.idesc,
                                           (kb | s | k) != 0 ? 1u : 0u);
                        if constexpr (kCG == 1) umma_commit<1>(&a_empty[sa]);
                        else umma_commit_2sm_mcast(&a_empty[sa], 0b11);
                        if (++sa == kStagesA) { sa = 0; pa ^= 1u; }
                    }
                    if constexpr (kCG == 1) umma_commit<1>(&b_empty[sb]);
                    else umma_commit_2sm_mcast(&b_empty[sb], 0b11);
                    if (++sb == kStagesB) { sb = 0; pb ^= 1u; }
                }
                if constexpr (kCG == 1) umma_commit<1>(&acc_full[buf]);
                else umma_commit_2sm_mcast(&acc_full[buf], 0b11);
            }
        }
    } else if (warp >= 4) {
        // ============================================ LIF epilogue (4 warps)
        const int q = warp - 4;                        // TMEM lane quadrant of this warp
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const int te = q * 32 + lane;                  // 0..127: channel of this thread inside the CTA's 128
        const int n_out = 5 * p.A;
        if (p.fuse_readout) {                          // this CTA's 128-channel slice of the two 1x1 readout convs
            const int c0 = static_cast<int>(rank) * 128;
            for (int i = te; i < kRoMaxOut * 128; i += 128) {
                const int o = i >> 7, cc = i & 127;
                float wv = 0.f;
                if (o < p.A) wv = p.w_cls[o * p.m_total + c0 + cc];
                else if (o < n_out) wv = p.w_bbox[(o - p.A) * p.m_total + c0 + cc];
                ro_w[o * kRoWStride + cc] = wv;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        uint32_t chunk_ctr = 0;
        uint32_t it = 0;
        for (int tile = group; tile < p.total_tiles; tile += n_groups, ++it) {
            const int ut = tile / p.m_tiles, mt = tile - ut * p.m_tiles;
            const int c = mt * 128 * kCG + static_cast<int>(rank) * 128 + q * 32 + lane;
            int H = 1, W = 1, n = 0, h0 = 0, w0 = 0, lvl = 0;
            uint8_t* trains = reinterpret_cast<uint8_t*>(p.trains);
            unsigned int tile_spikes = 0;
            if (p.conv) {
                while (lvl + 1 < p.n_levels && ut >= p.lv[lvl + 1].tile_begin) ++lvl;
                const LevelDesc& L = p.lv[lvl];
                int local = ut - L.tile_begin;
                const int per_img = L.tiles_w * L.tiles_h;
                n = local / per_img; local -= n * per_img;
                const int ty = local / L.tiles_w;
                h0 = ty * p.TH; w0 = (local - ty * L.tiles_w) * p.TW;
                H = L.H; W = L.W;
                trains = reinterpret_cast<uint8_t*>(L.trains);
            }
            const float wscale = __ldg(&p.w_scale[c]);
            const uint32_t buf = it & 1u;
            mbar_wait(&acc_full[buf], (it >> 1) & 1u);
            tcgen05_fence_after();
            const uint32_t acc = tmem_base + lane_addr + buf * 256u;

            for (int sub = 0; sub < kCG; ++sub) {
                for (int j0 = 0; j0 < p.Jh; j0 += CW) {
                    float v[CW], cu[CW];
                    float ii[CW], sk[CW];
                    uint32_t tr[CW];
#pragma unroll
                    for (int u = 0; u < CW; ++u) { v[u] = 0.f; ii[u] = 0.f; tr[u] = 0u; sk[u] = 0.f; }
                    for (int t = p.t0; t < p.T_total; ++t) {
                        const int tl = t - p.t0;
                        const bool live = tl < p.T_live;
                        if (live) {
                            tmem_ld<CW>(acc + static_cast<uint32_t>(sub * n_half + tl * p.Jh + j0),
                                        reinterpret_cast<uint32_t*>(cu));
                            tmem_ld_wait();
#pragma unroll
                            for (int u = 0; u < CW; ++u) cu[u] = __fmul_rn(cu[u], wscale);   // exact: power of two
                        }
                        const float kap = p.kappa[t];
#pragma unroll
                        for (int u = 0; u < CW; ++u) {
                            const float cur = live ? cu[u] : 0.0f;
                            const uint32_t z = lif_update(v[u], ii[u], cur);
                            tr[u] |= z << t;
                            sk[u] = z ? __fadd_rn(sk[u], kap) : sk[u];
                        }
                        if (p.dump != nullptr && live && !p.conv) {
#pragma unroll
                            for (int u = 0; u < CW; ++u) {
                                const int r = ut * p.J + sub * p.Jh + j0 + u;
                                if (r < p.rows)
                                    p.dump[(static_cast<size_t>(tl) * p.rows + r) * p.m_total + c] = cu[u];
                            }
                        }
                    }
                    // ---- emit spike trains (+ optional bf16 planes for the next layer)
#pragma unroll
                    for (int u = 0; u < CW; ++u) {
                        const int jh = j0 + u;
                        size_t r;
                        bool ok;
                        if (p.conv) {
                            const int ty = jh / p.TWh, tx = jh - ty * p.TWh;
                            const int h = h0 + sub * p.sub_dh + ty, w = w0 + sub * p.sub_dw + tx;
                            ok = (h < H) && (w < W);
                            r = (static_cast<size_t>(n) * H + h) * W + w;
                        } else {
                            const int rr = ut * p.J + sub * p.Jh + jh;
                            ok = rr < p.rows;
                            r = static_cast<size_t>(rr);
                        }
                        if (!ok) continue;
                        tile_spikes += __popc(tr[u]);
                        if (trains == nullptr) continue;
                        const size_t e = r * p.m_total + c;
                        if (p.train_bytes == 1) trains[e] = static_cast<uint8_t>(tr[u]);
                        else if (p.train_bytes == 2) reinterpret_cast<uint16_t*>(trains)[e] = static_cast<uint16_t>(tr[u]);
                        else reinterpret_cast<uint32_t*>(trains)[e] = tr[u];
                        if (p.spikes_out != nullptr) {
                            for (int t = p.spk_t_lo; t < p.spk_t_hi; ++t)
                                p.spikes_out[(static_cast<size_t>(t - p.spk_t_lo) * p.rows + r) * p.m_total + c] =
                                    static_cast<uint16_t>(((tr[u] >> t) & 1u) ? p.spike_one : 0u);
                        }
                    }
                    // ---- fused LI readout: out[o][px] += sum_{c in this CTA} W[o][c] * sk[c][px]
                    if (p.fuse_readout) {
                        float* S = ro_s + (chunk_ctr & 1u) * (128 * kRoSStride);
                        ++chunk_ctr;
#pragma unroll
                        for (int u = 0; u < CW; u += 4)
                            *reinterpret_cast<float4*>(&S[te * kRoSStride + u]) = make_float4(sk[u], sk[u + 1], sk[u + 2], sk[u + 3]);
                        asm volatile("bar.sync 1, 128;" ::: "memory");
                        constexpr int kGroups = 128 / CW;              // thread = (pixel u, output group og)
                        constexpr int kOPT = (kRoMaxOut + kGroups - 1) / kGroups;
                        const int u = te % CW, og = te / CW;
                        float acc[kOPT];
#pragma unroll
                        for (int k = 0; k < kOPT; ++k) acc[k] = 0.f;
                        if (og < kRoMaxOut) {
#pragma unroll 4
                            for (int cc = 0; cc < 128; cc += 4) {
                                const float s0 = S[(cc + 0) * kRoSStride + u], s1 = S[(cc + 1) * kRoSStride + u];
                                const float s2 = S[(cc + 2) * kRoSStride + u], s3 = S[(cc + 3) * kRoSStride + u];
#pragma unroll
                                for (int k = 0; k < kOPT; ++k) {
                                    const int o = og + k * kGroups;
                                    if (o < kRoMaxOut) {
                                        const float4 wv = *reinterpret_cast<const float4*>(&ro_w[o * kRoWStride + cc]);
                                        acc[k] = fmaf(wv.x, s0, acc[k]); acc[k] = fmaf(wv.y, s1, acc[k]);
                                        acc[k] = fmaf(wv.z, s2, acc[k]); acc[k] = fmaf(wv.w, s3, acc[k]);
                                    }
                                }
                            }
                            const int jh = j0 + u;
                            const int ty = jh / p.TWh, tx = jh - ty * p.TWh;
                            const int h = h0 + sub * p.sub_dh + ty, w = w0 + sub * p.sub_dw + tx;
                            if (h < H && w < W) {
                                const LevelDesc& L = p.lv[lvl];
                                const size_t hw = static_cast<size_t>(H) * W, pix = static_cast<size_t>(h) * W + w;
#pragma unroll
                                for (int k = 0; k < kOPT; ++k) {
                                    const int o = og + k * kGroups;
                                    if (o < p.A) atomicAdd(&L.logits[(static_cast<size_t>(n) * p.A + o) * hw + pix], acc[k]);
                                    else if (o < n_out)
                                        atomicAdd(&L.bbox[(static_cast<size_t>(n) * 4 * p.A + (o - p.A)) * hw + pix], acc[k]);
                                }
                            }
                        }
                    }
                }
            }
            if (p.conv && p.lv[lvl].counts != nullptr) {
                for (int o = 16; o > 0; o >>= 1) tile_spikes += __shfl_xor_sync(0xffffffffu, tile_spikes, o);
                if (lane == 0 && tile_spikes) atomicAdd(&p.lv[lvl].counts[n], static_cast<unsigned long long>(tile_spikes));
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (kCG == 1) mbar_arrive(&acc_empty[buf]);
                else mbar_arrive_cluster(&acc_empty[buf], 0);
            }
        }
    }

    tcgen05_fence_before();
    if constexpr (kCG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 2) tmem_dealloc<kCG>(tmem_base, 512);
}

}  // namespace snn
