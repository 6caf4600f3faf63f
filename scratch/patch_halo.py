p='snn_automotive_object_detection_b200/csrc/spike_gemm_lif.cuh'
s=open(p).read()

# ---------- header comment
s=s.replace('''//      fc   : words [R][K]            TMA box (64 words, Jh rows)
//      conv : words [N][H][W][C]      TMA box (64 words, TWh, THh, 1) per (tap, 64-channel block) at the
//             shifted pixel; the 3x3 halo and the image border are TMA out-of-bounds zero fill.
//    The word tile of a k-block is 1-2 KB per CTA (vs 14-16 KB for expanded planes), so the
//    L2->SM feed of the kernel is essentially the weight tiles alone.''','''//      fc   : words [R][K]            TMA box (64 words, Jh rows) per k-block; one spike tile per k-block.
//      conv : words [N][H][W][C]      ONE TMA box (64 words, 8+2, TH/kCG+2, 1) per (tile, 64-channel block): the
//             CTA's pixels plus a one-pixel halo (image border = TMA out-of-bounds zero fill).  It is expanded
//             ONCE into a halo'd spike tile with rows ordered (halo row, t, halo column); the 9 taps of the
//             3x3 conv are 9 MMA descriptors into that same tile: start row (dy * T_box * 10 + dx), 8-row
//             groups 10 rows (1280 B) apart.  This relies on the 128-byte swizzle of tcgen05.mma following
//             ABSOLUTE shared-memory address bits (scratch/swizzle_experiment.py: any 128-B row start and
//             group stride works with base_offset 0), so the producers swizzle by the absolute row address.
//    The word tile of a stage is 1-3 KB per CTA, so the L2->SM feed of the kernel is the weight tiles alone,
//    and the expansion work of the conv is 40/16 halo overhead x 1/9 = 0.28 of expanding every tap.''')

# ---------- params
s=s.replace('''    int n_pg;                 // producer groups''','''    int hrows;                // conv: halo rows per CTA = TH / kCG + 2 (halo columns = TWh + 2 = 10)
    int n_pg;                 // producer groups''')
s=s.replace('''    int dbg_shift, dbg_sbo, dbg_boff;   // swizzle experiment (scratch/swizzle_experiment.py): row shift, group stride, base-offset field''','''    int dbg_shift, dbg_sbo, dbg_boff;   // swizzle experiment (scratch/swizzle_experiment.py, fc only): row shift, group stride, base-offset field''')

# ---------- A TMA producer loop order
s=s.replace('''                for (int kb = 0; kb < p.kblocks; ++kb) {
                    for (int s = 0; s < p.nsplit; ++s) {
                        mbar_wait_parked(&a_empty[sa], pa ^ 1u);
                        if (rank == 0) mbar_expect_tx(&a_full[sa], kTileBytesA * kCG);
                        uint8_t* adst = a_ring + sa * kTileBytesA;
                        if constexpr (kCG == 1) tma_load_2d(adst, &p.tmA, &a_full[sa], kb * 64, s * p.m_total + m0);
                        else tma_load_2d_2sm(adst, &p.tmA, &a_full[sa], kb * 64, s * p.m_total + m0);
                        if (++sa == kStagesA) { sa = 0; pa ^= 1u; }
                    }
                }''','''                // k order: fc kb = 0..K/64; conv (64-channel block outer, tap inner) -- the order the MMA issuer uses
                for (int ko = 0; ko < n_outer; ++ko) {
                    for (int ki = 0; ki < n_inner; ++ki) {
                        const int kcol = kConv ? (ki * p.k_in + ko * 64) : ko * 64;
                        for (int s = 0; s < p.nsplit; ++s) {
                            mbar_wait_parked(&a_empty[sa], pa ^ 1u);
                            if (rank == 0) mbar_expect_tx(&a_full[sa], kTileBytesA * kCG);
                            uint8_t* adst = a_ring + sa * kTileBytesA;
                            if constexpr (kCG == 1) tma_load_2d(adst, &p.tmA, &a_full[sa], kcol, s * p.m_total + m0);
                            else tma_load_2d_2sm(adst, &p.tmA, &a_full[sa], kcol, s * p.m_total + m0);
                            if (++sa == kStagesA) { sa = 0; pa ^= 1u; }
                        }
                    }
                }''')
s=s.replace('''    const int n_half = p.n_mma / kCG;                 // B rows (= accumulator columns) produced per CTA
''','''    const int n_half = p.n_mma / kCG;                 // B rows (= accumulator columns) produced per CTA
    // spike-tile ring stages per tile, and MMA k-blocks per stage: fc one k-block per stage; conv one stage per
    // 64-channel block, read by the 9 taps
    const int n_outer = kConv ? p.cblocks : p.kblocks;
    const int n_inner = kConv ? 9 : 1;
''')

# ---------- MMA issuer
a=s.index('                for (int kb = 0; kb < p.kblocks; ++kb) {\n                    mbar_wait(&b_ready[sb], pb);')
b=s.index('                if constexpr (kCG == 1) umma_commit<1>(&acc_full[buf]);')
mma='''                for (int ko = 0; ko < n_outer; ++ko) {
                    mbar_wait(&b_ready[sb], pb);
                    if constexpr (kCG == 2) mbar_wait_cluster(&b_peer[sb], pb);
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
                            mbar_wait(&a_full[sa], pa);
                            tcgen05_fence_after();
                            const uint64_t a_desc = umma_desc_sw128(smem_u32(a_ring + sa * kTileBytesA));
#pragma unroll
                            for (int k = 0; k < 4; ++k)     // 4 x (K = 16 x 16-bit = 32 B) inside the 128-B swizzle span
                                umma_f16<kCG>(d_tmem, a_desc + 2u * k, b_desc + 2u * k, p.idesc,
                                              (ko | ki | s | k) != 0 ? 1u : 0u);
                            if constexpr (kCG == 1) umma_commit<1>(&a_empty[sa]);
                            else umma_commit_2sm_mcast(&a_empty[sa], 0b11);
                            if (++sa == kStagesA) { sa = 0; pa ^= 1u; }
                        }
                    }
                    if constexpr (kCG == 1) umma_commit<1>(&b_empty[sb]);
                    else umma_commit_2sm_mcast(&b_empty[sb], 0b11);
                    if (++sb == static_cast<uint32_t>(stages_b)) { sb = 0; pb ^= 1u; }
                }
'''
s=s[:a]+mma+s[b:]
# relay count
s=s.replace('''            const long long total_kb = static_cast<long long>(my_tiles) * p.kblocks;
            uint32_t pb = 0;
            for (long long i = lane; i < total_kb; i += stages_b, pb ^= 1u) {''','''            const long long total_kb = static_cast<long long>(my_tiles) * n_outer;
            uint32_t pb = 0;
            for (long long i = lane; i < total_kb; i += stages_b, pb ^= 1u) {''')

# ---------- word TMA producer
a=s.index('            const uint32_t w_bytes = static_cast<uint32_t>(p.Jh) * 64u * static_cast<uint32_t>(p.in_wb);')
b=s.index('    } else if (warp >= 8 && warp < 8 + kProducerWarps) {')
wt='''            const uint32_t w_bytes = (kConv ? static_cast<uint32_t>(p.hrows) * 10u : static_cast<uint32_t>(p.Jh)) * 64u *
                                     static_cast<uint32_t>(p.in_wb);
            for (int tile = group; tile < p.total_tiles; tile += n_groups) {
                const int ut = tile / p.m_tiles;
                const TilePos tp = decode_tile(p, ut);
                const int h0 = tp.h0 + static_cast<int>(rank) * p.sub_dh, w0 = tp.w0 + static_cast<int>(rank) * p.sub_dw;
                const int r0 = ut * p.J + static_cast<int>(rank) * p.Jh;
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
'''
s=s[:a]+wt+s[b:]

# ---------- producers: pairs / rows
s=s.replace('''        const int n_pairs = p.Jh * 8;
        const int wb = p.in_wb;''','''        // conv: every pixel of the halo'd region; fc: the CTA's units
        const int n_pairs = (kConv ? p.hrows * 10 : p.Jh) * 8;
        const int wb = p.in_wb;''')
s=s.replace('''        const long long total_kb = static_cast<long long>(my_tiles) * p.kblocks;
        const uint32_t b_base = smem_u32(b_ring), w_base = smem_u32(w_ring);
        const uint32_t row_step = static_cast<uint32_t>(p.Jh) * 128u;''','''        const long long total_kb = static_cast<long long>(my_tiles) * n_outer;
        const uint32_t b_base = smem_u32(b_ring), w_base = smem_u32(w_ring);
        // row of (unit j, step t): fc t * Jh + j; conv halo pixel (hh, ww): (hh * T_box + t) * 10 + ww
        const uint32_t row_step = (kConv ? 10u : static_cast<uint32_t>(p.Jh)) * 128u;''')
s=s.replace('''                const uint32_t j = pr >> 3, q = pr & 7;
                const uint32_t src = wslot + static_cast<uint32_t>(pr) * 8u * wb;
                uint32_t r = j, addr = slot + j * 128u;
                if (packed) {''','''                const uint32_t j = pr >> 3, q = pr & 7;
                const uint32_t src = wslot + static_cast<uint32_t>(pr) * 8u * wb;
                uint32_t r0 = j;
                if constexpr (kConv) { const uint32_t hh = j / 10u; r0 = hh * static_cast<uint32_t>(p.T_box) * 10u + (j - hh * 10u); }
                uint32_t addr = slot + r0 * 128u;      // slot is 1024-B aligned: (addr >> 7) & 7 == row & 7
                if (packed) {''')
s=s.replace('''#pragma unroll 4
                    for (int t = 0; t < p.T_box; ++t, r += p.Jh, addr += row_step) {
                        uint4 o;
                        o.x = ((P[0] >> t) & 0x00010001u) * one; o.y = ((P[1] >> t) & 0x00010001u) * one;
                        o.z = ((P[2] >> t) & 0x00010001u) * one; o.w = ((P[3] >> t) & 0x00010001u) * one;
                        if (p.dbg_sbo != 0) {  // swizzle experiment: absolute-address swizzle, rows in groups of 8 at stride dbg_sbo
                            const uint32_t ra = slot + static_cast<uint32_t>(p.dbg_shift) * 128u + (r >> 3) * p.dbg_sbo + (r & 7u) * 128u;
                            sts_v4(ra + ((q ^ ((ra >> 7) & 7u)) << 4), o);
                        } else
                        sts_v4(addr + ((q ^ (r & 7u)) << 4), o);
                    }''','''#pragma unroll 4
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
                    }''')
s=s.replace('''                    for (int t = 0; t < p.T_box; ++t, r += p.Jh, addr += row_step) {
                        uint4 o;
                        o.x = (((wv[0] >> t) & 1u) | (((wv[1] >> t) & 1u) << 16)) * one;
                        o.y = (((wv[2] >> t) & 1u) | (((wv[3] >> t) & 1u) << 16)) * one;
                        o.z = (((wv[4] >> t) & 1u) | (((wv[5] >> t) & 1u) << 16)) * one;
                        o.w = (((wv[6] >> t) & 1u) | (((wv[7] >> t) & 1u) << 16)) * one;
                        sts_v4(addr + ((q ^ (r & 7u)) << 4), o);
                    }''','''                    for (int t = 0; t < p.T_box; ++t, addr += row_step) {
                        uint4 o;
                        o.x = (((wv[0] >> t) & 1u) | (((wv[1] >> t) & 1u) << 16)) * one;
                        o.y = (((wv[2] >> t) & 1u) | (((wv[3] >> t) & 1u) << 16)) * one;
                        o.z = (((wv[4] >> t) & 1u) | (((wv[5] >> t) & 1u) << 16)) * one;
                        o.w = (((wv[6] >> t) & 1u) | (((wv[7] >> t) & 1u) << 16)) * one;
                        sts_v4(addr + ((q ^ ((addr >> 7) & 7u)) << 4), o);
                    }''')

# ---------- epilogue columns
s=s.replace('''                    uint32_t col = acc + static_cast<uint32_t>(sub * n_half + j0);
                    for (int tl = 0; tl < p.T_live; ++tl, col += p.Jh) {''','''                    // accumulator column of (unit j, step t): fc t * Jh + j; conv (tile row, t, tile column)
                    uint32_t col = acc + static_cast<uint32_t>(sub * n_half + (kConv ? (j0 >> 3) * p.T_box * 8 + (j0 & 7) : j0));
                    const uint32_t col_step = kConv ? 8u : static_cast<uint32_t>(p.Jh);
                    for (int tl = 0; tl < p.T_live; ++tl, col += col_step) {''')
open(p,'w').write(s)

p='snn_automotive_object_detection_b200/csrc/ptx.cuh'
s=open(p).read()
s=s.replace('''// instruction descriptor (kind::f16)''','''// the same with an arbitrary 128-B-aligned start row and 8-row-group stride (bytes): the swizzle follows the
// absolute shared-memory address bits, so shifted / strided windows of one tile are valid operands (base offset 0)
__device__ __forceinline__ uint64_t umma_desc_sw128_strided(uint32_t saddr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// instruction descriptor (kind::f16)''')
open(p,'w').write(s)

p='snn_automotive_object_detection_b200/csrc/snn_abi.cu'
s=open(p).read()
s=s.replace('''    p.slot_b = static_cast<int>(align_up(static_cast<size_t>(tc.n_mma / tc.cg) * 128, 1024));''','''    if (p.conv) {    // one halo'd spike tile per 64-channel block: (TH/cg + 2) x T_box x (8 + 2) rows of 128 B
        p.hrows = tc.THh + 2;
        p.slot_b = static_cast<int>(align_up(static_cast<size_t>(p.hrows) * tc.T_box * 10 * 128, 1024));
    } else {
        p.slot_b = static_cast<int>(align_up(static_cast<size_t>(tc.n_mma / tc.cg) * 128, 1024));
    }''')
s=s.replace('''    p.slot_w = static_cast<int>(align_up(static_cast<size_t>(tc.Jh) * 64 * p.in_wb, 128));''','''    p.slot_w = static_cast<int>(align_up(static_cast<size_t>(p.conv ? p.hrows * 10 : tc.Jh) * 64 * p.in_wb, 128));''')
s=s.replace('''        p.n_pg = m >= 2 ? 2 : 1;      // measured: 2 groups x 4 warps beat 4 x 2 (r01h vs r01f)''','''        p.n_pg = m >= 2 ? 2 : 1;      // measured: 2 groups x 4 warps beat 4 x 2 (r01h vs r01f)
        if (p.conv) p.n_pg = 1;       // one stage per 64-channel block serves 9 taps: all 8 warps fill it together''')
s=s.replace('''                cuuint32_t box[4] = {(cuuint32_t)(64 * wbz), (cuuint32_t)tc.TWh, (cuuint32_t)tc.THh, 1};''','''                cuuint32_t box[4] = {(cuuint32_t)(64 * wbz), (cuuint32_t)(tc.TWh + 2), (cuuint32_t)(tc.THh + 2), 1};''')
s=s.replace('''            {   // encoder words [N][H][W][C] as a byte tensor; one box = (64 words, TWh, THh) of one image''','''            {   // encoder words [N][H][W][C] as a byte tensor; one box = (64 words, TWh + 2, THh + 2) of one image''')
open(p,'w').write(s)
