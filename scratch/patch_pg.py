p='snn_automotive_object_detection_b200/csrc/spike_gemm_lif.cuh'
s=open(p).read()
s=s.replace('''constexpr int kProducerThreads = 128;       // per producer group
constexpr int kMaxPairs = 2;                // (unit, 8-channel chunk) pairs a producer thread expands per k-block''','''constexpr int kProducerWarps = 8;           // split into p.n_pg groups (1, 2 or 4); group g expands the k-blocks i = g (mod n_pg)
constexpr int kMaxPairs = 4;                // (unit, 8-channel chunk) pairs a producer thread expands per k-block
constexpr int kMaxUnitsPerCta = kMaxPairs * (kProducerWarps * 32 / 4) / 8;   // Jh <= 32: Jh * 8 pairs <= 4 x 64 threads''')
s=s.replace('''    int stages_w, slot_w;     // word ring geometry (slot = Jh units x 64 words)
''','''    int stages_w, slot_w;     // word ring geometry (slot = Jh units x 64 words)
    int n_pg;                 // producer groups (each owns every n_pg-th k-block): min(4, stages_b, stages_w) rounded to 1/2/4
''')
s=s.replace('''        for (int s = 0; s < kMaxStagesB; ++s) { mbar_init(&b_ready[s], 4); mbar_init(&b_peer[s], 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < kMaxStagesW; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 4); }''','''        const uint32_t wpg = static_cast<uint32_t>(kProducerWarps / p.n_pg);      // warps per producer group
        for (int s = 0; s < kMaxStagesB; ++s) { mbar_init(&b_ready[s], wpg); mbar_init(&b_peer[s], 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < kMaxStagesW; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], wpg); }''')
a=s.index('    } else if (warp >= 8 && warp < 16) {')
b=s.index('    } else if (warp >= 4) {')
prod=open('scratch/prod_section.txt').read()
s=s[:a]+prod+s[b:]
s=s.replace('// Raw spike-train words of 8 consecutive input neurons (8, 16 or 32 bytes).\nstruct RawWords { uint4 a, b; };\n\n','')
open(p,'w').write(s)

p='snn_automotive_object_detection_b200/csrc/snn_abi.cu'
s=open(p).read()
s=s.replace('''    const int maxJ = (kMaxPairs * kProducerThreads / 8) * cg;     // producers: Jh * 8 pairs <= kMaxPairs * 128''','''    const int maxJ = kMaxUnitsPerCta * cg;                        // producers: Jh * 8 pairs <= kMaxPairs x 64 threads''')
s=s.replace('''    // the two producer groups start on stages 0 and 1: a ring of one stage would let group 1 pass its first wait
    if (p.stages_b < 2 || p.stages_w < 2) return fail(SNN_E_ARG, "tile shape leaves fewer than two ring stages");''','''    // producer group g starts on stage g of both rings, so there are at most min(stages) groups (1, 2 or 4)
    {
        const int m = p.stages_b < p.stages_w ? p.stages_b : p.stages_w;
        p.n_pg = m >= 4 ? 4 : m >= 2 ? 2 : 1;
    }''')
open(p,'w').write(s)
