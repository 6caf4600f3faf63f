// C ABI (include/snn_heads.h) over the sm_100a kernels: workspace carving, TMA tensor maps,
// tile-shape selection and launches.  No device allocation, no synchronisation, no CPU fallback.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/snn_heads.h"
#include "aux_kernels.cuh"
#include "spike_gemm_lif.cuh"

using namespace snn;

namespace {

thread_local std::string g_err;
thread_local int g_launches = 0;
thread_local int g_force_cg = 0;
thread_local int g_fc_dual = 0;      // 0 auto, 1 never, 2 whenever the tile shape allows it (tests)
thread_local int g_fc_units = 0;     // > 0: cap on the units per fc tile (experiments)
thread_local int g_fc_split = 0;
thread_local int g_conv_mc = 0;       // 1: conv weight tiles multicast across clusters of two CTA pairs (experiment, off by default)
thread_local int g_roi_kernel = 0;   // 0 = two channel planes in flight per thread, 1 = four (A-B timing)
thread_local unsigned long long* g_role_cycles = nullptr;   // profiling: MMA-thread wait counters of the next launches
thread_local int g_role_phase = -1;                        // which launch gets them: 0 conv, 1 / 3 fc with K >= 4096 in dual / single tiles, 2 other fc     // 1: never run the last partial wave of dual tiles as single tiles (experiments)

// Optional per-phase device timing (CUDA events recorded on the launch stream around each phase).
enum Phase { PH_ENC_RPN = 0, PH_GEMM_RPN, PH_RO_RPN, PH_ENC_BOX, PH_GEMM_FC6, PH_GEMM_FC7, PH_RO_BOX, PH_COUNT };
constexpr int kMaxTimed = 256;
// Thread-local like every other knob of this file (one host thread drives one device at a time); the events belong to
// the device that was current when they were created and are re-created when the thread has moved to another device.
struct PhaseEvents { cudaEvent_t start[kMaxTimed], stop[kMaxTimed]; int created = 0, used = 0, device = -1; };
thread_local PhaseEvents g_ph[PH_COUNT];
thread_local unsigned g_profile = 0;          // bit ph: phase ph is timed
thread_local unsigned long long* g_clock_probe = nullptr;   // [pairs][2]: SM cycles / ns summed over the conv launches

void phase_begin(int ph, cudaStream_t st) {
    if (!((g_profile >> ph) & 1u)) return;
    PhaseEvents& e = g_ph[ph];
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    if (e.device != dev) {
        if (e.used > 0) return;                  // unread records of another device: do not mix
        for (int i = 0; i < e.created; ++i) { cudaEventDestroy(e.start[i]); cudaEventDestroy(e.stop[i]); }
        e.created = 0; e.device = dev;
    }
    if (e.used >= kMaxTimed) return;
    if (e.used >= e.created) {
        if (cudaEventCreate(&e.start[e.created]) != cudaSuccess || cudaEventCreate(&e.stop[e.created]) != cudaSuccess) return;
        ++e.created;
    }
    cudaEventRecord(e.start[e.used], st);
}
void phase_end(int ph, cudaStream_t st) {
    if (!((g_profile >> ph) & 1u)) return;
    PhaseEvents& e = g_ph[ph];
    if (e.used >= kMaxTimed || e.used >= e.created) return;
    cudaEventRecord(e.stop[e.used], st);
    ++e.used;
}

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess) return fail(SNN_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// With SNN_PDL=1 every kernel of a forward is launched with programmatic stream serialization: it may be scheduled
// while its predecessor drains and runs its prologue up to griddep_wait() (ptx.cuh).  Off by default: measured r01av
// (145 GPU tests green with it), sustained 770-791 img/s with vs 790-791 without -- the path is bound by the board's
// power cap, so the ~15 us of idle time it removes per kernel boundary come back as a lower SM clock.
bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("SNN_PDL"); return e && e[0] == '1'; }();
    return on;
}
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// --------------------------------------------------------------------------- device / driver
// Host-side state that is immutable after first use is cached: device attributes per device, the
// max-dynamic-shared-memory attribute per (kernel, device), and encoded TMA tensor maps keyed by everything
// cuTensorMapEncodeTiled takes.  A forward then makes no driver query and no descriptor encode in steady state.
struct DeviceInfo { int sms = 0; int cc_major = 0; int dev = -1; bool ok = false; };
constexpr int kMaxDevices = 64;
std::mutex g_host_mutex;
DeviceInfo g_dev_info[kMaxDevices];

int device_info(DeviceInfo& di) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 0 && dev < kMaxDevices) {
        std::lock_guard<std::mutex> lk(g_host_mutex);
        if (g_dev_info[dev].ok) { di = g_dev_info[dev]; return SNN_OK; }
    }
    CUDA_TRY(cudaDeviceGetAttribute(&di.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    CUDA_TRY(cudaDeviceGetAttribute(&di.sms, cudaDevAttrMultiProcessorCount, dev));
    if (di.cc_major != 10)
        return fail(SNN_E_ARCH, "device compute capability %d.x is not sm_100 (B200); there is no fallback path",
                    di.cc_major);
    di.dev = dev; di.ok = true;
    if (dev >= 0 && dev < kMaxDevices) {
        std::lock_guard<std::mutex> lk(g_host_mutex);
        g_dev_info[dev] = di;
    }
    return SNN_OK;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize once per (kernel instantiation, device)
cudaError_t ensure_dyn_smem(const void* kern, int bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    static std::vector<std::pair<const void*, int>> done;
    {
        std::lock_guard<std::mutex> lk(g_host_mutex);
        for (const auto& d : done) if (d.first == kern && d.second == dev) return cudaSuccess;
    }
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(g_host_mutex);
    done.emplace_back(kern, dev);
    return cudaSuccess;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// Tensor-map cache (thread-local: no lock on the forward path).  The key holds every input of the encode, so a hit
// returns a byte-identical descriptor; the map is a pure function of the key (it holds the address, not the data).
struct TmapKey {
    const void* base; int rank; int words;
    cuuint64_t dims[4]; cuuint64_t strides[3]; cuuint32_t box[4];
    bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
        uint64_t h = 1469598103934665603ull;
        for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) { h ^= w[i]; h *= 1099511628211ull; }
        return static_cast<size_t>(h);
    }
};
static_assert(sizeof(TmapKey) % 8 == 0, "TmapKey is hashed as 64-bit words");
constexpr size_t kTmapCacheMax = 1024;
thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmaps;
thread_local unsigned long long g_tmap_hits = 0, g_tmap_misses = 0;

// innermost dim contiguous, zero fill out of bounds.  words = false: 16-bit weight tensor, 128-byte swizzle
// (tensor-core operand tiles); words = true: byte tensor of spike-train words, no swizzle (dense box).
int make_tmap(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
              const cuuint32_t* box, bool words = false) {
    TmapKey key;
    memset(&key, 0, sizeof(key));
    key.base = base; key.rank = rank; key.words = words ? 1 : 0;
    for (int i = 0; i < rank && i < 4; ++i) { key.dims[i] = dims[i]; key.box[i] = box[i]; }
    for (int i = 0; i + 1 < rank && i < 3; ++i) key.strides[i] = strides_bytes[i];
    auto it = g_tmaps.find(key);
    if (it != g_tmaps.end()) { *m = it->second; ++g_tmap_hits; return SNN_OK; }
    ++g_tmap_misses;
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) return fail(SNN_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(m, words ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                    static_cast<cuuint32_t>(rank), const_cast<void*>(base), dims, strides_bytes, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, words ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(SNN_E_CUDA, "cuTensorMapEncodeTiled failed (%d), rank %d dims %llu %llu box %u %u", (int)r, rank,
                    (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    if (g_tmaps.size() >= kTmapCacheMax) g_tmaps.clear();      // shapes / addresses keep changing: start over
    g_tmaps.emplace(key, *m);
    return SNN_OK;
}

int nsplit_of(int mode) {
    switch (mode) {
        case SNN_MODE_FP32_EXACT: return 3;
        case SNN_MODE_BF16: return 1;
        case SNN_MODE_BF16X2: return 2;
        case SNN_MODE_FP16X2: return 2;
        case SNN_MODE_FP16: return 1;
        default: return 0;
    }
}
bool is_fp16(int mode) { return mode == SNN_MODE_FP16X2 || mode == SNN_MODE_FP16; }
uint32_t one_of(int mode) { return is_fp16(mode) ? 0x3C00u : 0x3F80u; }
// the per-row accumulator scales sit behind the 16-bit pieces of a prepared weight
const float* scale_of(const void* w_prep, int rows, int cols, int mode) {
    return reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(w_prep) +
                                          static_cast<size_t>(nsplit_of(mode)) * rows * cols * 2);
}


// sleep between polls of the conv producers' long wait (ptx.cuh mbar_wait_backoff).  r02q A/B on one box, 400 steps each:
// 0 ns 764.8 / 764.0 img/s, 100 ns 760.9, 300 ns 769.4 / 771.1, 1000 ns 763.2, 3000 ns 762.8 (profiles/r02/r02q_backoff.txt)
constexpr int kConvBackoffNs = 300;

// --------------------------------------------------------------------------- tile selection
struct TileCfg { int cg, T_box, J, Jh, TW, TH, TWh, THh, dw, dh, CW, n_mma; };

// Fold time into the MMA N dimension: N = T_box * J <= 256, N % 16 == 0, J units per tile.
// cap_units > 0: no more than that many units per tile.
bool pick_tile_cg(int T_live, bool conv, int cg, TileCfg& out, int cap_units = 0) {
    const int step = conv ? 8 * cg : 2 * cg;      // fc: units per CTA a multiple of 2 (epilogue chunks of 2, 4 or 8)
    int maxJ = kMaxUnitsPerCta * cg;                              // producers: Jh * 8 pairs <= kMaxPairs x 64 threads
    if (!conv && g_fc_units >= step && g_fc_units < maxJ) maxJ = g_fc_units;
    if (cap_units >= step && cap_units < maxJ) maxJ = cap_units;
    int bestJ = 0, bestT = 0;
    for (int J = step; J <= maxJ; J += step)
        for (int Tb = T_live; Tb <= T_live + 1; ++Tb) {
            const int n = Tb * J;
            if (n > 256 || (n % 16) != 0 || n < 16) continue;
            if (J > bestJ) { bestJ = J; bestT = Tb; }
        }
    if (bestJ == 0) return false;
    // fc: a padding step (T_box = T_live + 1) is tensor work on zero rows.  Prefer the widest tile WITHOUT padding
    // when it keeps N >= 160 (narrower MMAs are bound by the shared-memory reads of the weight operand); such a
    // tile runs as a dual tile (two accumulators per weight tile), which restores the weight reuse.
    if (!conv && cg == 2 && bestT != T_live) {
        for (int J = bestJ - step; J >= step; J -= step) {
            const int n = T_live * J;
            if (n <= 256 && (n % 16) == 0 && n >= 160) { bestJ = J; bestT = T_live; break; }
        }
    }
    out.cg = cg; out.T_box = bestT; out.J = bestJ; out.Jh = bestJ / cg; out.n_mma = bestT * bestJ;
    if (conv) {
        out.TW = 8; out.TH = bestJ / 8; out.TWh = 8; out.THh = out.TH / cg; out.dw = 0; out.dh = (cg == 2) ? out.THh : 0;
    } else {
        out.TW = out.TH = out.TWh = out.THh = 0; out.dw = out.dh = 0;
    }
    out.CW = (out.Jh % 8 == 0) ? 8 : (out.Jh % 4 == 0) ? 4 : 2;
    return true;
}
bool pick_tile(int T_live, bool conv, int m_total, int force_cg, TileCfg& out) {
    if (force_cg != 1 && (m_total % 256) == 0 && pick_tile_cg(T_live, conv, 2, out)) return true;
    if (force_cg == 2) return false;
    return pick_tile_cg(T_live, conv, 1, out);
}

template <int kCG>
cudaError_t launch_gemm_cw(const GemmLifParams& p, int CW, int grid, cudaStream_t st, bool mc = false) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = kGemmSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = mc ? 2 * kCG : kCG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
#define SNN_LAUNCH(CWV, CONV)                                                                                  \
    {                                                                                                          \
        auto kern = spike_gemm_lif_kernel<kCG, CWV, CONV>;                                                     \
        cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(kern), (int)kGemmSmemBytes);             \
        if (e != cudaSuccess) return e;                                                                        \
        return cudaLaunchKernelEx(&cfg, kern, p);                                                              \
    }
    if (p.conv) {
        if constexpr (kCG == 2) {
            if (mc && CW == 8) {
                auto kern = spike_gemm_lif_kernel<2, 8, true, false, true>;
                cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(kern), (int)kGemmSmemBytes);
                if (e != cudaSuccess) return e;
                return cudaLaunchKernelEx(&cfg, kern, p);
            }
        }
        if (CW == 8) SNN_LAUNCH(8, true) else SNN_LAUNCH(4, true)
    } else if (p.dual) {
        if constexpr (kCG == 2) {
#define SNN_LAUNCH_DUAL(CWV)                                                                                   \
    {                                                                                                          \
        auto kern = spike_gemm_lif_kernel<2, CWV, false, true>;                                                \
        cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(kern), (int)kGemmSmemBytes);             \
        if (e != cudaSuccess) return e;                                                                        \
        return cudaLaunchKernelEx(&cfg, kern, p);                                                              \
    }
            if (CW == 8) SNN_LAUNCH_DUAL(8) else if (CW == 4) SNN_LAUNCH_DUAL(4) else SNN_LAUNCH_DUAL(2)
#undef SNN_LAUNCH_DUAL
        } else {
            return cudaErrorInvalidValue;
        }
    } else {
        if (CW == 8) SNN_LAUNCH(8, false) else if (CW == 4) SNN_LAUNCH(4, false) else SNN_LAUNCH(2, false)
    }
#undef SNN_LAUNCH
}

int launch_gemm(GemmLifParams& p, const TileCfg& tc, const DeviceInfo& di, int mode, cudaStream_t st) {
    p.J = tc.J; p.Jh = tc.Jh; p.TW = tc.TW; p.TH = tc.TH; p.TWh = tc.TWh; p.THh = tc.THh;
    p.sub_dw = tc.dw; p.sub_dh = tc.dh; p.T_box = tc.T_box; p.n_mma = tc.n_mma;
    p.idesc = umma_idesc_f16(128 * tc.cg, tc.n_mma, !is_fp16(mode));
    p.spike_one = one_of(mode);
    if (p.conv) {    // one halo'd spike tile per 64-channel block: (TH/cg + 2) x T_box x (8 + 2) rows of 128 B
        p.hrows = tc.THh + 2;
        p.slot_b = static_cast<int>(align_up(static_cast<size_t>(p.hrows) * tc.T_box * 10 * 128, 1024));
    } else {
        p.slot_b = static_cast<int>(align_up(static_cast<size_t>(tc.n_mma / tc.cg) * 128, 1024)) * (p.dual ? 2 : 1);
    }
    // The weight ring (16 KB stages) and the spike-tile ring share 193 KB.
    const int ring_total = kRingBytesAB;
    if (p.conv) {
        // A conv spike tile (one 64-channel block of the halo'd region, read by 9 taps) is large: two of them in
        // flight (producers refill one while the MMAs read the other), the rest is weight stages -- the conv's MMA
        // thread waits mostly for weight tiles (r01ad role counters: 16 % of its time with 6 stages)
        p.stages_b = 2 * p.slot_b + 3 * kTileBytesA <= ring_total ? 2 : 1;
        if (p.slot_b <= 16 * 1024 && 3 * p.slot_b + 6 * kTileBytesA <= ring_total) p.stages_b = 3;
        p.stages_a = (ring_total - p.stages_b * p.slot_b) / kTileBytesA;
    } else {
        // fc: one spike tile per k-block, handed producer -> relay -> MMA -> commit across the CTA pair; that chain
        // is a few thousand cycles, so the spike ring wants depth more than the weight ring does (measured r01ad,
        // dual tiles of 22 KB, 2 pieces: 5 weight + 5 spike stages 0.632 ms, 6 + 4 0.666, 4 + 5 0.645, 3 + 6 0.700);
        // with 3 pieces per weight the weight ring is the one that must stay deep
        const int a_min = p.nsplit >= 3 ? 6 : 4;
        p.stages_b = (ring_total - a_min * kTileBytesA) / p.slot_b;
        if (p.stages_b > kMaxStagesB) p.stages_b = kMaxStagesB;
        if (p.stages_b < 1) p.stages_b = 1;
        p.stages_a = (ring_total - p.stages_b * p.slot_b) / kTileBytesA;
    }
    if (p.stages_a > kStagesA) p.stages_a = kStagesA;
    if (p.stages_a < 2 || p.stages_b < 1)
        return fail(SNN_E_ARG, "spike tile of %d bytes does not fit shared memory", p.slot_b);
    if (const char* e = getenv("SNN_DBG_SWIZZLE")) {       // "shift,sbo,boff" -- scratch/swizzle_experiment.py only
        int a = 0, b = 0, c = 0;
        if (sscanf(e, "%d,%d,%d", &a, &b, &c) == 3 && b >= 1024) {
            p.dbg_shift = a; p.dbg_sbo = b; p.dbg_boff = c;
            p.slot_b = static_cast<int>(align_up(static_cast<size_t>(tc.n_mma / tc.cg + 7) / 8 * b + 2048, 1024));
            p.stages_b = (ring_total - 4 * kTileBytesA) / p.slot_b;
            if (p.stages_b > kMaxStagesB) p.stages_b = kMaxStagesB;
            p.stages_a = 4;
        }
    }
    p.slot_w = static_cast<int>(align_up(static_cast<size_t>(p.conv ? p.hrows * 10 : tc.Jh * (p.dual ? 2 : 1)) * 64 * p.in_wb, 128));
    p.stages_w = kRingBytesW / p.slot_w;
    if (p.stages_w > kMaxStagesW) p.stages_w = kMaxStagesW;
    // producer group g starts on stage g of both rings, so there are at most min(stages) groups (1, 2 or 4)
    {
        const int m = p.stages_b < p.stages_w ? p.stages_b : p.stages_w;
        p.n_pg = m >= 2 ? 2 : 1;      // pair-loop producers: 2 groups x 4 warps beat 4 x 2 (r01h vs r01f)
        if (p.conv) p.n_pg = 1;       // one stage per 64-channel block serves 9 taps: all 8 warps fill it together
        if (!p.conv && tc.T_box <= 16) p.n_pg = 1;   // fc item-parallel producers: all 8 warps fill every stage
    }
    if (const char* e = getenv("SNN_DBG_STAGES")) {        // "a,b": ring depths, profiling experiments only
        int a = 0, b = 0;
        if (sscanf(e, "%d,%d", &a, &b) == 2 && a >= 2 && a <= kStagesA && b >= 1 && b <= kMaxStagesB &&
            static_cast<size_t>(a) * kTileBytesA + static_cast<size_t>(b) * p.slot_b <= static_cast<size_t>(ring_total) && !p.conv) {
            p.stages_a = a; p.stages_b = b;
        }
    }
    {
        const int phase = p.conv ? 0 : (p.kblocks >= 64 ? (p.dual ? 1 : 3) : 2);
        p.role_cycles = (g_role_cycles != nullptr && phase == g_role_phase) ? g_role_cycles : nullptr;
        p.clock_probe = (phase == 0) ? g_clock_probe : nullptr;
    }
    {   // conv producers: sleep between polls while the MMAs work through a spike tile (SNN_DBG_BACKOFF=ns overrides)
        static const int backoff = [] { const char* e = getenv("SNN_DBG_BACKOFF"); return e ? atoi(e) : kConvBackoffNs; }();
        p.wait_backoff_ns = p.conv ? backoff : 0;
        static const int reuse = [] { const char* e = getenv("SNN_DBG_PROD_REUSE"); return e ? atoi(e) : 1; }();
        p.prod_word_reuse = reuse;
    }
    p.m_tiles = p.m_total / (128 * tc.cg);
    p.total_tiles = p.unit_tiles * p.m_tiles;
    if (p.total_tiles <= 0) return SNN_OK;
    int groups = di.sms / tc.cg;
    if (groups > p.total_tiles) groups = p.total_tiles;
    // weight multicast (experiment): clusters of two CTA pairs, an even number of pairs
    const bool mc = p.conv && g_conv_mc == 1 && tc.cg == 2 && tc.CW == 8 && p.total_tiles >= 2;
    if (mc) groups &= ~1;
    const int grid = groups * tc.cg;
    cudaError_t e = (tc.cg == 2) ? launch_gemm_cw<2>(p, tc.CW, grid, st, mc) : launch_gemm_cw<1>(p, tc.CW, grid, st);
    if (e != cudaSuccess) return fail(SNN_E_CUDA, "spike_gemm_lif launch failed: %s", cudaGetErrorString(e));
    ++g_launches;
    return SNN_OK;
}

// kappa[t] = weight of a spike at step t in the last leaky-integrator membrane: 0.9^{T-t} - 0.8^{T-t}
KappaTable kappa_table(int T) {
    KappaTable kt;
    for (int t = 0; t < 32; ++t) kt.k[t] = (t < T) ? pow(0.9, T - t) - pow(0.8, T - t) : 0.0;
    return kt;
}

int word_bytes(int nbits) { return nbits <= 8 ? 1 : nbits <= 16 ? 2 : 4; }

// x [total] fp32 -> words of wb bytes (total a multiple of 16)
cudaError_t launch_encode_rows(const float* x, size_t total, int T_live, int wb, uint8_t* z, int sms, cudaStream_t st) {
    const size_t total16 = total / 16;
    const size_t want = (total16 + 255) / 256;
    const int blocks = static_cast<int>(want > static_cast<size_t>(sms) * 8 ? static_cast<size_t>(sms) * 8 : want);
    cudaError_t e = cudaSuccess;
    SNN_ENC_BUCKETS(T_live, {
        if (wb == 1) e = launch_pdl(encode_rows_kernel<NT, 1>, dim3(blocks), dim3(256), 0, st, x, total16, T_live, z);
        else if (wb == 2) e = launch_pdl(encode_rows_kernel<NT, 2>, dim3(blocks), dim3(256), 0, st, x, total16, T_live, z);
        else e = launch_pdl(encode_rows_kernel<NT, 4>, dim3(blocks), dim3(256), 0, st, x, total16, T_live, z);
    });
    return e;
}

// More than kPassSteps live steps run as several launches over the time axis with the neuron state carried in HBM
constexpr int kPassSteps = 16;
// steps per pass: the time axis is cut evenly (23 live steps = 12 + 11, not 16 + 7 with nine all-zero rows per unit)
int pass_steps(int T_live) {
    if (T_live <= kPassSteps) return T_live;
    const int n_pass = (T_live + kPassSteps - 1) / kPassSteps;
    return (T_live + n_pass - 1) / n_pass;
}

struct RpnWs { size_t z_off[kMaxLevels], tr_off[kMaxLevels], st_off[kMaxLevels], lut_off, total; };

int rpn_ws_layout(const int* H, const int* W, int L, int N, int C, int T, int mode, RpnWs& ws, TileCfg& tc) {
    if (L < 1 || L > kMaxLevels) return fail(SNN_E_ARG, "n_levels %d outside [1,%d]", L, kMaxLevels);
    if (T < 1 || T > 32) return fail(SNN_E_ARG, "num_steps %d outside [1,32]", T);
    if (C % 128 != 0 || C < 128) return fail(SNN_E_ARG, "in_channels %d must be a multiple of 128", C);
    if (nsplit_of(mode) == 0) return fail(SNN_E_ARG, "unknown mode %d", mode);
    if (N < 1) return fail(SNN_E_ARG, "batch size %d < 1", N);
    const int T_live = T - 1;
    tc = TileCfg{};
    if (T_live > 0 && !pick_tile(pass_steps(T_live), true, C, g_force_cg, tc)) return fail(SNN_E_ARG, "no tile shape for T=%d", T);
    const int tb = snn_train_word_bytes(T);
    size_t off = 0;
    for (int l = 0; l < L; ++l) {
        if (H[l] < 1 || W[l] < 1) return fail(SNN_E_ARG, "level %d has empty spatial size", l);
        ws.z_off[l] = off;
        off = align_up(off + static_cast<size_t>(N) * H[l] * W[l] * C * word_bytes(T_live), 1024);   // encoder words
    }
    for (int l = 0; l < L; ++l) {
        ws.tr_off[l] = off;
        off = align_up(off + static_cast<size_t>(N) * H[l] * W[l] * C * tb, 1024);
    }
    for (int l = 0; l < L; ++l) {       // carried neuron state (v, i, kappa sum, train word), multi-pass only
        ws.st_off[l] = off;
        if (T_live > kPassSteps) off = align_up(off + static_cast<size_t>(N) * H[l] * W[l] * C * sizeof(float4), 1024);
    }
    ws.lut_off = off;
    off += kMaxTrainBytes * 256 * sizeof(float);
    ws.total = align_up(off, 1024);
    return SNN_OK;
}

struct BoxWs { size_t z_off, tr6_off, tr7_off, st_off, lut_off, total; };

int box_ws_layout(int R, int K, int Hd, int T, int mode, bool stats, BoxWs& ws, TileCfg& t6, TileCfg& t7) {
    if (T < 1 || T > 32) return fail(SNN_E_ARG, "num_steps %d outside [1,32]", T);
    if (R < 1) return fail(SNN_E_ARG, "R %d < 1", R);
    if (K % 64 != 0) return fail(SNN_E_ARG, "in_channels %d must be a multiple of 64", K);
    if (Hd % 128 != 0) return fail(SNN_E_ARG, "representation_size %d must be a multiple of 128", Hd);
    if (nsplit_of(mode) == 0) return fail(SNN_E_ARG, "unknown mode %d", mode);
    // worst case over stats on/off so one workspace size serves both.  T < 3: no fc6 current reaches an output (the
    // reference's membranes are all zero there); fc6 still runs for T = 2 when its spikes are asked for.
    TileCfg a{}, b{};
    t6 = TileCfg{}; t7 = TileCfg{};
    if (T >= 2 && !pick_tile(pass_steps(T - 1), false, Hd, g_force_cg, a)) return fail(SNN_E_ARG, "no tile shape for T=%d", T);
    if (T >= 3 && (!pick_tile(pass_steps(T - 2), false, Hd, g_force_cg, b) || !pick_tile(pass_steps(T - 2), false, Hd, g_force_cg, t7)))
        return fail(SNN_E_ARG, "no tile shape for T=%d", T);
    t6 = stats ? a : b;
    const int tb = snn_train_word_bytes(T);
    size_t off = 0;
    ws.z_off = off; off = align_up(off + static_cast<size_t>(R) * K * word_bytes(T - 1), 1024);   // encoder words
    ws.tr6_off = off; off = align_up(off + static_cast<size_t>(R) * Hd * tb, 1024);
    ws.tr7_off = off; off = align_up(off + static_cast<size_t>(R) * Hd * tb, 1024);
    ws.st_off = off;                    // carried neuron state, multi-pass only (fc6's passes finish before fc7's start)
    if (T - 1 > kPassSteps) off = align_up(off + static_cast<size_t>(R) * Hd * sizeof(float4), 1024);
    ws.lut_off = off; off += kMaxTrainBytes * 256 * sizeof(float);
    ws.total = align_up(off, 1024);
    return SNN_OK;
}

// one launch over the units [0, R) of z_words / trains; dump (debug) is [T_live][dump_rows][M] and already offset
int fc_launch(const DeviceInfo& di, const void* z_words, int in_wb, int in_bit0, int R, int K, int M, int T, int t0,
              int T_live, int mode, const void* w_prep, void* trains, float* dump, int dump_rows, const TileCfg& tc,
              bool dual, float4* state, int state_load, int state_store, cudaStream_t st) {
    GemmLifParams p;
    memset(&p, 0, sizeof(p));
    const int nsplit = nsplit_of(mode);
    p.w_scale = scale_of(w_prep, M, K, mode);
    {
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)nsplit * M};
        cuuint64_t str[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {64, 128};
        int rc = make_tmap(&p.tmA, w_prep, 2, dims, str, box);
        if (rc) return rc;
    }
    p.n_levels = 1; p.conv = 0; p.n_images = 1;
    p.m_total = M; p.nsplit = nsplit; p.kblocks = K / 64; p.cblocks = 1; p.k_in = K;
    p.T_total = T; p.t0 = t0; p.T_live = T_live;
    p.dual = dual ? 1 : 0;
    const int tile_units = tc.J * (p.dual ? 2 : 1);
    p.rows = R; p.unit_tiles = (R + tile_units - 1) / tile_units;
    p.train_bytes = snn_train_word_bytes(T);
    p.in_wb = in_wb; p.in_bit0 = in_bit0;
    {   // input words [R][K] as a byte tensor
        cuuint64_t dims[2] = {(cuuint64_t)K * in_wb, (cuuint64_t)R};
        cuuint64_t str[1] = {(cuuint64_t)K * in_wb};
        cuuint32_t box[2] = {(cuuint32_t)(64 * in_wb), (cuuint32_t)(tc.Jh * (p.dual ? 2 : 1))};
        int rc = make_tmap(&p.tmW[0], z_words, 2, dims, str, box, true);
        if (rc) return rc;
    }
    p.trains = trains;
    p.dump = dump; p.dump_rows = dump_rows;
    p.state = state; p.state_load = state_load; p.state_store = state_store;
    return launch_gemm(p, tc, di, mode, st);
}

// One fully-connected spiking layer.  Dual tiles (2J units, both accumulator buffers fed from every weight tile) pay
// off for long contractions: the weight stream from L2 and the cross-CTA hand-offs per FLOP halve; the fc epilogue of
// a tile is then exposed, so short K stays single.  The persistent grid runs whole waves of tiles; when the last
// wave of dual tiles would be less than half full, its units are run as single tiles by a second launch (half the
// time of a dual wave).
int fc_layer(const DeviceInfo& di, const void* z_words, int in_wb, int in_bit0, int R, int K, int M, int T, int t0,
             int T_live, int mode, const void* w_prep, void* trains, float* dump, const TileCfg& tc, cudaStream_t st,
             float4* state = nullptr, int state_load = 0, int state_store = 0) {
    const bool dual_ok = tc.cg == 2 && tc.T_box <= 16 && (tc.n_mma / 2) % 8 == 0;
    bool dual = dual_ok && g_fc_dual != 1 && (g_fc_dual == 2 || K / 64 >= 64);
    int R1 = R;                                   // units [0, R1) in the first launch
    TileCfg tail = tc;                            // tile shape of the second launch
    bool tail_dual = false;
    if (dual && g_fc_dual == 0) {
        const int groups = di.sms / 2, m_tiles = M / 256, U2 = 2 * tc.J;
        const int ut_total = (R + U2 - 1) / U2;
        const int full = ut_total * m_tiles / groups;                 // complete waves of dual tiles
        if (full == 0) {
            if ((R + tc.J - 1) / tc.J * m_tiles <= groups) dual = false;   // one wave either way: single tiles are half as long
        } else if (ut_total * m_tiles > full * groups && g_fc_split != 1) {
            // the units left after the full waves fit ONE wave of smaller tiles: single tiles of the same shape
            // (measured r01ae: 0.622 ms), or -- tail_split 2 -- narrower dual tiles (N = 96: 0.631 ms)
            const int UT1 = full * groups / m_tiles;
            const int R2 = R - UT1 * U2;
            const int step = 2 * tc.cg;
            for (int cap = step; cap < tc.J && g_fc_split == 2; cap += step) {
                TileCfg t2{};
                if (!pick_tile_cg(T_live, false, 2, t2, cap) || t2.J != cap) continue;
                if ((t2.n_mma / 2) % 8 != 0 || t2.n_mma < 96) continue;
                if ((R2 + 2 * t2.J - 1) / (2 * t2.J) * m_tiles > groups) continue;
                tail = t2; tail_dual = true; R1 = UT1 * U2;
                break;
            }
            if (R1 == R && (R2 + tc.J - 1) / tc.J * m_tiles <= groups) R1 = UT1 * U2;
        }
    }
    const size_t tb = snn_train_word_bytes(T);
    int rc = fc_launch(di, z_words, in_wb, in_bit0, R1, K, M, T, t0, T_live, mode, w_prep, trains, dump, R, tc, dual,
                       state, state_load, state_store, st);
    if (rc || R1 == R) return rc;
    rc = fc_launch(di, reinterpret_cast<const uint8_t*>(z_words) + static_cast<size_t>(R1) * K * in_wb, in_wb, in_bit0,
                   R - R1, K, M, T, t0, T_live, mode, w_prep,
                   reinterpret_cast<uint8_t*>(trains) + static_cast<size_t>(R1) * M * tb,
                   dump ? dump + static_cast<size_t>(R1) * M : nullptr, R, tail, tail_dual,
                   state ? state + static_cast<size_t>(R1) * M : nullptr, state_load, state_store, st);
    return rc;
}

// The same over a long time axis: passes of <= kPassSteps live steps, the neuron state carried through `state`
// ([R][M] float4, caller's workspace) between the launches; only the last pass drains the synapse and emits.
int fc_layer_passes(const DeviceInfo& di, const void* z_words, int in_wb, int in_bit0, int R, int K, int M, int T, int t0,
                    int T_live, int mode, const void* w_prep, void* trains, const TileCfg& tc, float4* state,
                    cudaStream_t st) {
    if (T_live <= kPassSteps)
        return fc_layer(di, z_words, in_wb, in_bit0, R, K, M, T, t0, T_live, mode, w_prep, trains, nullptr, tc, st);
    const int per = pass_steps(T_live);
    for (int b = 0; b < T_live; b += per) {
        const int n = T_live - b < per ? T_live - b : per;
        const int rc = fc_layer(di, z_words, in_wb, in_bit0 + b, R, K, M, T, t0 + b, n, mode, w_prep, trains, nullptr, tc, st,
                                state, b > 0 ? 1 : 0, b + n < T_live ? 1 : 0);
        if (rc) return rc;
    }
    return SNN_OK;
}

template <typename T>
cudaError_t launch_readout_rpn(const void* trains, int C, int HW, int N, const float* wc, int n_a, const float* wb, int n_b,
                               const KappaTable& lut, float* lo, float* bo, unsigned long long* counts, cudaStream_t st) {
    const size_t smem = static_cast<size_t>(n_a + n_b) * C * 4 + 256 * sizeof(T) * 4 +
                        static_cast<size_t>(kRpnRoPx) * ((C * sizeof(T)) / 4 + 1) * 4;
    auto kern = readout_rpn_kernel<T>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // size varies with C / outputs
    if (e != cudaSuccess) return e;
    dim3 grid((HW + kRpnRoPx - 1) / kRpnRoPx, N);
    return launch_pdl(kern, grid, dim3(kRpnRoPx), smem, st, reinterpret_cast<const T*>(trains), C, HW, wc, wb, n_a, n_b, lut, lo, bo, counts);
}

template <typename T>
cudaError_t launch_readout_rows(const void* tr7, const void* tr6, int R, int Hd, const float* wc, int nc,
                                const float* wb, int nb, const KappaTable& lut, float* oc, float* ob, unsigned int* counts,
                                cudaStream_t st) {
    const size_t smem = static_cast<size_t>(kRoRows) * Hd * 4;
    auto kern = readout_rows_kernel<T>;
    if (smem > 40 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    return launch_pdl(kern, dim3((R + kRoRows - 1) / kRoRows), dim3(256), smem, st, reinterpret_cast<const T*>(tr7),
                      reinterpret_cast<const T*>(tr6), R, Hd, wc, nc, wb, nb, lut, oc, ob, counts);
}

}  // namespace

extern "C" {

int snn_version(void) { return SNN_ABI_VERSION; }
const char* snn_last_error(void) { return g_err.c_str(); }
int snn_last_launch_count(void) { return g_launches; }
void snn_set_cta_group(int cg) { g_force_cg = (cg == 1 || cg == 2) ? cg : 0; }
void snn_set_role_timers(unsigned long long* device_counters, int phase) {
    g_role_cycles = device_counters;
    g_role_phase = phase;
}
void snn_set_clock_probe(unsigned long long* device_counters) { g_clock_probe = device_counters; }
void snn_set_roi_kernel(int which) { g_roi_kernel = which == 1 ? 1 : 0; }
void snn_set_conv_multicast(int on) { g_conv_mc = on == 1 ? 1 : 0; }
void snn_host_cache_stats(unsigned long long* hits, unsigned long long* misses) {
    if (hits) *hits = g_tmap_hits;
    if (misses) *misses = g_tmap_misses;
}
void snn_set_fc_tiling(int dual, int max_units, int tail_split) {
    g_fc_dual = (dual == 1 || dual == 2) ? dual : 0;
    g_fc_units = max_units > 0 ? max_units : 0;
    g_fc_split = (tail_split == 1 || tail_split == 2) ? tail_split : 0;
}
int snn_train_word_bytes(int T) { return T <= 8 ? 1 : T <= 16 ? 2 : 4; }
int snn_mode_pieces(int mode) { return nsplit_of(mode); }

size_t snn_prepared_weight_bytes(int rows, int cols, int mode) {
    return static_cast<size_t>(nsplit_of(mode)) * rows * cols * 2 + static_cast<size_t>(rows) * sizeof(float);
}

int snn_prepare_conv3x3_weights(const float* w, int O, int C, int mode, void* out, snn_stream_t stream) {
    const int ns = nsplit_of(mode);
    if (ns == 0 || O < 1 || C < 1 || !w || !out) return fail(SNN_E_ARG, "prepare_conv3x3: bad argument");
    if ((C * 9) % 8 != 0) return fail(SNN_E_ARG, "prepare_conv3x3: 9*C must be a multiple of 8");
    float* scale = const_cast<float*>(scale_of(out, O, 9 * C, mode));
    prep_weights_kernel<true><<<O < 2048 ? O : 2048, 256, 0, (cudaStream_t)stream>>>(
        w, O, 9 * C, C, ns, is_fp16(mode) ? 1 : 0, reinterpret_cast<uint16_t*>(out), scale);
    CUDA_TRY(cudaGetLastError());
    return SNN_OK;
}

int snn_prepare_fc_weights(const float* w, int O, int K, int mode, void* out, snn_stream_t stream) {
    const int ns = nsplit_of(mode);
    if (ns == 0 || O < 1 || K < 1 || !w || !out) return fail(SNN_E_ARG, "prepare_fc: bad argument");
    if (K % 2 != 0) return fail(SNN_E_ARG, "prepare_fc: K must be even");
    float* scale = const_cast<float*>(scale_of(out, O, K, mode));
    prep_weights_kernel<false><<<O < 2048 ? O : 2048, 256, 0, (cudaStream_t)stream>>>(
        w, O, K, 0, ns, is_fp16(mode) ? 1 : 0, reinterpret_cast<uint16_t*>(out), scale);
    CUDA_TRY(cudaGetLastError());
    return SNN_OK;
}

size_t snn_rpn_head_workspace_bytes(const int* H, const int* W, int n_levels, int N, int C_in, int T, int mode) {
    RpnWs ws; TileCfg tc;
    if (rpn_ws_layout(H, W, n_levels, N, C_in, T, mode, ws, tc) != SNN_OK) return 0;
    return ws.total;
}

int snn_rpn_head_forward(const void* const* feat_ptrs, const int* H, const int* W, int n_levels, int N, int C_in,
                         int A, int T, int mode, const void* w_shared_prep, const float* w_cls, const float* w_bbox,
                         void* const* logits_out, void* const* bbox_out, void* const* spike_trains_out,
                         unsigned long long* spike_counts_out, void* workspace, size_t workspace_bytes,
                         snn_stream_t stream) {
    g_launches = 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (!feat_ptrs || !H || !W || !w_shared_prep || !w_cls || !w_bbox || !logits_out || !bbox_out || !workspace)
        return fail(SNN_E_ARG, "rpn_head_forward: null argument");
    if (A < 1) return fail(SNN_E_ARG, "num_anchors %d < 1", A);
    DeviceInfo di;
    int rc = device_info(di);
    if (rc) return rc;
    RpnWs ws; TileCfg tc;
    rc = rpn_ws_layout(H, W, n_levels, N, C_in, T, mode, ws, tc);
    if (rc) return rc;
    if (workspace_bytes < ws.total)
        return fail(SNN_E_WORKSPACE, "workspace %zu B < required %zu B", workspace_bytes, ws.total);
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return fail(SNN_E_ARG, "workspace must be 256-B aligned");
    uint8_t* wsp = reinterpret_cast<uint8_t*>(workspace);
    const int tb = snn_train_word_bytes(T);
    const int ns = nsplit_of(mode);
    const int T_live = T - 1;
    const KappaTable lut = kappa_table(T);

    // The LI readout (and the spike counts) are fused into the GEMM epilogue when a CTA pair covers all
    // output channels (cta_group 2, C_in == 256) and the 5A outputs fit one pass; otherwise a separate
    // readout kernel consumes the spike trains (and needs the kappa lookup table).
    const bool fused = T_live > 0 && tc.cg == 2 && tc.CW == 8 && C_in == 256 && 5 * A <= kRoMaxOut;
    void* trains[kMaxLevels];
    for (int l = 0; l < n_levels; ++l) {
        trains[l] = (spike_trains_out && spike_trains_out[l]) ? spike_trains_out[l] : (wsp + ws.tr_off[l]);
        if (!feat_ptrs[l] || !logits_out[l] || !bbox_out[l]) return fail(SNN_E_ARG, "level %d: null pointer", l);
    }

    if (T_live > 0) {
        if (fused) {
            // The two CTAs of a pair add their channel halves onto zeroed outputs.  Outputs the caller laid out back to
            // back (logits of all levels, then box deltas of all levels) are cleared with ONE memset -- issued before
            // the encoder, so that encoder -> GEMM stay adjacent kernels of the stream (programmatic dependent launch).
            bool adjacent = true;
            const uint8_t* expect = reinterpret_cast<const uint8_t*>(logits_out[0]);
            for (int l = 0; l < n_levels && adjacent; ++l) {
                adjacent = reinterpret_cast<const uint8_t*>(logits_out[l]) == expect;
                expect += static_cast<size_t>(N) * A * H[l] * W[l] * 4;
            }
            for (int l = 0; l < n_levels && adjacent; ++l) {
                adjacent = reinterpret_cast<const uint8_t*>(bbox_out[l]) == expect;
                expect += static_cast<size_t>(N) * 4 * A * H[l] * W[l] * 4;
            }
            if (adjacent)
                CUDA_TRY(cudaMemsetAsync(logits_out[0], 0, expect - reinterpret_cast<const uint8_t*>(logits_out[0]), st));
            for (int l = 0; l < n_levels && !adjacent; ++l) {
                CUDA_TRY(cudaMemsetAsync(logits_out[l], 0, static_cast<size_t>(N) * A * H[l] * W[l] * 4, st));
                CUDA_TRY(cudaMemsetAsync(bbox_out[l], 0, static_cast<size_t>(N) * 4 * A * H[l] * W[l] * 4, st));
            }
        }
        // 1) encoder: fp32 NCHW features -> NHWC spike-train words (bit t = z_t)
        phase_begin(PH_ENC_RPN, st);
        {
            EncParams ep;
            memset(&ep, 0, sizeof(ep));
            int chunks = 0;
            for (int l = 0; l < n_levels; ++l) {
                EncLevel& E = ep.lv[l];
                E.x = reinterpret_cast<const float*>(feat_ptrs[l]);
                E.z = wsp + ws.z_off[l];
                E.HW = H[l] * W[l]; E.chunks = (E.HW + kEncPx - 1) / kEncPx; E.chunk_begin = chunks;
                chunks += N * E.chunks;
            }
            ep.n_levels = n_levels; ep.N = N; ep.C = C_in; ep.T_live = T_live; ep.wb = word_bytes(T_live);
            ep.total_items = chunks * (C_in / kEncCh);
            const int blocks = (ep.total_items + 7) / 8 < di.sms * 8 ? (ep.total_items + 7) / 8 : di.sms * 8;
            cudaError_t le = cudaSuccess;
            SNN_ENC_BUCKETS(T_live, {
                if (ep.wb == 1) le = launch_pdl(encode_nchw_kernel<NT, 1>, dim3(blocks), dim3(256), 0, st, ep);
                else if (ep.wb == 2) le = launch_pdl(encode_nchw_kernel<NT, 2>, dim3(blocks), dim3(256), 0, st, ep);
                else le = launch_pdl(encode_nchw_kernel<NT, 4>, dim3(blocks), dim3(256), 0, st, ep);
            });
            if (le != cudaSuccess) return fail(SNN_E_CUDA, "encode_nchw launch failed: %s", cudaGetErrorString(le));
            ++g_launches;
        }
        phase_end(PH_ENC_RPN, st);
        // 2) all levels, all images: implicit-GEMM 3x3 conv + LIF recurrence in one persistent launch
        GemmLifParams p;
        memset(&p, 0, sizeof(p));
        {
            cuuint64_t dims[2] = {(cuuint64_t)9 * C_in, (cuuint64_t)ns * C_in};
            cuuint64_t str[1] = {(cuuint64_t)9 * C_in * 2};
            cuuint32_t box[2] = {64, 128};
            rc = make_tmap(&p.tmA, w_shared_prep, 2, dims, str, box);
            if (rc) return rc;
            if (g_conv_mc == 1) {          // experiment: half tiles for the multicast schedule
                cuuint32_t box_half[2] = {64, 64};
                rc = make_tmap(&p.tmA_half, w_shared_prep, 2, dims, str, box_half);
                if (rc) return rc;
            }
        }
        int tiles = 0;
        for (int l = 0; l < n_levels; ++l) {
            LevelDesc& L = p.lv[l];
            L.H = H[l]; L.W = W[l];
            L.tiles_w = (W[l] + tc.TW - 1) / tc.TW; L.tiles_h = (H[l] + tc.TH - 1) / tc.TH;
            L.tile_begin = tiles; L.trains = trains[l];
            {   // encoder words [N][H][W][C] as a byte tensor; one box = (64 words, TWh + 2, THh + 2) of one image
                const cuuint64_t wbz = (cuuint64_t)word_bytes(T_live);
                cuuint64_t dims[4] = {(cuuint64_t)C_in * wbz, (cuuint64_t)W[l], (cuuint64_t)H[l], (cuuint64_t)N};
                cuuint64_t str[3] = {(cuuint64_t)C_in * wbz, (cuuint64_t)W[l] * C_in * wbz, (cuuint64_t)H[l] * W[l] * C_in * wbz};
                cuuint32_t box[4] = {(cuuint32_t)(64 * wbz), (cuuint32_t)(tc.TWh + 2), (cuuint32_t)(tc.THh + 2), 1};
                rc = make_tmap(&p.tmW[l], wsp + ws.z_off[l], 4, dims, str, box, true);
                if (rc) return rc;
            }
            L.logits = reinterpret_cast<float*>(logits_out[l]); L.bbox = reinterpret_cast<float*>(bbox_out[l]);
            L.counts = spike_counts_out ? spike_counts_out + static_cast<size_t>(l) * N : nullptr;
            tiles += L.tiles_w * L.tiles_h * N;
        }
        if (fused) {
            p.fuse_readout = 1;
            // the two CTAs of a pair add their channel halves onto zeroed outputs.  Outputs the caller laid out
            // back to back (logits of all levels, then box deltas of all levels) are cleared with ONE memset.
            // (zeroed before the encoder launch, above)
            for (int l = 0; l < n_levels; ++l)
                if (!(spike_trains_out && spike_trains_out[l])) p.lv[l].trains = nullptr;   // nobody reads them
        } else {
            for (int l = 0; l < n_levels; ++l) p.lv[l].counts = nullptr;   // the readout kernel counts instead
        }
        p.A = A; p.w_cls = w_cls; p.w_bbox = w_bbox;
        for (int t = 0; t < 32; ++t)
            p.kappa[t] = (t < T) ? static_cast<float>(pow(0.9, T - t) - pow(0.8, T - t)) : 0.f;
        p.n_levels = n_levels; p.conv = 1; p.n_images = N;
        p.m_total = C_in; p.nsplit = ns; p.cblocks = C_in / 64; p.kblocks = 9 * p.cblocks; p.k_in = C_in;
        p.in_wb = word_bytes(T_live); p.in_bit0 = 0;
        p.T_total = T; p.t0 = 0; p.T_live = T_live;
        p.rows = 0; p.unit_tiles = tiles; p.train_bytes = tb;
        p.w_scale = scale_of(w_shared_prep, C_in, 9 * C_in, mode);
        phase_begin(PH_GEMM_RPN, st);
        if (T_live <= kPassSteps) {
            rc = launch_gemm(p, tc, di, mode, st);
        } else {
            // long time axis: passes of <= 16 live steps, the neuron state carried through the workspace; only the
            // last pass drains, emits the spike trains and runs the fused readout
            const int fuse = p.fuse_readout;
            for (int l = 0; l < n_levels; ++l) p.lv[l].state = reinterpret_cast<float4*>(wsp + ws.st_off[l]);
            const int per = pass_steps(T_live);
            for (int b = 0; b < T_live && rc == SNN_OK; b += per) {
                const int n = T_live - b < per ? T_live - b : per;
                const bool last = b + n >= T_live;
                p.t0 = b; p.in_bit0 = b; p.T_live = n;
                p.state_load = b > 0 ? 1 : 0; p.state_store = last ? 0 : 1;
                p.fuse_readout = last ? fuse : 0;
                rc = launch_gemm(p, tc, di, mode, st);
            }
        }
        phase_end(PH_GEMM_RPN, st);
        if (rc) return rc;
    } else {
        for (int l = 0; l < n_levels; ++l)
            CUDA_TRY(cudaMemsetAsync(trains[l], 0, static_cast<size_t>(N) * H[l] * W[l] * C_in * tb, st));
    }
    // 3) leaky-integrator readouts (objectness + box deltas) from the spike trains (unless fused above)
    phase_begin(PH_RO_RPN, st);
    for (int l = 0; l < n_levels && !fused; ++l) {
        unsigned long long* cnt = spike_counts_out ? spike_counts_out + static_cast<size_t>(l) * N : nullptr;
        cudaError_t e;
        float* lo = reinterpret_cast<float*>(logits_out[l]);
        float* bo = reinterpret_cast<float*>(bbox_out[l]);
        if (tb == 1) e = launch_readout_rpn<uint8_t>(trains[l], C_in, H[l] * W[l], N, w_cls, A, w_bbox, 4 * A, lut, lo, bo, cnt, st);
        else if (tb == 2) e = launch_readout_rpn<uint16_t>(trains[l], C_in, H[l] * W[l], N, w_cls, A, w_bbox, 4 * A, lut, lo, bo, cnt, st);
        else e = launch_readout_rpn<uint32_t>(trains[l], C_in, H[l] * W[l], N, w_cls, A, w_bbox, 4 * A, lut, lo, bo, cnt, st);
        if (e != cudaSuccess) return fail(SNN_E_CUDA, "readout_rpn launch failed: %s", cudaGetErrorString(e));
        ++g_launches;
    }
    phase_end(PH_RO_RPN, st);
    return SNN_OK;
}

size_t snn_box_head_workspace_bytes(int R, int K, int Hdim, int T, int mode) {
    BoxWs ws; TileCfg a, b;
    if (box_ws_layout(R, K, Hdim, T, mode, true, ws, a, b) != SNN_OK) return 0;
    return ws.total;
}

static int box_head_forward_impl(const void* x, bool x_is_words, int R, int K, int Hdim, int C, int n_box_out, int T,
                                 int mode, const void* w6_prep, const void* w7_prep, const float* w_cls,
                                 const float* w_bbox, float* cls_out, float* bbox_out, void* spk6_trains,
                                 void* spk7_trains, unsigned int* spike_counts_out, void* workspace,
                                 size_t workspace_bytes, snn_stream_t stream) {
    g_launches = 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (!x || !w6_prep || !w7_prep || !w_cls || !w_bbox || !cls_out || !bbox_out || !workspace)
        return fail(SNN_E_ARG, "box_head_forward: null argument");
    if (C < 1 || n_box_out < 1) return fail(SNN_E_ARG, "box_head_forward: bad output sizes");
    DeviceInfo di;
    int rc = device_info(di);
    if (rc) return rc;
    // the last lif6 step feeds no output; it is evaluated only when its spikes are asked for
    const bool stats = spike_counts_out != nullptr || spk6_trains != nullptr;
    BoxWs ws; TileCfg t6, t7;
    rc = box_ws_layout(R, K, Hdim, T, mode, stats, ws, t6, t7);
    if (rc) return rc;
    if (workspace_bytes < ws.total)
        return fail(SNN_E_WORKSPACE, "workspace %zu B < required %zu B", workspace_bytes, ws.total);
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return fail(SNN_E_ARG, "workspace must be 256-B aligned");
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) return fail(SNN_E_ARG, "x must be 16-B aligned");
    uint8_t* wsp = reinterpret_cast<uint8_t*>(workspace);
    const int tb = snn_train_word_bytes(T);
    const int ns = nsplit_of(mode);
    const KappaTable lut = kappa_table(T);
    void* tr6 = spk6_trains ? spk6_trains : (wsp + ws.tr6_off);
    void* tr7 = spk7_trains ? spk7_trains : (wsp + ws.tr7_off);
    // x_is_words: the caller already holds the encoder's spike-train words (snn_roi_align_encode), [R][K] words of
    // word_bytes(T - 1) bytes; bits beyond the live steps are ignored by the producers
    const void* z = x_is_words ? x : static_cast<const void*>(wsp + ws.z_off);

    // fc6 is live for steps 0..T-3 (its last two steps never reach the outputs); one more step when
    // the fc6 spike statistics are wanted.  fc7 is live for steps 1..T-2.
    const int T_live6 = stats ? T - 1 : T - 2;
    const int T_live7 = T - 2;
    const size_t tr_bytes = static_cast<size_t>(R) * Hdim * tb;
    float4* carried = reinterpret_cast<float4*>(wsp + ws.st_off);
    if (T_live6 >= 1) {
        if (!x_is_words) {
            phase_begin(PH_ENC_BOX, st);
            cudaError_t le = launch_encode_rows(reinterpret_cast<const float*>(x), static_cast<size_t>(R) * K, T_live6,
                                                word_bytes(T - 1), wsp + ws.z_off, di.sms, st);
            phase_end(PH_ENC_BOX, st);
            if (le != cudaSuccess) return fail(SNN_E_CUDA, "encode_rows launch failed: %s", cudaGetErrorString(le));
            ++g_launches;
        }
        phase_begin(PH_GEMM_FC6, st);
        rc = fc_layer_passes(di, z, word_bytes(T - 1), 0, R, K, Hdim, T, 0, T_live6, mode, w6_prep, tr6, t6, carried, st);
        phase_end(PH_GEMM_FC6, st);
        if (rc) return rc;
    } else {
        CUDA_TRY(cudaMemsetAsync(tr6, 0, tr_bytes, st));           // T = 1 (or T = 2 without statistics): lif6 never spikes
    }
    if (T_live7 >= 1) {
        phase_begin(PH_GEMM_FC7, st);
        // fc7 contracts lif6's spike-train words directly: its step t0 = 1 is bit 1 of the word
        rc = fc_layer_passes(di, tr6, tb, 1, R, Hdim, Hdim, T, 1, T_live7, mode, w7_prep, tr7, t7, carried, st);
        phase_end(PH_GEMM_FC7, st);
        if (rc) return rc;
    } else {
        // T < 3: a lif6 spike (step 1 at the earliest) reaches lif7's membrane at step 2 -- after the last step; the
        // reference's membranes are exactly zero (faster_rcnn.py:492-516 with num_steps 1 or 2), and so is the readout
        // of all-zero lif7 trains below
        CUDA_TRY(cudaMemsetAsync(tr7, 0, tr_bytes, st));
    }
    phase_begin(PH_RO_BOX, st);
    cudaError_t e;
    const void* tr6_for_counts = spike_counts_out ? tr6 : nullptr;
    if (tb == 1) e = launch_readout_rows<uint8_t>(tr7, tr6_for_counts, R, Hdim, w_cls, C, w_bbox, n_box_out, lut, cls_out, bbox_out, spike_counts_out, st);
    else if (tb == 2) e = launch_readout_rows<uint16_t>(tr7, tr6_for_counts, R, Hdim, w_cls, C, w_bbox, n_box_out, lut, cls_out, bbox_out, spike_counts_out, st);
    else e = launch_readout_rows<uint32_t>(tr7, tr6_for_counts, R, Hdim, w_cls, C, w_bbox, n_box_out, lut, cls_out, bbox_out, spike_counts_out, st);
    if (e != cudaSuccess) return fail(SNN_E_CUDA, "readout_rows launch failed: %s", cudaGetErrorString(e));
    ++g_launches;
    phase_end(PH_RO_BOX, st);
    return SNN_OK;
}

int snn_box_head_forward(const void* x, int R, int K, int Hdim, int C, int n_box_out, int T, int mode,
                         const void* w6_prep, const void* w7_prep, const float* w_cls, const float* w_bbox,
                         float* cls_out, float* bbox_out, void* spk6_trains, void* spk7_trains,
                         unsigned int* spike_counts_out, void* workspace, size_t workspace_bytes,
                         snn_stream_t stream) {
    return box_head_forward_impl(x, false, R, K, Hdim, C, n_box_out, T, mode, w6_prep, w7_prep, w_cls, w_bbox, cls_out,
                                 bbox_out, spk6_trains, spk7_trains, spike_counts_out, workspace, workspace_bytes, stream);
}

int snn_box_head_forward_encoded(const void* words, int R, int K, int Hdim, int C, int n_box_out, int T, int mode,
                                 const void* w6_prep, const void* w7_prep, const float* w_cls, const float* w_bbox,
                                 float* cls_out, float* bbox_out, void* spk6_trains, void* spk7_trains,
                                 unsigned int* spike_counts_out, void* workspace, size_t workspace_bytes,
                                 snn_stream_t stream) {
    return box_head_forward_impl(words, true, R, K, Hdim, C, n_box_out, T, mode, w6_prep, w7_prep, w_cls, w_bbox, cls_out,
                                 bbox_out, spk6_trains, spk7_trains, spike_counts_out, workspace, workspace_bytes, stream);
}

int snn_roi_align_encode(const void* const* feat_ptrs, const int* H, const int* W, const float* scales, int n_levels,
                         int C, const float* rois, const int* roi_level, int R, int pooled_size, int sampling_ratio,
                         int T_live, void* words_out, float* pooled_out, snn_stream_t stream) {
    if (!feat_ptrs || !H || !W || !scales || !rois || !roi_level || !words_out)
        return fail(SNN_E_ARG, "roi_align_encode: null argument");
    if (n_levels < 1 || n_levels > kRoiMaxLevels || C < 1 || R < 1 || pooled_size < 1 || T_live < 1 || T_live > 32 ||
        (C * pooled_size * pooled_size) % 8 != 0)
        return fail(SNN_E_ARG, "roi_align_encode: unsupported sizes (levels %d, C %d, R %d, pooled %d, T_live %d)", n_levels, C,
                    R, pooled_size, T_live);
    RoiEncParams p;
    memset(&p, 0, sizeof(p));
    for (int l = 0; l < n_levels; ++l) {
        if (!feat_ptrs[l] || H[l] < 1 || W[l] < 1) return fail(SNN_E_ARG, "roi_align_encode: level %d: bad argument", l);
        p.lv[l].x = reinterpret_cast<const float*>(feat_ptrs[l]); p.lv[l].H = H[l]; p.lv[l].W = W[l]; p.lv[l].scale = scales[l];
    }
    p.n_levels = n_levels; p.C = C; p.R = R; p.P = pooled_size; p.sampling = sampling_ratio; p.T_live = T_live;
    p.wb = word_bytes(T_live);
    p.rois = rois; p.roi_level = roi_level; p.words = reinterpret_cast<uint8_t*>(words_out); p.pooled = pooled_out;
    const size_t items = static_cast<size_t>(R) * pooled_size * pooled_size * ((C + kRoiChPerThread - 1) / kRoiChPerThread);
    const int blocks = static_cast<int>((items + 255) / 256 > 148 * 32 ? 148 * 32 : (items + 255) / 256);
    if (g_roi_kernel == 1) {
        SNN_ENC_BUCKETS(T_live, (roi_align_encode_kernel<NT, 4><<<blocks, 256, 0, (cudaStream_t)stream>>>(p)));
    } else {
        SNN_ENC_BUCKETS(T_live, (roi_align_encode_kernel<NT, 2><<<blocks, 256, 0, (cudaStream_t)stream>>>(p)));
    }
    CUDA_TRY(cudaGetLastError());
    return SNN_OK;
}

int snn_rpn_decode_selected(const void* const* logits, const void* const* deltas, const float* const* base_anchors,
                            const int* H, const int* W, const int* stride_h, const int* stride_w, const int* k_per_level,
                            int n_levels, int N, int A, const long long* idx, float* boxes_out, float* scores_out,
                            float* logits_out, long long* ref_index_out, snn_stream_t stream) {
    if (!logits || !deltas || !base_anchors || !H || !W || !stride_h || !stride_w || !k_per_level || !idx || !boxes_out ||
        !scores_out)
        return fail(SNN_E_ARG, "rpn_decode_selected: null argument");
    if (n_levels < 1 || n_levels > kPropMaxLevels || N < 1 || A < 1)
        return fail(SNN_E_ARG, "rpn_decode_selected: bad sizes (levels %d, N %d, A %d)", n_levels, N, A);
    PropParams p;
    memset(&p, 0, sizeof(p));
    int kb = 0;
    long long ab = 0;
    for (int l = 0; l < n_levels; ++l) {
        PropLevel& L = p.lv[l];
        if (!logits[l] || !deltas[l] || !base_anchors[l] || H[l] < 1 || W[l] < 1 || k_per_level[l] < 0 ||
            k_per_level[l] > A * H[l] * W[l])
            return fail(SNN_E_ARG, "rpn_decode_selected: level %d: bad argument", l);
        L.logits = reinterpret_cast<const float*>(logits[l]); L.deltas = reinterpret_cast<const float*>(deltas[l]);
        L.base = base_anchors[l];
        L.H = H[l]; L.W = W[l]; L.k = k_per_level[l]; L.k_begin = kb; L.stride_h = stride_h[l]; L.stride_w = stride_w[l];
        L.anchor_begin = ab;
        kb += k_per_level[l];
        ab += static_cast<long long>(A) * H[l] * W[l];
    }
    if (kb == 0) return SNN_OK;
    p.n_levels = n_levels; p.N = N; p.A = A; p.K_total = kb;
    p.clip = static_cast<float>(log(1000.0 / 16.0));
    p.idx = idx; p.boxes = boxes_out; p.scores = scores_out; p.logit_out = logits_out; p.ref_index = ref_index_out;
    const int total = N * kb;
    rpn_decode_selected_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p);
    CUDA_TRY(cudaGetLastError());
    return SNN_OK;
}

int snn_rpn_topk_keys(const void* const* logits, const int* H, const int* W, int n_levels, int N, int A,
                      long long* const* keys_out, snn_stream_t stream) {
    if (!logits || !H || !W || !keys_out) return fail(SNN_E_ARG, "rpn_topk_keys: null argument");
    if (n_levels < 1 || n_levels > kPropMaxLevels || N < 1 || A < 1)
        return fail(SNN_E_ARG, "rpn_topk_keys: bad sizes (levels %d, N %d, A %d)", n_levels, N, A);
    KeyParams p;
    memset(&p, 0, sizeof(p));
    long long total = 0;
    for (int l = 0; l < n_levels; ++l) {
        if (!logits[l] || !keys_out[l] || H[l] < 1 || W[l] < 1) return fail(SNN_E_ARG, "rpn_topk_keys: level %d: bad argument", l);
        p.lv[l].logits = reinterpret_cast<const float*>(logits[l]); p.lv[l].keys = keys_out[l];
        p.lv[l].A = A; p.lv[l].HW = H[l] * W[l]; p.lv[l].begin = total;
        total += static_cast<long long>(N) * A * H[l] * W[l];
    }
    p.n_levels = n_levels; p.N = N; p.total = total;
    const long long want = (total + 255) / 256;
    rpn_topk_keys_kernel<<<static_cast<int>(want < 148 * 8 ? want : 148 * 8), 256, 0, (cudaStream_t)stream>>>(p);
    CUDA_TRY(cudaGetLastError());
    return SNN_OK;
}

// ---- per-level top-k of the objectness logits, all levels and images in eight launches (radix select + sort)
struct TopkWs { size_t prefix, k_rem, done, hist, n_cand, cand, zero_begin, zero_end, total; };
static TopkWs topk_ws_layout(int n_levels, int N) {
    const size_t segs = static_cast<size_t>(n_levels) * N;
    TopkWs w{};
    size_t off = 0;
    w.prefix = off; off = align_up(off + segs * 8, 256);
    w.k_rem = off; off = align_up(off + segs * 4, 256);
    w.cand = off; off = align_up(off + segs * kTopkMaxK * 8, 256);
    w.zero_begin = off;                                   // everything from here on is cleared by one memset per call
    w.done = off; off = align_up(off + segs * kTopkPasses * 4, 256);
    w.n_cand = off; off = align_up(off + segs * 4, 256);
    w.hist = off; off = align_up(off + segs * kTopkPasses * kTopkBins * 4, 256);
    w.zero_end = off;
    w.total = off;
    return w;
}

size_t snn_rpn_topk_workspace_bytes(int n_levels, int N) {
    if (n_levels < 1 || n_levels > kPropMaxLevels || N < 1) return 0;
    return topk_ws_layout(n_levels, N).total;
}

namespace {
struct RpnNmsWs { size_t keys, cnt, done, total; };
RpnNmsWs rpn_nms_ws_layout(int n_levels, int N) {
    RpnNmsWs w;
    size_t off = 0;
    w.keys = off; off = align_up(off + static_cast<size_t>(N) * n_levels * kRpnNmsMaxLevel * sizeof(unsigned long long), 256);
    w.cnt = off; off = align_up(off + static_cast<size_t>(N) * n_levels * sizeof(int), 256);
    w.done = off; off = align_up(off + static_cast<size_t>(N) * sizeof(unsigned int), 256);
    w.total = off;
    return w;
}
}  // namespace

size_t snn_rpn_nms_workspace_bytes(int n_levels, int N) {
    if (n_levels < 1 || n_levels > kPropMaxLevels || N < 1) return 0;
    return rpn_nms_ws_layout(n_levels, N).total;
}

int snn_rpn_nms(const float* proposals, const float* probs, const int* level_sizes, const int* img_h, const int* img_w,
                int n_levels, int N, float min_size, float score_thresh, float nms_thresh, int post_nms_top_n,
                float* out_boxes, float* out_scores, int* out_counts, void* workspace, size_t workspace_bytes,
                snn_stream_t stream) {
    if (!proposals || !probs || !level_sizes || !img_h || !img_w || !out_boxes || !out_scores || !out_counts || !workspace)
        return fail(SNN_E_ARG, "rpn_nms: null argument");
    if (n_levels < 1 || n_levels > kPropMaxLevels || N < 1 || post_nms_top_n < 1)
        return fail(SNN_E_ARG, "rpn_nms: bad sizes (levels %d, N %d, post_nms_top_n %d)", n_levels, N, post_nms_top_n);
    RpnNmsParams p;
    memset(&p, 0, sizeof(p));
    long long K = 0, kept_max = 0;
    for (int l = 0; l < n_levels; ++l) {
        if (level_sizes[l] < 0 || level_sizes[l] > kRpnNmsMaxLevel)
            return fail(SNN_E_ARG, "rpn_nms: level %d has %d entries per image (limit %d)", l, level_sizes[l], kRpnNmsMaxLevel);
        p.k_begin[l] = static_cast<int>(K);
        K += level_sizes[l];
        kept_max += level_sizes[l] < post_nms_top_n ? level_sizes[l] : post_nms_top_n;
    }
    p.k_begin[n_levels] = static_cast<int>(K);
    if (kept_max > kRpnNmsMaxKept) return fail(SNN_E_ARG, "rpn_nms: up to %lld keepers per image (limit %d)", kept_max, kRpnNmsMaxKept);
    if (((reinterpret_cast<uintptr_t>(proposals) | reinterpret_cast<uintptr_t>(out_boxes)) & 15) != 0)
        return fail(SNN_E_ARG, "rpn_nms: boxes must be 16-byte aligned");
    const RpnNmsWs w = rpn_nms_ws_layout(n_levels, N);
    if (workspace_bytes < w.total) return fail(SNN_E_WORKSPACE, "workspace %zu B < required %zu B", workspace_bytes, w.total);
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return fail(SNN_E_ARG, "workspace must be 256-B aligned");
    DeviceInfo di;
    if (int rc = device_info(di)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(ensure_dyn_smem(reinterpret_cast<const void*>(rpn_nms_kernel), kRpnNmsSmemBytes));
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    CUDA_TRY(cudaMemsetAsync(ws + w.done, 0, static_cast<size_t>(N) * sizeof(unsigned int), st));
    p.L = n_levels; p.K = static_cast<int>(K); p.post_n = post_nms_top_n;
    p.min_size = min_size; p.score_thresh = score_thresh; p.nms_thresh = nms_thresh;
    for (int b0 = 0; b0 < N; b0 += kDetMaxImages) {
        const int nb = N - b0 < kDetMaxImages ? N - b0 : kDetMaxImages;
        p.N = nb;
        p.props = proposals + static_cast<size_t>(b0) * K * 4; p.probs = probs + static_cast<size_t>(b0) * K;
        p.out_boxes = out_boxes + static_cast<size_t>(b0) * post_nms_top_n * 4;
        p.out_scores = out_scores + static_cast<size_t>(b0) * post_nms_top_n; p.out_counts = out_counts + b0;
        p.kept_keys = reinterpret_cast<unsigned long long*>(ws + w.keys) + static_cast<size_t>(b0) * n_levels * kRpnNmsMaxLevel;
        p.kept_cnt = reinterpret_cast<int*>(ws + w.cnt) + static_cast<size_t>(b0) * n_levels;
        p.done = reinterpret_cast<unsigned int*>(ws + w.done) + b0;
        for (int b = 0; b < nb; ++b) {
            if (img_h[b0 + b] < 0 || img_w[b0 + b] < 0) return fail(SNN_E_ARG, "rpn_nms: image %d: bad size", b0 + b);
            p.img_h[b] = static_cast<float>(img_h[b0 + b]); p.img_w[b] = static_cast<float>(img_w[b0 + b]);
        }
        rpn_nms_kernel<<<dim3(n_levels, nb), kDetThreads, kRpnNmsSmemBytes, st>>>(p);
    }
    CUDA_TRY(cudaGetLastError());
    return SNN_OK;
}

int snn_det_postprocess_max_candidates(void) { return kDetMaxCand; }

int snn_det_postprocess(const float* scores, const float* boxes, const int* rois_per_image, const int* img_h,
                        const int* img_w, int N, int C, float score_thresh, float nms_thresh, float min_size,
                        int detections_per_img, int cap, float* all_boxes, float* out_boxes, float* out_scores,
                        long long* out_labels, int* out_counts, snn_stream_t stream) {
    if (!rois_per_image || !img_h || !img_w || !out_counts) return fail(SNN_E_ARG, "det_postprocess: null argument");
    if (N < 1 || C < 2 || detections_per_img < 0 || cap < 1) return fail(SNN_E_ARG, "det_postprocess: bad sizes (N %d, C %d, cap %d)", N, C, cap);
    long long total = 0;
    int max_rows = 0;
    for (int b = 0; b < N; ++b) {
        if (rois_per_image[b] < 0 || img_h[b] < 0 || img_w[b] < 0) return fail(SNN_E_ARG, "det_postprocess: image %d: bad argument", b);
        if (static_cast<long long>(rois_per_image[b]) * (C - 1) > kDetMaxCand)
            return fail(SNN_E_ARG, "det_postprocess: image %d has %d RoIs x %d classes > %d candidates", b, rois_per_image[b], C - 1, kDetMaxCand);
        total += rois_per_image[b];
        if (rois_per_image[b] > max_rows) max_rows = rois_per_image[b];
    }
    if (total > 0x7FFFFFFFll / (4ll * C)) return fail(SNN_E_ARG, "det_postprocess: too many RoIs");
    if (cap < detections_per_img + max_rows) return fail(SNN_E_ARG, "det_postprocess: cap %d < detections_per_img + max RoIs per image (%d)", cap, detections_per_img + max_rows);
    if (total > 0 && (!scores || !boxes || !all_boxes)) return fail(SNN_E_ARG, "det_postprocess: null tensor");
    if (!out_boxes || !out_scores || !out_labels) return fail(SNN_E_ARG, "det_postprocess: null output");
    if (((reinterpret_cast<uintptr_t>(boxes) | reinterpret_cast<uintptr_t>(all_boxes)) & 15) != 0)
        return fail(SNN_E_ARG, "det_postprocess: boxes must be 16-byte aligned");
    DeviceInfo di;
    if (int rc = device_info(di)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(ensure_dyn_smem(reinterpret_cast<const void*>(det_postprocess_kernel), kDetSmemBytes));
    int row0 = 0;
    for (int b0 = 0; b0 < N; b0 += kDetMaxImages) {
        const int nb = N - b0 < kDetMaxImages ? N - b0 : kDetMaxImages;
        DetParams p;
        memset(&p, 0, sizeof(p));
        p.scores = scores; p.boxes = boxes; p.all_boxes = all_boxes;
        p.out_boxes = out_boxes + static_cast<size_t>(b0) * cap * 4; p.out_scores = out_scores + static_cast<size_t>(b0) * cap;
        p.out_labels = out_labels + static_cast<size_t>(b0) * cap; p.out_counts = out_counts + static_cast<size_t>(b0) * 2;
        p.N = nb; p.C = C; p.cap = cap; p.det_per_img = detections_per_img;
        p.score_thresh = score_thresh; p.nms_thresh = nms_thresh; p.min_size = min_size;
        for (int b = 0; b < nb; ++b) {
            p.row0[b] = row0; p.rows[b] = rois_per_image[b0 + b];
            p.img_h[b] = static_cast<float>(img_h[b0 + b]); p.img_w[b] = static_cast<float>(img_w[b0 + b]);
            row0 += rois_per_image[b0 + b];
        }
        det_postprocess_kernel<<<nb, kDetThreads, kDetSmemBytes, st>>>(p);
    }
    CUDA_TRY(cudaGetLastError());
    return SNN_OK;
}

int snn_rpn_topk_select(const void* const* logits, const int* H, const int* W, int n_levels, int N, int A, int k,
                        long long* idx_out, void* workspace, size_t workspace_bytes, snn_stream_t stream) {
    if (!logits || !H || !W || !idx_out || !workspace) return fail(SNN_E_ARG, "rpn_topk_select: null argument");
    if (n_levels < 1 || n_levels > kPropMaxLevels || N < 1 || A < 1 || k < 1 || k > kTopkMaxK)
        return fail(SNN_E_ARG, "rpn_topk_select: bad sizes (levels %d, N %d, A %d, k %d; k <= %d)", n_levels, N, A, k, kTopkMaxK);
    if (static_cast<long long>(n_levels) * N > 65535) return fail(SNN_E_ARG, "rpn_topk_select: too many (level, image) segments");
    const TopkWs w = topk_ws_layout(n_levels, N);
    if (workspace_bytes < w.total) return fail(SNN_E_WORKSPACE, "workspace %zu B < required %zu B", workspace_bytes, w.total);
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return fail(SNN_E_ARG, "workspace must be 256-B aligned");
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    TopkParams p;
    memset(&p, 0, sizeof(p));
    int kb = 0, max_chunks = 0;
    for (int l = 0; l < n_levels; ++l) {
        if (!logits[l] || H[l] < 1 || W[l] < 1) return fail(SNN_E_ARG, "rpn_topk_select: level %d: bad argument", l);
        TopkLevel& L = p.lv[l];
        const long long n = static_cast<long long>(A) * H[l] * W[l];
        if (n > 0x7FFFFFFFll) return fail(SNN_E_ARG, "rpn_topk_select: level %d too large", l);
        L.logits = reinterpret_cast<const float*>(logits[l]); L.A = A; L.HW = H[l] * W[l]; L.n = static_cast<int>(n);
        L.k = k < L.n ? k : L.n; L.k_begin = kb; L.chunks = (L.n + kTopkChunk - 1) / kTopkChunk;
        kb += L.k;
        if (L.chunks > max_chunks) max_chunks = L.chunks;
    }
    p.n_levels = n_levels; p.N = N; p.K_total = kb;
    p.prefix = reinterpret_cast<unsigned long long*>(ws + w.prefix); p.k_rem = reinterpret_cast<unsigned int*>(ws + w.k_rem);
    p.done = reinterpret_cast<unsigned int*>(ws + w.done); p.hist = reinterpret_cast<unsigned int*>(ws + w.hist);
    p.n_cand = reinterpret_cast<unsigned int*>(ws + w.n_cand); p.cand = reinterpret_cast<unsigned long long*>(ws + w.cand);
    p.idx = idx_out;
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(ws + w.zero_begin, 0, w.zero_end - w.zero_begin, st));
    const dim3 grid(max_chunks, n_levels * N);
    rpn_topk_pass_kernel<0><<<grid, kTopkThreads, 0, st>>>(p);
    rpn_topk_pass_kernel<1><<<grid, kTopkThreads, 0, st>>>(p);
    rpn_topk_pass_kernel<2><<<grid, kTopkThreads, 0, st>>>(p);
    rpn_topk_pass_kernel<3><<<grid, kTopkThreads, 0, st>>>(p);
    rpn_topk_pass_kernel<4><<<grid, kTopkThreads, 0, st>>>(p);
    rpn_topk_pass_kernel<5><<<grid, kTopkThreads, 0, st>>>(p);
    rpn_topk_collect_kernel<<<grid, kTopkThreads, 0, st>>>(p);
    rpn_topk_sort_kernel<<<n_levels * N, 1024, 0, st>>>(p);
    CUDA_TRY(cudaGetLastError());
    return SNN_OK;
}

// ---- linear statistics of spike trains (the spike-rate / energy report, rates.py)
int snn_li_readout_nhwc(const void* trains, int train_bytes, int N, int HW, int C, const double* step_weights,
                        const float* w_a, int n_a, const float* w_b, int n_b, float* out_a, float* out_b,
                        snn_stream_t stream) {
    if (!trains || !step_weights || !w_a || !w_b || !out_a || !out_b) return fail(SNN_E_ARG, "li_readout_nhwc: null argument");
    if (N < 1 || HW < 1 || C < 1 || n_a < 1 || n_b < 1 || (train_bytes != 1 && train_bytes != 2 && train_bytes != 4) ||
        (C * train_bytes) % 4 != 0)
        return fail(SNN_E_ARG, "li_readout_nhwc: unsupported sizes (N %d, HW %d, C %d, outputs %d + %d, %d-byte words)", N, HW, C,
                    n_a, n_b, train_bytes);
    KappaTable kt;
    for (int t = 0; t < 32; ++t) kt.k[t] = (t < 8 * train_bytes) ? step_weights[t] : 0.0;
    cudaError_t e;
    cudaStream_t st = (cudaStream_t)stream;
    if (train_bytes == 1) e = launch_readout_rpn<uint8_t>(trains, C, HW, N, w_a, n_a, w_b, n_b, kt, out_a, out_b, nullptr, st);
    else if (train_bytes == 2) e = launch_readout_rpn<uint16_t>(trains, C, HW, N, w_a, n_a, w_b, n_b, kt, out_a, out_b, nullptr, st);
    else e = launch_readout_rpn<uint32_t>(trains, C, HW, N, w_a, n_a, w_b, n_b, kt, out_a, out_b, nullptr, st);
    if (e != cudaSuccess) return fail(SNN_E_CUDA, "li_readout_nhwc launch failed: %s", cudaGetErrorString(e));
    return SNN_OK;
}

int snn_li_readout_rows(const void* trains, int train_bytes, int R, int Hd, const double* step_weights, const float* w_a,
                        int n_a, const float* w_b, int n_b, float* out_a, float* out_b, snn_stream_t stream) {
    if (!trains || !step_weights || !w_a || !w_b || !out_a || !out_b) return fail(SNN_E_ARG, "li_readout_rows: null argument");
    if (R < 1 || Hd < 1 || n_a < 1 || n_b < 1 || (train_bytes != 1 && train_bytes != 2 && train_bytes != 4) ||
        (Hd * train_bytes) % 16 != 0 || Hd % 4 != 0)
        return fail(SNN_E_ARG, "li_readout_rows: unsupported sizes (R %d, Hd %d, outputs %d + %d, %d-byte words)", R, Hd, n_a, n_b,
                    train_bytes);
    KappaTable kt;
    for (int t = 0; t < 32; ++t) kt.k[t] = (t < 8 * train_bytes) ? step_weights[t] : 0.0;
    cudaError_t e;
    cudaStream_t st = (cudaStream_t)stream;
    if (train_bytes == 1) e = launch_readout_rows<uint8_t>(trains, nullptr, R, Hd, w_a, n_a, w_b, n_b, kt, out_a, out_b, nullptr, st);
    else if (train_bytes == 2) e = launch_readout_rows<uint16_t>(trains, nullptr, R, Hd, w_a, n_a, w_b, n_b, kt, out_a, out_b, nullptr, st);
    else e = launch_readout_rows<uint32_t>(trains, nullptr, R, Hd, w_a, n_a, w_b, n_b, kt, out_a, out_b, nullptr, st);
    if (e != cudaSuccess) return fail(SNN_E_CUDA, "li_readout_rows launch failed: %s", cudaGetErrorString(e));
    return SNN_OK;
}

void snn_profile_enable(int on) {
    g_profile = on == 1 ? 0xFFFFFFFFu : on <= 0 ? 0u : (static_cast<unsigned>(on) >> 1);
    for (int k = 0; k < PH_COUNT; ++k) g_ph[k].used = 0;
}

int snn_profile_read(float* ms_out, int* counts_out) {
    for (int k = 0; k < PH_COUNT; ++k) {
        float tot = 0.f;
        PhaseEvents& e = g_ph[k];
        for (int i = 0; i < e.used; ++i) {
            CUDA_TRY(cudaEventSynchronize(e.stop[i]));
            float ms = 0.f;
            CUDA_TRY(cudaEventElapsedTime(&ms, e.start[i], e.stop[i]));
            tot += ms;
        }
        if (ms_out) ms_out[k] = tot;
        if (counts_out) counts_out[k] = e.used;
        e.used = 0;
    }
    return SNN_OK;
}

int snn_encode_rows(const float* x, int R, int K, int T_live, void* z_words, snn_stream_t stream) {
    if (!x || !z_words || R < 1 || K % 16 != 0 || T_live < 1 || T_live > 32)
        return fail(SNN_E_ARG, "encode_rows: bad argument");
    DeviceInfo di;
    int rc = device_info(di);
    if (rc) return rc;
    CUDA_TRY(launch_encode_rows(x, static_cast<size_t>(R) * K, T_live, word_bytes(T_live),
                                reinterpret_cast<uint8_t*>(z_words), di.sms, (cudaStream_t)stream));
    return SNN_OK;
}

void snn_encoder_table(float* thresholds33, unsigned int* deltas33) {
    for (int n = 0; n <= 32; ++n) {
        if (thresholds33) thresholds33[n] = kEncTableHost.thr[n];
        if (deltas33) deltas33[n] = kEncTableHost.delta[n];
    }
}

int snn_encoder_selftest(int T_live, unsigned long long* mismatches, snn_stream_t stream) {
    if (!mismatches || T_live < 1 || T_live > 32) return fail(SNN_E_ARG, "encoder_selftest: bad argument");
    SNN_ENC_BUCKETS(T_live, (encoder_selftest_kernel<NT><<<148 * 16, 256, 0, (cudaStream_t)stream>>>(T_live, mismatches)));
    CUDA_TRY(cudaGetLastError());
    return SNN_OK;
}

int snn_fc_lif_layer(const void* z_words, int in_word_bytes, int in_bit0, int R, int K, int M, int T, int t0,
                     int T_live, int mode, const void* w_prep, void* trains, float* dump, int cta_group,
                     snn_stream_t stream) {
    g_launches = 0;
    if (!z_words || !w_prep || !trains) return fail(SNN_E_ARG, "fc_lif_layer: null argument");
    if (K % 64 != 0 || M % 128 != 0 || R < 1 || T < 1 || T > 32 || T_live < 1 || t0 < 0 || t0 + T_live > T)
        return fail(SNN_E_ARG, "fc_lif_layer: unsupported shape R=%d K=%d M=%d T=%d t0=%d T_live=%d", R, K, M, T, t0, T_live);
    if ((in_word_bytes != 1 && in_word_bytes != 2 && in_word_bytes != 4) || in_bit0 < 0 ||
        in_bit0 + T_live > 8 * in_word_bytes)
        return fail(SNN_E_ARG, "fc_lif_layer: %d steps from bit %d do not fit %d-byte words", T_live, in_bit0, in_word_bytes);
    if (nsplit_of(mode) == 0) return fail(SNN_E_ARG, "unknown mode %d", mode);
    DeviceInfo di;
    int rc = device_info(di);
    if (rc) return rc;
    TileCfg tc;
    if (!pick_tile(T_live, false, M, cta_group, tc)) return fail(SNN_E_ARG, "no tile shape for T_live=%d cta_group=%d", T_live, cta_group);
    return fc_layer(di, z_words, in_word_bytes, in_bit0, R, K, M, T, t0, T_live, mode, w_prep, trains, dump, tc,
                    (cudaStream_t)stream);
}

}  // extern "C"
