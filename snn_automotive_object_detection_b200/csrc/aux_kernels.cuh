// HBM-bound companions of spike_gemm_lif: weight preparation, the constant-current
// LIF encoder (Norse lif_current_encoder; rpn.py:101, faster_rcnn.py:494) and the
// leaky-integrator readouts (Norse LICell; rpn.py:110-115, faster_rcnn.py:505-510).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>

#include "ptx.cuh"

namespace snn {

// ------------------------------------------------------------ weight prep
// One block per output row.  bf16 modes: w = hi + mid + lo, each piece a bf16 (residuals are exact in
// fp32), scale = 1.  fp16 modes: the row is first multiplied by 2^S with max|row| * 2^S in [2^14, 2^15)
// (exact; keeps every low piece of a weight down to 2^-16 of the row maximum out of the fp16 subnormals),
// then split into fp16 pieces; the epilogue multiplies the accumulator by scale = 2^-S (exact).
// Output layout: [nsplit][O][K] 16-bit pieces, then float scale[O].
// kConv: source is [O][C][3][3] and the prepared k index is (ky*3+kx)*C + c; else source is [O][K].
template <bool kConv>
__global__ void __launch_bounds__(256) prep_weights_kernel(const float* __restrict__ w, int O, int K, int C,
                                                           int nsplit, int fp16, uint16_t* __restrict__ out,
                                                           float* __restrict__ scale) {
    __shared__ float s_max[8];
    const size_t plane = static_cast<size_t>(O) * K;
    for (int o = blockIdx.x; o < O; o += gridDim.x) {
        const float* wrow = w + static_cast<size_t>(o) * K;
        int S = 0;
        if (fp16) {
            float m = 0.f;
            for (int k = threadIdx.x; k < K; k += blockDim.x) m = fmaxf(m, fabsf(wrow[k]));
            for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
            __syncthreads();
            if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
            __syncthreads();
            m = s_max[0];
            for (int q = 1; q < 8; ++q) m = fmaxf(m, s_max[q]);
            if (m > 0.f && m < 3.0e38f) {
                int e;
                frexpf(m, &e);                     // m = f * 2^e, f in [0.5, 1)
                S = 15 - e;
                S = S > 100 ? 100 : (S < -100 ? -100 : S);
            }
        }
        if (threadIdx.x == 0) scale[o] = ldexpf(1.0f, -S);
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            float src;
            if (kConv) {
                const int tap = k / C, c = k - tap * C;
                src = wrow[c * 9 + tap];
            } else {
                src = wrow[k];
            }
            float r = ldexpf(src, S);
            const size_t idx = static_cast<size_t>(o) * K + k;
            for (int s = 0; s < nsplit; ++s) {
                if (fp16) {
                    const __half h = __float2half_rn(r);
                    out[s * plane + idx] = __half_as_ushort(h);
                    r = __fsub_rn(r, __half2float(h));
                } else {
                    const __nv_bfloat16 b = __float2bfloat16_rn(r);
                    out[s * plane + idx] = __bfloat16_as_ushort(b);
                    r = __fsub_rn(r, __bfloat162float(b));
                }
            }
        }
    }
}

// ---------------------------------------------------------------- encoder
// Norse lif_current_encoder, op for op: v += 0.1f*((0-v)+x); z = (v-0.25f > 0); v -= z*v.
// Returns the spike train of the first T steps as a bit word (bit t = z_t).  (x = +inf spikes once: the arithmetic reset
// inf - inf leaves NaN, which never crosses the threshold again -- reproduced, see encode_word.)
// The threshold test is evaluated as v > 0.25f: for fp32 numbers a and b, fl(a - b) > 0 <=> a > b (the
// difference of two floats near the threshold is a multiple of 2^-26, never flushed), so it is the same bit.
__device__ __forceinline__ uint32_t encode_train(float x, int T) {
    float v = 0.f;
    uint32_t w = 0u;
    for (int t = 0; t < T; ++t) {
        v = __fadd_rn(v, __fmul_rn(0.1f, __fsub_rn(x, v)));
        const bool z = v > 0.25f;
        if (z) w |= 1u << t;
        v = z ? __fsub_rn(v, v) : v;          // Norse resets arithmetically, v - z * (v - v_reset): 0 for finite v, NaN for +inf
    }
    return w;
}

// ---- the encoder as a comparator bank
// The input current x is constant over the steps and a spike resets v to exactly 0.f = the initial state, so the
// trajectory repeats: a neuron whose first spike falls on step n (1-based) fires at steps n, 2n, 3n, ... -- its whole
// train is a function of n(x) alone.  n(x) is non-increasing in x (every op of the update is monotone in x while no
// spike has occurred; checked for EVERY fp32 input by snn_encoder_selftest), so
//     n(x) = min { n : x >= thr[n] },   thr[n] = the smallest fp32 x whose first spike comes at step <= n
// and the train word is a telescoping XOR over the thresholds passed:
//     word = XOR_{n : x >= thr[n]} (full[n] ^ full[n+1]),   full[n] = bits n-1, 2n-1, 3n-1, ... of a 32-step train
// (x >= thr[n] implies x >= thr[m] for m > n, so the XOR collapses to full[n(x)]); the caller masks the live steps.
// 2 instructions per step and neuron instead of the 6 of the simulation.  The thresholds are found at COMPILE time by
// bisection on the simulation itself (constexpr fp32 arithmetic: one IEEE add/sub/mul per op, the same roundings as
// __fadd_rn/__fmul_rn); NaN compares false everywhere = never spikes, as simulated.
// Pipes: FSETP and LOP3 both issue to the ALU pipe (2 cycles per warp instruction), which bounded the first version
// (ncu r01ab: ALU 73 % busy, FMA 20 %).  For words of <= 16 bits the telescoping sum is therefore taken in fp32 on the
// FMA pipe: acc = 2^23 + SUM_{n : x >= thr[n]} (full16[n] - full16[n+1]) -- small integers, exact in fp32 -- and the
// word is the low mantissa bits of acc: one FSETP (ALU) + one predicated FADD (FMA) per step, the pipes in parallel.
struct EncTable {
    float thr[33]; uint32_t delta[33];      // index n = 1..32; [0] unused
    float deltaf[17], lastf[17];            // fp32 path (n <= 16): full16[n] - full16[n+1], and full16[n] for the bucket's last n
};

__host__ __device__ constexpr int enc_first_spike(float x, int nmax) {
    float v = 0.f;
    for (int n = 1; n <= nmax; ++n) {
        const float d = x - v;
        const float dv = 0.1f * d;
        v = v + dv;
        if (v > 0.25f) return n;
    }
    return nmax + 1;
}
__host__ __device__ constexpr uint32_t enc_full_train(int n) {
    uint32_t w = 0u;
    for (int t = n - 1; t < 32; t += n) w |= 1u << t;
    return n <= 32 ? w : 0u;
}
__host__ __device__ constexpr EncTable make_enc_table() {
    EncTable tb{};
    for (int n = 1; n <= 32; ++n) {
        float lo = 0.25f, hi = 4.0f;            // first spike of lo comes after step n (never), of hi at step 1
        for (int it = 0; it < 64; ++it) {
            const float mid = 0.5f * (lo + hi);
            if (!(mid > lo && mid < hi)) break; // lo and hi are adjacent floats
            if (enc_first_spike(mid, n) <= n) hi = mid; else lo = mid;
        }
        tb.thr[n] = hi;
        tb.delta[n] = enc_full_train(n) ^ enc_full_train(n + 1);
    }
    tb.thr[0] = 0.f; tb.delta[0] = 0u;
    for (int n = 1; n <= 16; ++n) {
        const int a = static_cast<int>(enc_full_train(n) & 0xFFFFu), b = static_cast<int>(enc_full_train(n + 1) & 0xFFFFu);
        tb.deltaf[n] = static_cast<float>(a - b);
        tb.lastf[n] = static_cast<float>(a);
    }
    tb.deltaf[0] = 0.f; tb.lastf[0] = 0.f;
    return tb;
}
constexpr EncTable kEncTableHost = make_enc_table();
__constant__ EncTable c_enc = make_enc_table();

// if (x >= th) w ^= delta: one FSETP + one predicated LOP3 (both ALU pipe)
__device__ __forceinline__ void xor_if_ge(uint32_t& w, float x, float th, uint32_t delta) {
    asm("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %1, %2;\n\t@p xor.b32 %0, %0, %3;\n\t}" : "+r"(w) : "f"(x), "f"(th), "r"(delta));
}
// if (x >= th) acc += d: one FSETP (ALU pipe) + one predicated FADD (FMA pipe)
__device__ __forceinline__ void add_if_ge(float& acc, float x, float th, float d) {
    asm("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %1, %2;\n\t@p add.rn.f32 %0, %0, %3;\n\t}" : "+f"(acc) : "f"(x), "f"(th), "f"(d));
}

// N inputs at once (independent chains for the scheduler).  NT >= the number of live steps (a compile-time bucket);
// tmask = the live steps (bits at steps >= T_live come out of the wider bucket and are masked).
template <int NT, int N>
__device__ __forceinline__ void encode_words(const float (&x)[N], uint32_t tmask, uint32_t (&w)[N]) {
    if constexpr (NT <= 16) {
        float acc[N];
#pragma unroll
        for (int k = 0; k < N; ++k) acc[k] = 8388608.f;            // 2^23: integers up to 2^23 on top of it sit in the mantissa
#pragma unroll
        for (int n = 1; n <= NT; ++n) {
            const float th = c_enc.thr[n];
            const float d = (n == NT) ? c_enc.lastf[n] : c_enc.deltaf[n];
#pragma unroll
            for (int k = 0; k < N; ++k) add_if_ge(acc[k], x[k], th, d);
        }
#pragma unroll
        for (int k = 0; k < N; ++k) w[k] = __float_as_uint(acc[k]) & tmask;
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) w[k] = 0u;
#pragma unroll
        for (int n = 1; n <= NT; ++n) {
            const float th = c_enc.thr[n];
            const uint32_t dl = c_enc.delta[n] & tmask;
#pragma unroll
            for (int k = 0; k < N; ++k) xor_if_ge(w[k], x[k], th, dl);
        }
    }
    // the one input whose train is not periodic: +inf spikes at step 0 and leaves v = inf - inf = NaN (Norse's
    // arithmetic reset), which never spikes again
#pragma unroll
    for (int k = 0; k < N; ++k)
        if (x[k] == __int_as_float(0x7f800000)) w[k] = 1u & tmask;
}
template <int NT>
__device__ __forceinline__ uint32_t encode_word(float x, uint32_t tmask) {
    const float xs[1] = {x};
    uint32_t w[1];
    encode_words<NT, 1>(xs, tmask, w);
    return w[0];
}

// Tried in round 2 and dropped (profiles/r02/r02k_enc_lut_experiment.txt): the encoder as a table lookup.  ncu r02c shows
// both encoder kernels instruction-bound rather than HBM-bound (issue slots 73-75 % busy, ALU pipe 65-81 %, DRAM 4.1-4.2
// of 6.5 TB/s), so a lookup keyed by the fp16 image of the clamped input (4097 values; the rounding interval of one
// fp16 value contains at most one threshold, so one byte lookup + one 16-byte lookup + one exact compare finish, bit-exact
// for all 2^32 inputs) looked attractive at 11.5 instructions per neuron against 2 per step.  Measured: the box encoder
// (11 steps) went from 0.034 to 0.051 ms, a variant with flagged entries and an out-of-line exact path from 0.034 to
// 0.041 ms and the RPN encoder from 0.063 to 0.070 ms -- 32 lanes reading random shared-memory addresses cost more issue
// and wavefront cycles than 22 register-only instructions.  The comparator bank stays.

// dispatch a kernel template on the bucket of live steps (exact for the step counts of the reference's sweeps)
#define SNN_ENC_BUCKETS(T_live, ...)                                            \
    do {                                                                        \
        if ((T_live) <= 4) { constexpr int NT = 4; __VA_ARGS__; }               \
        else if ((T_live) <= 7) { constexpr int NT = 7; __VA_ARGS__; }          \
        else if ((T_live) <= 8) { constexpr int NT = 8; __VA_ARGS__; }          \
        else if ((T_live) <= 10) { constexpr int NT = 10; __VA_ARGS__; }        \
        else if ((T_live) <= 11) { constexpr int NT = 11; __VA_ARGS__; }        \
        else if ((T_live) <= 12) { constexpr int NT = 12; __VA_ARGS__; }        \
        else if ((T_live) <= 16) { constexpr int NT = 16; __VA_ARGS__; }        \
        else if ((T_live) <= 24) { constexpr int NT = 24; __VA_ARGS__; }        \
        else { constexpr int NT = 32; __VA_ARGS__; }                            \
    } while (0)

constexpr int kEncPx = 32;       // pixels per work item (one per lane)
constexpr int kEncCh = 32;       // channels per work item (all held by one thread: a full 32-byte sector of 1-byte words)
constexpr int kEncMaxLevels = 8;

struct EncLevel {
    const float* x;             // [N][C][H*W] fp32 (contiguous NCHW)
    uint8_t* z;                 // [N][H*W][C] spike-train words of `wb` bytes (bit t = z_t, t < T_live)
    int HW, chunks, chunk_begin, pad_;   // chunks of 32 consecutive pixels per image (the last one ragged)
};
struct EncParams {
    EncLevel lv[kEncMaxLevels];
    int n_levels, N, C, T_live, wb, total_items;
};

template <int WB>
__device__ __forceinline__ void store_words16(uint8_t* dst, const uint32_t (&w)[16]) {
    uint4* d = reinterpret_cast<uint4*>(dst);
    if constexpr (WB == 1) {
        d[0] = make_uint4(w[0] | (w[1] << 8) | (w[2] << 16) | (w[3] << 24), w[4] | (w[5] << 8) | (w[6] << 16) | (w[7] << 24),
                          w[8] | (w[9] << 8) | (w[10] << 16) | (w[11] << 24), w[12] | (w[13] << 8) | (w[14] << 16) | (w[15] << 24));
    } else if constexpr (WB == 2) {
        d[0] = make_uint4(w[0] | (w[1] << 16), w[2] | (w[3] << 16), w[4] | (w[5] << 16), w[6] | (w[7] << 16));
        d[1] = make_uint4(w[8] | (w[9] << 16), w[10] | (w[11] << 16), w[12] | (w[13] << 16), w[14] | (w[15] << 16));
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) d[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
    }
}

// All FPN levels in one launch, no shared memory.  A level's channel planes and its NHWC output are both contiguous
// in the flat pixel index, so the work item of a WARP is (32 consecutive flat pixels, 32 consecutive channels):
//   lane = pixel; its 32 loads (one per channel plane, each a coalesced 128-byte row across the warp) are all issued
//   before the comparator bank runs, and its 32 words are 32 / 64 / 128 contiguous bytes of the NHWC output -- whole
//   32-byte sectors, written with 16-byte stores.  Consecutive warps take the channel groups of the same pixels, so a
//   block completes whole pixel rows.  Grid-stride over the items of all levels and images.
// HBM traffic: 4 B read + wb B written per input neuron (the per-timestep planes never exist).
template <int NT, int WB>
__global__ void __launch_bounds__(256) encode_nchw_kernel(const __grid_constant__ EncParams p) {
    griddep_launch_dependents();
    griddep_wait();             // the words buffer may still be read by the previous forward's GEMM
    const int lane = threadIdx.x & 31;
    const int n_warps = gridDim.x * (blockDim.x >> 5);
    const int cgroups = p.C / kEncCh;
    const uint32_t tmask = (p.T_live >= 32) ? 0xFFFFFFFFu : ((1u << p.T_live) - 1u);
    for (int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); item < p.total_items; item += n_warps) {
        const int chunk = item / cgroups, c0 = (item - chunk * cgroups) * kEncCh;
        int lvl = 0;
        while (lvl + 1 < p.n_levels && chunk >= p.lv[lvl + 1].chunk_begin) ++lvl;
        const EncLevel& L = p.lv[lvl];
        int local = chunk - L.chunk_begin;
        const int n = local / L.chunks;
        local -= n * L.chunks;
        const int px = local * kEncPx + lane;
        const bool ok = px < L.HW;
        const size_t hw = static_cast<size_t>(L.HW);
        // byte pointer of channel c0, then one 32 x 32 -> 64-bit multiply-add per channel plane
        const char* src = reinterpret_cast<const char*>(L.x + (static_cast<size_t>(n) * p.C + c0) * hw + px);
        const uint32_t plane = static_cast<uint32_t>(L.HW) * 4u;
        float xv[kEncCh];
#pragma unroll
        for (int k = 0; k < kEncCh; ++k)
            xv[k] = ok ? __ldg(reinterpret_cast<const float*>(src + static_cast<uint64_t>(plane) * static_cast<uint32_t>(k))) : 0.f;
        uint8_t* dst = L.z + ((static_cast<size_t>(n) * hw + px) * p.C + c0) * WB;
#pragma unroll
        for (int h = 0; h < kEncCh / 16; ++h) {
            float xs[16];
            uint32_t w[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) xs[k] = xv[16 * h + k];
            encode_words<NT, 16>(xs, tmask, w);
            if (ok) store_words16<WB>(dst + 16 * h * WB, w);
        }
    }
}

// x [R][K] fp32 -> words [R][K] of WB bytes; 16 consecutive k per thread and iteration (4 x float4 in flight)
template <int NT, int WB>
__global__ void __launch_bounds__(256) encode_rows_kernel(const float* __restrict__ x, size_t total16, int T_live,
                                                          uint8_t* __restrict__ z) {
    griddep_launch_dependents();
    griddep_wait();
    const uint32_t tmask = (T_live >= 32) ? 0xFFFFFFFFu : ((1u << T_live) - 1u);
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total16;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float4* src = reinterpret_cast<const float4*>(x) + 4 * i;
        const float4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2), d = __ldg(src + 3);
        const float xs[16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
        uint32_t w[16];
        encode_words<NT, 16>(xs, tmask, w);
        store_words16<WB>(z + i * 16 * WB, w);
    }
}

// ---------------------------------------------------------------- readout
// LICell is linear with zero initial state, so its last-step membrane is
//   mem_{T-1} = W . s,   s[c] = sum_t kappa_{T-1-t} spk_t[c],  kappa_n = 0.9^{n+1} - 0.8^{n+1}
// s[c] is a function of the neuron's spike-train word -> byte-indexed lookup tables
// lut[b][256] (b = byte position in the word), built on the host in float64.
constexpr int kMaxTrainBytes = 4;

// kappa[t] = 0.9^{T-t} - 0.8^{T-t} in float64 (computed on the host), passed by value; each block builds its own
// byte-indexed tables from it: lut[b][v] = float( sum_{bit j of v} kappa[8b + j] )  (summed in float64)
struct KappaTable { double k[32]; };
__device__ __forceinline__ void build_kappa_lut(const KappaTable& kt, int nbytes, float* s_lut) {
    for (int i = threadIdx.x; i < nbytes * 256; i += blockDim.x) {
        const int b = i >> 8, v = i & 255;
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if ((v >> j) & 1) acc += kt.k[8 * b + j];
        s_lut[i] = static_cast<float>(acc);
    }
}

template <typename TrainT>
__device__ __forceinline__ float lut_weight(const float* lut, TrainT tr) {
    float s = lut[tr & 0xFFu];
    if constexpr (sizeof(TrainT) >= 2) s += lut[256 + ((tr >> 8) & 0xFFu)];
    if constexpr (sizeof(TrainT) == 4) { s += lut[512 + ((tr >> 16) & 0xFFu)]; s += lut[768 + ((tr >> 24) & 0xFFu)]; }
    return s;
}

constexpr int kRpnRoPx = 128;      // pixels per block
constexpr int kRpnMaxOut = 16;     // A + 4A <= 16 per pass (A = 3 -> 15)

// trains [N][H*W][C] -> logits [N][n_a][H][W], bbox [N][n_b][H][W] (the RPN head: n_a = A, n_b = 4A); one thread per
// pixel.  Also counts spikes per image (popc of the train words) into counts[n].  `kt` is the weight of a spike at
// step t: kappa_{T-1-t} for the last LI membrane, other tables for other linear statistics of the trains (rates.py).
template <typename TrainT>
__global__ void __launch_bounds__(kRpnRoPx) readout_rpn_kernel(const TrainT* __restrict__ trains, int C, int HW,
                                                               const float* __restrict__ w_cls,
                                                               const float* __restrict__ w_bbox, int n_a, int n_b,
                                                               const __grid_constant__ KappaTable kt,
                                                               float* __restrict__ logits, float* __restrict__ bbox,
                                                               unsigned long long* __restrict__ counts) {
    extern __shared__ uint8_t s_raw[];
    const int n_out = n_a + n_b;
    const int row_words = (C * static_cast<int>(sizeof(TrainT))) / 4 + 1;   // +1 word: conflict-free row stride
    float* s_w = reinterpret_cast<float*>(s_raw);                  // [n_out][C]
    float* s_lut = s_w + n_out * C;                                // [sizeof(TrainT)][256]
    uint32_t* s_tr = reinterpret_cast<uint32_t*>(s_lut + 256 * sizeof(TrainT));   // [kRpnRoPx][row_words]
    __shared__ unsigned int s_cnt;

    const int n = blockIdx.y;
    const int p0 = blockIdx.x * kRpnRoPx;
    const int npx = min(kRpnRoPx, HW - p0);
    if (threadIdx.x == 0) s_cnt = 0;
    for (int i = threadIdx.x; i < n_out * C; i += blockDim.x)
        s_w[i] = (i < n_a * C) ? w_cls[i] : w_bbox[i - n_a * C];
    build_kappa_lut(kt, static_cast<int>(sizeof(TrainT)), s_lut);
    griddep_wait();             // the spike trains of the preceding GEMM (no-op without a programmatic launch)
    // coalesced tile load: npx rows of C*sizeof(TrainT) bytes are contiguous in global memory
    const uint32_t* src = reinterpret_cast<const uint32_t*>(trains + (static_cast<size_t>(n) * HW + p0) * C);
    const int wpr = row_words - 1;
    unsigned int cnt = 0;
    for (int i = threadIdx.x; i < npx * wpr; i += blockDim.x) {
        const int r = i / wpr, k = i - r * wpr;
        const uint32_t wv = src[i];
        cnt += __popc(wv);
        s_tr[r * row_words + k] = wv;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);

    const int px = threadIdx.x;
    if (px < npx) {
        constexpr int per_word = 4 / static_cast<int>(sizeof(TrainT));
        const size_t pix = static_cast<size_t>(p0 + px);
        for (int o0 = 0; o0 < n_out; o0 += kRpnMaxOut) {       // A = 3 -> a single pass of 15 outputs
            float acc[kRpnMaxOut];
#pragma unroll
            for (int o = 0; o < kRpnMaxOut; ++o) acc[o] = 0.f;
            for (int k = 0; k < wpr; ++k) {
                const uint32_t wv = s_tr[px * row_words + k];
#pragma unroll
                for (int e = 0; e < per_word; ++e) {
                    const TrainT tr = static_cast<TrainT>(wv >> (8 * static_cast<int>(sizeof(TrainT)) * e));
                    const float s = lut_weight<TrainT>(s_lut, tr);
                    const int c = k * per_word + e;
#pragma unroll
                    for (int o = 0; o < kRpnMaxOut; ++o)
                        if (o0 + o < n_out) acc[o] = fmaf(s_w[(o0 + o) * C + c], s, acc[o]);
                }
            }
#pragma unroll
            for (int o = 0; o < kRpnMaxOut; ++o) {
                const int oo = o0 + o;
                if (oo < n_a) logits[(static_cast<size_t>(n) * n_a + oo) * HW + pix] = acc[o];
                else if (oo < n_out) bbox[(static_cast<size_t>(n) * n_b + (oo - n_a)) * HW + pix] = acc[o];
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && counts != nullptr && s_cnt) atomicAdd(&counts[n], static_cast<unsigned long long>(s_cnt));
}

constexpr int kRoRows = 8;         // RoIs per block

// trains [R][Hd] -> out_cls [R][n_cls], out_box [R][n_box].  One block = 8 RoIs: their kappa-weighted
// spike sums s[8][Hd] are staged in shared memory (LUT per spike-train byte); each warp then owns
// every 8th output row of [w_cls; w_box], streams it once (coalesced) against all 8 RoIs and finishes
// with a warp-shuffle reduction.  Also per-RoI spike counts of this layer and of an optional
// second layer `trains_b` (fc6): counts[0][R] = layer b, counts[1][R] = this layer.
// The kernel is latency-bound (250 blocks, a few KB each), so every global read is a 16-byte vector and all the
// vectors of a row are requested before the first is used: one memory latency per row instead of one per 32 words.
template <typename TrainT>
__global__ void __launch_bounds__(256) readout_rows_kernel(const TrainT* __restrict__ trains,
                                                           const TrainT* __restrict__ trains_b, int R, int Hd,
                                                           const float* __restrict__ w_cls, int n_cls,
                                                           const float* __restrict__ w_box, int n_box,
                                                           const __grid_constant__ KappaTable kt,
                                                           float* __restrict__ out_cls, float* __restrict__ out_box,
                                                           unsigned int* __restrict__ counts) {
    __shared__ float s_lut[256 * sizeof(TrainT)];
    extern __shared__ float s_s[];                 // [kRoRows][Hd]
    griddep_launch_dependents();
    constexpr int kWpv = 16 / static_cast<int>(sizeof(TrainT));      // words per 16-byte vector
    constexpr int kMaxVec = 8;                                       // vectors per lane and pass (Hd <= 8 * 32 * kWpv per pass)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = blockIdx.x * kRoRows;
    build_kappa_lut(kt, static_cast<int>(sizeof(TrainT)), s_lut);
    griddep_wait();             // the spike trains of the preceding GEMM
    __syncthreads();
    {   // warp q stages row r0 + q (8 warps <-> 8 rows) and counts its spikes
        const int r = r0 + warp;
        unsigned int ca = 0, cb = 0;
        const int n_vec = Hd / kWpv;
        for (int v0 = 0; v0 < n_vec; v0 += 32 * kMaxVec) {
            uint4 va[kMaxVec], vb[kMaxVec];
#pragma unroll
            for (int i = 0; i < kMaxVec; ++i) {
                const int v = v0 + lane + 32 * i;
                va[i] = make_uint4(0u, 0u, 0u, 0u); vb[i] = va[i];
                if (r < R && v < n_vec) {
                    va[i] = __ldg(reinterpret_cast<const uint4*>(trains + static_cast<size_t>(r) * Hd) + v);
                    if (trains_b != nullptr) vb[i] = __ldg(reinterpret_cast<const uint4*>(trains_b + static_cast<size_t>(r) * Hd) + v);
                }
            }
#pragma unroll
            for (int i = 0; i < kMaxVec; ++i) {
                const int v = v0 + lane + 32 * i;
                if (v >= n_vec) break;
                ca += __popc(va[i].x) + __popc(va[i].y) + __popc(va[i].z) + __popc(va[i].w);
                cb += __popc(vb[i].x) + __popc(vb[i].y) + __popc(vb[i].z) + __popc(vb[i].w);
                const uint32_t w32[4] = {va[i].x, va[i].y, va[i].z, va[i].w};
                float* dst = &s_s[warp * Hd + v * kWpv];
#pragma unroll
                for (int e = 0; e < kWpv; ++e) {
                    constexpr int per32 = 4 / static_cast<int>(sizeof(TrainT));
                    const TrainT tr = static_cast<TrainT>(w32[e / per32] >> (8 * static_cast<int>(sizeof(TrainT)) * (e % per32)));
                    dst[e] = lut_weight<TrainT>(s_lut, tr);
                }
            }
        }
        if (counts != nullptr) {
            for (int off = 16; off > 0; off >>= 1) {
                ca += __shfl_xor_sync(0xffffffffu, ca, off);
                cb += __shfl_xor_sync(0xffffffffu, cb, off);
            }
            if (lane == 0 && r < R) { counts[r] = cb; counts[R + r] = ca; }
        }
    }
    __syncthreads();
    const int n_out = n_cls + n_box;
    const int n_v4 = Hd / 4;
    for (int o = warp; o < n_out; o += 8) {
        const float4* wrow = reinterpret_cast<const float4*>((o < n_cls) ? (w_cls + static_cast<size_t>(o) * Hd)
                                                                         : (w_box + static_cast<size_t>(o - n_cls) * Hd));
        float acc[kRoRows];
#pragma unroll
        for (int q = 0; q < kRoRows; ++q) acc[q] = 0.f;
        for (int v0 = 0; v0 < n_v4; v0 += 32 * kMaxVec) {
            float4 wv[kMaxVec];
#pragma unroll
            for (int i = 0; i < kMaxVec; ++i) {
                const int v = v0 + lane + 32 * i;
                wv[i] = (v < n_v4) ? __ldg(wrow + v) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < kMaxVec; ++i) {
                const int v = v0 + lane + 32 * i;
                if (v >= n_v4) break;
#pragma unroll
                for (int q = 0; q < kRoRows; ++q) {
                    const float4 sv = *reinterpret_cast<const float4*>(&s_s[q * Hd + 4 * v]);
                    acc[q] = fmaf(wv[i].x, sv.x, acc[q]); acc[q] = fmaf(wv[i].y, sv.y, acc[q]);
                    acc[q] = fmaf(wv[i].z, sv.z, acc[q]); acc[q] = fmaf(wv[i].w, sv.w, acc[q]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < kRoRows; ++q) {
            for (int off = 16; off > 0; off >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], off);
            const int r = r0 + q;
            if (lane == 0 && r < R) {
                if (o < n_cls) out_cls[static_cast<size_t>(r) * n_cls + o] = acc[q];
                else out_box[static_cast<size_t>(r) * n_box + (o - n_cls)] = acc[q];
            }
        }
    }
}


// ------------------------------------------------------- RPN proposal decode
// The step after RPNHeadSNN in RegionProposalNetwork.forward (rpn.py:636-670): the reference permutes all
// A*H*W logits and 4*A*H*W deltas of every level to (H, W, A) order, generates every anchor, decodes every
// box and only then keeps the per-level top-k.  Here the top-k runs on the head's native NCHW logits
// (index = (a*H + h)*W + w) and this kernel touches ONLY the selected entries: it regenerates their anchors
// analytically (torchvision AnchorGenerator: rounded base anchor + (w*stride_w, h*stride_h)), gathers the 4
// deltas from the NCHW tensor, decodes with BoxCoder(1,1,1,1) (det_utils.BoxCoder.decode_single, dw/dh clamped
// to log(1000/16)) and applies the sigmoid -- one thread per selected anchor.
constexpr int kPropMaxLevels = 8;
struct PropLevel {
    const float* logits;      // [N][A][H][W]
    const float* deltas;      // [N][4A][H][W]
    const float* base;        // [A][4] base anchors of this level (x1, y1, x2, y2)
    int H, W, k, k_begin;     // k selected per image, offset of this level inside the K_total selected of an image
    int stride_h, stride_w;
    long long anchor_begin;   // index of the level's first anchor in the reference's concatenated (level, h, w, a) order
};
struct PropParams {
    PropLevel lv[kPropMaxLevels];
    int n_levels, N, A, K_total;
    float clip;               // log(1000 / 16)
    const long long* idx;     // [N][K_total] selected index inside the level's [A][H][W] logits of that image
    float* boxes;             // [N][K_total][4]
    float* scores;            // [N][K_total] sigmoid(objectness)
    float* logit_out;         // [N][K_total] raw objectness (nullable)
    long long* ref_index;     // [N][K_total] index in the reference's (level, h, w, a) anchor order (nullable)
};

__global__ void __launch_bounds__(256) rpn_decode_selected_kernel(const __grid_constant__ PropParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.N * p.K_total) return;
    const int n = i / p.K_total, kk = i - n * p.K_total;
    int l = 0;
    while (l + 1 < p.n_levels && kk >= p.lv[l + 1].k_begin) ++l;
    const PropLevel& L = p.lv[l];
    const long long id = p.idx[i];
    const int hw = L.H * L.W;
    const int a = static_cast<int>(id / hw);
    const int rem = static_cast<int>(id - static_cast<long long>(a) * hw);
    const int h = rem / L.W, w = rem - h * L.W;
    const float logit = L.logits[(static_cast<size_t>(n) * p.A + a) * hw + rem];
    const float* d = L.deltas + (static_cast<size_t>(n) * 4 * p.A + 4 * a) * hw + rem;
    const float dx = d[0], dy = d[hw];
    const float dw = fminf(d[2 * static_cast<size_t>(hw)], p.clip), dh = fminf(d[3 * static_cast<size_t>(hw)], p.clip);
    const float sx = static_cast<float>(w * L.stride_w), sy = static_cast<float>(h * L.stride_h);
    const float x1 = L.base[4 * a + 0] + sx, y1 = L.base[4 * a + 1] + sy;
    const float x2 = L.base[4 * a + 2] + sx, y2 = L.base[4 * a + 3] + sy;
    // det_utils.BoxCoder.decode_single, weights (1, 1, 1, 1), op for op (no FMA contraction)
    const float widths = __fsub_rn(x2, x1), heights = __fsub_rn(y2, y1);
    const float ctr_x = __fadd_rn(x1, __fmul_rn(0.5f, widths)), ctr_y = __fadd_rn(y1, __fmul_rn(0.5f, heights));
    const float pcx = __fadd_rn(__fmul_rn(dx, widths), ctr_x), pcy = __fadd_rn(__fmul_rn(dy, heights), ctr_y);
    const float pw = __fmul_rn(expf(dw), widths), ph = __fmul_rn(expf(dh), heights);
    const float cw = __fmul_rn(0.5f, pw), ch = __fmul_rn(0.5f, ph);
    float4 o;
    o.x = __fsub_rn(pcx, cw); o.y = __fsub_rn(pcy, ch); o.z = __fadd_rn(pcx, cw); o.w = __fadd_rn(pcy, ch);
    reinterpret_cast<float4*>(p.boxes)[i] = o;
    p.scores[i] = 1.0f / (1.0f + expf(-logit));
    if (p.logit_out != nullptr) p.logit_out[i] = logit;
    if (p.ref_index != nullptr) p.ref_index[i] = L.anchor_begin + (static_cast<long long>(h) * L.W + w) * p.A + a;
}

// Sort keys for the per-level top-k that precedes the decode: the reference takes top-k on the (H, W, A)-flattened
// logits (rpn.py:248-259, 468-472); taking it on the head's NCHW tensor picks the same VALUES but may break ties --
// e.g. the exactly-zero membranes of pixels where no shared_lif neuron spiked -- in another order.  The key is unique
// per anchor: (order-preserving integer image of the fp32 logit) << 32 | (2^32 - 1 - index in the reference's order),
// so top-k on the keys = "largest logit first, ties by the lowest reference index", independent of how the top-k
// implementation treats equal elements, and still without the permute / reshape copy of the logits.
struct KeyLevel { const float* logits; long long* keys; int A, HW; long long begin; };   // begin: first flat item of the level
struct KeyParams { KeyLevel lv[kPropMaxLevels]; int n_levels, N; long long total; };

__global__ void __launch_bounds__(256) rpn_topk_keys_kernel(const __grid_constant__ KeyParams p) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < p.total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        int l = 0;
        while (l + 1 < p.n_levels && i >= p.lv[l + 1].begin) ++l;
        const KeyLevel& L = p.lv[l];
        const long long local = i - L.begin;                       // n * (A * HW) + a * HW + rem
        const int per = L.A * L.HW;
        const int pos = static_cast<int>(local % per);
        const int a = pos / L.HW, rem = pos - a * L.HW;
        const int b = __float_as_int(__fadd_rn(L.logits[local], 0.0f));          // -0 -> +0
        const int k32 = b ^ ((b >> 31) & 0x7FFFFFFF);                            // integer order == float order
        const unsigned int ref = static_cast<unsigned int>(rem) * L.A + a;
        L.keys[local] = (static_cast<long long>(k32) << 32) | static_cast<long long>(0xFFFFFFFFu - ref);
    }
}

// ---- the per-level top-k itself, for all levels and images at once
// torch.topk on those int64 keys is one call per level, each a handful of kernels working on N (= 2) rows: 0.56 ms for
// the Cityscapes batch (r02g) -- the top-k, not the decode, was the cost of the proposal selection.  Here: an exact radix
// select over all L x N segments at once.  The key of an anchor is recomputed from its logit on every pass (no key
// tensor); six histogram passes (digits of 11, 11, 11, 11, 10, 10 bits from the top, 2048-bin shared-memory histograms
// with warp-aggregated atomics, so a plateau of equal logits costs one atomic per warp) narrow the k-th largest key down
// to all its 64 bits -- the last block of a segment to finish a pass (ticket counter) scans the segment's histogram and
// publishes the digit -- then one pass collects the k keys >= that threshold (keys are unique, so there are exactly k)
// and one block per segment sorts them (bitonic, shared memory) and writes the positions, largest key first.
constexpr int kTopkPasses = 6;
constexpr int kTopkBins = 2048;
constexpr int kTopkChunk = 4096;            // anchors per block and pass
constexpr int kTopkThreads = 256;
constexpr int kTopkMaxK = 2048;             // candidates sorted per segment (pre_nms_top_n: 1000 at test time, 2000 in training)
__host__ __device__ constexpr int topk_bits(int pass) { return pass < 4 ? 11 : 10; }
__host__ __device__ constexpr int topk_shift(int pass) { return pass < 4 ? 64 - 11 * (pass + 1) : 10 * (5 - pass); }

struct TopkLevel { const float* logits; int A, HW, n, k, k_begin, chunks; };      // n = A * HW anchors per image
struct TopkParams {
    TopkLevel lv[kPropMaxLevels];
    int n_levels, N, K_total;
    unsigned long long* prefix;     // [segments] digits of the k-th largest key found so far (segment = level * N + image)
    unsigned int* k_rem;            // [segments] rank of that key among the anchors that share the prefix
    unsigned int* done;             // [segments][kTopkPasses] blocks that finished the pass
    unsigned int* hist;             // [segments][kTopkPasses][kTopkBins]
    unsigned int* n_cand;           // [segments]
    unsigned long long* cand;       // [segments][kTopkMaxK]
    long long* idx;                 // [N][K_total] selected positions inside the level's [A][H][W] logits
};

// unsigned 64-bit key: larger = better; (order-preserving image of the logit, -0 folded into +0) then the LOWER index in
// the reference's (H, W, A) flattening wins
__device__ __forceinline__ unsigned long long topk_key(float logit, int pos, int A, int HW) {
    const int a = pos / HW, rem = pos - a * HW;
    const int b = __float_as_int(__fadd_rn(logit, 0.0f));
    const unsigned int hi = static_cast<unsigned int>(b ^ ((b >> 31) & 0x7FFFFFFF)) ^ 0x80000000u;
    return (static_cast<unsigned long long>(hi) << 32) | (0xFFFFFFFFu - (static_cast<unsigned int>(rem) * A + a));
}

template <int PASS>
__global__ void __launch_bounds__(kTopkThreads) rpn_topk_pass_kernel(const __grid_constant__ TopkParams p) {
    constexpr int shift = topk_shift(PASS), bits = topk_bits(PASS);
    const int seg = blockIdx.y, l = seg / p.N, n = seg - l * p.N;
    const TopkLevel& L = p.lv[l];
    if (static_cast<int>(blockIdx.x) >= L.chunks) return;
    __shared__ unsigned int h[kTopkBins];
    __shared__ unsigned int s_part[kTopkThreads];
    __shared__ unsigned int s_ticket;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int b = tid; b < kTopkBins; b += kTopkThreads) h[b] = 0u;
    __syncthreads();
    const unsigned long long prefix = PASS ? p.prefix[seg] : 0ull;
    const float* base = L.logits + static_cast<size_t>(n) * L.n;
    const int i0 = blockIdx.x * kTopkChunk, i1 = min(L.n, i0 + kTopkChunk);
    for (int i = i0 + tid; i < i0 + kTopkChunk; i += kTopkThreads) {      // uniform trip count: the warp votes below are full
        unsigned int d = 0xFFFFFFFFu;
        if (i < i1) {
            const unsigned long long key = topk_key(__ldg(base + i), i, L.A, L.HW);
            bool in = true;
            if constexpr (PASS > 0) in = (key >> (shift + bits)) == (prefix >> (shift + bits));
            if (in) d = static_cast<unsigned int>(key >> shift) & ((1u << bits) - 1u);
        }
        const unsigned int m = __match_any_sync(0xFFFFFFFFu, d);
        if (d != 0xFFFFFFFFu && lane == __ffs(static_cast<int>(m)) - 1) atomicAdd(&h[d], static_cast<unsigned int>(__popc(m)));
    }
    __syncthreads();
    unsigned int* gh = p.hist + (static_cast<size_t>(seg) * kTopkPasses + PASS) * kTopkBins;
    for (int b = tid; b < kTopkBins; b += kTopkThreads)
        if (h[b]) atomicAdd(&gh[b], h[b]);
    __threadfence();
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(&p.done[seg * kTopkPasses + PASS], 1u);
    __syncthreads();
    if (s_ticket != static_cast<unsigned int>(L.chunks - 1)) return;
    // ---- last block of the segment: the digit of the k-th largest key = the highest bin whose suffix count reaches k
    __threadfence();
    const unsigned int k_rem = PASS ? p.k_rem[seg] : static_cast<unsigned int>(L.k);
    constexpr int per = kTopkBins / kTopkThreads;                 // bins per thread, thread 0 owns the TOP bins
    unsigned int mine[per], sum = 0u;
#pragma unroll
    for (int j = 0; j < per; ++j) {
        mine[j] = __ldcg(&gh[kTopkBins - 1 - (tid * per + j)]);
        sum += mine[j];
    }
    s_part[tid] = sum;
    __syncthreads();
    if (tid == 0) {                                               // 256 partial sums: serial exclusive scan from the top
        unsigned int run = 0u;
        for (int t = 0; t < kTopkThreads; ++t) { const unsigned int v = s_part[t]; s_part[t] = run; run += v; }
    }
    __syncthreads();
    unsigned int above = s_part[tid];                             // anchors in bins above this thread's
#pragma unroll
    for (int j = 0; j < per; ++j) {
        if (above < k_rem && above + mine[j] >= k_rem) {          // exactly one (thread, j) satisfies this
            const unsigned long long digit = static_cast<unsigned long long>(kTopkBins - 1 - (tid * per + j));
            p.prefix[seg] = prefix | (digit << shift);
            p.k_rem[seg] = k_rem - above;
        }
        above += mine[j];
    }
}

__global__ void __launch_bounds__(kTopkThreads) rpn_topk_collect_kernel(const __grid_constant__ TopkParams p) {
    const int seg = blockIdx.y, l = seg / p.N, n = seg - l * p.N;
    const TopkLevel& L = p.lv[l];
    if (static_cast<int>(blockIdx.x) >= L.chunks) return;
    const unsigned long long thr = p.prefix[seg];
    const float* base = L.logits + static_cast<size_t>(n) * L.n;
    const int i0 = blockIdx.x * kTopkChunk, i1 = min(L.n, i0 + kTopkChunk);
    for (int i = i0 + threadIdx.x; i < i1; i += kTopkThreads) {
        const unsigned long long key = topk_key(__ldg(base + i), i, L.A, L.HW);
        if (key >= thr) {
            const unsigned int slot = atomicAdd(&p.n_cand[seg], 1u);
            if (slot < static_cast<unsigned int>(kTopkMaxK)) p.cand[static_cast<size_t>(seg) * kTopkMaxK + slot] = key;
        }
    }
}

__global__ void __launch_bounds__(1024) rpn_topk_sort_kernel(const __grid_constant__ TopkParams p) {
    __shared__ unsigned long long s[kTopkMaxK];
    const int seg = blockIdx.x, l = seg / p.N, n = seg - l * p.N;
    const TopkLevel& L = p.lv[l];
    const int tid = threadIdx.x;
    const int cnt = min(static_cast<int>(p.n_cand[seg]), kTopkMaxK);
    for (int i = tid; i < kTopkMaxK; i += 1024) s[i] = i < cnt ? p.cand[static_cast<size_t>(seg) * kTopkMaxK + i] : 0ull;
    __syncthreads();
    for (int size = 2; size <= kTopkMaxK; size <<= 1)             // bitonic sort, descending
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < kTopkMaxK / 2; t += 1024) {
                const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const unsigned long long a = s[lo], b = s[hi];
                if ((a < b) == desc) { s[lo] = b; s[hi] = a; }
            }
            __syncthreads();
        }
    for (int j = tid; j < L.k && j < cnt; j += 1024) {
        const unsigned int ref = 0xFFFFFFFFu - static_cast<unsigned int>(s[j] & 0xFFFFFFFFull);
        const unsigned int a = ref % static_cast<unsigned int>(L.A), rem = ref / static_cast<unsigned int>(L.A);
        p.idx[static_cast<size_t>(n) * p.K_total + L.k_begin + j] = static_cast<long long>(a) * L.HW + rem;
    }
}

// ------------------------------------------------- RoIAlign fused with the encoder
// The step before FastRCNNPredictorSNNFull in RoIHeadsSNN.forward (roi_heads.py:1217 -> faster_rcnn.py:474-494):
// MultiScaleRoIAlign writes [R][C][7][7] fp32 only for the encoder to threshold it into spikes.  Here one thread
// computes 8 consecutive k = c*49 + ph*7 + pw of a RoI with torchvision's roi_align arithmetic (aligned = False,
// sampling_ratio samples per bin, bilinear_interpolate with its border rules) and feeds them straight into the
// lock-step encoder: the pooled tensor never exists in HBM (50 MB written + read per image).
constexpr int kRoiMaxLevels = 8;
struct RoiLevel { const float* x; int H, W; float scale; };
struct RoiEncParams {
    RoiLevel lv[kRoiMaxLevels];
    int n_levels, C, R, P, sampling, T_live, wb;
    const float* rois;        // [R][5]: batch index, x1, y1, x2, y2 (image coordinates)
    const int* roi_level;     // [R] FPN level of each RoI (torchvision LevelMapper)
    uint8_t* words;           // [R][C*P*P] spike-train words of wb bytes
    float* pooled;            // nullable (tests): [R][C*P*P] the RoIAlign values themselves
};

__device__ __forceinline__ float roi_bilinear(const float* __restrict__ f, int H, int W, float y, float x) {
    if (y < -1.0f || y > static_cast<float>(H) || x < -1.0f || x > static_cast<float>(W)) return 0.f;
    if (y <= 0.f) y = 0.f;
    if (x <= 0.f) x = 0.f;
    int y_low = static_cast<int>(y), x_low = static_cast<int>(x);
    int y_high, x_high;
    if (y_low >= H - 1) { y_high = y_low = H - 1; y = static_cast<float>(y_low); } else { y_high = y_low + 1; }
    if (x_low >= W - 1) { x_high = x_low = W - 1; x = static_cast<float>(x_low); } else { x_high = x_low + 1; }
    const float ly = y - y_low, lx = x - x_low, hy = 1.f - ly, hx = 1.f - lx;
    const float v1 = __ldg(f + y_low * W + x_low), v2 = __ldg(f + y_low * W + x_high);
    const float v3 = __ldg(f + y_high * W + x_low), v4 = __ldg(f + y_high * W + x_high);
    const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
    return w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4;
}

// One bilinear sample of a bin, prepared once and applied to every channel: corner offsets inside a channel plane and
// corner weights (all zero for a sample outside the map, as roi_bilinear returns 0 there).
struct RoiSample { int o1, o2, o3, o4; float w1, w2, w3, w4; };

__device__ __forceinline__ RoiSample roi_prepare(int H, int W, float y, float x) {
    RoiSample s{0, 0, 0, 0, 0.f, 0.f, 0.f, 0.f};
    if (y < -1.0f || y > static_cast<float>(H) || x < -1.0f || x > static_cast<float>(W)) return s;
    if (y <= 0.f) y = 0.f;
    if (x <= 0.f) x = 0.f;
    int y_low = static_cast<int>(y), x_low = static_cast<int>(x);
    int y_high, x_high;
    if (y_low >= H - 1) { y_high = y_low = H - 1; y = static_cast<float>(y_low); } else { y_high = y_low + 1; }
    if (x_low >= W - 1) { x_high = x_low = W - 1; x = static_cast<float>(x_low); } else { x_high = x_low + 1; }
    const float ly = y - y_low, lx = x - x_low, hy = 1.f - ly, hx = 1.f - lx;
    s.o1 = y_low * W + x_low; s.o2 = y_low * W + x_high; s.o3 = y_high * W + x_low; s.o4 = y_high * W + x_high;
    s.w1 = hy * hx; s.w2 = hy * lx; s.w3 = ly * hx; s.w4 = ly * lx;
    return s;
}

constexpr int kRoiChPerThread = 64;

// Thread = (RoI, bin (ph, pw), group of 64 channels).  The sampling grid of a bin -- positions, border rules and the
// four bilinear weights of each sample -- depends on the RoI and the bin only, so it is prepared ONCE (up to 2 x 2
// samples, torchvision's sampling_ratio = 2 of Faster R-CNN) and applied to the thread's 64 channel planes: 16 loads
// + 16 multiply-adds + the comparator-bank encoder per output instead of re-deriving the grid for every channel
// (the first version: 8 consecutive k = c*PP + bin per thread).  Consecutive threads are consecutive bins of a RoI, so
// a warp's loads fall into a few rows of one channel plane and its stores are consecutive words.
// Other sampling grids (adaptive, or more than 4 samples per bin) take the per-channel path.
// What bounds it (r02b): not DRAM (8.6 %) but the L1 data path.  A 2 x 2 x 49 sampling grid over a window of 15-29
// pixels a side (torchvision's level mapping puts a RoI of 112-224 px on the stride-8 level, i.e. 14-28 feature
// pixels) reads almost every pixel of the window exactly once per channel, so there is no reuse for shared-memory
// staging to exploit -- a version with one block per RoI and the window staged in shared memory was 4x SLOWER
// (1.21 vs 0.30 ms, profiles/r02/r02b_next_rows.json) and was dropped -- and every warp-level load touches ~8 cache lines
// for its 32 lanes: 401 M lane-loads = 12.5 M warp-loads x ~8 wavefronts / (148 SMs x 1 wavefront/clk) = 0.36 ms at
// 1.9 GHz, the measured 0.30 ms.  The useful bytes alone (<= 784 floats per RoI and channel, 1.6 GB per 2000 RoIs)
// are 0.28 ms of L2 -> SM bandwidth.  kRoiUnroll channel planes per iteration keep 16 x kRoiUnroll loads in flight.
template <int NT, int kRoiUnroll>
__global__ void __launch_bounds__(256, kRoiUnroll <= 2 ? 3 : 2) roi_align_encode_kernel(const __grid_constant__ RoiEncParams p) {
    const uint32_t tmask = (p.T_live >= 32) ? 0xFFFFFFFFu : ((1u << p.T_live) - 1u);
    const int PP = p.P * p.P;
    const int K = p.C * PP;
    const int groups = (p.C + kRoiChPerThread - 1) / kRoiChPerThread;
    const size_t bins = static_cast<size_t>(p.R) * PP;
    const size_t total = bins * groups;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i / bins);
        const size_t rb = i - static_cast<size_t>(g) * bins;
        const int r = static_cast<int>(rb / PP);
        const int bin = static_cast<int>(rb - static_cast<size_t>(r) * PP);
        const int ph = bin / p.P, pw = bin - ph * p.P;
        const float* roi = p.rois + 5 * static_cast<size_t>(r);
        const RoiLevel& L = p.lv[p.roi_level[r]];
        const int b = static_cast<int>(roi[0]);
        const float rsw = roi[1] * L.scale, rsh = roi[2] * L.scale, rew = roi[3] * L.scale, reh = roi[4] * L.scale;
        const float roi_w = fmaxf(rew - rsw, 1.f), roi_h = fmaxf(reh - rsh, 1.f);
        const float bin_h = roi_h / static_cast<float>(p.P), bin_w = roi_w / static_cast<float>(p.P);
        const int gh = p.sampling > 0 ? p.sampling : static_cast<int>(ceilf(roi_h / p.P));
        const int gw = p.sampling > 0 ? p.sampling : static_cast<int>(ceilf(roi_w / p.P));
        const float count = fmaxf(static_cast<float>(gh * gw), 1.f);
        const int c0 = g * kRoiChPerThread, c1 = min(p.C, c0 + kRoiChPerThread);
        const size_t plane = static_cast<size_t>(L.H) * L.W;
        const float* f = L.x + (static_cast<size_t>(b) * p.C + c0) * plane;
        const size_t out0 = static_cast<size_t>(r) * K + static_cast<size_t>(c0) * PP + bin;
        const bool fast = gh <= 2 && gw <= 2;
        RoiSample sm[4];
        if (fast) {
#pragma unroll
            for (int iy = 0; iy < 2; ++iy)
#pragma unroll
                for (int ix = 0; ix < 2; ++ix) {
                    const float y = rsh + ph * bin_h + static_cast<float>(iy + .5f) * bin_h / static_cast<float>(gh);
                    const float x = rsw + pw * bin_w + static_cast<float>(ix + .5f) * bin_w / static_cast<float>(gw);
                    sm[iy * 2 + ix] = (iy < gh && ix < gw) ? roi_prepare(L.H, L.W, y, x) : RoiSample{0, 0, 0, 0, 0.f, 0.f, 0.f, 0.f};
                }
        }
        // kRoiUnroll channel planes per iteration: their 16 loads each are all requested before the first is used
        // (ncu r01ba: 78 % of the stall samples of the one-plane loop were waits for these loads)
        size_t out = out0;
        for (int c = c0; c < c1; c += kRoiUnroll, f += kRoiUnroll * plane, out += static_cast<size_t>(kRoiUnroll) * PP) {
            float val[kRoiUnroll];
            if (fast) {
                float v[kRoiUnroll][16];
#pragma unroll
                for (int u = 0; u < kRoiUnroll; ++u) {
                    const float* fu = f + (c + u < c1 ? u : 0) * plane;       // past the last channel: re-read plane c
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        v[u][4 * q + 0] = __ldg(fu + sm[q].o1); v[u][4 * q + 1] = __ldg(fu + sm[q].o2);
                        v[u][4 * q + 2] = __ldg(fu + sm[q].o3); v[u][4 * q + 3] = __ldg(fu + sm[q].o4);
                    }
                }
#pragma unroll
                for (int u = 0; u < kRoiUnroll; ++u) {
                    float acc = 0.f;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        acc += sm[q].w1 * v[u][4 * q + 0] + sm[q].w2 * v[u][4 * q + 1] + sm[q].w3 * v[u][4 * q + 2] + sm[q].w4 * v[u][4 * q + 3];
                    val[u] = acc / count;
                }
            } else {
#pragma unroll
                for (int u = 0; u < kRoiUnroll; ++u) {
                    const float* fu = f + (c + u < c1 ? u : 0) * plane;
                    float acc = 0.f;
                    for (int iy = 0; iy < gh; ++iy) {
                        const float y = rsh + ph * bin_h + static_cast<float>(iy + .5f) * bin_h / static_cast<float>(gh);
                        for (int ix = 0; ix < gw; ++ix) {
                            const float x = rsw + pw * bin_w + static_cast<float>(ix + .5f) * bin_w / static_cast<float>(gw);
                            acc += roi_bilinear(fu, L.H, L.W, y, x);
                        }
                    }
                    val[u] = acc / count;
                }
            }
#pragma unroll
            for (int u = 0; u < kRoiUnroll; ++u) {
                if (c + u >= c1) break;
                const size_t o = out + static_cast<size_t>(u) * PP;
                if (p.pooled != nullptr) p.pooled[o] = val[u];
                const uint32_t w = encode_word<NT>(val[u], tmask);
                if (p.wb == 1) p.words[o] = static_cast<uint8_t>(w);
                else if (p.wb == 2) reinterpret_cast<uint16_t*>(p.words)[o] = static_cast<uint16_t>(w);
                else reinterpret_cast<uint32_t*>(p.words)[o] = w;
            }
        }
    }
}

// Exhaustive check of the comparator bank against the simulation: every one of the 2^32 fp32 bit patterns.
template <int NT>
__global__ void __launch_bounds__(256) encoder_selftest_kernel(int T_live, unsigned long long* __restrict__ mismatches) {
    const uint32_t tmask = (T_live >= 32) ? 0xFFFFFFFFu : ((1u << T_live) - 1u);
    unsigned int bad = 0;
    for (unsigned long long b = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x; b < (1ull << 32);
         b += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
        const float x = __uint_as_float(static_cast<uint32_t>(b));
        bad += encode_word<NT>(x, tmask) != encode_train(x, T_live);
    }
    for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(mismatches, static_cast<unsigned long long>(bad));
}

// ---------------------------------------------------------------------------------------------------------------------
// Detector post-processing: RoIHeadsSNN.postprocess_detections, roi_heads.py:1075-1176, after the softmax
// and the box decode (those stay torch: element-wise, no synchronisation).  One block per image does what the reference
// does with ~40 small kernels and ~10 host synchronisations per image: clip the boxes to the image (l.1105), select the
// (RoI, class >= 1) pairs with score > score_thresh (l.1127-1128), keep the background box of every RoI none of whose
// classes passed (l.1139-1150), drop boxes below min_size (l.1153-1158), batched NMS of both sets (l.1161-1162), keep
// the detections_per_img best objects (l.1164) and write objects, then background (l.1170-1172).
// The NMS reproduces torchvision's CUDA path operation for operation, so the kept set and its order are the ones the
// reference gets on the same device: a stable descending sort of the scores (ties: the lower index first); greedy
// suppression inside a class; the IoU of torchvision's devIoU as its sm_100 SASS computes it (the keeper's area is an
// FMUL, the sum of the two areas one FFMA with it as the addend, IEEE division, `> threshold`); and batched_nms's own
// switch between the coordinate trick (boxes shifted by label * (max coordinate + 1), with the fp32 rounding that
// implies) up to 5000 boxes and per-class NMS on the unshifted boxes above that.
constexpr int kDetMaxCand = 8192;           // (RoI, class) pairs / RoIs per image the block sorts in shared memory
constexpr int kDetThreads = 1024;
constexpr int kDetMaxImages = 64;           // images per launch (the parameter block holds their geometry)
constexpr int kDetWords = kDetMaxCand / 32;    // bitmap words (alive / kept)

struct DetParams {
    const float* scores;            // [R_total][C] softmax scores
    const float* boxes;             // [R_total][C][4] decoded boxes, not clipped
    float* all_boxes;               // [R_total][C][4] clipped (the reference's `all_boxes`)
    float* out_boxes;               // [N][cap][4]
    float* out_scores;              // [N][cap]
    long long* out_labels;          // [N][cap]
    int* out_counts;                // [N][2] objects, background boxes written
    int N, C, cap, det_per_img;
    float score_thresh, nms_thresh, min_size;
    int row0[kDetMaxImages], rows[kDetMaxImages];
    float img_h[kDetMaxImages], img_w[kDetMaxImages];
};

// torch.clamp(min=0, max=hi): NaN stays NaN
__device__ __forceinline__ float det_clamp(float v, float hi) { return v < 0.f ? 0.f : (v > hi ? hi : v); }
__device__ __forceinline__ float4 det_clip(const float* b, float w, float h) {
    const float4 v = *reinterpret_cast<const float4*>(b);
    return make_float4(det_clamp(v.x, w), det_clamp(v.y, h), det_clamp(v.z, w), det_clamp(v.w, h));
}
// ascending sort key: higher score first (scores are softmax outputs, >= +0), then the lower index
__device__ __forceinline__ unsigned long long det_key(float score, unsigned int idx) {
    const int b = __float_as_int(score);
    const unsigned int hi = static_cast<unsigned int>(b ^ ((b >> 31) & 0x7FFFFFFF)) ^ 0x80000000u;
    return (static_cast<unsigned long long>(~hi) << 32) | idx;
}

// torchvision's devIoU(keeper a, candidate b) > threshold, as nvcc compiled it for sm_100 (see above)
__device__ __forceinline__ bool det_suppresses(const float4 a, const float Sa, const float4 b, const float thr) {
    const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z), top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
    // disjoint boxes (most pairs): the intersection is 0 (or NaN), the quotient 0, -0 or NaN, never > a threshold >= 0
    if (thr >= 0.f && (right <= left || bottom <= top)) return false;
    const float w = fmaxf(__fsub_rn(right, left), 0.f), h = fmaxf(__fsub_rn(bottom, top), 0.f);
    const float inter = __fmul_rn(w, h);
    const float sum = __fmaf_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y), Sa);
    const float den = __fsub_rn(sum, inter);
    // The correctly rounded quotient is only needed within 2^-20 of the threshold: with t = fl(thr * den) (relative error
    // 2^-24), inter > t (1 + 2^-20) implies inter / den > thr (1 + 2^-21), whose rounding is still > thr, and likewise
    // below -- the same boolean as the division for every input, without the division's ~40-instruction slow path on the
    // (almost all) pairs that are nowhere near the threshold.
    const float t = __fmul_rn(thr, den);
    if (thr >= 0.f && t > 1.0e-30f && t < 1.0e38f) {
        if (inter > __fmul_rn(t, 1.0f + 0x1p-20f)) return true;
        if (inter < __fmul_rn(t, 1.0f - 0x1p-20f)) return false;
    }
    return __fdiv_rn(inter, den) > thr;
}
__device__ __forceinline__ float det_area(const float4 a) { return __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y)); }

// Working set of one NMS problem in shared memory (boxes in sorted order) -- shared by the detector post-processing and
// the RPN proposal filter.
struct NmsSmem {
    float4* sbox;                   // [n] the boxes NMS sees (shifted by the coordinate trick or not)
    unsigned short* lab;            // [n] class / level: boxes of different labels never suppress each other
    unsigned int* alive;            // [words] not yet looked at and not suppressed
    unsigned int* keptb;            // [words] kept
    float4* kbox; float* karea; int* klab;      // [32] this round's keepers
    int* bidx;                      // [32] this round's candidates
    unsigned int* supp;             // [32] pair tests of a round
    int* scan;                      // [32] warp totals
    int* pref;                      // [words] keepers before a word
    int* misc;                      // [2] batch size, [3] keepers of the round, [4] cursor
};
constexpr int kNmsFixedBytes = 32 * (16 + 4 + 4 + 4 + 4 + 4) + 32;     // kbox .. scan, misc
__device__ __forceinline__ NmsSmem nms_carve(unsigned char* base, int max_n) {       // base 16-byte aligned, max_n % 32 == 0
    NmsSmem m;
    const int words = max_n / 32;
    m.sbox = reinterpret_cast<float4*>(base);
    m.lab = reinterpret_cast<unsigned short*>(base + static_cast<size_t>(max_n) * 16);
    m.alive = reinterpret_cast<unsigned int*>(base + static_cast<size_t>(max_n) * 18);
    m.keptb = m.alive + words;
    m.pref = reinterpret_cast<int*>(m.keptb + words);
    m.kbox = reinterpret_cast<float4*>(m.pref + words);
    m.karea = reinterpret_cast<float*>(m.kbox + 32);
    m.klab = reinterpret_cast<int*>(m.karea + 32);
    m.bidx = m.klab + 32;
    m.supp = reinterpret_cast<unsigned int*>(m.bidx + 32);
    m.scan = reinterpret_cast<int*>(m.supp + 32);
    m.misc = m.scan + 32;
    return m;
}
__host__ __device__ constexpr int nms_smem_bytes(int max_n) { return max_n * 18 + (max_n / 32) * 12 + kNmsFixedBytes; }

// ascending bitonic sort of P (a power of two) unique 64-bit keys in shared memory; all kDetThreads threads call it
__device__ __forceinline__ void block_bitonic_sort(unsigned long long* keys, int P) {
    for (int size = 2; size <= P; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < (P >> 1); t += kDetThreads) {
                const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1)), hi = lo | stride;
                const unsigned long long a = keys[lo], b = keys[hi];
                const bool up = (lo & size) == 0;
                if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
            }
            __syncthreads();
        }
}

// Greedy NMS over m.sbox[0, n) in order (all kDetThreads threads call it; m.sbox / m.lab filled and synchronised):
// afterwards m.keptb flags the first max_keep boxes kept.  32 candidates a round:
//  A  warp 0 takes the next 32 boxes still alive;
//  B  the block tests every pair of them (thread (l, m): would m suppress l?);
//  C  warp 0 reads the keepers off in order (a box is kept iff no KEPT earlier box of the round suppresses it; earlier
//     rounds' keepers have been applied to it already) and publishes them;
//  D  the block tests every later box still alive against the round's keepers (one warp per box, one lane per keeper).
__device__ __forceinline__ void nms_rounds(const NmsSmem& m, const int n, const int max_keep, const float thr) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_words = (n + 31) >> 5;
    for (int w = tid; w < n_words; w += kDetThreads) {
        m.alive[w] = (w * 32 + 32 <= n) ? 0xFFFFFFFFu : ((1u << (n - w * 32)) - 1u);
        m.keptb[w] = 0u;
    }
    if (tid == 0) m.misc[4] = 0;
    __syncthreads();
    int nk = 0;
    while (nk < max_keep) {
        if (warp == 0) {                            // A
            const int cursor = m.misc[4];
            int cnt = 0;
            for (int w0 = cursor >> 5; cnt < 32 && w0 < n_words; w0 += 32) {
                const int w = w0 + lane;
                unsigned int bits = w < n_words ? m.alive[w] : 0u;
                if (w == (cursor >> 5)) bits &= 0xFFFFFFFFu << (cursor & 31);
                const int c = __popc(bits);
                int incl = c;
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
                int pos = cnt + incl - c;
                while (bits && pos < 32) {
                    m.bidx[pos++] = w * 32 + __ffs(static_cast<int>(bits)) - 1;
                    bits &= bits - 1u;
                }
                cnt = min(32, cnt + __shfl_sync(0xffffffffu, incl, 31));
            }
            if (lane == 0) m.misc[2] = cnt;
        }
        __syncthreads();
        const int cnt = m.misc[2];
        if (cnt == 0) break;
        {                                           // B: l = warp, m = lane
            bool sup = false;
            if (warp < cnt && lane < warp) {
                const int il = m.bidx[warp], im = m.bidx[lane];
                if (m.lab[il] == m.lab[im]) {
                    const float4 a = m.sbox[im];
                    sup = det_suppresses(a, det_area(a), m.sbox[il], thr);
                }
            }
            const unsigned int bal = __ballot_sync(0xffffffffu, sup);
            if (lane == 0) m.supp[warp] = bal;
        }
        __syncthreads();
        if (warp == 0) {                            // C
            const unsigned int supp = m.supp[lane];
            unsigned int kept_mask = 0u, seen = 0u;
            int k_here = 0;
            for (int q = 0; q < cnt; ++q) {
                const unsigned int sq = __shfl_sync(0xffffffffu, supp, q);
                if (nk + k_here >= max_keep) break;                 // only the first max_keep are wanted
                seen |= 1u << q;
                if ((sq & kept_mask) == 0u) { kept_mask |= 1u << q; ++k_here; }
            }
            // every candidate looked at leaves the alive set; keepers are flagged and published for the block
            if (lane < cnt && ((seen >> lane) & 1u)) {
                const int me = m.bidx[lane];
                atomicAnd(&m.alive[me >> 5], ~(1u << (me & 31)));
                if ((kept_mask >> lane) & 1u) {
                    atomicOr(&m.keptb[me >> 5], 1u << (me & 31));
                    const int q = __popc(kept_mask & ((1u << lane) - 1u));
                    const float4 b = m.sbox[me];
                    m.kbox[q] = b; m.karea[q] = det_area(b); m.klab[q] = m.lab[me];
                }
            }
            if (lane == 0) { m.misc[3] = k_here; m.misc[4] = m.bidx[cnt - 1] + 1; }
        }
        __syncthreads();
        const int nkb = m.misc[3], from = m.misc[4];
        nk += nkb;
        if (nk >= max_keep) break;
        {                                           // D
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            float Sa = 0.f;
            int la = -1;
            if (lane < nkb) { a = m.kbox[lane]; Sa = m.karea[lane]; la = m.klab[lane]; }
            for (int jj = from + warp; jj < n; jj += kDetThreads / 32) {
                if (!((m.alive[jj >> 5] >> (jj & 31)) & 1u)) continue;          // uniform over the warp
                const bool sup = la == static_cast<int>(m.lab[jj]) && det_suppresses(a, Sa, m.sbox[jj], thr);
                if (__any_sync(0xffffffffu, sup) && lane == 0) atomicAnd(&m.alive[jj >> 5], ~(1u << (jj & 31)));
            }
        }
        __syncthreads();
    }
    __syncthreads();
}

// rank of every keeper among the keepers (m.pref[word] = keepers before the word); returns their number.  n <= 32 * kDetThreads
__device__ __forceinline__ int nms_kept_ranks(const NmsSmem& m, const int n) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_words = (n + 31) >> 5;
    const unsigned int bits = tid < n_words ? m.keptb[tid] : 0u;
    int incl = __popc(bits);
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) m.scan[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int v = m.scan[lane];
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
        m.scan[lane] = v;                           // inclusive totals of the warps
    }
    __syncthreads();
    if (tid < n_words) m.pref[tid] = incl - __popc(bits) + (warp ? m.scan[warp - 1] : 0);
    const int total = m.scan[31];
    __syncthreads();
    return total;
}

constexpr int kDetSmemBytes = kDetMaxCand * 8 + nms_smem_bytes(kDetMaxCand);

__global__ void __launch_bounds__(kDetThreads) det_postprocess_kernel(const __grid_constant__ DetParams p) {
    extern __shared__ __align__(16) unsigned char det_smem[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(det_smem);                      // [kDetMaxCand]
    const NmsSmem m = nms_carve(det_smem + static_cast<size_t>(kDetMaxCand) * 8, kDetMaxCand);
    int* s_cnt = m.misc;                                           // [0] candidates, [1] max coordinate bits
    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int C = p.C, Cm1 = C - 1, R = p.rows[img];
    const size_t r0 = static_cast<size_t>(p.row0[img]);
    const float W = p.img_w[img], H = p.img_h[img];
    const float* sc = p.scores + r0 * C;
    const float* bx = p.boxes + r0 * C * 4;

    // the reference's `all_boxes`: every class's box of every RoI, clipped
    for (int e = tid; e < R * C; e += kDetThreads)
        *reinterpret_cast<float4*>(p.all_boxes + (r0 * C + e) * 4) = det_clip(bx + static_cast<size_t>(e) * 4, W, H);

    int n_written = 0;
    for (int set = 0; set < 2; ++set) {                 // 0: objects (classes >= 1), 1: background (class 0)
        const int n_src = set == 0 ? R * Cm1 : R;
        if (tid < 2) s_cnt[tid] = 0;
        __syncthreads();
        // ---- candidates (any order: the sort below fixes it; the key carries the reference's flat index
        //      r * (C - 1) + (c - 1), or r)
        float my_max = 0.f;
        for (int f0 = 0; f0 < n_src; f0 += kDetThreads) {
            const int f = f0 + tid;
            bool pass = false;
            float s = 0.f;
            if (f < n_src) {
                const int r = set == 0 ? f / Cm1 : f;
                const int c = set == 0 ? 1 + (f - r * Cm1) : 0;
                s = sc[static_cast<size_t>(r) * C + c];
                if (set == 0) pass = s > p.score_thresh;
                else {
                    pass = s >= 0.f;                    // (keeps NaN rows out, as the reference's torch.where does)
                    for (int k = 1; k < C; ++k) pass = pass && !(sc[static_cast<size_t>(r) * C + k] > p.score_thresh);
                }
                if (pass) {
                    const float4 b = det_clip(bx + (static_cast<size_t>(r) * C + c) * 4, W, H);
                    pass = (b.z - b.x >= p.min_size) && (b.w - b.y >= p.min_size);     // remove_small_boxes
                    if (pass) my_max = fmaxf(my_max, fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
                }
            }
            const unsigned int bal = __ballot_sync(0xffffffffu, pass);
            int base = 0;
            if (lane == 0 && bal) base = atomicAdd(&s_cnt[0], __popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (pass) keys[base + __popc(bal & ((1u << lane) - 1u))] = det_key(s, static_cast<unsigned int>(f));
        }
        for (int o = 16; o > 0; o >>= 1) my_max = fmaxf(my_max, __shfl_xor_sync(0xffffffffu, my_max, o));
        if (lane == 0) atomicMax(&s_cnt[1], __float_as_int(my_max));                   // clipped coordinates are >= 0
        __syncthreads();
        const int n = s_cnt[0];
        const float shift1 = __fadd_rn(__int_as_float(s_cnt[1]), 1.0f);
        int P = 2;
        while (P < n) P <<= 1;
        for (int t = n + tid; t < P; t += kDetThreads) keys[t] = ~0ull;
        __syncthreads();
        block_bitonic_sort(keys, P);                    // stable descending sort of the scores (unique keys)
        // ---- the boxes NMS sees: batched_nms shifts them by label * (max + 1) unless there are more than 5000
        const bool trick = n * 4 <= 20000;
        for (int t = tid; t < n; t += kDetThreads) {
            const unsigned int f = static_cast<unsigned int>(keys[t]);
            const int r = set == 0 ? static_cast<int>(f) / Cm1 : static_cast<int>(f);
            const int c = set == 0 ? 1 + (static_cast<int>(f) - r * Cm1) : 0;
            float4 b = det_clip(bx + (static_cast<size_t>(r) * C + c) * 4, W, H);
            if (trick) {
                const float off = __fmul_rn(static_cast<float>(c), shift1);
                b.x = __fadd_rn(b.x, off); b.y = __fadd_rn(b.y, off); b.z = __fadd_rn(b.z, off); b.w = __fadd_rn(b.w, off);
            }
            m.sbox[t] = b;
            m.lab[t] = static_cast<unsigned short>(c);
        }
        __syncthreads();
        nms_rounds(m, n, set == 0 ? p.det_per_img : 0x7FFFFFFF, p.nms_thresh);
        // ---- write the keepers in sorted order
        const int total = nms_kept_ranks(m, n);
        for (int t = tid; t < n; t += kDetThreads) {
            const unsigned int wbits = m.keptb[t >> 5];
            if (!((wbits >> (t & 31)) & 1u)) continue;
            const int row = n_written + m.pref[t >> 5] + __popc(wbits & ((1u << (t & 31)) - 1u));
            if (row >= p.cap) continue;
            const unsigned int f = static_cast<unsigned int>(keys[t]);
            const int r = set == 0 ? static_cast<int>(f) / Cm1 : static_cast<int>(f);
            const int c = set == 0 ? 1 + (static_cast<int>(f) - r * Cm1) : 0;
            const size_t o = static_cast<size_t>(img) * p.cap + row;
            *reinterpret_cast<float4*>(p.out_boxes + o * 4) = det_clip(bx + (static_cast<size_t>(r) * C + c) * 4, W, H);
            p.out_scores[o] = sc[static_cast<size_t>(r) * C + c];
            p.out_labels[o] = c;
        }
        if (tid == 0) p.out_counts[img * 2 + set] = total;
        n_written += total;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The tail of RegionProposalNetwork.filter_proposals (rpn.py:493-525) on the selected and decoded entries of
// snn_rpn_topk_select / snn_rpn_decode_selected: clip to the image, remove_small_boxes, score threshold, batched NMS over
// the levels, the post_nms_top_n best.  torchvision's NMS spends milliseconds here (its keeper scan is one serial pass
// over all 4864 boxes of an image).  batched NMS never lets boxes of different levels interact, so every (image, level)
// is its own NMS problem: one block each sorts its <= 2048 candidates and runs the rounds above -- on the boxes as
// batched_nms shifts them (level * (max coordinate of the IMAGE + 1), fp32 rounding included; no shift above 5000
// candidates per image), so the kept set is torchvision's -- and writes its keepers' (score, index) keys; the last
// block of an image to finish (ticket) merges the levels by one more sort and writes the post_nms_top_n best, score
// descending, ties by the lower index, as the stable sort inside torchvision's nms orders them.
constexpr int kRpnNmsMaxLevel = 2048;       // candidates per (image, level) = kTopkMaxK
constexpr int kRpnNmsMaxKept = 8192;        // keepers per image the merging block sorts
constexpr int kRpnNmsSmemBytes = kRpnNmsMaxKept * 8 + nms_smem_bytes(kRpnNmsMaxLevel);

struct RpnNmsParams {
    const float* props;             // [N][K][4] decoded boxes, not clipped
    const float* probs;             // [N][K]
    float* out_boxes;               // [N][post_n][4]
    float* out_scores;              // [N][post_n]
    int* out_counts;                // [N]
    unsigned long long* kept_keys;  // workspace [N][L][kRpnNmsMaxLevel]
    int* kept_cnt;                  // workspace [N][L]
    unsigned int* done;             // workspace [N], zero before the launch
    int N, L, K, post_n;
    float min_size, score_thresh, nms_thresh;
    int k_begin[kPropMaxLevels + 1];
    float img_h[kDetMaxImages], img_w[kDetMaxImages];
};

__global__ void __launch_bounds__(kDetThreads) rpn_nms_kernel(const __grid_constant__ RpnNmsParams p) {
    extern __shared__ __align__(16) unsigned char det_smem[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(det_smem);                      // [kRpnNmsMaxKept]
    const NmsSmem m = nms_carve(det_smem + static_cast<size_t>(kRpnNmsMaxKept) * 8, kRpnNmsMaxLevel);
    int* s_cnt = m.misc;                                // [0] candidates of the level, [1] max coordinate bits, [5] of the image, [6] ticket
    const int l = blockIdx.x, img = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    const float W = p.img_w[img], H = p.img_h[img];
    const float* bx = p.props + static_cast<size_t>(img) * p.K * 4;
    const float* sc = p.probs + static_cast<size_t>(img) * p.K;
    const int kb = p.k_begin[l], ke = p.k_begin[l + 1];
    if (tid < 8) s_cnt[tid] = 0;
    __syncthreads();
    // ---- the image's candidates (all levels): their number picks batched_nms's variant, their largest coordinate its shift
    {
        int cnt = 0;
        float my_max = 0.f;
        for (int k = tid; k < p.K; k += kDetThreads) {
            const float4 b = det_clip(bx + static_cast<size_t>(k) * 4, W, H);
            if ((b.z - b.x >= p.min_size) && (b.w - b.y >= p.min_size) && sc[k] >= p.score_thresh) {
                ++cnt;
                my_max = fmaxf(my_max, fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            my_max = fmaxf(my_max, __shfl_xor_sync(0xffffffffu, my_max, o));
        }
        if (lane == 0) { atomicAdd(&s_cnt[5], cnt); atomicMax(&s_cnt[1], __float_as_int(my_max)); }
    }
    // ---- this level's candidates; the key carries the entry's position in the image's level-major list
    for (int k0 = kb; k0 < ke; k0 += kDetThreads) {
        const int k = k0 + tid;
        bool pass = false;
        float s = 0.f;
        if (k < ke) {
            const float4 b = det_clip(bx + static_cast<size_t>(k) * 4, W, H);
            s = sc[k];
            pass = (b.z - b.x >= p.min_size) && (b.w - b.y >= p.min_size) && s >= p.score_thresh;
        }
        const unsigned int bal = __ballot_sync(0xffffffffu, pass);
        int base = 0;
        if (lane == 0 && bal) base = atomicAdd(&s_cnt[0], __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (pass) keys[base + __popc(bal & ((1u << lane) - 1u))] = det_key(s, static_cast<unsigned int>(k));
    }
    __syncthreads();
    const int n = s_cnt[0];
    const bool trick = s_cnt[5] * 4 <= 20000;
    const float off = trick ? __fmul_rn(static_cast<float>(l), __fadd_rn(__int_as_float(s_cnt[1]), 1.0f)) : 0.f;
    int P = 2;
    while (P < n) P <<= 1;
    for (int t = n + tid; t < P; t += kDetThreads) keys[t] = ~0ull;
    __syncthreads();
    block_bitonic_sort(keys, P);
    for (int t = tid; t < n; t += kDetThreads) {
        float4 b = det_clip(bx + static_cast<size_t>(static_cast<unsigned int>(keys[t])) * 4, W, H);
        if (trick) { b.x = __fadd_rn(b.x, off); b.y = __fadd_rn(b.y, off); b.z = __fadd_rn(b.z, off); b.w = __fadd_rn(b.w, off); }
        m.sbox[t] = b;
        m.lab[t] = 0;
    }
    __syncthreads();
    nms_rounds(m, n, p.post_n, p.nms_thresh);
    const int total = nms_kept_ranks(m, n);
    unsigned long long* mine = p.kept_keys + (static_cast<size_t>(img) * p.L + l) * kRpnNmsMaxLevel;
    for (int t = tid; t < n; t += kDetThreads) {
        const unsigned int wbits = m.keptb[t >> 5];
        if ((wbits >> (t & 31)) & 1u) mine[m.pref[t >> 5] + __popc(wbits & ((1u << (t & 31)) - 1u))] = keys[t];
    }
    if (tid == 0) p.kept_cnt[img * p.L + l] = total;
    // ---- the last block of the image merges the levels
    __threadfence();
    __syncthreads();
    if (tid == 0) s_cnt[6] = static_cast<int>(atomicAdd(&p.done[img], 1u));
    __syncthreads();
    if (s_cnt[6] != p.L - 1) return;
    __threadfence();
    // every level's keepers are sorted (score descending, then position): a keeper's place in the merged order is its
    // place in its own list plus, for every other list, the number of keys below it (binary search; keys are unique)
    int* s_off = m.pref;                                // [L + 1] starts of the levels' lists in keys[]
    if (tid == 0) {
        int acc = 0;
        for (int ll = 0; ll < p.L; ++ll) { s_off[ll] = acc; acc += __ldcg(&p.kept_cnt[img * p.L + ll]); }
        s_off[p.L] = acc;
    }
    __syncthreads();
    const int n_all = s_off[p.L];
    for (int ll = 0; ll < p.L; ++ll) {
        const unsigned long long* src = p.kept_keys + (static_cast<size_t>(img) * p.L + ll) * kRpnNmsMaxLevel;
        for (int t = s_off[ll] + tid; t < s_off[ll + 1]; t += kDetThreads) keys[t] = __ldcg(&src[t - s_off[ll]]);
    }
    __syncthreads();
    const int n_out = min(n_all, p.post_n);
    for (int t = tid; t < n_all; t += kDetThreads) {
        const unsigned long long key = keys[t];
        int rank = 0;
        for (int ll = 0; ll < p.L; ++ll) {
            int lo = s_off[ll], hi = s_off[ll + 1];
            if (t >= lo && t < hi) { rank += t - lo; continue; }
            while (lo < hi) {                           // first position of the list whose key is not below `key`
                const int mid = (lo + hi) >> 1;
                if (keys[mid] < key) lo = mid + 1; else hi = mid;
            }
            rank += lo - s_off[ll];
        }
        if (rank < n_out) {
            const unsigned int k = static_cast<unsigned int>(key);
            const size_t o = static_cast<size_t>(img) * p.post_n + rank;
            *reinterpret_cast<float4*>(p.out_boxes + o * 4) = det_clip(bx + static_cast<size_t>(k) * 4, W, H);
            p.out_scores[o] = sc[k];
        }
    }
    if (tid == 0) p.out_counts[img] = n_out;
}

}  // namespace snn
