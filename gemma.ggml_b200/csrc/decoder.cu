// decoder.cu — the persistent decode kernel: ONE cooperative launch per token (v5: reference-order arithmetic).
//
// Why: a Gemma-2B Q4_0 token is 1.41 GB of weights + 38 MB of KV = ~220 us of HBM time on a B200, spread
// over 163 mat-vecs and ~450 small nodes (ref: src/gemma_model.cpp:688-740).  One CTA per SM stays resident
// for the whole token; what limits it is (a) how cleanly the weights stream and (b) what every phase change costs.
//
// RESULTS ARE BIT-IDENTICAL TO THE SCALAR ggml ORDER of the reference's CPU path (the oracle's only order): every
// mat-vec output is the fp32 sum of its per-block terms added one by one in block order, attention is the
// reference's two-pass soft_max_ext with fp16 P and double-accumulated dots.  How that is made cheap:
//
//   * PASS STREAM.  At create time every weight matrix is re-laid out as a stream of PASSES: one pass = two rows x
//     one 64-block K segment (2304 B Q4_0 / 4352 B Q8_0), payloads then scales = what one consumer warp dots in one
//     go.  A CTA's share of a phase is ONE contiguous byte range, ordered K-SEGMENT-MAJOR (segment 0 of all its row
//     pairs, then segment 1, ...), so a trivial producer loop (lane 0 of warp 17: cp.async.bulk, SASS UBLKCP, <= 16
//     passes per ring slot, 5 slots, mbarrier completion) runs AHEAD across phase boundaries.
//   * CONSUMERS = 16 warps.  Warp w owns pass w of every slot: 4 x LDS.128 weights, the lane's two activation
//     blocks in registers, dp4a, the scalar reference's per-block fp32 term.  The 128 terms of a pass go to a
//     skewed (conflict-free) shared-memory table — NO in-warp reduction.
//   * CHAIN WARPS (warps 0, 18, 19).  Per ring slot one of them adds the terms to each row's accumulator strictly in
//     block order (64 dependent FADDs per lane, 32 rows at once) and FINISHES rows as soon as their last segment is
//     in: residual, flagged output, arg-max candidates; GATE_UP rows are staged and the consumer warps do GeGLU +
//     quantize_row_q8_0 per 32 outputs after their passes.  Three loop forms (chain_phase): single-segment phases
//     (a slot = one round), row-lane segment-driven (O / DOWN: lane = row, register accumulator, one round per K
//     segment or several), and the general loop (quad layouts, accumulators in shared memory, in-order commit).
//     Consumers and chain warps hand the term buffers over through mbarriers.
//   * NO GRID BARRIERS.  Every value that crosses CTAs travels as an 8-byte word {payload, tag} written with one
//     relaxed store (the NCCL "LL" idea): tag = the token's sequence number, buffers are per layer, a consumer
//     re-reads a word until its tag matches.  Data and flag arrive together: no fence, no separate wait.  A thread
//     requests all its words at once; if some are missing, the lanes that wait together spin on ONE shared word (a
//     spinning load is an L2 request that competes with the weight stream of the CTAs still working) and then re-read
//     only their own missing words (ll_gather).  Serial polls on a critical path cost a round trip each.
//   * ATTENTION = two phases without a combine: SCORE (items = kv head x 32-position chunk: kq in double, the new
//     position's K/V roped, rounded and stored) and PV (items = head x 16/32-dim slice: the WHOLE soft_max_ext row of
//     the head - global max, fp16 exp, exact double sum, fp16 P - then P.V in double).  Double accumulators make
//     the sums order-insensitive (1e-16), so no partial needs a second order of its own.
//   * TENSOR PARALLEL (one process per GPU).  Phases that carry the bytes (GATE_UP, DOWN, LOGITS) are row-sharded and
//     push their flagged outputs into every rank's buffers with peer stores over NVLink; the attention block is
//     sharded by KV head when the head counts divide (MHA: Q/K/V rows, cache, SCORE, PV per head, O rows) and is
//     REPLICATED otherwise (MQA: a few MB per layer cost less than a cross-GPU hop).  2 or 4 hops per layer.
//
// Phases per layer: QKV | SCORE | PV | O (+ residual) | GATE_UP (+ GeGLU + Q8_0 quantise) | DOWN (+ residual);
// then LOGITS (+ arg-max) | SAMPLE.
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace {

constexpr int kCW = 16;             // consumer warps
constexpr int kCT = kCW * 32;       // consumer threads
// warp roles: chain warps 0, 18, 19; consumers 1..16; producer 17.  The pass loop keeps the schedulers ~80 % busy, so
// the chain work (one warp's ~150 instructions per slot) is dealt slot by slot to three warps on three schedulers
constexpr int kProdWarp = 17, kC0 = 32;
constexpr int kNCH = 3;             // chain warps: warps 0, 18, 19 (one more warp on three of the four schedulers)
constexpr int kThreads = 20 * 32;
constexpr int kMaxBuf = 3;          // term buffers (as many as fit behind the phase's activation row)
constexpr int kSlots = 5;
constexpr int kMaxOutstanding = 2;  // ring copies in flight per SM (default; GGB_MAX_OUTSTANDING overrides)
constexpr int kPrefetchChunks = 0;  // L2 look-ahead of the weight stream in ring chunks (GGB_PREFETCH_CHUNKS overrides)
constexpr int kChunk = GGB_DEV_ATTN_CHUNK;
constexpr int kSeg = GGB_DEV_SEG_BLOCKS; // 64 blocks per K segment
constexpr int kMaxE = 4096;         // longest norm-fed activation row
constexpr int kMaxK = 24576;        // longest activation row kept in shared memory
constexpr int kTermRow = 68;        // floats per chain lane in a term buffer: 64 block terms + 4 of padding (bank skew)
constexpr int kTermFloats = 32 * kTermRow; // per term buffer: 16 passes x 128 block terms
constexpr int kAccFloats = 1792;    // row accumulators of multi-segment phases (rows of one CTA)
constexpr int kGuFloats = 512;      // finished gate / up rows of one CTA waiting for GeGLU
constexpr int kAccSmall = 64;       // the same for phases whose activation row covers the accumulator overlay (K > 8192)
constexpr int kMaxCtx = 8192;       // PV keeps one soft-max row in shared memory

enum PhaseKind { PH_QKV = 0, PH_SCORE, PH_PV, PH_O, PH_GATE_UP, PH_DOWN, PH_LOGITS, PH_SAMPLE };

template <int TYPE> struct Fmt {
    static constexpr int PB = TYPE == GGB_TYPE_Q4_0 ? 16 : 32;        // payload bytes per block
    static constexpr int ROW_PB = kSeg * PB;                           // payload bytes per row segment
    static constexpr int PASS_BYTES = 2 * kSeg * (PB + 2);             // 2304 / 4352
    static constexpr int PPS = TYPE == GGB_TYPE_Q4_0 ? 16 : 8;         // passes per ring slot
    static constexpr int SLOT_BYTES = PPS * PASS_BYTES;                // 36864 / 34816
};

typedef unsigned long long ll_t; // flagged word: low 32 bits payload, high 32 bits tag

struct Phase {
    int kind, layer;
    int n_units;     // schedulable units of THIS RANK (row pairs; quad: 4 rows; GATE_UP: 32-row groups of gate + up)
    int unit_passes; // passes per unit
    int unit_rows;   // logical rows per unit
    int K;           // activation length
    int n_lr;        // logical rows of this rank's stream (rows past it are padding)
    int rot;         // CTA rotation so remainders do not always land on the same SMs
    int uq, ur;      // n_units / grid, n_units % grid
    // pass geometry: a CTA's passes run K-segment-major: for seg < kA: its row pairs (plain passes: 2 rows x 64
    // blocks); then, quad layout only (K % 2048 in (0, 1024]: the last segment has <= 32 blocks), one pass per 4 rows
    // holding the short segment of all four (Gemma-7B: K = 3072)
    int quad, kf, ksegs, kA; // kA = quad ? kf : ksegs
    int tail_blocks;         // blocks of the last section: quad: nb - 64 kf; plain: nb - 64 (ksegs - 1)
    int u_off;               // first unit of this rank (GATE_UP output indexing)
    int bcast;               // flagged outputs of this phase go to every rank (tensor parallel)
    int resid_embed;         // residual = embedding row of the token (layer 0, O phase)
    // rank-local logical row -> output index: up to three runs (Q | K | V of a head-sharded QKV phase)
    int seg_rows[3], seg_off[3];
    int pad_;
    const uint8_t *stream; // pass stream of the phase (this rank's part)
    // flagged inputs: QKV / GATE_UP / LOGITS: fp32 x [E] (null in layer 0: embedding row); O: fp32 kqv [H*d];
    // DOWN: packed Q8_0 (4 codes per word) [F/4] then fp16 scales [F/32]; SCORE / PV: fp32 q | k | v
    const ll_t *in_ll;
    const float *norm_w;
    ll_t *out_ll;          // flagged output (QKV: qkv; O: sa; DOWN: next layer's x; GATE_UP: packed Q8_0; PV: kqv)
    float *out;            // plain fp32 copy for tensor_dump-style captures / logits (may be null)
    const ll_t *resid_ll;  // flagged residual
    float *f32_out;        // PV: kqv dump; GATE_UP: optional ffn_act dump; LOGITS: result_norm
    __half *kcache, *vcache;
    ll_t *s_ll;            // SCORE -> PV: kq scores of this layer [H][n_ctx]
};
static_assert(sizeof(Phase) % 4 == 0, "Phase is copied word-wise");
constexpr int kPhWords = (int)(sizeof(Phase) / 4);
static_assert(kPhWords <= kCT, "phase descriptor copied by one pass of the consumer threads");

struct DecoderParams {
    const Phase *phases;
    int n_phases;
    int type;
    int E, H, Hkv, d, F, V, n_ctx;
    // host-precomputed so that no phase starts with a chain of integer divisions
    int Gq, hsplit, Gh, hd_shift;
    int gq_shift;          // log2(query heads per KV head)
    int h0, nh, g0, ng, ng_shift; // attention heads / KV heads of this rank (all of them unless sharded by head)
    int pv_w, pv_oct_shift, pv_slices, pv_sl_shift; // PV: dims per item (16 | 32), log2(octets), slices per head
    float eps, q_scale, emb_scale;
    const uint8_t *emb_qs; // planar embedding table (layer-0 get_rows)
    const __half *emb_d;
    int32_t *state, *history;
    ll_t *cand_ll;         // [2][R * grid * kNCH]: arg-max candidates (value, index) of every chain warp of every rank
    int max_outstanding, prefetch_chunks; // producer tuning (see producer_loop)
    int rank, R;           // tensor parallel: this GPU's rank and the number of GPUs (1 = single GPU)
    long long peer_delta[8]; // byte offset from a local flagged word to the same word in rank r's memory (0 for self)
    uint32_t *seq;         // token sequence number = tag of the previous launch
    const uint16_t *exp_table, *gelu_table;
    int exp_compute;       // 1 / 2: exp_f16_computed / exp_f16_fast reproduce exp_table bit for bit (checked at create)
    int gelu_compute;      // 1: gelu_f16_computed reproduces gelu_table bit for bit (checked at create)
    const float *rope_cos, *rope_sin;
    long long *trace; // optional [n_phases][grid][4] %globaltimer stamps (ns): start, after prologue (inputs arrived), end, end
    long long *prof;  // PROF instantiation only: [grid][32] SM-clock sums inside phase prof_phase (see tools_trace.py)
    int prof_phase;
};

// union region: the views are never live at the same time
//   mat-vec : act_q [K] | act_d half[K/32] | bias int[K/32] | (K > 6144) terms float[2..3][kTermFloats]
//             (K <= 6144) at +8192: xs float[kMaxE] during the norm prologue, then acc float[kAccFloats] | gu float[kGuFloats] | terms float[3][kTermFloats]
//   score   : Kc half[2][32][256] | Sq double[4][256]
//   pv      : p float[n_ctx <= 8192] | part double[16][32]
constexpr int U_XS = 8192;
constexpr int U_ACC = U_XS, U_GU = U_ACC + kAccFloats * 4;
constexpr int U_KC = 0, U_SQ = 32768;
constexpr int U_PVP = 0;
constexpr int U_BYTES = kMaxK + kMaxK / 16 + kMaxK / 8 + 2 * kTermFloats * 4;
constexpr int kAccMaxK = 6144;      // activation rows up to this length (+ scales + biases) end below the overlay at U_ACC
// the phase's activation row: codes [K] | fp16 scales [K/32] | int biases [K/32]; then (large K) or behind the
// accumulator overlay (small K) the term buffers
__host__ __device__ __forceinline__ int act_d_off(int K) { return K; }
__host__ __device__ __forceinline__ int act_bias_off(int K) { return K + (K >> 4); }
__host__ __device__ __forceinline__ int terms_off(int K) {
    return K <= kAccMaxK ? U_XS + (kAccFloats + kGuFloats) * 4 : ((K + (K >> 4) + (K >> 3) + 127) & ~127);
}
__host__ __device__ __forceinline__ int terms_nbuf(int K) {
    const int n = (U_BYTES - terms_off(K)) / (kTermFloats * 4);
    return n < kMaxBuf ? n : kMaxBuf;
}
static_assert(U_SQ + 4 * 256 * 8 <= U_BYTES, "union region too small for the score view");
static_assert(kMaxCtx * 4 + 16 * 32 * 8 + kCT * 16 <= U_BYTES, "union region too small for the pv view");
static_assert(kMaxE <= U_XS && kMaxE <= kAccMaxK, "xs overlaps the activation row");
static_assert(kAccMaxK + kAccMaxK / 16 + kAccMaxK / 8 <= U_XS, "small activation rows must end below the overlay");

template <int TYPE> struct Smem {
    static constexpr int SLOTS = kSlots;
    static constexpr int U_SIZE = U_BYTES;
    alignas(1024) uint8_t ring[SLOTS][Fmt<TYPE>::SLOT_BYTES];
    alignas(128) uint8_t u[U_SIZE];
    double red[kCW];
    float redf[kCW];
    float acc_small[kAccSmall];
    alignas(8) unsigned long long full_bar[kSlots];
    alignas(8) unsigned long long empty_bar[kSlots];
    alignas(8) unsigned long long kv_bar[2];  // SCORE: the K chunk rows of buffer 0 / 1 have landed
    alignas(8) unsigned long long tfull[kMaxBuf];   // term buffer b holds a whole slot's terms (16 consumer-warp arrivals)
    alignas(8) unsigned long long tempty[kMaxBuf];  // a chain warp has consumed term buffer b
    volatile int done_slot; // slots (numbered through the token) whose rows are all chained, committed in order
    long long tr[2]; // trace stamps of the current phase (thread 0)
    int wst[kCW + kNCH][5]; // per-warp state that lives across phases: ring slot, ring parity, term slot counter, K-buffer parities
    int sched[6];    // [0] = chunks, [3] = token, [4] = number of positions (kept out of registers: the kernel runs
                     // at the 96-register limit of 18 warps per SM and a spill costs an L2 round trip)
    alignas(16) Phase ph[2];
};

extern __shared__ __align__(1024) uint8_t ggb_dec_smem[];

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// GGB_WATCHDOG: a wait that does not end (a lost flag, a mis-counted barrier) traps instead of hanging the GPU: the
// launch then fails with an error the host reports.  The bounds are seconds of spinning, far beyond any legal wait.
#ifndef GGB_WATCHDOG
#define GGB_WATCHDOG 0
#endif
// try_wait with a suspend-time hint: the waiting warp sleeps in hardware until the phase completes (or the hint runs
// out) instead of re-issuing the poll - sixteen consumer warps polling a term-buffer barrier took two thirds of the
// issue slots away from the chain warp they were waiting for.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
#if GGB_WATCHDOG
    for (uint32_t spins = 0; !done; spins++) {
#else
    while (!done) {
#endif
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(1000000u)
            : "memory");
#if GGB_WATCHDOG
        if (spins > (1u << 22)) __trap();
#endif
    }
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier.  The weight stream is read once
// per token and is 11x the L2: it carries an evict-first policy so that it does not push the small hot data
// (exp / gelu tables, flagged buffers, norm weights, KV cache) out of L2 between two layers.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_plain(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// small async copy global -> shared (the issuing warp does not wait for the data)
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
// orders this thread's (and, through a preceding barrier, the CTA's) generic-proxy accesses before later async-proxy ones
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kCT) : "memory"); }
__device__ __forceinline__ void phase_bar() { asm volatile("bar.sync 2, %0;" ::"n"(kCT + 32 * kNCH) : "memory"); } // consumers + chain warps
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ float lds_half(uint32_t addr) {
    unsigned short h;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(addr));
    return __half2float(__ushort_as_half(h));
}
__device__ __forceinline__ int lds_s32(uint32_t addr) {
    int r;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(r) : "r"(addr));
    return r;
}
__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c) { // unsigned bytes x signed bytes
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// (the fp16 exp / gelu tables are pinned in L2 by the launch's access-policy window: under the weight stream a demand
// miss to HBM costs ~1 us of queueing, and the lookups sit on every layer's critical path)

// exp with the reference's fp16-table semantics, table[h] = fp16(expf(fp32(h))), COMPUTED: exp() in double, rounded
// to fp32, then to fp16.  ggb_decoder_create checks this against the uploaded table for every non-NaN fp16 input
// and only then sets P.exp_compute; the lookup it replaces is an L2 round trip on the critical path of ATTN and
// COMBINE in every layer.
__device__ __forceinline__ float exp_f16_computed(float x_f16_valued) {
    return half_bits_to_float(float_to_half_bits(__double2float_rn(exp((double)x_f16_valued))));
}
// the cheap form: fp32 expf (a few ulp of fp32), rounded to fp16 - identical to the table only if no input lands within
// that error of an fp16 rounding boundary, which the same exhaustive check decides (exp_compute == 2)
__device__ __forceinline__ float exp_f16_fast(float x_f16_valued) {
    return half_bits_to_float(float_to_half_bits(expf(x_f16_valued)));
}
__device__ __forceinline__ float exp_f16(const DecoderParams &P, float x) {
    const uint16_t hb = float_to_half_bits(x);
    if (!P.exp_compute || (hb & 0x7FFFu) > 0x7C00u) return half_bits_to_float(__ldg(P.exp_table + hb)); // no proof / NaN
    if (P.exp_compute == 2) return exp_f16_fast(half_bits_to_float(hb));
    return exp_f16_computed(half_bits_to_float(hb));
}
// gelu with the reference's fp16-table semantics, table[h] = fp16(gelu_f32(fp32(h))) ([UPSTREAM] ggml_gelu_f32: the tanh
// approximation in fp32), COMPUTED with the device tanhf and every operator rounded as the host rounds it; used only
// after the same exhaustive check (the lookup it replaces is an L2 round trip inside the chain warps' GeGLU)
__device__ __forceinline__ uint16_t gelu_f16_computed(float x) {
    const float A = 0.044715f, S = 0.79788456080286535587989211986876f;
    const float arg = __fmul_rn(__fmul_rn(S, x), __fadd_rn(1.0f, __fmul_rn(__fmul_rn(A, x), x)));
    return float_to_half_bits(__fmul_rn(__fmul_rn(0.5f, x), __fadd_rn(1.0f, tanhf(arg))));
}
__global__ void exp_check_kernel(const uint16_t *table, const uint16_t *gelu_table, int *mismatches) {
    // [0]: double exp, [1]: fp32 exp, [2]: gelu
    const uint32_t hb = blockIdx.x * blockDim.x + threadIdx.x;
    if (hb >= 65536u || (hb & 0x7FFFu) > 0x7C00u) return;
    const float x = half_bits_to_float((uint16_t)hb);
    if (float_to_half_bits(exp_f16_computed(x)) != table[hb]) atomicAdd(mismatches, 1);
    if (float_to_half_bits(exp_f16_fast(x)) != table[hb]) atomicAdd(mismatches + 1, 1);
    if (gelu_f16_computed(x) != gelu_table[hb]) atomicAdd(mismatches + 2, 1);
}

// ---- flagged words.  An aligned 8-byte access is single-copy atomic, so payload and tag are seen together;
// relaxed gpu-scope accesses go to L2 (never a stale L1 line).  ll_get = the word if its tag matches, else poll.
// bcast (outputs of a row-sharded phase, tensor parallel): the word is also written to every peer's copy of the
// buffer (one-shot push all-gather over NVLink peer stores, the consumer polls its local memory).
__device__ __forceinline__ void ll_put(const DecoderParams &pe, bool bcast, ll_t *p, uint32_t payload, uint32_t tag) {
    const ll_t v = ((ll_t)tag << 32) | payload;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    if (bcast) {
#pragma unroll 1
        for (int r = 1; r < pe.R; r++) { // peer_delta[0] is this rank (0); entries 1 .. R-1 are the peer mappings
            const char *q = reinterpret_cast<const char *>(p) + pe.peer_delta[r];
            asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(q), "l"(v) : "memory");
        }
    }
}
__device__ __forceinline__ void ll_put_f(const DecoderParams &pe, bool bcast, ll_t *p, float x, uint32_t tag) {
    ll_put(pe, bcast, p, __float_as_uint(x), tag);
}
__device__ __forceinline__ ll_t ll_peek(const ll_t *p) {
    ll_t v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ll_get(ll_t v, const ll_t *p, uint32_t tag) {
#if GGB_WATCHDOG
    for (uint32_t spins = 0; (uint32_t)(v >> 32) != tag; spins++) {
        v = ll_peek(p);
        if (spins > (1u << 24)) __trap();
    }
#else
    while ((uint32_t)(v >> 32) != tag) v = ll_peek(p);
#endif
    return (uint32_t)v;
}
__device__ __forceinline__ float ll_get_f(ll_t v, const ll_t *p, uint32_t tag) { return __uint_as_float(ll_get(v, p, tag)); }
// A thread's N flagged words p[0], p[stride], ... (the first n of them), requested together.  When the data is there this
// is ONE L2 round trip.  When it is not, polling the words one after the other (ll_get on each) would cost a round trip
// PER WORD after the data has arrived - every stale word has to be re-read, serially; instead the thread spins on one
// missing word and then re-reads all of them at once: the wait ends at most two round trips after the last store.
template <int N>
__device__ __forceinline__ void ll_gather(ll_t (&v)[N], const ll_t *p, int stride, int n, uint32_t tag) {
#pragma unroll
    for (int u = 0; u < N; u++) v[u] = u < n ? ll_peek(p + (size_t)stride * u) : 0ull;
    // The lanes that called together wait together: while anything is missing they all spin on ONE word (the first
    // missing word of the lowest such lane - 32 loads of one address are one L2 request, where 32 private spins were 32
    // requests per round trip, from every warp of every waiting CTA, competing with the weight stream of the CTAs that
    // are still working), then every lane re-reads only its own missing words, together.
    const unsigned am = __activemask();
    for (;;) {
        int bad = -1;
#pragma unroll
        for (int u = N - 1; u >= 0; u--) bad = (u < n && (uint32_t)(v[u] >> 32) != tag) ? u : bad;
        const unsigned need = __ballot_sync(am, bad >= 0);
        if (!need) return;
        const ll_t *mine = p + (size_t)stride * (bad < 0 ? 0 : bad);
        const ll_t *pb = reinterpret_cast<const ll_t *>(__shfl_sync(am, (unsigned long long)mine, __ffs(need) - 1));
#if GGB_WATCHDOG
        for (uint32_t spins = 0; (uint32_t)(ll_peek(pb) >> 32) != tag; spins++)
            if (spins > (1u << 24)) __trap();
#else
        while ((uint32_t)(ll_peek(pb) >> 32) != tag) {}
#endif
#pragma unroll
        for (int u = 0; u < N; u++) // only the words that were missing (the others would be N reads per round for nothing)
            if (u < n && (uint32_t)(v[u] >> 32) != tag) v[u] = ll_peek(p + (size_t)stride * u);
    }
}
__device__ __forceinline__ unsigned long long gtime() { // ns, one clock for the whole GPU (SM clocks are not aligned)
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
struct Best {
    float v;
    int i;
};
__device__ __forceinline__ Best better(Best a, Best b) { // first maximum (ref: src/gemma_model.cpp:538-543)
    if (b.i < 0) return a;
    if (a.i < 0) return b;
    if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
}

// CTA c's share [u0, u1) of n units: contiguous and balanced (q = n / G, r = n % G come from the host and
// rot < G: no division here, a 64-bit one costs 0.2 us on the critical path of every phase)
__device__ __forceinline__ void unit_range(int cta, int G, int rot, int q, int r, int &u0, int &u1) {
    int c = cta + rot;
    if (c >= G) c -= G;
    u0 = c * q + min(c, r);
    u1 = u0 + q + (c < r ? 1 : 0);
}

__device__ __forceinline__ bool is_matvec(int kind) {
    return kind == PH_QKV || kind == PH_O || kind == PH_GATE_UP || kind == PH_DOWN || kind == PH_LOGITS;
}

// get_rows(token_embd, token)[i] * sqrt(n_embd)   (ref: src/gemma_model.cpp:675-679)
template <int TYPE> __device__ __forceinline__ float embed_value(const DecoderParams &P, int token, int i) {
    const size_t blk = (size_t)token * (P.E / 32) + (i >> 5);
    const int e = i & 31;
    const float dd = __half2float(P.emb_d[blk]);
    int qv;
    if (TYPE == GGB_TYPE_Q4_0) {
        const int byte = P.emb_qs[blk * 16 + (e & 15)];
        qv = (e < 16 ? (byte & 0x0F) : (byte >> 4)) - 8;
    } else {
        qv = ((const int8_t *)P.emb_qs)[blk * 32 + e];
    }
    return __fmul_rn(__fmul_rn((float)qv, dd), P.emb_scale);
}

// ---------------------------------------------------------------- producer
// The weight stream of this CTA, flattened over the token: chunk after chunk of <= PPS passes, never across a phase.
template <int TYPE> struct ChunkWalk {
    const DecoderParams &P;
    int cta, G, pi, n;
    const uint8_t *src;
    __device__ __forceinline__ ChunkWalk(const DecoderParams &P_, int cta_, int G_) : P(P_), cta(cta_), G(G_), pi(-1), n(0), src(nullptr) {}
    // next chunk: false when the token's stream is exhausted
    __device__ __forceinline__ bool next(const uint8_t *&p, uint32_t &bytes) {
        using F = Fmt<TYPE>;
        while (n <= 0) {
            if (++pi >= P.n_phases) return false;
            const Phase *ph = P.phases + pi;
            if (!is_matvec(ph->kind)) continue;
            int u0, u1;
            unit_range(cta, G, ph->rot, ph->uq, ph->ur, u0, u1);
            const int up = ph->unit_passes;
            n = (u1 - u0) * up;
            src = ph->stream + (size_t)u0 * up * F::PASS_BYTES;
        }
        const int cnt = min(n, F::PPS);
        bytes = (uint32_t)cnt * F::PASS_BYTES;
        p = src;
        src += bytes;
        n -= cnt;
        return true;
    }
};

// Two walkers over the same chunk sequence.  The LOADER fills the shared-memory ring (cp.async.bulk, completion
// on the slot's mbarrier).  The PREFETCHER runs P.prefetch_chunks chunks ahead of it and only asks L2 to fetch
// (cp.async.bulk.prefetch.L2): the ring (180 KB per SM) is full during the latency-bound phases (ATTN, COMBINE,
// prologues) and HBM would idle; with the look-ahead in L2 the stream keeps flowing and GATE_UP / DOWN find their
// tail in L2 instead of waiting for HBM.  The look-ahead is bounded (chunks x 36 KB x 148 SMs, well inside the
// 126 MB L2), and the later ring load marks the lines evict-first, so a consumed chunk is the first to leave.
template <int TYPE>
__device__ __forceinline__ void producer_loop(const DecoderParams &P, Smem<TYPE> &S, int cta, int G) {
    using F = Fmt<TYPE>;
    constexpr int SLOTS = Smem<TYPE>::SLOTS;
    const uint32_t ring = smem_u32(&S.ring[0][0]);
    const uint32_t full0 = smem_u32(&S.full_bar[0]), empty0 = smem_u32(&S.empty_bar[0]);
    int slot = 0, issued = 0;
    uint32_t parity = 0;
    const uint64_t pol = l2_evict_first_policy();
    const int max_out = P.max_outstanding;
    ChunkWalk<TYPE> load(P, cta, G), pre(P, cta, G);
    const uint8_t *src, *psrc;
    uint32_t bytes, pbytes;
    bool more = P.prefetch_chunks > 0;
    for (int i = 0; i < P.prefetch_chunks && more; i++) {
        more = pre.next(psrc, pbytes);
        if (more && i >= SLOTS) bulk_prefetch_l2(psrc, pbytes); // the first kSlots chunks go straight to the ring
    }
#pragma unroll 1
    while (load.next(src, bytes)) {
        if (more) {
            more = pre.next(psrc, pbytes);
            if (more) bulk_prefetch_l2(psrc, pbytes);
        }
        mbar_wait(empty0 + 8 * slot, parity ^ 1u);
        // at most max_out copies in flight per SM: enough to cover the HBM latency at full rate, and it bounds the
        // queueing delay every small latency-critical access (flag polls, table lookups) sees
        if (issued >= max_out) {
            const int old = issued - max_out;
            mbar_wait(full0 + 8 * (old % SLOTS), (uint32_t)((old / SLOTS) & 1));
        }
#ifdef GGB_KO_STREAM // knock-out build (timing only, wrong results): the slot "arrives" without a copy
        mbar_expect_tx(full0 + 8 * slot, 0u);
#else
        mbar_expect_tx(full0 + 8 * slot, bytes);
        bulk_g2s(ring + slot * F::SLOT_BYTES, src, bytes, full0 + 8 * slot, pol);
#endif
        issued++;
        if (++slot == SLOTS) {
            slot = 0;
            parity ^= 1u;
        }
    }
}

// ---------------------------------------------------------------- prologues (each ends with the activation row in
// shared memory: act_q int8[K], act_d fp16[K/32], bias int[K/32] = -8 * sum of the block's codes)

// quantize_row_q8_0 of FOUR consecutive 32-element blocks per warp: lane l holds elements 4l .. 4l+3 of the
// 128, so 8 lanes share a block.  The serial instruction chain (two IEEE divisions, roundf) is what a block
// costs (~0.3 us at 4 warps per scheduler), so a warp runs it once per four blocks instead of once per block.
// amax / code sum: three xor-shuffles inside each 8-lane group (a sub-warp REDUX serialises the four groups
// behind WARPSYNC.EXCLUSIVE and costs ~0.5 us).  max is exact: same value as the reference's scan.
// (activation codes live in TWO planes - bytes 0..15 of every block, then bytes 16..31 at + hi_off = K / 2 - so that
// load_act_regs' LDS.128 are 16-byte-strided across the warp: conflict-free)
__device__ __forceinline__ void quant_quad_to_smem(uint32_t u, int quad, int lane, const float (&v)[4], int hi_off, int K) {
    float amax = fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3])));
#pragma unroll
    for (int off = 1; off <= 4; off <<= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off)); // 8-lane groups
    const float d = __fdiv_rn(amax, 127.0f);
    const float id = d != 0.0f ? __fdiv_rn(1.0f, d) : 0.0f;
    uint32_t packed = 0;
#pragma unroll
    for (int e = 0; e < 4; e++) packed |= ((uint32_t)(int)roundf(__fmul_rn(v[e], id)) & 0xFFu) << (8 * e);
    // word j = lane & 7 of block b = 4 quad + (lane >> 3): plane j >> 2, offset 16 b + 4 (j & 3)
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(u + ((lane & 4) ? hi_off : 0) + 64 * quad + 16 * (lane >> 3) + 4 * (lane & 3)), "r"(packed) : "memory");
    int qs = __dp4a((int)packed, 0x01010101, 0);
#pragma unroll
    for (int off = 1; off <= 4; off <<= 1) qs += __shfl_xor_sync(0xffffffffu, qs, off);
    if ((lane & 7) == 0) {
        const int b = 4 * quad + (lane >> 3);
        const unsigned short hd = __half_as_ushort(__float2half_rn(d));
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(u + act_d_off(K) + 2 * b), "h"(hd) : "memory");
        asm volatile("st.shared.s32 [%0], %1;" ::"r"(u + act_bias_off(K) + 4 * b), "r"(-8 * qs) : "memory");
    }
}

// rms_norm (+ weight mul) + quantize_row_q8_0 of the phase input; every CTA recomputes it (16 KB of flagged
// words from L2) instead of making it a phase of its own.  Sum of squares ([UPSTREAM] ggml_float = double, of the
// fp32-rounded products x * x): the sixteen warps share the row - thread (warp w, lane l) takes the lower (w < 8) or
// upper half of the elements 32 * (w & 7) + l + 256 k -, a butterfly per warp, the sixteen warp sums in order.  A
// double sum of these terms does not depend on its grouping at the precision the rounding to fp32 keeps, so the result
// is the scalar loop's (bitwise in every test); the warps then split the quantisation by blocks.
template <int TYPE>
__device__ __forceinline__ void norm_prologue(const DecoderParams &P, Smem<TYPE> &S, const Phase &ph, int token,
                                              uint32_t tag, float *f32_out) {
    const int tid = (int)threadIdx.x - kC0, lane = tid & 31, warp = tid >> 5;
    const int E = P.E, nk = E / GGB_DEV_NORM_THREADS;
    const int base = 32 * (warp & 7) + lane;
    float *xs = reinterpret_cast<float *>(S.u + U_XS);
    const ll_t *in = ph.in_ll;
    // the norm weights of this warp's first quad of blocks: requested before the (long) wait for x
    const float *norm_w = ph.norm_w;
    const int nquad = E / 128;
    float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (warp < nquad) w4 = *reinterpret_cast<const float4 *>(norm_w + 128 * warp + 4 * lane);
    double acc = 0.0;
    // the sixteen warps share the row: warp half h takes the upper or lower half of a thread column's words (both halves
    // used to gather ALL of them: twice the polling traffic for nothing).  A thread's words are requested together
    // (all of them for E <= 2048): ONE L2 round trip when x is there
    constexpr int NV = 4;
    const int nk_half = (nk + 1) >> 1, k_lo = (warp >> 3) * nk_half, k_hi = min(nk, k_lo + nk_half);
    if (in) {
#pragma unroll 1
        for (int k0 = k_lo; k0 < k_hi; k0 += NV) {
            ll_t v[NV];
            const ll_t *pw = in + GGB_DEV_NORM_THREADS * k0 + base;
            ll_gather<NV>(v, pw, GGB_DEV_NORM_THREADS, k_hi - k0, tag);
#pragma unroll
            for (int u = 0; u < NV; u++) {
                if (k0 + u < k_hi) {
                    const float x = __uint_as_float((uint32_t)v[u]);
                    xs[GGB_DEV_NORM_THREADS * (k0 + u) + base] = x;
                    acc = __dadd_rn(acc, (double)__fmul_rn(x, x));
                }
            }
        }
    } else { // layer 0: the embedding row of the token
#pragma unroll 1
        for (int k = k_lo; k < k_hi; k++) {
            const int i = GGB_DEV_NORM_THREADS * k + base;
            const float x = embed_value<TYPE>(P, token, i);
            xs[i] = x;
            acc = __dadd_rn(acc, (double)__fmul_rn(x, x));
        }
    }
    acc = warp_sum_f64(acc);
    if (lane == 0) S.red[warp] = acc;
    consumer_bar();
    double sum = S.red[0];
#pragma unroll
    for (int i = 1; i < kCW; i++) sum = __dadd_rn(sum, S.red[i]);
    // sum / E: a power-of-two E scales exactly (same bits as the division, which is a ~600-cycle routine)
    const float mean = __double2float_rn((E & (E - 1)) == 0 ? __dmul_rn(sum, 1.0 / (double)E) : __ddiv_rn(sum, (double)E));
    const float scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, P.eps)));
    const uint32_t ua = smem_u32(S.u);
#pragma unroll 1
    for (int quad = warp; quad < nquad; quad += kCW) { // elements 128 quad + 4 lane .. + 3 (staged by other threads: bar above)
        const int i = 128 * quad + 4 * lane;
        const float4 x4 = *reinterpret_cast<const float4 *>(xs + i);
        if (quad != warp) w4 = *reinterpret_cast<const float4 *>(norm_w + i);
        float v[4];
        v[0] = __fmul_rn(__fmul_rn(x4.x, scale), w4.x);
        v[1] = __fmul_rn(__fmul_rn(x4.y, scale), w4.y);
        v[2] = __fmul_rn(__fmul_rn(x4.z, scale), w4.z);
        v[3] = __fmul_rn(__fmul_rn(x4.w, scale), w4.w);
        if (f32_out) *reinterpret_cast<float4 *>(f32_out + i) = make_float4(v[0], v[1], v[2], v[3]);
        quant_quad_to_smem(ua, quad, lane, v, E / 2, E);
    }
    consumer_bar();
}

// O: the flagged fp32 row written by COMBINE -> quantize_row_q8_0 (ggml mul_mat INIT stage)
template <int TYPE>
__device__ __forceinline__ void quant_prologue(Smem<TYPE> &S, const ll_t *in, int K, uint32_t tag) {
    const int lane = threadIdx.x & 31, warp = ((int)threadIdx.x - kC0) >> 5;
    const uint32_t ua = smem_u32(S.u);
#pragma unroll 1
    for (int quad = warp; quad < K / 128; quad += kCW) {
        const ll_t *p = in + 128 * quad + 4 * lane;
        ll_t w[4];
        ll_gather<4>(w, p, 1, 4, tag);
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; e++) v[e] = __uint_as_float((uint32_t)w[e]);
        quant_quad_to_smem(ua, quad, lane, v, K / 2, K);
    }
    consumer_bar();
}

// DOWN: packed Q8_0 row written by GATE_UP: 4 codes per flagged word, then one word per block with the fp16 scale
// (low half) and the sum of the block's codes (high half, int16: |sum| <= 32 * 127), from which the Q4_0 bias follows
template <int TYPE>
__device__ __forceinline__ void packed_prologue(Smem<TYPE> &S, const ll_t *in, int K, uint32_t tag) {
    const int tid = (int)threadIdx.x - kC0;
    const int nw = K / 4, ns = K / 32;
    uint32_t *aq = reinterpret_cast<uint32_t *>(S.u);
    __half *ad = reinterpret_cast<__half *>(S.u + act_d_off(K));
    int *bias = reinterpret_cast<int *>(S.u + act_bias_off(K));
    constexpr int NW = 5; // words in flight per thread (9 = a whole Gemma-2B DOWN row at once measured 9-14 % SLOWER per token, before and after the polling fixes; 4 and 5 tie)
#pragma unroll 1
    for (int i0 = tid; i0 < nw + ns; i0 += NW * kCT) {
        ll_t v[NW];
        ll_gather<NW>(v, in + i0, kCT, (nw + ns - i0 + kCT - 1) / kCT, tag);
#pragma unroll
        for (int u = 0; u < NW; u++) {
            const int i = i0 + kCT * u;
            if (i < nw + ns) {
                const uint32_t w = (uint32_t)v[u];
                if (i < nw) aq[((i & 4) ? nw / 2 : 0) + 4 * (i >> 3) + (i & 3)] = w; // two planes (see quant_quad_to_smem)
                else {
                    ad[i - nw] = __ushort_as_half((unsigned short)w);
                    if (TYPE == GGB_TYPE_Q4_0) bias[i - nw] = -8 * ((int)w >> 16);
                }
            }
        }
    }
    consumer_bar();
}

// ---------------------------------------------------------------- one pass: 2 rows x one K segment -> 128 block terms
// Lane l owns blocks l and l + 32 of the segment for both rows.  Per-block term = the scalar reference's
//   Q4_0: ((float)sumi * d_x) * d_y      Q8_0: (float)sumi * (d_x * d_y)
// ([UPSTREAM] ggml_vec_dot_q4_0_q8_0 / ggml_vec_dot_q8_0_q8_0 scalar branches).  Nothing is added here: the four
// terms of a lane go to the term table and the chain warp adds them in block order.
struct ActRegs { // the codes of the lane's two activation blocks of one K segment
    int a[2][8];
};
template <int TYPE>
__device__ __forceinline__ void load_act_regs(ActRegs &A, uint32_t act, int seg, int nb, int lane) {
#pragma unroll
    for (int t = 0; t < 2; t++) {
        const int b = seg * kSeg + lane + 32 * t;
        const int bc = b < nb ? b : 0; // blocks past the row: any valid address (their terms are never read)
        const uint4 lo = lds_u4(act + bc * 16), hi = lds_u4(act + nb * 16 + bc * 16); // planes: bytes 0..15 | 16..31
        A.a[t][0] = lo.x; A.a[t][1] = lo.y; A.a[t][2] = lo.z; A.a[t][3] = lo.w;
        A.a[t][4] = hi.x; A.a[t][5] = hi.y; A.a[t][6] = hi.z; A.a[t][7] = hi.w;
    }
}
// QUADPASS: the pass holds the short last segment (<= 32 blocks) of FOUR rows; lane l owns block l of each row.
// The byte layout of a pass is the same in both cases: four groups of 32 blocks, payloads then scales; group k is
// (row k >> 1, blocks 32 (k & 1) ..) of a plain pass and row k of a quad pass.
// b0 = seg * 64 + lane: the lane's first activation block.  Its scale and Q4_0 bias (and those of block b0 + 32) are
// re-read from shared memory every pass - requested first, used last - instead of living in registers: the kernel
// runs at the 96-register limit and a spilled value would come back from L2 in the middle of the dp4a chain.
// (A block index past the row reads some other shared-memory word; its term is never used.)
template <int TYPE, bool QUADPASS>
__device__ __forceinline__ void pass_terms(uint32_t base, uint32_t act_d, uint32_t act_bias, int b0, const ActRegs &A, int lane,
                                           float (&term)[4]) {
    using F = Fmt<TYPE>;
    const int(&a)[2][8] = A.a;
    float dy[2];
    int bias[2];
#pragma unroll
    for (int t = 0; t < (QUADPASS ? 1 : 2); t++) {
        dy[t] = lds_half(act_d + 2 * (b0 + 32 * t));
        bias[t] = TYPE == GGB_TYPE_Q4_0 ? lds_s32(act_bias + 4 * (b0 + 32 * t)) : 0;
    }
    if (QUADPASS) {
        dy[1] = dy[0];
        bias[1] = bias[0];
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int t = QUADPASS ? 0 : (k & 1);
        // Q8_0 groups hold the first 16 codes of their 32 blocks, then the second 16 (two planes of 512 B): both
        // LDS.128 of a lane are then 16-byte-strided across the warp, i.e. conflict-free
        const uint32_t pay = base + k * (32 * F::PB) + lane * 16;
        const float dx = lds_half(base + 2 * F::ROW_PB + k * 64 + lane * 2);
        if (TYPE == GGB_TYPE_Q4_0) {
            const uint4 w = lds_u4(pay);
            const uint32_t q[4] = {w.x, w.y, w.z, w.w};
            int lo = 0, hi = 0; // high nibbles stay in place (x16) and are divided out once: exact
#pragma unroll
            for (int i = 0; i < 4; i++) {
                lo = __dp4a((int)(q[i] & 0x0F0F0F0Fu), a[t][i], lo);
                hi = dp4a_us(q[i] & 0xF0F0F0F0u, a[t][4 + i], hi);
            }
            const int sumi = lo + (hi >> 4) + bias[t];
            term[k] = __fmul_rn(__fmul_rn((float)sumi, dx), dy[t]);
        } else {
            const uint4 w0 = lds_u4(pay), w1 = lds_u4(pay + 512);
            int s0 = 0, s1 = 0;
            s0 = __dp4a((int)w0.x, a[t][0], s0); s1 = __dp4a((int)w1.x, a[t][4], s1);
            s0 = __dp4a((int)w0.y, a[t][1], s0); s1 = __dp4a((int)w1.y, a[t][5], s1);
            s0 = __dp4a((int)w0.z, a[t][2], s0); s1 = __dp4a((int)w1.z, a[t][6], s1);
            s0 = __dp4a((int)w0.w, a[t][3], s0); s1 = __dp4a((int)w1.w, a[t][7], s1);
            term[k] = __fmul_rn((float)(s0 + s1), __fmul_rn(dx, dy[t]));
        }
    }
}
// Term table of one slot: pass j of the slot, group k, lane l -> chain lane c = 2 j + (k >> 1), block b = l + 32 (k & 1)
// of that lane's 64-term run, stored at [c][b] with rows of 68 floats: the consumers' stores (32 lanes = 32 consecutive
// blocks of one c) and the chain warp's 128-bit loads (lane stride 68 words = 4 banks: quarter-warps cover all 32 banks)
// are both conflict-free.
__device__ __forceinline__ void store_terms(uint32_t tbase, int j, int lane, const float (&term)[4]) {
    const uint32_t r0 = tbase + 4u * (uint32_t)(2 * j * kTermRow + lane);
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(r0), "f"(term[0]) : "memory");
    asm volatile("st.shared.f32 [%0+128], %1;" ::"r"(r0), "f"(term[1]) : "memory");
    asm volatile("st.shared.f32 [%0+272], %1;" ::"r"(r0), "f"(term[2]) : "memory");
    asm volatile("st.shared.f32 [%0+400], %1;" ::"r"(r0), "f"(term[3]) : "memory");
}
static_assert(kTermRow * 4 == 272, "store_terms offsets");
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
    return r;
}
// acc + term[b0] + term[b0 + 1] + ... (len terms, one by one: the reference's  sumf += term  loop); row = shared address
// of the lane's 64-term run; len is the same for every lane of the warp
// acc + term[b0] + term[b0 + 1] + ... (len terms, one by one: the reference's  sumf += term  loop); row = shared address
// of the lane's 64-term run; len is the same for every lane of the warp
__device__ __forceinline__ float chain_add(uint32_t row, int b0, int len, float acc) {
#ifdef GGB_KO_FADD // knock-out build (timing only, wrong results): one add per round instead of len
    { const float4 t0 = lds_f4(row); return __fadd_rn(acc, t0.x); }
#endif
    if (len == 64) { // the common case, fully unrolled: 16 loads up front, then the 64 dependent adds
        float4 t[16];
#pragma unroll
        for (int u = 0; u < 16; u++) t[u] = lds_f4(row + 16u * (uint32_t)u);
#pragma unroll
        for (int u = 0; u < 16; u++) {
            acc = __fadd_rn(acc, t[u].x);
            acc = __fadd_rn(acc, t[u].y);
            acc = __fadd_rn(acc, t[u].z);
            acc = __fadd_rn(acc, t[u].w);
        }
        return acc;
    }
#pragma unroll 1
    for (int i0 = 0; i0 < len; i0 += 16) {
        float4 t[4];
#pragma unroll
        for (int u = 0; u < 4; u++) t[u] = lds_f4(row + 4u * (uint32_t)(b0 + i0) + 16u * (uint32_t)u);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (i0 + 4 * u + 0 < len) acc = __fadd_rn(acc, t[u].x);
            if (i0 + 4 * u + 1 < len) acc = __fadd_rn(acc, t[u].y);
            if (i0 + 4 * u + 2 < len) acc = __fadd_rn(acc, t[u].z);
            if (i0 + 4 * u + 3 < len) acc = __fadd_rn(acc, t[u].w);
        }
    }
    return acc;
}


// GeGLU (fp16 gelu table) + quantize_row_q8_0 of a 32-row group whose gate and up rows are finished (gu[0..31] = gate
// rows, gu[32..63] = up rows): lane = row of the group.  A lane's code joins three neighbours' in one flagged word of
// the packed row; the block's fp16 scale and the sum of its codes (the Q4_0 bias of DOWN's dot products) share one more.
__device__ __forceinline__ void geglu_unit(const DecoderParams &P, const float *gu, int ug, ll_t *out_ll, float *act_dump,
                                           bool bcast, uint32_t tag, int lane) {
    const uint16_t gb = float_to_half_bits(gu[lane]);
    const float gl = half_bits_to_float((P.gelu_compute && (gb & 0x7FFFu) <= 0x7C00u) ? gelu_f16_computed(half_bits_to_float(gb))
                                                                                    : __ldg(P.gelu_table + gb));
    const float x = __fmul_rn(gl, gu[32 + lane]);
    if (act_dump) act_dump[(size_t)ug * 32 + lane] = x;
    const float amax = warp_max_f32(fabsf(x));
    const float dd = __fdiv_rn(amax, 127.0f);
    const float id = dd != 0.0f ? __fdiv_rn(1.0f, dd) : 0.0f;
    const int qi = (int)roundf(__fmul_rn(x, id));
    uint32_t w = (uint32_t)qi & 0xFFu;
    w |= (__shfl_down_sync(0xffffffffu, w, 1) & 0xFFu) << 8;
    w |= (__shfl_down_sync(0xffffffffu, w, 2) & 0xFFFFu) << 16;
    if ((lane & 3) == 0) ll_put(P, bcast, out_ll + (size_t)ug * 8 + (lane >> 2), w, tag);
    const int qs = __reduce_add_sync(0xffffffffu, qi);
    if (lane == 0)
        ll_put(P, bcast, out_ll + P.F / 4 + ug, ((uint32_t)qs << 16) | (uint32_t)__half_as_ushort(__float2half_rn(dd)), tag);
}

// ---------------------------------------------------------------- a chain warp's share of a mat-vec phase
// The slots of a CTA's share go round-robin to the kNCH chain warps.  For its slots a chain warp waits for the consumers'
// terms, adds them to the rows' accumulators in block order (rounds = the passes of one K segment inside the slot: a
// row occurs once per round) and finishes the rows whose last segment this was.  Rows that continue an accumulator
// from an earlier slot (multi-segment phases) wait until every earlier slot is committed; slots commit in order
// (S.done_slot), and the warp that commits the slot which completes a 32-row GATE_UP group does its GeGLU + quantize.
template <int TYPE, bool QUAD, bool PROF>
__device__ __forceinline__ void chain_phase(const DecoderParams &P, int phi, int ci, int cta, int G, uint32_t tag, bool prof) {
    long long pt0 = 0, p_wait = 0, p_chain = 0, p_fin = 0, p_rounds = 0, p_begin = 0, p_add = 0, p_row = 0, pta = 0;
    if (PROF) p_begin = clock64();
    using F = Fmt<TYPE>;
    Smem<TYPE> &S = *reinterpret_cast<Smem<TYPE> *>(ggb_dec_smem);
    const Phase &ph = S.ph[phi];
    const int lane = threadIdx.x & 31;
    uint32_t tmask = (uint32_t)S.wst[kCW + ci][2]; // bit b: parity of term buffer b's barriers at its next use
    int gslot = S.wst[kCW + ci][1];               // slots of earlier phases (the numbering of S.done_slot)
    int u0, u1;
    unit_range(cta, G, ph.rot, ph.uq, ph.ur, u0, u1);
    const int n_passes = (u1 - u0) * ph.unit_passes;
    const int npairs = ((u1 - u0) * ph.unit_rows) >> 1;
    const int kA = ph.kA, kind = ph.kind, K = ph.K;
    const bool quad = QUAD && ph.quad != 0;
    const int row0 = u0 * ph.unit_rows;
    const bool bcast = ph.bcast != 0;
    // the descriptor lives in shared memory and the waits below are memory barriers to the compiler: what the row
    // finish needs is read once, here
    const int n_lr = ph.n_lr, sr0 = ph.seg_rows[0], sr1 = ph.seg_rows[1], so0 = ph.seg_off[0], so1 = ph.seg_off[1], so2 = ph.seg_off[2];
    const int tail_blocks = ph.tail_blocks, kf = ph.kf, resid_embed = ph.resid_embed;
    float *const out_f = ph.out;
    ll_t *const out_ll = ph.out_ll;
    const ll_t *const resid = ph.resid_ll;
    auto oidx = [&](int lr) {
        if (lr < sr0) return so0 + lr;
        lr -= sr0;
        return lr < sr1 ? so1 + lr : so2 + lr - sr1;
    };
    float *A = K <= kAccMaxK ? reinterpret_cast<float *>(S.u + U_ACC) : S.acc_small;
    float *GU = reinterpret_cast<float *>(S.u + U_GU);
    const uint32_t tf0 = smem_u32(&S.tfull[0]), te0 = smem_u32(&S.tempty[0]);
    const uint32_t terms0 = smem_u32(S.u) + (uint32_t)terms_off(K);
    const int nbuf = terms_nbuf(K);
    const int j = lane >> 1, ab = lane & 1;
    int segc = 0, rem = 0;
    Best best{-INFINITY, -1};

    auto finish_row = [&](int rl, float total, ll_t rv) {
        const int lr = row0 + rl;
        if (lr >= n_lr) return;
        if (kind == PH_GATE_UP) {
            GU[rl] = total;
            return;
        }
        const int gi = oidx(lr);
        if (resid) total = __fadd_rn(total, ll_get_f(rv, resid + gi, tag));
        else if (resid_embed) total = __fadd_rn(total, embed_value<TYPE>(P, S.sched[3], gi));
        if (out_f) out_f[gi] = total;
        if (out_ll) ll_put_f(P, bcast, out_ll + gi, total, tag);
        if (total > best.v || best.i < 0) best = Best{total, gi};
    };
    // ROW-LANE MODE (multi-segment phases whose CTA share is <= 32 rows: O / DOWN): lane = row for the whole phase, the
    // accumulator stays in a register from segment to segment, and ONE chain warp takes every slot (the segments of a
    // row are a dependent chain anyway: handing it from warp to warp only adds latency)
    const bool rowlane = (kA > 1 || (quad && kf > 0)) && 2 * npairs <= 32 && kind != PH_GATE_UP;
    // memory fences only where rows are handed from warp to warp (accumulators of multi-segment rows, gate / up rows)
    const bool handoff = kind == PH_GATE_UP || kA > 1 || (quad && kf > 0);
    float racc = 0.0f;
    int tb = 0, turn = 0;
    // SEGMENT-DRIVEN ROW-LANE CHAIN (O / DOWN of the Gemma shapes): one round per K segment, for every row at once - a
    // row's 512 (DOWN, 2B) terms are 8 rounds of 64 dependent adds and nothing else.  Round k starts when the slot
    // holding the last pass of segment k has its terms; a lane reads its pass wherever it lies (a segment's passes
    // span at most two slots, so two term buffers are live at a time), and a buffer is handed back as soon as the
    // rounds have consumed all of its passes.  The slot-driven loop below costs a round per (slot, segment) piece -
    // 11 instead of 8 for DOWN - and ~300 cycles of scalar bookkeeping per round.
    // (n_passes > 0: a CTA without rows in the phase - tensor parallel with many ranks, thin matrices - has nothing to wait for)
    const bool seg_driven = rowlane && !quad && n_passes > 0 && n_passes == npairs * kA && (npairs + F::PPS - 2) / F::PPS + 1 <= nbuf;
    if (seg_driven) {
        const int n_slots = (n_passes + F::PPS - 1) / F::PPS;
        if (ci == 0) {
            // (unsigned arithmetic and non-unrolled inner loops on purpose: this is scalar control code that one warp
            // runs at ~5 cycles per instruction between two rounds of adds; signed divisions by PPS and an unrolled
            // two-trip loop tripled it)
            constexpr uint32_t SH = F::PPS == 16 ? 4 : 3;
            static_assert((1u << SH) == (uint32_t)F::PPS, "PPS must be 8 or 16");
            const uint32_t pair = (uint32_t)lane >> 1, np_u = (uint32_t)npairs, nbuf_u = (uint32_t)nbuf;
            const bool act2 = pair < np_u;
            uint32_t wmask = tmask;
            uint32_t wb = 0, rb = 0, waited = 0, released = 0;
            uint32_t p = act2 ? pair : 0u, ls = 0, lb = 0; // this lane's pass of segment k, its slot and that slot's buffer
            uint32_t done = 0;
            ll_t rv = 0; // the residual word: requested now, used after the last round
            if (act2 && resid && row0 + lane < n_lr) rv = ll_peek(resid + oidx(row0 + lane));
            // several segments per round when a slot holds them (Gemma-2B DOWN: 7 pairs per segment, 16 passes per slot
            // -> 2): the wait / release / loop code between two runs of adds is paid once per slot instead of per segment
            const int spr = min(4, max(1, (int)(F::PPS / np_u)));
#pragma unroll 1
            for (int k = 0; k < kA; k += spr) {
                const int nseg = min(spr, kA - k);
                done += np_u * (uint32_t)nseg; // passes consumed once this round is over
                const uint32_t need = (done - 1u) >> SH;
                if (PROF) pt0 = clock64();
#pragma unroll 1
                while (waited <= need) {
                    mbar_wait(tf0 + 8u * wb, (wmask >> wb) & 1u);
                    if (PROF && prof && lane == 0 && waited < 8) P.prof[64 * cta + 48 + 2 * waited] = clock64() - p_begin;
                    wmask ^= 1u << wb;
                    wb = wb + 1u == nbuf_u ? 0u : wb + 1u;
                    waited++;
                }
                if (PROF) { const long long t1 = clock64(); p_wait += t1 - pt0; pt0 = t1; p_rounds++; }
                const bool fin = k + nseg == kA;
#pragma unroll 1
                for (int q = 0; q < nseg; q++, p += np_u) {
                    const uint32_t ns = p >> SH;
#pragma unroll 1
                    while (ls < ns) {
                        ls++;
                        lb = lb + 1u == nbuf_u ? 0u : lb + 1u;
                    }
                    const uint32_t addr = terms0 + lb * (uint32_t)(kTermFloats * 4) + (2u * (p & (uint32_t)(F::PPS - 1)) + (uint32_t)ab) * (uint32_t)(kTermRow * 4);
                    if (PROF) pta = clock64();
                    racc = chain_add(addr, 0, k + q == kA - 1 ? tail_blocks : kSeg, racc);
                    if (PROF) { const long long t1 = clock64(); p_add += t1 - pta; if (prof && lane == 0 && k + q < 8) P.prof[64 * cta + 24 + k + q] = t1 - p_begin; }
                }
                if (fin && act2) finish_row(lane, racc, rv);
                __syncwarp();
#pragma unroll 1
                while (released < waited && (((released + 1u) << SH) <= done || fin)) {
                    if (lane == 0) mbar_arrive(te0 + 8u * rb);
                    if (PROF && prof && lane == 0 && released < 8) P.prof[64 * cta + 49 + 2 * released] = clock64() - p_begin;
                    rb = rb + 1u == nbuf_u ? 0u : rb + 1u;
                    released++;
                }
                if (PROF) p_chain += clock64() - pt0;
            }
        }
        for (int sl = 0; sl < n_slots; sl++) {
            tmask ^= 1u << tb;
            if (++tb == nbuf) tb = 0;
        }
        gslot += n_slots;
    } else if (!rowlane && kA == 1 && !quad) {
        // SINGLE-SEGMENT PHASES (K <= 2048: QKV / GATE_UP / LOGITS of Gemma-2B): a slot is one round, lane = row of the
        // slot, nothing is carried from slot to slot.  Same arithmetic as the general loop below with its bookkeeping
        // (a single warp runs scalar control code at ~5 cycles per instruction) stripped to what this case needs.
#pragma unroll 1
        for (int m0 = 0; m0 < n_passes; m0 += F::PPS, gslot++) {
            if (turn == ci) {
                const int cnt = min(F::PPS, n_passes - m0);
                // the residual word is requested BEFORE the wait for the terms (it is there since the previous phase): its
                // L2 round trip used to sit behind the 64 adds, on the tail of O
                const int rl = 2 * m0 + lane;
                const bool active = j < cnt;
                ll_t rv = 0;
                if (active && resid && row0 + rl < n_lr) rv = ll_peek(resid + oidx(row0 + rl));
                if (PROF) pt0 = clock64();
                mbar_wait(tf0 + 8 * tb, (tmask >> tb) & 1u);
                if (PROF) { const long long t1 = clock64(); p_wait += t1 - pt0; pt0 = t1; p_rounds++; const int k = m0 / F::PPS; if (prof && lane == 0 && ci == 0 && k < 8) P.prof[64 * cta + 48 + 2 * k] = t1 - p_begin; }
                const float acc = chain_add(terms0 + (uint32_t)tb * (uint32_t)(kTermFloats * 4) + (uint32_t)(lane * kTermRow * 4), 0, tail_blocks, 0.0f);
                if (PROF) { const long long t1 = clock64(); p_add += t1 - pt0; }
                if (active) finish_row(rl, acc, rv);
                __syncwarp();
                if (lane == 0) mbar_arrive(te0 + 8 * tb);
                if (PROF) { const long long t1 = clock64(); p_chain += t1 - pt0; pt0 = t1; }
                if (handoff) __threadfence_block();
                while (S.done_slot != gslot) __nanosleep(20);
                __syncwarp();
                if (lane == 0) S.done_slot = gslot + 1;
                if (PROF) { const long long t1 = clock64(); p_fin += t1 - pt0; const int k = m0 / F::PPS; if (prof && lane == 0 && ci == 0 && k < 8) P.prof[64 * cta + 49 + 2 * k] = t1 - p_begin; }
            }
            tmask ^= 1u << tb;
            if (++tb == nbuf) tb = 0;
            if (++turn == kNCH) turn = 0;
        }
    } else
#pragma unroll 1
    for (int m0 = 0; m0 < n_passes; m0 += F::PPS, gslot++) {
        const int cnt = min(F::PPS, n_passes - m0);
        const bool mine = rowlane ? ci == 0 : turn == ci;
        const uint32_t tpar = (tmask >> tb) & 1u;
        if (mine) {
            if (PROF) pt0 = clock64();
            mbar_wait(tf0 + 8 * tb, tpar);
            if (PROF) { const long long t1 = clock64(); p_wait += t1 - pt0; const int k = m0 / F::PPS; if (prof && lane == 0 && ci == 0 && k < 8) P.prof[64 * cta + 48 + 2 * k] = t1 - p_begin; }
        }
        const uint32_t tbase = terms0 + (uint32_t)tb * (uint32_t)(kTermFloats * 4);
        int jj = 0;
#pragma unroll 1
        while (jj < cnt) {
            const bool isB = segc >= kA;
            const int run = isB ? cnt - jj : min(cnt - jj, npairs - rem);
            if (mine) {
                if (PROF) { pt0 = clock64(); p_rounds++; }
                if (rowlane) {
                    if (!isB) {
                        const int pair = lane >> 1;
                        const bool act2 = pair >= rem && pair < rem + run;
                        const bool last = !quad && segc == kA - 1;
                        const int len = last ? tail_blocks : kSeg;
                        ll_t rv = 0;
                        if (last && act2 && resid && row0 + lane < n_lr) rv = ll_peek(resid + oidx(row0 + lane));
                        const int c = 2 * (jj + pair - rem) + ab;
                        if (PROF) pta = clock64();
                        const float r = chain_add(tbase + (uint32_t)((c & 31) * kTermRow * 4), 0, len, segc > 0 ? racc : 0.0f);
                        if (PROF) { const long long t1 = clock64(); p_add += t1 - pta; if (prof && lane == 0 && p_rounds <= 8) P.prof[64 * cta + 23 + p_rounds] = t1 - p_begin; }
                        if (act2) {
                            racc = r;
                            if (last) finish_row(lane, r, rv);
                        }
                    } else {
                        const int q = lane >> 2, kq = lane & 3;
                        const bool act2 = q >= rem && q < rem + run;
                        ll_t rv = 0;
                        if (act2 && resid && row0 + lane < n_lr) rv = ll_peek(resid + oidx(row0 + lane));
                        const int c = 2 * (jj + q - rem) + (kq >> 1);
                        const float r = chain_add(tbase + (uint32_t)((c & 31) * kTermRow * 4), 32 * (kq & 1), tail_blocks, kf > 0 ? racc : 0.0f);
                        if (act2) finish_row(lane, r, rv);
                    }
                } else {
                    const bool active = j >= jj && j < jj + run;
                    const int idx = rem + (j - jj);
                    // rows that continue an accumulator: every earlier slot must be committed
                    if (isB ? kf > 0 : segc > 0) {
                        while (S.done_slot < gslot) __nanosleep(20);
                        __threadfence_block();
                    }
                    if (!isB) {
                        const bool last = !quad && segc == kA - 1;
                        const int len = last ? tail_blocks : kSeg;
                        const int rl = 2 * idx + ab;
                        ll_t rv = 0;
                        if (last && active && resid && row0 + rl < n_lr) rv = ll_peek(resid + oidx(row0 + rl));
                        float acc = (active && segc > 0) ? A[rl] : 0.0f;
                        if (PROF) pta = clock64();
                        acc = chain_add(tbase + (uint32_t)(lane * kTermRow * 4), 0, len, acc);
                        if (PROF) { const long long t1 = clock64(); p_add += t1 - pta; pta = t1; }
                        if (active) {
                            if (last) finish_row(rl, acc, rv);
                            else A[rl] = acc;
                        }
                        if (PROF) p_row += clock64() - pta;
                    } else { // the short segment of rows 4 idx + 2 ab and + 1: two 32-term runs of this lane
                        const int rl = 4 * idx + 2 * ab;
                        ll_t rv0 = 0, rv1 = 0;
                        if (active && resid) {
                            if (row0 + rl < n_lr) rv0 = ll_peek(resid + oidx(row0 + rl));
                            if (row0 + rl + 1 < n_lr) rv1 = ll_peek(resid + oidx(row0 + rl + 1));
                        }
                        float a0 = (active && kf > 0) ? A[rl] : 0.0f, a1 = (active && kf > 0) ? A[rl + 1] : 0.0f;
                        a0 = chain_add(tbase + (uint32_t)(lane * kTermRow * 4), 0, tail_blocks, a0);
                        a1 = chain_add(tbase + (uint32_t)(lane * kTermRow * 4), 32, tail_blocks, a1);
                        if (active) {
                            finish_row(rl, a0, rv0);
                            finish_row(rl + 1, a1, rv1);
                        }
                    }
                }
                __syncwarp();
                if (PROF) p_chain += clock64() - pt0;
            }
            jj += run;
            rem += run;
            if (!isB && rem == npairs) {
                rem = 0;
                segc++;
            }
        }
        if (mine) {
            if (PROF) pt0 = clock64();
            __syncwarp();
            if (lane == 0) mbar_arrive(te0 + 8 * tb); // the terms are consumed: the consumers may refill the buffer
            if (!rowlane) { // commit in slot order; then the GATE_UP groups this slot completed
                if (handoff) __threadfence_block(); // this warp's accumulator / staging writes before its commit
                while (S.done_slot != gslot) __nanosleep(20);
                __syncwarp();
                if (lane == 0) S.done_slot = gslot + 1;
            }
            if (PROF) { const long long t1 = clock64(); p_fin += t1 - pt0; const int k = m0 / F::PPS; if (prof && lane == 0 && ci == 0 && k < 8) P.prof[64 * cta + 49 + 2 * k] = t1 - p_begin; }
        }
        tmask ^= 1u << tb;
        if (++tb == nbuf) tb = 0;
        if (++turn == kNCH) turn = 0;
    }
    if (PROF && prof && lane == 0 && ci == 0) {
        long long *o = P.prof + 64 * cta + 16;
        o[0] = p_wait; o[1] = p_chain; o[2] = p_fin; o[3] = p_rounds; o[4] = clock64() - p_begin; o[5] = n_passes;
        o[6] = p_add; o[7] = p_row;
    }
    if (lane == 0) {
        S.wst[kCW + ci][2] = (int)tmask;
        S.wst[kCW + ci][1] = gslot;
        if (rowlane && ci == 0) S.done_slot = gslot; // (the other chain warps did not touch this phase's slots)
    }
    if (kind == PH_LOGITS) { // this warp's arg-max candidate (strict '>' / lowest index, ref: src/gemma_model.cpp:538-543)
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            Best o;
            o.v = __shfl_xor_sync(0xffffffffu, best.v, off);
            o.i = __shfl_xor_sync(0xffffffffu, best.i, off);
            best = better(best, o);
        }
        if (lane == 0) {
            const int ce = (P.rank * G + cta) * kNCH + ci;
            ll_put_f(P, true, P.cand_ll + ce, best.v, tag);
            ll_put(P, true, P.cand_ll + P.R * G * kNCH + ce, (uint32_t)best.i, tag);
        }
    }
}

#define PROF_STAMP(i) do { if (PROF && prof && (int)threadIdx.x == kC0) P.prof[64 * cta + (i)] = clock64(); } while (0)
// ---------------------------------------------------------------- SCORE
// kq = mul_mat(k, q) (ref: src/gemma_model.cpp:465-474) for one (kv head, 32-position chunk, half of the query heads)
// per item; rope + Q scale are applied to q on load, and the item that owns the newest position ropes K, rounds K / V to
// fp16, stores them in the cache and uses the K row.  The K rows of older positions do not depend on this token: they
// travel as bulk copies (TMA) while the CTA still waits for q, double-buffered so that the next item's chunk lands
// during this item's dots.  fp16 x fp16 products are exact; they are accumulated in DOUBLE and rounded to fp32 once, as
// the reference's vec_dot_f16 does (ref: src/kernals.cl:158-164) - the order of a double sum is invisible after that
// rounding.  A warp scores its two positions against the item's (<= 4) heads and reduces the 8 lane partials with one
// folded butterfly.  Scores leave as flagged words.
// MHA_OPT (= the QUAD instantiation, the one Gemma-7B runs): items in head-major order, a contiguous range per CTA, so
// that a CTA with several items (16 kv heads x n_kv / 32 chunks = 7 per CTA at n_kv 2048) walks consecutive chunks of one
// head and stages that head's q once instead of once per item; and a one-head form of the dots and the fold.  Kept out
// of the other instantiation because its extra live state costs the Gemma-2B kernel 1 % (registers: 96, spills move
// into the pass loop) for nothing - an MQA model has one item per CTA.
template <int TYPE, bool PROF, bool MHA_OPT>
__device__ __forceinline__ void score_phase(const DecoderParams &P, int phi, int cta, int G, int n_pos, uint32_t tag, bool prof) {
    Smem<TYPE> &S = *reinterpret_cast<Smem<TYPE> *>(ggb_dec_smem);
    const Phase &ph = S.ph[phi];
    const int tid = (int)threadIdx.x - kC0, lane = tid & 31, warp = tid >> 5;
    uint32_t kvp = (uint32_t)S.wst[warp][3]; // bit b = parity of K buffer b's barrier
    const int d = P.d, halfd = d >> 1, Gq = P.Gq, hsplit = P.hsplit, Gh = P.Gh, ng = P.ng;
    const int nchunk = S.sched[0];
    const int n_items = ng * nchunk * hsplit;
    const int pos = n_pos - 1;
    const size_t kv_stride = (size_t)ng * d; // this rank's cache holds its own KV heads only
    double *Sq = reinterpret_cast<double *>(S.u + U_SQ); // [4][d] fp16-rounded q
    const uint32_t kc0 = smem_u32(S.u + U_KC), kvbar0 = smem_u32(&S.kv_bar[0]);
    const uint32_t row_bytes = (uint32_t)d * 2;
    int first = cta + ph.rot;
    if (first >= G) first -= G;

    auto decode = [&](int it, int &hh, int &gl, int &c) {
        hh = hsplit == 2 ? (it & 1) : 0;
        const int gc = hsplit == 2 ? (it >> 1) : it;
        if (MHA_OPT) { // head-major
            gl = gc / nchunk;
            c = gc - gl * nchunk;
        } else {
            gl = gc & (ng - 1);
            c = gc >> P.ng_shift;
        }
    };
    // the CTA's items: first, first + G, ... or (MHA_OPT) the contiguous range [first * per, first * per + per)
    const int per = (n_items + G - 1) / G;
    const int it_begin = MHA_OPT ? min(first * per, n_items) : first;
    const int it_end = MHA_OPT ? min(it_begin + per, n_items) : n_items;
    const int it_step = MHA_OPT ? 1 : G;
    int q_h0 = -1; // (MHA_OPT) the head whose q is staged in Sq
    // K rows of an item's chunk (the new position's row is computed, not loaded).  One KV head: the rows of a chunk are
    // contiguous in the cache, so warps 0 and 1 issue one copy each (rows before | after the new one); otherwise the
    // warp that scores rows w, w + 16 copies them.  Every warp arrives once on the buffer's barrier.
    auto issue_k = [&](int it, int buf) {
        if (lane != 0) return;
        int hh, gl, c;
        decode(it, hh, gl, c);
        const int p0 = c * kChunk, np = min(kChunk, n_pos - p0);
        const int jnew = pos - p0;
        const bool owns_new = jnew >= 0 && jnew < np;
        const uint32_t kc = kc0 + (uint32_t)buf * 16384u, bar = kvbar0 + 8u * (uint32_t)buf;
        fence_proxy_async(); // the buffer was last touched through the generic proxy (ordered by a barrier)
        if (ng == 1) {
            const int lo = (warp & 1) ? (owns_new ? jnew + 1 : np) : 0;
            const int hi = (warp & 1) ? np : (owns_new ? jnew : np);
            const uint32_t bytes = warp < 2 && hi > lo ? (uint32_t)(hi - lo) * row_bytes : 0u;
            mbar_expect_tx(bar, bytes);
            if (bytes) bulk_g2s_plain(kc + (uint32_t)lo * row_bytes, ph.kcache + (size_t)(p0 + lo) * d, bytes, bar);
        } else {
            uint32_t bytes = 0;
            for (int jr = warp; jr < np; jr += kCW) bytes += jr == jnew ? 0u : row_bytes;
            mbar_expect_tx(bar, bytes);
            for (int jr = warp; jr < np; jr += kCW) {
                if (jr == jnew) continue;
                bulk_g2s_plain(kc + (uint32_t)jr * row_bytes, ph.kcache + (size_t)(p0 + jr) * kv_stride + (size_t)gl * d, row_bytes, bar);
            }
        }
    };

    int buf = 0;
    if (it_begin < it_end) issue_k(it_begin, 0);
#pragma unroll 1
    for (int it = it_begin; it < it_end; it += it_step) {
        int hh, gl, c;
        decode(it, hh, gl, c);
        const int g = P.g0 + gl; // global KV head
        const int p0 = c * kChunk, np = min(kChunk, n_pos - p0);
        const int h0 = g * Gq + hh * Gh;
        const int jnew = pos - p0;
        const bool owns_new = jnew >= 0 && jnew < np;
        __half *Kc = reinterpret_cast<__half *>(S.u + U_KC + buf * 16384);
        PROF_STAMP(0);
        {
        const float *cosr = P.rope_cos + (size_t)pos * halfd, *sinr = P.rope_sin + (size_t)pos * halfd;
        const ll_t *q_all = ph.in_ll, *k_all = q_all + (size_t)P.H * d, *v_all = k_all + (size_t)P.Hkv * d;
        // q: rope (NEOX pairs i, i + d/2), scale, round to fp16 (ggml mul_mat INIT converts src1 to fp16), widen
        for (int i = tid; i < ((!MHA_OPT || h0 != q_h0) ? Gh * halfd : 0); i += kCT) {
            const int jh = i >> P.hd_shift, e = i & (halfd - 1);
            const ll_t *qh = q_all + (size_t)(h0 + jh) * d;
            const float cs = cosr[e], sn = sinr[e];
            ll_t w[2];
            ll_gather<2>(w, qh + e, halfd, 2, tag);
            const float x0 = __uint_as_float((uint32_t)w[0]), x1 = __uint_as_float((uint32_t)w[1]);
            const float r0 = __fmul_rn(__fsub_rn(__fmul_rn(x0, cs), __fmul_rn(x1, sn)), P.q_scale);
            const float r1 = __fmul_rn(__fadd_rn(__fmul_rn(x0, sn), __fmul_rn(x1, cs)), P.q_scale);
            // a lane's dims 8l .. 8l+7 as four 16-byte pieces, piece k of all lanes contiguous: conflict-free LDS.128
            const int e1 = e + halfd;
            Sq[jh * 256 + ((e >> 1) & 3) * 64 + (e >> 3) * 2 + (e & 1)] = (double)__half2float(__float2half_rn(r0));
            Sq[jh * 256 + ((e1 >> 1) & 3) * 64 + (e1 >> 3) * 2 + (e1 & 1)] = (double)__half2float(__float2half_rn(r1));
        }
        if (owns_new) { // KV store of the new position (ref: src/gemma_model.cpp:499-518)
            for (int e = tid; e < halfd; e += kCT) {
                const ll_t *kp = k_all + (size_t)g * d + e;
                const ll_t w0 = ll_peek(kp), w1 = ll_peek(kp + halfd);
                const float cs = cosr[e], sn = sinr[e];
                const float x0 = ll_get_f(w0, kp, tag), x1 = ll_get_f(w1, kp + halfd, tag);
                const __half k0 = __float2half_rn(__fsub_rn(__fmul_rn(x0, cs), __fmul_rn(x1, sn)));
                const __half k1 = __float2half_rn(__fadd_rn(__fmul_rn(x0, sn), __fmul_rn(x1, cs)));
                Kc[jnew * d + e] = k0;
                Kc[jnew * d + e + halfd] = k1;
                if (hh == 0) {
                    __half *krow = ph.kcache + (size_t)pos * kv_stride + (size_t)gl * d;
                    krow[e] = k0;
                    krow[e + halfd] = k1;
                }
            }
            if (hh == 0) {
                for (int e = tid; e < d; e += kCT) {
                    const ll_t *vp = v_all + (size_t)g * d + e;
                    ph.vcache[(size_t)pos * kv_stride + (size_t)gl * d + e] = __float2half_rn(ll_get_f(ll_peek(vp), vp, tag));
                }
            }
        }
        }
        PROF_STAMP(1);
        mbar_wait(kvbar0 + 8u * (uint32_t)buf, (kvp >> buf) & 1u);
        kvp ^= 1u << buf;
        consumer_bar();
        PROF_STAMP(2);
        if (MHA_OPT) q_h0 = h0;
        if (it + it_step < it_end) issue_k(it + it_step, buf ^ 1); // the other buffer was last read before the barrier above
        {
            const bool in = 8 * lane < d;
            const int j0 = warp, j1 = warp + kCW;
            // both positions of the warp against one read of q (the fp64 q rows in shared memory are the phase's
            // bandwidth: 4 wavefronts per LDS.128, once per (head, dim quad) instead of once per position)
            double acc[2][4];
            double kd[2][8];
#pragma unroll
            for (int p = 0; p < 2; p++) {
                const int jr = p ? j1 : j0;
                uint4 raw = make_uint4(0, 0, 0, 0);
                if (in && jr < np) raw = *reinterpret_cast<const uint4 *>(Kc + jr * d + 8 * lane);
                const __half2 *h2 = reinterpret_cast<const __half2 *>(&raw);
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const float2 f = __half22float2(h2[e]);
                    kd[p][2 * e] = (double)f.x;
                    kd[p][2 * e + 1] = (double)f.y;
                }
            }
            if (MHA_OPT && Gh == 1) {
                // one query head per item (MHA): 16 DFMA instead of 64, and a two-value fold (one exchange, four plain
                // steps) instead of the eight-value one
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    double2 q2 = make_double2(0.0, 0.0);
                    if (in) q2 = *reinterpret_cast<const double2 *>(Sq + k * 64 + lane * 2);
                    a0 = __fma_rn(kd[0][2 * k], q2.x, a0);
                    a0 = __fma_rn(kd[0][2 * k + 1], q2.y, a0);
                    a1 = __fma_rn(kd[1][2 * k], q2.x, a1);
                    a1 = __fma_rn(kd[1][2 * k + 1], q2.y, a1);
                }
                const bool hb4 = lane & 16;
                double r = __dadd_rn(hb4 ? a1 : a0, __shfl_xor_sync(0xffffffffu, hb4 ? a0 : a1, 16));
#pragma unroll
                for (int off = 8; off >= 1; off >>= 1) r = __dadd_rn(r, __shfl_xor_sync(0xffffffffu, r, off));
                const int js1 = hb4 ? j1 : j0;
                if ((lane & 15) == 0 && js1 < np)
                    ll_put_f(P, false, ph.s_ll + (size_t)h0 * P.n_ctx + p0 + js1, __double2float_rn(r), tag);
            } else {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    double2 q2 = make_double2(0.0, 0.0);
                    if (u < Gh && in) q2 = *reinterpret_cast<const double2 *>(Sq + u * 256 + k * 64 + lane * 2);
                    a0 = __fma_rn(kd[0][2 * k], q2.x, a0);
                    a0 = __fma_rn(kd[0][2 * k + 1], q2.y, a0);
                    a1 = __fma_rn(kd[1][2 * k], q2.x, a1);
                    a1 = __fma_rn(kd[1][2 * k + 1], q2.y, a1);
                }
                acc[0][u] = a0;
                acc[1][u] = a1;
            }
            // folded butterfly: xor 16 keeps one position per half-warp, xor 8 two heads, xor 4 one head
            const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
            double k4[4], k2[2];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const double send = b4 ? acc[0][u] : acc[1][u], keep = b4 ? acc[1][u] : acc[0][u];
                k4[u] = __dadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 16));
            }
#pragma unroll
            for (int v = 0; v < 2; v++) {
                const double send = b3 ? k4[v] : k4[2 + v], keep = b3 ? k4[2 + v] : k4[v];
                k2[v] = __dadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 8));
            }
            double k1;
            {
                const double send = b2 ? k2[0] : k2[1], keep = b2 ? k2[1] : k2[0];
                k1 = __dadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 4));
            }
            k1 = __dadd_rn(k1, __shfl_xor_sync(0xffffffffu, k1, 2));
            k1 = __dadd_rn(k1, __shfl_xor_sync(0xffffffffu, k1, 1));
            const int js = b4 ? j1 : j0, us = (b3 ? 2 : 0) + (b2 ? 1 : 0);
            if ((lane & 3) == 0 && js < np && us < Gh)
                ll_put_f(P, false, ph.s_ll + (size_t)(h0 + us) * P.n_ctx + p0 + js, __double2float_rn(k1), tag);
            }
        }
        PROF_STAMP(3);
        consumer_bar(); // Sq and (after one more item) this K buffer are rewritten
        buf ^= 1;
    }
    if (lane == 0) S.wst[warp][3] = (int)kvp;
}

// ---------------------------------------------------------------- PV
// soft_max_ext + kqv = mul_mat(v, kq) (ref: src/gemma_model.cpp:476-489) for one (head, W-dim slice) per item.  Every
// item rebuilds the head's whole soft-max row from the flagged scores ([UPSTREAM] ggml_compute_forward_soft_max: global
// max, p = exp_table[f16(s - max)], sum in double - exact for fp16-valued terms in any order -, p *= (float)(1/sum)),
// rounds P to fp16 (ggml mul_mat INIT of the kqv node) and accumulates out[i] = sum_p V16[p][i] * P16[p] in double.
// Thread (o, pl): dims 8 o .. 8 o + 7 of the slice, positions pl, pl + PL, ...; the V rows of older positions do not
// depend on this token and are requested before the scores arrive.
template <int TYPE, bool PROF, bool MHA_OPT>
__device__ __forceinline__ void pv_phase(const DecoderParams &P, int phi, int cta, int G, int n_pos, uint32_t tag, bool prof) {
    Smem<TYPE> &S = *reinterpret_cast<Smem<TYPE> *>(ggb_dec_smem);
    const Phase &ph = S.ph[phi];
    constexpr int NPRE = 4; // V rows per thread held in registers, after the NST rows staged in shared memory
    const int tid = (int)threadIdx.x - kC0, lane = tid & 31, warp = tid >> 5;
    const int d = P.d, W = P.pv_w;
    // the thread's first NST (<= 4) V rows wait in a private 16-byte slot each of the union region, copied there
    // asynchronously (no register is held while the soft-max row is built): [row k][thread]
    const int sp_bytes = (P.n_ctx * 4 + 127) & ~127;
    const int NST = min(4, (U_BYTES - sp_bytes - 16 * 32 * 8) / (kCT * 16));
    const uint32_t stage = smem_u32(S.u) + (uint32_t)(sp_bytes + 16 * 32 * 8) + 16u * (uint32_t)tid;
    const int n_items = P.nh * P.pv_slices;
    const int pos = n_pos - 1;
    const size_t kv_stride = (size_t)P.ng * d;
    float *sp = reinterpret_cast<float *>(S.u + U_PVP);
    double *part = reinterpret_cast<double *>(S.u + sp_bytes); // [16 warps][32]
    const int o = tid & ((1 << P.pv_oct_shift) - 1), pl = tid >> P.pv_oct_shift, PL = kCT >> P.pv_oct_shift;
    int first = cta + ph.rot;
    if (first >= G) first -= G;
#pragma unroll 1
    for (int it = first; it < n_items; it += G) {
        PROF_STAMP(0);
        const int hl = it >> P.pv_sl_shift, slice = it & (P.pv_slices - 1);
        const int h = P.h0 + hl, g = h >> P.gq_shift, gl = g - P.g0;
        const __half *vbase = ph.vcache + (size_t)gl * d + (size_t)slice * W + 8 * o;
        for (int k = 0; k < NST; k++) {
            const int p = pl + PL * k;
            if (p < pos) cp_async16(stage + (uint32_t)k * (kCT * 16), vbase + (size_t)p * kv_stride);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        // the head's scores: 4 flagged words in flight per thread
        const ll_t *srow = ph.s_ll + (size_t)h * P.n_ctx;
        float mx = -INFINITY;
#pragma unroll 1
        for (int i0 = tid; i0 < n_pos; i0 += 4 * kCT) {
            ll_t v[4];
            ll_gather<4>(v, srow + i0, kCT, (n_pos - i0 + kCT - 1) / kCT, tag);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + kCT * u;
                if (i < n_pos) {
                    const float s = __uint_as_float((uint32_t)v[u]);
                    sp[i] = s;
                    mx = fmaxf(mx, s);
                }
            }
        }
        PROF_STAMP(1);
        mx = warp_max_f32(mx);
        if (lane == 0) S.redf[warp] = mx;
        consumer_bar();
        PROF_STAMP(2);
        mx = S.redf[0];
#pragma unroll
        for (int w = 1; w < kCW; w++) mx = fmaxf(mx, S.redf[w]);
        double sum = 0.0;
        for (int i = tid; i < n_pos; i += kCT) {
            const float s = sp[i];
            float e = 0.0f;
            if (s != -INFINITY) {
                e = exp_f16(P, __fsub_rn(s, mx));
                sum = __dadd_rn(sum, (double)e);
            }
            sp[i] = e;
        }
        PROF_STAMP(3);
        // the next NPRE V rows of the thread: requested now (after the register-hungry exp), used after the staged ones
        uint4 vr[NPRE];
        const int pr0 = pl + PL * NST;
#pragma unroll
        for (int k = 0; k < NPRE; k++) {
            const int p = pr0 + PL * k;
            vr[k] = p < pos ? __ldcg(reinterpret_cast<const uint4 *>(vbase + (size_t)p * kv_stride)) : make_uint4(0, 0, 0, 0);
        }
        sum = warp_sum_f64(sum);
        if (lane == 0) S.red[warp] = sum;
        consumer_bar();
        PROF_STAMP(4);
        sum = S.red[0];
#pragma unroll
        for (int w = 1; w < kCW; w++) sum = __dadd_rn(sum, S.red[w]);
        const float fs = __double2float_rn(__ddiv_rn(1.0, sum));
        for (int i = tid; i < n_pos; i += kCT) sp[i] = __half2float(__float2half_rn(__fmul_rn(sp[i], fs)));
        consumer_bar();
        PROF_STAMP(5);
        double acc[8];
#pragma unroll
        for (int e = 0; e < 8; e++) acc[e] = 0.0;
        auto fma8 = [&](const uint4 &raw, int p) {
            const double pd = (double)sp[p];
            const __half2 *h2 = reinterpret_cast<const __half2 *>(&raw);
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const float2 f = __half22float2(h2[e]);
                acc[2 * e] = __fma_rn((double)f.x, pd, acc[2 * e]);
                acc[2 * e + 1] = __fma_rn((double)f.y, pd, acc[2 * e + 1]);
            }
        };
        if (MHA_OPT) {
            // (the Gemma-7B instantiation) rows beyond the staged and the early-requested ones - 16 rows per thread at
            // n_kv 2048 with 32-dim slices, 64 at 8192: two register batches in flight alternately, the next batch is
            // requested before the current one is used (otherwise every batch exposes an L2 round trip)
            auto load_rows = [&](uint4 (&v)[NPRE], int q0) {
#pragma unroll
                for (int k = 0; k < NPRE; k++) {
                    const int p = q0 + PL * k;
                    v[k] = p < pos ? __ldcg(reinterpret_cast<const uint4 *>(vbase + (size_t)p * kv_stride)) : make_uint4(0, 0, 0, 0);
                }
            };
            auto fma_rows = [&](const uint4 (&v)[NPRE], int q0) {
#pragma unroll
                for (int k = 0; k < NPRE; k++)
                    if (q0 + PL * k < pos) fma8(v[k], q0 + PL * k);
            };
            uint4 vb[NPRE];
            int q0 = pr0 + PL * NPRE;
            load_rows(vb, q0);
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            for (int k = 0; k < NST; k++) {
                const int p = pl + PL * k;
                if (p < pos) fma8(lds_u4(stage + (uint32_t)k * (kCT * 16)), p);
            }
            fma_rows(vr, pr0);
#pragma unroll 1
            while (q0 < pos) {
                const int q1 = q0 + PL * NPRE;
                load_rows(vr, q1);
                fma_rows(vb, q0);
                if (q1 >= pos) break;
                q0 = q1 + PL * NPRE;
                load_rows(vb, q0);
                fma_rows(vr, q1);
            }
        } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        for (int k = 0; k < NST; k++) {
            const int p = pl + PL * k;
            if (p < pos) fma8(lds_u4(stage + (uint32_t)k * (kCT * 16)), p);
        }
#pragma unroll
        for (int k = 0; k < NPRE; k++)
            if (pr0 + PL * k < pos) fma8(vr[k], pr0 + PL * k);
#pragma unroll 1
        for (int p0 = pr0 + PL * NPRE; p0 < pos; p0 += PL * NPRE) { // long contexts: NPRE rows in flight
#pragma unroll
            for (int k = 0; k < NPRE; k++) {
                const int p = p0 + PL * k;
                vr[k] = p < pos ? __ldcg(reinterpret_cast<const uint4 *>(vbase + (size_t)p * kv_stride)) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int k = 0; k < NPRE; k++)
                if (p0 + PL * k < pos) fma8(vr[k], p0 + PL * k);
        }
        }
        if (pl == (pos & (PL - 1))) { // the new position's V row: from the flagged v words, rounded to fp16 like the cache
            const ll_t *vp = ph.in_ll + (size_t)P.H * d + (size_t)P.Hkv * d + (size_t)g * d + (size_t)slice * W + 8 * o;
            uint4 raw;
            __half2 *h2 = reinterpret_cast<__half2 *>(&raw);
            ll_t vw[8]; // the eight words together: one round trip (one after the other they were eight, in every PV item)
            ll_gather<8>(vw, vp, 1, 8, tag);
#pragma unroll
            for (int e = 0; e < 4; e++)
                h2[e] = __halves2half2(__float2half_rn(__uint_as_float((uint32_t)vw[2 * e])), __float2half_rn(__uint_as_float((uint32_t)vw[2 * e + 1])));
            fma8(raw, pos);
        }
        PROF_STAMP(6);
        // lanes with the same octet (they differ in lane bits oct_shift .. 4), then the 16 warps (any order: double).
        // Folded butterfly: a lane hands over half of its sums at every step (4 + 2 + 1 doubles instead of 8 + 8 + 8)
        // and ends with dim e = 4 bit4 + 2 bit3 + bit2 of its octet.
        {
            const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
            double k4[4], k2[2];
#pragma unroll
            for (int v = 0; v < 4; v++) {
                const double send = b4 ? acc[v] : acc[4 + v], keep = b4 ? acc[4 + v] : acc[v];
                k4[v] = __dadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 16));
            }
#pragma unroll
            for (int v = 0; v < 2; v++) {
                const double send = b3 ? k4[v] : k4[2 + v], keep = b3 ? k4[2 + v] : k4[v];
                k2[v] = __dadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 8));
            }
            double k1 = __dadd_rn(b2 ? k2[1] : k2[0], __shfl_xor_sync(0xffffffffu, b2 ? k2[0] : k2[1], 4));
            if (P.pv_oct_shift == 1) k1 = __dadd_rn(k1, __shfl_xor_sync(0xffffffffu, k1, 2));
            const int oct = lane & ((1 << P.pv_oct_shift) - 1), e = (b4 ? 4 : 0) + (b3 ? 2 : 0) + (b2 ? 1 : 0);
            if (P.pv_oct_shift == 2 || !(lane & 2)) part[warp * 32 + 8 * oct + e] = k1;
        }
        consumer_bar();
        if (tid < W) {
            double t = part[tid];
#pragma unroll
            for (int w = 1; w < kCW; w++) t = __dadd_rn(t, part[w * 32 + tid]);
            const float r = __double2float_rn(t);
            const size_t idx = (size_t)h * d + (size_t)slice * W + tid;
            ph.f32_out[idx] = r;
            ll_put_f(P, ph.bcast != 0, ph.out_ll + idx, r, tag);
        }
        PROF_STAMP(7);
        consumer_bar();
    }
}

// ---------------------------------------------------------------- a consumer warp's share of a mat-vec phase
template <int TYPE, bool QUAD, bool PROF>
__device__ __forceinline__ void matvec_consumer(const DecoderParams &P, int phi, int cta, int G, uint32_t tag, bool prof) {
    long long pt0 = 0, p_full = 0, p_te = 0, p_comp = 0, p_begin = 0, p_pro = 0;
    if (PROF) p_begin = clock64();
    using F = Fmt<TYPE>;
    using SM = Smem<TYPE>;
    SM &S = *reinterpret_cast<SM *>(ggb_dec_smem);
    const Phase &ph = S.ph[phi];
    const int tid = (int)threadIdx.x - kC0, warp = tid >> 5, lane = tid & 31;
    const int kind = ph.kind;
    const uint32_t ring = smem_u32(&S.ring[0][0]);
    const uint32_t full0 = smem_u32(&S.full_bar[0]), empty0 = smem_u32(&S.empty_bar[0]);
    const uint32_t tf0 = smem_u32(&S.tfull[0]), te0 = smem_u32(&S.tempty[0]);
    const uint32_t act = smem_u32(S.u);
    int slot = S.wst[warp][0];
    uint32_t parity = (uint32_t)S.wst[warp][1], tmask = (uint32_t)S.wst[warp][2];
    const int gslot0 = S.wst[warp][4]; // slots of earlier phases (the numbering of S.done_slot)
    int u0, u1;
    unit_range(cta, G, ph.rot, ph.uq, ph.ur, u0, u1);
    const int n_passes = (u1 - u0) * ph.unit_passes;
    const int npairs = ((u1 - u0) * ph.unit_rows) >> 1;
    const int nb = ph.K >> 5, kA = ph.kA;
    const uint32_t terms0 = act + (uint32_t)terms_off(ph.K), act_d = act + (uint32_t)act_d_off(ph.K), act_bias = act + (uint32_t)act_bias_off(ph.K);
    const int nbuf = terms_nbuf(ph.K);
    if (n_passes > 0 || (cta == 0 && kind == PH_LOGITS)) { // CTA 0 also writes the result_norm dump
        if (kind == PH_O) quant_prologue<TYPE>(S, ph.in_ll, ph.K, tag);
        else if (kind == PH_DOWN) packed_prologue<TYPE>(S, ph.in_ll, ph.K, tag);
        else norm_prologue<TYPE>(P, S, ph, S.sched[3], tag, (kind == PH_LOGITS && cta == 0) ? ph.f32_out : nullptr);
    }
    if (P.trace && tid == 0) S.tr[1] = (long long)gtime();
    if (PROF) p_pro = clock64() - p_begin;
    // ---- passes: warp w owns passes w, w + 16, ... of the CTA's share (Q4_0: one per slot; Q8_0: 8 passes per
    // slot, so the two warp halves alternate slots).  Every warp waits for and releases every slot and every
    // term buffer.  The activation blocks live in registers and are reloaded only when the K segment changes
    // (the share is segment-major: at most ksegs times per phase).
    if (n_passes > 0) {
        const bool quad = QUAD && ph.quad != 0;
        int rem = warp, segc = 0, cur_seg = -1; // pass `warp` = pair / quad `rem` of section `segc`
        while (segc < kA && rem >= npairs) {
            rem -= npairs;
            segc++;
        }
        ActRegs A;
        int tb = 0;
#pragma unroll 1
        for (int p0 = 0, i = 0; p0 < n_passes; p0 += F::PPS, i++) {
            const int cnt = min(F::PPS, n_passes - p0);
            const int j = F::PPS == kCW ? warp : (((warp >> 3) == (i & 1)) ? (warp & 7) : F::PPS);
            const int aseg = segc < kA ? segc : ph.kf; // the short segment of a quad pass follows the kf full ones
            if (j < cnt && aseg != cur_seg) {
                load_act_regs<TYPE>(A, act, aseg, nb, lane);
                cur_seg = aseg;
            }
            if (PROF) pt0 = clock64();
            mbar_wait(full0 + 8 * slot, parity);
            if (PROF) { const long long t1 = clock64(); p_full += t1 - pt0; pt0 = t1; if (prof && tid == 0 && i < 4) P.prof[64 * cta + 32 + 3 * i] = t1 - p_begin; }
            // the chain warp is done with this term buffer's previous slot.  EVERY warp waits, also one without a
            // pass in this slot: its arrival below must not be counted for the buffer's previous use
#ifndef GGB_KO_TEMPTY
            mbar_wait(te0 + 8 * tb, ((tmask >> tb) & 1u) ^ 1u);
#endif
            if (PROF) { const long long t1 = clock64(); p_te += t1 - pt0; pt0 = t1; if (prof && tid == 0 && i < 4) P.prof[64 * cta + 33 + 3 * i] = t1 - p_begin; }
            if (j < cnt) {
                const uint32_t pass = ring + slot * F::SLOT_BYTES + j * F::PASS_BYTES;
                float term[4];
#ifdef GGB_KO_COMPUTE // knock-out build (timing only, wrong results): no term arithmetic, a loaded word is stored
                term[0] = term[1] = term[2] = term[3] = __uint_as_float(lds_u4(pass + 16 * lane).x & 0x3fffffffu);
#else
                if (quad && segc >= kA) pass_terms<TYPE, true>(pass, act_d, act_bias, aseg * kSeg + lane, A, lane, term);
                else pass_terms<TYPE, false>(pass, act_d, act_bias, aseg * kSeg + lane, A, lane, term);
#endif
                store_terms(terms0 + (uint32_t)tb * (uint32_t)(kTermFloats * 4), j, lane, term);
                rem += kCW;
                while (segc < kA && rem >= npairs) {
                    rem -= npairs;
                    segc++;
                }
            }
            __syncwarp();
            if (PROF) { const long long t1 = clock64(); p_comp += t1 - pt0; if (prof && tid == 0 && i < 4) P.prof[64 * cta + 34 + 3 * i] = t1 - p_begin; }
            if (lane == 0) {
                mbar_arrive(empty0 + 8 * slot);
                mbar_arrive(tf0 + 8 * tb);
            }
            tmask ^= 1u << tb;
            if (++tb == nbuf) tb = 0;
            if (++slot == SM::SLOTS) {
                slot = 0;
                parity ^= 1u;
            }
        }
    }
    if (lane == 0) {
        S.wst[warp][0] = slot;
        S.wst[warp][1] = (int)parity;
        S.wst[warp][2] = (int)tmask;
        S.wst[warp][4] = gslot0 + (n_passes + F::PPS - 1) / F::PPS;
    }
    if (kind == PH_GATE_UP) {
        // GeGLU + quantize of the CTA's 32-row groups, dealt to the consumer warps once their passes are done (the
        // chain warps only finish rows: their time per slot is what the pass loop waits for).  Group u is complete when
        // the slot holding the last pass of its rows is committed.
        const bool quad = QUAD && ph.quad != 0;
        const int base_last = quad ? kA * npairs : (kA - 1) * npairs, per = quad ? 16 : 32;
        const float *GU = reinterpret_cast<const float *>(S.u + U_GU);
#pragma unroll 1
        for (int u = warp; u < u1 - u0; u += kCW) {
            const int m_end = base_last + per * (u + 1) - 1;
            const int need = gslot0 + m_end / F::PPS + 1;
            while (S.done_slot < need) __nanosleep(20);
            __threadfence_block();
            geglu_unit(P, GU + 64 * u, ph.u_off + u0 + u, ph.out_ll, ph.f32_out, ph.bcast != 0, tag, lane);
        }
    }
    if (PROF && prof && tid == 0) {
        long long *o = P.prof + 64 * cta;
        o[0] = p_pro; o[1] = p_full; o[2] = p_te; o[3] = p_comp; o[4] = clock64() - p_begin; o[5] = n_passes;
    }
}

// ---------------------------------------------------------------- the kernel
// QUAD: the model has phases in the quad layout (a short last K segment).  A compile-time switch, because the pass loop
// is issue-bound: the extra branch of the quad layout costs a model without such phases (Gemma-2B) ~1 %.
template <int TYPE, bool QUAD, bool PROF>
__global__ void __launch_bounds__(kThreads, 1) decoder_kernel(const __grid_constant__ DecoderParams P) {
    using F = Fmt<TYPE>;
    using SM = Smem<TYPE>;
    SM &S = *reinterpret_cast<SM *>(ggb_dec_smem);
    const int cta = blockIdx.x, G = gridDim.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < SM::SLOTS; i++) {
            mbar_init(smem_u32(&S.full_bar[i]), 1);
            mbar_init(smem_u32(&S.empty_bar[i]), kCW);
        }
        for (int i = 0; i < 2; i++) mbar_init(smem_u32(&S.kv_bar[i]), kCW);
        for (int i = 0; i < kMaxBuf; i++) {
            mbar_init(smem_u32(&S.tfull[i]), kCW);
            mbar_init(smem_u32(&S.tempty[i]), 1);
        }
        S.done_slot = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n_pos = min(max(P.state[1], 0), P.n_ctx - 1) + 1;
    const uint32_t tag = *P.seq + 1u;
    if (tid < (kCW + kNCH) * 5) (&S.wst[0][0])[tid] = 0;
    if (tid == 0) {
        S.sched[0] = (n_pos + kChunk - 1) / kChunk;
        S.sched[3] = min(max(P.state[0], 0), P.V - 1);
        S.sched[4] = n_pos;
    }

    if (warp == kProdWarp) {
        if (lane == 0) producer_loop<TYPE>(P, S, cta, G);
        return;
    }
    if (tid < kPhWords) reinterpret_cast<uint32_t *>(&S.ph[0])[tid] = reinterpret_cast<const uint32_t *>(&P.phases[0])[tid];
    phase_bar();

    if (warp == 0 || warp > kProdWarp) { // ---- the chain warps
        const int ci = warp == 0 ? 0 : warp - kProdWarp;
#pragma unroll 1
        for (int pi = 0; pi < P.n_phases; pi++) {
            const Phase &ph = S.ph[pi & 1];
            if (is_matvec(ph.kind)) chain_phase<TYPE, QUAD, PROF>(P, pi & 1, ci, cta, G, tag, PROF && pi == P.prof_phase);
            phase_bar();
        }
        return;
    }

    // ---- consumers (thread / warp index among the consumers)
    const int ctid = tid - kC0, cwarp = ctid >> 5;
#pragma unroll 1
    for (int pi = 0; pi < P.n_phases; pi++) {
        const Phase &ph = S.ph[pi & 1];
        // prefetch the next descriptor with an ASYNC copy (a plain load + store would stall warps in order for one
        // L2 round trip at the start of every phase); waited for at the phase's end
        if (pi + 1 < P.n_phases && ctid < kPhWords)
            cp_async4(smem_u32(reinterpret_cast<uint32_t *>(&S.ph[(pi + 1) & 1]) + ctid),
                      reinterpret_cast<const uint32_t *>(&P.phases[pi + 1]) + ctid);
        if (P.trace && ctid == 0) S.tr[0] = S.tr[1] = (long long)gtime();
        const int kind = ph.kind;
        if (is_matvec(kind)) {
            matvec_consumer<TYPE, QUAD, PROF>(P, pi & 1, cta, G, tag, PROF && pi == P.prof_phase);
        } else if (kind == PH_SCORE) {
            score_phase<TYPE, PROF, QUAD>(P, pi & 1, cta, G, S.sched[4], tag, PROF && pi == P.prof_phase);
        } else if (kind == PH_PV) {
            pv_phase<TYPE, PROF, QUAD>(P, pi & 1, cta, G, S.sched[4], tag, PROF && pi == P.prof_phase);
        } else { // PH_SAMPLE: CTA 0 (of every rank) picks the first maximum of all ranks' per-CTA candidates
            if (cta == 0) { // all consumer warps: one candidate per thread, i.e. ONE round trip instead of one per 32
                const int nc = P.R * G * kNCH;
                Best b{-INFINITY, -1};
                for (int i = ctid; i < nc; i += kCT) {
                    ll_t w[2];
                    ll_gather<2>(w, P.cand_ll + i, nc, 2, tag);
                    b = better(b, Best{__uint_as_float((uint32_t)w[0]), (int)(uint32_t)w[1]});
                }
                int *redi = reinterpret_cast<int *>(S.red);
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    Best o;
                    o.v = __shfl_xor_sync(0xffffffffu, b.v, off);
                    o.i = __shfl_xor_sync(0xffffffffu, b.i, off);
                    b = better(b, o);
                }
                if (lane == 0) {
                    S.redf[cwarp] = b.v;
                    redi[cwarp] = b.i;
                }
                consumer_bar();
                if (cwarp == 0) {
                    b = lane < kCW ? Best{S.redf[lane], redi[lane]} : Best{-INFINITY, -1};
#pragma unroll
                    for (int off = 8; off >= 1; off >>= 1) {
                        Best o;
                        o.v = __shfl_xor_sync(0xffffffffu, b.v, off);
                        o.i = __shfl_xor_sync(0xffffffffu, b.i, off);
                        b = better(b, o);
                    }
                }
                if (cwarp == 0 && lane == 0) {
                    // every CTA has published its candidate, i.e. has long read state and seq.  The position stops at
                    // n_ctx - 1 (the KV cache, the rope table and history[n_ctx + 2] end there).
                    int np = P.state[1] + P.state[2];
                    np = min(max(np, 0), P.n_ctx - 1);
                    P.state[0] = b.i;
                    P.state[1] = np;
                    if (P.history) P.history[np] = b.i;
                    *P.seq = tag;
                }
            }
        }
        // the union region is rewritten by the next phase, and the prefetched descriptor is read
        cp_async_wait_all();
        phase_bar();
        if (P.trace && ctid == 0) {
            long long *t = P.trace + ((size_t)pi * G + cta) * 4;
            const long long tr2 = (long long)gtime();
            t[0] = S.tr[0]; t[1] = S.tr[1]; t[2] = tr2; t[3] = tr2;
        }
    }
}

// ---------------------------------------------------------------- pass-stream builder
struct RepackSrc {
    const uint8_t *qs;
    const __half *d;
    int row0; // first row of the matrix that belongs to the stream
    int rows; // rows of the matrix in the stream
};
struct RepackArgs {
    RepackSrc m[3];
    int n_mats;
    int gate_up;     // logical row r: group r / 64, j = r % 64: j < 32 -> m[0] row 32*group + j, else m[1] row 32*group + j - 32
    int n_lr;        // logical rows of the stream (rows past it are zero padding)
    int row_off;     // logical row of the stream's first row inside the concatenated matrices (contiguous row sharding)
    int nb;
    int quad, kf, kA;
    int unit_passes, unit_rows, uq, ur; // how units are dealt to CTAs (see unit_range)
    uint8_t *dst;
};
// one block per pass; a pass is four groups of 32 blocks: thread t < 128 copies block (t & 31) of group (t >> 5).
// Plain pass: groups = (row 0 | row 1) x (blocks 0-31 | 32-63) of one 64-block segment; quad pass: groups = rows 0..3,
// blocks = the short last segment.  A CTA's passes are segment-major (see Phase).
template <int TYPE> __global__ void repack_kernel(RepackArgs A) {
    using F = Fmt<TYPE>;
    const long long p = blockIdx.x;
    const int t = threadIdx.x;
    if (t >= 128) return;
    const int k = t >> 5, i = t & 31;
    // the CTA share this pass belongs to (rotated CTA index c: the first ur shares have uq + 1 units)
    const int unit = (int)(p / A.unit_passes);
    int c, u0;
    if (unit < A.ur * (A.uq + 1)) {
        c = unit / (A.uq + 1);
        u0 = c * (A.uq + 1);
    } else {
        c = A.ur + (unit - A.ur * (A.uq + 1)) / (A.uq > 0 ? A.uq : 1);
        u0 = A.ur * (A.uq + 1) + (c - A.ur) * A.uq;
    }
    const int nun = c < A.ur ? A.uq + 1 : A.uq;
    const int m = (int)(p - (long long)u0 * A.unit_passes);
    const int npairs = (nun * A.unit_rows) >> 1;
    const int row_base = u0 * A.unit_rows;
    int lr, blk;
    if (m < A.kA * npairs) {
        const int seg = m / npairs, pair = m % npairs;
        lr = row_base + 2 * pair + (k >> 1);
        blk = seg * kSeg + (k & 1) * 32 + i;
    } else {
        const int q = m - A.kA * npairs;
        lr = row_base + 4 * q + k;
        blk = A.kf * kSeg + i;
    }
    int mat = 0, row = lr + A.row_off;
    const bool valid = lr < A.n_lr && blk < A.nb;
    if (A.gate_up) {
        const int grp = row / 64, j = row % 64;
        mat = j < 32 ? 0 : 1;
        row = 32 * grp + (j & 31);
    } else {
        while (mat + 1 < A.n_mats && row >= A.m[mat].rows) {
            row -= A.m[mat].rows;
            mat++;
        }
        row += A.m[mat].row0;
    }
    uint8_t *pass = A.dst + (size_t)p * F::PASS_BYTES;
    uint4 *pay = reinterpret_cast<uint4 *>(pass + k * (32 * F::PB) + i * 16); // Q8_0: second half of the block at + 512 B
    __half *scl = reinterpret_cast<__half *>(pass + 2 * F::ROW_PB + k * 64) + i;
    if (valid) {
        const size_t b = (size_t)row * A.nb + blk;
        const uint4 *src = reinterpret_cast<const uint4 *>(A.m[mat].qs + b * F::PB);
        pay[0] = src[0];
        if (F::PB == 32) pay[32] = src[1];
        *scl = A.m[mat].d[b];
    } else {
        pay[0] = make_uint4(0, 0, 0, 0);
        if (F::PB == 32) pay[32] = make_uint4(0, 0, 0, 0);
        *scl = __float2half(0.0f);
    }
}

} // namespace

// ---------------------------------------------------------------- host side
struct ggb_decoder {
    DecoderParams P{};
    std::vector<Phase> phases;
    std::vector<void *> streams;
    Phase *d_phases = nullptr;
    void *scratch = nullptr;
    int grid = 0;
    size_t smem = 0;
    size_t stream_bytes = 0;
    std::vector<void *> peer_maps; // cudaIpcOpenMemHandle mappings of the other ranks' scratch
    bool any_quad = false;         // some phase uses the quad layout -> the QUAD instantiation of the kernel
};

static const void *decoder_fn(int type, bool quad, bool prof = false) {
    if (prof) { // the profiling instantiation (SM-clock sums inside one phase); Q4_0 only keeps the build short
        if (type == GGB_TYPE_Q4_0)
            return quad ? (const void *)decoder_kernel<GGB_TYPE_Q4_0, true, true> : (const void *)decoder_kernel<GGB_TYPE_Q4_0, false, true>;
    }
    if (type == GGB_TYPE_Q4_0)
        return quad ? (const void *)decoder_kernel<GGB_TYPE_Q4_0, true, false> : (const void *)decoder_kernel<GGB_TYPE_Q4_0, false, false>;
    return quad ? (const void *)decoder_kernel<GGB_TYPE_Q8_0, true, false> : (const void *)decoder_kernel<GGB_TYPE_Q8_0, false, false>;
}
static size_t decoder_smem(int type) {
    return type == GGB_TYPE_Q4_0 ? sizeof(Smem<GGB_TYPE_Q4_0>) : sizeof(Smem<GGB_TYPE_Q8_0>);
}

static void decoder_release(ggb_decoder *D) {
    for (void *pm : D->peer_maps) cudaIpcCloseMemHandle(pm);
    for (void *s : D->streams) cudaFree(s);
    cudaFree(D->d_phases);
    cudaFree(D->scratch);
    delete D;
}

// all ranks have finished every store into every peer's buffers / have mapped them: a tiny allgather + stream sync
static int tp_barrier(int size) {
    if (size <= 1) return GGB_OK;
    int *d = nullptr;
    GGB_CUDA(cudaMalloc(&d, sizeof(int) * (size_t)size));
    GGB_CUDA(cudaMemsetAsync(d, 0, sizeof(int) * (size_t)size, ggb_cuda_stream()));
    const int rc = ggb_op_allgather(d + ggb_comm_rank(), d, sizeof(int));
    cudaError_t e = cudaStreamSynchronize(ggb_cuda_stream());
    cudaFree(d);
    if (rc != GGB_OK) return rc;
    if (e != cudaSuccess) {
        ggb_set_error("decoder: tensor-parallel barrier -> %s", cudaGetErrorString(e));
        return GGB_ERR_CUDA;
    }
    return GGB_OK;
}

extern "C" {

// KV heads a rank's cache holds under the fused tensor-parallel lowering: its own heads when the head counts divide
// the world size (attention sharded by KV head), all of them otherwise (attention block replicated)
int ggb_decoder_kv_heads_local(int n_head, int n_head_kv, int size) {
    if (size > 1 && n_head_kv % size == 0 && n_head % size == 0) return n_head_kv / size;
    return n_head_kv;
}

int ggb_decoder_create(const ggb_decoder_dims *dims, const ggb_decoder_layer *layers, const ggb_weight *token_embd,
                       const float *output_norm, const ggb_decoder_io *io, ggb_decoder **out) {
    return ggb_decoder_create_tp(dims, layers, token_embd, output_norm, io, 0, 1, out);
}

int ggb_decoder_create_tp(const ggb_decoder_dims *dims, const ggb_decoder_layer *layers, const ggb_weight *token_embd,
                          const float *output_norm, const ggb_decoder_io *io, int rank, int size, ggb_decoder **out) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(size >= 1 && size <= 8 && rank >= 0 && rank < size, "bad tensor-parallel rank %d / size %d", rank, size);
    GGB_CHECK_ARG(dims && layers && token_embd && output_norm && io && out, "null pointer");
    const int E = dims->n_embd, L = dims->n_layer, H = dims->n_head, Hkv = dims->n_head_kv, d = dims->head_dim;
    const int F = dims->n_ff, V = dims->n_vocab;
    const int type = token_embd->type;
    auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    if (!(E % 256 == 0 && E <= kMaxE && (H * d) % 256 == 0 && H * d <= kMaxE && F % 256 == 0 && F <= kMaxK &&
          d >= 32 && d <= 256 && pow2(d) && pow2(Hkv) && H % Hkv == 0 && pow2(H / Hkv) && H / Hkv <= 8 &&
          dims->n_ctx <= kMaxCtx && (type == GGB_TYPE_Q4_0 || type == GGB_TYPE_Q8_0))) {
        ggb_set_error("decoder: shape outside the fused kernel's envelope (E=%d H=%d Hkv=%d d=%d F=%d ctx=%d)", E, H, Hkv,
                      d, F, dims->n_ctx);
        return GGB_ERR_UNSUPPORTED;
    }
    GGB_CHECK_ARG(ggb_rope_cos_dev() && ggb_rope_n_dims() == d && ggb_rope_n_ctx() >= dims->n_ctx,
                  "rope table not initialised for head_dim %d / ctx %d", d, dims->n_ctx);
    ggb_decoder *D = new ggb_decoder();
    DecoderParams &P = D->P;
    P.type = type; P.E = E; P.H = H; P.Hkv = Hkv; P.d = d; P.F = F; P.V = V; P.n_ctx = dims->n_ctx;
    P.eps = dims->eps;
    P.q_scale = 1.0f / sqrtf((float)d);
    P.emb_scale = sqrtf((float)E);
    P.emb_qs = token_embd->qs; P.emb_d = token_embd->d;
    P.state = io->state; P.history = io->history;
    P.exp_table = ggb_exp_table_dev(); P.gelu_table = ggb_gelu_table_dev();
    P.rope_cos = ggb_rope_cos_dev(); P.rope_sin = ggb_rope_sin_dev();
    {   // does the computed exp reproduce the table?  (GGB_EXP_TABLE=1 forces the lookup)
        int *d_mis = nullptr, h_mis[3] = {-1, -1, -1};
        const char *force = getenv("GGB_EXP_TABLE");
        if (!(force && atoi(force)) && cudaMalloc(&d_mis, 3 * sizeof(int)) == cudaSuccess) {
            cudaMemsetAsync(d_mis, 0, 3 * sizeof(int), ggb_cuda_stream());
            exp_check_kernel<<<256, 256, 0, ggb_cuda_stream()>>>(P.exp_table, P.gelu_table, d_mis);
            if (cudaMemcpyAsync(h_mis, d_mis, 3 * sizeof(int), cudaMemcpyDeviceToHost, ggb_cuda_stream()) != cudaSuccess ||
                cudaStreamSynchronize(ggb_cuda_stream()) != cudaSuccess)
                h_mis[0] = h_mis[1] = h_mis[2] = -1;
            cudaFree(d_mis);
        }
        P.gelu_compute = h_mis[2] == 0 ? 1 : 0;
        P.exp_compute = h_mis[1] == 0 ? 2 : (h_mis[0] == 0 ? 1 : 0);
        if (const char *e2 = getenv("GGB_EXP_MODE")) P.exp_compute = std::min(P.exp_compute, std::max(atoi(e2), 0));
    }
    P.rank = rank; P.R = size;
    auto ilog2 = [](int v) { int s = 0; while ((1 << s) < v) s++; return s; };
    P.Gq = H / Hkv; P.hsplit = P.Gq > 4 ? 2 : 1; P.Gh = P.Gq / P.hsplit;
    P.hd_shift = ilog2(d / 2); P.gq_shift = ilog2(P.Gq);
    // tensor parallel: the attention block (QKV, SCORE, PV, O) is sharded by KV head when the head counts divide the
    // world size; otherwise (MQA) every rank runs all of it - a few MB per layer cost less than the cross-GPU hops
    const bool attn_sharded = size > 1 && ggb_decoder_kv_heads_local(H, Hkv, size) != Hkv;
    P.ng = attn_sharded ? Hkv / size : Hkv; P.g0 = attn_sharded ? rank * P.ng : 0; P.ng_shift = ilog2(P.ng);
    P.nh = P.ng * P.Gq; P.h0 = P.g0 * P.Gq;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    D->grid = sms;
    P.pv_w = (P.nh * (d / 16) > sms && d >= 32) ? 32 : 16;
    P.pv_oct_shift = ilog2(P.pv_w / 8); P.pv_slices = d / P.pv_w; P.pv_sl_shift = ilog2(P.pv_slices);
    // ring copies in flight: 2 for MQA models (bounds the queueing delay the flag polls see; measured best for Gemma-2B),
    // the whole ring for MHA models, whose long attention phases leave the stream more to catch up (Gemma-7B: + 1.2 %)
    P.max_outstanding = Hkv > 1 ? kSlots : kMaxOutstanding; P.prefetch_chunks = kPrefetchChunks;
    if (const char *e = getenv("GGB_MAX_OUTSTANDING")) P.max_outstanding = std::min(std::max(atoi(e), 1), kSlots);
    if (const char *e = getenv("GGB_PREFETCH_CHUNKS")) P.prefetch_chunks = std::min(std::max(atoi(e), 0), 64);
    // scratch: flagged buffers, one set per layer (tag = token sequence number, so a word of the previous
    // token never matches and nothing is rewritten while another CTA may still read it)
    const size_t QD = (size_t)H * d, KD = (size_t)Hkv * d;
    size_t off = 0;
    auto carve = [&](size_t words) { size_t o = off; off = (off + words * 8 + 255) & ~(size_t)255; return o; };
    const size_t o_x = carve((size_t)L * E);                       // slot il = output of layer il
    const size_t o_qkv = carve((size_t)L * (QD + 2 * KD));
    const size_t o_s = carve((size_t)L * H * dims->n_ctx);
    const size_t o_kqv = carve((size_t)L * QD), o_sa = carve((size_t)L * E);
    const size_t o_ff = carve((size_t)L * (F / 4 + F / 32));
    const size_t o_cand = carve((size_t)2 * sms * size * kNCH), o_seq = carve(1);
    cudaError_t e = cudaMalloc(&D->scratch, off);
    if (e != cudaSuccess) {
        ggb_set_error("decoder scratch cudaMalloc(%zu) -> %s", off, cudaGetErrorString(e));
        delete D;
        return GGB_ERR_CUDA;
    }
    cudaMemsetAsync(D->scratch, 0, off, ggb_cuda_stream());
    for (int r = 0; r < 8; r++) P.peer_delta[r] = 0;
    if (size > 1) {
        // every rank maps every other rank's flagged buffers (cudaIpc over NVLink peer access); the 64-byte handles
        // travel through the NCCL communicator, which also orders the memset above before any peer store
        cudaIpcMemHandle_t mine;
        void *d_h = nullptr;
        int rc = GGB_OK;
        std::vector<cudaIpcMemHandle_t> all((size_t)size);
        cudaError_t ce = cudaIpcGetMemHandle(&mine, D->scratch);
        if (ce == cudaSuccess) ce = cudaMalloc(&d_h, sizeof(mine) * (size_t)size);
        if (ce == cudaSuccess)
            ce = cudaMemcpyAsync((char *)d_h + sizeof(mine) * (size_t)rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, ggb_cuda_stream());
        if (ce == cudaSuccess) rc = ggb_op_allgather((char *)d_h + sizeof(mine) * (size_t)rank, d_h, sizeof(mine));
        if (ce == cudaSuccess && rc == GGB_OK)
            ce = cudaMemcpyAsync(all.data(), d_h, sizeof(mine) * (size_t)size, cudaMemcpyDeviceToHost, ggb_cuda_stream());
        if (ce == cudaSuccess && rc == GGB_OK) ce = cudaStreamSynchronize(ggb_cuda_stream());
        cudaFree(d_h);
        if (ce != cudaSuccess || rc != GGB_OK) {
            if (ce != cudaSuccess) ggb_set_error("decoder: peer handle exchange -> %s", cudaGetErrorString(ce));
            decoder_release(D);
            return ce != cudaSuccess ? GGB_ERR_CUDA : rc;
        }
        int slot = 1; // peer_delta[0] = self; the others in rank order
        for (int r = 0; r < size; r++) {
            if (r == rank) continue;
            void *pm = nullptr;
            cudaError_t ie = cudaIpcOpenMemHandle(&pm, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess);
            if (ie != cudaSuccess) {
                ggb_set_error("decoder: cudaIpcOpenMemHandle(rank %d) -> %s", r, cudaGetErrorString(ie));
                decoder_release(D);
                return GGB_ERR_CUDA;
            }
            D->peer_maps.push_back(pm);
            P.peer_delta[slot++] = (long long)((char *)pm - (char *)D->scratch);
        }
    }
    uint8_t *sb = (uint8_t *)D->scratch;
    ll_t *x_ll = (ll_t *)(sb + o_x), *qkv_ll = (ll_t *)(sb + o_qkv), *s_ll = (ll_t *)(sb + o_s);
    ll_t *kqv_ll = (ll_t *)(sb + o_kqv), *sa_ll = (ll_t *)(sb + o_sa);
    ll_t *ff_ll = (ll_t *)(sb + o_ff);
    P.cand_ll = (ll_t *)(sb + o_cand);
    P.seq = (uint32_t *)(sb + o_seq);

    const int pass_bytes = type == GGB_TYPE_Q4_0 ? Fmt<GGB_TYPE_Q4_0>::PASS_BYTES : Fmt<GGB_TYPE_Q8_0>::PASS_BYTES;
    bool ok = true;
    const char *why = "";
    // builds the pass stream of one mat-vec phase and fills the schedule fields.  shard: 0 = every rank computes all
    // rows; 1 = contiguous row blocks per rank; 2 = QKV by head (rows of this rank's query / KV heads)
    auto matvec_phase = [&](int kind, int layer, int n, const ggb_weight *const *w, bool gate_up, int shard) -> Phase {
        Phase ph{};
        ph.kind = kind; ph.layer = layer;
        const int nb = (int)w[0]->nb;
        ph.K = (int)w[0]->cols;
        const int nbp = nb % kSeg;                       // blocks of the short last segment (0: none)
        ph.quad = (nbp > 0 && nbp <= 32) ? 1 : 0;        // it fits one group of a pass: quad layout
        ph.kf = nb / kSeg;
        ph.ksegs = (nb + kSeg - 1) / kSeg;
        ph.kA = ph.quad ? ph.kf : ph.ksegs;
        ph.tail_blocks = ph.quad ? nbp : nb - kSeg * (ph.ksegs - 1);
        RepackArgs A{};
        A.n_mats = n; A.gate_up = gate_up ? 1 : 0; A.nb = nb; A.quad = ph.quad; A.kf = ph.kf; A.kA = ph.kA;
        int rows = 0;
        const int row_mult = ph.quad ? 4 : 2;
        for (int i = 0; i < n; i++) {
            A.m[i].qs = w[i]->qs; A.m[i].d = w[i]->d; A.m[i].row0 = 0; A.m[i].rows = (int)w[i]->rows;
            if (w[i]->nb != nb || w[i]->type != type) { ok = false; why = "mixed shapes/types in one phase"; }
        }
        if (shard == 2) { // this rank's heads of Q, K, V
            A.m[0].row0 = P.h0 * d; A.m[0].rows = P.nh * d;
            A.m[1].row0 = P.g0 * d; A.m[1].rows = P.ng * d;
            A.m[2].row0 = P.g0 * d; A.m[2].rows = P.ng * d;
            ph.seg_rows[0] = P.nh * d; ph.seg_off[0] = P.h0 * d;
            ph.seg_rows[1] = P.ng * d; ph.seg_off[1] = (int)QD + P.g0 * d;
            ph.seg_rows[2] = P.ng * d; ph.seg_off[2] = (int)(QD + KD) + P.g0 * d;
        }
        for (int i = 0; i < n; i++) {
            rows += A.m[i].rows;
            if (i + 1 < n && !gate_up && (A.m[i].rows % row_mult)) { ok = false; why = "row count inside a fused phase not a multiple of the unit"; }
        }
        int u_total;
        if (gate_up) {
            u_total = rows / 64; ph.unit_rows = 64;
            if (rows % 64) { ok = false; why = "n_ff not a multiple of 32"; }
        } else if (ph.quad) {
            u_total = (rows + 3) / 4; ph.unit_rows = 4;
        } else {
            u_total = (rows + 1) / 2; ph.unit_rows = 2;
        }
        ph.unit_passes = (ph.unit_rows / 2) * ph.kA + (ph.quad ? ph.unit_rows / 4 : 0);
        // contiguous row sharding: units are split by rank first (balanced), then by CTA
        ph.u_off = shard == 1 ? (int)(((long long)rank * u_total) / size) : 0;
        ph.n_units = shard == 1 ? (int)(((long long)(rank + 1) * u_total) / size) - ph.u_off : u_total;
        ph.n_lr = std::max(0, std::min(ph.n_units * ph.unit_rows, rows - ph.u_off * ph.unit_rows));
        if (shard != 2) {
            ph.seg_rows[0] = 1 << 30; ph.seg_off[0] = ph.u_off * ph.unit_rows;
        }
        ph.bcast = (shard == 1 && size > 1) ? 1 : 0;
        ph.uq = ph.n_units / sms; ph.ur = ph.n_units % sms;
        const long long max_units = (ph.n_units + sms - 1) / sms;
        const long long max_rows = max_units * ph.unit_rows;
        const bool multi = ph.kA > 1 || (ph.quad && ph.kf > 0); // rows keep an accumulator between slots
        if (gate_up && max_rows > kGuFloats) { ok = false; why = "GATE_UP rows per CTA do not fit the staging"; }
        if (multi && max_rows > (ph.K <= kAccMaxK ? kAccFloats : kAccSmall)) { ok = false; why = "rows per CTA do not fit the accumulators"; }
        if (gate_up && ph.K > kAccMaxK) { ok = false; why = "GATE_UP activation row too long"; }
        if (terms_nbuf(ph.K) < 2) { ok = false; why = "no room for two term buffers"; }
        if (ph.K > kMaxK) { ok = false; why = "activation row too long"; }
        if (!ok) return ph;
        const size_t n_passes = (size_t)ph.n_units * ph.unit_passes; // this rank's passes only
        A.n_lr = ph.n_lr; A.row_off = shard == 1 ? ph.u_off * ph.unit_rows : 0;
        A.unit_passes = ph.unit_passes; A.unit_rows = ph.unit_rows; A.uq = ph.uq; A.ur = ph.ur;
        if (n_passes == 0) return ph;
        void *stream = nullptr;
        cudaError_t ce = cudaMalloc(&stream, n_passes * pass_bytes);
        if (ce != cudaSuccess) { ok = false; why = "pass stream cudaMalloc failed"; return ph; }
        D->streams.push_back(stream);
        D->stream_bytes += n_passes * pass_bytes;
        A.dst = (uint8_t *)stream;
        if (type == GGB_TYPE_Q4_0) repack_kernel<GGB_TYPE_Q4_0><<<(unsigned)n_passes, 128, 0, ggb_cuda_stream()>>>(A);
        else repack_kernel<GGB_TYPE_Q8_0><<<(unsigned)n_passes, 128, 0, ggb_cuda_stream()>>>(A);
        ph.stream = (const uint8_t *)stream;
        return ph;
    };
    const int attn_shard = attn_sharded ? 2 : 0;
    for (int il = 0; il < L && ok; il++) {
        const ggb_decoder_layer &ly = layers[il];
        const ll_t *x_in = il == 0 ? nullptr : x_ll + (size_t)(il - 1) * E;
        float *sa = io->sa_out + (size_t)il * E, *lo = io->l_out + (size_t)il * E;
        ll_t *qkv_l = qkv_ll + (size_t)il * (QD + 2 * KD);
        ll_t *s_l = s_ll + (size_t)il * H * dims->n_ctx;
        ll_t *kqv_l = kqv_ll + (size_t)il * QD, *sa_l = sa_ll + (size_t)il * E;
        ll_t *ff_l = ff_ll + (size_t)il * (F / 4 + F / 32);
        {   // QKV
            const ggb_weight *w[3] = {ly.attn_q, ly.attn_k, ly.attn_v};
            Phase ph = matvec_phase(PH_QKV, il, 3, w, false, attn_shard);
            ph.in_ll = x_in; ph.norm_w = ly.attn_norm; ph.out_ll = qkv_l;
            if (ph.K != E) { ok = false; why = "attn_q input width != n_embd"; }
            if ((int)ly.attn_q->rows != (int)QD || (int)ly.attn_k->rows != (int)KD || (int)ly.attn_v->rows != (int)KD) { ok = false; why = "attn_q/k/v row counts"; }
            D->phases.push_back(ph);
        }
        {   // SCORE
            Phase ph{};
            ph.kind = PH_SCORE; ph.layer = il; ph.kcache = (__half *)ly.k_cache; ph.vcache = (__half *)ly.v_cache;
            ph.in_ll = qkv_l; ph.s_ll = s_l;
            D->phases.push_back(ph);
        }
        {   // PV
            Phase ph{};
            ph.kind = PH_PV; ph.layer = il; ph.vcache = (__half *)ly.v_cache; ph.in_ll = qkv_l; ph.s_ll = s_l;
            ph.f32_out = io->kqv + (size_t)il * QD; ph.out_ll = kqv_l; ph.bcast = attn_sharded ? 1 : 0;
            D->phases.push_back(ph);
        }
        {   // O + residual
            const ggb_weight *w[1] = {ly.attn_output};
            Phase ph = matvec_phase(PH_O, il, 1, w, false, attn_sharded ? 1 : 0);
            ph.in_ll = kqv_l; ph.resid_ll = x_in; ph.resid_embed = il == 0 ? 1 : 0; ph.out = sa; ph.out_ll = sa_l;
            if (ph.K != (int)QD) { ok = false; why = "attn_output input width != n_head*head_dim"; }
            D->phases.push_back(ph);
        }
        {   // GATE_UP
            const ggb_weight *w[2] = {ly.ffn_gate, ly.ffn_up};
            Phase ph = matvec_phase(PH_GATE_UP, il, 2, w, true, size > 1 ? 1 : 0);
            ph.in_ll = sa_l; ph.norm_w = ly.ffn_norm; ph.out_ll = ff_l;
            ph.f32_out = io->ffn_act ? io->ffn_act + (size_t)il * F : nullptr;
            if (ph.K != E) { ok = false; why = "ffn_gate input width != n_embd"; }
            D->phases.push_back(ph);
        }
        {   // DOWN + residual
            const ggb_weight *w[1] = {ly.ffn_down};
            Phase ph = matvec_phase(PH_DOWN, il, 1, w, false, size > 1 ? 1 : 0);
            ph.in_ll = ff_l; ph.resid_ll = sa_l; ph.out = lo; ph.out_ll = x_ll + (size_t)il * E;
            if (ph.K != F) { ok = false; why = "ffn_down input width != n_ff"; }
            D->phases.push_back(ph);
        }
    }
    if (ok) {   // LOGITS (tied embedding) + SAMPLE
        const ggb_weight *w[1] = {token_embd};
        Phase ph = matvec_phase(PH_LOGITS, L, 1, w, false, size > 1 ? 1 : 0);
        ph.in_ll = x_ll + (size_t)(L - 1) * E; ph.norm_w = output_norm; ph.f32_out = io->result_norm;
        ph.out = io->logits;
        D->phases.push_back(ph);
        Phase sp{};
        sp.kind = PH_SAMPLE; sp.layer = L;
        D->phases.push_back(sp);
    }
    if (!ok) {
        ggb_set_error("decoder: %s", why);
        cudaStreamSynchronize(ggb_cuda_stream());
        decoder_release(D);
        return GGB_ERR_UNSUPPORTED;
    }
    for (size_t i = 0; i < D->phases.size(); i++) D->phases[i].rot = (int)((i * 61) % (size_t)sms);
    e = cudaMalloc(&D->d_phases, D->phases.size() * sizeof(Phase));
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(D->d_phases, D->phases.data(), D->phases.size() * sizeof(Phase), cudaMemcpyHostToDevice,
                            ggb_cuda_stream());
    if (e != cudaSuccess) {
        ggb_set_error("decoder phase table upload -> %s", cudaGetErrorString(e));
        decoder_release(D);
        return GGB_ERR_CUDA;
    }
    P.phases = D->d_phases;
    P.n_phases = (int)D->phases.size();
    D->smem = decoder_smem(type);
    for (const Phase &ph : D->phases) D->any_quad = D->any_quad || ph.quad;
    const void *fn = decoder_fn(type, D->any_quad);
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D->smem);
    int occ = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kThreads, D->smem);
    if (e != cudaSuccess || occ < 1) {
        ggb_set_error("decoder kernel cannot be resident (smem %zu B): %s", D->smem,
                      e != cudaSuccess ? cudaGetErrorString(e) : "occupancy 0");
        decoder_release(D);
        return GGB_ERR_CUDA;
    }
    e = cudaStreamSynchronize(ggb_cuda_stream());
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        ggb_set_error("decoder pass-stream build -> %s", cudaGetErrorString(e));
        decoder_release(D);
        return GGB_ERR_CUDA;
    }
    // every rank has mapped every peer's buffers (and zeroed its own) before the first token pushes words into them
    const int brc = tp_barrier(size);
    if (brc != GGB_OK) {
        decoder_release(D);
        return brc;
    }
    *out = D;
    return GGB_OK;
}

int ggb_decoder_step(ggb_decoder *D) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(D != nullptr, "null decoder");
    void *args[] = {(void *)&D->P};
    const void *fn = decoder_fn(D->P.type, D->any_quad, D->P.prof != nullptr);
    // cooperative launch: the flagged data flow needs every CTA resident (a waiting CTA spins).  The fp16
    // exp / gelu tables are pinned in the persisting part of L2 for this kernel.
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)D->grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = D->smem;
    cfg.stream = ggb_cuda_stream();
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    at[1].id = cudaLaunchAttributeAccessPolicyWindow;
    at[1].val.accessPolicyWindow.base_ptr = (void *)ggb_exp_table_dev();
    at[1].val.accessPolicyWindow.num_bytes = ggb_tables_bytes();
    at[1].val.accessPolicyWindow.hitRatio = 1.0f;
    at[1].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    at[1].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cfg.attrs = at;
    cfg.numAttrs = 2;
    GGB_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
    ggb_count_launch();
    return GGB_OK;
}

// one traced step: per phase and CTA four %globaltimer stamps in ns (start, after the prologue = inputs
// arrived, end, end) plus the phase kinds; the cudaEvent/ncu numbers say how long, this says where.
int ggb_decoder_trace(ggb_decoder *D, long long *stamps, int32_t *kinds, int max_phases, int *n_phases, int *grid) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(D && stamps && kinds && n_phases && grid, "null pointer");
    const int np = (int)D->phases.size();
    GGB_CHECK_ARG(max_phases >= np, "trace buffer holds %d phases, need %d", max_phases, np);
    const size_t bytes1 = (size_t)np * D->grid * 4 * sizeof(long long);
    const size_t bytes = bytes1 + (size_t)D->grid * 64 * sizeof(long long); // + [grid][32] sums inside phase GGB_PROF_PHASE
    GGB_CHECK_ARG((size_t)max_phases * D->grid * 4 * sizeof(long long) >= bytes, "trace buffer too small");
    long long *dev = nullptr;
    GGB_CUDA(cudaMalloc(&dev, bytes));
    GGB_CUDA(cudaMemsetAsync(dev, 0, bytes, ggb_cuda_stream()));
    D->P.trace = dev;
    if (const char *env = getenv("GGB_PROF_PHASE")) {
        if (D->P.type == GGB_TYPE_Q4_0 && atoi(env) >= 0) {
            D->P.prof = dev + bytes1 / sizeof(long long);
            D->P.prof_phase = atoi(env);
            cudaFuncSetAttribute(decoder_fn(D->P.type, D->any_quad, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D->smem);
        }
    }
    int rc = ggb_decoder_step(D);
    D->P.trace = nullptr;
    D->P.prof = nullptr;
    if (rc == GGB_OK) {
        cudaError_t e = cudaMemcpyAsync(stamps, dev, bytes, cudaMemcpyDeviceToHost, ggb_cuda_stream());
        if (e == cudaSuccess) e = cudaStreamSynchronize(ggb_cuda_stream());
        if (e != cudaSuccess) {
            ggb_set_error("decoder trace -> %s", cudaGetErrorString(e));
            rc = GGB_ERR_CUDA;
        }
    }
    cudaFree(dev);
    for (int i = 0; i < np; i++) kinds[i] = D->phases[(size_t)i].kind;
    *n_phases = np;
    *grid = D->grid;
    return rc;
}

int ggb_decoder_exp_computed(const ggb_decoder *D) { return D ? D->P.exp_compute : 0; }

// bytes of pass stream THIS rank reads per token (its shards + the phases it runs replicated)
long long ggb_decoder_stream_bytes(const ggb_decoder *D) { return D ? (long long)D->stream_bytes : 0; }
// KV heads whose cache rows this rank reads per token
int ggb_decoder_kv_heads(const ggb_decoder *D) { return D ? D->P.ng : 0; }

int ggb_decoder_free(ggb_decoder *D) {
    if (!D) return GGB_OK;
    cudaStreamSynchronize(ggb_cuda_stream());
    // a faster rank must not unmap / free its buffers while a peer's last token still pushes words into them
    tp_barrier(D->P.R);
    decoder_release(D);
    return GGB_OK;
}

} // extern "C"
