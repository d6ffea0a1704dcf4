// decoder.cu — the persistent decode kernel: ONE cooperative launch per token (v4: pass stream + flagged data flow).
//
// Why: a Gemma-2B Q4_0 token is 1.41 GB of weights + 38 MB of KV = ~220 us of HBM time on a B200, spread
// over 163 mat-vecs and ~450 small nodes (ref: src/gemma_model.cpp:688-740).  One CTA per SM stays resident
// for the whole token; what limits it is (a) how cleanly the weights stream and (b) what every phase change
// costs.  Design rules, each measured (tools/ubench_stream.cu, ubench_bar.cu, ubench_dp4a.cu, tools_trace.py, tools/ab.sh):
//
//   * PASS STREAM.  At create time every weight matrix is re-laid out as a stream of PASSES: one pass =
//     two rows x one 64-block K segment (2304 B Q4_0 / 4352 B Q8_0), payloads then scales, i.e. exactly what
//     one consumer warp dots in one go.  A phase's passes are contiguous in HBM in schedule order, so a CTA's
//     share of a phase is ONE contiguous byte range and the producer is a trivial loop of bulk copies.
//   * PRODUCER = lane 0 of warp 16: cp.async.bulk (TMA, SASS UBLKCP) of up to 16 passes (36 KB) per ring slot,
//     5 slots, completion on mbarriers.  Weights do not depend on activations, so it runs AHEAD across phase
//     boundaries: while consumers wait for their inputs, HBM keeps streaming.
//   * CONSUMERS = 16 warps.  Warp w owns pass w of every slot: 4 x LDS.128 weights + the two activation
//     blocks of its segment from shared memory, dp4a, the scalar reference's per-block fp32 term, a folded
//     xor-butterfly for the two rows.  Segment sums go to a shared-memory partial table; ONE block barrier per
//     phase, then an epilogue finishes rows (segments in order, residual, GeGLU + Q8_0, arg-max).
//     The stand-alone loop sustains 6.0-6.2 TB/s of 6.6 TB/s read peak (tools/ubench_stream.cu).
//   * NO GRID BARRIERS.  A grid barrier costs 1.4 us on B200 (0.6 us of it the release fence,
//     tools/ubench_bar.cu) and a token has 108 phase changes.  Instead every value that crosses CTAs travels
//     as an 8-byte word {payload, tag} written with one store (the NCCL "LL" idea): tag = the token's sequence
//     number, the buffers are per layer, and a consumer simply re-reads a word until its tag matches.  Data
//     and flag arrive together, so no fence and no separate wait: a phase change costs one L2 write + one L2
//     read.  Each phase's prologue polls exactly the inputs it needs.
//   * ISSUE-BOUND, AT THE REGISTER LIMIT.  The pass loop is bound by instruction issue (IDP.4A every 2 cycles per SM
//     sub-partition, ~185 instructions per pass: ~85 GB/s per SM against the 44 GB/s HBM delivers), the phases
//     between the mat-vecs by chains of ~30-cycle steps.  17 warps per SM cap the kernel at 96 registers, shared
//     memory leaves no L1, so a spill is an L2 round trip: token-constant scalars live in shared memory, every hot
//     LDS.128 is conflict-free (two-plane layouts), no integer division at a phase start, the soft-max exp is
//     computed (proved equal to the reference's fp16 table at create time).  Instruction-cache warmth is NOT a
//     factor (a warm second run of a phase takes the same time).
//   * TEMPLATE SWITCHES.  QUAD: models with a short last K segment (Gemma-7B, K = 3072) share one pass between four
//     rows for it; GROUPS = 2: MHA models at long context run two ATTN items per CTA in two 8-warp groups.
//
// Phases per layer: QKV | ATTN (split-K partials; rope, Q scale and the KV-cache store of the new position
// folded in) | COMBINE | O (+ residual) | GATE_UP (+ GeGLU + Q8_0 quantise) | DOWN (+ residual);
// then LOGITS (+ per-CTA arg-max) | SAMPLE.
// Numerics are bit-identical to the per-operator kernels (same DEV order; tests compare both and the oracle).
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace {

constexpr int kCW = 16;             // consumer warps
constexpr int kCT = kCW * 32;       // consumer threads
constexpr int kThreads = kCT + 32;  // + producer warp
constexpr int kSlots = 5;
constexpr int kMaxOutstanding = 2;  // ring copies in flight per SM (default; GGB_MAX_OUTSTANDING overrides)
constexpr int kPrefetchChunks = 0;  // L2 look-ahead of the weight stream in ring chunks (GGB_PREFETCH_CHUNKS overrides)
constexpr int kChunk = GGB_DEV_ATTN_CHUNK;
constexpr int kSeg = GGB_DEV_SEG_BLOCKS; // 64 blocks per K segment
constexpr int kMaxG = 8;
constexpr int kMaxE = 4096;         // longest norm-fed activation row
constexpr int kMaxK = 24576;        // longest activation row kept in shared memory
constexpr int kPartFloats = 4096;   // partial table: (rows of this CTA in a phase) x (K segments)
constexpr int kMaxChunks = 256;     // split-K chunks per head the COMBINE staging holds (n_ctx <= 8192)

enum PhaseKind { PH_QKV = 0, PH_ATTN, PH_COMBINE, PH_O, PH_GATE_UP, PH_DOWN, PH_LOGITS, PH_SAMPLE };

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
    int n_units;     // schedulable units (row pairs; GATE_UP: 32-row groups of gate + up)
    int unit_passes; // passes per unit
    int unit_rows;   // logical rows per unit
    int ksegs;       // K segments per row
    int K;           // activation length
    int n_rows;      // logical rows of the phase
    int rot;         // CTA rotation so remainders do not always land on the same SMs
    int uq, ur;      // n_units / grid, n_units % grid
    // how a warp's position advances from pass to pass (it takes every 16th pass of the CTA's share), all from the
    // host so that no division runs on the critical path.  Position = (seg, pidx): index inside the unit of `wrapn`
    // passes, and the unit's base in the partial table.  Per pass: pidx += dstep, seg += sstep, and on seg >= wrapn:
    // seg -= wrapn, pidx += winc.  Plain phases: unit = row pair (wrapn = ksegs passes, pidx includes seg).
    // QUAD phases (K % 2048 in (0, 1024]: the last K segment has <= 32 blocks): unit = 4 rows =
    // kf passes of rows 0-1 | kf passes of rows 2-3 | ONE pass holding the short segment of all 4 rows, so the
    // short segment costs half a pass per row pair instead of a zero-padded full one (Gemma-7B: K = 3072).
    int sstep, dstep, wrapn, winc, winit;
    int quad, kf;    // quad layout; number of full 64-block segments per row
    int u_off;       // first unit of this rank (tensor parallel: units are split by rank, then by CTA)
    int pad2_;
    int resid_embed; // residual = embedding row of the token (layer 0, O phase)
    const uint8_t *stream; // pass stream of the phase
    // flagged inputs: QKV / GATE_UP / LOGITS: fp32 x [E] (null in layer 0: embedding row); O: fp32 kqv [H*d];
    // DOWN: packed Q8_0 (4 codes per word) [F/4] then fp16 scales [F/32]; ATTN: fp32 q | k | v
    const ll_t *in_ll;
    const float *norm_w;
    ll_t *out_ll;          // flagged output (QKV: qkv; O: sa; DOWN: next layer's x; GATE_UP: packed Q8_0; COMBINE: kqv)
    float *out;            // plain fp32 copy for tensor_dump-style captures / logits (may be null)
    const ll_t *resid_ll;  // flagged residual
    float *f32_out;        // COMBINE: kqv dump; GATE_UP: optional ffn_act dump; LOGITS: result_norm
    __half *kcache, *vcache;
    ll_t *acc_ll, *m_ll, *l_ll; // ATTN -> COMBINE split-K partials of this layer
};
static_assert(sizeof(Phase) % 4 == 0, "Phase is copied word-wise");
constexpr int kPhWords = (int)(sizeof(Phase) / 4);

struct DecoderParams {
    const Phase *phases;
    int n_phases;
    int type;
    int E, H, Hkv, d, F, V, n_ctx;
    // host-precomputed so that no phase starts with a chain of integer divisions (each ~100 cycles of dependent
    // MUFU / I2F / uniform-register round trips): query heads per KV head, the ATTN head split, log2 of d/2, d/32, Hkv
    int Gq, hsplit, Gh, hd_shift, sl_shift, kvh_shift;
    float eps, q_scale, emb_scale;
    const uint8_t *emb_qs; // planar embedding table (layer-0 get_rows)
    const __half *emb_d;
    int32_t *state, *history;
    int max_chunks;
    ll_t *cand_ll;         // [2][R * grid]: per-CTA arg-max candidates (value, index) of every rank
    int max_outstanding, prefetch_chunks; // producer tuning (see producer_loop)
    int rank, R;           // tensor parallel: this GPU's rank and the number of GPUs (1 = single GPU)
    long long peer_delta[8]; // byte offset from a local flagged word to the same word in rank r's memory (0 for self)
    uint32_t *seq;         // token sequence number = tag of the previous launch
    const uint16_t *exp_table, *gelu_table;
    int exp_compute;       // 1: exp_f16_computed reproduces exp_table bit for bit (checked at create)
    const float *rope_cos, *rope_sin;
    long long *trace; // optional [n_phases][grid][4] %globaltimer stamps (ns): start, after prologue (inputs arrived), end, end
    long long *dbg;   // optional [grid][16] SM-clock stamps inside phase dbg_phase
    int dbg_phase;
    int dbg_repeat; // debug: run ATTN / COMBINE this many times (idempotent), stamps from the last run (warm caches)
};

// union region: the views are never live at the same time
//   mat-vec : act_q [kMaxK] | act_d half[kMaxK/32] | bias int[kMaxK/32] | part float[kPartFloats]
//   norm    : xs float[kMaxE] at +8192 (act_q only uses [0, kMaxE) in a norm-fed phase)
//   attn    : Kc half[32][256] | Vc half[32][256] | Sq float[4][256] | s float[4][32] | e float[32][4]
//   combine : stage float[kMaxChunks][32] | m float[kMaxChunks] | l float[kMaxChunks] | w float[kMaxChunks]
constexpr int U_ACT_D = kMaxK;
constexpr int U_BIAS = U_ACT_D + kMaxK / 32 * 2;
constexpr int U_PART = U_BIAS + kMaxK / 32 * 4;
constexpr int U_XS = 8192;
constexpr int U_KC = 0, U_VC = 16384, U_SQ = 32768, U_S = 36864, U_E = 37888;
constexpr int U_CM = kMaxChunks * 128, U_CL = U_CM + kMaxChunks * 4, U_CW = U_CL + kMaxChunks * 4;
constexpr int U_BYTES = U_PART + kPartFloats * 4;
static_assert(U_E + 1024 <= U_BYTES, "union region too small for the attention view");
static_assert(U_XS + kMaxE * 4 <= U_ACT_D && kMaxE <= U_XS, "xs overlaps the activation row");
static_assert(U_CW + kMaxChunks * 4 <= U_BYTES, "combine staging too large");

// GROUPS: how many ATTN items a CTA works on at once.  1 (MQA models): all 16 warps on one item, weight ring of 5 slots.
// 2 (MHA models: many one-head items per CTA, each a ~3 us latency chain): two groups of 8 warps, each with its own K/V
// buffers, barrier and item; the ring gives up one slot for the second attention view.
constexpr int kAttnView = U_E + 1024; // bytes of one group's attention view
template <int TYPE, int GROUPS> struct Smem {
    static constexpr int SLOTS = GROUPS == 2 ? kSlots - 1 : kSlots;
    static constexpr int U_SIZE = GROUPS * kAttnView > U_BYTES ? GROUPS * kAttnView : U_BYTES;
    alignas(1024) uint8_t ring[SLOTS][Fmt<TYPE>::SLOT_BYTES];
    alignas(128) uint8_t u[U_SIZE];
    double red[kCW];
    float cand_v[kCW];
    int cand_i[kCW];
    alignas(8) unsigned long long full_bar[kSlots];
    alignas(8) unsigned long long empty_bar[kSlots];
    alignas(8) unsigned long long kv_bar[2]; // ATTN: a group's K / V chunk rows have landed (its warps' arrivals + bytes)
    long long tr[2]; // trace stamps of the current phase (thread 0)
    alignas(8) ll_t resid[32]; // residual words of this CTA's first 32 rows, fetched asynchronously before the passes
    int sched[6]; // [3] = token, [4] = number of positions; the token's attention schedule: chunks, this rank's ATTN items, items % grid (kept out of registers:
                  // the kernel runs at the 96-register limit of 17 warps per SM and a spill costs an L2 round trip)
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}\n" ::"r"(bar), "r"(parity) : "memory");
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
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
// orders this thread's (and, through a preceding barrier, the CTA's) generic-proxy accesses before later async-proxy ones
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kCT) : "memory"); }
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
__device__ __forceinline__ float exp_f16(const DecoderParams &P, float x) {
    const uint16_t hb = float_to_half_bits(x);
    if (!P.exp_compute || (hb & 0x7FFFu) > 0x7C00u) return half_bits_to_float(__ldg(P.exp_table + hb)); // no proof / NaN
    return exp_f16_computed(half_bits_to_float(hb));
}
__global__ void exp_check_kernel(const uint16_t *table, int *mismatches) {
    const uint32_t hb = blockIdx.x * blockDim.x + threadIdx.x;
    if (hb >= 65536u || (hb & 0x7FFFu) > 0x7C00u) return;
    const float got = exp_f16_computed(half_bits_to_float((uint16_t)hb));
    if (float_to_half_bits(got) != table[hb]) atomicAdd(mismatches, 1);
}

// ---- flagged words.  An aligned 8-byte access is single-copy atomic, so payload and tag are seen together;
// relaxed gpu-scope accesses go to L2 (never a stale L1 line).  ll_get = the word if its tag matches, else poll.
// A flagged word is written to EVERY rank's copy of the buffer (tensor parallel: one-shot push all-gather over
// NVLink peer stores, the consumer polls its local memory); with one GPU this is a single local store.
__device__ __forceinline__ void ll_put(const DecoderParams &pe, ll_t *p, uint32_t payload, uint32_t tag) {
    const ll_t v = ((ll_t)tag << 32) | payload;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#pragma unroll 1
    for (int r = 1; r < pe.R; r++) { // peer_delta[0] is this rank (0); entries 1 .. R-1 are the peer mappings
        const char *q = reinterpret_cast<const char *>(p) + pe.peer_delta[r];
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(q), "l"(v) : "memory");
    }
}
__device__ __forceinline__ void ll_put_f(const DecoderParams &pe, ll_t *p, float x, uint32_t tag) { ll_put(pe, p, __float_as_uint(x), tag); }
__device__ __forceinline__ ll_t ll_peek(const ll_t *p) {
    ll_t v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ll_get(ll_t v, const ll_t *p, uint32_t tag) {
    while ((uint32_t)(v >> 32) != tag) v = ll_peek(p);
    return (uint32_t)v;
}
__device__ __forceinline__ float ll_get_f(ll_t v, const ll_t *p, uint32_t tag) { return __uint_as_float(ll_get(v, p, tag)); }

__device__ __forceinline__ unsigned long long gtime() { // ns, one clock for the whole GPU (SM clocks are not aligned)
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// fine-grained SM-clock stamps of ONE phase (debug builds of the trace): thread 0 of every CTA
// (dbg is a flag, not a pointer, and the trace stamps live in shared memory: the kernel has no registers to spare)
#define DBG_STAMP(i) do { if (dbg && threadIdx.x == 0) P.dbg[16 * blockIdx.x + (i)] = clock64(); } while (0)

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
template <int TYPE, int GROUPS>
__device__ __forceinline__ void producer_loop(const DecoderParams &P, Smem<TYPE, GROUPS> &S, int cta, int G) {
    using F = Fmt<TYPE>;
    constexpr int SLOTS = Smem<TYPE, GROUPS>::SLOTS;
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
        mbar_expect_tx(full0 + 8 * slot, bytes);
        bulk_g2s(ring + slot * F::SLOT_BYTES, src, bytes, full0 + 8 * slot, pol);
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
__device__ __forceinline__ void quant_quad_to_smem(uint32_t u, int quad, int lane, const float (&v)[4], int hi_off) {
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
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(u + U_ACT_D + 2 * b), "h"(hd) : "memory");
        asm volatile("st.shared.s32 [%0], %1;" ::"r"(u + U_BIAS + 4 * b), "r"(-8 * qs) : "memory");
    }
}

// rms_norm (+ weight mul) + quantize_row_q8_0 of the phase input; every CTA recomputes it (16 KB of flagged
// words from L2) instead of making it a phase of its own.  Sum of squares in the DEV order: 256 strided double
// partials (partial t = elements t, t + 256, ...), butterfly per warp, the 8 warp sums in order.  Thread
// (warp w, lane l) loads exactly the elements of partial 32 * (w & 7) + l, so both warp halves hold the
// partials in registers and no staging pass is needed; the halves then split the quantisation by blocks.
template <int TYPE, int GROUPS>
__device__ __forceinline__ void norm_prologue(const DecoderParams &P, Smem<TYPE, GROUPS> &S, const Phase &ph, int token,
                                              uint32_t tag, float *f32_out, bool dbg) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int E = P.E, nk = E / GGB_DEV_NORM_THREADS;
    DBG_STAMP(1);
    const int base = 32 * (warp & 7) + lane;
    float *xs = reinterpret_cast<float *>(S.u + U_XS);
    const ll_t *in = ph.in_ll;
    // the norm weights of this warp's first quad of blocks: requested before the (long) wait for x
    const float *norm_w = ph.norm_w;
    const int nquad = E / 128;
    float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (warp < nquad) w4 = *reinterpret_cast<const float4 *>(norm_w + 128 * warp + 4 * lane);
    double acc = 0.0;
    // a thread's words are requested 8 at a time (all of them for E <= 2048): ONE L2 round trip when x is there
    constexpr int NV = GROUPS == 2 ? 4 : 8; // (the two-group kernel has fewer registers to spare)
#pragma unroll 1
    for (int k0 = 0; k0 < nk; k0 += NV) {
        ll_t v[NV];
#pragma unroll
        for (int u = 0; u < NV; u++)
            if (in && k0 + u < nk) v[u] = ll_peek(in + GGB_DEV_NORM_THREADS * (k0 + u) + base);
#pragma unroll
        for (int u = 0; u < NV; u++) {
            if (k0 + u < nk) {
                const int i = GGB_DEV_NORM_THREADS * (k0 + u) + base;
                const float x = in ? ll_get_f(v[u], in + i, tag) : embed_value<TYPE>(P, token, i);
                xs[i] = x;
                acc = __dadd_rn(acc, (double)__fmul_rn(x, x));
            }
        }
    }
    DBG_STAMP(2);
    acc = warp_sum_f64(acc);
    if (warp < 8 && lane == 0) S.red[warp] = acc;
    DBG_STAMP(3);
    consumer_bar();
    DBG_STAMP(4);
    double sum = S.red[0];
#pragma unroll
    for (int i = 1; i < GGB_DEV_NORM_THREADS / 32; i++) sum = __dadd_rn(sum, S.red[i]);
    // sum / E: a power-of-two E scales exactly (same bits as the division, which is a ~600-cycle routine)
    const float mean = __double2float_rn((E & (E - 1)) == 0 ? __dmul_rn(sum, 1.0 / (double)E) : __ddiv_rn(sum, (double)E));
    const float scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, P.eps)));
    DBG_STAMP(5);
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
        quant_quad_to_smem(ua, quad, lane, v, E / 2);
    }
    DBG_STAMP(6);
    consumer_bar();
    DBG_STAMP(7);
}

// O: the flagged fp32 row written by COMBINE -> quantize_row_q8_0 (ggml mul_mat INIT stage)
template <int TYPE, int GROUPS>
__device__ __forceinline__ void quant_prologue(Smem<TYPE, GROUPS> &S, const ll_t *in, int K, uint32_t tag) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ua = smem_u32(S.u);
#pragma unroll 1
    for (int quad = warp; quad < K / 128; quad += kCW) {
        const ll_t *p = in + 128 * quad + 4 * lane;
        ll_t w[4];
#pragma unroll
        for (int e = 0; e < 4; e++) w[e] = ll_peek(p + e);
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; e++) v[e] = ll_get_f(w[e], p + e, tag);
        quant_quad_to_smem(ua, quad, lane, v, K / 2);
    }
    consumer_bar();
}

// DOWN: packed Q8_0 row written by GATE_UP: 4 codes per flagged word, then one word per block with the fp16 scale
// (low half) and the sum of the block's codes (high half, int16: |sum| <= 32 * 127), from which the Q4_0 bias follows
template <int TYPE, int GROUPS>
__device__ __forceinline__ void packed_prologue(Smem<TYPE, GROUPS> &S, const ll_t *in, int K, uint32_t tag) {
    const int tid = threadIdx.x;
    const int nw = K / 4, ns = K / 32;
    uint32_t *aq = reinterpret_cast<uint32_t *>(S.u);
    __half *ad = reinterpret_cast<__half *>(S.u + U_ACT_D);
    int *bias = reinterpret_cast<int *>(S.u + U_BIAS);
#pragma unroll 1
    for (int i0 = tid; i0 < nw + ns; i0 += 4 * kCT) { // 4 words in flight per thread (more spills at the 96-register limit)
        ll_t v[4];
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (i0 + kCT * u < nw + ns) v[u] = ll_peek(in + i0 + kCT * u);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = i0 + kCT * u;
            if (i < nw + ns) {
                const uint32_t w = ll_get(v[u], in + i, tag);
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

// ---------------------------------------------------------------- one pass: 2 rows x one K segment
// Lane l owns blocks l and l + 32 of the segment for both rows.  Per-block term = the scalar reference's
//   Q4_0: ((float)sumi * d_x) * d_y      Q8_0: (float)sumi * (d_x * d_y)
// lane partial = blocks in order from 0.0f; the two rows' butterflies (xor 16,8,4,2,1) are folded into one:
// lanes < 16 end with row A's sum, lanes >= 16 with row B's (same additions, DEV order).
struct ActRegs { // the lane's two activation blocks of one K segment
    int a[2][8];
    uint32_t dy2; // the two fp16 activation scales in one register (the kernel runs at its register limit)
    int bias[2];
    bool live[2];
};
template <int TYPE>
__device__ __forceinline__ void load_act_regs(ActRegs &A, uint32_t act, int seg, int nb, int lane) {
#pragma unroll
    for (int t = 0; t < 2; t++) {
        const int b = seg * kSeg + lane + 32 * t;
        A.live[t] = b < nb;
        const int bc = A.live[t] ? b : 0;
        const uint4 lo = lds_u4(act + bc * 16), hi = lds_u4(act + nb * 16 + bc * 16); // planes: bytes 0..15 | 16..31
        A.a[t][0] = lo.x; A.a[t][1] = lo.y; A.a[t][2] = lo.z; A.a[t][3] = lo.w;
        A.a[t][4] = hi.x; A.a[t][5] = hi.y; A.a[t][6] = hi.z; A.a[t][7] = hi.w;
        unsigned short hd;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(hd) : "r"(act + U_ACT_D + bc * 2));
        A.dy2 = t == 0 ? (uint32_t)hd : (A.dy2 | ((uint32_t)hd << 16));
        A.bias[t] = TYPE == GGB_TYPE_Q4_0 ? lds_s32(act + U_BIAS + bc * 4) : 0;
    }
}
// QUAD = true: the pass holds the short last segment (<= 32 blocks) of FOUR rows; lane l owns block l of each row
// (its "l + 32" is past the end of the row), lane partial = 0.0f + term, and the four rows' butterflies are folded
// into one (xor 16 keeps two rows per half-warp, xor 8 one row per quarter): the same additions as the plain
// butterfly of each row.  Returns row 2 * (lane bit 4) + (lane bit 3)'s segment sum.  The byte layout of a pass is
// the same in both cases: four groups of 32 blocks, payloads then scales.
template <int TYPE, bool QUAD>
__device__ __forceinline__ float pass_dot(uint32_t base, const ActRegs &A, int lane) {
    using F = Fmt<TYPE>;
    const int(&a)[2][8] = A.a;
    const float dy[2] = {__half2float(__ushort_as_half((unsigned short)(A.dy2 & 0xFFFFu))),
                         __half2float(__ushort_as_half((unsigned short)(A.dy2 >> 16)))};
    const int(&bias)[2] = A.bias;
    const bool(&live)[2] = A.live;
    float term[4]; // plain: [row][block]; quad: [row]
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int t = QUAD ? 0 : (k & 1);
        // Q8_0 groups hold the first 16 codes of their 32 blocks, then the second 16 (two planes of 512 B): both
        // LDS.128 of a lane are then 16-byte-strided across the warp, i.e. conflict-free
        const uint32_t pay = base + k * (32 * F::PB) + lane * 16;
        const float dx = lds_half(base + 2 * F::ROW_PB + k * 64 + lane * 2);
        if (TYPE == GGB_TYPE_Q4_0) {
            const uint4 w = lds_u4(pay);
            const uint32_t q[4] = {w.x, w.y, w.z, w.w};
            int lo = bias[t], hi = 0; // high nibbles stay in place (x16) and are divided out once: exact
#pragma unroll
            for (int i = 0; i < 4; i++) {
                lo = __dp4a((int)(q[i] & 0x0F0F0F0Fu), a[t][i], lo);
                hi = dp4a_us(q[i] & 0xF0F0F0F0u, a[t][4 + i], hi);
            }
            const int sumi = lo + (hi >> 4);
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
    const bool b4 = lane & 16;
    if (QUAD) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = live[0] ? __fadd_rn(0.0f, term[k]) : 0.0f;
        const bool b3 = lane & 8;
        const float k20 = __fadd_rn(b4 ? v[2] : v[0], __shfl_xor_sync(0xffffffffu, b4 ? v[0] : v[2], 16));
        const float k21 = __fadd_rn(b4 ? v[3] : v[1], __shfl_xor_sync(0xffffffffu, b4 ? v[1] : v[3], 16));
        float keep = __fadd_rn(b3 ? k21 : k20, __shfl_xor_sync(0xffffffffu, b3 ? k20 : k21, 8));
#pragma unroll
        for (int off = 4; off >= 1; off >>= 1) keep = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, keep, off));
        return keep;
    }
    float a0 = 0.0f, a1 = 0.0f;
    a0 = live[0] ? __fadd_rn(a0, term[0]) : a0;
    a0 = live[1] ? __fadd_rn(a0, term[1]) : a0;
    a1 = live[0] ? __fadd_rn(a1, term[2]) : a1;
    a1 = live[1] ? __fadd_rn(a1, term[3]) : a1;
    const float send = b4 ? a0 : a1;
    float keep = __fadd_rn(b4 ? a1 : a0, __shfl_xor_sync(0xffffffffu, send, 16));
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) keep = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, keep, off));
    return keep;
}

// ---------------------------------------------------------------- ATTN
// items (kv head g, chunk of 32 positions, half of the query heads).  Same arithmetic as attention.cu
// attn_partial_kernel; rope + Q scale are applied to q on load, and the item that owns the newest position ropes K,
// rounds K/V to fp16, stores them in the cache and uses them.  The K/V rows of older positions do not depend on
// this token: they travel as bulk copies (TMA, one per row, issued by lane 0 of the warp that will score the row;
// completion on S.kv_bar) while the CTA still waits for q.  Everything here is issue-bound, so the instruction
// count is what was minimised: q is staged as fp32 (converted once), a warp scores its two positions against all
// heads of the item at once and reduces the 8 lane partials with ONE folded butterfly (9 shuffles instead of 40;
// every sum adds the same pairs in the same order as the plain xor-butterfly), and P.V runs with one thread per
// pair of dims for all heads (V converted once, the heads' weights in one LDS.128).
// (the group index is a run-time value: making it a template parameter - two inlined copies - or calling the groups'
// code as real functions both measured slower on Gemma-7B; the two-group kernel has ~50 bytes of spills either way)
template <int TYPE, int GROUPS>
__device__ __forceinline__ void attn_group(const DecoderParams &P, Smem<TYPE, GROUPS> &S, const Phase &ph, int cta, int G,
                                           int n_pos, int n_items, uint32_t tag, uint32_t &kv_parity, bool dbg) {
    // a group = kCW / GROUPS warps working on one item: thread / warp index inside the group, the group's view
    constexpr int NW = kCW / GROUPS, NT = kCT / GROUPS;
    const int lane = threadIdx.x & 31, grp = GROUPS == 2 ? (threadIdx.x >> 5) / NW : 0;
    const int warp = (threadIdx.x >> 5) - grp * NW, tid = threadIdx.x - grp * NT;
    uint8_t *const view = S.u + grp * kAttnView;
    auto group_bar = [&]() {
        if (GROUPS == 1) consumer_bar();
        else asm volatile("bar.sync %0, %1;" ::"r"(2 + grp), "n"(NT) : "memory");
    };
    const int d = P.d, halfd = d >> 1, Hkv = P.Hkv, Gq = P.Gq;
    // with more than 4 query heads per KV head an item is split in two (heads 0..Gq/2-1 | the rest): the K/V
    // chunk is read twice (L2), the score / PV work per CTA halves and twice as many SMs take part
    const int hsplit = P.hsplit, Gh = P.Gh;
    // (chunk, kv head) pairs gc = chunk * Hkv + g are dealt round-robin to the ranks (pair gc belongs to rank
    // gc % R for every later token as well: its K/V rows, including the new position's, live only there); a rank's
    // n_items items go round-robin to its CTAs
    const int pos = n_pos - 1;
    const size_t kv_stride = (size_t)Hkv * d;
    const float *cosr = P.rope_cos + (size_t)pos * halfd, *sinr = P.rope_sin + (size_t)pos * halfd;
    const ll_t *q_all = ph.in_ll, *k_all = q_all + (size_t)P.H * d, *v_all = k_all + (size_t)Hkv * d;
    __half *Kc = reinterpret_cast<__half *>(view + U_KC), *Vc = reinterpret_cast<__half *>(view + U_VC);
    float *Sq = reinterpret_cast<float *>(view + U_SQ);              // [4][d] fp16-rounded q, widened
    float(*Ss)[kChunk] = reinterpret_cast<float(*)[kChunk]>(view + U_S); // [4][32] scores
    float *Se = reinterpret_cast<float *>(view + U_E);               // [32][4] soft-max numerators, heads innermost
    const uint32_t kc_addr = smem_u32(Kc), vc_addr = smem_u32(Vc), kvbar = smem_u32(&S.kv_bar[grp]);
    const uint32_t row_bytes = (uint32_t)d * 2;
#pragma unroll 1
    for (int it = (cta + ph.rot >= G ? cta + ph.rot - G : cta + ph.rot) + grp * G; it < n_items; it += GROUPS * G) {
        const int hh = hsplit == 2 ? (it & 1) : 0;
        const int gc = (hsplit == 2 ? (it >> 1) : it) * P.R + P.rank;
        const int g = gc & (Hkv - 1), c = gc >> P.kvh_shift;
        const int p0 = c * kChunk, np = min(kChunk, n_pos - p0);
        const int h0 = g * Gq + hh * Gh;
        const int jnew = pos - p0; // row of the new position inside this chunk (if 0 <= jnew < np)
        const bool owns_new = jnew >= 0 && jnew < np;
        // K / V chunk rows (the new position's row is computed below, not loaded).  MQA (one KV head): the rows of a
        // chunk are contiguous in the cache, so warps 0..3 issue one copy each (K | V) x (rows before | after the new
        // one); otherwise the warp that scores rows w, w + 16 copies them.  Every warp arrives once on kv_bar.
        if (lane == 0) {
#ifndef GGB_EXP_NOFENCE
            fence_proxy_async(); // the union region was last touched through the generic proxy (ordered by the phase barrier)
#endif
            if (Hkv == 1) {
                const int lo = (warp & 1) ? (owns_new ? jnew + 1 : np) : 0;
                const int hi = (warp & 1) ? np : (owns_new ? jnew : np);
                const uint32_t bytes = warp < 4 && hi > lo ? (uint32_t)(hi - lo) * row_bytes : 0u;
                mbar_expect_tx(kvbar, bytes);
                if (bytes) {
                    const __half *cache = warp < 2 ? ph.kcache : ph.vcache;
                    bulk_g2s_plain((warp < 2 ? kc_addr : vc_addr) + (uint32_t)lo * row_bytes, cache + (size_t)(p0 + lo) * d, bytes, kvbar);
                }
            } else {
                uint32_t bytes = 0;
                for (int j = warp; j < np; j += NW) bytes += j == jnew ? 0u : 2u * row_bytes;
                mbar_expect_tx(kvbar, bytes); // arrive (+ expected bytes, possibly 0): the barrier counts the kCW warps
                for (int j = warp; j < np; j += NW) {
                    if (j == jnew) continue;
                    const size_t src = (size_t)(p0 + j) * kv_stride + (size_t)g * d;
                    bulk_g2s_plain(kc_addr + (uint32_t)j * row_bytes, ph.kcache + src, row_bytes, kvbar);
                    bulk_g2s_plain(vc_addr + (uint32_t)j * row_bytes, ph.vcache + src, row_bytes, kvbar);
                }
            }
        }
        DBG_STAMP(1);
        // q: rope (NEOX pairs i, i + d/2), scale, round to fp16 (ggml mul_mat INIT converts src1 to fp16)
        for (int i = tid; i < Gh * halfd; i += NT) {
            const int jh = i >> P.hd_shift, e = i & (halfd - 1);
            const ll_t *qh = q_all + (size_t)(h0 + jh) * d;
            const ll_t w0 = ll_peek(qh + e), w1 = ll_peek(qh + e + halfd);
            const float cs = cosr[e], sn = sinr[e];
            const float x0 = ll_get_f(w0, qh + e, tag), x1 = ll_get_f(w1, qh + e + halfd, tag);
            const float r0 = __fmul_rn(__fsub_rn(__fmul_rn(x0, cs), __fmul_rn(x1, sn)), P.q_scale);
            const float r1 = __fmul_rn(__fadd_rn(__fmul_rn(x0, sn), __fmul_rn(x1, cs)), P.q_scale);
            // two planes per head so that the score loop's LDS.128 are conflict-free (a lane's dims 8l..8l+3 at 16 l of
            // plane 0, dims 8l+4..8l+7 at 16 l of plane 1; one 32-byte-strided row would take two wavefronts per load)
            const int e1 = e + halfd;
            Sq[jh * d + ((e >> 2) & 1) * halfd + (e >> 3) * 4 + (e & 3)] = __half2float(__float2half_rn(r0));
            Sq[jh * d + ((e1 >> 2) & 1) * halfd + (e1 >> 3) * 4 + (e1 & 3)] = __half2float(__float2half_rn(r1));
        }
        DBG_STAMP(2);
        if (owns_new) { // KV store of the new position (ref: src/gemma_model.cpp:499-518)
            for (int e = tid; e < halfd; e += NT) {
                const ll_t *kp = k_all + (size_t)g * d + e;
                const ll_t w0 = ll_peek(kp), w1 = ll_peek(kp + halfd);
                const float cs = cosr[e], sn = sinr[e];
                const float x0 = ll_get_f(w0, kp, tag), x1 = ll_get_f(w1, kp + halfd, tag);
                const __half k0 = __float2half_rn(__fsub_rn(__fmul_rn(x0, cs), __fmul_rn(x1, sn)));
                const __half k1 = __float2half_rn(__fadd_rn(__fmul_rn(x0, sn), __fmul_rn(x1, cs)));
                Kc[jnew * d + e] = k0;
                Kc[jnew * d + e + halfd] = k1;
                if (hh == 0) {
                    __half *krow = ph.kcache + ((size_t)pos * Hkv + g) * d;
                    krow[e] = k0;
                    krow[e + halfd] = k1;
                }
            }
            for (int e = tid; e < d; e += NT) {
                const ll_t *vp = v_all + (size_t)g * d + e;
                const __half vv = __float2half_rn(ll_get_f(ll_peek(vp), vp, tag));
                Vc[jnew * d + e] = vv;
                if (hh == 0) ph.vcache[((size_t)pos * Hkv + g) * d + e] = vv;
            }
        }
        DBG_STAMP(3);
        mbar_wait(kvbar, kv_parity);
        kv_parity ^= 1u;
        group_bar();
        DBG_STAMP(4);
        // scores: warp w takes positions w and w + NW (and, with two groups, w + 2 NW and w + 3 NW) against the item's (<= 4) heads; lane = 8 dims.  fp16 x fp16
        // products are exact in fp32, so fmaf(k, q, acc) == acc + k*q rounded once
#pragma unroll 1
        for (int rp = 0; rp < GROUPS; rp++) { // a warp takes kChunk / NW positions, two at a time
            const bool in = 8 * lane < d;
            const int j0 = warp + 2 * rp * NW, j1 = j0 + NW;
            float kf[2][8];
#pragma unroll
            for (int p = 0; p < 2; p++) {
                const int j = p ? j1 : j0;
                uint4 raw = make_uint4(0, 0, 0, 0);
                if (in && j < np) raw = *reinterpret_cast<const uint4 *>(Kc + j * d + 8 * lane);
                const __half2 *h2 = reinterpret_cast<const __half2 *>(&raw);
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const float2 f = __half22float2(h2[e]);
                    kf[p][2 * e] = f.x;
                    kf[p][2 * e + 1] = f.y;
                }
            }
            float acc[2][4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                float4 qa = make_float4(0.f, 0.f, 0.f, 0.f), qb = qa;
                if (u < Gh && in) {
                    qa = *reinterpret_cast<const float4 *>(Sq + u * d + 4 * lane);
                    qb = *reinterpret_cast<const float4 *>(Sq + u * d + halfd + 4 * lane);
                }
                const float qf[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    float a = 0.0f;
#pragma unroll
                    for (int e = 0; e < 8; e++) a = __fmaf_rn(kf[p][e], qf[e], a);
                    acc[p][u] = a;
                }
            }
            // folded butterfly: xor 16 keeps one position per half-warp, xor 8 two heads, xor 4 one head
            const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
            float k4[4], k2[2];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const float send = b4 ? acc[0][u] : acc[1][u], keep = b4 ? acc[1][u] : acc[0][u];
                k4[u] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 16));
            }
#pragma unroll
            for (int v = 0; v < 2; v++) {
                const float send = b3 ? k4[v] : k4[2 + v], keep = b3 ? k4[2 + v] : k4[v];
                k2[v] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 8));
            }
            float k1;
            {
                const float send = b2 ? k2[0] : k2[1], keep = b2 ? k2[1] : k2[0];
                k1 = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 4));
            }
            k1 = __fadd_rn(k1, __shfl_xor_sync(0xffffffffu, k1, 2));
            k1 = __fadd_rn(k1, __shfl_xor_sync(0xffffffffu, k1, 1));
            const int js = b4 ? j1 : j0, us = (b3 ? 2 : 0) + (b2 ? 1 : 0);
            if ((lane & 3) == 0 && js < np && us < Gh) Ss[us][js] = k1;
        }
        DBG_STAMP(5);
        group_bar();
        // per-head chunk soft-max through the fp16 exp table (heads the item does not have get weight 0)
        if (warp < 4) {
            float e = 0.0f;
            if (warp < Gh) {
                const float sv = lane < np ? Ss[warp][lane] : -INFINITY;
                const float m = warp_max_f32(sv);
                e = lane < np ? exp_f16(P, __fsub_rn(sv, m)) : 0.0f;
                const float l = warp_sum_f32(e);
                if (lane == 0) {
                    ll_put_f(P, ph.m_ll + (size_t)(h0 + warp) * P.max_chunks + c, m, tag);
                    ll_put_f(P, ph.l_ll + (size_t)(h0 + warp) * P.max_chunks + c, l, tag);
                }
            }
            Se[lane * 4 + warp] = e;
        }
        DBG_STAMP(6);
        group_bar();
        DBG_STAMP(7);
        // acc[h][i] = sum_p e[h][p] * V[p][i], sequential in p; thread -> dims 2 tid, 2 tid + 1 for every head
        // (e is an fp16 value: e*v is exact, fmaf == mul + add)
        if (tid < halfd) {
            float a0[4], a1[4];
#pragma unroll
            for (int u = 0; u < 4; u++) a0[u] = a1[u] = 0.0f;
#pragma unroll 4
            for (int j = 0; j < np; j++) {
                const float2 vf = __half22float2(*reinterpret_cast<const __half2 *>(Vc + j * d + 2 * tid));
                const float4 e4 = *reinterpret_cast<const float4 *>(Se + 4 * j);
                a0[0] = __fmaf_rn(e4.x, vf.x, a0[0]); a1[0] = __fmaf_rn(e4.x, vf.y, a1[0]);
                a0[1] = __fmaf_rn(e4.y, vf.x, a0[1]); a1[1] = __fmaf_rn(e4.y, vf.y, a1[1]);
                a0[2] = __fmaf_rn(e4.z, vf.x, a0[2]); a1[2] = __fmaf_rn(e4.z, vf.y, a1[2]);
                a0[3] = __fmaf_rn(e4.w, vf.x, a0[3]); a1[3] = __fmaf_rn(e4.w, vf.y, a1[3]);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (u < Gh) {
                    ll_t *out = ph.acc_ll + ((size_t)(h0 + u) * P.max_chunks + c) * d + 2 * tid;
                    ll_put_f(P, out, a0[u], tag);
                    ll_put_f(P, out + 1, a1[u], tag);
                }
            }
        }
        DBG_STAMP(8);
        group_bar();
    }
}

template <int TYPE, int GROUPS>
__device__ __forceinline__ void attn_phase(const DecoderParams &P, Smem<TYPE, GROUPS> &S, const Phase &ph, int cta, int G,
                                           int n_pos, int n_items, uint32_t tag, uint32_t &kv_parity, bool dbg) {
    attn_group<TYPE, GROUPS>(P, S, ph, cta, G, n_pos, n_items, tag, kv_parity, dbg);
}

// COMBINE: one CTA per (head, 32-dim slice): M = max m_c; w_c = exp_table[f16(m_c - M)]; L = sum l_c*w_c;
// O = sum acc_c*w_c (sequential in c); out = O / L.  All 16 warps stage the flagged partials (one L2 round
// trip), warp 0 runs the sequential chain from shared memory (lane = dim).  Tasks start on the CTAs after the
// ones that own ATTN items, which are free earlier.
template <int TYPE, int GROUPS>
__device__ __forceinline__ void combine_phase(const DecoderParams &P, Smem<TYPE, GROUPS> &S, const Phase &ph, int cta, int G,
                                              int nchunk, int items_rem, uint32_t tag, bool dbg) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = P.d, slices = d >> 5;
    const int n_tasks = P.H * slices;
    float *stage = reinterpret_cast<float *>(S.u);
    float *sm = reinterpret_cast<float *>(S.u + U_CM), *sl = reinterpret_cast<float *>(S.u + U_CL);
    float *sw = reinterpret_cast<float *>(S.u + U_CW);
    int tl = cta + ph.rot; // tasks start on the CTAs after the ones that own this rank's ATTN items (items_rem = items % G)
    if (tl >= G) tl -= G;
    tl -= items_rem;
    if (tl < 0) tl += G;
#pragma unroll 1
    for (; tl * P.R + P.rank < n_tasks; tl += G) {
        const int task = tl * P.R + P.rank; // tasks are dealt round-robin to the ranks
        const int h = task >> P.sl_shift, sli = task & (slices - 1);
        const ll_t *pa = ph.acc_ll + (size_t)h * P.max_chunks * d + sli * 32 + lane;
        const ll_t *pm = ph.m_ll + (size_t)h * P.max_chunks, *pl = ph.l_ll + (size_t)h * P.max_chunks;
#pragma unroll 1
        for (int c0 = warp; c0 < nchunk; c0 += 4 * kCW) {
            ll_t v[4];
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (c0 + kCW * u < nchunk) v[u] = ll_peek(pa + (size_t)(c0 + kCW * u) * d);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int c = c0 + kCW * u;
                if (c < nchunk) stage[c * 32 + lane] = ll_get_f(v[u], pa + (size_t)c * d, tag);
            }
        }
        for (int c = tid; c < nchunk; c += kCT) {
            const ll_t vm = ll_peek(pm + c), vl = ll_peek(pl + c);
            sm[c] = ll_get_f(vm, pm + c, tag);
            sl[c] = ll_get_f(vl, pl + c, tag);
        }
        DBG_STAMP(1);
        consumer_bar();
        DBG_STAMP(2);
        if (warp == 0) {
            float mx = -INFINITY;
            for (int c = lane; c < nchunk; c += 32) mx = fmaxf(mx, sm[c]);
            mx = warp_max_f32(mx);
            for (int c = lane; c < nchunk; c += 32)
                sw[c] = exp_f16(P, __fsub_rn(sm[c], mx));
            __syncwarp();
            DBG_STAMP(3);
            float L = 0.0f, O = 0.0f;
#pragma unroll 1
            for (int c0 = 0; c0 < nchunk; c0 += 16) { // products of 16 chunks first (independent), then the chain
                float pl[16], po[16];
#pragma unroll
                for (int u = 0; u < 16; u++) {
                    const int c = min(c0 + u, nchunk - 1);
                    const float w = sw[c];
                    pl[u] = __fmul_rn(sl[c], w);
                    po[u] = __fmul_rn(stage[c * 32 + lane], w);
                }
#pragma unroll
                for (int u = 0; u < 16; u++) {
                    if (c0 + u < nchunk) {
                        L = __fadd_rn(L, pl[u]);
                        O = __fadd_rn(O, po[u]);
                    }
                }
            }
            const float r = __fdiv_rn(O, L);
            const size_t idx = (size_t)h * d + sli * 32 + lane;
            ph.f32_out[idx] = r;
            ll_put_f(P, ph.out_ll + idx, r, tag);
            DBG_STAMP(4);
        }
        consumer_bar();
    }
}

// ---------------------------------------------------------------- the kernel
// QUAD: the model has phases in the quad layout (a short last K segment).  A compile-time switch, because the pass loop
// is issue-bound: the two extra branches of the quad layout cost a model without such phases (Gemma-2B) 1.5 %.
template <int TYPE, bool QUAD, int GROUPS>
__global__ void __launch_bounds__(kThreads, 1) decoder_kernel(const __grid_constant__ DecoderParams P) {
    using F = Fmt<TYPE>;
    using SM = Smem<TYPE, GROUPS>;
    SM &S = *reinterpret_cast<SM *>(ggb_dec_smem);
    const int cta = blockIdx.x, G = gridDim.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < SM::SLOTS; i++) {
            mbar_init(smem_u32(&S.full_bar[i]), 1);
            mbar_init(smem_u32(&S.empty_bar[i]), kCW);
        }
        mbar_init(smem_u32(&S.kv_bar[0]), kCW / GROUPS);
        mbar_init(smem_u32(&S.kv_bar[1]), kCW / GROUPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int token = min(max(P.state[0], 0), P.V - 1);
    const int n_pos = min(max(P.state[1], 0), P.n_ctx - 1) + 1;
    const uint32_t tag = *P.seq + 1u;
    if (tid == 0) { // the token's attention schedule (the same for every layer)
        const int nchunk = (n_pos + kChunk - 1) / kChunk;
        const int items = (P.R == 1 ? P.Hkv * nchunk : (P.Hkv * nchunk - P.rank + P.R - 1) / P.R) * P.hsplit;
        S.sched[0] = nchunk;
        S.sched[1] = items;
        S.sched[2] = items % G;
        S.sched[3] = token;
        S.sched[4] = n_pos;
    }

    if (warp == kCW) {
        if (lane == 0) producer_loop<TYPE, GROUPS>(P, S, cta, G);
        return;
    }

    // ---- consumers
    const uint32_t ring = smem_u32(&S.ring[0][0]);
    const uint32_t full0 = smem_u32(&S.full_bar[0]), empty0 = smem_u32(&S.empty_bar[0]);
    const uint32_t act = smem_u32(S.u);
    float *part = reinterpret_cast<float *>(S.u + U_PART);
    int slot = 0;
    uint32_t parity = 0, kv_parity = 0;
    if (tid < kPhWords) reinterpret_cast<uint32_t *>(&S.ph[0])[tid] = reinterpret_cast<const uint32_t *>(&P.phases[0])[tid];
    consumer_bar();
#pragma unroll 1
    for (int pi = 0; pi < P.n_phases; pi++) {
        const Phase &ph = S.ph[pi & 1];
        // prefetch the next descriptor with an ASYNC copy (a plain load + store would stall warps 0-1 in order for one
        // L2 round trip, ~0.8 us under the weight stream, at the start of every phase); waited for at the phase's end
        if (pi + 1 < P.n_phases && tid < kPhWords)
            cp_async4(smem_u32(reinterpret_cast<uint32_t *>(&S.ph[(pi + 1) & 1]) + tid),
                      reinterpret_cast<const uint32_t *>(&P.phases[pi + 1]) + tid);
        if (P.trace && tid == 0) S.tr[0] = S.tr[1] = (long long)gtime();
        const int kind = ph.kind;
        const bool dbg = P.dbg && pi == P.dbg_phase;
        DBG_STAMP(0);
        if (is_matvec(kind)) {
            int u0, u1;
            unit_range(cta, G, ph.rot, ph.uq, ph.ur, u0, u1);
            const int n_passes = (u1 - u0) * ph.unit_passes;
            const int ksegs = ph.ksegs, nb = ph.K >> 5;
            if (n_passes > 0 || (cta == 0 && kind == PH_LOGITS)) { // CTA 0 also writes the result_norm dump
                if (kind == PH_O) quant_prologue<TYPE, GROUPS>(S, ph.in_ll, ph.K, tag);
                else if (kind == PH_DOWN) packed_prologue<TYPE, GROUPS>(S, ph.in_ll, ph.K, tag);
                else norm_prologue<TYPE, GROUPS>(P, S, ph, S.sched[3], tag, (kind == PH_LOGITS && cta == 0) ? ph.f32_out : nullptr, dbg);
            }
            if (P.trace && tid == 0) S.tr[1] = (long long)gtime();
            // the residual of this thread's first row: requested before the passes, consumed in the epilogue
            // (an async 8-byte copy into shared memory: no register is held across the pass loop, no warp waits; a
            // word that had not arrived yet, or a stale L1 line, simply fails the tag check later and is polled then)
            {
                const int row0 = (ph.u_off + u0) * ph.unit_rows, nrows = (u1 - u0) * ph.unit_rows;
                if (ph.resid_ll && kind != PH_GATE_UP && tid < 32 && tid < nrows && row0 + tid < ph.n_rows)
                    cp_async8(smem_u32(&S.resid[tid]), ph.resid_ll + row0 + tid);
            }
            // ---- passes: warp w owns passes w, w + 16, ... (Q4_0: one per slot; Q8_0: 8 passes per slot, so the
            // two warp halves alternate slots).  Every warp waits for and releases every slot.  The activation
            // blocks live in registers and are reloaded only when the warp's K segment changes (never for
            // K <= 2048, and never when 16 % ksegs == 0).
            {
                const int sstep = ph.sstep, dstep = ph.dstep, wrapn = ph.wrapn, winc = ph.winc;
                const int kf = ph.kf;
                const bool quad = QUAD && ph.quad != 0;
                // this warp's first pass is number `warp` of the CTA's share (by subtraction: at most 15 steps)
                int seg = warp, pidx = 0, cur_seg = -1;
                while (seg >= wrapn) {
                    seg -= wrapn;
                    pidx += ph.winit;
                }
                if (!quad) pidx += seg;
                ActRegs A;
#pragma unroll 1
                for (int p0 = 0, i = 0; p0 < n_passes; p0 += F::PPS, i++) {
                    const int cnt = min(F::PPS, n_passes - p0);
                    const int j = F::PPS == kCW ? warp : (((warp >> 3) == (i & 1)) ? (warp & 7) : F::PPS);
                    // activation segment of this pass (quad units: rows 0-1 | rows 2-3 | the short segment of all four)
                    int aseg = seg;
                    if (quad) aseg = seg < kf ? seg : (seg >= 2 * kf ? kf : seg - kf);
                    if (j < cnt && aseg != cur_seg) {
                        load_act_regs<TYPE>(A, act, aseg, nb, lane);
                        cur_seg = aseg;
                    }
                    mbar_wait(full0 + 8 * slot, parity);
                    if (j < cnt) {
                        const uint32_t pass = ring + slot * F::SLOT_BYTES + j * F::PASS_BYTES;
                        if (quad && seg >= 2 * kf) {
                            const float keep = pass_dot<TYPE, true>(pass, A, lane);
                            if ((lane & 7) == 0) part[pidx + kf + (lane >> 3) * ksegs] = keep;
                        } else {
                            const float keep = pass_dot<TYPE, false>(pass, A, lane);
                            int out = pidx;
                            if (quad) out += seg < kf ? seg : 2 * ksegs + seg - kf;
                            if ((lane & 15) == 0) part[out + ((lane >> 4) ? ksegs : 0)] = keep;
                        }
                        pidx += dstep;
                        seg += sstep;
                        if (seg >= wrapn) { // next unit
                            seg -= wrapn;
                            pidx += winc;
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty0 + 8 * slot);
                    if (++slot == SM::SLOTS) {
                        slot = 0;
                        parity ^= 1u;
                    }
                }
            }
            DBG_STAMP(8);
            consumer_bar();
            DBG_STAMP(9);
            // ---- epilogue: finish this CTA's rows (segments added in order)
            if (kind == PH_GATE_UP) {
                // GeGLU (fp16 gelu table) + quantize_row_q8_0 in one go: 8 lanes per 32-row group (a lane owns 4 rows:
                // four table lookups in flight at once = one L2 round trip), 4 groups per warp.  A lane's four codes
                // are exactly one flagged word of the packed row; the block's fp16 scale and the sum of its codes
                // (the Q4_0 bias of DOWN's dot products) share one more.
                ll_t *oq = ph.out_ll, *od = ph.out_ll + P.F / 4;
                const int nu = u1 - u0;
#pragma unroll 1
                for (int q0 = 4 * warp; q0 < nu; q0 += 4 * kCW) {
                    const int u = q0 + (lane >> 3), i0 = 4 * (lane & 7);
                    const bool ok = u < nu;
                    float x[4];
                    {
                        float gs[4], us[4];
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const float *pg = part + ((ok ? u : 0) * 64 + i0 + e) * ksegs, *pu = pg + 32 * ksegs;
                            gs[e] = pg[0];
                            us[e] = pu[0];
                            for (int k = 1; k < ksegs; k++) {
                                gs[e] = __fadd_rn(gs[e], pg[k]);
                                us[e] = __fadd_rn(us[e], pu[k]);
                            }
                        }
                        uint16_t gl[4];
#pragma unroll
                        for (int e = 0; e < 4; e++) gl[e] = __ldg(P.gelu_table + float_to_half_bits(gs[e]));
#pragma unroll
                        for (int e = 0; e < 4; e++) x[e] = ok ? __fmul_rn(half_bits_to_float(gl[e]), us[e]) : 0.0f;
                    }
                    if (ph.f32_out && ok)
                        *reinterpret_cast<float4 *>(ph.f32_out + (size_t)(ph.u_off + u0 + u) * 32 + i0) = make_float4(x[0], x[1], x[2], x[3]);
                    float amax = fmaxf(fmaxf(fabsf(x[0]), fabsf(x[1])), fmaxf(fabsf(x[2]), fabsf(x[3])));
#pragma unroll
                    for (int off = 1; off <= 4; off <<= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off));
                    const float dd = __fdiv_rn(amax, 127.0f);
                    const float id = dd != 0.0f ? __fdiv_rn(1.0f, dd) : 0.0f;
                    const uint32_t w = ((uint32_t)(int)roundf(__fmul_rn(x[0], id)) & 0xFFu) |
                                       (((uint32_t)(int)roundf(__fmul_rn(x[1], id)) & 0xFFu) << 8) |
                                       (((uint32_t)(int)roundf(__fmul_rn(x[2], id)) & 0xFFu) << 16) |
                                       (((uint32_t)(int)roundf(__fmul_rn(x[3], id)) & 0xFFu) << 24);
                    if (ok) ll_put(P, oq + (size_t)(ph.u_off + u0 + u) * 8 + (lane & 7), w, tag);
                    int qs = __dp4a((int)w, 0x01010101, 0);
#pragma unroll
                    for (int off = 1; off <= 4; off <<= 1) qs += __shfl_xor_sync(0xffffffffu, qs, off);
                    if (ok && (lane & 7) == 0)
                        ll_put(P, od + (ph.u_off + u0 + u), ((uint32_t)qs << 16) | (uint32_t)__half_as_ushort(__float2half_rn(dd)), tag);
                }
            } else {
                Best best{-INFINITY, -1};
                const ll_t *resid = ph.resid_ll;
                const int row0 = (ph.u_off + u0) * ph.unit_rows, nrows = (u1 - u0) * ph.unit_rows; // global rows of this CTA
                if (resid) cp_async_wait_all();
#pragma unroll 1
                for (int r = tid; r < nrows; r += kCT) {
                    const int grow = row0 + r;
                    if (grow >= ph.n_rows) break;
                    const float *pp = part + r * ksegs;
                    float total = pp[0];
                    for (int k = 1; k < ksegs; k++) total = __fadd_rn(total, pp[k]);
                    if (resid) total = __fadd_rn(total, ll_get_f(r < 32 ? S.resid[r] : ll_peek(resid + grow), resid + grow, tag));
                    else if (ph.resid_embed) total = __fadd_rn(total, embed_value<TYPE>(P, S.sched[3], grow));
                    if (ph.out) ph.out[grow] = total;
                    if (ph.out_ll) ll_put_f(P, ph.out_ll + grow, total, tag);
                    if (total > best.v || best.i < 0) best = Best{total, grow};
                }
                if (kind == PH_LOGITS) { // this CTA's arg-max candidate (strict '>' / lowest index, ref: :538-543)
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) {
                        Best o;
                        o.v = __shfl_xor_sync(0xffffffffu, best.v, off);
                        o.i = __shfl_xor_sync(0xffffffffu, best.i, off);
                        best = better(best, o);
                    }
                    if (lane == 0) {
                        S.cand_v[warp] = best.v;
                        S.cand_i[warp] = best.i;
                    }
                    consumer_bar();
                    if (tid == 0) {
                        Best b{S.cand_v[0], S.cand_i[0]};
                        for (int w = 1; w < kCW; w++) b = better(b, Best{S.cand_v[w], S.cand_i[w]});
                        ll_put_f(P, P.cand_ll + P.rank * G + cta, b.v, tag);
                        ll_put(P, P.cand_ll + P.R * G + P.rank * G + cta, (uint32_t)b.i, tag);
                    }
                }
            }
        } else if (kind == PH_ATTN) {
#pragma unroll 1
            for (int rep = P.dbg_repeat; rep >= 1; rep--) {
                const bool dbg_r = dbg && rep == 1;
                if (dbg_r && threadIdx.x == 0) P.dbg[16 * cta] = clock64();
                attn_phase<TYPE, GROUPS>(P, S, ph, cta, G, S.sched[4], S.sched[1], tag, kv_parity, dbg_r);
            }
        } else if (kind == PH_COMBINE) {
#pragma unroll 1
            for (int rep = P.dbg_repeat; rep >= 1; rep--) {
                const bool dbg_r = dbg && rep == 1;
                if (dbg_r && threadIdx.x == 0) P.dbg[16 * cta] = clock64();
                combine_phase<TYPE, GROUPS>(P, S, ph, cta, G, S.sched[0], S.sched[2], tag, dbg_r);
            }
        } else { // PH_SAMPLE: CTA 0 (of every rank) picks the first maximum of all ranks' per-CTA candidates
            if (cta == 0 && warp == 0) {
                Best b{-INFINITY, -1};
                for (int i = lane; i < P.R * G; i += 32) {
                    const ll_t *pv = P.cand_ll + i, *pi2 = P.cand_ll + P.R * G + i;
                    const ll_t wv = ll_peek(pv), wi = ll_peek(pi2);
                    b = better(b, Best{ll_get_f(wv, pv, tag), (int)ll_get(wi, pi2, tag)});
                }
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    Best o;
                    o.v = __shfl_xor_sync(0xffffffffu, b.v, off);
                    o.i = __shfl_xor_sync(0xffffffffu, b.i, off);
                    b = better(b, o);
                }
                if (lane == 0) {
                    // every CTA has published its candidate, i.e. has long read state and seq
                    P.state[0] = b.i;
                    P.state[1] += P.state[2];
                    if (P.history) P.history[P.state[1]] = b.i;
                    *P.seq = tag;
                }
            }
        }
        DBG_STAMP(10);
        // the union region is rewritten by the next phase's prologue, and the prefetched descriptor is read
        cp_async_wait_all();
        consumer_bar();
        DBG_STAMP(11);
        if (P.trace && tid == 0) {
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
    int rows;
};
struct RepackArgs {
    RepackSrc m[3];
    int n_mats;
    int gate_up; // logical row r: group r / 64, j = r % 64: j < 32 -> m[0] row 32*group + j, else m[1] row 32*group + j - 32
    int n_rows;  // logical rows
    int nb, ksegs;
    int quad, kf; // quad layout (see Phase): unit = 4 rows = 2 * kf + 1 passes
    int row_off;  // first logical row of the stream (tensor parallel: a rank streams only its own rows)
    uint8_t *dst;
};
// one block per pass; a pass is four groups of 32 blocks: thread t < 128 copies block (t & 31) of group (t >> 5).
// Plain pass: groups = (row 0 | row 1) x (blocks 0-31 | 32-63) of one 64-block segment; quad pass: groups = rows 0..3,
// blocks = the short last segment.
template <int TYPE> __global__ void repack_kernel(RepackArgs A) {
    using F = Fmt<TYPE>;
    const long long p = blockIdx.x;
    const int t = threadIdx.x;
    if (t >= 128) return;
    const int k = t >> 5, i = t & 31;
    int lr, blk;
    if (!A.quad) {
        const int pair = (int)(p / A.ksegs), seg = (int)(p % A.ksegs);
        lr = A.row_off + 2 * pair + (k >> 1);
        blk = seg * kSeg + (k & 1) * 32 + i;
    } else {
        const int upq = 2 * A.kf + 1;
        const int q = (int)(p / upq), m = (int)(p % upq);
        if (m < 2 * A.kf) {
            lr = A.row_off + 4 * q + (m < A.kf ? 0 : 2) + (k >> 1);
            blk = (m < A.kf ? m : m - A.kf) * kSeg + (k & 1) * 32 + i;
        } else {
            lr = A.row_off + 4 * q + k;
            blk = A.kf * kSeg + i;
        }
    }
    int mat = 0, row = lr;
    const bool valid = lr < A.n_rows && blk < A.nb;
    if (A.gate_up) {
        const int grp = lr / 64, j = lr % 64;
        mat = j < 32 ? 0 : 1;
        row = 32 * grp + (j & 31);
    } else {
        while (mat + 1 < A.n_mats && row >= A.m[mat].rows) {
            row -= A.m[mat].rows;
            mat++;
        }
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
    int groups = 1;                // ATTN items a CTA works on at once (2 for MHA models)
};

// instantiations: (plain, 1 group) | (quad, 1 group) | (quad, 2 groups: also serves MHA models without quad phases)
static const void *decoder_fn(int type, bool quad, int groups) {
    if (type == GGB_TYPE_Q4_0) {
        if (groups == 2) return (const void *)decoder_kernel<GGB_TYPE_Q4_0, true, 2>;
        return quad ? (const void *)decoder_kernel<GGB_TYPE_Q4_0, true, 1> : (const void *)decoder_kernel<GGB_TYPE_Q4_0, false, 1>;
    }
    if (groups == 2) return (const void *)decoder_kernel<GGB_TYPE_Q8_0, true, 2>;
    return quad ? (const void *)decoder_kernel<GGB_TYPE_Q8_0, true, 1> : (const void *)decoder_kernel<GGB_TYPE_Q8_0, false, 1>;
}
static size_t decoder_smem(int type, int groups) {
    if (type == GGB_TYPE_Q4_0) return groups == 2 ? sizeof(Smem<GGB_TYPE_Q4_0, 2>) : sizeof(Smem<GGB_TYPE_Q4_0, 1>);
    return groups == 2 ? sizeof(Smem<GGB_TYPE_Q8_0, 2>) : sizeof(Smem<GGB_TYPE_Q8_0, 1>);
}

static void decoder_release(ggb_decoder *D) {
    for (void *pm : D->peer_maps) cudaIpcCloseMemHandle(pm);
    for (void *s : D->streams) cudaFree(s);
    cudaFree(D->d_phases);
    cudaFree(D->scratch);
    delete D;
}

extern "C" {

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
    const int max_chunks = (dims->n_ctx + kChunk - 1) / kChunk;
    if (!(E % 256 == 0 && E <= kMaxE && (H * d) % 256 == 0 && H * d <= kMaxE && F % 256 == 0 && F <= kMaxK &&
          d >= 32 && d <= 256 && (d & (d - 1)) == 0 && Hkv > 0 && (Hkv & (Hkv - 1)) == 0 && H % Hkv == 0 &&
          H / Hkv <= kMaxG && max_chunks <= kMaxChunks &&
          (type == GGB_TYPE_Q4_0 || type == GGB_TYPE_Q8_0))) {
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
    P.max_chunks = max_chunks;
    {   // does the computed exp reproduce the table?  (GGB_EXP_TABLE=1 forces the lookup)
        int *d_mis = nullptr, h_mis = -1;
        const char *force = getenv("GGB_EXP_TABLE");
        if (!(force && atoi(force)) && cudaMalloc(&d_mis, sizeof(int)) == cudaSuccess) {
            cudaMemsetAsync(d_mis, 0, sizeof(int), ggb_cuda_stream());
            exp_check_kernel<<<256, 256, 0, ggb_cuda_stream()>>>(P.exp_table, d_mis);
            if (cudaMemcpyAsync(&h_mis, d_mis, sizeof(int), cudaMemcpyDeviceToHost, ggb_cuda_stream()) != cudaSuccess ||
                cudaStreamSynchronize(ggb_cuda_stream()) != cudaSuccess)
                h_mis = -1;
            cudaFree(d_mis);
        }
        P.exp_compute = h_mis == 0 ? 1 : 0;
    }
    P.rank = rank; P.R = size;
    auto ilog2 = [](int v) { int s = 0; while ((1 << s) < v) s++; return s; };
    P.Gq = H / Hkv; P.hsplit = P.Gq > 4 ? 2 : 1; P.Gh = P.Gq / P.hsplit;
    P.hd_shift = ilog2(d / 2); P.sl_shift = ilog2(d / 32); P.kvh_shift = ilog2(Hkv);
    // ring copies in flight: 2 for MQA models (bounds the queueing delay the flag polls see; measured best for Gemma-2B),
    // the whole ring for MHA models, whose long ATTN phases leave the stream more to catch up (Gemma-7B: + 1.2 %)
    P.max_outstanding = Hkv > 1 ? kSlots : kMaxOutstanding; P.prefetch_chunks = kPrefetchChunks; P.dbg_repeat = 1;
    if (const char *e = getenv("GGB_MAX_OUTSTANDING")) P.max_outstanding = std::min(std::max(atoi(e), 1), kSlots);
    if (const char *e = getenv("GGB_PREFETCH_CHUNKS")) P.prefetch_chunks = std::min(std::max(atoi(e), 0), 64);
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    D->grid = sms;
    // scratch: flagged buffers, one set per layer (tag = token sequence number, so a word of the previous
    // token never matches and nothing is rewritten while another CTA may still read it)
    const size_t QD = (size_t)H * d, KD = (size_t)Hkv * d;
    size_t off = 0;
    auto carve = [&](size_t words) { size_t o = off; off = (off + words * 8 + 255) & ~(size_t)255; return o; };
    const size_t o_x = carve((size_t)L * E);                       // slot il = output of layer il
    const size_t o_qkv = carve((size_t)L * (QD + 2 * KD));
    const size_t o_m = carve((size_t)L * H * max_chunks), o_l = carve((size_t)L * H * max_chunks);
    const size_t o_acc = carve((size_t)L * H * max_chunks * d);
    const size_t o_kqv = carve((size_t)L * QD), o_sa = carve((size_t)L * E);
    const size_t o_ff = carve((size_t)L * (F / 4 + F / 32));
    const size_t o_cand = carve((size_t)2 * sms * size), o_seq = carve(1);
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
        GGB_CUDA(cudaIpcGetMemHandle(&mine, D->scratch));
        void *d_h = nullptr;
        GGB_CUDA(cudaMalloc(&d_h, sizeof(mine) * (size_t)size));
        GGB_CUDA(cudaMemcpyAsync((char *)d_h + sizeof(mine) * (size_t)rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, ggb_cuda_stream()));
        int rc = ggb_op_allgather((char *)d_h + sizeof(mine) * (size_t)rank, d_h, sizeof(mine));
        std::vector<cudaIpcMemHandle_t> all((size_t)size);
        if (rc == GGB_OK) {
            GGB_CUDA(cudaMemcpyAsync(all.data(), d_h, sizeof(mine) * (size_t)size, cudaMemcpyDeviceToHost, ggb_cuda_stream()));
            GGB_CUDA(cudaStreamSynchronize(ggb_cuda_stream()));
        }
        cudaFree(d_h);
        if (rc != GGB_OK) { decoder_release(D); return rc; }
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
    ll_t *x_ll = (ll_t *)(sb + o_x), *qkv_ll = (ll_t *)(sb + o_qkv), *m_ll = (ll_t *)(sb + o_m), *l_ll = (ll_t *)(sb + o_l);
    ll_t *acc_ll = (ll_t *)(sb + o_acc), *kqv_ll = (ll_t *)(sb + o_kqv), *sa_ll = (ll_t *)(sb + o_sa);
    ll_t *ff_ll = (ll_t *)(sb + o_ff);
    P.cand_ll = (ll_t *)(sb + o_cand);
    P.seq = (uint32_t *)(sb + o_seq);

    const int pass_bytes = type == GGB_TYPE_Q4_0 ? Fmt<GGB_TYPE_Q4_0>::PASS_BYTES : Fmt<GGB_TYPE_Q8_0>::PASS_BYTES;
    bool ok = true;
    const char *why = "";
    // builds the pass stream of one mat-vec phase and fills the schedule fields
    auto matvec_phase = [&](int kind, int layer, int n, const ggb_weight *const *w, bool gate_up) -> Phase {
        Phase ph{};
        ph.kind = kind; ph.layer = layer;
        const int nb = (int)w[0]->nb;
        ph.K = (int)w[0]->cols;
        const int nbp = nb % kSeg;                       // blocks of the short last segment (0: none)
        ph.quad = (nbp > 0 && nbp <= 32) ? 1 : 0;        // it fits one group of a pass: quad layout
        ph.kf = nb / kSeg;
        ph.ksegs = (nb + kSeg - 1) / kSeg;
        const int upq = 2 * ph.kf + 1;                   // passes per 4 rows (quad layout)
        RepackArgs A{};
        A.n_mats = n; A.gate_up = gate_up ? 1 : 0; A.nb = nb; A.ksegs = ph.ksegs; A.quad = ph.quad; A.kf = ph.kf;
        int rows = 0;
        const int row_mult = ph.quad ? 4 : 2;
        for (int i = 0; i < n; i++) {
            A.m[i].qs = w[i]->qs; A.m[i].d = w[i]->d; A.m[i].rows = (int)w[i]->rows;
            rows += (int)w[i]->rows;
            if (w[i]->nb != nb || w[i]->type != type) { ok = false; why = "mixed shapes/types in one phase"; }
            if (i + 1 < n && !gate_up && (w[i]->rows % row_mult)) { ok = false; why = "row count inside a fused phase not a multiple of the unit"; }
        }
        ph.n_rows = rows; A.n_rows = rows;
        if (gate_up) {
            ph.n_units = rows / 64; ph.unit_rows = 64;
            ph.unit_passes = ph.quad ? 16 * upq : 32 * ph.ksegs;
            if (rows % 64) { ok = false; why = "n_ff not a multiple of 32"; }
        } else if (ph.quad) {
            ph.n_units = (rows + 3) / 4; ph.unit_passes = upq; ph.unit_rows = 4;
        } else {
            ph.n_units = (rows + 1) / 2; ph.unit_passes = ph.ksegs; ph.unit_rows = 2;
        }
        if (ph.quad) {
            ph.wrapn = upq; ph.sstep = kCW % upq; ph.dstep = 4 * ph.ksegs * (kCW / upq); ph.winc = 4 * ph.ksegs;
            ph.winit = 4 * ph.ksegs;
        } else {
            ph.wrapn = ph.ksegs; ph.sstep = kCW % ph.ksegs; ph.dstep = 2 * (kCW / ph.ksegs) * ph.ksegs + ph.sstep;
            ph.winc = ph.ksegs; ph.winit = 2 * ph.ksegs;
        }
        // tensor parallel: units are split by rank first (contiguous, balanced), then by CTA
        const int u_total = ph.n_units;
        ph.u_off = (int)(((long long)rank * u_total) / size);
        ph.n_units = (int)(((long long)(rank + 1) * u_total) / size) - ph.u_off;
        ph.uq = ph.n_units / sms; ph.ur = ph.n_units % sms;
        const long long max_units = (ph.n_units + sms - 1) / sms;
        if (max_units * ph.unit_rows * ph.ksegs > kPartFloats) { ok = false; why = "phase does not fit the partial table"; }
        if (ph.K > kMaxK) { ok = false; why = "activation row too long"; }
        if (!ok) return ph;
        const size_t n_passes = (size_t)ph.n_units * ph.unit_passes; // this rank's passes only
        A.row_off = ph.u_off * ph.unit_rows;
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
    for (int il = 0; il < L && ok; il++) {
        const ggb_decoder_layer &ly = layers[il];
        const ll_t *x_in = il == 0 ? nullptr : x_ll + (size_t)(il - 1) * E;
        float *sa = io->sa_out + (size_t)il * E, *lo = io->l_out + (size_t)il * E;
        ll_t *qkv_l = qkv_ll + (size_t)il * (QD + 2 * KD);
        ll_t *m_l = m_ll + (size_t)il * H * max_chunks, *l_l = l_ll + (size_t)il * H * max_chunks;
        ll_t *acc_l = acc_ll + (size_t)il * H * max_chunks * d;
        ll_t *kqv_l = kqv_ll + (size_t)il * QD, *sa_l = sa_ll + (size_t)il * E;
        ll_t *ff_l = ff_ll + (size_t)il * (F / 4 + F / 32);
        {   // QKV
            const ggb_weight *w[3] = {ly.attn_q, ly.attn_k, ly.attn_v};
            Phase ph = matvec_phase(PH_QKV, il, 3, w, false);
            ph.in_ll = x_in; ph.norm_w = ly.attn_norm; ph.out_ll = qkv_l;
            if (ph.K != E) { ok = false; why = "attn_q input width != n_embd"; }
            D->phases.push_back(ph);
        }
        {   // ATTN
            Phase ph{};
            ph.kind = PH_ATTN; ph.layer = il; ph.kcache = (__half *)ly.k_cache; ph.vcache = (__half *)ly.v_cache;
            ph.in_ll = qkv_l; ph.acc_ll = acc_l; ph.m_ll = m_l; ph.l_ll = l_l;
            D->phases.push_back(ph);
        }
        {   // COMBINE
            Phase ph{};
            ph.kind = PH_COMBINE; ph.layer = il; ph.f32_out = io->kqv + (size_t)il * QD; ph.out_ll = kqv_l;
            ph.acc_ll = acc_l; ph.m_ll = m_l; ph.l_ll = l_l;
            D->phases.push_back(ph);
        }
        {   // O + residual
            const ggb_weight *w[1] = {ly.attn_output};
            Phase ph = matvec_phase(PH_O, il, 1, w, false);
            ph.in_ll = kqv_l; ph.resid_ll = x_in; ph.resid_embed = il == 0 ? 1 : 0; ph.out = sa; ph.out_ll = sa_l;
            if (ph.K != (int)QD) { ok = false; why = "attn_output input width != n_head*head_dim"; }
            D->phases.push_back(ph);
        }
        {   // GATE_UP
            const ggb_weight *w[2] = {ly.ffn_gate, ly.ffn_up};
            Phase ph = matvec_phase(PH_GATE_UP, il, 2, w, true);
            ph.in_ll = sa_l; ph.norm_w = ly.ffn_norm; ph.out_ll = ff_l;
            ph.f32_out = io->ffn_act ? io->ffn_act + (size_t)il * F : nullptr;
            if (ph.K != E) { ok = false; why = "ffn_gate input width != n_embd"; }
            D->phases.push_back(ph);
        }
        {   // DOWN + residual
            const ggb_weight *w[1] = {ly.ffn_down};
            Phase ph = matvec_phase(PH_DOWN, il, 1, w, false);
            ph.in_ll = ff_l; ph.resid_ll = sa_l; ph.out = lo; ph.out_ll = x_ll + (size_t)il * E;
            if (ph.K != F) { ok = false; why = "ffn_down input width != n_ff"; }
            D->phases.push_back(ph);
        }
    }
    if (ok) {   // LOGITS (tied embedding) + SAMPLE
        const ggb_weight *w[1] = {token_embd};
        Phase ph = matvec_phase(PH_LOGITS, L, 1, w, false);
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
    // two ATTN groups pay off when a CTA has many items per layer (MHA at long context: Gemma-7B Q8_0 n_kv 8192
    // 4.67 -> 4.03 ms/token); at n_kv <= 2048 the one-group kernel with its 5-slot ring is as fast or faster
    D->groups = (Hkv > 1 && dims->n_ctx >= 4096) ? 2 : 1;
    if (const char *e = getenv("GGB_ATTN_GROUPS")) D->groups = atoi(e) == 2 ? 2 : 1;
    D->smem = decoder_smem(type, D->groups);
    P.max_outstanding = std::min(P.max_outstanding, D->groups == 2 ? kSlots - 1 : kSlots);
    for (const Phase &ph : D->phases) D->any_quad = D->any_quad || ph.quad;
    const void *fn = decoder_fn(type, D->any_quad, D->groups);
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
    *out = D;
    return GGB_OK;
}

int ggb_decoder_step(ggb_decoder *D) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(D != nullptr, "null decoder");
    void *args[] = {(void *)&D->P};
    const void *fn = decoder_fn(D->P.type, D->any_quad, D->groups);
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
    const size_t bytes = bytes1 + (size_t)D->grid * 16 * sizeof(long long); // + [grid][16] stamps inside one phase
    GGB_CHECK_ARG((size_t)max_phases * D->grid * 4 * sizeof(long long) >= bytes, "trace buffer too small");
    long long *dev = nullptr;
    GGB_CUDA(cudaMalloc(&dev, bytes));
    GGB_CUDA(cudaMemsetAsync(dev, 0, bytes, ggb_cuda_stream()));
    D->P.trace = dev;
    D->P.dbg = dev + bytes1 / sizeof(long long);
    if (const char *env = getenv("GGB_DBG_PHASE")) D->P.dbg_phase = atoi(env);
    else D->P.dbg_phase = 30;
    D->P.dbg_repeat = getenv("GGB_DBG_REPEAT") ? atoi(getenv("GGB_DBG_REPEAT")) : 1;
    int rc = ggb_decoder_step(D);
    D->P.trace = nullptr;
    D->P.dbg = nullptr;
    D->P.dbg_repeat = 1;
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

int ggb_decoder_free(ggb_decoder *D) {
    if (!D) return GGB_OK;
    cudaStreamSynchronize(ggb_cuda_stream());
    decoder_release(D);
    return GGB_OK;
}

} // extern "C"
