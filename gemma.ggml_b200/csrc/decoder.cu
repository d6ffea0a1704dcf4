// decoder.cu — the persistent decode kernel: ONE cooperative launch per token.
//
// Why: a Gemma-2B Q4_0 token is 1.41 GB of weights + 38 MB of KV = ~220 us of HBM time on a B200,
// spread over 163 mat-vecs and ~450 small nodes (ref: src/gemma_model.cpp:688-740).  Launched as
// separate kernels the HBM pipe drains and refills ~170 times per token and the step takes 1.6 ms.
// Here one CTA per SM stays resident for the whole token:
//
//   * warp 12 lane 0 is the PRODUCER: it walks the static per-CTA tile schedule of the whole token
//     and streams every weight tile (and KV chunk) into a 7 x 27 KiB shared-memory ring with
//     cp.async.bulk (TMA, SASS UBLKCP) completing on mbarriers.  Weights do not depend on
//     activations, so the producer runs AHEAD across phase boundaries: while the consumers sit in a
//     grid barrier or a norm prologue, HBM keeps streaming the next phase's tiles.
//   * warps 0-11 are CONSUMERS: per phase they build their activation registers (rms_norm + Q8_0
//     quantisation is recomputed by every CTA, 8 KB from L2, instead of being a phase of its own),
//     then dot ring tiles exactly like matvec.cu (same lane->block map, same DEV summation order).
//   * phases are separated by a grid barrier (monotonic 64-bit counter in global memory).
//
// Phases per layer: QKV | ATTN (split-K partials; rope, Q scale and the KV-cache store of the new
// position are folded in) | COMBINE (+ Q8_0 quantise) | O (+ residual) | GATE_UP (+ GeGLU + quantise)
// | DOWN (+ residual); then LOGITS (+ per-CTA argmax) | SAMPLE.
// Numerics are bit-identical to the per-operator kernels (same DEV order; tests compare both).
#include <cooperative_groups.h>

#include <vector>

#include "dot.cuh"

namespace {

constexpr int kConsumerWarps = 16;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kThreads = kConsumerThreads + 32;
constexpr int kSlotBytes = 36864; // 32 rows of K=2048 Q4_0 (1152 B/row); 21 of K=3072; 4 of K=16384; 2 of K=24576
constexpr int kSlots = 5;
constexpr int kChunk = GGB_DEV_ATTN_CHUNK;
constexpr int kMaxG = 8;
constexpr int kMaxE = 4096; // n_embd limit of the fused path (one K segment for the norm-fed mat-vecs)

enum PhaseKind { PH_QKV = 0, PH_ATTN, PH_COMBINE, PH_O, PH_GATE_UP, PH_DOWN, PH_LOGITS, PH_SAMPLE };

struct Mat {
    const uint8_t *qs;
    const __half *d;
    int rows;
    int n_tiles;
};

struct PhaseDesc {
    int kind, layer;
    Mat m[3];
    int n_mats;
    int nb;            // blocks per weight row
    int cols;          // nb * 32
    int rows_per_tile; // weight rows per ring slot (GATE_UP: gate rows per sub-tile; the slot also holds the up rows)
    int ksegs;         // ceil(nb / 128)
    int n_tiles;       // schedulable units: tiles (mat-vec), super-tiles of 32 rows (GATE_UP)
    int rot;           // CTA rotation so small phases do not always land on the same SMs
    const float *x_in;    // fp32 input of a norm-fed phase (null in layer 0: embedding row)
    const float *norm_w;
    const uint8_t *act_in; // planar Q8_0 input (O, DOWN)
    float *out[3];
    const float *resid;
    uint8_t *act_out; // planar Q8_0 output (COMBINE, GATE_UP)
    float *f32_out;   // COMBINE: kqv; GATE_UP: optional ffn_act dump; LOGITS: result_norm
    __half *kcache, *vcache;
};

struct DecoderParams {
    const PhaseDesc *phases;
    int n_phases;
    int type; // weight type
    int E, H, Hkv, d, F, V, n_ctx;
    float eps, q_scale, emb_scale;
    // embedding
    const uint8_t *emb_qs;
    const __half *emb_d;
    float *x0;
    // decode state
    int32_t *state, *history;
    // scratch
    float *q, *k, *v; // fp32 Q/K/V of the current token
    float *attn_m, *attn_l, *attn_acc;
    int max_chunks;
    float *cand_v;
    int *cand_i;
    unsigned long long *bar_count, *bar_base;
    const uint16_t *exp_table, *gelu_table;
    const float *rope_cos, *rope_sin;
    long long *trace; // optional [n_phases][grid][4] SM-clock stamps: start, after prologue, after tiles, after barrier
    long long *trace2; // optional [grid][32][3]: LOGITS tiles: before wait, after wait, after compute
};

struct SmemLayout {
    alignas(128) uint8_t ring[kSlots][kSlotBytes];
    alignas(16) uint8_t act[kMaxE + kMaxE / 16]; // planar Q8_0 activation built by the norm prologue
    // 16 KB scratch: the norm prologue's fp32 input row, the attention's rope'd q (8 x 256), the combine's
    // staged partials (2 x 8 KB) — never live at the same time
    alignas(16) float scratch[kMaxE];
    float s[kMaxG][kChunk];
    float e[kMaxG][kChunk];
    alignas(16) __half knew[256];
    alignas(16) __half vnew[256];
    float part[2][2 * kConsumerWarps][kConsumerWarps];
    float stage_g[2][32], stage_u[2][32];
    double red[kConsumerWarps];
    float cand_v[kConsumerWarps];
    int cand_i[kConsumerWarps];
    alignas(8) unsigned long long full_bar[kSlots];
    alignas(8) unsigned long long empty_bar[kSlots];
    alignas(16) PhaseDesc ph[2]; // current / next phase descriptor (global reads would re-miss after every fence)
    alignas(16) PhaseDesc ph_prod; // the producer's own copy (it runs phases ahead of the consumers)
    alignas(16) DecoderParams P;   // kernel parameters (see decoder_kernel)
};
static_assert(sizeof(DecoderParams) % 4 == 0, "DecoderParams is copied word-wise");

// Every function re-derives the layout from the extern __shared__ symbol instead of receiving a reference:
// an out-of-line function that gets a SmemLayout& only sees a generic pointer and emits LD.E/ST.E.
extern __shared__ __align__(1024) uint8_t ggb_dec_smem[];
__device__ __forceinline__ SmemLayout &smem() { return *reinterpret_cast<SmemLayout *>(ggb_dec_smem); }
static_assert(sizeof(PhaseDesc) % 4 == 0, "PhaseDesc is copied word-wise");

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// Producer-side wait: poll with back-off.  The producer is warp 12, which shares scheduler 0 with consumer
// warps 0/4/8 and (highest warp id first) out-prioritises them: a tight try_wait loop there starves a
// quarter of the consumers.  The producer runs tiles ahead, so a little wake-up latency costs nothing.
__device__ __forceinline__ void mbar_wait_backoff(unsigned long long *bar, uint32_t parity) {
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return;
        __nanosleep(500);
    }
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// Ampere-style async copy global -> shared, L2 only (.cg): used to stage small rows without holding registers
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory"); }
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// grid barrier over the consumer threads of all CTAs: bar.sync orders the CTA's writes before
// thread 0's release-add (cumulative at gpu scope); the acquire-poll + bar.sync orders every
// consumer's later reads after the other CTAs' writes.
// Polling uses relaxed loads (an acquire load costs an L1 invalidate — CCTL.IVALL — per poll) and a single
// acquire fence once the count is reached.
__device__ __forceinline__ void grid_barrier(unsigned long long *bar_count, unsigned long long target) {
    consumer_bar();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(bar_count), "l"(1ull) : "memory");
        unsigned long long v;
        do {
            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar_count) : "memory");
        } while (v < target);
        // no acquire needed here: every read of another CTA's data after a barrier is an L2 access
        // (__ldcg / cp.async.cg / TMA), never an L1-cached load, and the writers released at gpu scope
    }
    consumer_bar();
}

// ---------------------------------------------------------------- schedule helpers (producer and consumers agree)
struct TileInfo {
    int mat, row0, nrows;
};
__device__ __forceinline__ TileInfo locate_tile(const PhaseDesc &ph, int t) {
    TileInfo ti;
    int m = 0;
    while (m + 1 < ph.n_mats && t >= ph.m[m].n_tiles) {
        t -= ph.m[m].n_tiles;
        m++;
    }
    ti.mat = m;
    ti.row0 = t * ph.rows_per_tile;
    ti.nrows = min(ph.rows_per_tile, ph.m[m].rows - ti.row0);
    return ti;
}
__device__ __forceinline__ int first_unit(const PhaseDesc &ph, int cta, int G) {
    return ((cta - ph.rot) % G + G) % G;
}

struct RingPos {
    uint32_t seq;
    __device__ __forceinline__ int slot() const { return (int)(seq % kSlots); }
    __device__ __forceinline__ uint32_t parity() const { return (seq / kSlots) & 1u; }
};

// ---------------------------------------------------------------- producer
template <int TYPE>
__device__ __noinline__ void producer_loop(int cta, int G, int n_pos) {
    SmemLayout &S = smem();
    const DecoderParams &P = S.P;
    constexpr int PB = TYPE == GGB_TYPE_Q4_0 ? 16 : 32;
    RingPos rp{0};
    const int nchunk = (n_pos + kChunk - 1) / kChunk;
    for (int pi = 0; pi < P.n_phases; pi++) {
        // private shared-memory copy of the descriptor: one burst of independent loads per phase
        // (a by-value struct would live in local memory because ph.m[] is indexed dynamically)
        {
            constexpr int kWords = (int)(sizeof(PhaseDesc) / 4);
            const uint32_t *src = reinterpret_cast<const uint32_t *>(&P.phases[pi]);
            uint32_t *dst = reinterpret_cast<uint32_t *>(&S.ph_prod);
#pragma unroll 10
            for (int i = 0; i < kWords; i++) dst[i] = src[i];
        }
        const PhaseDesc &ph = S.ph_prod;
        if (ph.kind == PH_QKV || ph.kind == PH_O || ph.kind == PH_DOWN || ph.kind == PH_LOGITS) {
            for (int t = first_unit(ph, cta, G); t < ph.n_tiles; t += G) {
                const TileInfo ti = locate_tile(ph, t);
                const uint32_t pbytes = (uint32_t)ti.nrows * ph.nb * PB, sbytes = (uint32_t)ti.nrows * ph.nb * 2;
                const int s = rp.slot();
                mbar_wait_backoff(&S.empty_bar[s], rp.parity() ^ 1u);
                mbar_expect_tx(&S.full_bar[s], pbytes + sbytes);
                bulk_g2s(S.ring[s], ph.m[ti.mat].qs + (size_t)ti.row0 * ph.nb * PB, pbytes, &S.full_bar[s]);
                bulk_g2s(S.ring[s] + pbytes, ph.m[ti.mat].d + (size_t)ti.row0 * ph.nb, sbytes, &S.full_bar[s]);
                rp.seq++;
            }
        } else if (ph.kind == PH_GATE_UP) {
            for (int st = first_unit(ph, cta, G); st < ph.n_tiles; st += G) {
                for (int r0 = 0; r0 < 32; r0 += ph.rows_per_tile) {
                    const int nrows = min(ph.rows_per_tile, 32 - r0);
                    const size_t row = (size_t)st * 32 + r0;
                    const uint32_t pbytes = (uint32_t)nrows * ph.nb * PB, sbytes = (uint32_t)nrows * ph.nb * 2;
                    const int s = rp.slot();
                    mbar_wait_backoff(&S.empty_bar[s], rp.parity() ^ 1u);
                    mbar_expect_tx(&S.full_bar[s], 2 * (pbytes + sbytes));
                    uint8_t *dst = S.ring[s];
                    bulk_g2s(dst, ph.m[0].qs + row * ph.nb * PB, pbytes, &S.full_bar[s]);
                    bulk_g2s(dst + pbytes, ph.m[1].qs + row * ph.nb * PB, pbytes, &S.full_bar[s]);
                    bulk_g2s(dst + 2 * pbytes, ph.m[0].d + row * ph.nb, sbytes, &S.full_bar[s]);
                    bulk_g2s(dst + 2 * pbytes + sbytes, ph.m[1].d + row * ph.nb, sbytes, &S.full_bar[s]);
                    rp.seq++;
                }
            }
        } else if (ph.kind == PH_ATTN) {
            const int n_items = P.Hkv * nchunk;
            const uint32_t row_bytes = (uint32_t)P.d * 2;
            const size_t kv_stride = (size_t)P.Hkv * P.d;
            for (int it = first_unit(ph, cta, G); it < n_items; it += G) {
                const int g = it / nchunk, c = it % nchunk;
                const int p0 = c * kChunk, np = min(kChunk, n_pos - p0);
                for (int which = 0; which < 2; which++) {
                    const __half *base = (which == 0 ? ph.kcache : ph.vcache) + (size_t)p0 * kv_stride + (size_t)g * P.d;
                    const int s = rp.slot();
                    mbar_wait_backoff(&S.empty_bar[s], rp.parity() ^ 1u);
                    mbar_expect_tx(&S.full_bar[s], (uint32_t)np * row_bytes);
                    if (P.Hkv == 1) {
                        bulk_g2s(S.ring[s], base, (uint32_t)np * row_bytes, &S.full_bar[s]);
                    } else {
                        for (int j = 0; j < np; j++)
                            bulk_g2s(S.ring[s] + (size_t)j * row_bytes, base + (size_t)j * kv_stride, row_bytes,
                                     &S.full_bar[s]);
                    }
                    rp.seq++;
                }
            }
        }
    }
}

// ---------------------------------------------------------------- consumer pieces

// rms_norm (+ weight mul) + quantize_row_q8_0 of the phase input into S.act (every CTA recomputes it).
// Sum of squares in the DEV order: threads 0..255 hold the strided double partials.
// rms_norm (+ weight mul) + quantize_row_q8_0 of the phase input into S.act (every CTA recomputes it).
// Deliberately compact (the persistent kernel is instruction-fetch sensitive): the fp32 row is staged in
// shared memory with cp.async (one L2 round trip, no registers), the sum of squares follows the DEV order
// (thread t < 256 sums elements t, t+256, ... in double; butterfly; warps in order), and the 12 warps then
// quantise one 32-element block per iteration.
__device__ __noinline__ void norm_prologue(int pb, int token, bool write_x0, float *f32_out) {
    SmemLayout &S = smem();
    const DecoderParams &P = S.P;
    const PhaseDesc &ph = S.ph[pb];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int E = P.E;
    float *xs = S.scratch;
    const float *x_in = ph.x_in, *norm_w = ph.norm_w;
    if (x_in) {
        const uint32_t xs_addr = smem_u32(xs);
        for (int v = tid; v < E / 4; v += kConsumerThreads) cp_async16(xs_addr + 16u * v, x_in + 4 * v);
        cp_async_wait_all();
    } else {
        // layer 0: get_rows(token_embd, token) * sqrt(n_embd)   (ref: src/gemma_model.cpp:675-679)
        const int nb = E / 32;
        const size_t blk0 = (size_t)token * nb;
        for (int i = tid; i < E; i += kConsumerThreads) {
            const int b = i >> 5, e = i & 31;
            const float dd = __half2float(P.emb_d[blk0 + b]);
            int qv;
            if (P.type == GGB_TYPE_Q4_0) {
                const int byte = P.emb_qs[(blk0 + b) * 16 + (e & 15)];
                qv = (e < 16 ? (byte & 0x0F) : (byte >> 4)) - 8;
            } else {
                qv = ((const int8_t *)P.emb_qs)[(blk0 + b) * 32 + e];
            }
            xs[i] = __fmul_rn(__fmul_rn((float)qv, dd), P.emb_scale);
        }
    }
    consumer_bar();
    if (tid < GGB_DEV_NORM_THREADS) {
        double acc = 0.0;
#pragma unroll 4
        for (int i = tid; i < E; i += GGB_DEV_NORM_THREADS) {
            const float v = xs[i];
            acc = __dadd_rn(acc, (double)__fmul_rn(v, v));
        }
        acc = warp_sum_f64(acc);
        if (lane == 0) S.red[warp] = acc;
    }
    consumer_bar();
    double sum = S.red[0];
#pragma unroll
    for (int i = 1; i < GGB_DEV_NORM_THREADS / 32; i++) sum = __dadd_rn(sum, S.red[i]);
    const float mean = __double2float_rn(__ddiv_rn(sum, (double)E));
    const float scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, P.eps)));
    // one Q8_0 block (32 consecutive elements) per warp per iteration; next block's weight prefetched
    float w_next = norm_w[32 * warp + lane];
#pragma unroll 1
    for (int b = warp; b < E / 32; b += kConsumerWarps) {
        const int i = 32 * b + lane;
        const float w = w_next;
        if (b + kConsumerWarps < E / 32) w_next = norm_w[i + 32 * kConsumerWarps];
        const float x = xs[i];
        const float v = __fmul_rn(__fmul_rn(x, scale), w);
        if (write_x0) P.x0[i] = x;
        if (f32_out) f32_out[i] = v;
        float d;
        const int q = quantize_q8_lane(v, d);
        ((int8_t *)S.act)[i] = (int8_t)q;
        if (lane == 0) reinterpret_cast<__half *>(S.act + E)[b] = __float2half_rn(d);
    }
    consumer_bar();
}

struct Best {
    float v;
    int i;
};
// per-thread state crosses the out-of-line phase functions BY VALUE (registers): a reference would live in
// local memory, be flushed/reloaded around every asm("memory") barrier, and miss the tiny L1 every time
struct MvRet {
    uint32_t seq;
    Best best;
};
__device__ __forceinline__ Best better(Best a, Best b) {
    if (b.i < 0) return a;
    if (a.i < 0) return b;
    if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
}

// ONE routine for every mat-vec phase (QKV, O, GATE_UP, DOWN, LOGITS).  The persistent kernel executes each
// phase's code once per layer, so its instruction footprint per layer has to stay cache-resident: separate,
// heavily unrolled functions per phase made the kernel instruction-fetch bound (a cold norm prologue took
// ~4 us, the same code re-run hot 0.5 us).
//
// K is cut into segments of 64 blocks (GGB_DEV_SEG_BLOCKS): lane l owns blocks l and l+32 of its warp's
// segment, for EVERY row — 20 registers of activations, loaded once per phase.  A "pass" covers two rows of
// that segment (4 block slots per lane): the four dp4a chains advance in lock step, the two lane partials
// ("blocks in order from 0.0f") go through ONE folded xor-butterfly (lanes < 16 end with row 0's sum, lanes
// >= 16 with row 1's; same additions as two separate butterflies), i.e. the DEV order.  With more than one
// segment (K > 2048) the per-segment sums meet in shared memory and are added in segment order.
// Ring tiles are read with explicit ld.shared; the pass is branch-free.
// GATE_UP slots hold [gate payload][up payload][gate scales][up scales], i.e. they look like a 2R-row tile.
template <int TYPE>
__device__ __noinline__ MvRet matvec_phase(int pb, int cta, int G, uint32_t seq, bool act_in_smem, Best best) {
    SmemLayout &S = smem();
    const DecoderParams &P = S.P;
    const PhaseDesc &ph = S.ph[pb];
    constexpr int PB = TYPE == GGB_TYPE_Q4_0 ? 16 : 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kind = ph.kind, nb = ph.nb, ksegs = ph.ksegs, n_units = ph.n_tiles, rpt = ph.rows_per_tile;
    const bool gate_up = kind == PH_GATE_UP;
    const float *resid = ph.resid;
    const int n_mats = ph.n_mats, rows0 = ph.m[0].rows;
    const int unit0 = first_unit(ph, cta, G);
    if (unit0 >= n_units) return MvRet{seq, best};

    const int kseg = warp % ksegs, group = warp / ksegs; // this warp's segment and row-pair group
    const int groups = kConsumerWarps / ksegs;
    const bool warp_live = group < groups;
    const uint32_t row_pb = (uint32_t)nb * PB, row_sb = (uint32_t)nb * 2;
    const int blk0 = kseg * GGB_DEV_SEG_BLOCKS + lane;
    ActBlock ab[2];
    bool blk_live[2];
#pragma unroll
    for (int b = 0; b < 2; b++) {
        blk_live[b] = warp_live && blk0 + 32 * b < nb;
        if (act_in_smem) load_act(ab[b], S.act, ph.cols, blk0 + 32 * b, blk_live[b]);
        else load_act_cg(ab[b], ph.act_in, ph.cols, blk0 + 32 * b, blk_live[b]);
    }
    const uint32_t ring0 = smem_u32(&S.ring[0][0]);
    const uint32_t lane_pay = (uint32_t)blk0 * PB, lane_scl = (uint32_t)blk0 * 2;
    const int half = lane >> 4; // which row of a pass this lane publishes
    const int n_sub = gate_up ? (32 + rpt - 1) / rpt : 1;
    int pbuf = 0, sbuf = 0, ord = 0;
    int slot = (int)(seq % kSlots);
    uint32_t parity = (seq / kSlots) & 1u;

    // hand one finished row (index r of the slot) to its consumer: residual add + store, or the GeGLU staging
    auto publish = [&](int r, float v, float rv, int mat, int row0, int nrows) {
        if (gate_up) {
            if (r < nrows) S.stage_g[sbuf][row0 + r] = v;
            else S.stage_u[sbuf][row0 + r - nrows] = v;
        } else {
            if (resid) v = __fadd_rn(v, rv);
            ph.out[mat][row0 + r] = v;
            if (kind == PH_LOGITS) best = better(best, Best{v, row0 + r});
        }
    };

    for (int u = unit0; u < n_units; u += G, ord++) {
        for (int sub = 0; sub < n_sub; sub++) {
            int mat = 0, row0, nrows;
            if (gate_up) {
                row0 = sub * rpt;
                nrows = min(rpt, 32 - row0);
            } else if (n_mats == 1) {
                row0 = u * rpt;
                nrows = min(rpt, rows0 - row0);
            } else {
                const TileInfo ti = locate_tile(ph, u);
                mat = ti.mat; row0 = ti.row0; nrows = ti.nrows;
            }
            const int slot_rows = gate_up ? 2 * nrows : nrows;
            const int s = slot;
            const uint32_t par = parity;
            const uint32_t pay0 = ring0 + (uint32_t)s * kSlotBytes + lane_pay;
            const uint32_t scl0 = ring0 + (uint32_t)s * kSlotBytes + (uint32_t)slot_rows * row_pb + lane_scl;
            // residual of the row this lane will publish, requested before the tile is waited for
            const int my_first = 2 * group;
            const int my_row = ksegs == 1 ? my_first + half : tid; // ksegs > 1: thread t < slot_rows finishes row t
            const bool my_pub = ksegs == 1 ? ((lane & 15) == 0 && my_row < slot_rows) : (tid < slot_rows);
            float rv = 0.0f;
            if (my_pub && resid) rv = __ldcg(resid + row0 + my_row);
            const bool dbg = P.trace2 && kind == PH_LOGITS && ord < 31 && tid == 0;
            if (dbg) P.trace2[((size_t)cta * 32 + ord) * 3 + 0] = clock64();
            mbar_wait(&S.full_bar[s], par);
            if (dbg) P.trace2[((size_t)cta * 32 + ord) * 3 + 1] = clock64();
#pragma unroll 1
            for (int first = my_first; first < slot_rows; first += 2 * groups) {
                uint4 w0[4], w1[4];
                float dx[4];
                const uint32_t pbase = pay0 + (uint32_t)first * row_pb, sbase = scl0 + (uint32_t)first * row_sb;
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const uint32_t po = (uint32_t)(t >> 1) * row_pb + (uint32_t)(t & 1) * 32 * PB;
                    w0[t] = lds_u4(pbase + po);
                    if (TYPE == GGB_TYPE_Q8_0) w1[t] = lds_u4(pbase + po + 16);
                    dx[t] = lds_half_as_float(sbase + (uint32_t)(t >> 1) * row_sb + (uint32_t)(t & 1) * 64);
                }
                float term[4];
                pass_terms<TYPE, 2>(w0, w1, dx, ab, term);
                const bool row1 = first + 1 < slot_rows;
                // lane partials: blocks in order from 0.0f, dead blocks dropped
                float a0 = 0.0f, a1 = 0.0f;
                a0 = blk_live[0] ? __fadd_rn(a0, term[0]) : a0;
                a0 = blk_live[1] ? __fadd_rn(a0, term[1]) : a0;
                a1 = (blk_live[0] && row1) ? __fadd_rn(a1, term[2]) : a1;
                a1 = (blk_live[1] && row1) ? __fadd_rn(a1, term[3]) : a1;
                // folded xor-butterfly 16,8,4,2,1
                const float send = half ? a0 : a1;
                float keep = __fadd_rn(half ? a1 : a0, __shfl_xor_sync(0xffffffffu, send, 16));
#pragma unroll
                for (int off = 8; off >= 1; off >>= 1) keep = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, keep, off));
                const int r = first + half;
                if ((lane & 15) == 0 && r < slot_rows && warp_live) {
                    if (ksegs > 1) S.part[pbuf][r][kseg] = keep;
                    else publish(r, keep, first == my_first ? rv : (resid ? __ldcg(resid + row0 + r) : 0.0f), mat, row0, nrows);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.empty_bar[s]);
            if (dbg) P.trace2[((size_t)cta * 32 + ord) * 3 + 2] = clock64();
            if (ksegs > 1) {
                consumer_bar();
                if (tid < slot_rows) { // thread r: segment sums of row r added in segment order
                    float total = S.part[pbuf][tid][0];
                    for (int k = 1; k < ksegs; k++) total = __fadd_rn(total, S.part[pbuf][tid][k]);
                    publish(tid, total, rv, mat, row0, nrows);
                }
                pbuf ^= 1;
            }
            seq++;
            if (++slot == kSlots) {
                slot = 0;
                parity ^= 1u;
            }
        }
        if (gate_up) { // GeGLU (fp16 gelu table) + quantize_row_q8_0 of the super-tile's 32 products, lane-parallel
            consumer_bar();
            if (warp == 0) {
                const float gl = half_bits_to_float(P.gelu_table[float_to_half_bits(S.stage_g[sbuf][lane])]);
                const float x = __fmul_rn(gl, S.stage_u[sbuf][lane]);
                const size_t row = (size_t)u * 32 + lane;
                if (ph.f32_out) ph.f32_out[row] = x;
                float d;
                const int q = quantize_q8_lane(x, d);
                ((int8_t *)ph.act_out)[row] = (int8_t)q;
                if (lane == 0) reinterpret_cast<__half *>(ph.act_out + P.F)[u] = __float2half_rn(d);
            }
            sbuf ^= 1;
        }
    }
    return MvRet{seq, best};
}

// ATTN: items (kv head g, chunk of 32 positions); K chunk and V chunk arrive in two ring slots.
// Same arithmetic as attention.cu attn_partial_kernel; rope + Q scale are applied to q on load, and the
// item that owns the newest position ropes K, rounds K/V to fp16, stores them in the cache and uses them.
__device__ __noinline__ uint32_t attn_phase(int pb, int cta, int G, uint32_t seq_in, int n_pos) {
    SmemLayout &S = smem();
    const DecoderParams &P = S.P;
    const PhaseDesc &ph = S.ph[pb];
    RingPos rp{seq_in};
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = P.d, halfd = d / 2, Gq = P.H / P.Hkv;
    const int nchunk = (n_pos + kChunk - 1) / kChunk;
    const int n_items = P.Hkv * nchunk;
    const int pos = n_pos - 1;
    const float *cosr = P.rope_cos + (size_t)pos * halfd, *sinr = P.rope_sin + (size_t)pos * halfd;
    float(*Sq)[256] = reinterpret_cast<float(*)[256]>(S.scratch);
#pragma unroll 1
    for (int it = first_unit(ph, cta, G); it < n_items; it += G) {
        const int g = it / nchunk, c = it % nchunk;
        const int p0 = c * kChunk, np = min(kChunk, n_pos - p0);
        const int h0 = g * Gq;
        const bool owns_new = (pos >= p0 && pos < p0 + np);
        // q: rope (NEOX pairs i, i + d/2), scale, round to fp16 (ggml mul_mat INIT converts src1 to fp16)
        {
            constexpr int kQIter = (kMaxG * 128 + kConsumerThreads - 1) / kConsumerThreads;
            float x0[kQIter], x1[kQIter], cs[kQIter], sn[kQIter];
#pragma unroll
            for (int it = 0; it < kQIter; it++) { // all loads first: one L2 round trip
                const int i = tid + kConsumerThreads * it;
                if (i < Gq * halfd) {
                    const int jh = i / halfd, e = i % halfd;
                    const float *qh = P.q + (size_t)(h0 + jh) * d;
                    x0[it] = __ldcg(qh + e);
                    x1[it] = __ldcg(qh + e + halfd);
                    cs[it] = cosr[e];
                    sn[it] = sinr[e];
                }
            }
#pragma unroll
            for (int it = 0; it < kQIter; it++) {
                const int i = tid + kConsumerThreads * it;
                if (i < Gq * halfd) {
                    const int jh = i / halfd, e = i % halfd;
                    const float r0 = __fmul_rn(__fsub_rn(__fmul_rn(x0[it], cs[it]), __fmul_rn(x1[it], sn[it])), P.q_scale);
                    const float r1 = __fmul_rn(__fadd_rn(__fmul_rn(x0[it], sn[it]), __fmul_rn(x1[it], cs[it])), P.q_scale);
                    Sq[jh][e] = __half2float(__float2half_rn(r0));
                    Sq[jh][e + halfd] = __half2float(__float2half_rn(r1));
                }
            }
        }
        if (owns_new) { // KV store of the new position (ref: src/gemma_model.cpp:499-518)
            for (int e = tid; e < halfd; e += kConsumerThreads) {
                const float x0 = __ldcg(P.k + (size_t)g * d + e), x1 = __ldcg(P.k + (size_t)g * d + e + halfd);
                const float cs = cosr[e], sn = sinr[e];
                const __half k0 = __float2half_rn(__fsub_rn(__fmul_rn(x0, cs), __fmul_rn(x1, sn)));
                const __half k1 = __float2half_rn(__fadd_rn(__fmul_rn(x0, sn), __fmul_rn(x1, cs)));
                S.knew[e] = k0;
                S.knew[e + halfd] = k1;
                __half *krow = ph.kcache + ((size_t)pos * P.Hkv + g) * d;
                krow[e] = k0;
                krow[e + halfd] = k1;
            }
            for (int e = tid; e < d; e += kConsumerThreads) {
                const __half vv = __float2half_rn(__ldcg(P.v + (size_t)g * d + e));
                S.vnew[e] = vv;
                ph.vcache[((size_t)pos * P.Hkv + g) * d + e] = vv;
            }
        }
        consumer_bar();
        // scores: warp w takes positions w, w+12, w+24
        {
            const int s = rp.slot();
            mbar_wait(&S.full_bar[s], rp.parity());
            const __half *kchunk = reinterpret_cast<const __half *>(S.ring[s]);
            for (int j = warp; j < np; j += kConsumerWarps) {
                float kf[8];
                if (8 * lane < d) {
                    const __half *krow = (owns_new && p0 + j == pos) ? S.knew : kchunk + (size_t)j * d;
                    const uint4 raw = *reinterpret_cast<const uint4 *>(krow + 8 * lane);
                    const __half2 *h2 = reinterpret_cast<const __half2 *>(&raw);
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const float2 f = __half22float2(h2[e]);
                        kf[2 * e] = f.x;
                        kf[2 * e + 1] = f.y;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 8; e++) kf[e] = 0.f;
                }
                // (fp16 x fp16 products are exact in fp32, so fmaf(k, q, acc) == acc + k*q rounded once: same bits)
                // heads in groups of 4: lane partials first, then the group's butterflies interleaved (ILP
                // without unrolling all 8 heads: code size matters here); each sum follows the DEV order
#pragma unroll 1
                for (int j0 = 0; j0 < Gq; j0 += 4) {
                    float acc[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        acc[u] = 0.0f;
                        if (j0 + u < Gq && 8 * lane < d) {
                            const float4 qa = *reinterpret_cast<const float4 *>(&Sq[j0 + u][8 * lane]);
                            const float4 qb = *reinterpret_cast<const float4 *>(&Sq[j0 + u][8 * lane + 4]);
                            acc[u] = __fmaf_rn(kf[0], qa.x, acc[u]);
                            acc[u] = __fmaf_rn(kf[1], qa.y, acc[u]);
                            acc[u] = __fmaf_rn(kf[2], qa.z, acc[u]);
                            acc[u] = __fmaf_rn(kf[3], qa.w, acc[u]);
                            acc[u] = __fmaf_rn(kf[4], qb.x, acc[u]);
                            acc[u] = __fmaf_rn(kf[5], qb.y, acc[u]);
                            acc[u] = __fmaf_rn(kf[6], qb.z, acc[u]);
                            acc[u] = __fmaf_rn(kf[7], qb.w, acc[u]);
                        }
                    }
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
                        for (int u = 0; u < 4; u++) acc[u] = __fadd_rn(acc[u], __shfl_xor_sync(0xffffffffu, acc[u], off));
                    }
                    if (lane == 0) {
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            if (j0 + u < Gq) S.s[j0 + u][j] = acc[u];
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.empty_bar[s]);
            rp.seq++;
        }
        consumer_bar();
        // per-head chunk soft-max through the fp16 exp table
        if (warp < Gq) {
            const float sv = lane < np ? S.s[warp][lane] : -INFINITY;
            const float m = warp_max_f32(sv);
            const float e = lane < np ? half_bits_to_float(P.exp_table[float_to_half_bits(__fsub_rn(sv, m))]) : 0.0f;
            const float l = warp_sum_f32(e);
            S.e[warp][lane] = e;
            if (lane == 0) {
                P.attn_m[(size_t)(h0 + warp) * P.max_chunks + c] = m;
                P.attn_l[(size_t)(h0 + warp) * P.max_chunks + c] = l;
            }
        }
        consumer_bar();
        // acc[h][i] = sum_p e[h][p] * V[p][i], sequential in p
        {
            const int s = rp.slot();
            mbar_wait(&S.full_bar[s], rp.parity());
            const __half *vchunk = reinterpret_cast<const __half *>(S.ring[s]);
            // 4 thread groups of 128: group hs takes heads hs, hs+4; thread -> one pair of dims
            constexpr int kHG = kConsumerThreads / 128; // 4
            const int dp = tid % 128, hs = tid / 128;
            if (dp < halfd) {
                float a0[kMaxG / kHG], a1[kMaxG / kHG];
#pragma unroll
                for (int u = 0; u < kMaxG / kHG; u++) a0[u] = a1[u] = 0.0f;
#pragma unroll 2
                for (int j = 0; j < np; j++) {
                    const __half *vrow = (owns_new && p0 + j == pos) ? S.vnew : vchunk + (size_t)j * d;
                    const float2 vf = __half22float2(*reinterpret_cast<const __half2 *>(vrow + 2 * dp));
#pragma unroll
                    for (int u = 0; u < kMaxG / kHG; u++) {
                        const int jh = kHG * u + hs;
                        if (jh < Gq) { // e is an fp16 value: e*v is exact, fmaf == mul + add
                            const float e = S.e[jh][j];
                            a0[u] = __fmaf_rn(e, vf.x, a0[u]);
                            a1[u] = __fmaf_rn(e, vf.y, a1[u]);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < kMaxG / kHG; u++) {
                    const int jh = kHG * u + hs;
                    if (jh < Gq) {
                        float *out = P.attn_acc + ((size_t)(h0 + jh) * P.max_chunks + c) * d + 2 * dp;
                        *reinterpret_cast<float2 *>(out) = make_float2(a0[u], a1[u]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.empty_bar[s]);
            rp.seq++;
        }
        consumer_bar();
    }
    return rp.seq;
}

// COMBINE: one warp per (head, 32-dim slice): M = max m_c; w_c = exp_table[f16(m_c - M)];
// L = sum l_c*w_c; O = sum acc_c*w_c (sequential in c); out = O / L; the slice is one Q8_0 block.
// Compact on purpose: the warp stages its 64-chunk window of partials in shared memory with cp.async
// (one L2 round trip), then runs the sequential chain from shared memory.
__device__ __noinline__ void combine_phase(int pb, int cta, int G, int n_pos) {
    SmemLayout &S = smem();
    const DecoderParams &P = S.P;
    const PhaseDesc &ph = S.ph[pb];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d = P.d, slices = d / 32;
    const int n_tasks = P.H * slices;
    const int nchunk = (n_pos + kChunk - 1) / kChunk;
    if (warp >= 2) return; // two 8 KB staging windows -> at most two tasks in flight per CTA
    float *win = S.scratch + warp * 2048; // 64 chunks x 32 floats
    const uint32_t win_addr = smem_u32(win);
#pragma unroll 1
    for (int task = warp * G + ((cta - ph.rot) % G + G) % G; task < n_tasks; task += 2 * G) {
        const int h = task / slices, sl = task % slices;
        const float *pm = P.attn_m + (size_t)h * P.max_chunks;
        const float *pl = P.attn_l + (size_t)h * P.max_chunks;
        const float *pa = P.attn_acc + (size_t)h * P.max_chunks * d + sl * 32;
        float mx = -INFINITY;
#pragma unroll 4
        for (int c = lane; c < nchunk; c += 32) mx = fmaxf(mx, __ldcg(pm + c));
        mx = warp_max_f32(mx);
        float L = 0.0f, O = 0.0f;
#pragma unroll 1
        for (int c0 = 0; c0 < nchunk; c0 += 64) {
            const int nc = min(64, nchunk - c0);
            // stage nc x 128 B: 8 x 16 B pieces per chunk
            for (int v = lane; v < nc * 8; v += 32)
                cp_async16(win_addr + 16u * v, pa + (size_t)(c0 + (v >> 3)) * d + 4 * (v & 7));
            float wl[2], ll[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int c = c0 + 32 * u + lane;
                const bool ok = 32 * u + lane < nc;
                wl[u] = ok ? half_bits_to_float(P.exp_table[float_to_half_bits(__fsub_rn(__ldcg(pm + c), mx))]) : 0.0f;
                ll[u] = ok ? __ldcg(pl + c) : 0.0f;
            }
            cp_async_wait_all();
            __syncwarp();
#pragma unroll 1
            for (int u = 0; u < 2; u++) {
                const int n = min(32, nc - 32 * u);
#pragma unroll 4
                for (int j = 0; j < n; j++) {
                    const float w = __shfl_sync(0xffffffffu, wl[u], j);
                    const float l = __shfl_sync(0xffffffffu, ll[u], j);
                    L = __fadd_rn(L, __fmul_rn(l, w));
                    O = __fadd_rn(O, __fmul_rn(win[(32 * u + j) * 32 + lane], w));
                }
            }
            __syncwarp();
        }
        const float r = __fdiv_rn(O, L);
        const size_t idx = (size_t)h * d + sl * 32 + lane;
        ph.f32_out[idx] = r;
        float dd;
        const int qv = quantize_q8_lane(r, dd);
        ((int8_t *)ph.act_out)[idx] = (int8_t)qv;
        if (lane == 0) reinterpret_cast<__half *>(ph.act_out + (size_t)P.H * d)[idx >> 5] = __float2half_rn(dd);
    }
}

template <int TYPE>
__device__ __forceinline__ MvRet run_matvec(int pb, int cta, int G, uint32_t seq, bool act_in_smem, Best best) {
    return matvec_phase<TYPE>(pb, cta, G, seq, act_in_smem, best);
}

// 17 warps: the register file is allocated in units of 4 warps (20 here), so the per-thread budget is 96.  With
// 212 KB of the SM's 228 KB carved out as shared memory the L1 left for local memory is tiny and a spill
// is an L2 round trip: the phase functions are kept out of line and templated so none of them spills.
template <int TYPE>
__global__ void __launch_bounds__(kThreads, 1) decoder_kernel(const __grid_constant__ DecoderParams Pk) {
    SmemLayout &S = smem();
    // the parameter block is copied to shared memory: the out-of-line phase functions take it by reference,
    // and a reference to kernel parameters would be a local-memory copy
    {
        constexpr int kWords = (int)(sizeof(DecoderParams) / 4);
        for (int i = threadIdx.x; i < kWords; i += kThreads)
            reinterpret_cast<uint32_t *>(&S.P)[i] = reinterpret_cast<const uint32_t *>(&Pk)[i];
    }
    __syncthreads();
    const DecoderParams &P = S.P;
    const int cta = blockIdx.x, G = gridDim.x;
    const int tid = threadIdx.x, warp = tid >> 5;

    if (tid == 0) {
        for (int i = 0; i < kSlots; i++) {
            mbar_init(&S.full_bar[i], 1);
            mbar_init(&S.empty_bar[i], kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int token = min(max(P.state[0], 0), P.V - 1);
    const int n_pos = min(max(P.state[1], 0), P.n_ctx - 1) + 1;
    const unsigned long long base = *P.bar_base;
    unsigned long long *const bar_count = P.bar_count;

    if (warp == kConsumerWarps) {
        if (tid == kConsumerThreads) producer_loop<TYPE>(cta, G, n_pos);
        return;
    }

    // ---- consumers
    uint32_t seq = 0;
    unsigned long long bar_target = base;
    Best best{-INFINITY, -1};
    constexpr int kPhWords = (int)(sizeof(PhaseDesc) / 4);
    if (tid < kPhWords) reinterpret_cast<uint32_t *>(&S.ph[0])[tid] = reinterpret_cast<const uint32_t *>(&P.phases[0])[tid];
    consumer_bar();
    for (int pi = 0; pi < P.n_phases; pi++) {
        const PhaseDesc &ph = S.ph[pi & 1];
        // prefetch the next descriptor; it is published by the barriers at the end of this phase
        if (pi + 1 < P.n_phases && tid < kPhWords)
            reinterpret_cast<uint32_t *>(&S.ph[(pi + 1) & 1])[tid] =
                reinterpret_cast<const uint32_t *>(&P.phases[pi + 1])[tid];
        long long tr0 = 0, tr1 = 0;
        if (P.trace) tr0 = tr1 = clock64();
        switch (ph.kind) {
        case PH_QKV:
        case PH_LOGITS: {
            const bool mine = first_unit(ph, cta, G) < ph.n_tiles;
            const bool write_x0 = (ph.kind == PH_QKV && ph.x_in == nullptr && cta == 0);
            // LOGITS has rot = 0, so CTA 0 owns tile 0 and also writes the result_norm dump
            if (mine || write_x0)
                norm_prologue(pi & 1, token, write_x0, (ph.kind == PH_LOGITS && cta == 0) ? ph.f32_out : nullptr);
            if (P.trace) tr1 = clock64();
            {
                const MvRet r = run_matvec<TYPE>(pi & 1, cta, G, seq, true, best);
                seq = r.seq;
                best = r.best;
            }
            break;
        }
        case PH_GATE_UP: {
            if (first_unit(ph, cta, G) < ph.n_tiles) norm_prologue(pi & 1, token, false, nullptr);
            if (P.trace) tr1 = clock64();
            seq = run_matvec<TYPE>(pi & 1, cta, G, seq, true, best).seq;
            break;
        }
        case PH_O:
        case PH_DOWN:
            {
                const MvRet r = run_matvec<TYPE>(pi & 1, cta, G, seq, false, best);
                seq = r.seq;
                best = r.best;
            }
            break;
        case PH_ATTN:
            seq = attn_phase(pi & 1, cta, G, seq, n_pos);
            break;
        case PH_COMBINE:
            combine_phase(pi & 1, cta, G, n_pos);
            break;
        case PH_SAMPLE: {
            // per-CTA candidates were published before the previous barrier; CTA 0 picks the first maximum
            if (cta == 0 && warp == 0) {
                const int lane = tid & 31;
                Best b{-INFINITY, -1};
                for (int i = lane; i < G; i += 32) b = better(b, Best{__ldcg(P.cand_v + i), __ldcg(P.cand_i + i)});
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    Best o;
                    o.v = __shfl_xor_sync(0xffffffffu, b.v, off);
                    o.i = __shfl_xor_sync(0xffffffffu, b.i, off);
                    b = better(b, o);
                }
                if (lane == 0) {
                    P.state[0] = b.i;
                    P.state[1] += P.state[2];
                    if (P.history) P.history[P.state[1]] = b.i;
                }
            }
            break;
        }
        }
        if (ph.kind == PH_LOGITS) { // publish this CTA's arg-max candidate (strict '>' / lowest index, ref: :538-543)
            const int lane = tid & 31;
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
                for (int w = 1; w < kConsumerWarps; w++) b = better(b, Best{S.cand_v[w], S.cand_i[w]});
                P.cand_v[cta] = b.v;
                P.cand_i[cta] = b.i;
            }
        }
        long long tr2 = 0;
        if (P.trace) tr2 = clock64();
        if (ph.kind != PH_SAMPLE) {
            bar_target += (unsigned long long)G;
            grid_barrier(bar_count, bar_target);
        }
        if (P.trace && tid == 0) {
            long long *t = P.trace + ((size_t)pi * G + cta) * 4;
            t[0] = tr0; t[1] = tr1; t[2] = tr2; t[3] = clock64();
        }
    }
    if (cta == 0 && tid == 0) *P.bar_base = bar_target;
}

} // namespace

// ---------------------------------------------------------------- host side
struct ggb_decoder {
    DecoderParams P{};
    std::vector<PhaseDesc> phases;
    PhaseDesc *d_phases = nullptr;
    void *scratch = nullptr;
    int grid = 0;
    size_t smem = 0;
    int n_barriers = 0;
};

extern "C" {

int ggb_decoder_create(const ggb_decoder_dims *dims, const ggb_decoder_layer *layers, const ggb_weight *token_embd,
                       const float *output_norm, const ggb_decoder_io *io, ggb_decoder **out) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(dims && layers && token_embd && output_norm && io && out, "null pointer");
    const int E = dims->n_embd, L = dims->n_layer, H = dims->n_head, Hkv = dims->n_head_kv, d = dims->head_dim;
    const int F = dims->n_ff, V = dims->n_vocab;
    const int type = token_embd->type;
    const int PB = ggb_payload_bytes(type);
    if (!(E % 256 == 0 && E <= kMaxE && (H * d) % 256 == 0 && H * d <= kMaxE && F % 256 == 0 && d % 32 == 0 &&
          d <= 256 && Hkv > 0 && H % Hkv == 0 && H / Hkv <= kMaxG && F % 32 == 0 && F / 32 * 2 % 16 == 0)) {
        ggb_set_error("decoder: shape outside the fused kernel's envelope (E=%d H=%d Hkv=%d d=%d F=%d)", E, H, Hkv, d, F);
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
    P.max_chunks = (dims->n_ctx + kChunk - 1) / kChunk;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    D->grid = sms;
    // scratch: x0 | q | k | v | attn m,l,acc | cand | barrier words | planar activations
    const size_t QD = (size_t)H * d, KD = (size_t)Hkv * d;
    const size_t act_q_bytes = ggb_q8_row_bytes((int64_t)QD), act_f_bytes = ggb_q8_row_bytes((int64_t)F);
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
    const size_t o_x0 = carve((size_t)E * 4), o_q = carve(QD * 4), o_k = carve(KD * 4), o_v = carve(KD * 4);
    const size_t o_m = carve((size_t)H * P.max_chunks * 4), o_l = carve((size_t)H * P.max_chunks * 4);
    const size_t o_acc = carve((size_t)H * P.max_chunks * d * 4);
    const size_t o_cv = carve((size_t)sms * 4), o_ci = carve((size_t)sms * 4), o_bar = carve(16);
    const size_t o_aq = carve(act_q_bytes), o_af = carve(act_f_bytes);
    cudaError_t e = cudaMalloc(&D->scratch, off);
    if (e != cudaSuccess) {
        ggb_set_error("decoder scratch cudaMalloc(%zu) -> %s", off, cudaGetErrorString(e));
        delete D;
        return GGB_ERR_CUDA;
    }
    cudaMemsetAsync(D->scratch, 0, off, ggb_cuda_stream());
    uint8_t *sb = (uint8_t *)D->scratch;
    P.x0 = (float *)(sb + o_x0); P.q = (float *)(sb + o_q); P.k = (float *)(sb + o_k); P.v = (float *)(sb + o_v);
    P.attn_m = (float *)(sb + o_m); P.attn_l = (float *)(sb + o_l); P.attn_acc = (float *)(sb + o_acc);
    P.cand_v = (float *)(sb + o_cv); P.cand_i = (int *)(sb + o_ci);
    P.bar_count = (unsigned long long *)(sb + o_bar); P.bar_base = P.bar_count + 1;
    uint8_t *act_q = sb + o_aq, *act_f = sb + o_af;

    auto fit_rows = [&](int nb, int mats) { // rows per ring slot
        const int row_bytes = nb * (PB + 2) * mats;
        int r = kSlotBytes / row_bytes;
        return r < 1 ? 0 : r;
    };
    auto matvec_desc = [&](int kind, int layer, int n, const ggb_weight *const *w, float *const *outs) {
        PhaseDesc ph{};
        ph.kind = kind; ph.layer = layer; ph.n_mats = n;
        ph.nb = (int)w[0]->nb; ph.cols = (int)w[0]->cols;
        ph.ksegs = (ph.nb + GGB_DEV_SEG_BLOCKS - 1) / GGB_DEV_SEG_BLOCKS;
        int r = fit_rows(ph.nb, 1);
        const int rpr = kConsumerWarps / ph.ksegs > 0 ? kConsumerWarps / ph.ksegs : 1;
        if (r > 2 * kConsumerWarps) r = 2 * kConsumerWarps;
        if (r >= rpr) r = r / rpr * rpr; // whole rounds of the consumer warps
        ph.rows_per_tile = r;
        ph.n_tiles = 0;
        for (int i = 0; i < n; i++) {
            ph.m[i].qs = w[i]->qs; ph.m[i].d = w[i]->d; ph.m[i].rows = (int)w[i]->rows;
            ph.m[i].n_tiles = r > 0 ? (int)((w[i]->rows + r - 1) / r) : 0;
            ph.n_tiles += ph.m[i].n_tiles;
            ph.out[i] = outs[i];
        }
        return ph;
    };
    bool ok = true;
    for (int il = 0; il < L && ok; il++) {
        const ggb_decoder_layer &ly = layers[il];
        const float *x_in = il == 0 ? nullptr : io->l_out + (size_t)(il - 1) * E;
        const float *x_res = il == 0 ? P.x0 : x_in;
        float *sa = io->sa_out + (size_t)il * E, *lo = io->l_out + (size_t)il * E;
        {   // QKV
            const ggb_weight *w[3] = {ly.attn_q, ly.attn_k, ly.attn_v};
            float *o[3] = {P.q, P.k, P.v};
            PhaseDesc ph = matvec_desc(PH_QKV, il, 3, w, o);
            ph.x_in = x_in; ph.norm_w = ly.attn_norm;
            ok = ok && ph.rows_per_tile > 0 && ph.ksegs <= kConsumerWarps;
            D->phases.push_back(ph);
        }
        {   // ATTN
            PhaseDesc ph{};
            ph.kind = PH_ATTN; ph.layer = il; ph.kcache = (__half *)ly.k_cache; ph.vcache = (__half *)ly.v_cache;
            D->phases.push_back(ph);
        }
        {   // COMBINE
            PhaseDesc ph{};
            ph.kind = PH_COMBINE; ph.layer = il; ph.act_out = act_q; ph.f32_out = io->kqv + (size_t)il * QD;
            D->phases.push_back(ph);
        }
        {   // O + residual
            const ggb_weight *w[1] = {ly.attn_output};
            float *o[1] = {sa};
            PhaseDesc ph = matvec_desc(PH_O, il, 1, w, o);
            ph.act_in = act_q; ph.resid = x_res;
            ok = ok && ph.rows_per_tile > 0;
            D->phases.push_back(ph);
        }
        {   // GATE_UP
            PhaseDesc ph{};
            ph.kind = PH_GATE_UP; ph.layer = il; ph.n_mats = 2;
            ph.nb = (int)ly.ffn_gate->nb; ph.cols = (int)ly.ffn_gate->cols;
            ph.ksegs = (ph.nb + GGB_DEV_SEG_BLOCKS - 1) / GGB_DEV_SEG_BLOCKS;
            ph.m[0].qs = ly.ffn_gate->qs; ph.m[0].d = ly.ffn_gate->d; ph.m[0].rows = F;
            ph.m[1].qs = ly.ffn_up->qs; ph.m[1].d = ly.ffn_up->d; ph.m[1].rows = F;
            int r = fit_rows(ph.nb, 2);
            if (r > kConsumerWarps) r = kConsumerWarps;
            if (r > 8 && r < kConsumerWarps) r = 8;
            ph.rows_per_tile = r;
            ph.n_tiles = F / 32;
            ph.x_in = sa; ph.norm_w = ly.ffn_norm; ph.act_out = act_f;
            ph.f32_out = io->ffn_act ? io->ffn_act + (size_t)il * F : nullptr;
            ok = ok && r > 0 && ph.ksegs <= kConsumerWarps;
            D->phases.push_back(ph);
        }
        {   // DOWN + residual
            const ggb_weight *w[1] = {ly.ffn_down};
            float *o[1] = {lo};
            PhaseDesc ph = matvec_desc(PH_DOWN, il, 1, w, o);
            ph.act_in = act_f; ph.resid = sa;
            ok = ok && ph.rows_per_tile > 0 && ph.ksegs <= kConsumerWarps;
            D->phases.push_back(ph);
        }
    }
    {   // LOGITS (tied embedding) + SAMPLE
        const ggb_weight *w[1] = {token_embd};
        float *o[1] = {io->logits};
        PhaseDesc ph = matvec_desc(PH_LOGITS, L, 1, w, o);
        ph.x_in = io->l_out + (size_t)(L - 1) * E; ph.norm_w = output_norm; ph.f32_out = io->result_norm;
        ok = ok && ph.rows_per_tile > 0 && ph.ksegs <= kConsumerWarps;
        D->phases.push_back(ph);
        PhaseDesc sp{};
        sp.kind = PH_SAMPLE; sp.layer = L;
        D->phases.push_back(sp);
    }
    if (!ok) {
        ggb_set_error("decoder: a weight row does not fit the ring slot (%d B) for this shape", kSlotBytes);
        cudaFree(D->scratch);
        delete D;
        return GGB_ERR_UNSUPPORTED;
    }
    for (size_t i = 0; i < D->phases.size(); i++) D->phases[i].rot = (int)((i * 61) % (size_t)sms);
    D->phases[D->phases.size() - 2].rot = 0; // LOGITS: result_norm dump is written by the CTA owning tile 0
    D->n_barriers = (int)D->phases.size() - 1;
    e = cudaMalloc(&D->d_phases, D->phases.size() * sizeof(PhaseDesc));
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(D->d_phases, D->phases.data(), D->phases.size() * sizeof(PhaseDesc), cudaMemcpyHostToDevice,
                            ggb_cuda_stream());
    if (e != cudaSuccess) {
        ggb_set_error("decoder phase table upload -> %s", cudaGetErrorString(e));
        cudaFree(D->scratch);
        delete D;
        return GGB_ERR_CUDA;
    }
    P.phases = D->d_phases;
    P.n_phases = (int)D->phases.size();
    D->smem = sizeof(SmemLayout);
    const void *fn = type == GGB_TYPE_Q4_0 ? (const void *)decoder_kernel<GGB_TYPE_Q4_0>
                                           : (const void *)decoder_kernel<GGB_TYPE_Q8_0>;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D->smem);
    int occ = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kThreads, D->smem);
    if (e != cudaSuccess || occ < 1) {
        ggb_set_error("decoder kernel cannot be resident (smem %zu B): %s", D->smem,
                      e != cudaSuccess ? cudaGetErrorString(e) : "occupancy 0");
        cudaFree(D->d_phases);
        cudaFree(D->scratch);
        delete D;
        return GGB_ERR_CUDA;
    }
    GGB_CUDA(cudaStreamSynchronize(ggb_cuda_stream()));
    *out = D;
    return GGB_OK;
}

int ggb_decoder_step(ggb_decoder *D) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(D != nullptr, "null decoder");
    void *args[] = {(void *)&D->P};
    const void *fn = D->P.type == GGB_TYPE_Q4_0 ? (const void *)decoder_kernel<GGB_TYPE_Q4_0>
                                                : (const void *)decoder_kernel<GGB_TYPE_Q8_0>;
    GGB_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)D->grid), dim3(kThreads), args, D->smem, ggb_cuda_stream()));
    ggb_count_launch();
    return GGB_OK;
}

// one traced step: per phase and CTA four SM-clock stamps (start, after prologue, after tiles, after
// the grid barrier) plus the phase kinds; the cudaEvent/ncu numbers say how long, this says where.
int ggb_decoder_trace(ggb_decoder *D, long long *stamps, int32_t *kinds, int max_phases, int *n_phases, int *grid) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(D && stamps && kinds && n_phases && grid, "null pointer");
    const int np = (int)D->phases.size();
    GGB_CHECK_ARG(max_phases >= np, "trace buffer holds %d phases, need %d", max_phases, np);
    // stamps: [np][grid][4] phase stamps followed by [grid][32][3] LOGITS tile stamps
    const size_t bytes1 = (size_t)np * D->grid * 4 * sizeof(long long);
    const size_t bytes2 = (size_t)D->grid * 32 * 3 * sizeof(long long);
    const size_t bytes = bytes1 + bytes2 + (size_t)D->grid * 8 * sizeof(long long);
    GGB_CHECK_ARG((size_t)max_phases * D->grid * 4 * sizeof(long long) >= bytes, "trace buffer too small");
    long long *dev = nullptr;
    GGB_CUDA(cudaMalloc(&dev, bytes));
    GGB_CUDA(cudaMemsetAsync(dev, 0, bytes, ggb_cuda_stream()));
    D->P.trace = dev;
    D->P.trace2 = dev + bytes1 / sizeof(long long);
    int rc = ggb_decoder_step(D);
    D->P.trace = nullptr;
    D->P.trace2 = nullptr;
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

int ggb_decoder_free(ggb_decoder *D) {
    if (!D) return GGB_OK;
    cudaStreamSynchronize(ggb_cuda_stream());
    cudaFree(D->d_phases);
    cudaFree(D->scratch);
    delete D;
    return GGB_OK;
}

} // extern "C"
