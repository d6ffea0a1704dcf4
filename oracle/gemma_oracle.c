/*
 * gemma_oracle.c — CPU ORACLE, model level (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 * See gemma_oracle.h.  Follows ref: src/gemma_model.cpp (graph), src/hpc.cpp (mul_mat
 * dispatch), src/thread_pool.cpp (worker pool).
 */
#define _GNU_SOURCE
#include "gemma_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------ thread pool
 * ref: src/thread_pool.{h,cpp}: mutex + condvar task queue, enqueue -> future.  Here the
 * "future" is a completion counter the dispatcher waits on (future.get() loop,
 * ref: src/hpc.cpp:271-273). */
typedef struct task {
    void (*fn)(void *);
    void *arg;
    struct task *next;
} task;

typedef struct {
    pthread_t *workers;
    int n;
    task *head, *tail;
    pthread_mutex_t mu;
    pthread_cond_t cv, done_cv;
    int pending, stop;
} pool;

static void *pool_worker(void *p_) {
    pool *p = (pool *)p_;
    for (;;) {
        pthread_mutex_lock(&p->mu);
        while (!p->stop && !p->head) pthread_cond_wait(&p->cv, &p->mu);
        if (p->stop && !p->head) {
            pthread_mutex_unlock(&p->mu);
            return NULL;
        }
        task *t = p->head;
        p->head = t->next;
        if (!p->head) p->tail = NULL;
        pthread_mutex_unlock(&p->mu);
        t->fn(t->arg);
        free(t);
        pthread_mutex_lock(&p->mu);
        if (--p->pending == 0) pthread_cond_broadcast(&p->done_cv);
        pthread_mutex_unlock(&p->mu);
    }
}

static pool *pool_create(int n) {
    pool *p = (pool *)calloc(1, sizeof(pool));
    p->n = n;
    pthread_mutex_init(&p->mu, NULL);
    pthread_cond_init(&p->cv, NULL);
    pthread_cond_init(&p->done_cv, NULL);
    p->workers = (pthread_t *)calloc((size_t)n, sizeof(pthread_t));
    for (int i = 0; i < n; i++) pthread_create(&p->workers[i], NULL, pool_worker, p);
    return p;
}

static void pool_enqueue(pool *p, void (*fn)(void *), void *arg) {
    task *t = (task *)malloc(sizeof(task));
    t->fn = fn;
    t->arg = arg;
    t->next = NULL;
    pthread_mutex_lock(&p->mu);
    if (p->tail) p->tail->next = t;
    else p->head = t;
    p->tail = t;
    p->pending++;
    pthread_mutex_unlock(&p->mu);
    pthread_cond_signal(&p->cv);
}

static void pool_wait(pool *p) {
    pthread_mutex_lock(&p->mu);
    while (p->pending) pthread_cond_wait(&p->done_cv, &p->mu);
    pthread_mutex_unlock(&p->mu);
}

static void pool_destroy(pool *p) {
    pthread_mutex_lock(&p->mu);
    p->stop = 1;
    pthread_mutex_unlock(&p->mu);
    pthread_cond_broadcast(&p->cv);
    for (int i = 0; i < p->n; i++) pthread_join(p->workers[i], NULL);
    free(p->workers);
    free(p);
}

/* ------------------------------------------------------------------ model */

typedef struct {
    void *data;
    int64_t rows, cols, nbytes;
    int type;
} otensor;

struct orc_model {
    orc_config cfg;
    pool *tp;
    int n_threads;
    int n_tensors;
    otensor *t;
    uint16_t **kc, **vc; /* per layer [n_ctx][Hkv*d] fp16 */
    /* captures */
    float *cap_l_out, *cap_kqv, *cap_sa, *cap_ffn_act, *cap_norm, *cap_logits;
};

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

typedef struct {
    orc_model *m;
    otensor *t;
    int tid;
    int64_t r0, r1;
} gen_job;

static void gen_rows(void *a) {
    gen_job *j = (gen_job *)a;
    otensor *t = j->t;
    float *row = (float *)malloc(sizeof(float) * (size_t)t->cols);
    const size_t rs = orc_row_size(t->type, t->cols);
    for (int64_t r = j->r0; r < j->r1; r++) {
        const float sigma = j->tid == ORC_T_TOKEN_EMBD ? j->m->cfg.sigma_embd : j->m->cfg.sigma_mat;
        if (t->type == ORC_TYPE_Q4_K || t->type == ORC_TYPE_Q6_K) {
            /* K-quant weights: every bit pattern is a valid block, so the row is counter-based random bytes with the
             * fp16 super-scales set to sigma-sized values (no K-quant quantiser is restated: the search-based
             * make_qkx2_quants of ggml is not on the inference path) */
            unsigned char *b = (unsigned char *)t->data + r * rs;
            const uint64_t key = j->m->cfg.seed * 0x9E3779B97F4A7C15ull + (uint64_t)j->tid * 0xD1B54A32D192ED03ull + (uint64_t)r;
            for (size_t i = 0; i < rs; i += 8) {
                uint64_t z = key + (uint64_t)i * 0xBF58476D1CE4E5B9ull;
                z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
                z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
                z ^= z >> 31;
                memcpy(b + i, &z, rs - i < 8 ? rs - i : 8);
            }
            const int64_t nb = t->cols / ORC_QK_K;
            for (int64_t i = 0; i < nb; i++) {
                if (t->type == ORC_TYPE_Q4_K) {
                    orc_block_q4_K *blk = (orc_block_q4_K *)b + i;
                    const float d = sigma * (1.0f / 256.0f) * (float)(1 + (blk->d & 7));
                    blk->d = orc_fp32_to_fp16(d);
                    blk->dmin = orc_fp32_to_fp16(7.5f * d); /* E[d sc q] = E[dmin m]: zero-mean weights */
                } else {
                    orc_block_q6_K *blk = (orc_block_q6_K *)b + i;
                    blk->d = orc_fp32_to_fp16(sigma * (1.0f / 1024.0f) * (float)(1 + (blk->d & 7)));
                }
            }
            continue;
        }
        orc_synth_row(j->m->cfg.seed, (uint64_t)j->tid, r, t->cols, sigma, 0.0f, row);
        if (t->type == ORC_TYPE_Q4_0) orc_quantize_row_q4_0(row, (orc_block_q4_0 *)((char *)t->data + r * rs), t->cols);
        else orc_quantize_row_q8_0(row, (orc_block_q8_0 *)((char *)t->data + r * rs), t->cols);
    }
    free(row);
    free(j);
}

static void make_matrix(orc_model *m, int tid, int64_t rows, int64_t cols) {
    otensor *t = &m->t[tid];
    t->rows = rows;
    t->cols = cols;
    /* wtype Q4_K = the mix of a q4_k_m file in miniature: matrices Q4_K, the (tied) embedding / output matrix Q6_K */
    t->type = (m->cfg.wtype == ORC_TYPE_Q4_K && tid == ORC_T_TOKEN_EMBD) ? ORC_TYPE_Q6_K : m->cfg.wtype;
    t->nbytes = (int64_t)(orc_row_size(t->type, cols) * (size_t)rows);
    t->data = malloc((size_t)t->nbytes);
    const int64_t chunk = (rows + m->n_threads * 4 - 1) / (m->n_threads * 4);
    for (int64_t r0 = 0; r0 < rows; r0 += chunk) {
        gen_job *j = (gen_job *)malloc(sizeof(gen_job));
        j->m = m; j->t = t; j->tid = tid; j->r0 = r0; j->r1 = r0 + chunk < rows ? r0 + chunk : rows;
        pool_enqueue(m->tp, gen_rows, j);
    }
}

static void make_norm(orc_model *m, int tid, int64_t n) {
    otensor *t = &m->t[tid];
    t->rows = 1;
    t->cols = n;
    t->type = ORC_TYPE_F32;
    t->nbytes = n * 4;
    t->data = malloc((size_t)t->nbytes);
    orc_synth_row(m->cfg.seed, (uint64_t)tid, 0, n, ORC_SYNTH_SIGMA, 1.0f, (float *)t->data);
}

orc_model *orc_model_create(const orc_config *cfg, int n_threads) {
    orc_model *m = (orc_model *)calloc(1, sizeof(orc_model));
    m->cfg = *cfg;
    if (n_threads < 1) n_threads = 1;
    m->n_threads = n_threads;
    m->tp = pool_create(n_threads);
    const int L = cfg->n_layer;
    const int64_t E = cfg->n_embd, F = cfg->n_ff, V = cfg->n_vocab;
    const int64_t QD = (int64_t)cfg->n_head * cfg->head_dim, KD = (int64_t)cfg->n_head_kv * cfg->head_dim;
    m->n_tensors = ORC_T_LAYER0 + ORC_T_PER_LAYER * L;
    m->t = (otensor *)calloc((size_t)m->n_tensors, sizeof(otensor));
    make_matrix(m, ORC_T_TOKEN_EMBD, V, E);
    make_norm(m, ORC_T_OUTPUT_NORM, E);
    for (int il = 0; il < L; il++) {
        const int b = ORC_T_LAYER0 + ORC_T_PER_LAYER * il;
        make_norm(m, b + ORC_L_ATTN_NORM, E);
        make_matrix(m, b + ORC_L_ATTN_Q, QD, E);
        make_matrix(m, b + ORC_L_ATTN_K, KD, E);
        make_matrix(m, b + ORC_L_ATTN_V, KD, E);
        make_matrix(m, b + ORC_L_ATTN_OUT, E, QD);
        make_norm(m, b + ORC_L_FFN_NORM, E);
        make_matrix(m, b + ORC_L_FFN_GATE, F, E);
        make_matrix(m, b + ORC_L_FFN_UP, F, E);
        make_matrix(m, b + ORC_L_FFN_DOWN, E, F);
    }
    pool_wait(m->tp);
    m->kc = (uint16_t **)calloc((size_t)L, sizeof(uint16_t *));
    m->vc = (uint16_t **)calloc((size_t)L, sizeof(uint16_t *));
    for (int il = 0; il < L; il++) {
        m->kc[il] = (uint16_t *)calloc((size_t)cfg->n_ctx * (size_t)KD, 2);
        m->vc[il] = (uint16_t *)calloc((size_t)cfg->n_ctx * (size_t)KD, 2);
    }
    m->cap_l_out = (float *)calloc((size_t)L * (size_t)E, 4);
    m->cap_kqv = (float *)calloc((size_t)L * (size_t)QD, 4);
    m->cap_sa = (float *)calloc((size_t)L * (size_t)E, 4);
    m->cap_ffn_act = (float *)calloc((size_t)L * (size_t)F, 4);
    m->cap_norm = (float *)calloc((size_t)E, 4);
    m->cap_logits = (float *)calloc((size_t)V, 4);
    return m;
}

void orc_model_free(orc_model *m) {
    if (!m) return;
    for (int i = 0; i < m->n_tensors; i++) free(m->t[i].data);
    for (int il = 0; il < m->cfg.n_layer; il++) {
        free(m->kc[il]);
        free(m->vc[il]);
    }
    free(m->kc); free(m->vc); free(m->t);
    free(m->cap_l_out); free(m->cap_kqv); free(m->cap_sa); free(m->cap_ffn_act); free(m->cap_norm); free(m->cap_logits);
    pool_destroy(m->tp);
    free(m);
}

int orc_model_threads(const orc_model *m) { return m->n_threads; }

const void *orc_model_tensor(const orc_model *m, int tensor_id, int64_t *rows, int64_t *cols, int64_t *nbytes) {
    if (tensor_id < 0 || tensor_id >= m->n_tensors) return NULL;
    if (rows) *rows = m->t[tensor_id].rows;
    if (cols) *cols = m->t[tensor_id].cols;
    if (nbytes) *nbytes = m->t[tensor_id].nbytes;
    return m->t[tensor_id].data;
}

void orc_model_reset(orc_model *m) {
    const size_t n = (size_t)m->cfg.n_ctx * (size_t)m->cfg.n_head_kv * (size_t)m->cfg.head_dim * 2;
    for (int il = 0; il < m->cfg.n_layer; il++) {
        memset(m->kc[il], 0, n);
        memset(m->vc[il], 0, n);
    }
}

/* KV cache pre-filled with seeded ~N(0,1) values rounded to fp16 (bandwidth sweeps);
 * same generator and tensor ids (2*layer, 2*layer+1) as the device-side fill. */
void orc_model_fill_kv_synth(orc_model *m, uint64_t seed) {
    const int64_t n = (int64_t)m->cfg.n_ctx * m->cfg.n_head_kv * m->cfg.head_dim;
    for (int il = 0; il < m->cfg.n_layer; il++) {
        for (int64_t i = 0; i < n; i++) {
            m->kc[il][i] = orc_fp32_to_fp16(0.0f + orc_synth_value(seed, (uint64_t)(2 * il), (uint64_t)i, 1.0f));
            m->vc[il][i] = orc_fp32_to_fp16(0.0f + orc_synth_value(seed, (uint64_t)(2 * il + 1), (uint64_t)i, 1.0f));
        }
    }
}

uint16_t *orc_model_kcache(orc_model *m, int layer) { return m->kc[layer]; }
uint16_t *orc_model_vcache(orc_model *m, int layer) { return m->vc[layer]; }

/* ----------------------------------------------------- hpc.cpp mul_mat */

typedef struct {
    int64_t start_row, end_row, shared_edge, col_num, ne01, ne1, nb01, nb1, nb2, row_size;
    const char *src0, *src1;
    char *dst;
    int type, order;
} mm_job;

static void mm_run(void *a) {
    mm_job *j = (mm_job *)a;
    orc_mul_mat_sub(j->start_row, j->end_row, j->shared_edge, j->col_num, j->ne01, j->ne1, j->nb01, j->nb1,
                    j->nb2, j->row_size, j->src0, j->src1, j->dst, j->type, j->order);
    free(j);
}

/* ref: src/hpc.cpp:216-277 with CPU_SPLIT_RATIO = 1.0 (the reference has no Q4_0/Q8_0 device
 * kernel, src/hpc.cpp:132-143,162-166) and N_THREADS_MUL_MAT_CPU = pool size:
 * cpu_row_num = ceil(ne01*ratio / N) * N; row_per_core = cpu_row_num / N; thread i gets
 * [i*row_per_core, (i+1)*row_per_core), clipped at ne01 inside mul_mat_sub (:34). */
void orc_hpc_mul_mat(orc_model *m, int64_t ne01, int64_t ne11, int64_t ne12, int64_t nb01, int64_t ne1,
                     int64_t nb1, int64_t nb2, size_t row_size, int64_t shared_edge, const void *src0,
                     void *dst, int src0_type, const char *wdata, int order) {
    const int N = m->n_threads;
    const int64_t col_num = ne11 * ne12;
    const int64_t cpu_row_num = ((int64_t)ceil(((double)ne01 * 1.0) / (double)N)) * N;
    const int64_t row_per_core = cpu_row_num / N;
    for (int i = 0; i < N; i++) {
        mm_job *j = (mm_job *)malloc(sizeof(mm_job));
        j->start_row = i * row_per_core;
        j->end_row = (i + 1) * row_per_core;
        j->shared_edge = shared_edge; j->col_num = col_num; j->ne01 = ne01; j->ne1 = ne1;
        j->nb01 = nb01; j->nb1 = nb1; j->nb2 = nb2; j->row_size = (int64_t)row_size;
        j->src0 = (const char *)src0; j->src1 = wdata; j->dst = (char *)dst;
        j->type = src0_type; j->order = order;
        pool_enqueue(m->tp, mm_run, j);
    }
    pool_wait(m->tp);
}

/* ggml_mul_mat node: INIT stage quantises the T activation rows to Q8_0 ([UPSTREAM]
 * ggml_compute_forward_mul_mat, "mmstage_init" ref: notes.md:91), COMPUTE = hpc mul_mat.
 * x: [T][K] f32, out: [T][N] f32. */
static void node_mul_mat(orc_model *m, const otensor *w, const float *x, int T, float *out, int order) {
    const int64_t K = w->cols, N = w->rows;
    /* vec_dot_type of the weight type: Q8_0 for Q4_0 / Q8_0, Q8_K for the K-quants */
    const int kq = w->type == ORC_TYPE_Q4_K || w->type == ORC_TYPE_Q6_K;
    const size_t rs = orc_row_size(kq ? ORC_TYPE_Q8_K : ORC_TYPE_Q8_0, K);
    char *wdata = (char *)calloc(rs, (size_t)T);
    for (int t = 0; t < T; t++) {
        if (kq) orc_quantize_row_q8_K(x + (size_t)t * K, (orc_block_q8_K *)(wdata + rs * t), K);
        else orc_quantize_row_q8_0(x + (size_t)t * K, (orc_block_q8_0 *)(wdata + rs * t), K);
    }
    orc_hpc_mul_mat(m, N, T, 1, (int64_t)orc_row_size(w->type, K), T, N * 4, N * 4 * T, rs, K, w->data, out,
                    w->type, wdata, order);
    free(wdata);
}

static void node_norm(const orc_model *m, const float *x, const float *w, int T, int64_t E, float *y, int order) {
    (void)order;
    for (int t = 0; t < T; t++) {
        orc_rms_norm(E, x + (size_t)t * E, m->cfg.eps, y + (size_t)t * E);
        orc_mul(E, y + (size_t)t * E, w, y + (size_t)t * E);
    }
}

typedef struct {
    orc_model *m;
    int il, order, t0, t1, pos0;
    const float *q;
    float *out;
} attn_job;

static void attn_run(void *a) {
    attn_job *j = (attn_job *)a;
    const orc_config *c = &j->m->cfg;
    const int64_t QD = (int64_t)c->n_head * c->head_dim, KD = (int64_t)c->n_head_kv * c->head_dim;
    for (int t = j->t0; t < j->t1; t++) {
        const int64_t n_pos = j->pos0 + t + 1;
        orc_attn_decode_ref(c->n_head, c->n_head_kv, c->head_dim, n_pos, c->n_ctx, j->q + (size_t)t * QD,
                                j->m->kc[j->il], j->m->vc[j->il], KD, j->out + (size_t)t * QD);
    }
    free(j);
}

/* ref: src/gemma_model.cpp:665-747 (build_compute_graph) executed for T tokens. */
int orc_model_forward(orc_model *m, const int32_t *tokens, int T, int pos0, int order, int logits_all,
                      float *logits_last) {
    const orc_config *c = &m->cfg;
    const int64_t E = c->n_embd, F = c->n_ff, V = c->n_vocab, d = c->head_dim;
    const int64_t QD = (int64_t)c->n_head * d, KD = (int64_t)c->n_head_kv * d;
    if (T < 1 || pos0 < 0 || pos0 + T > c->n_ctx) return -1;
    float *x = (float *)malloc(sizeof(float) * (size_t)T * E);
    float *cur = (float *)malloc(sizeof(float) * (size_t)T * E);
    float *q = (float *)malloc(sizeof(float) * (size_t)T * QD);
    float *k = (float *)malloc(sizeof(float) * (size_t)T * KD);
    float *v = (float *)malloc(sizeof(float) * (size_t)T * KD);
    float *kqv = (float *)malloc(sizeof(float) * (size_t)T * QD);
    float *sa = (float *)malloc(sizeof(float) * (size_t)T * E);
    float *gate = (float *)malloc(sizeof(float) * (size_t)T * F);
    float *up = (float *)malloc(sizeof(float) * (size_t)T * F);
    float *tmp = (float *)malloc(sizeof(float) * (size_t)T * (E > QD ? E : QD));

    /* get_rows + scale sqrt(n_embd)  (:675-679) */
    const otensor *emb = &m->t[ORC_T_TOKEN_EMBD];
    const size_t emb_rs = orc_row_size(emb->type, E);
    const float emb_scale = sqrtf((float)c->n_embd);
    for (int t = 0; t < T; t++) {
        if (tokens[t] < 0 || tokens[t] >= V) return -2;
        orc_dequantize_row(emb->type, (const char *)emb->data + emb_rs * (size_t)tokens[t], x + (size_t)t * E, E);
        orc_scale(E, x + (size_t)t * E, emb_scale, x + (size_t)t * E);
    }
    const float q_scale = 1.0f / sqrtf((float)d);

    for (int il = 0; il < c->n_layer; il++) {
        const otensor *lt = &m->t[ORC_T_LAYER0 + ORC_T_PER_LAYER * il];
        node_norm(m, x, (const float *)lt[ORC_L_ATTN_NORM].data, T, E, cur, order); /* :690 */
        node_mul_mat(m, &lt[ORC_L_ATTN_Q], cur, T, q, order);                        /* :692 */
        node_mul_mat(m, &lt[ORC_L_ATTN_K], cur, T, k, order);                        /* :694 */
        node_mul_mat(m, &lt[ORC_L_ATTN_V], cur, T, v, order);                        /* :696 */
        for (int t = 0; t < T; t++) {
            const int32_t pos = pos0 + t;
            orc_rope_neox((int)d, c->n_head, pos, c->rope_base, q + (size_t)t * QD, q + (size_t)t * QD); /* :698 */
            orc_scale(QD, q + (size_t)t * QD, q_scale, q + (size_t)t * QD);                              /* :708 */
            orc_rope_neox((int)d, c->n_head_kv, pos, c->rope_base, k + (size_t)t * KD, k + (size_t)t * KD); /* :710 */
            /* KV store, F32 -> F16 (:499-518) */
            for (int64_t i = 0; i < KD; i++) {
                m->kc[il][(size_t)pos * KD + i] = orc_fp32_to_fp16(k[(size_t)t * KD + i]);
                m->vc[il][(size_t)pos * KD + i] = orc_fp32_to_fp16(v[(size_t)t * KD + i]);
            }
        }
        /* kq -> soft_max_ext -> kqv (:454-489), tokens spread over the pool */
        {
            const int per = (T + m->n_threads - 1) / m->n_threads;
            for (int t0 = 0; t0 < T; t0 += per) {
                attn_job *j = (attn_job *)malloc(sizeof(attn_job));
                j->m = m; j->il = il; j->order = order; j->t0 = t0; j->t1 = t0 + per < T ? t0 + per : T;
                j->pos0 = pos0; j->q = q; j->out = kqv;
                pool_enqueue(m->tp, attn_run, j);
            }
            pool_wait(m->tp);
        }
        memcpy(m->cap_kqv + (size_t)il * QD, kqv + (size_t)(T - 1) * QD, sizeof(float) * (size_t)QD);
        node_mul_mat(m, &lt[ORC_L_ATTN_OUT], kqv, T, tmp, order); /* :493 */
        orc_add((int64_t)T * E, tmp, x, sa);                      /* :723 sa_out = cur + inpL */
        memcpy(m->cap_sa + (size_t)il * E, sa + (size_t)(T - 1) * E, sizeof(float) * (size_t)E);
        node_norm(m, sa, (const float *)lt[ORC_L_FFN_NORM].data, T, E, cur, order); /* :724 */
        node_mul_mat(m, &lt[ORC_L_FFN_UP], cur, T, up, order);                       /* :446 */
        node_mul_mat(m, &lt[ORC_L_FFN_GATE], cur, T, gate, order);                   /* :447 */
        orc_gelu((int64_t)T * F, gate, gate);                                        /* :448 */
        orc_mul((int64_t)T * F, gate, up, gate);                                     /* :449 */
        memcpy(m->cap_ffn_act + (size_t)il * F, gate + (size_t)(T - 1) * F, sizeof(float) * (size_t)F);
        node_mul_mat(m, &lt[ORC_L_FFN_DOWN], gate, T, tmp, order);                   /* :450 */
        orc_add((int64_t)T * E, tmp, sa, x);                                         /* :731 */
        memcpy(m->cap_l_out + (size_t)il * E, x + (size_t)(T - 1) * E, sizeof(float) * (size_t)E);
    }
    node_norm(m, x, (const float *)m->t[ORC_T_OUTPUT_NORM].data, T, E, cur, order); /* :738 */
    memcpy(m->cap_norm, cur + (size_t)(T - 1) * E, sizeof(float) * (size_t)E);
    if (logits_all && T > 1) { /* :740 computes every token's logits */
        float *all = (float *)malloc(sizeof(float) * (size_t)T * V);
        node_mul_mat(m, emb, cur, T, all, order);
        memcpy(m->cap_logits, all + (size_t)(T - 1) * V, sizeof(float) * (size_t)V);
        free(all);
    } else {
        node_mul_mat(m, emb, cur + (size_t)(T - 1) * E, 1, m->cap_logits, order);
    }
    if (logits_last) memcpy(logits_last, m->cap_logits, sizeof(float) * (size_t)V);
    free(x); free(cur); free(q); free(k); free(v); free(kqv); free(sa); free(gate); free(up); free(tmp);
    return 0;
}

int orc_model_generate(orc_model *m, const int32_t *prompt, int n_prompt, int n_decode, int order,
                       int32_t *out_tokens, double *prefill_s, double *decode_s) {
    orc_model_reset(m);
    double t0 = now_s();
    int rc = orc_model_forward(m, prompt, n_prompt, 0, order, 1, NULL);
    if (rc) return rc;
    int32_t tok = orc_argmax(m->cfg.n_vocab, m->cap_logits);
    double t1 = now_s();
    if (prefill_s) *prefill_s = t1 - t0;
    int n = 0;
    if (n_decode > 0) out_tokens[n++] = tok;
    while (n < n_decode) {
        rc = orc_model_forward(m, &tok, 1, n_prompt + n - 1, order, 0, NULL);
        if (rc) return rc;
        tok = orc_argmax(m->cfg.n_vocab, m->cap_logits);
        out_tokens[n++] = tok;
    }
    if (decode_s) *decode_s = now_s() - t1;
    return 0;
}

const float *orc_model_capture(const orc_model *m, const char *name, int layer, int64_t *n) {
    const orc_config *c = &m->cfg;
    const int64_t E = c->n_embd, QD = (int64_t)c->n_head * c->head_dim;
    if (!strcmp(name, "l_out")) { if (n) *n = E; return m->cap_l_out + (size_t)layer * E; }
    if (!strcmp(name, "sa_out")) { if (n) *n = E; return m->cap_sa + (size_t)layer * E; }
    if (!strcmp(name, "kqv_out")) { if (n) *n = QD; return m->cap_kqv + (size_t)layer * QD; }
    if (!strcmp(name, "ffn_act")) { if (n) *n = c->n_ff; return m->cap_ffn_act + (size_t)layer * c->n_ff; }
    if (!strcmp(name, "result_norm")) { if (n) *n = E; return m->cap_norm; }
    if (!strcmp(name, "result_output")) { if (n) *n = c->n_vocab; return m->cap_logits; }
    return NULL;
}
