/* ORACLE -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see stark.h): restates the winterfell v0.3 prover and
 * verifier from the published protocol; every [RECALLED] detail is marked.  Plain C + OpenMP, base field only. */
#include "stark.h"
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
void stark_free(void *p) { free(p); }
/* representation changes for test drivers: canonical <-> Montgomery over arrays */
void fe_array_to_mont(const uint64_t *in, fe *out, size_t n) { for (size_t i = 0; i < n; i++) out[i] = fe_from_u64(in[i]); }
void fe_array_from_mont(const fe *in, uint64_t *out, size_t n) { for (size_t i = 0; i < n; i++) out[i] = fe_to_u64(in[i]); }

/* ------------------------------------------------------------------ byte buffers (winterfell ByteWriter: little endian) */
typedef struct { uint8_t *p; size_t len, cap; } buf_t;
static void buf_put(buf_t *b, const void *src, size_t n) {
    if (b->len + n > b->cap) { b->cap = (b->len + n) * 2 + 64; b->p = realloc(b->p, b->cap); }
    memcpy(b->p + b->len, src, n); b->len += n;
}
static void buf_u8(buf_t *b, uint8_t v) { buf_put(b, &v, 1); }
static void buf_u16(buf_t *b, uint16_t v) { uint8_t t[2] = {(uint8_t)v, (uint8_t)(v >> 8)}; buf_put(b, t, 2); }
static void buf_u32(buf_t *b, uint32_t v) { uint8_t t[4]; for (int i = 0; i < 4; i++) t[i] = (uint8_t)(v >> (8 * i)); buf_put(b, t, 4); }
static void buf_u64(buf_t *b, uint64_t v) { uint8_t t[8]; for (int i = 0; i < 8; i++) t[i] = (uint8_t)(v >> (8 * i)); buf_put(b, t, 8); }
static void buf_fe(buf_t *b, fe v) { buf_u64(b, fe_to_u64(v)); } /* BaseElement::write_into = canonical LE bytes */

/* ------------------------------------------------------------------ NTT (any correct transform gives identical field values) */
static void bit_reverse(fe *a, size_t n) {
    unsigned l = ilog2(n);
    for (size_t i = 0; i < n; i++) {
        size_t j = 0;
        for (unsigned b = 0; b < l; b++) j |= ((i >> b) & 1) << (l - 1 - b);
        if (j > i) { fe t = a[i]; a[i] = a[j]; a[j] = t; }
    }
}
void ntt_natural(fe *a, size_t n, int inverse) {
    if (n == 1) return;
    unsigned l = ilog2(n);
    fe w = fe_root_of_unity(l);
    if (inverse) w = fe_inv(w);
    fe *tw = malloc((n / 2) * sizeof(fe));
    tw[0] = FE_ONE;
    for (size_t i = 1; i < n / 2; i++) tw[i] = fe_mul(tw[i - 1], w);
    bit_reverse(a, n);
    for (size_t h = 1; h < n; h <<= 1) {
        size_t step = n / (2 * h);
        for (size_t s = 0; s < n; s += 2 * h)
            for (size_t k = 0; k < h; k++) {
                fe u = a[s + k], v = fe_mul(a[s + k + h], tw[k * step]);
                a[s + k] = fe_add(u, v); a[s + k + h] = fe_sub(u, v);
            }
    }
    if (inverse) { fe ninv = fe_inv(fe_from_u64(n)); for (size_t i = 0; i < n; i++) a[i] = fe_mul(a[i], ninv); }
    free(tw);
}
/* coefficients (n) -> evaluations over offset * <w_{n*blowup}>, natural order (winterfell fft::evaluate_poly_with_offset) */
static void eval_with_offset(const fe *coeffs, size_t n, size_t blowup, fe *out) {
    size_t m = n * blowup;
    fe off = fe_from_u64(F63_GENERATOR), s = FE_ONE;
    for (size_t i = 0; i < n; i++) { out[i] = fe_mul(coeffs[i], s); s = fe_mul(s, off); }
    memset(out + n, 0, (m - n) * sizeof(fe));
    ntt_natural(out, m, 0);
}
void lde_column(const fe *evals, size_t n, size_t blowup, fe *out) {
    fe *c = malloc(n * sizeof(fe));
    memcpy(c, evals, n * sizeof(fe));
    ntt_natural(c, n, 1);
    eval_with_offset(c, n, blowup, out);
    free(c);
}
static fe poly_eval(const fe *c, size_t n, fe x) { fe r = 0; for (size_t i = n; i-- > 0;) r = fe_add(fe_mul(r, x), c[i]); return r; }

/* ------------------------------------------------------------------ hashing of elements, Merkle trees */
void hash_elements(int hash_fn, const fe *e, size_t n, uint8_t out[32]) {
    uint8_t stackbuf[1024] = {0}, *b = n * 8 <= sizeof stackbuf ? stackbuf : malloc(n * 8);
    for (size_t i = 0; i < n; i++) { uint64_t v = fe_to_u64(e[i]); memcpy(b + 8 * i, &v, 8); }
    hash_bytes(hash_fn, b, n * 8, out);
    if (b != stackbuf) free(b);
}
/* nodes[1] is the root, nodes[i] = H(nodes[2i] || nodes[2i+1]), leaves at nodes[nleaves + j]  (winterfell MerkleTree::new) */
void merkle_build(int hash_fn, const uint8_t *leaves, size_t nl, uint8_t *nodes) {
    memset(nodes, 0, 32);
    memcpy(nodes + nl * 32, leaves, nl * 32);
    for (size_t lvl = nl / 2; lvl >= 1; lvl /= 2) {
#pragma omp parallel for schedule(static) if (lvl >= 1024)
        for (size_t i = lvl; i < 2 * lvl; i++) hash_bytes(hash_fn, nodes + 2 * i * 32, 64, nodes + i * 32);
    }
}
static int cmp_size(const void *a, const void *b) { size_t x = *(const size_t *)a, y = *(const size_t *)b; return x < y ? -1 : x > y; }
/* winterfell MerkleTree::prove_batch + BatchMerkleProof::serialize_nodes [RECALLED]; returns the byte length written */
size_t merkle_prove_batch(const uint8_t *nodes, size_t nl, const size_t *positions, size_t npos, uint8_t *out) {
    /* normalize: even member of each leaf pair, sorted, unique */
    size_t *idx = malloc(npos * sizeof(size_t)), cnt = 0;
    for (size_t i = 0; i < npos; i++) idx[i] = positions[i] & ~(size_t)1;
    qsort(idx, npos, sizeof(size_t), cmp_size);
    for (size_t i = 0; i < npos; i++) if (i == 0 || idx[i] != idx[i - 1]) idx[cnt++] = idx[i];
    unsigned depth = ilog2(nl);
    /* per-slot node vectors; a slot can receive at most depth+1 digests */
    uint8_t *vec = calloc(cnt * (depth + 1), 32);
    uint8_t *vlen = calloc(cnt, 1);
    size_t *cur = malloc(cnt * sizeof(size_t)), *nxt = malloc(cnt * sizeof(size_t)), ncur = 0;
    for (size_t i = 0; i < cnt; i++) {
        for (size_t j = idx[i]; j < idx[i] + 2; j++) {
            int queried = 0;
            for (size_t k = 0; k < npos; k++) if (positions[k] == j) { queried = 1; break; }
            if (!queried) { memcpy(vec + (i * (depth + 1) + vlen[i]) * 32, nodes + (nl + j) * 32, 32); vlen[i]++; }
        }
        cur[ncur++] = (idx[i] + nl) >> 1;
    }
    for (unsigned d = 1; d < depth; d++) {
        size_t nn = 0, i = 0;
        while (i < ncur) {
            size_t sib = cur[i] ^ 1;
            if (i + 1 < ncur && cur[i + 1] == sib) i++;
            else { memcpy(vec + (i * (depth + 1) + vlen[i]) * 32, nodes + sib * 32, 32); vlen[i]++; }
            nxt[nn++] = sib >> 1;
            i++;
        }
        memcpy(cur, nxt, nn * sizeof(size_t)); ncur = nn;
    }
    size_t o = 0;
    out[o++] = (uint8_t)cnt;
    for (size_t i = 0; i < cnt; i++) {
        out[o++] = vlen[i];
        memcpy(out + o, vec + i * (depth + 1) * 32, (size_t)vlen[i] * 32); o += (size_t)vlen[i] * 32;
    }
    free(idx); free(vec); free(vlen); free(cur); free(nxt);
    return o;
}
/* winterfell BatchMerkleProof::get_root [RECALLED]; leaves[k] is the digest at positions[k]. returns 0 on success */
static int merkle_batch_root(int hash_fn, const uint8_t *paths, size_t plen, const size_t *positions, const uint8_t *leaves, size_t npos,
                             unsigned depth, uint8_t root[32]) {
    size_t nl = (size_t)1 << depth;
    size_t *idx = malloc(npos * sizeof(size_t)), cnt = 0;
    for (size_t i = 0; i < npos; i++) {
        for (size_t k = 0; k < i; k++) if (positions[k] == positions[i]) { free(idx); return 1; }
        if (positions[i] >= nl) { free(idx); return 1; }
        idx[i] = positions[i] & ~(size_t)1;
    }
    qsort(idx, npos, sizeof(size_t), cmp_size);
    for (size_t i = 0; i < npos; i++) if (i == 0 || idx[i] != idx[i - 1]) idx[cnt++] = idx[i];
    /* parse node vectors */
    if (plen < 1 || paths[0] != cnt) { free(idx); return 2; }
    const uint8_t **vptr = malloc(cnt * sizeof(*vptr));
    uint8_t *vlen = malloc(cnt), *ptr = calloc(cnt, 1);
    size_t o = 1;
    for (size_t i = 0; i < cnt; i++) {
        if (o >= plen) goto bad;
        vlen[i] = paths[o++]; vptr[i] = paths + o; o += (size_t)vlen[i] * 32;
        if (o > plen) goto bad;
    }
    if (o != plen) goto bad;
    /* node values of the current level, parallel to cur[] */
    size_t *cur = malloc(cnt * sizeof(size_t)), *nxt = malloc(cnt * sizeof(size_t)), ncur = 0;
    uint8_t *val = malloc(cnt * 32), *nval = malloc(cnt * 32);
    for (size_t i = 0; i < cnt; i++) {
        uint8_t pair[64];
        for (size_t j = 0; j < 2; j++) {
            const uint8_t *src = NULL;
            for (size_t k = 0; k < npos; k++) if (positions[k] == idx[i] + j) { src = leaves + k * 32; break; }
            if (!src) { if (ptr[i] >= vlen[i]) goto bad2; src = vptr[i] + (size_t)ptr[i] * 32; ptr[i]++; }
            memcpy(pair + 32 * j, src, 32);
        }
        hash_bytes(hash_fn, pair, 64, val + ncur * 32);
        cur[ncur++] = (nl + idx[i]) >> 1;
    }
    for (unsigned d = 1; d < depth; d++) {
        size_t nn = 0, i = 0;
        while (i < ncur) {
            size_t node = cur[i], sib = node ^ 1;
            uint8_t pair[64];
            const uint8_t *nv = val + i * 32, *sv;
            if (i + 1 < ncur && cur[i + 1] == sib) { sv = val + (i + 1) * 32; i++; }
            else { size_t slot = i; if (ptr[slot] >= vlen[slot]) goto bad2; sv = vptr[slot] + (size_t)ptr[slot] * 32; ptr[slot]++; }
            if (node & 1) { memcpy(pair, sv, 32); memcpy(pair + 32, nv, 32); } else { memcpy(pair, nv, 32); memcpy(pair + 32, sv, 32); }
            hash_bytes(hash_fn, pair, 64, nval + nn * 32);
            nxt[nn++] = node >> 1;
            i++;
        }
        memcpy(cur, nxt, nn * sizeof(size_t)); memcpy(val, nval, nn * 32); ncur = nn;
    }
    int ok = (ncur == 1 && cur[0] == 1);
    if (ok) memcpy(root, val, 32);
    free(cur); free(nxt); free(val); free(nval); free(idx); free(vptr); free(vlen); free(ptr);
    return ok ? 0 : 3;
bad2:
    free(cur); free(nxt); free(val); free(nval);
bad:
    free(idx); free(vptr); free(vlen); free(ptr);
    return 2;
}

/* ------------------------------------------------------------------ public coin (winterfell RandomCoin [RECALLED]) */
typedef struct { int hash_fn; uint8_t seed[32]; uint64_t counter; } coin_t;
static void coin_init(coin_t *c, int hash_fn, const uint8_t *bytes, size_t n) { c->hash_fn = hash_fn; hash_bytes(hash_fn, bytes, n, c->seed); c->counter = 0; }
static void coin_reseed(coin_t *c, const uint8_t d[32]) { uint8_t t[64]; memcpy(t, c->seed, 32); memcpy(t + 32, d, 32); hash_bytes(c->hash_fn, t, 64, c->seed); c->counter = 0; }
static void merge_with_int(int hash_fn, const uint8_t seed[32], uint64_t v, uint8_t out[32]) {
    uint8_t t[40]; memcpy(t, seed, 32); for (int i = 0; i < 8; i++) t[32 + i] = (uint8_t)(v >> (8 * i));
    hash_bytes(hash_fn, t, 40, out);
}
static void coin_reseed_int(coin_t *c, uint64_t v) { uint8_t t[32]; merge_with_int(c->hash_fn, c->seed, v, t); memcpy(c->seed, t, 32); c->counter = 0; }
static uint64_t coin_next_u64(coin_t *c) { uint8_t t[32]; c->counter++; merge_with_int(c->hash_fn, c->seed, c->counter, t); uint64_t v; memcpy(&v, t, 8); return v; }
/* rejection sampling of the first ELEMENT_BYTES of each output until it is a canonical element */
static int coin_draw(coin_t *c, fe *out) {
    for (int i = 0; i < 1000; i++) { uint64_t v = coin_next_u64(c); if (v < F63_P) { *out = fe_from_u64(v); return 0; } }
    return 1;
}
static int coin_draw_integers(coin_t *c, size_t k, size_t domain, size_t *out) {
    size_t got = 0;
    for (int i = 0; i < 1000 && got < k; i++) {
        size_t v = (size_t)(coin_next_u64(c) & (uint64_t)(domain - 1));
        int dup = 0;
        for (size_t j = 0; j < got; j++) if (out[j] == v) dup = 1;
        if (!dup) out[got++] = v;
    }
    return got == k ? 0 : 1;
}
static unsigned ctz64(uint64_t v) { return v ? (unsigned)__builtin_ctzll(v) : 64; }
static unsigned coin_check_leading_zeros(const coin_t *c, uint64_t v) { uint8_t t[32]; merge_with_int(c->hash_fn, c->seed, v, t); uint64_t h; memcpy(&h, t, 8); return ctz64(h); }
static unsigned coin_leading_zeros(const coin_t *c) { uint64_t h; memcpy(&h, c->seed, 8); return ctz64(h); }

/* ------------------------------------------------------------------ context / options */
static void write_context(buf_t *b, uint32_t width, size_t n, const stark_options *o) {
    buf_u8(b, (uint8_t)width); buf_u8(b, (uint8_t)ilog2(n));
    buf_u16(b, 0);                       /* trace meta: empty for TraceTable::new */
    buf_u8(b, 8); buf_u64(b, F63_P);     /* field modulus bytes */
    buf_u8(b, (uint8_t)o->num_queries); buf_u8(b, (uint8_t)ilog2(o->blowup_factor)); buf_u8(b, (uint8_t)o->grinding_factor);
    buf_u8(b, (uint8_t)o->hash_fn); buf_u8(b, (uint8_t)o->field_extension);
    buf_u8(b, (uint8_t)ilog2(o->fri_folding_factor)); buf_u8(b, (uint8_t)ilog2(o->fri_max_remainder_size));
}
#define CONTEXT_BYTES 20
static size_t num_fri_layers(const stark_options *o, size_t domain) { size_t r = 0; while (domain > o->fri_max_remainder_size) { domain /= o->fri_folding_factor; r++; } return r; }
static int options_ok(const stark_options *o) {
    return o->field_extension >= 1 && o->field_extension <= 3 && o->fri_folding_factor == 4 && o->num_queries > 0 && o->num_queries < 256 && o->blowup_factor >= 2 &&
           (o->blowup_factor & (o->blowup_factor - 1)) == 0 && (o->hash_fn == HASH_BLAKE3_256 || o->hash_fn == HASH_SHA3_256) &&
           o->fri_max_remainder_size >= 4 && (o->fri_max_remainder_size & (o->fri_max_remainder_size - 1)) == 0 && o->grinding_factor < 32;
}

/* ------------------------------------------------------------------ constraint bookkeeping shared by prover and verifier */
typedef struct { size_t eval_degree; uint64_t adj; } tgroup_t;
typedef struct {
    size_t num_steps; fe offset; uint64_t adj;       /* divisor x^num_steps - offset ; degree adjustment */
    size_t first_step, stride;
} bgroup_t;
typedef struct {
    const air_t *air; size_t n, ce, ce_n;
    fe *t_alpha, *t_beta; uint32_t *t_group; tgroup_t *tgroups; uint32_t ntg;
    uint32_t na; air_assertion *sa; fe **apoly; fe *axoff; fe *b_alpha, *b_beta; uint32_t *a_group; bgroup_t *bgroups; uint32_t nbg;
    fe g, g_inv_last; /* trace domain generator, g^(n-1) */
} cons_t;
static int cmp_assert(const void *a, const void *b) {
    const air_assertion *x = a, *y = b;
    if (x->stride != y->stride) return x->stride < y->stride ? -1 : 1;
    if (x->first_step != y->first_step) return x->first_step < y->first_step ? -1 : 1;
    return x->column < y->column ? -1 : x->column > y->column;
}
/* draws the composition coefficients in winterfell's order (Air::get_constraint_composition_coefficients) and builds
 * transition groups (Air::get_transition_constraints) and boundary groups (Air::get_boundary_constraints) [RECALLED] */
static int cons_init(cons_t *k, const air_t *a, coin_t *coin) {
    memset(k, 0, sizeof *k);
    k->air = a; k->n = a->trace_len; k->ce = air_ce_blowup(a); k->ce_n = k->n * k->ce;
    k->g = fe_root_of_unity(ilog2(k->n)); k->g_inv_last = fe_exp(k->g, k->n - 1);
    uint32_t nc = a->num_constraints;
    k->t_alpha = malloc(nc * sizeof(fe)); k->t_beta = malloc(nc * sizeof(fe)); k->t_group = malloc(nc * sizeof(uint32_t));
    for (uint32_t i = 0; coin && i < nc; i++) if (coin_draw(coin, &k->t_alpha[i]) || coin_draw(coin, &k->t_beta[i])) return 1;
    k->na = a->num_assertions;
    k->sa = malloc((k->na ? k->na : 1) * sizeof(air_assertion));
    memcpy(k->sa, a->assertions, k->na * sizeof(air_assertion));
    qsort(k->sa, k->na, sizeof(air_assertion), cmp_assert);
    k->b_alpha = malloc((k->na + 1) * sizeof(fe)); k->b_beta = malloc((k->na + 1) * sizeof(fe));
    for (uint32_t i = 0; coin && i < k->na; i++) if (coin_draw(coin, &k->b_alpha[i]) || coin_draw(coin, &k->b_beta[i])) return 1;
    /* transition groups keyed by evaluation degree, ascending */
    size_t comp_degree = k->ce_n - 1, target = comp_degree + (k->n - 1);
    k->tgroups = malloc(nc * sizeof(tgroup_t));
    for (uint32_t i = 0; i < nc; i++) {
        size_t ed = air_eval_degree(&a->degrees[i], k->n);
        uint32_t gi = 0;
        while (gi < k->ntg && k->tgroups[gi].eval_degree != ed) gi++;
        if (gi == k->ntg) { k->tgroups[gi].eval_degree = ed; k->tgroups[gi].adj = target - ed; k->ntg++; }
        k->t_group[i] = gi;
    }
    /* boundary groups keyed by (stride, first_step); sorted assertions make them contiguous */
    k->bgroups = malloc((k->na + 1) * sizeof(bgroup_t)); k->a_group = malloc((k->na + 1) * sizeof(uint32_t));
    k->apoly = calloc(k->na + 1, sizeof(fe *)); k->axoff = malloc((k->na + 1) * sizeof(fe));
    fe g_inv = fe_inv(k->g);
    for (uint32_t i = 0; i < k->na; i++) {
        air_assertion *s = &k->sa[i];
        if (i == 0 || s->stride != k->sa[i - 1].stride || s->first_step != k->sa[i - 1].first_step) {
            bgroup_t *bg = &k->bgroups[k->nbg++];
            bg->stride = s->stride; bg->first_step = s->first_step;
            bg->num_steps = s->stride == 0 ? 1 : k->n / s->stride;
            bg->offset = fe_exp(k->g, (uint64_t)bg->num_steps * s->first_step);
            bg->adj = comp_degree + bg->num_steps - (k->n - 1);
        }
        k->a_group[i] = k->nbg - 1;
        k->axoff[i] = FE_ONE;
        if (s->nvalues > 1) { /* value polynomial: interpolate, evaluate at x * g^-first_step */
            k->apoly[i] = malloc(s->nvalues * sizeof(fe));
            memcpy(k->apoly[i], s->values, s->nvalues * sizeof(fe));
            ntt_natural(k->apoly[i], s->nvalues, 1);
            if (s->first_step != 0) k->axoff[i] = fe_exp(g_inv, s->first_step);
        }
    }
    return 0;
}
static void cons_free(cons_t *k) {
    for (uint32_t i = 0; i < k->na; i++) free(k->apoly[i]);
    free(k->t_alpha); free(k->t_beta); free(k->t_group); free(k->tgroups); free(k->sa); free(k->apoly); free(k->axoff);
    free(k->b_alpha); free(k->b_beta); free(k->a_group); free(k->bgroups);
}
/* merged value of all constraints at the point x:  T(x)/Z_T(x) + sum_g B_g(x)/Z_g(x)
 * (ConstraintEvaluator::evaluate + ConstraintEvaluationTable::into_poly on the prover, evaluate_constraints on the verifier) */
static fe cons_combine(const cons_t *k, fe x, const fe *t_evals, const fe *cur_row) {
    fe xp[16];
    for (uint32_t g = 0; g < k->ntg; g++) xp[g] = fe_exp(x, k->tgroups[g].adj);
    fe t = 0;
    for (uint32_t i = 0; i < k->air->num_constraints; i++)
        t = fe_add(t, fe_mul(fe_add(k->t_alpha[i], fe_mul(k->t_beta[i], xp[k->t_group[i]])), t_evals[i]));
    /* transition divisor (x^n - 1) / (x - g^(n-1)) */
    fe result = fe_mul(fe_mul(t, fe_sub(x, k->g_inv_last)), fe_inv(fe_sub(fe_exp(x, k->n), FE_ONE)));
    uint32_t i = 0;
    for (uint32_t g = 0; g < k->nbg; g++) {
        const bgroup_t *bg = &k->bgroups[g];
        fe xpb = fe_exp(x, bg->adj), acc = 0;
        for (; i < k->na && k->a_group[i] == g; i++) {
            const air_assertion *s = &k->sa[i];
            fe v = s->nvalues == 1 ? s->values[0] : poly_eval(k->apoly[i], s->nvalues, fe_mul(x, k->axoff[i]));
            acc = fe_add(acc, fe_mul(fe_sub(cur_row[s->column], v), fe_add(k->b_alpha[i], fe_mul(k->b_beta[i], xpb))));
        }
        result = fe_add(result, fe_mul(acc, fe_inv(fe_sub(fe_exp(x, bg->num_steps), bg->offset))));
    }
    return result;
}

/* ------------------------------------------------------------------ FRI helpers (folding factor 4) */
/* degree-respecting projection of one transposed row (winterfell fri::folding::apply_drp [RECALLED]):
 * interpolate the 4 values on x*{1,z,z^2,z^3} (z a primitive 4th root of unity) and evaluate at alpha */
static fe fold_row(const fe v[4], fe x_inv, fe alpha, fe zeta_inv, fe quarter) {
    fe s02 = fe_add(v[0], v[2]), d02 = fe_sub(v[0], v[2]), s13 = fe_add(v[1], v[3]), d13 = fe_mul(fe_sub(v[1], v[3]), zeta_inv);
    fe d[4] = {fe_add(s02, s13), fe_add(d02, d13), fe_sub(s02, s13), fe_sub(d02, d13)};
    fe y = fe_mul(alpha, x_inv), r = 0;
    for (int j = 3; j >= 0; j--) r = fe_add(fe_mul(r, y), d[j]);
    return fe_mul(r, quarter);
}
void fri_fold4(const fe *e, size_t m, fe alpha, fe *out) {
    size_t q = m / 4;
    fe g = fe_root_of_unity(ilog2(m)), ginv = fe_inv(g), zeta_inv = fe_exp(ginv, q), quarter = fe_inv(fe_from_u64(4));
    fe off_inv = fe_inv(fe_from_u64(F63_GENERATOR)); /* the SAME domain offset at every layer [RECALLED: v0.3 passes options.domain_offset()] */
#pragma omp parallel if (q >= 4096)
    {
#ifdef _OPENMP
        int nt = omp_get_num_threads(), tid = omp_get_thread_num();
#else
        int nt = 1, tid = 0;
#endif
        size_t lo = q * tid / nt, hi = q * (tid + 1) / nt;
        fe xi = fe_mul(off_inv, fe_exp(ginv, lo));
        for (size_t i = lo; i < hi; i++) {
            fe v[4] = {e[i], e[i + q], e[i + 2 * q], e[i + 3 * q]};
            out[i] = fold_row(v, xi, alpha, zeta_inv, quarter);
            xi = fe_mul(xi, ginv);
        }
    }
}
static size_t fold_positions(const size_t *pos, size_t np, size_t domain, size_t *out) {
    size_t target = domain / 4, n = 0;
    for (size_t i = 0; i < np; i++) {
        size_t p = pos[i] % target; int dup = 0;
        for (size_t j = 0; j < n; j++) if (out[j] == p) dup = 1;
        if (!dup) out[n++] = p;
    }
    return n;
}

/* ================================================================== PROVER */
typedef struct { uint8_t *nodes; fe *evals; size_t m; } fri_layer_t;
static int stark_prove_ext(int air_id, const uint64_t *trace, size_t n, const uint64_t *pub, size_t npub, const stark_options *opt, int d,
                           uint8_t **proof_out, size_t *proof_len, stark_debug *dbg);
static int stark_verify_ext(int air_id, const uint64_t *pub, size_t npub, const uint8_t *proof, size_t proof_len);

int stark_prove(int air_id, const uint64_t *trace, size_t n, const uint64_t *pub, size_t npub, const stark_options *opt,
                uint8_t **proof_out, size_t *proof_len, stark_debug *dbg) {
    if (!options_ok(opt) || n < 8 || (n & (n - 1))) return -1;
    if (opt->field_extension != 1) return stark_prove_ext(air_id, trace, n, pub, npub, opt, (int)opt->field_extension, proof_out, proof_len, dbg);
    air_t *air = air_new(air_id, n, pub, npub);
    if (!air) return -2;
    const uint32_t w = air->width;
    const size_t b = opt->blowup_factor, lde_n = n * b;
    const int hf = (int)opt->hash_fn;
    double t0 = now_s(), t_start = t0;
    stark_debug local_dbg; if (!dbg) dbg = &local_dbg; memset(dbg, 0, sizeof *dbg);

    /* 0. channel: coin seeded with pub-input bytes || context bytes (ProverChannel::new) */
    buf_t seed = {0};
    for (size_t i = 0; i < npub; i++) buf_u64(&seed, pub[i]);
    write_context(&seed, w, n, opt);
    coin_t coin; coin_init(&coin, hf, seed.p, seed.len); free(seed.p);

    /* 1. trace polynomials and LDE (Trace::extend) */
    fe *polys = malloc((size_t)w * n * sizeof(fe)), *lde = malloc((size_t)w * lde_n * sizeof(fe));
#pragma omp parallel for schedule(dynamic)
    for (uint32_t c = 0; c < w; c++) {
        fe *p = polys + (size_t)c * n;
        for (size_t i = 0; i < n; i++) p[i] = fe_from_u64(trace[(size_t)c * n + i]);
        ntt_natural(p, n, 1);
        eval_with_offset(p, n, b, lde + (size_t)c * lde_n);
    }
    dbg->t_lde = now_s() - t0; t0 = now_s();

    /* 2. commit to the LDE rows */
    uint8_t *t_nodes = malloc(2 * lde_n * 32);
    {
        uint8_t *leaves = malloc(lde_n * 32);
#pragma omp parallel
        {
            fe *row = malloc(w * sizeof(fe));
#pragma omp for schedule(static)
            for (size_t j = 0; j < lde_n; j++) {
                for (uint32_t c = 0; c < w; c++) row[c] = lde[(size_t)c * lde_n + j];
                hash_elements(hf, row, w, leaves + j * 32);
            }
            free(row);
        }
        merkle_build(hf, leaves, lde_n, t_nodes);
        free(leaves);
    }
    memcpy(dbg->trace_root, t_nodes + 32, 32);
    coin_reseed(&coin, t_nodes + 32);
    dbg->t_commit_trace = now_s() - t0; t0 = now_s();

    /* 3. constraint evaluation over the ce domain, merged and divided (ConstraintEvaluator::evaluate + into_poly) */
    cons_t K;
    if (cons_init(&K, air, &coin)) return -3;
    const size_t ce = K.ce, ce_n = K.ce_n, lde_stride = b / ce;
    if (ce > b) return -4;
    fe *combined = malloc(ce_n * sizeof(fe));
    {
        /* periodic column values on the ce domain: column of cycle P repeats with period P*ce */
        fe **ptab = malloc(air->num_periodic * sizeof(fe *));
        fe offset = fe_from_u64(F63_GENERATOR), g_ce = fe_root_of_unity(ilog2(ce_n));
        for (uint32_t c = 0; c < air->num_periodic; c++) {
            size_t P = air->periodic_len[c], per = P * ce;
            fe *poly = malloc(P * sizeof(fe));
            memcpy(poly, air->periodic[c], P * sizeof(fe));
            ntt_natural(poly, P, 1);
            /* values at y_s = offset^(n/P) * w_per^s, s < per: one size-`per` NTT of the coefficients scaled by offset^(n/P * i) */
            ptab[c] = calloc(per, sizeof(fe));
            fe y0 = fe_exp(offset, n / P), sc = FE_ONE;
            for (size_t i = 0; i < P; i++) { ptab[c][i] = fe_mul(poly[i], sc); sc = fe_mul(sc, y0); }
            ntt_natural(ptab[c], per, 0);
            free(poly);
        }
#pragma omp parallel
        {
            fe *cur = malloc(w * sizeof(fe)), *nxt = malloc(w * sizeof(fe)), *res = malloc(air->num_constraints * sizeof(fe));
            fe pv[64];
#pragma omp for schedule(static)
            for (size_t s = 0; s < ce_n; s++) {
                size_t j = s * lde_stride, jn = (j + b) % lde_n;
                for (uint32_t c = 0; c < w; c++) { cur[c] = lde[(size_t)c * lde_n + j]; nxt[c] = lde[(size_t)c * lde_n + jn]; }
                for (uint32_t c = 0; c < air->num_periodic; c++) pv[c] = ptab[c][s % (air->periodic_len[c] * ce)];
                memset(res, 0, air->num_constraints * sizeof(fe));
                air->eval(air, cur, nxt, pv, res);
                fe x = fe_mul(offset, fe_exp(g_ce, s));
                combined[s] = cons_combine(&K, x, res, cur);
            }
            free(cur); free(nxt); free(res);
        }
        for (uint32_t c = 0; c < air->num_periodic; c++) free(ptab[c]);
        free(ptab);
    }
    dbg->t_constraints = now_s() - t0; t0 = now_s();

    /* 4. composition polynomial: interpolate over the ce coset, split into ce columns, LDE, commit */
    fe *cpolys = malloc(ce_n * sizeof(fe)), *clde = malloc(ce * lde_n * sizeof(fe));
    uint8_t *c_nodes = malloc(2 * lde_n * 32);
    {
        ntt_natural(combined, ce_n, 1);
        fe oinv = fe_inv(fe_from_u64(F63_GENERATOR)), s = FE_ONE;
        for (size_t i = 0; i < ce_n; i++) { combined[i] = fe_mul(combined[i], s); s = fe_mul(s, oinv); }
        for (size_t i = 0; i < ce_n; i++) cpolys[(i % ce) * n + i / ce] = combined[i]; /* CompositionPoly::new -> transpose */
#pragma omp parallel for
        for (size_t r = 0; r < ce; r++) eval_with_offset(cpolys + r * n, n, b, clde + r * lde_n);
        uint8_t *leaves = malloc(lde_n * 32);
#pragma omp parallel for schedule(static)
        for (size_t j = 0; j < lde_n; j++) {
            fe row[16];
            for (size_t r = 0; r < ce; r++) row[r] = clde[r * lde_n + j];
            hash_elements(hf, row, ce, leaves + j * 32);
        }
        merkle_build(hf, leaves, lde_n, c_nodes);
        free(leaves);
    }
    free(combined);
    memcpy(dbg->constraint_root, c_nodes + 32, 32);
    coin_reseed(&coin, c_nodes + 32);
    dbg->t_composition = now_s() - t0; t0 = now_s();

    /* 5. out-of-domain point and frame */
    fe z; if (coin_draw(&coin, &z)) return -3;
    dbg->z = fe_to_u64(z);
    fe zg = fe_mul(z, K.g), zm = fe_exp(z, ce);
    fe *ood_cur = malloc(w * sizeof(fe)), *ood_next = malloc(w * sizeof(fe)), ood_comp[16];
#pragma omp parallel for
    for (uint32_t c = 0; c < w; c++) { ood_cur[c] = poly_eval(polys + (size_t)c * n, n, z); ood_next[c] = poly_eval(polys + (size_t)c * n, n, zg); }
    for (size_t r = 0; r < ce; r++) ood_comp[r] = poly_eval(cpolys + r * n, n, zm);
    { uint8_t d[32]; hash_elements(hf, ood_cur, w, d); coin_reseed(&coin, d); hash_elements(hf, ood_next, w, d); coin_reseed(&coin, d);
      hash_elements(hf, ood_comp, ce, d); coin_reseed(&coin, d); }

    /* 6. DEEP composition polynomial (DeepCompositionPoly) and its LDE */
    fe *dc_a = malloc(w * sizeof(fe)), *dc_b = malloc(w * sizeof(fe)), dc_c[16], dc_l, dc_m, unused;
    for (uint32_t c = 0; c < w; c++) if (coin_draw(&coin, &dc_a[c]) || coin_draw(&coin, &dc_b[c]) || coin_draw(&coin, &unused)) return -3;
    for (size_t r = 0; r < ce; r++) if (coin_draw(&coin, &dc_c[r])) return -3;
    if (coin_draw(&coin, &dc_l) || coin_draw(&coin, &dc_m)) return -3;
    fe *deep_evals = malloc(lde_n * sizeof(fe));
    {
        fe *t1 = calloc(n, sizeof(fe)), *t2 = calloc(n, sizeof(fe)), *t3 = calloc(n, sizeof(fe));
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n; i++) {
            fe a = 0, bb = 0, cc = 0;
            for (uint32_t c = 0; c < w; c++) { fe p = polys[(size_t)c * n + i]; a = fe_add(a, fe_mul(p, dc_a[c])); bb = fe_add(bb, fe_mul(p, dc_b[c])); }
            for (size_t r = 0; r < ce; r++) cc = fe_add(cc, fe_mul(cpolys[r * n + i], dc_c[r]));
            t1[i] = a; t2[i] = bb; t3[i] = cc;
        }
        for (uint32_t c = 0; c < w; c++) { t1[0] = fe_sub(t1[0], fe_mul(ood_cur[c], dc_a[c])); t2[0] = fe_sub(t2[0], fe_mul(ood_next[c], dc_b[c])); }
        for (size_t r = 0; r < ce; r++) t3[0] = fe_sub(t3[0], fe_mul(ood_comp[r], dc_c[r]));
        fe *num[3] = {t1, t2, t3}, pt[3] = {z, zg, zm};
        for (int q = 0; q < 3; q++) { /* synthetic division by (x - pt) */
            fe carry = 0;
            for (size_t i = n; i-- > 0;) { fe t = fe_add(num[q][i], fe_mul(pt[q], carry)); num[q][i] = carry; carry = t; }
        }
        fe *deep = malloc(n * sizeof(fe));
        for (size_t i = 0; i < n; i++) t1[i] = fe_add(fe_add(t1[i], t2[i]), t3[i]);
        for (size_t i = 0; i < n; i++) deep[i] = fe_add(fe_mul(t1[i], dc_l), i ? fe_mul(t1[i - 1], dc_m) : 0); /* adjust_degree: * (l + m x) */
        eval_with_offset(deep, n, b, deep_evals);
        free(t1); free(t2); free(t3); free(deep);
    }
    dbg->t_deep = now_s() - t0; t0 = now_s();

    /* 7. FRI commit phase (FriProver::build_layers): num_fri_layers folds + the committed remainder layer */
    size_t nlayers = num_fri_layers(opt, lde_n) + 1;
    fri_layer_t *layers = calloc(nlayers, sizeof *layers);
    {
        fe *cur = deep_evals; size_t m = lde_n;
        for (size_t l = 0; l < nlayers; l++) {
            size_t q = m / 4;
            uint8_t *leaves = malloc(q * 32);
#pragma omp parallel for schedule(static) if (q >= 1024)
            for (size_t i = 0; i < q; i++) { fe row[4] = {cur[i], cur[i + q], cur[i + 2 * q], cur[i + 3 * q]}; hash_elements(hf, row, 4, leaves + i * 32); }
            layers[l].nodes = malloc(2 * q * 32); layers[l].evals = cur; layers[l].m = m;
            merkle_build(hf, leaves, q, layers[l].nodes);
            free(leaves);
            const uint8_t *root = q == 1 ? layers[l].nodes + 32 : layers[l].nodes + 32;
            coin_reseed(&coin, root);
            memcpy(dbg->fri_roots[l], root, 32);
            fe alpha; if (coin_draw(&coin, &alpha)) return -3;
            dbg->fri_alphas[l] = fe_to_u64(alpha);
            if (l + 1 < nlayers) { fe *nx = malloc(q * sizeof(fe)); fri_fold4(cur, m, alpha, nx); cur = nx; m = q; }
        }
        dbg->num_fri_layers = (uint32_t)nlayers;
    }
    dbg->t_fri = now_s() - t0; t0 = now_s();

    /* 8. proof of work and query positions */
    uint64_t nonce = 1;
    while (coin_check_leading_zeros(&coin, nonce) < opt->grinding_factor) nonce++;
    coin_reseed_int(&coin, nonce);
    dbg->pow_nonce = nonce;
    size_t nq = opt->num_queries, *pos = malloc(nq * sizeof(size_t));
    if (coin_draw_integers(&coin, nq, lde_n, pos)) return -3;
    dbg->num_positions = (uint32_t)nq;
    for (size_t i = 0; i < nq && i < 256; i++) dbg->positions[i] = pos[i];

    /* 9. assemble the proof (StarkProof::to_bytes) */
    buf_t P = {0};
    write_context(&P, w, n, opt);
    buf_u16(&P, (uint16_t)((2 + nlayers) * 32));
    buf_put(&P, t_nodes + 32, 32); buf_put(&P, c_nodes + 32, 32);
    for (size_t l = 0; l < nlayers; l++) buf_put(&P, layers[l].nodes + 32, 32);
    uint8_t *pathbuf = malloc(1 + nq * (1 + 32 * (ilog2(lde_n) + 1)));
    { /* trace queries */
        buf_u32(&P, (uint32_t)(nq * w * 8));
        for (size_t i = 0; i < nq; i++) for (uint32_t c = 0; c < w; c++) buf_fe(&P, lde[(size_t)c * lde_n + pos[i]]);
        size_t pl = merkle_prove_batch(t_nodes, lde_n, pos, nq, pathbuf);
        buf_u32(&P, (uint32_t)pl); buf_put(&P, pathbuf, pl);
    }
    { /* constraint queries */
        buf_u32(&P, (uint32_t)(nq * ce * 8));
        for (size_t i = 0; i < nq; i++) for (size_t r = 0; r < ce; r++) buf_fe(&P, clde[r * lde_n + pos[i]]);
        size_t pl = merkle_prove_batch(c_nodes, lde_n, pos, nq, pathbuf);
        buf_u32(&P, (uint32_t)pl); buf_put(&P, pathbuf, pl);
    }
    buf_u16(&P, (uint16_t)(w * 8));
    for (uint32_t c = 0; c < w; c++) buf_fe(&P, ood_cur[c]);
    for (uint32_t c = 0; c < w; c++) buf_fe(&P, ood_next[c]);
    buf_u16(&P, (uint16_t)(ce * 8));
    for (size_t r = 0; r < ce; r++) buf_fe(&P, ood_comp[r]);
    { /* FRI proof (FriProver::build_proof) */
        buf_u8(&P, (uint8_t)(nlayers - 1));
        size_t *fp = malloc(nq * sizeof(size_t)), *fp2 = malloc(nq * sizeof(size_t)), nfp = nq, domain = lde_n;
        memcpy(fp, pos, nq * sizeof(size_t));
        for (size_t l = 0; l + 1 < nlayers; l++) {
            nfp = fold_positions(fp, nfp, domain, fp2);
            memcpy(fp, fp2, nfp * sizeof(size_t));
            size_t q = domain / 4; const fe *e = layers[l].evals;
            buf_u32(&P, (uint32_t)(nfp * 4 * 8));
            for (size_t i = 0; i < nfp; i++) for (int k2 = 0; k2 < 4; k2++) buf_fe(&P, e[fp[i] + (size_t)k2 * q]);
            size_t pl = merkle_prove_batch(layers[l].nodes, q, fp, nfp, pathbuf);
            buf_u32(&P, (uint32_t)pl); buf_put(&P, pathbuf, pl);
            domain = q;
        }
        const fri_layer_t *last = &layers[nlayers - 1];
        buf_u16(&P, (uint16_t)(last->m * 8));
        for (size_t i = 0; i < last->m; i++) buf_fe(&P, last->evals[i]); /* un-transposed remainder = natural order */
        buf_u8(&P, 1); /* num_partitions */
        free(fp); free(fp2);
    }
    buf_u64(&P, nonce);
    dbg->t_queries = now_s() - t0; dbg->t_total = now_s() - t_start;

    for (size_t l = 0; l < nlayers; l++) { free(layers[l].nodes); free(layers[l].evals); }
    free(layers); free(pathbuf); free(pos); free(dc_a); free(dc_b); free(ood_cur); free(ood_next);
    free(cpolys); free(clde); free(c_nodes); free(t_nodes); free(polys); free(lde);
    cons_free(&K); air_free(air);
    *proof_out = P.p; *proof_len = P.len;
    return 0;
}

/* ================================================================== VERIFIER */
typedef struct { const uint8_t *p; size_t len, off; int err; } rd_t;
static const uint8_t *rd_take(rd_t *r, size_t n) { if (r->err || r->off + n > r->len) { r->err = 1; return NULL; } const uint8_t *q = r->p + r->off; r->off += n; return q; }
static uint64_t rd_uint(rd_t *r, int bytes) { const uint8_t *q = rd_take(r, bytes); uint64_t v = 0; if (q) for (int i = 0; i < bytes; i++) v |= (uint64_t)q[i] << (8 * i); return v; }
static int rd_fe(rd_t *r, fe *out) { uint64_t v = rd_uint(r, 8); if (r->err || v >= F63_P) { r->err = 1; return 1; } *out = fe_from_u64(v); return 0; }

int stark_verify(int air_id, const uint64_t *pub, size_t npub, const uint8_t *proof, size_t proof_len) {
    if (proof_len > 18 && proof[2] == 0 && proof[3] == 0 && proof[17] != 1) return stark_verify_ext(air_id, pub, npub, proof, proof_len);   /* field_extension byte */
    rd_t R = {proof, proof_len, 0, 0};
    /* context */
    uint32_t w = (uint32_t)rd_uint(&R, 1); unsigned logn = (unsigned)rd_uint(&R, 1);
    size_t meta = rd_uint(&R, 2); rd_take(&R, meta);
    size_t modlen = rd_uint(&R, 1); const uint8_t *mod = rd_take(&R, modlen);
    stark_options o;
    o.num_queries = (uint32_t)rd_uint(&R, 1); o.blowup_factor = 1u << rd_uint(&R, 1); o.grinding_factor = (uint32_t)rd_uint(&R, 1);
    o.hash_fn = (uint32_t)rd_uint(&R, 1); o.field_extension = (uint32_t)rd_uint(&R, 1);
    o.fri_folding_factor = 1u << rd_uint(&R, 1); o.fri_max_remainder_size = 1u << rd_uint(&R, 1);
    if (R.err || modlen != 8 || !options_ok(&o) || logn < 3 || logn > 40) return -1;
    { uint64_t m; memcpy(&m, mod, 8); if (m != F63_P) return -1; }
    const size_t n = (size_t)1 << logn, b = o.blowup_factor, lde_n = n * b, nq = o.num_queries;
    const int hf = (int)o.hash_fn;
    air_t *air = air_new(air_id, n, pub, npub);
    if (!air) return -2;
    if (air->width != w) { air_free(air); return -2; }
    const size_t nlayers = num_fri_layers(&o, lde_n) + 1;
    /* commitments */
    size_t clen = rd_uint(&R, 2); const uint8_t *commits = rd_take(&R, clen);
    if (R.err || clen != (2 + nlayers) * 32) { air_free(air); return -1; }
    /* queries */
    size_t tv_len = rd_uint(&R, 4); const uint8_t *tv = rd_take(&R, tv_len); size_t tp_len = rd_uint(&R, 4); const uint8_t *tp = rd_take(&R, tp_len);
    size_t cv_len = rd_uint(&R, 4); const uint8_t *cv = rd_take(&R, cv_len); size_t cp_len = rd_uint(&R, 4); const uint8_t *cp = rd_take(&R, cp_len);
    if (R.err) { air_free(air); return -1; }
    int rc = 0;
    buf_t seed = {0};
    for (size_t i = 0; i < npub; i++) buf_u64(&seed, pub[i]);
    write_context(&seed, w, n, &o);
    coin_t coin; coin_init(&coin, hf, seed.p, seed.len); free(seed.p);
    cons_t K; memset(&K, 0, sizeof K);
    fe *ood_cur = malloc(w * sizeof(fe)), *ood_next = malloc(w * sizeof(fe)), ood_comp[16];
    fe *t_rows = NULL, *c_rows = NULL, *deep = NULL; size_t *pos = NULL;
    fe *dc_a = malloc(w * sizeof(fe)), *dc_b = malloc(w * sizeof(fe)), dc_c[16], dc_l, dc_m, unused;

    /* 1. trace commitment, constraint coefficients */
    coin_reseed(&coin, commits);
    if (cons_init(&K, air, &coin)) { rc = -3; goto done; }
    const size_t ce = K.ce;
    if (tv_len != nq * w * 8 || cv_len != nq * ce * 8) { rc = -1; goto done; }
    /* 2. constraint commitment, OOD point */
    coin_reseed(&coin, commits + 32);
    fe z; if (coin_draw(&coin, &z)) { rc = -3; goto done; }
    /* 3. OOD frame and consistency check */
    if (rd_uint(&R, 2) != w * 8) { rc = -1; goto done; }
    for (uint32_t c = 0; c < w; c++) rd_fe(&R, &ood_cur[c]);
    for (uint32_t c = 0; c < w; c++) rd_fe(&R, &ood_next[c]);
    if (rd_uint(&R, 2) != ce * 8) { rc = -1; goto done; }
    for (size_t r = 0; r < ce; r++) rd_fe(&R, &ood_comp[r]);
    if (R.err) { rc = -1; goto done; }
    {
        fe pv[64], *res = calloc(air->num_constraints, sizeof(fe));
        for (uint32_t c = 0; c < air->num_periodic; c++) {
            size_t Pn = air->periodic_len[c];
            fe *poly = malloc(Pn * sizeof(fe)); memcpy(poly, air->periodic[c], Pn * sizeof(fe));
            ntt_natural(poly, Pn, 1);
            pv[c] = poly_eval(poly, Pn, fe_exp(z, n / Pn));
            free(poly);
        }
        air->eval(air, ood_cur, ood_next, pv, res);
        fe e1 = cons_combine(&K, z, res, ood_cur), e2 = 0, zp = FE_ONE;
        free(res);
        for (size_t r = 0; r < ce; r++) { e2 = fe_add(e2, fe_mul(zp, ood_comp[r])); zp = fe_mul(zp, z); }
        uint8_t d[32];
        hash_elements(hf, ood_cur, w, d); coin_reseed(&coin, d); hash_elements(hf, ood_next, w, d); coin_reseed(&coin, d);
        hash_elements(hf, ood_comp, ce, d); coin_reseed(&coin, d);
        if (e1 != e2) { rc = 3; goto done; } /* InconsistentOodConstraintEvaluations */
    }
    /* 4. DEEP coefficients, FRI layer commitments and alphas */
    for (uint32_t c = 0; c < w; c++) if (coin_draw(&coin, &dc_a[c]) || coin_draw(&coin, &dc_b[c]) || coin_draw(&coin, &unused)) { rc = -3; goto done; }
    for (size_t r = 0; r < ce; r++) if (coin_draw(&coin, &dc_c[r])) { rc = -3; goto done; }
    if (coin_draw(&coin, &dc_l) || coin_draw(&coin, &dc_m)) { rc = -3; goto done; }
    fe alphas[16];
    for (size_t l = 0; l < nlayers; l++) { coin_reseed(&coin, commits + (2 + l) * 32); if (coin_draw(&coin, &alphas[l])) { rc = -3; goto done; } }
    /* parse the FRI proof + nonce */
    size_t fl = rd_uint(&R, 1);
    if (R.err || fl != nlayers - 1) { rc = -1; goto done; }
    const uint8_t *lv[16], *lp[16]; size_t lv_len[16], lp_len[16];
    for (size_t l = 0; l < fl; l++) { lv_len[l] = rd_uint(&R, 4); lv[l] = rd_take(&R, lv_len[l]); lp_len[l] = rd_uint(&R, 4); lp[l] = rd_take(&R, lp_len[l]); }
    size_t rem_len = rd_uint(&R, 2); const uint8_t *rem_bytes = rd_take(&R, rem_len);
    size_t nparts = rd_uint(&R, 1); uint64_t nonce = rd_uint(&R, 8);
    if (R.err || R.off != R.len || nparts != 1) { rc = -1; goto done; }
    /* 5. proof of work, query positions, openings against both commitments */
    coin_reseed_int(&coin, nonce);
    if (coin_leading_zeros(&coin) < o.grinding_factor) { rc = 5; goto done; }
    pos = malloc(nq * sizeof(size_t));
    if (coin_draw_integers(&coin, nq, lde_n, pos)) { rc = -3; goto done; }
    t_rows = malloc(nq * w * sizeof(fe)); c_rows = malloc(nq * ce * sizeof(fe));
    {
        rd_t T = {tv, tv_len, 0, 0}, C = {cv, cv_len, 0, 0};
        uint8_t *lh = malloc(nq * 32), root[32];
        for (size_t i = 0; i < nq; i++) { for (uint32_t c = 0; c < w; c++) rd_fe(&T, &t_rows[i * w + c]); hash_elements(hf, t_rows + i * w, w, lh + i * 32); }
        if (T.err || merkle_batch_root(hf, tp, tp_len, pos, lh, nq, ilog2(lde_n), root) || memcmp(root, commits, 32)) { free(lh); rc = 6; goto done; }
        for (size_t i = 0; i < nq; i++) { for (size_t r = 0; r < ce; r++) rd_fe(&C, &c_rows[i * ce + r]); hash_elements(hf, c_rows + i * ce, ce, lh + i * 32); }
        if (C.err || merkle_batch_root(hf, cp, cp_len, pos, lh, nq, ilog2(lde_n), root) || memcmp(root, commits + 32, 32)) { free(lh); rc = 7; goto done; }
        free(lh);
    }
    /* 6. DEEP composition at the queried points (DeepComposer) */
    deep = malloc(nq * sizeof(fe));
    {
        fe offset = fe_from_u64(F63_GENERATOR), g_lde = fe_root_of_unity(ilog2(lde_n)), zg = fe_mul(z, K.g), zm = fe_exp(z, ce);
        for (size_t i = 0; i < nq; i++) {
            fe x = fe_mul(offset, fe_exp(g_lde, pos[i])), a = 0, bb = 0, cc = 0;
            for (uint32_t c = 0; c < w; c++) {
                a = fe_add(a, fe_mul(dc_a[c], fe_sub(t_rows[i * w + c], ood_cur[c])));
                bb = fe_add(bb, fe_mul(dc_b[c], fe_sub(t_rows[i * w + c], ood_next[c])));
            }
            for (size_t r = 0; r < ce; r++) cc = fe_add(cc, fe_mul(dc_c[r], fe_sub(c_rows[i * ce + r], ood_comp[r])));
            fe sum = fe_add(fe_add(fe_mul(a, fe_inv(fe_sub(x, z))), fe_mul(bb, fe_inv(fe_sub(x, zg)))), fe_mul(cc, fe_inv(fe_sub(x, zm))));
            deep[i] = fe_mul(sum, fe_add(dc_l, fe_mul(dc_m, x)));
        }
    }
    /* 7. FRI verification (FriVerifier::verify) */
    {
        size_t *p1 = malloc(nq * sizeof(size_t)), *p2 = malloc(nq * sizeof(size_t)), np1 = nq, domain = lde_n;
        fe *ev = malloc(nq * sizeof(fe)), *ev2 = malloc(nq * sizeof(fe));
        memcpy(p1, pos, nq * sizeof(size_t)); memcpy(ev, deep, nq * sizeof(fe));
        size_t max_deg_plus_1 = n; /* trace_poly_degree + 1 */
        fe off_inv = fe_inv(fe_from_u64(F63_GENERATOR)), quarter = fe_inv(fe_from_u64(4));
        for (size_t l = 0; l < fl && !rc; l++) {
            size_t q = domain / 4, np2 = fold_positions(p1, np1, domain, p2);
            fe ginv = fe_inv(fe_root_of_unity(ilog2(domain))), zeta_inv = fe_exp(ginv, q);
            if (lv_len[l] != np2 * 32) { rc = 8; break; }
            fe *vals = malloc(np2 * 4 * sizeof(fe)); uint8_t *lh = malloc(np2 * 32), root[32];
            rd_t V = {lv[l], lv_len[l], 0, 0};
            for (size_t i = 0; i < np2; i++) { for (int k2 = 0; k2 < 4; k2++) rd_fe(&V, &vals[i * 4 + k2]); hash_elements(hf, vals + i * 4, 4, lh + i * 32); }
            if (V.err || merkle_batch_root(hf, lp[l], lp_len[l], p2, lh, np2, ilog2(q), root) || memcmp(root, commits + (2 + l) * 32, 32)) rc = 8;
            /* the value carried from the previous layer must sit in the opened row */
            for (size_t i = 0; i < np1 && !rc; i++) {
                size_t fpos = p1[i] % q, row = 0;
                while (p2[row] != fpos) row++;
                if (vals[row * 4 + p1[i] / q] != ev[i]) rc = 9; /* InvalidLayerFolding */
            }
            for (size_t i = 0; i < np2; i++) ev2[i] = fold_row(vals + i * 4, fe_mul(off_inv, fe_exp(ginv, p2[i])), alphas[l], zeta_inv, quarter);
            free(vals); free(lh);
            if (max_deg_plus_1 % 4) rc = 10;
            max_deg_plus_1 /= 4; domain = q;
            memcpy(p1, p2, np2 * sizeof(size_t)); memcpy(ev, ev2, np2 * sizeof(fe)); np1 = np2;
        }
        if (!rc) { /* remainder: commitment, consistency with the last fold, degree */
            if (rem_len != domain * 8) rc = -1;
            else {
                fe *rem = malloc(domain * sizeof(fe)); rd_t Q = {rem_bytes, rem_len, 0, 0};
                for (size_t i = 0; i < domain; i++) rd_fe(&Q, &rem[i]);
                size_t q = domain / 4; uint8_t *leaves = malloc(q * 32), *nodes = malloc(2 * q * 32);
                for (size_t i = 0; i < q; i++) { fe row[4] = {rem[i], rem[i + q], rem[i + 2 * q], rem[i + 3 * q]}; hash_elements(hf, row, 4, leaves + i * 32); }
                merkle_build(hf, leaves, q, nodes);
                if (Q.err || memcmp(nodes + 32, commits + (2 + fl) * 32, 32)) rc = 11; /* RemainderCommitmentMismatch */
                for (size_t i = 0; i < np1 && !rc; i++) if (rem[p1[i]] != ev[i]) rc = 12;
                if (!rc) {
                    size_t max_degree = max_deg_plus_1 - 1;
                    if (max_degree >= domain - 1) rc = 13;
                    else { ntt_natural(rem, domain, 1); for (size_t i = max_degree + 1; i < domain; i++) if (rem[i]) rc = 14; }
                }
                free(rem); free(leaves); free(nodes);
            }
        }
        free(p1); free(p2); free(ev); free(ev2);
    }
done:
    cons_free(&K);
    free(ood_cur); free(ood_next); free(t_rows); free(c_rows); free(deep); free(pos); free(dc_a); free(dc_b);
    air_free(air);
    return rc;
}

#include "stark_ext.inc"
