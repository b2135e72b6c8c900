/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Restates /root/reference/src/utils/rescue.rs (see rescue.h). */
#include "rescue.h"
#include "ref_constants.h"
#include <string.h>

static fe MDS[196], INV_MDS[196], ARK[8][28];
static int tables_ready = 0;

void rescue_init_tables(void) {
    if (tables_ready) return;
    for (int i = 0; i < 196; i++) { MDS[i] = fe_from_u64(REF_MDS[i]); INV_MDS[i] = fe_from_u64(REF_INV_MDS[i]); }
    for (int r = 0; r < 8; r++) for (int j = 0; j < 28; j++) ARK[r][j] = fe_from_u64(REF_ARK[r * 28 + j]);
    tables_ready = 1;
}
const fe *rescue_ark(unsigned row) { rescue_init_tables(); return ARK[row % 8]; }

/* rescue.rs:327-333  x -> x^3 */
static void sbox(fe *s) { for (int i = 0; i < 14; i++) s[i] = fe_mul(s[i], fe_sqr(s[i])); }
/* rescue.rs:335-341  x -> x^(1/3) */
static void inv_sbox(fe *s) { for (int i = 0; i < 14; i++) s[i] = fe_exp(s[i], REF_INV_ALPHA); }
/* rescue.rs:343-375 */
static void mat_apply(const fe *m, fe *s) {
    fe r[14];
    for (int i = 0; i < 14; i++) {
        fe acc = 0;
        for (int j = 0; j < 14; j++) acc = fe_add(acc, fe_mul(m[i * 14 + j], s[j]));
        r[i] = acc;
    }
    memcpy(s, r, sizeof r);
}

void rescue_apply_round(fe *state, size_t step) {
    rescue_init_tables();
    const fe *ark = ARK[step % RESCUE_CYCLE];
    sbox(state); mat_apply(MDS, state);
    for (int i = 0; i < 14; i++) state[i] = fe_add(state[i], ark[i]);
    inv_sbox(state); mat_apply(MDS, state);
    for (int i = 0; i < 14; i++) state[i] = fe_add(state[i], ark[14 + i]);
}
void rescue_apply_permutation(fe *state) { for (int i = 0; i < RESCUE_NUM_ROUNDS; i++) rescue_apply_round(state, i); }

void rescue_digest(const fe *data, size_t n, fe out[7]) {
    fe st[14] = {0};
    size_t i = 0;
    for (size_t k = 0; k < n; k++) {
        st[i] = fe_add(st[i], data[k]);
        i++;
        if (i % RESCUE_RATE_WIDTH == 0) { rescue_apply_permutation(st); i = 0; }
    }
    if (i > 0) rescue_apply_permutation(st);
    memcpy(out, st, 7 * sizeof(fe));
}
void rescue_merge(const fe a[7], const fe b[7], fe out[7]) {
    fe st[14];
    memcpy(st, a, 7 * sizeof(fe)); memcpy(st + 7, b, 7 * sizeof(fe));
    rescue_apply_permutation(st);
    memcpy(out, st, 7 * sizeof(fe));
}
void rescue_enforce_round(fe *result, const fe *cur, const fe *next, const fe *ark, fe flag) {
    rescue_init_tables();
    fe s1[14], s2[14];
    memcpy(s1, cur, sizeof s1);
    sbox(s1); mat_apply(MDS, s1);
    for (int i = 0; i < 14; i++) s1[i] = fe_add(s1[i], ark[i]);
    for (int i = 0; i < 14; i++) s2[i] = fe_sub(next[i], ark[14 + i]);
    mat_apply(INV_MDS, s2); sbox(s2);
    for (int i = 0; i < 14; i++) result[i] = fe_add(result[i], fe_mul(flag, fe_sub(s2[i], s1[i])));
}
