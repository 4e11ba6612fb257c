/*
 * sor_dispatch.c -- glue compiled INTO oracle/_ref/libsf_ref.so next to the reference's own,
 * unmodified objects.  TEST INFRASTRUCTURE ONLY.
 *
 * The reference's solver.c is compiled with -Dsor_coupled=sor_coupled_lex, so its lexicographic
 * solver keeps its code but gets the name sor_coupled_lex.  This file provides the symbol
 * `sor_coupled` that variational.c:57 and variational_mt.cpp:368 call, and forwards either to the
 * reference's solver (mode 0) or to a red-black re-ordering of the same update (mode 1) -- the
 * "CPU red-black reference mode" that isolates the ordering change of the GPU solver.
 */
#include <stdlib.h>

typedef struct image_s {
    int width, height, stride;
    float *data;
} image_t;

void sor_coupled_lex(image_t *du, image_t *dv, image_t *a11, image_t *a12, image_t *a22, image_t *b1, image_t *b2,
                     image_t *dpsis_horiz, image_t *dpsis_vert, const int iterations, const float omega);

static __thread int g_sor_mode = 0; /* 0 = reference lexicographic, 1 = red-black */

void sf_ref_set_sor_mode(int mode) { g_sor_mode = mode; }
int sf_ref_get_sor_mode(void) { return g_sor_mode; }

static void sor_coupled_redblack(image_t *du, image_t *dv, image_t *a11, image_t *a12, image_t *a22, image_t *b1,
                                 image_t *b2, image_t *ph, image_t *pv, const int iterations, const float omega) {
    const int W = du->width, H = du->height, S = du->stride;
    /* block inverse, same expression order as solver.c:101-106 */
    for (int j = 0; j < H; j++)
        for (int i = 0; i < W; i++) {
            const size_t o = (size_t)j * S + i;
            float sp = ((i > 0) ? ph->data[o - 1] : 0.0f) + ph->data[o];
            if (j > 0) sp = sp + pv->data[o - S];
            if (j < H - 1) sp = sp + pv->data[o];
            const float A11 = a22->data[o] + sp, A22 = a11->data[o] + sp;
            const float det = A11 * A22 - a12->data[o] * a12->data[o];
            a11->data[o] = A11 / det;
            a22->data[o] = A22 / det;
            a12->data[o] = a12->data[o] / -det;
        }
    for (int it = 0; it < iterations; it++)
        for (int colour = 0; colour < 2; colour++)
            for (int j = 0; j < H; j++)
                for (int i = (j + colour) & 1; i < W; i += 2) {
                    const size_t o = (size_t)j * S + i;
                    const float dur = (i < W - 1) ? du->data[o + 1] : 0.0f, dvr = (i < W - 1) ? dv->data[o + 1] : 0.0f;
                    float s1 = ph->data[o] * dur, s2 = ph->data[o] * dvr;
                    if (j > 0) { s1 = s1 + pv->data[o - S] * du->data[o - S]; s2 = s2 + pv->data[o - S] * dv->data[o - S]; }
                    if (j < H - 1) { s1 = s1 + pv->data[o] * du->data[o + S]; s2 = s2 + pv->data[o] * dv->data[o + S]; }
                    s1 = s1 + b1->data[o];
                    s2 = s2 + b2->data[o];
                    float B1 = s1, B2 = s2;
                    if (i > 0) { B1 = ph->data[o - 1] * du->data[o - 1] + s1; B2 = ph->data[o - 1] * dv->data[o - 1] + s2; }
                    du->data[o] += omega * (a11->data[o] * B1 + a12->data[o] * B2 - du->data[o]);
                    dv->data[o] += omega * (a12->data[o] * B1 + a22->data[o] * B2 - dv->data[o]);
                }
}

void sor_coupled(image_t *du, image_t *dv, image_t *a11, image_t *a12, image_t *a22, image_t *b1, image_t *b2,
                 image_t *dpsis_horiz, image_t *dpsis_vert, const int iterations, const float omega) {
    if (g_sor_mode == 1 && du->width >= 2 && du->height >= 2 && iterations >= 1)
        sor_coupled_redblack(du, dv, a11, a12, a22, b1, b2, dpsis_horiz, dpsis_vert, iterations, omega);
    else
        sor_coupled_lex(du, dv, a11, a12, a22, b1, b2, dpsis_horiz, dpsis_vert, iterations, omega);
}
