// epic_api.cpp -- C entry points over the reference's UNMODIFIED EPIC interpolation (epic_flow_extended/epic.cpp +
// epic_aux.cpp, compiled in place into oracle/_ref/libsf_ref.so; LAPACK's sgels_ comes from lapack_stub.c).
// TEST INFRASTRUCTURE ONLY.
#include <string.h>

#include "epic.h"     // the reference's own header: epic_params_t, epic(), float_image
#include "epic_aux.h" // dist_trf_nnfield_subset & co. for operator-level checks

extern "C" {

void sf_ref_epic_params_default(epic_params_t *p) { epic_params_default(p); }

// epic() exactly as epicflow.cpp:125 / slow_flow.cpp:819 call it.  matches: n x tx floats (x1 y1 x2 y2 ...); edges: W*H
// costs, modified in place like in the reference (+= euc).
void sf_ref_epic(image_t *flowx, image_t *flowy, const color_image_t *im, float *matches, int n, int tx, float *edges,
                 const epic_params_t *params) {
    float_image m = {matches, tx, n};
    float_image e = {edges, im->width, im->height};
    epic(flowx, flowy, im, &m, &e, params, 1);
}

// geodesic k-nearest-seed search (epic_aux.cpp:350-401) with the seeds themselves as query points, plus the label map
void sf_ref_dist_trf_nnfield(int *best, float *dist, int *labels, const int *seeds, int ns, int nn, float *cost, int w, int h) {
    int_image b = {best, nn, ns}, l = {labels, w, h}, s = {(int *)seeds, 2, ns};
    float_image d = {dist, nn, ns}, c = {cost, w, h};
    dist_trf_nnfield_subset(&b, &d, &l, &s, &c, NULL, &s, 1);
}

image_t *sf_ref_saliency(const color_image_t *im, float sigma_image, float sigma_matrix) { return saliency(im, sigma_image, sigma_matrix); }

} // extern "C"
