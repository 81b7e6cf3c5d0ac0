/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle of the rotated NMS of the GenComm post-processing.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may load this library.
 *
 * gc_ref_nms_rotated restates opencood/utils/box_utils.py:915-960 (nms_rotated: sort by score, top 1000, greedy
 * suppression at IoU > threshold compared in float32) with the polygon IoU of opencood/utils/common_utils.py:230-252
 * (compute_iou over shapely Polygons, :255-270).  shapely (GEOS) is a third-party dependency absent from this image and
 * not version-pinned by the reference: "parity unpinned" for the polygon area; it is restated here as Sutherland-Hodgman
 * clipping of convex quadrilaterals in double precision (the same formulation as oracle/ref_ops.py, which pins it).
 * Score ties are ordered larger-index-first (numpy's argsort()[::-1] under a stable sort).
 */
#include <math.h>
#include <stdlib.h>

static double poly_area(const double *p, int n) {
    double a = 0.0;
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1) % n;
        a += p[2 * i] * p[2 * j + 1] - p[2 * j] * p[2 * i + 1];
    }
    return 0.5 * a;
}

static void load_ccw(const double *src, double *dst) {
    if (poly_area(src, 4) < 0.0) {
        for (int i = 0; i < 4; ++i) { dst[2 * i] = src[2 * (3 - i)]; dst[2 * i + 1] = src[2 * (3 - i) + 1]; }
    } else {
        for (int i = 0; i < 8; ++i) dst[i] = src[i];
    }
}

double gc_ref_quad_intersection_area(const double *p_in, const double *q_in) {
    double p[8], q[8], a[32], b[32], side[16];
    load_ccw(p_in, p);
    load_ccw(q_in, q);
    double *in = a, *out = b;
    int n = 4;
    for (int i = 0; i < 8; ++i) in[i] = p[i];
    for (int e = 0; e < 4 && n > 0; ++e) {
        const double ax = q[2 * e], ay = q[2 * e + 1];
        const double ex = q[2 * ((e + 1) % 4)] - ax, ey = q[2 * ((e + 1) % 4) + 1] - ay;
        for (int k = 0; k < n; ++k) side[k] = ex * (in[2 * k + 1] - ay) - ey * (in[2 * k] - ax);
        int m = 0;
        for (int k = 0; k < n; ++k) {
            const int k1 = (k + 1) % n;
            const double sc = side[k], sn = side[k1];
            if (sc >= 0.0) { out[2 * m] = in[2 * k]; out[2 * m + 1] = in[2 * k + 1]; ++m; }
            if ((sc >= 0.0) != (sn >= 0.0)) {
                const double t = sc / (sc - sn);
                out[2 * m] = in[2 * k] + t * (in[2 * k1] - in[2 * k]);
                out[2 * m + 1] = in[2 * k + 1] + t * (in[2 * k1 + 1] - in[2 * k + 1]);
                ++m;
            }
        }
        double *t2 = in; in = out; out = t2;
        n = m;
    }
    return n < 3 ? 0.0 : fabs(poly_area(in, n));
}

float gc_ref_quad_iou(const double *p, const double *q) {
    const double inter = gc_ref_quad_intersection_area(p, q);
    const double uni = fabs(poly_area(p, 4)) + fabs(poly_area(q, 4)) - inter;
    return (float)(inter / uni);
}

typedef struct { float s; int i; } key_t_;
static int cmp_desc(const void *a, const void *b) {
    const key_t_ *x = (const key_t_ *)a, *y = (const key_t_ *)b;
    if (x->s != y->s) return x->s > y->s ? -1 : 1;
    return x->i > y->i ? -1 : (x->i < y->i ? 1 : 0);
}

/* quads [n][4][2] f64, scores [n] f32 -> pick [<= top] (indices into the input, highest score first); returns the count */
int gc_ref_nms_rotated(const double *quads, const float *scores, int n, float threshold, int top, int *pick) {
    if (n <= 0) return 0;
    key_t_ *keys = (key_t_ *)malloc(sizeof(key_t_) * (size_t)n);
    for (int i = 0; i < n; ++i) { keys[i].s = scores[i]; keys[i].i = i; }
    qsort(keys, (size_t)n, sizeof(key_t_), cmp_desc);
    const int m = n < top ? n : top;
    char *dead = (char *)calloc((size_t)m, 1);
    int count = 0;
    for (int a = 0; a < m; ++a) {
        if (dead[a]) continue;
        pick[count++] = keys[a].i;
        for (int b = a + 1; b < m; ++b) {
            if (dead[b]) continue;
            if (gc_ref_quad_iou(quads + 8 * (size_t)keys[a].i, quads + 8 * (size_t)keys[b].i) > threshold) dead[b] = 1;
        }
    }
    free(keys);
    free(dead);
    return count;
}
