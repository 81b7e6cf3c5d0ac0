/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the PointPillars front end of the GenComm hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (gencomm_b200/) never links, imports or calls it.
 *
 * PARITY STATUS
 *   gc_ref_voxelize : "parity unpinned".  The arithmetic lives in the third-party package
 *       spconv (+cumm), which is neither vendored under /root/reference nor version-pinned by it
 *       (README.md:116 "pip install spconv-cu116"; requirements.txt has no entry; setup.py:27
 *       install_requires=[]).  The call sites are
 *       opencood/data_utils/pre_processor/sp_voxel_preprocessor.py:47-60 (constructor arguments)
 *       and :62-85 (point_to_voxel / generate).  This file restates the published algorithm of
 *       spconv Point2VoxelCPU3d.point_to_voxel == spconv 1.x VoxelGeneratorV2.generate
 *       (SURVEY.md App. A.1): a single sequential pass over the points in input order.
 *       The reference holds no golden vector for it.
 *   gc_ref_scatter  : restates opencood/models/sub_modules/point_pillar_scatter.py:45-73; pinned
 *       against the reference class by oracle/gen_golden.py (bit-exact).
 *   gc_ref_pillar_vfe : restates opencood/models/sub_modules/pillar_vfe.py:105-155 (+ PFNLayer
 *       :31-53) with ONE fixed floating-point evaluation order (documented below, an algebraic
 *       regrouping of the same sum) that the CUDA
 *       kernel also follows, so kernel-vs-oracle comparison is bit-exact; the oracle itself is
 *       pinned against the reference class (torch evaluation order) to <=1e-5 relative by
 *       oracle/gen_golden.py / tests/test_oracle_cpu.py.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no implicit FMA, every fused
 * multiply-add below is an explicit fmaf()).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------------------------------------
 * Voxelizer.  spconv semantics, restated:
 *   for each point i in input order:
 *       c_j = floor((p[i][j] - range_min[j]) / vsize[j])   (float32 arithmetic), j = x,y,z
 *       skip the point if any c_j < 0 or c_j >= grid[j]
 *       v = map[cz][cy][cx]                                  (dense int map, -1 = empty)
 *       if v == -1: if num_voxels >= max_voxels: skip;  v = num_voxels++; coords[v] = (cz,cy,cx)
 *       if num_points[v] < max_points: voxels[v][num_points[v]++] = p[i]
 * grid[j] = round((range_max[j]-range_min[j]) / vsize[j])   (sp_voxel_preprocessor.py:41-43,
 * computed by the caller in float64 and passed in).
 * Outputs are zero padded: voxels [max_voxels][max_points][4], coords [max_voxels][3] (z,y,x),
 * num_points [max_voxels].  Returns the number of voxels M.
 * ------------------------------------------------------------------------------------------- */
int gc_ref_voxelize(const float *points, int n_points, const float *range6, const float *vsize3,
                    const int *grid3 /* nx,ny,nz */, int max_points, int max_voxels,
                    float *voxels, int32_t *coords, int32_t *num_points) {
    const int nx = grid3[0], ny = grid3[1], nz = grid3[2];
    const size_t ncell = (size_t)nx * ny * nz;
    int32_t *map = (int32_t *)malloc(ncell * sizeof(int32_t));
    if (!map) return -1;
    for (size_t i = 0; i < ncell; ++i) map[i] = -1;
    memset(voxels, 0, (size_t)max_voxels * max_points * 4 * sizeof(float));
    memset(coords, 0, (size_t)max_voxels * 3 * sizeof(int32_t));
    memset(num_points, 0, (size_t)max_voxels * sizeof(int32_t));
    int num_voxels = 0;
    for (int i = 0; i < n_points; ++i) {
        const float *p = points + (size_t)i * 4;
        int c[3];
        int ok = 1;
        for (int j = 0; j < 3; ++j) {
            volatile float d = p[j] - range6[j];
            volatile float q = d / vsize3[j];
            float f = floorf(q);
            /* compare in float first: a huge |q| must not overflow the int conversion */
            if (!(f >= 0.0f) || !(f < (float)grid3[j])) { ok = 0; break; }
            c[j] = (int)f;
        }
        if (!ok) continue;
        size_t cell = ((size_t)c[2] * ny + c[1]) * nx + c[0];
        int v = map[cell];
        if (v == -1) {
            if (num_voxels >= max_voxels) continue;
            v = num_voxels++;
            map[cell] = v;
            coords[v * 3 + 0] = c[2];
            coords[v * 3 + 1] = c[1];
            coords[v * 3 + 2] = c[0];
        }
        int n = num_points[v];
        if (n < max_points) {
            memcpy(voxels + ((size_t)v * max_points + n) * 4, p, 4 * sizeof(float));
            num_points[v] = n + 1;
        }
    }
    free(map);
    return num_voxels;
}

/* ---------------------------------------------------------------------------------------------
 * PillarVFE, fixed evaluation order ("kernel order").  32 slots per pillar, n valid points.
 * The reference computes, per slot s and output channel k (pillar_vfe.py:118-149, :39, :42-46):
 *     lin = W0 x + W1 y + W2 z + W3 i + W4 (x-mx) + W5 (y-my) + W6 (z-mz) + W7 (x-cx) + W8 (y-cy) + W9 (z-cz)
 *     out = max_s relu(lin * scale + shift)            (eval BatchNorm folded: scale, shift)
 * with (mx,my,mz) the pillar mean and (cx,cy,cz) the pillar centre.  With r = p - centre (exact in
 * fp32 by Sterbenz whenever |p| >= voxel) and m' = mean(r) = m - centre this is algebraically
 *     lin*scale + shift = [A0 xr + A1 yr + A2 zr + A3 i] + [shift + B0 cx + B1 cy + B2 cz + D0 m'x + D1 m'y + D2 m'z]
 *     A_j = ((W_j + W_{4+j}) + W_{7+j}) * scale (j<3),  A_3 = W_3 * scale,  B_j = W_j * scale,  D_j = (-W_{4+j}) * scale
 * i.e. 4 instead of 10 FMAs per (point, channel), a per-pillar bias, and -- because rounding is
 * monotone -- the bias can be added after the max over the points.  The 12 constants per channel are
 * computed once on the host in fp32 (gencomm_b200.ops.pack_pfn; pack_pfn_row() below is the same
 * arithmetic).  Evaluation order (every rounding listed; the CUDA kernel is identical):
 *   centre_j  = (float)c_j * voxel_j + offset_j          (round after mul, round after add; :123-132)
 *   r_j[s]    = p_j[s] - centre_j   for s < n, 0 for padded slots
 *   sum_j     = xor-butterfly tree over the 32 slots: v[s] += v[s^16]; ^8; ^4; ^2; ^1
 *   m'_j      = sum_j * (1.0f / (float)n)
 *   b_k       = fmaf(B0,cx,shift); b = fmaf(B1,cy,b); b = fmaf(B2,cz,b);
 *               b = fmaf(D0,m'x,b); b = fmaf(D1,m'y,b); b = fmaf(D2,m'z,b)
 *   acc_k[s]  = fmaf(A0, xr, fmaf(A1, yr, fmaf(A2, zr, A3 * i)))
 *   out_k     = max(max_{s<n} acc_k[s] + b_k, 0), and max with max(shift_k, 0) when n < 32
 *               (padded slots contribute relu(bn(0)), :46)
 * coords are (b,z,y,x) rows of 4 int32.
 * ------------------------------------------------------------------------------------------- */
static float tree_sum32(const float *v) {
    float t[32];
    memcpy(t, v, sizeof(t));
    for (int m = 16; m >= 1; m >>= 1) {
        float u[32];
        for (int s = 0; s < 32; ++s) {
            volatile float r = t[s] + t[s ^ m];
            u[s] = r;
        }
        memcpy(t, u, sizeof(t));
    }
    return t[0];
}

/* row layout shared with include/gencomm_b200.h: [0..3] A, [4..6] B, [7..9] D, [10] shift, [11] max(shift,0) */
static void pack_pfn_row(const float *w /* [10] */, float scale, float shift, float *row /* [12] */) {
    for (int j = 0; j < 3; ++j) {
        volatile float t = w[j] + w[4 + j];
        t = t + w[7 + j];
        volatile float a = t * scale;
        volatile float b = w[j] * scale;
        volatile float d = (-w[4 + j]) * scale;
        row[j] = a;
        row[4 + j] = b;
        row[7 + j] = d;
    }
    volatile float a3 = w[3] * scale;
    row[3] = a3;
    row[10] = shift;
    row[11] = shift > 0.0f ? shift : 0.0f;
}

void gc_ref_pillar_vfe(const float *voxels /* [M][32][4] */, const int32_t *num_points,
                       const int32_t *coords4 /* [M][4] b,z,y,x */, int M,
                       const float *W /* [64][10] */, const float *scale, const float *shift,
                       const float *vsize3, const float *offset3 /* x,y,z centre offsets */,
                       float *out /* [M][64] */) {
    float tab[64][12];
    for (int k = 0; k < 64; ++k) pack_pfn_row(W + k * 10, scale[k], shift[k], tab[k]);
    for (int m = 0; m < M; ++m) {
        const float *vx = voxels + (size_t)m * 32 * 4;
        const int n = num_points[m];
        float centre[3], mp[3];
        /* x uses coords[:,3], y coords[:,2], z coords[:,1] */
        for (int j = 0; j < 3; ++j) {
            volatile float a = (float)coords4[m * 4 + (3 - j)] * vsize3[j];
            volatile float b = a + offset3[j];
            centre[j] = b;
        }
        float r[3][32];
        for (int s = 0; s < 32; ++s)
            for (int j = 0; j < 3; ++j) {
                volatile float d = vx[s * 4 + j] - centre[j];
                r[j][s] = (s < n) ? d : 0.0f;
            }
        volatile float inv_n = 1.0f / (float)n;
        for (int j = 0; j < 3; ++j) {
            volatile float q = tree_sum32(r[j]) * inv_n;
            mp[j] = q;
        }
        for (int k = 0; k < 64; ++k) {
            const float *t = tab[k];
            float b = fmaf(t[4], centre[0], t[10]);
            b = fmaf(t[5], centre[1], b);
            b = fmaf(t[6], centre[2], b);
            b = fmaf(t[7], mp[0], b);
            b = fmaf(t[8], mp[1], b);
            b = fmaf(t[9], mp[2], b);
            float best = -INFINITY;
            for (int s = 0; s < n && s < 32; ++s) {
                volatile float a3 = t[3] * vx[s * 4 + 3];
                float acc = fmaf(t[2], r[2][s], a3);
                acc = fmaf(t[1], r[1][s], acc);
                acc = fmaf(t[0], r[0][s], acc);
                if (acc > best) best = acc;
            }
            volatile float y = best + b;
            float o = y > 0.0f ? y : 0.0f;
            if (n < 32 && t[11] > o) o = t[11];
            out[(size_t)m * 64 + k] = o;
        }
    }
}

/* ---------------------------------------------------------------------------------------------
 * PointPillarScatter (point_pillar_scatter.py:45-73): canvas[b][c][z + y*nx + x] = feat[m][c],
 * everything else zero.  canvas is [n_batch][C][ny*nx] (nz == 1 asserted at :17).
 * ------------------------------------------------------------------------------------------- */
void gc_ref_scatter(const float *pillar_feat /* [M][C] */, const int32_t *coords4, int M, int C,
                    int nx, int ny, int n_batch, float *canvas) {
    const size_t plane = (size_t)nx * ny;
    memset(canvas, 0, (size_t)n_batch * C * plane * sizeof(float));
    for (int m = 0; m < M; ++m) {
        const int b = coords4[m * 4 + 0];
        const size_t idx = (size_t)coords4[m * 4 + 1] + (size_t)coords4[m * 4 + 2] * nx +
                           (size_t)coords4[m * 4 + 3];
        float *dst = canvas + (size_t)b * C * plane + idx;
        for (int c = 0; c < C; ++c) dst[(size_t)c * plane] = pillar_feat[(size_t)m * C + c];
    }
}
