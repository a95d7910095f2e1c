/*
 * pdm_oracle.c -- CPU restatement of pyDEM's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the *checker* for the CUDA product in pydem_b200/csrc: tests/, the smoke
 * test and bench.py's cpu_baseline / --impl reference legs are the only users.  The
 * product never links or calls it.
 *
 * Parity status: PINNED.  Every function below is validated (tests/test_oracle_vs_reference.py,
 * run where /root/reference exists) against the unmodified reference imported in place,
 * and against the committed fixtures in tests/golden/ (made by tests/golden/make_golden.py
 * from the reference, incl. its own 5x5 known-answer vectors test_end_to_end.py:152-287).
 *
 * Written as a per-cell restatement of the algorithm (the reference is whole-array NumPy
 * plus one Cython file); each function cites the reference lines whose arithmetic it
 * follows.  All floating point is IEEE double in the reference's operation order; build
 * with -ffp-contract=off.
 *
 * Conventions: C-order grids [R][C], flat index i*C+j, row 0 = north.  dX,dY have R-1
 * entries ("fence" between row f and f+1).  thA[f]=atan2(dY[f],dX[f]), thB[f]=atan2(dX[f],dY[f])
 * are passed in by the caller (computed with numpy.arctan2 like the reference does).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;
typedef uint8_t u8;

#define PI_ 3.141592653589793

/* facet tables: dem_processing.py:173-193 (facets: cardinal e1, diagonal e2; ang_adj: (q,a)) */
static const int E1R[8] = {0, -1, -1, 0, 0, 1, 1, 0};
static const int E1C[8] = {1, 0, 0, -1, -1, 0, 0, 1};
static const int E2R[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
static const int E2C[8] = {1, 1, -1, -1, -1, -1, 1, 1};
static const int QK[8] = {0, 1, 1, 2, 2, 3, 3, 4};
static const int AK[8] = {1, -1, 1, -1, 1, -1, 1, -1};
static const int TYPE_A[8] = {1, 0, 0, 1, 1, 0, 0, 1}; /* d1=dX for facets 0,3,4,7 (1913) */

/* ------------------------------------------------------------------------------------
 * a1: one facet of one cell.  _calc_direction dem_processing.py:1942-1991 with the
 * spacing selection of _get_d1_d2 1905-1938 (upper facets use fence i-1, lower fence i;
 * 'top' -> fence 0 and 'bot' -> fence R-2 are the same rule at rows 0 and R-1).
 * Updates (*m2,*dr) only on a strictly larger squared magnitude (1986-1989).
 * ------------------------------------------------------------------------------------ */
static void facet(const double *E, i64 R, i64 C, i64 i, i64 j, int k,
                  const double *dX, const double *dY, const double *thA, const double *thB,
                  double *m2, double *dr)
{
    i64 i1 = i + E1R[k], j1 = j + E1C[k], i2 = i + E2R[k], j2 = j + E2C[k];
    if (i1 < 0 || i1 >= R || j1 < 0 || j1 >= C || i2 < 0 || i2 >= R || j2 < 0 || j2 >= C)
        return;
    i64 f = (k < 4) ? i - 1 : i;
    double d1, d2, th;
    if (TYPE_A[k]) { d1 = dX[f]; d2 = dY[f]; th = thA[f]; }
    else           { d1 = dY[f]; d2 = dX[f]; th = thB[f]; }
    double e0 = E[i * C + j], e1 = E[i1 * C + j1], e2 = E[i2 * C + j2];
    double s1 = (e0 - e1) / d1;
    double s2 = (e1 - e2) / d2;
    double s1_2 = s1 * s1;
    double sd = (e0 - e2) / sqrt(d1 * d1 + d2 * d2);
    double r = atan2(s2, s1);
    double rad2 = s1_2 + s2 * s2;
    int s1_le = s1 <= 0, s2_le = s2 <= 0, s1_gt = s1 > 0, s2_gt = s2 > 0;
    if ((s1_le && s2_gt) || (r > th)) { rad2 = sd * sd; r = th; }      /* 1973-1976 */
    if ((s1_gt && s2_le) || (r < 0)) { rad2 = s1_2; r = 0; }           /* 1978-1981 */
    if (s1_le && (s2_le || (s2_gt && (sd <= 0)))) rad2 = -1;           /* 1983-1984 */
    if (rad2 > *m2) {                                                  /* 1986-1989 */
        *m2 = rad2;
        *dr = r * (double)AK[k] + (double)QK[k] * PI_ / 2;
    }
}

/* a1: _tarboton_slopes_directions dem_processing.py:1753-1903 */
int orc_slopes_directions(const double *E, i64 R, i64 C, const double *dX, const double *dY,
                          const double *thA, const double *thB, double *mag, double *dir)
{
    if (R < 3 || C < 3) return 1;
    i64 N = R * C;
    for (i64 n = 0; n < N; n++) { mag[n] = -1; dir[n] = -1; }
    /* interior, all 8 facets in ascending order (1764-1777) */
    for (i64 i = 1; i < R - 1; i++)
        for (i64 j = 1; j < C - 1; j++)
            for (int k = 0; k < 8; k++)
                facet(E, R, C, i, j, k, dX, dY, thA, thB, &mag[i * C + j], &dir[i * C + j]);
    /* copy from the inward neighbour when it flows toward the border; four sequential
       whole-row / whole-column passes (1782-1795) */
    const double hp = PI_ / 2, thp = 3 * PI_ / 2, tp = 2 * PI_;
    for (i64 i = 0; i < R; i++) {
        double d = dir[i * C + 1];
        if (d > hp && d < thp) { dir[i * C] = d; mag[i * C] = mag[i * C + 1]; }
    }
    for (i64 i = 0; i < R; i++) {
        double d = dir[i * C + C - 2];
        if (d < hp || d > thp) { dir[i * C + C - 1] = d; mag[i * C + C - 1] = mag[i * C + C - 2]; }
    }
    for (i64 j = 0; j < C; j++) {
        double d = dir[C + j];
        if (d > 0 && d < PI_) { dir[j] = d; mag[j] = mag[C + j]; }
    }
    for (i64 j = 0; j < C; j++) {
        double d = dir[(R - 2) * C + j];
        if (d > PI_ && d < tp) { dir[(R - 1) * C + j] = d; mag[(R - 1) * C + j] = mag[(R - 2) * C + j]; }
    }
    /* border cells: in-bounds facets only, ascending (1800-1899) */
    for (i64 i = 0; i < R; i++)
        for (i64 j = 0; j < C; j++) {
            if (i > 0 && i < R - 1 && j > 0 && j < C - 1) { j = C - 2; continue; }
            for (int k = 0; k < 8; k++)
                facet(E, R, C, i, j, k, dX, dY, thA, thB, &mag[i * C + j], &dir[i * C + j]);
        }
    for (i64 n = 0; n < N; n++)
        if (mag[n] > 0) mag[n] = sqrt(mag[n]);                          /* 1901 */
    return 0;
}

/* ------------------------------------------------------------------------------------
 * a2: _find_flats_edges dem_processing.py:657-680 (+ get_adjacent_index utils.py:270-311).
 * flat = mag==-1, 8-connected regions numbered in raster order of their first cell
 * (scipy.ndimage.label), then for regions in label order: every 8-neighbour J of a
 * region cell gets flat[J] = (elev[J]==elev[first cell]).  The sequential overwrite means
 * the highest-numbered adjacent region decides.
 * ------------------------------------------------------------------------------------ */
int orc_flats(const double *E, const double *mag, i64 R, i64 C, u8 *flats)
{
    i64 N = R * C;
    int32_t *lab = (int32_t *)calloc((size_t)N, sizeof(int32_t));
    i64 *stack = (i64 *)malloc((size_t)N * sizeof(i64));
    i64 *first = NULL;
    i64 nlab = 0, cap = 0;
    if (!lab || !stack) { free(lab); free(stack); return 2; }
    for (i64 n = 0; n < N; n++) flats[n] = (mag[n] == -1);
    for (i64 n = 0; n < N; n++) {
        if (!flats[n] || lab[n]) continue;
        if (nlab == cap) { cap = cap ? cap * 2 : 1024; first = (i64 *)realloc(first, (size_t)cap * sizeof(i64)); }
        first[nlab++] = n;
        i64 sp = 0; stack[sp++] = n; lab[n] = (int32_t)nlab;
        while (sp) {
            i64 c = stack[--sp], ci = c / C, cj = c % C;
            for (int di = -1; di <= 1; di++) for (int dj = -1; dj <= 1; dj++) {
                i64 ni = ci + di, nj = cj + dj;
                if ((di || dj) && ni >= 0 && ni < R && nj >= 0 && nj < C) {
                    i64 m = ni * C + nj;
                    if (flats[m] && !lab[m]) { lab[m] = (int32_t)nlab; stack[sp++] = m; }
                }
            }
        }
    }
    for (i64 i = 0; i < R; i++) for (i64 j = 0; j < C; j++) {
        int32_t best = 0;
        for (int di = -1; di <= 1; di++) for (int dj = -1; dj <= 1; dj++) {
            i64 ni = i + di, nj = j + dj;
            if ((di || dj) && ni >= 0 && ni < R && nj >= 0 && nj < C) {
                int32_t l = lab[ni * C + nj];
                if (l > best) best = l;
            }
        }
        if (best) flats[i * C + j] = (E[i * C + j] == E[first[best - 1]]);
    }
    free(lab); free(stack); free(first);
    return 0;
}

/* ------------------------------------------------------------------------------------
 * a3: _calc_uca_section_proportion dem_processing.py:1021-1070.
 * theta per row: facet-0 theta of fences 0..R-3 padded with its first/last entry
 * (1031-1033) => row 0 -> thA[0], row i -> thA[i-1], row R-1 -> thA[R-3].
 * ------------------------------------------------------------------------------------ */
static double row_theta(const double *thA, i64 R, i64 i)
{
    if (i == 0) return thA[0];
    if (i == R - 1) return thA[R - 3];
    return thA[i - 1];
}

int orc_section_proportion(const double *dir, const u8 *flats, const double *thA, i64 R, i64 C,
                           int8_t *section, double *prop)
{
    static const int ADJ[8] = {1, -1, 1, -1, 1, -1, 1, -1};
    int bad = 0;
    for (i64 i = 0; i < R; i++) {
        double th = row_theta(thA, R, i);
        double cth = PI_ / 2 - th;
        for (i64 j = 0; j < C; j++) {
            i64 n = i * C + j;
            double d = dir[n];
            int q = (int)(int8_t)floor(d / PI_ * 2.0);                 /* 1035 */
            double quad = d - PI_ / 2.0 * (double)q;                   /* 1037 */
            int odd = ((q % 2) + 2) % 2;                               /* numpy modulo */
            int sec = q * 2 + ((quad > th) && !odd) + ((quad > cth) && odd); /* 1040-1043 */
            double p = NAN;
            int I1 = (sec == 0 || sec == 1 || sec == 4 || sec == 5);   /* 1050 */
            if (I1 && quad <= th) p = quad / th;
            if (I1 && quad > th) p = (quad - th) / (PI_ / 2 - th);
            if (!I1 && quad <= cth) p = quad / (PI_ / 2 - th);
            if (!I1 && quad > cth) p = (quad - (PI_ / 2 - th)) / th;
            if (flats[n]) { sec = -1; p = NAN; }                       /* 1064-1065 */
            if (sec == 8) sec = 0;                                     /* 1067 */
            int idx = sec < 0 ? sec + 8 : sec;                         /* python negative index */
            if (idx < 0 || idx > 7) { bad = 1; idx = 0; }
            double a = (double)ADJ[idx];
            section[n] = (int8_t)sec;
            prop[n] = (1 + a) / 2.0 - a * p;                           /* 1068 */
        }
    }
    return bad ? 3 : 0;
}

/* a4: _mk_connectivity dem_processing.py:1155-1267 -- receivers of facets[section];
 * the edge/corner tables there reduce to "set iff that neighbour is inside the grid". */
int orc_receivers(const int8_t *section, i64 R, i64 C, i64 *j1, i64 *j2)
{
    for (i64 i = 0; i < R; i++) for (i64 j = 0; j < C; j++) {
        i64 n = i * C + j; int s = section[n];
        j1[n] = -1; j2[n] = -1;
        if (s < 0 || s > 7) continue;
        i64 a = i + E1R[s], b = j + E1C[s];
        if (a >= 0 && a < R && b >= 0 && b < C) j1[n] = a * C + b;
        a = i + E2R[s]; b = j + E2C[s];
        if (a >= 0 && a < R && b >= 0 && b < C) j2[n] = a * C + b;
    }
    return 0;
}

/* numpy's pairwise float64 summation (numpy/_core/src/umath/loops_utils.h.src,
 * DOUBLE_pairwise_sum), which is what add.reduce applies to a contiguous float64 vector
 * (checked against np.sum in tests/test_oracle_golden.py); third-party
 * arithmetic that np.mean/np.sum in _mk_connectivity_pits (1346-1370) rely on. */
static double np_pairwise(const double *a, i64 n)
{
    if (n < 8) { double r = -0.0; for (i64 i = 0; i < n; i++) r += a[i]; return r; }
    if (n <= 128) {
        double r[8]; i64 i;
        for (int k = 0; k < 8; k++) r[k] = a[k];
        for (i = 8; i < n - (n % 8); i += 8) for (int k = 0; k < 8; k++) r[k] += a[i + k];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    }
    i64 n2 = n / 2; n2 -= n2 % 8;
    return np_pairwise(a, n2) + np_pairwise(a + n2, n - n2);
}
static double np_sum(const double *a, i64 n)
{
    if (n == 0) return 0.0;
    return np_pairwise(a, n);
}
double orc_np_sum(const double *a, i64 n) { return np_sum(a, n); }

/* ------------------------------------------------------------------------------------
 * a5: _mk_connectivity_pits dem_processing.py:1269-1382 (+ get_border_index utils.py:313-340,
 * _get_dX_mean 1993-1997, make_slice utils.py:404-408).
 * Output lists are malloc'ed; caller frees with orc_free.  flats/mag are updated in place
 * for drained pits (1370-1371).
 * ------------------------------------------------------------------------------------ */
typedef struct {
    i64 max_iter;       /* drain_pits_max_iter (300) */
    i64 max_dist;       /* drain_pits_max_dist (32); 0 = no index-distance filter */
    double max_dist_xy; /* drain_pits_max_dist_XY; <=0 / NaN = off */
    int min_border;     /* drain_pits_min_border */
} orc_pit_params;

static int cmp_i64(const void *a, const void *b)
{
    i64 x = *(const i64 *)a, y = *(const i64 *)b;
    return (x > y) - (x < y);
}

/* min with numpy semantics (NaN propagates) */
static double np_min(const double *a, i64 n)
{
    double m = a[0];
    if (m != m) return m;
    for (i64 i = 1; i < n; i++) { if (a[i] != a[i]) return a[i]; if (a[i] < m) m = a[i]; }
    return m;
}

void orc_free(void *p) { free(p); }

int orc_pits(const double *E, u8 *flats, double *mag, const double *dX, const double *dY,
             i64 R, i64 C, const orc_pit_params *pp,
             i64 *n_edges_out, i64 **pit_i_out, i64 **pit_j_out, double **pit_prop_out,
             i64 *n_warn_out)
{
    i64 N = R * C;
    u8 *pits_bool = (u8 *)malloc((size_t)N);
    int32_t *mark = (int32_t *)calloc((size_t)N, sizeof(int32_t)); /* 0 none, stamp*2 region, stamp*2+1 border */
    i64 ecap = 1024, ne = 0, nwarn = 0;
    i64 *pi = (i64 *)malloc((size_t)ecap * sizeof(i64)), *pj = (i64 *)malloc((size_t)ecap * sizeof(i64));
    double *pw = (double *)malloc((size_t)ecap * sizeof(double));
    i64 rcap = 4096, bcap = 4096;
    i64 *region = (i64 *)malloc((size_t)rcap * sizeof(i64));
    i64 *border = (i64 *)malloc((size_t)bcap * sizeof(i64));
    double *eb = (double *)malloc((size_t)bcap * sizeof(double));
    double *tmp = (double *)malloc((size_t)bcap * sizeof(double));
    int32_t stamp = 0;
    for (i64 n = 0; n < N; n++) pits_bool[n] = flats[n] && (E[n] > 0);  /* 1284 */

    for (i64 pit = 0; pit < N; pit++) {
        if (!pits_bool[pit]) continue;  /* result does not depend on the visiting order (1286-1287) */
        stamp++;
        i64 nr = 0, nb = 0, ndrain = 0;
        region[nr++] = pit; mark[pit] = stamp * 2;
        double epit = E[pit], epit_border = epit;
        int found = 0;
        i64 *drain = NULL;
        for (i64 it = -1; it < pp->max_iter; it++) {
            /* it == -1 is the pre-loop border of the bare pit (1292-1297), only needed for
               drain_pits_min_border */
            if (it == -1 && !pp->min_border) continue;
            /* border := 8-neighbours of the region not in it, ascending index (setdiff1d) */
            nb = 0;
            for (i64 t = 0; t < nr; t++) {
                i64 c = region[t], ci = c / C, cj = c % C;
                for (int di = -1; di <= 1; di++) for (int dj = -1; dj <= 1; dj++) {
                    i64 ni = ci + di, nj = cj + dj;
                    if (!(di || dj) || ni < 0 || ni >= R || nj < 0 || nj >= C) continue;
                    i64 m = ni * C + nj;
                    if (mark[m] == stamp * 2 || mark[m] == stamp * 2 + 1) continue;
                    mark[m] = stamp * 2 + 1;
                    if (nb == bcap) {
                        bcap *= 2;
                        border = (i64 *)realloc(border, (size_t)bcap * sizeof(i64));
                        eb = (double *)realloc(eb, (size_t)bcap * sizeof(double));
                        tmp = (double *)realloc(tmp, (size_t)bcap * sizeof(double));
                    }
                    border[nb++] = m;
                }
            }
            for (i64 t = 0; t < nb; t++) mark[border[t]] = 0; /* border is rebuilt every iteration */
            qsort(border, (size_t)nb, sizeof(i64), cmp_i64);
            for (i64 t = 0; t < nb; t++) eb[t] = E[border[t]];
            if (it == -1) { if (nb) epit_border = np_min(eb, nb); continue; }
            if (nb == 0) break;                                         /* 1304-1305 */
            double emin = np_min(eb, nb);                               /* 1307 */
            i64 n_np = 0, n_p = 0;
            for (i64 t = 0; t < nb; t++) if (!pits_bool[border[t]]) tmp[n_np++] = eb[t];
            if (n_np > 0 && np_min(tmp, n_np) < epit_border) {          /* 1312-1316 */
                drain = (i64 *)malloc((size_t)nb * sizeof(i64));
                for (i64 t = 0; t < nb; t++)
                    if (!pits_bool[border[t]] && eb[t] < epit_border) drain[ndrain++] = border[t];
                found = 1; break;
            }
            for (i64 t = 0; t < nb; t++) if (pits_bool[border[t]]) tmp[n_p++] = eb[t];
            if (n_p > 0 && np_min(tmp, n_p) < epit) {                   /* 1317-1320 */
                drain = (i64 *)malloc((size_t)nb * sizeof(i64));
                for (i64 t = 0; t < nb; t++)
                    if (pits_bool[border[t]] && eb[t] < epit) drain[ndrain++] = border[t];
                found = 1; break;
            }
            for (i64 t = 0; t < nb; t++)                                /* 1322-1323 */
                if (eb[t] == emin) {
                    if (nr == rcap) { rcap *= 2; region = (i64 *)realloc(region, (size_t)rcap * sizeof(i64)); }
                    region[nr++] = border[t]; mark[border[t]] = stamp * 2;
                }
        }
        for (i64 t = 0; t < nr; t++) mark[region[t]] = 0;
        if (stamp > 1000000000) { memset(mark, 0, (size_t)N * sizeof(int32_t)); stamp = 0; }
        if (!found) { nwarn++; continue; }                              /* 1327-1329 */

        i64 ip = pit / C, jp = pit % C;
        if (pp->max_dist > 0) {                                         /* 1335-1343 */
            i64 k = 0;
            for (i64 t = 0; t < ndrain; t++) {
                i64 di = ip - drain[t] / C, dj = jp - drain[t] % C;
                if (sqrt((double)(di * di + dj * dj)) <= (double)pp->max_dist) drain[k++] = drain[t];
            }
            ndrain = k;
            if (!ndrain) { nwarn++; free(drain); continue; }
        }
        double *dxy = (double *)malloc((size_t)ndrain * sizeof(double));
        double *s = (double *)malloc((size_t)ndrain * sizeof(double));
        for (i64 t = 0; t < ndrain; t++) {                              /* 1346-1349 */
            i64 id = drain[t] / C, jd = drain[t] % C;
            double dxm, dy;
            if (ip == id) { i64 f = ip < R - 2 ? ip : R - 2; dxm = dX[f]; dy = 0.0; }
            else {
                i64 lo = ip < id ? ip : id, hi = ip < id ? id : ip;
                dxm = np_sum(dX + lo, hi - lo) / (double)(hi - lo);
                dy = np_sum(dY + lo, hi - lo);
            }
            double dx = dxm * (double)(jp - jd);
            dxy[t] = sqrt(dx * dx + dy * dy);
        }
        if (pp->max_dist_xy > 0) {                                      /* 1352-1358 */
            i64 k = 0;
            for (i64 t = 0; t < ndrain; t++)
                if (dxy[t] <= pp->max_dist_xy) { drain[k] = drain[t]; dxy[k] = dxy[t]; k++; }
            ndrain = k;
            if (!ndrain) { nwarn++; free(drain); free(dxy); free(s); continue; }
        }
        for (i64 t = 0; t < ndrain; t++) s[t] = fabs(E[pit] - E[drain[t]]) / dxy[t]; /* 1361 */
        double ssum = np_sum(s, ndrain);
        for (i64 t = 0; t < ndrain; t++) {                              /* 1365-1367 */
            if (ne == ecap) {
                ecap *= 2;
                pi = (i64 *)realloc(pi, (size_t)ecap * sizeof(i64));
                pj = (i64 *)realloc(pj, (size_t)ecap * sizeof(i64));
                pw = (double *)realloc(pw, (size_t)ecap * sizeof(double));
            }
            pi[ne] = pit; pj[ne] = drain[t]; pw[ne] = s[t] / ssum; ne++;
        }
        mag[pit] = ssum / (double)ndrain;                               /* 1370: np.mean */
        flats[pit] = 0;                                                 /* 1371 */
        free(drain); free(dxy); free(s);
    }
    free(pits_bool); free(mark); free(region); free(border); free(eb); free(tmp);
    *n_edges_out = ne; *pit_i_out = pi; *pit_j_out = pj; *pit_prop_out = pw; *n_warn_out = nwarn;
    return 0;
}

/* ------------------------------------------------------------------------------------
 * a4 (matrix part): _mk_adjacency_matrix dem_processing.py:1072-1153.
 * Edge list = [(i,j1[i],prop[i]) for all i] ++ [(i,j2[i],1-prop[i])] ++ pit edges, with the
 * pits' own j1/j2 crossed out (1099-1100), filtered by 1136-1137, stored as
 * CSC (column = source) and CSR (row = receiver) like scipy (1147, 879).
 * ------------------------------------------------------------------------------------ */
typedef struct {
    i64 N, nnz;
    i64 *cptr, *cidx; double *cdat;  /* CSC: receivers of source c */
    i64 *rptr, *ridx;                /* CSR: sources of receiver r */
} orc_graph;

orc_graph *orc_graph_build(const double *E, i64 N, const i64 *j1, const i64 *j2, const double *prop,
                           i64 npit, const i64 *pit_i, const i64 *pit_j, const double *pit_prop)
{
    orc_graph *g = (orc_graph *)calloc(1, sizeof(orc_graph));
    g->N = N;
    u8 *crossed = (u8 *)calloc((size_t)N, 1);
    for (i64 t = 0; t < npit; t++) crossed[pit_i[t]] = 1;
    g->cptr = (i64 *)calloc((size_t)N + 1, sizeof(i64));
    g->rptr = (i64 *)calloc((size_t)N + 1, sizeof(i64));
#define KEEP(i, j, w) (!((w) != (w)) && (j) != -1 && (w) > 1e-8 && E[(j)] <= E[(i)])
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) {
            for (i64 n = 0; n < N; n++) { g->cptr[n + 1] += g->cptr[n]; g->rptr[n + 1] += g->rptr[n]; }
            g->nnz = g->cptr[N];
            g->cidx = (i64 *)malloc((size_t)(g->nnz + 1) * sizeof(i64));
            g->cdat = (double *)malloc((size_t)(g->nnz + 1) * sizeof(double));
            g->ridx = (i64 *)malloc((size_t)(g->nnz + 1) * sizeof(i64));
        }
        /* pass 0 counts into ptr[n+1]; pass 1 fills using running cursors */
        i64 *ccur = NULL, *rcur = NULL;
        if (pass == 1) {
            ccur = (i64 *)malloc((size_t)N * sizeof(i64)); rcur = (i64 *)malloc((size_t)N * sizeof(i64));
            memcpy(ccur, g->cptr, (size_t)N * sizeof(i64)); memcpy(rcur, g->rptr, (size_t)N * sizeof(i64));
        }
        for (int which = 0; which < 3; which++) {
            i64 cnt = which < 2 ? N : npit;
            for (i64 t = 0; t < cnt; t++) {
                i64 i, j; double w;
                if (which == 0) { i = t; j = crossed[t] ? -1 : j1[t]; w = prop[t]; }
                else if (which == 1) { i = t; j = crossed[t] ? -1 : j2[t]; w = 1 - prop[t]; }
                else { i = pit_i[t]; j = pit_j[t]; w = pit_prop[t]; }
                if (!KEEP(i, j, w)) continue;
                if (pass == 0) { g->cptr[i + 1]++; g->rptr[j + 1]++; }
                else {
                    g->cidx[ccur[i]] = j; g->cdat[ccur[i]] = w; ccur[i]++;
                    g->ridx[rcur[j]++] = i;
                }
            }
        }
        free(ccur); free(rcur);
    }
#undef KEEP
    free(crossed);
    return g;
}
i64 orc_graph_nnz(const orc_graph *g) { return g->nnz; }
void orc_graph_export(const orc_graph *g, i64 *cptr, i64 *cidx, double *cdat, i64 *rptr, i64 *ridx)
{
    memcpy(cptr, g->cptr, (size_t)(g->N + 1) * sizeof(i64));
    memcpy(rptr, g->rptr, (size_t)(g->N + 1) * sizeof(i64));
    memcpy(cidx, g->cidx, (size_t)g->nnz * sizeof(i64));
    memcpy(cdat, g->cdat, (size_t)g->nnz * sizeof(double));
    memcpy(ridx, g->ridx, (size_t)g->nnz * sizeof(i64));
}
/* per-cell sums used by _calc_uca_chunk 882, 913-930: inflow (row sum) and outflow (column sum) */
void orc_graph_sums(const orc_graph *g, double *inflow, double *outflow, i64 *indeg)
{
    for (i64 n = 0; n < g->N; n++) { inflow[n] = 0; outflow[n] = 0; indeg[n] = g->rptr[n + 1] - g->rptr[n]; }
    for (i64 c = 0; c < g->N; c++)
        for (i64 t = g->cptr[c]; t < g->cptr[c + 1]; t++) { outflow[c] += g->cdat[t]; inflow[g->cidx[t]] += g->cdat[t]; }
}
void orc_graph_free(orc_graph *g)
{
    if (!g) return;
    free(g->cptr); free(g->cidx); free(g->cdat); free(g->rptr); free(g->ridx); free(g);
}

/* _check_id_on_edge cyutils.pyx:207-226 */
static int on_edge(i64 r, i64 R, i64 C)
{
    return r < C || r >= C * R - C || (r % C) == 0 || (r % C) == C - 1;
}

/* ------------------------------------------------------------------------------------
 * a7: cyutils._drain_area cyutils.pyx:119-187.  Same rounds, same visiting order
 * (ascending source index inside a round), same readiness test (all CSR sources done),
 * same border skip rule; the per-round full-array scans are replaced by sorted frontier
 * lists (cost O(edges) instead of O(cells x rounds)).  Returns the number of rounds.
 * ------------------------------------------------------------------------------------ */
i64 orc_drain_area(const orc_graph *g, double *area, u8 *done, const u8 *ids0, i64 R, i64 C,
                   double *edge_todo /* or NULL */, int skip_edge, i64 *n_drained_out)
{
    i64 N = g->N, ncur = 0, nnext = 0, rounds = 0, drained = 0;
    i64 *cur = (i64 *)malloc((size_t)N * sizeof(i64)), *next = (i64 *)malloc((size_t)N * sizeof(i64));
    u8 *innext = (u8 *)calloc((size_t)N, 1);
    for (i64 n = 0; n < N; n++) if (ids0[n]) cur[ncur++] = n;
    for (;;) {
        for (i64 t = 0; t < ncur; t++) done[cur[t]] = 1;               /* 138-140 */
        nnext = 0;
        for (i64 t = 0; t < ncur; t++) {
            i64 i = cur[t];
            for (i64 e = g->cptr[i]; e < g->cptr[i + 1]; e++) {
                i64 r = g->cidx[e]; double f = g->cdat[e];
                if ((skip_edge || done[r]) && on_edge(r, R, C)) continue; /* 157-159 */
                area[r] += area[i] * f;                                 /* 161 */
                if (edge_todo) edge_todo[r] += edge_todo[i] * f;        /* 163-164 */
                int wait = 0;
                for (i64 k = g->rptr[r]; k < g->rptr[r + 1]; k++)
                    if (done[g->ridx[k]] < 1) { wait = 1; break; }      /* 171-177 */
                if (!wait && !innext[r]) { innext[r] = 1; next[nnext++] = r; }
            }
        }
        drained += ncur; rounds++;
        qsort(next, (size_t)nnext, sizeof(i64), cmp_i64);
        int same = (nnext == ncur);
        if (same) for (i64 t = 0; t < ncur; t++) if (cur[t] != next[t]) { same = 0; break; }
        for (i64 t = 0; t < nnext; t++) innext[next[t]] = 0;
        i64 *sw = cur; cur = next; next = sw; ncur = nnext;
        if (same) break;                                                /* 187 */
    }
    free(cur); free(next); free(innext);
    if (n_drained_out) *n_drained_out = drained;
    return rounds;
}

/* a8: cyutils._drain_connections cyutils.pyx:49-72 */
i64 orc_drain_connections(const orc_graph *g, u8 *arr, const u8 *ids0, int set_to)
{
    i64 N = g->N, ncur = 0, nnext = 0, rounds = 0;
    i64 *cur = (i64 *)malloc((size_t)N * sizeof(i64)), *next = (i64 *)malloc((size_t)N * sizeof(i64));
    u8 *innext = (u8 *)calloc((size_t)N, 1);
    u8 tf = (u8)set_to;
    for (i64 n = 0; n < N; n++) if (ids0[n]) cur[ncur++] = n;
    for (;;) {
        nnext = 0;
        for (i64 t = 0; t < ncur; t++) {
            i64 i = cur[t];
            for (i64 e = g->cptr[i]; e < g->cptr[i + 1]; e++) {
                i64 r = g->cidx[e];
                if (arr[r] != tf && !innext[r]) { innext[r] = 1; next[nnext++] = r; }
                arr[r] = tf;
            }
        }
        rounds++;
        qsort(next, (size_t)nnext, sizeof(i64), cmp_i64);
        int same = (nnext == ncur);
        if (same) for (i64 t = 0; t < ncur; t++) if (cur[t] != next[t]) { same = 0; break; }
        for (i64 t = 0; t < nnext; t++) innext[next[t]] = 0;
        i64 *sw = cur; cur = next; next = sw; ncur = nnext;
        if (same) break;
    }
    free(cur); free(next); free(innext);
    return rounds;
}

/* a9: calc_twi dem_processing.py:1647-1677 (returns the un-scaled twi; caller stores 10*twi) */
int orc_twi(const double *uca, const double *mag, i64 N, double min_slope, double min_area,
            double sat_limit, int limit_uca, int limit_twi, double *twi)
{
    double cap = sat_limit * min_area;
    double twi_sat = log(sat_limit * min_area / min_slope);
    for (i64 n = 0; n < N; n++) {
        double u = uca[n];
        if (limit_uca && u > cap) u = cap;
        double t = log(u / (mag[n] + min_slope));
        if (limit_twi && t > twi_sat) t = twi_sat;
        twi[n] = t;
    }
    return 0;
}
