/* priority_flood.c -- synthetic-data helper (host): hydrological conditioning of a test DEM with
 * priority-flood + epsilon (Barnes, Lehman & Mulla 2014): every cell ends up with a strictly
 * descending 8-connected path to the border, so the conditioned fractal has long rivers and no
 * interior pits or flats -- the kind of input pyDEM sees after its own conditioning stage.
 * Used by pydem_b200.synth.conditioned_fractal_dem for tests and benchmarks only. */
#include <stdint.h>
#include <stdlib.h>

typedef struct { double e; int64_t i; } item;

static void push(item *h, int64_t *n, item x)
{
    int64_t k = (*n)++;
    while (k > 0) {
        int64_t p = (k - 1) >> 1;
        if (h[p].e < x.e || (h[p].e == x.e && h[p].i <= x.i)) break;
        h[k] = h[p]; k = p;
    }
    h[k] = x;
}

static item pop(item *h, int64_t *n)
{
    item top = h[0], x = h[--(*n)];
    int64_t k = 0, m = *n;
    for (;;) {
        int64_t c = 2 * k + 1;
        if (c >= m) break;
        if (c + 1 < m && (h[c + 1].e < h[c].e || (h[c + 1].e == h[c].e && h[c + 1].i < h[c].i))) c++;
        if (x.e < h[c].e || (x.e == h[c].e && x.i <= h[c].i)) break;
        h[k] = h[c]; k = c;
    }
    h[k] = x;
    return top;
}

/* in place; returns the number of raised cells, -1 on allocation failure.
 * wrap_rows != 0: rows wrap around (row R-1 is adjacent to row 0) and only the left/right columns
 * are outlets -- the conditioned block can then be stacked vertically any number of times into one
 * seamless, already conditioned DEM (row-sharded weak-scaling runs). */
int64_t pdm_priority_flood_eps2(double *E, int64_t R, int64_t C, double eps, int wrap_rows)
{
    int64_t N = R * C, nh = 0, raised = 0;
    item *h = (item *)malloc((size_t)N * sizeof(item));
    uint8_t *closed = (uint8_t *)calloc((size_t)N, 1);
    if (!h || !closed) { free(h); free(closed); return -1; }
    for (int64_t i = 0; i < R; i++)
        for (int64_t j = 0; j < C; j++)
            if (j == 0 || j == C - 1 || (!wrap_rows && (i == 0 || i == R - 1))) {
                item x = {E[i * C + j], i * C + j};
                closed[i * C + j] = 1; push(h, &nh, x);
            }
    while (nh) {
        item c = pop(h, &nh);
        int64_t ci = c.i / C, cj = c.i % C;
        for (int di = -1; di <= 1; di++)
            for (int dj = -1; dj <= 1; dj++) {
                int64_t ni = ci + di, nj = cj + dj;
                if (wrap_rows) ni = (ni + R) % R;
                if ((!di && !dj) || ni < 0 || ni >= R || nj < 0 || nj >= C) continue;
                int64_t n = ni * C + nj;
                if (closed[n]) continue;
                closed[n] = 1;
                if (E[n] < c.e + eps) { E[n] = c.e + eps; raised++; }
                item x = {E[n], n};
                push(h, &nh, x);
            }
    }
    free(h); free(closed);
    return raised;
}

int64_t pdm_priority_flood_eps(double *E, int64_t R, int64_t C, double eps)
{
    return pdm_priority_flood_eps2(E, R, C, eps, 0);
}
