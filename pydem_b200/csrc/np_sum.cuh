// np_sum.cuh -- NumPy's float64 summation order, restated for the device.
#pragma once
#include <stdint.h>

// numpy pairwise summation (numpy/_core/src/umath/loops_utils.h.src): what np.sum / np.mean
// apply to the contiguous float64 vectors at 1346-1370.
static __device__ double np_pairwise(const double *a, int64_t n)
{
    if (n < 8) {
        double r = -0.0;
        for (int64_t i = 0; i < n; i++) r = __dadd_rn(r, a[i]);
        return r;
    }
    if (n <= 128) {
        double r[8];
        int64_t i;
        for (int k = 0; k < 8; k++) r[k] = a[k];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; k++) r[k] = __dadd_rn(r[k], a[i + k]);
        double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                               __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (; i < n; i++) res = __dadd_rn(res, a[i]);
        return res;
    }
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    return __dadd_rn(np_pairwise(a, n2), np_pairwise(a + n2, n - n2));
}
static __device__ double np_sum(const double *a, int64_t n)
{
    if (n == 0) return 0.0;
    return np_pairwise(a, n);
}

