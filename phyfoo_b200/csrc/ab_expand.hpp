// ab_expand.hpp -- table layout and gap expansion of the alignment-based models of order >= 1 (see ab.cuh); host and device,
// so that tests/emul can run the same lines on the CPU against a statement-by-statement restatement of the Java.
#pragma once
#if defined(__CUDACC__)
#define AB_HD __host__ __device__ inline
#else
#define AB_HD inline
#endif

#define AB_MAX_ORDER 2
#define AB_MAX_EXPAND 100                                  // 4 + 16 + 80: the background list after three gaps

AB_HD int ab_size(int p, int order) { return 4 << (2 * (p < order ? p : order)); }
AB_HD int ab_off(int p, int order) {
    int o = 0;
    for (int i = 0; i < p; ++i) o += ab_size(i, order);
    return o;
}

// sym[k] = code of column l - k (4 = unobserved), k = 0..kmax.  variant 0: the motif's table and both training loops (a gap
// copies the entries of the previous group), 1: the background's likelihood (a gap copies every earlier entry).
// Result: hashes idx[pre..n).
AB_HD void ab_expand(const int* sym, int kmax, int variant, int* idx, int& pre, int& n) {
    n = 0; pre = 0;
    int pw = 1;
    for (int k = 0; k <= kmax; ++k, pw *= 4) {
        if (sym[k] == 4) {
            const int old_pre = pre;
            pre = n;
            if (n == 0) { for (int a = 0; a < 4; ++a) idx[n++] = a * pw; }
            else {
                const int lo = variant ? 0 : old_pre, hi = pre;
                for (int i = lo; i < hi; ++i)
                    for (int a = 0; a < 4; ++a) idx[n++] = idx[i] + a * pw;
            }
        } else {
            if (n == 0) { idx[0] = 0; n = 1; }
            // motif: the current group only (:208-212); background: every entry (:199-201) -- the entries before `pre` matter
            // only as copy sources of variant 1
            for (int i = variant ? 0 : pre; i < n; ++i) idx[i] += sym[k] * pw;
        }
    }
}

