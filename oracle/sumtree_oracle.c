/*
 * oracle/sumtree_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded CPU restatement of the five routines of the
 * reference kernel module `starr._cython`
 * (reference: starr/src/_sumtree_array.pyx).  It exists so that the CUDA path can be
 * checked bit-for-bit on any machine (the GPU box has no /root/reference) and so that
 * bench.py can time a CPU baseline.  Nothing under starr_b200/ may import, link or
 * call it: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do.
 *
 * Parity of this restatement is PINNED: tests/test_oracle.py checks it against the
 * reference's own golden vectors (tests/test_cython_sumtree_array.py:11-79, README
 * doc values) and, bit-for-bit on random inputs, against the unmodified reference
 * module compiled into oracle/_ref/ by oracle/build_ref.sh.
 *
 * Semantics restated (all arithmetic in the element type T, in this exact order):
 *   update : _sumtree_array.pyx:22-49   -- ordered per-update delta walk to the root
 *   build  : _sumtree_array.pyx:51-64   -- node = left + right, i = n-1 .. 1
 *   search : _sumtree_array.pyx:66-122  -- left-child descent, `>=` goes right
 *   range  : _sumtree_array.pyx:131-170 -- covering node minus outside siblings
 *   strided: _sumtree_array.pyx:176-194 -- one range sum per stride block
 * Layout: array[i] = leaf i, sumtree[h] = heap node h in [1, n), sumtree[0] unused (0);
 * child index c >= n means leaf c - n (_sumtree_array.pyx:15-20).
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/Makefile).
 * Return codes: 0 ok, 1 negative value, 2 index out of range, 3 empty vals.
 */
#include <stddef.h>
#include <stdint.h>

#define ORC_OK 0
#define ORC_NEGATIVE 1
#define ORC_INDEX 2
#define ORC_EMPTY_VALS 3

/* child lookup across the (sumtree | array) pair, _sumtree_array.pyx:15-20 */
#define NODE(tree, leaves, n, c) ((c) >= (n) ? (leaves)[(c) - (n)] : (tree)[(c)])

#define DEFINE_ORACLE(SUF, T, IS_SIGNED)                                                   \
                                                                                           \
    /* _sumtree_array.pyx:22-49.  Negative scan over ALL vals first (:37-39), then one     \
     * delta walk per update in batch order (:41-49); vals cycles with i % nv (:44). */    \
    int orc_update_##SUF(const int64_t *idxs, int64_t ni, const T *vals, int64_t nv,       \
                         T *leaves, T *tree, int64_t n)                                    \
    {                                                                                      \
        if (IS_SIGNED)                                                                     \
            for (int64_t i = 0; i < nv; ++i)                                               \
                if (vals[i] < (T)0) return ORC_NEGATIVE;                                   \
        if (ni > 0 && nv == 0) return ORC_EMPTY_VALS;                                      \
        for (int64_t i = 0; i < ni; ++i) {                                                 \
            int64_t leaf = idxs[i];                                                        \
            if (leaf < 0 || leaf >= n) return ORC_INDEX;                                   \
            T delta = (T)(vals[i % nv] - leaves[leaf]);                                    \
            leaves[leaf] = (T)(leaves[leaf] + delta);                                      \
            for (int64_t h = (leaf + n) / 2; h > 0; h /= 2)                                \
                tree[h] = (T)(tree[h] + delta);                                            \
        }                                                                                  \
        return ORC_OK;                                                                     \
    }                                                                                      \
                                                                                           \
    /* _sumtree_array.pyx:51-64.  min(array) < 0 rejected before any write (:58-59). */    \
    int orc_build_##SUF(const T *leaves, T *tree, int64_t n)                               \
    {                                                                                      \
        if (IS_SIGNED)                                                                     \
            for (int64_t i = 0; i < n; ++i)                                                \
                if (leaves[i] < (T)0) return ORC_NEGATIVE;                                 \
        for (int64_t h = n - 1; h >= 1; --h) {                                             \
            T l = NODE(tree, leaves, n, 2 * h);                                            \
            T r = NODE(tree, leaves, n, 2 * h + 1);                                        \
            tree[h] = (T)(l + r);                                                          \
        }                                                                                  \
        return ORC_OK;                                                                     \
    }                                                                                      \
                                                                                           \
    /* _sumtree_array.pyx:84-120.  Only the LEFT child is read at each step (:96-104);     \
     * `prefix >= left` descends right and subtracts (:107-113). */                        \
    void orc_search_##SUF(int64_t *out, const T *vals, int64_t m, const T *leaves,         \
                          const T *tree, int64_t n)                                        \
    {                                                                                      \
        for (int64_t q = 0; q < m; ++q) {                                                  \
            T prefix = vals[q];                                                            \
            int64_t h = 1;                                                                 \
            while (h < n) {                                                                \
                int64_t c = 2 * h;                                                         \
                T lv = NODE(tree, leaves, n, c);                                           \
                if (prefix >= lv) {                                                        \
                    prefix = (T)(prefix - lv);                                             \
                    h = c + 1;                                                             \
                } else {                                                                   \
                    h = c;                                                                 \
                }                                                                          \
            }                                                                              \
            out[q] = h - n;                                                                \
        }                                                                                  \
    }                                                                                      \
                                                                                           \
    /* _sumtree_array.pyx:131-170: sum of leaves [a, b) as (covering node) - (siblings     \
     * outside the range), accumulated a-side then b-side per level; the `wrapped`         \
     * branch (:146-155, :167-168) handles ranges whose ends sit at different depths. */   \
    T orc_sum_over_##SUF(const T *leaves, const T *tree, int64_t n, int64_t a, int64_t b)  \
    {                                                                                      \
        b -= 1;                                                                            \
        if (a == b) return leaves[a];                                                      \
        if (a > b) return (T)0;                                                            \
        a += n;                                                                            \
        b += n;                                                                            \
        T sub = (T)0;                                                                      \
        int wrapped = a < (a ^ b);                                                         \
        if (wrapped) {                                                                     \
            if ((b & 1) == 0) sub = (T)(sub + NODE(tree, leaves, n, b + 1));               \
            b /= 2;                                                                        \
        }                                                                                  \
        while ((a / 2) != (b / 2)) {                                                       \
            if ((a & 1) == 1) sub = (T)(sub + NODE(tree, leaves, n, a - 1));               \
            if ((b & 1) == 0) sub = (T)(sub + NODE(tree, leaves, n, b + 1));               \
            a /= 2;                                                                        \
            b /= 2;                                                                        \
        }                                                                                  \
        if (wrapped) return (T)(tree[1] - sub);                                            \
        return (T)(NODE(tree, leaves, n, a / 2) - sub);                                    \
    }                                                                                      \
                                                                                           \
    /* _sumtree_array.pyx:176-194: out[i] = sum_over(stride*i, min(stride*(i+1), n)),      \
     * nout = ceil(n / stride). */                                                         \
    void orc_strided_sum_##SUF(const T *leaves, const T *tree, int64_t n, int64_t stride,  \
                               T *out, int64_t nout)                                       \
    {                                                                                      \
        for (int64_t i = 0; i < nout; ++i) {                                               \
            int64_t hi = stride * (i + 1);                                                 \
            if (hi > n) hi = n;                                                            \
            out[i] = orc_sum_over_##SUF(leaves, tree, n, stride * i, hi);                  \
        }                                                                                  \
    }

DEFINE_ORACLE(i16, int16_t, 1)
DEFINE_ORACLE(i32, int32_t, 1)
DEFINE_ORACLE(i64, int64_t, 1)
DEFINE_ORACLE(u8, uint8_t, 0)
DEFINE_ORACLE(u16, uint16_t, 0)
DEFINE_ORACLE(u32, uint32_t, 0)
DEFINE_ORACLE(u64, uint64_t, 0)
DEFINE_ORACLE(f32, float, 1)
DEFINE_ORACLE(f64, double, 1)

/*
 * Sequential row-order fold used by the reference's non-contiguous `sum` fallback
 * (sumtree_array.py:514 -> ndarray.sum over a leading/middle axis).  NumPy adds row r
 * into the output accumulators for r = 0..R-1, i.e. out[a,c] = ((x[a,0,c] + x[a,1,c]) +
 * ...); small integer dtypes are widened to 64 bit first (NumPy's default sum dtype).
 */
#define DEFINE_FOLD(SUF, T, ACC)                                                           \
    void orc_axis_fold_##SUF(const T *x, int64_t A, int64_t R, int64_t C, ACC *out)        \
    {                                                                                      \
        for (int64_t a = 0; a < A; ++a)                                                    \
            for (int64_t c = 0; c < C; ++c) {                                              \
                ACC acc = (ACC)0; /* add.reduce starts from the identity: an all -0.0 column gives +0.0 */ \
                for (int64_t r = 0; r < R; ++r) acc = (ACC)(acc + (ACC)x[(a * R + r) * C + c]); \
                out[a * C + c] = acc;                                                      \
            }                                                                              \
    }

DEFINE_FOLD(i16, int16_t, int64_t)
DEFINE_FOLD(i32, int32_t, int64_t)
DEFINE_FOLD(i64, int64_t, int64_t)
DEFINE_FOLD(u8, uint8_t, uint64_t)
DEFINE_FOLD(u16, uint16_t, uint64_t)
DEFINE_FOLD(u32, uint32_t, uint64_t)
DEFINE_FOLD(u64, uint64_t, uint64_t)
DEFINE_FOLD(f32, float, float)
DEFINE_FOLD(f64, double, double)

/* ------------------------------------------------------------------------------------------
 * starr.experimental: the C++ twin, starr/experimental/src/sumtree_array.cpp.
 * float64 values, int32 indices.  Differences from orc_update_f64 that are restated here:
 * the leaf is ASSIGNED (`*array = val`, cpp:23) and the negative check is per element inside
 * the loop (cpp:15-17) -- updates in front of a negative value stay applied (the reference then
 * throws through a nogil extern and the process aborts; this restatement returns ORC_NEGATIVE).
 * `vals` cycles by a wrap-around cursor (cpp:41-50), the same as i % V.
 * Pinned against the compiled reference module in oracle/_ref (tests/test_oracle.py). */
int orc_exp_update_tree_multi(const int32_t *idxs, int32_t I, const double *vals, int32_t V,
                              double *array, int32_t n, double *sumtree)
{
    int32_t v = 0;
    for (int32_t i = 0; i < I; ++i) {
        const double val = vals[v];
        if (val < 0) return ORC_NEGATIVE;                 /* cpp:15-17 */
        int32_t idx = idxs[i];
        const double diff = val - array[idx];            /* cpp:21 */
        array[idx] = val;                                /* cpp:22-23 */
        for (idx = (idx + n) / 2; idx > 0; idx /= 2)     /* cpp:20, 25-29 */
            sumtree[idx] += diff;
        if (++v == V) v = 0;                             /* cpp:46-50 */
    }
    return ORC_OK;
}

/* cpp:54-85: the same left-child descent as orc_search_f64 with int32 outputs */
void orc_exp_get_prefix_sum_idx_multi(int32_t *out, const double *vals, int32_t m, const double *array,
                                      int32_t n, const double *sumtree)
{
    for (int32_t q = 0; q < m; ++q) {
        double val = vals[q];
        int32_t i = 1;
        while (i < n) {
            const int32_t left = 2 * i;
            const double left_val = left < n ? sumtree[left] : array[left - n];
            if (val >= left_val) {
                i = left + 1;
                val -= left_val;
            } else {
                i = left;
            }
        }
        out[q] = i - n;
    }
}

/* ------------------------------------------------------------------------------------------
 * MinTree: starr/experimental/slow.py:53-78.  data[2n], data[idx + n] = val, then
 * every ancestor = min(left, right) with Python's min (returns the right operand only if it is
 * smaller).  One call = a batch of sequential __setitem__ (slow.py:62-68). */
void orc_min_set_f64(double *data, int64_t n, const int64_t *idxs, int64_t ni, const double *vals, int64_t nv)
{
    for (int64_t i = 0; i < ni; ++i) {
        int64_t idx = idxs[i] + n;
        data[idx] = vals[i % nv];
        for (idx /= 2; idx > 0; idx /= 2) {
            const double l = data[2 * idx], r = data[2 * idx + 1];
            data[idx] = r < l ? r : l;
        }
    }
}
