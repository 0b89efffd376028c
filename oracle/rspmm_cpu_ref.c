/* CPU restatement of torchdrug's compiled rspmm path.  TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * The reference's hot path calls `functional.generalized_rspmm` (reference ultra/layer.py:134-167,
 * 336-369), whose arithmetic is torchdrug's C++ extension
 * `torchdrug/layers/functional/extension/rspmm.{h,cpp}` (torchdrug>=0.2.1, reference
 * requirements.txt:3).  That source is NOT under /root/reference and the package is not installed,
 * so it cannot be compiled here; this file restates its published algorithm:
 *
 *   forward : parallel over destination rows of the CSR built from the coalesced COO; per row
 *             out = zero; for each edge (col, layer, val) in coalesced order, for each feature d:
 *             out[d] = Nary::forward(out[d], val * Binary::forward(relation[layer,d], input[col,d]))
 *   backward: same loop nest; per (edge, d) recompute y, gate = Nary::backward(out, y),
 *             relation_grad[layer,d] += g*gate*val*dlhs   and   input_grad[col,d] += g*gate*val*drhs.
 *             torchdrug serialises those two updates with one std::mutex per element; this restatement
 *             uses an OpenMP atomic add per element instead (same result set, cheaper - i.e. a
 *             GENEROUS stand-in when timed as the CPU baseline).
 *
 * Semantics mirrored in the math: reference ultra/layer.py:52-109 (message + aggregate).
 * "kind": "port" in bench.py's cpu_baseline - parity of this file is pinned by tests/test_oracle.py
 * against oracle/rspmm_oracle.py and the golden vectors in tests/golden/.
 *
 * Build: see oracle/Makefile (gcc -O3 -fopenmp -ffp-contract=off; no fast-math so that max/min
 * messages are bit-faithful IEEE fp32).
 */
#include <float.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { SUM_ADD = 0, SUM_MIN = 1, SUM_MAX = 2 };
enum { MUL_MUL = 0, MUL_ADD = 1 };

int rspmm_ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void rspmm_ref_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* coo2csr3d: `row` must be sorted (coalesced).  row_ptr has n_out + 1 entries. */
int rspmm_ref_coo2csr(int64_t n_out, int64_t nnz, const int64_t *row, int64_t *row_ptr) {
    memset(row_ptr, 0, (size_t)(n_out + 1) * sizeof(int64_t));
    for (int64_t e = 0; e < nnz; ++e) {
        if (row[e] < 0 || row[e] >= n_out) return 1;
        if (e && row[e] < row[e - 1]) return 2;
        row_ptr[row[e] + 1] += 1;
    }
    for (int64_t i = 0; i < n_out; ++i) row_ptr[i + 1] += row_ptr[i];
    return 0;
}

static inline float nary_zero(int sum_op) {
    return sum_op == SUM_ADD ? 0.0f : (sum_op == SUM_MAX ? -FLT_MAX : FLT_MAX);
}

int rspmm_ref_forward_f32(int64_t n_out, int64_t dim, const int64_t *row_ptr, const int64_t *col,
                          const int64_t *layer, const float *val, const float *relation,
                          const float *input, float *output, int sum_op, int mul_op) {
    if (sum_op < 0 || sum_op > 2 || mul_op < 0 || mul_op > 1) return 1;
    const float zero = nary_zero(sum_op);
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n_out; ++i) {
        float *out = output + i * dim;
        for (int64_t d = 0; d < dim; ++d) out[d] = zero;
        for (int64_t p = row_ptr[i]; p < row_ptr[i + 1]; ++p) {
            const float *rel = relation + layer[p] * dim;
            const float *in = input + col[p] * dim;
            const float w = val[p];
            if (sum_op == SUM_ADD && mul_op == MUL_MUL) {
                for (int64_t d = 0; d < dim; ++d) out[d] = out[d] + w * (rel[d] * in[d]);
            } else if (sum_op == SUM_ADD) {
                for (int64_t d = 0; d < dim; ++d) out[d] = out[d] + w * (rel[d] + in[d]);
            } else {
                for (int64_t d = 0; d < dim; ++d) {
                    const float x = mul_op == MUL_MUL ? rel[d] * in[d] : rel[d] + in[d];
                    const float y = w * x;
                    if (sum_op == SUM_MAX) out[d] = out[d] > y ? out[d] : y;
                    else out[d] = out[d] < y ? out[d] : y;
                }
            }
        }
    }
    return 0;
}

int rspmm_ref_backward_f32(int64_t n_out, int64_t n_in, int64_t n_rel, int64_t dim,
                           const int64_t *row_ptr, const int64_t *col, const int64_t *layer,
                           const float *val, const float *relation, const float *input,
                           const float *output, const float *grad_output, float *grad_relation,
                           float *grad_input, int sum_op, int mul_op) {
    if (sum_op < 0 || sum_op > 2 || mul_op < 0 || mul_op > 1) return 1;
    memset(grad_relation, 0, (size_t)(n_rel * dim) * sizeof(float));
    memset(grad_input, 0, (size_t)(n_in * dim) * sizeof(float));
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n_out; ++i) {
        const float *g = grad_output + i * dim;
        const float *out = output + i * dim;
        for (int64_t p = row_ptr[i]; p < row_ptr[i + 1]; ++p) {
            const float *rel = relation + layer[p] * dim;
            const float *in = input + col[p] * dim;
            float *g_rel = grad_relation + layer[p] * dim;
            float *g_in = grad_input + col[p] * dim;
            const float w = val[p];
            for (int64_t d = 0; d < dim; ++d) {
                float gate = 1.0f;
                if (sum_op != SUM_ADD) {
                    const float x = mul_op == MUL_MUL ? rel[d] * in[d] : rel[d] + in[d];
                    gate = (out[d] == w * x) ? 1.0f : 0.0f;
                }
                const float up = g[d] * gate * w;
                const float to_rel = mul_op == MUL_MUL ? up * in[d] : up;
                const float to_in = mul_op == MUL_MUL ? up * rel[d] : up;
#pragma omp atomic
                g_rel[d] += to_rel;
#pragma omp atomic
                g_in[d] += to_in;
            }
        }
    }
    return 0;
}

/* ---- variants used by the full-size parity tests (tests/test_full_size_gpu.py) -------------------
 * Same loop nest as above, evaluated on a column subsample of the operands:
 *   rspmm_ref_forward_arg_f32 : min/max forward that also records the arg-index = position (in coalesced
 *                               order) of the FIRST edge attaining the extremum, -1 for empty rows;
 *   rspmm_ref_forward_f64     : fp32 operands up-cast to double, accumulated in double - the "true value"
 *                               the fp32 sums of the CUDA path are compared with (its summation order
 *                               legitimately differs from any sequential one);
 *   rspmm_ref_backward_f64    : likewise for the two gradients; the min/max gate compares the saved fp32
 *                               output with the message recomputed in fp32, as the reference does. */
int rspmm_ref_forward_arg_f32(int64_t n_out, int64_t dim, const int64_t *row_ptr, const int64_t *col,
                              const int64_t *layer, const float *val, const float *relation,
                              const float *input, float *output, int64_t *argidx, int sum_op, int mul_op) {
    if (sum_op != SUM_MIN && sum_op != SUM_MAX) return 1;
    if (mul_op < 0 || mul_op > 1) return 1;
    const float zero = nary_zero(sum_op);
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n_out; ++i) {
        float *out = output + i * dim;
        int64_t *arg = argidx + i * dim;
        for (int64_t d = 0; d < dim; ++d) { out[d] = zero; arg[d] = -1; }
        for (int64_t p = row_ptr[i]; p < row_ptr[i + 1]; ++p) {
            const float *rel = relation + layer[p] * dim;
            const float *in = input + col[p] * dim;
            const float w = val[p];
            for (int64_t d = 0; d < dim; ++d) {
                const float x = mul_op == MUL_MUL ? rel[d] * in[d] : rel[d] + in[d];
                const float y = w * x;
                const int better = sum_op == SUM_MAX ? (y > out[d]) : (y < out[d]);
                /* strict improvement moves the candidate, ties keep the earlier edge; a message equal to the
                 * identity (+-FLT_MAX) is still attained by an edge */
                if (better || (arg[d] < 0 && y == out[d])) arg[d] = p;
                if (sum_op == SUM_MAX) out[d] = out[d] > y ? out[d] : y;
                else out[d] = out[d] < y ? out[d] : y;
            }
        }
    }
    return 0;
}

int rspmm_ref_forward_f64(int64_t n_out, int64_t dim, const int64_t *row_ptr, const int64_t *col,
                          const int64_t *layer, const float *val, const float *relation,
                          const float *input, double *output, int mul_op) {
    if (mul_op < 0 || mul_op > 1) return 1;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n_out; ++i) {
        double *out = output + i * dim;
        for (int64_t d = 0; d < dim; ++d) out[d] = 0.0;
        for (int64_t p = row_ptr[i]; p < row_ptr[i + 1]; ++p) {
            const float *rel = relation + layer[p] * dim;
            const float *in = input + col[p] * dim;
            const double w = val[p];
            if (mul_op == MUL_MUL)
                for (int64_t d = 0; d < dim; ++d) out[d] += w * ((double)rel[d] * (double)in[d]);
            else
                for (int64_t d = 0; d < dim; ++d) out[d] += w * ((double)rel[d] + (double)in[d]);
        }
    }
    return 0;
}

int rspmm_ref_backward_f64(int64_t n_out, int64_t n_in, int64_t n_rel, int64_t dim,
                           const int64_t *row_ptr, const int64_t *col, const int64_t *layer,
                           const float *val, const float *relation, const float *input,
                           const float *output, const float *grad_output, double *grad_relation,
                           double *grad_input, int sum_op, int mul_op) {
    if (sum_op < 0 || sum_op > 2 || mul_op < 0 || mul_op > 1) return 1;
    memset(grad_relation, 0, (size_t)(n_rel * dim) * sizeof(double));
    memset(grad_input, 0, (size_t)(n_in * dim) * sizeof(double));
    /* feature columns are independent: every thread owns a block of columns and walks all edges in coalesced
     * order, so the sums are sequential (deterministic) and need no atomics */
#pragma omp parallel
    {
#ifdef _OPENMP
        const int64_t threads = omp_get_num_threads(), self = omp_get_thread_num();
#else
        const int64_t threads = 1, self = 0;
#endif
        const int64_t d0 = dim * self / threads, d1 = dim * (self + 1) / threads;
        for (int64_t i = 0; i < n_out && d0 < d1; ++i) {
            const float *g = grad_output + i * dim;
            const float *out = output ? output + i * dim : 0;
            for (int64_t p = row_ptr[i]; p < row_ptr[i + 1]; ++p) {
                const float *rel = relation + layer[p] * dim;
                const float *in = input + col[p] * dim;
                double *g_rel = grad_relation + layer[p] * dim;
                double *g_in = grad_input + col[p] * dim;
                const float w = val[p];
                for (int64_t d = d0; d < d1; ++d) {
                    if (sum_op != SUM_ADD) {
                        const float x = mul_op == MUL_MUL ? rel[d] * in[d] : rel[d] + in[d];
                        if (!(out[d] == w * x)) continue;
                    }
                    const double up = (double)g[d] * (double)w;
                    g_rel[d] += mul_op == MUL_MUL ? up * (double)in[d] : up;
                    g_in[d] += mul_op == MUL_MUL ? up * (double)rel[d] : up;
                }
            }
        }
    }
    return 0;
}
