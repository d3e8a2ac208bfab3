/*
 * CPU oracle in C (OpenMP) -- TEST / BASELINE INFRASTRUCTURE ONLY, never linked into the product.
 *
 * Restates the same semantics as oracle/qiskit_semantics.py (which documents the reference call sites and the upstream
 * algorithms it follows; parity unpinned at the primitive boundary, see oracle/__init__.py):
 *   - little-endian statevector evolution of `u` / `cu3`-style gates: a 2x2 complex matrix on one target qubit, optionally
 *     controlled by one qubit ([upstream] Statevector._evolve_instruction; gate matrices from
 *     /root/reference/queasars/minimum_eigensolvers/evqe/quantum_circuit/quantum_gate.py:96-102, 157-165)
 *   - diagonal expectation sum_k |psi_k|^2 E(k) ([upstream] expval_pauli_no_x summed over the terms of the operator handed to
 *     /root/reference/queasars/circuit_evaluation/circuit_evaluation.py:200-215)
 * It exists to give bench.py a CPU baseline that uses all host cores the way a compiled simulator (Qiskit-Aer) would.
 *
 * Build: gcc -O3 -fopenmp -shared -fPIC -o oracle/c/liboracle.so oracle/c/statevector.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double re, im; } cplx;

static inline cplx cmul(cplx a, cplx b) { cplx r = {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; return r; }
static inline cplx cadd(cplx a, cplx b) { cplx r = {a.re + b.re, a.im + b.im}; return r; }

/* state: 2^n complex128; m: row-major 2x2 (m00, m01, m10, m11); control < 0: uncontrolled */
void oracle_apply_1q(cplx* state, int n, int target, int control, const cplx* m) {
    const int64_t half = (int64_t)1 << (n - 1);
    const int64_t tbit = (int64_t)1 << target;
    const int64_t cbit = control >= 0 ? (int64_t)1 << control : 0;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < half; ++i) {
        /* insert a zero at the target position */
        const int64_t lo = ((i >> target) << (target + 1)) | (i & (tbit - 1));
        if (cbit && !(lo & cbit)) continue;
        const int64_t hi = lo | tbit;
        const cplx x = state[lo], y = state[hi];
        state[lo] = cadd(cmul(m[0], x), cmul(m[1], y));
        state[hi] = cadd(cmul(m[2], x), cmul(m[3], y));
    }
}

void oracle_init_zero(cplx* state, int n) {
    const int64_t size = (int64_t)1 << n;
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < size; ++k) state[k].re = 0.0, state[k].im = 0.0;
    state[0].re = 1.0;
}

/* gates: n_gates records of (target, control, 8 doubles) ; returns sum_k |psi_k|^2 table[k] */
double oracle_run_circuit(cplx* state, int n, int n_gates, const int32_t* targets, const int32_t* controls, const double* matrices,
                          const double* table) {
    oracle_init_zero(state, n);
    for (int g = 0; g < n_gates; ++g) oracle_apply_1q(state, n, targets[g], controls[g], (const cplx*)(matrices + 8 * (size_t)g));
    if (!table) return 0.0;
    const int64_t size = (int64_t)1 << n;
    double acc = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : acc)
    for (int64_t k = 0; k < size; ++k) acc += (state[k].re * state[k].re + state[k].im * state[k].im) * table[k];
    return acc;
}

/* E(k) = sum_t c_t (-1)^{popcount(k & z_t)} */
void oracle_diag_table(double* table, int n, int n_terms, const uint64_t* z, const double* c) {
    const int64_t size = (int64_t)1 << n;
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < size; ++k) {
        double e = 0.0;
        for (int t = 0; t < n_terms; ++t) e += (__builtin_popcountll((uint64_t)k & z[t]) & 1) ? -c[t] : c[t];
        table[k] = e;
    }
}

/* Re sum_t (cr_t + i ci_t) <psi|P_t|psi> with P_t = i^{popcount(x_t & z_t)} X^{x_t} Z^{z_t}, i.e. P|k> = i^{nY} (-1)^{pc(k & z)} |k ^ x>
 * ([upstream] Statevector.expectation_value -> expval_pauli_no_x / expval_pauli_with_x, ONE pass over the state per term, which
 * is what the operator handed to /root/reference/queasars/circuit_evaluation/circuit_evaluation.py:200-215 costs on the CPU;
 * the imaginary part is dropped at :215).  Same formula as oracle/qiskit_semantics.py:pauli_expectation. */
double oracle_pauli_sum(const cplx* state, int n, int n_terms, const uint64_t* x, const uint64_t* z, const double* cr, const double* ci) {
    const int64_t size = (int64_t)1 << n;
    double total = 0.0;
    for (int t = 0; t < n_terms; ++t) {
        const uint64_t xm = x[t], zm = z[t];
        double sr = 0.0, si = 0.0;
        if (xm == 0) {
#pragma omp parallel for schedule(static) reduction(+ : sr)
            for (int64_t k = 0; k < size; ++k) {
                const double p = state[k].re * state[k].re + state[k].im * state[k].im;
                sr += (__builtin_popcountll((uint64_t)k & zm) & 1) ? -p : p;
            }
        } else {
#pragma omp parallel for schedule(static) reduction(+ : sr, si)
            for (int64_t k = 0; k < size; ++k) {
                /* conj(psi[k ^ x]) * sign * psi[k] */
                const cplx a = state[k], b = state[(uint64_t)k ^ xm];
                const double s = (__builtin_popcountll((uint64_t)k & zm) & 1) ? -1.0 : 1.0;
                sr += s * (b.re * a.re + b.im * a.im);
                si += s * (b.re * a.im - b.im * a.re);
            }
        }
        /* multiply by i^{nY} and by the coefficient, keep the real part */
        const int ny = __builtin_popcountll(xm & zm) & 3;
        double pr = sr, pi = si;
        if (ny == 1) pr = -si, pi = sr;
        else if (ny == 2) pr = -sr, pi = -si;
        else if (ny == 3) pr = si, pi = -sr;
        total += cr[t] * pr - (ci ? ci[t] : 0.0) * pi;
    }
    return total;
}

/* [upstream] Statevector.sample_memory -> numpy Generator.choice(p=probs): cdf = probs.cumsum() (sequential, left to right, like
 * numpy's cumsum); cdf /= cdf[-1]; idx = cdf.searchsorted(uniforms, side='right')  (measure_quasi_distributions,
 * /root/reference/queasars/circuit_evaluation/circuit_evaluation.py:29-59).  `cdf` is caller-provided scratch of 2^n doubles. */
void oracle_sample(const cplx* state, int n, int shots, const double* uniforms, int64_t* out, double* cdf) {
    const int64_t size = (int64_t)1 << n;
    double run = 0.0;
    for (int64_t k = 0; k < size; ++k) {
        run += state[k].re * state[k].re + state[k].im * state[k].im;
        cdf[k] = run;
    }
    const double last = cdf[size - 1];
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < size; ++k) cdf[k] /= last;
#pragma omp parallel for schedule(static)
    for (int s = 0; s < shots; ++s) {
        const double u = uniforms[s];
        int64_t lo = 0, hi = size; /* first index with cdf[i] > u */
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (cdf[mid] > u) hi = mid;
            else lo = mid + 1;
        }
        out[s] = lo;
    }
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_threads(int n) {
#ifdef _OPENMP
    extern void omp_set_num_threads(int);
    if (n > 0) omp_set_num_threads(n);
#endif
}
