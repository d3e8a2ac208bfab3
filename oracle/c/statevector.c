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
