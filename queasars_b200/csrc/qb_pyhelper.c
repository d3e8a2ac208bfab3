/* Host-side helper (CPython C API, loaded with ctypes.PyDLL): turns the reference's ``parameter_values: list[list[float]]``
 * (/root/reference/queasars/circuit_evaluation/circuit_evaluation.py:62-87; built with ``ndarray.tolist()`` at
 * evqe/evolutionary_algorithm/mutation.py:64) into the flat float64 buffer the C-ABI takes, in one pass and without one NumPy
 * call per row.  Pure data marshalling: no arithmetic of the hot path lives here, and the engine works without it (NumPy path).
 *
 * Build: gcc -O2 -shared -fPIC -I<python include dir> -o libqb_pyhelper.so qb_pyhelper.c
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>

/* rows: sequence of n_rows sequences of numbers; expected[i] = number of values row i must have; out: sum(expected) doubles.
 * Returns the number of doubles written, -(i + 1) when row i has the wrong length, LLONG_MIN on a type error (Python exception
 * cleared). */
long long qb_pack_rows(PyObject* rows, double* out, const long long* expected, long long n_rows) {
    PyObject* outer = PySequence_Fast(rows, "rows must be a sequence");
    if (!outer) {
        PyErr_Clear();
        return LLONG_MIN;
    }
    if (PySequence_Fast_GET_SIZE(outer) != n_rows) {
        Py_DECREF(outer);
        return LLONG_MIN;
    }
    long long pos = 0;
    for (long long i = 0; i < n_rows; ++i) {
        PyObject* row = PySequence_Fast(PySequence_Fast_GET_ITEM(outer, i), "row must be a sequence");
        if (!row) {
            PyErr_Clear();
            Py_DECREF(outer);
            return LLONG_MIN;
        }
        const Py_ssize_t n = PySequence_Fast_GET_SIZE(row);
        if ((long long)n != expected[i]) {
            Py_DECREF(row);
            Py_DECREF(outer);
            return -(i + 1);
        }
        PyObject** items = PySequence_Fast_ITEMS(row);
        for (Py_ssize_t j = 0; j < n; ++j) {
            PyObject* v = items[j];
            double d = PyFloat_CheckExact(v) ? PyFloat_AS_DOUBLE(v) : PyFloat_AsDouble(v);
            if (d == -1.0 && PyErr_Occurred()) {
                PyErr_Clear();
                Py_DECREF(row);
                Py_DECREF(outer);
                return LLONG_MIN;
            }
            out[pos++] = d;
        }
        Py_DECREF(row);
    }
    Py_DECREF(outer);
    return pos;
}

/* One circuit with one parameter row -- the optimizer loop's call (SPSA / NFT evaluate one or two points at a time,
 * /root/reference/queasars/minimum_eigensolvers/evqe/evolutionary_algorithm/mutation.py:59-77): packs ``row`` and calls
 * ``qb_evaluate_expectation`` (its address in ``fn``) with batch 1, the interpreter lock released for the duration of the native call.
 * Returns a Python float (the value), or a Python int: the native status when it is not 0, -1000 when the row has the wrong
 * length, -2000 when the row cannot be read here (the caller then takes the NumPy path). */
typedef int (*qb_eval_fn)(void* ctx, int batch, const long long* plan_ids, const double* params, const long long* offsets, long long ham_id,
                          double* out_values);

PyObject* qb_single_expectation(void* fn, void* ctx, long long plan_id, long long n_params, PyObject* row, long long ham_id) {
    double stack_buf[256];
    double* buf = stack_buf;
    PyObject* seq = PySequence_Fast(row, "row must be a sequence");
    if (!seq) {
        PyErr_Clear();
        return PyLong_FromLong(-2000);
    }
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
    if ((long long)n != n_params) {
        Py_DECREF(seq);
        return PyLong_FromLong(-1000);
    }
    if (n > 256) {
        buf = (double*)PyMem_RawMalloc(sizeof(double) * (size_t)n);
        if (!buf) {
            Py_DECREF(seq);
            return PyLong_FromLong(-2000);
        }
    }
    PyObject** items = PySequence_Fast_ITEMS(seq);
    for (Py_ssize_t j = 0; j < n; ++j) {
        PyObject* v = items[j];
        const double d = PyFloat_CheckExact(v) ? PyFloat_AS_DOUBLE(v) : PyFloat_AsDouble(v);
        if (d == -1.0 && PyErr_Occurred()) {
            PyErr_Clear();
            Py_DECREF(seq);
            if (buf != stack_buf) PyMem_RawFree(buf);
            return PyLong_FromLong(-2000);
        }
        buf[j] = d;
    }
    Py_DECREF(seq);
    const long long ids[1] = {plan_id}, offsets[2] = {0, n_params};
    double value = 0.0;
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = ((qb_eval_fn)fn)(ctx, 1, ids, buf, offsets, ham_id, &value);
    Py_END_ALLOW_THREADS
    if (buf != stack_buf) PyMem_RawFree(buf);
    return rc == 0 ? PyFloat_FromDouble(value) : PyLong_FromLong(rc);
}
