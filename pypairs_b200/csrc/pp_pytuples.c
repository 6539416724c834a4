/* pp_pytuples.c -- host-side helper, CPython extension (no CUDA): builds the list of (gene_high, gene_low)
 * tuples of a sandbag result (pypairs/tools/sandbag.py:131-140) from index arrays in one C loop.
 * pair_list(names: list, rows: int64 buffer, cols: int64 buffer) -> [(names[rows[i]], names[cols[i]]), ...]
 * Optional: pypairs_b200/sandbag.py falls back to list(zip(...)) when this module is not built. */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

static PyObject *pair_list(PyObject *self, PyObject *args) {
    PyObject *names, *out = NULL;
    Py_buffer rows, cols;
    (void)self;
    if (!PyArg_ParseTuple(args, "O!y*y*", &PyList_Type, &names, &rows, &cols)) return NULL;
    const Py_ssize_t n = rows.len / (Py_ssize_t)sizeof(int64_t), n_names = PyList_GET_SIZE(names);
    if (cols.len != rows.len) {
        PyErr_SetString(PyExc_ValueError, "rows and cols differ in length");
        goto done;
    }
    const int64_t *r = (const int64_t *)rows.buf, *c = (const int64_t *)cols.buf;
    out = PyList_New(n);
    if (!out) goto done;
    for (Py_ssize_t i = 0; i < n; ++i) {
        if (r[i] < 0 || r[i] >= n_names || c[i] < 0 || c[i] >= n_names) {
            Py_CLEAR(out);
            PyErr_SetString(PyExc_IndexError, "gene index out of range");
            goto done;
        }
        PyObject *a = PyList_GET_ITEM(names, r[i]), *b = PyList_GET_ITEM(names, c[i]);
        PyObject *t = PyTuple_New(2);
        if (!t) {
            Py_CLEAR(out);
            goto done;
        }
        Py_INCREF(a);
        Py_INCREF(b);
        PyTuple_SET_ITEM(t, 0, a);
        PyTuple_SET_ITEM(t, 1, b);
        PyList_SET_ITEM(out, i, t);
    }
done:
    PyBuffer_Release(&rows);
    PyBuffer_Release(&cols);
    return out;
}

/* pair_lists(names: list, rows: int64 buffer, cols: int64 buffer, cls: int32 buffer, n_classes) ->
 * [list of (names[rows[i]], names[cols[i]]) for the entries of class k, in input order] for k in range(n_classes).
 * One pass, with the cyclic collector paused: the new tuples only hold strings. */
static PyObject *pair_lists(PyObject *self, PyObject *args) {
    PyObject *names, *out = NULL;
    Py_buffer rows, cols, cls;
    int n_classes = 0;
    (void)self;
    if (!PyArg_ParseTuple(args, "O!y*y*y*i", &PyList_Type, &names, &rows, &cols, &cls, &n_classes)) return NULL;
    const Py_ssize_t n = rows.len / (Py_ssize_t)sizeof(int64_t), n_names = PyList_GET_SIZE(names);
    const int64_t *r = (const int64_t *)rows.buf, *c = (const int64_t *)cols.buf;
    const int32_t *k = (const int32_t *)cls.buf;
    Py_ssize_t *count = NULL, *fill = NULL;
    const int gc_was = PyGC_Disable();
    if (cols.len != rows.len || cls.len / (Py_ssize_t)sizeof(int32_t) != n || n_classes < 0) {
        PyErr_SetString(PyExc_ValueError, "rows, cols and cls differ in length");
        goto done;
    }
    count = (Py_ssize_t *)calloc((size_t)n_classes + 1, sizeof(Py_ssize_t));
    fill = (Py_ssize_t *)calloc((size_t)n_classes + 1, sizeof(Py_ssize_t));
    if (!count || !fill) {
        PyErr_NoMemory();
        goto done;
    }
    for (Py_ssize_t i = 0; i < n; ++i) {
        if (k[i] < 0 || k[i] >= n_classes || r[i] < 0 || r[i] >= n_names || c[i] < 0 || c[i] >= n_names) {
            PyErr_SetString(PyExc_IndexError, "class or gene index out of range");
            goto done;
        }
        count[k[i]]++;
    }
    out = PyList_New(n_classes);
    if (!out) goto done;
    for (int j = 0; j < n_classes; ++j) {
        PyObject *l = PyList_New(count[j]);
        if (!l) {
            Py_CLEAR(out);
            goto done;
        }
        PyList_SET_ITEM(out, j, l);
    }
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject *a = PyList_GET_ITEM(names, r[i]), *b = PyList_GET_ITEM(names, c[i]);
        PyObject *t = PyTuple_New(2);
        if (!t) {
            Py_CLEAR(out);
            goto done;
        }
        Py_INCREF(a);
        Py_INCREF(b);
        PyTuple_SET_ITEM(t, 0, a);
        PyTuple_SET_ITEM(t, 1, b);
        PyList_SET_ITEM(PyList_GET_ITEM(out, k[i]), fill[k[i]]++, t);
    }
done:
    if (gc_was) PyGC_Enable();
    free(count);
    free(fill);
    PyBuffer_Release(&rows);
    PyBuffer_Release(&cols);
    PyBuffer_Release(&cls);
    return out;
}

/* pair_indices(pairs: sequence of 2-sequences, name_to_idx: dict) -> (bytes of int64[n], bytes of int64[n]):
 * name_to_idx.get(p[0], -1), name_to_idx.get(p[1], -1) for every pair p, the look-ups of pypairs/tools/cyclone.py:291-300
 * in one C loop (a marker dict found by sandbag holds millions of pairs). */
static PyObject *pair_indices(PyObject *self, PyObject *args) {
    PyObject *pairs, *map, *fast = NULL, *ba = NULL, *bb = NULL, *out = NULL;
    (void)self;
    if (!PyArg_ParseTuple(args, "OO!", &pairs, &PyDict_Type, &map)) return NULL;
    fast = PySequence_Fast(pairs, "pairs must be a sequence");
    if (!fast) return NULL;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
    ba = PyBytes_FromStringAndSize(NULL, n * (Py_ssize_t)sizeof(int64_t));
    bb = PyBytes_FromStringAndSize(NULL, n * (Py_ssize_t)sizeof(int64_t));
    if (!ba || !bb) goto done;
    {
        int64_t *a = (int64_t *)PyBytes_AS_STRING(ba), *b = (int64_t *)PyBytes_AS_STRING(bb);
        for (Py_ssize_t i = 0; i < n; ++i) {
            PyObject *p = PySequence_Fast_GET_ITEM(fast, i);
            PyObject *g[2];
            int own = 0;
            if (PyTuple_CheckExact(p) && PyTuple_GET_SIZE(p) >= 2) {
                g[0] = PyTuple_GET_ITEM(p, 0);
                g[1] = PyTuple_GET_ITEM(p, 1);
            } else if (PyList_CheckExact(p) && PyList_GET_SIZE(p) >= 2) {
                g[0] = PyList_GET_ITEM(p, 0);
                g[1] = PyList_GET_ITEM(p, 1);
            } else {
                g[0] = PySequence_GetItem(p, 0);
                g[1] = g[0] ? PySequence_GetItem(p, 1) : NULL;
                if (!g[0] || !g[1]) {
                    Py_XDECREF(g[0]);
                    goto done;
                }
                own = 1;
            }
            for (int j = 0; j < 2; ++j) {
                PyObject *v = PyDict_GetItemWithError(map, g[j]);  /* borrowed */
                int64_t idx = -1;
                if (v) {
                    idx = (int64_t)PyLong_AsLongLong(v);
                } else if (PyErr_Occurred()) {
                    if (own) { Py_DECREF(g[0]); Py_DECREF(g[1]); }
                    goto done;
                }
                if (idx == -1 && PyErr_Occurred()) {
                    if (own) { Py_DECREF(g[0]); Py_DECREF(g[1]); }
                    goto done;
                }
                (j == 0 ? a : b)[i] = idx;
            }
            if (own) { Py_DECREF(g[0]); Py_DECREF(g[1]); }
        }
    }
    out = PyTuple_Pack(2, ba, bb);
done:
    Py_XDECREF(fast);
    Py_XDECREF(ba);
    Py_XDECREF(bb);
    return out;
}

static PyMethodDef methods[] = {{"pair_list", pair_list, METH_VARARGS, "list of (names[rows[i]], names[cols[i]]) tuples"},
                                {"pair_lists", pair_lists, METH_VARARGS, "per-class lists of (names[rows[i]], names[cols[i]])"},
                                {"pair_indices", pair_indices, METH_VARARGS, "(int64 bytes, int64 bytes) of name_to_idx.get(p[0], -1), .get(p[1], -1)"},
                                {NULL, NULL, 0, NULL}};
static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "_pytuples", NULL, -1, methods, NULL, NULL, NULL, NULL};
PyMODINIT_FUNC PyInit__pytuples(void) { return PyModule_Create(&moduledef); }
