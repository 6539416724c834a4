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

static PyMethodDef methods[] = {{"pair_list", pair_list, METH_VARARGS, "list of (names[rows[i]], names[cols[i]]) tuples"},
                                {"pair_lists", pair_lists, METH_VARARGS, "per-class lists of (names[rows[i]], names[cols[i]])"},
                                {NULL, NULL, 0, NULL}};
static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "_pytuples", NULL, -1, methods, NULL, NULL, NULL, NULL};
PyMODINIT_FUNC PyInit__pytuples(void) { return PyModule_Create(&moduledef); }
