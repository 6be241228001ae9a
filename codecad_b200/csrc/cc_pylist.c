/* CPython helper of codecad_b200.subdivision.LeafBlocks: builds the reference's list of leaf blocks
 *     [(grid_dims, corner Vector(float64), step, int_corner Vector(int), int_step), ...]
 * (subdivision.py:97-111) from the two arrays the library returns.  12 000 blocks are ~110 000 Python
 * objects; built here with the C API in a third of the time of the fastest pure-Python form
 * (zip / map / tuple.__new__).  Compiled by codecad_b200/build.py into _cc_pylist.so; optional — without
 * it LeafBlocks uses the Python form. */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

static PyObject *vector_of(PyTypeObject *type, PyObject *a, PyObject *b, PyObject *c)
{
    /* a namedtuple's __new__ is tuple.__new__: allocate the subclass instance and fill it */
    PyObject *t = type->tp_alloc(type, 3);
    if (!t || !a || !b || !c) {
        Py_XDECREF(t); Py_XDECREF(a); Py_XDECREF(b); Py_XDECREF(c);
        return NULL;
    }
    PyTuple_SET_ITEM(t, 0, a);
    PyTuple_SET_ITEM(t, 1, b);
    PyTuple_SET_ITEM(t, 2, c);
    return t;
}

/* leaf_blocks(vector_type, dims, corners: buffer of float64[n][3], step, int_corners: buffer of int64[n][3], int_step) */
static PyObject *leaf_blocks(PyObject *self, PyObject *args)
{
    PyObject *type_obj, *dims, *step, *int_step;
    Py_buffer corners, int_corners;
    if (!PyArg_ParseTuple(args, "OOy*Oy*O", &type_obj, &dims, &corners, &step, &int_corners, &int_step)) return NULL;
    PyObject *list = NULL;
    if (!PyType_Check(type_obj) || !PyType_IsSubtype((PyTypeObject *)type_obj, &PyTuple_Type)) {
        PyErr_SetString(PyExc_TypeError, "vector_type must be a tuple subclass");
        goto done;
    }
    if (corners.len % 24 || corners.len != int_corners.len) {
        PyErr_SetString(PyExc_ValueError, "corners and int_corners must be [n][3] arrays of 8-byte items");
        goto done;
    }
    {
        const Py_ssize_t n = corners.len / 24;
        const double *c = (const double *)corners.buf;
        const int64_t *ic = (const int64_t *)int_corners.buf;
        PyTypeObject *type = (PyTypeObject *)type_obj;
        list = PyList_New(n);
        if (!list) goto done;
        /* 3 n tuples would trigger a young-generation collection every few hundred of them, each one
         * walking what was just built; nothing allocated here can be garbage or part of a cycle */
        const int gc_was_enabled = PyGC_Disable();
        for (Py_ssize_t i = 0; i < n; ++i) {
            PyObject *corner = vector_of(type, PyFloat_FromDouble(c[3 * i]), PyFloat_FromDouble(c[3 * i + 1]),
                                         PyFloat_FromDouble(c[3 * i + 2]));
            PyObject *icorner = corner ? vector_of(type, PyLong_FromLongLong(ic[3 * i]), PyLong_FromLongLong(ic[3 * i + 1]),
                                                   PyLong_FromLongLong(ic[3 * i + 2])) : NULL;
            PyObject *item = icorner ? PyTuple_New(5) : NULL;
            if (!item) {
                Py_XDECREF(corner); Py_XDECREF(icorner);
                Py_CLEAR(list);
                if (gc_was_enabled) PyGC_Enable();
                goto done;
            }
            Py_INCREF(dims); Py_INCREF(step); Py_INCREF(int_step);
            PyTuple_SET_ITEM(item, 0, dims);
            PyTuple_SET_ITEM(item, 1, corner);
            PyTuple_SET_ITEM(item, 2, step);
            PyTuple_SET_ITEM(item, 3, icorner);
            PyTuple_SET_ITEM(item, 4, int_step);
            PyList_SET_ITEM(list, i, item);
        }
        if (gc_was_enabled) PyGC_Enable();
    }
done:
    PyBuffer_Release(&corners);
    PyBuffer_Release(&int_corners);
    return list;
}

static PyMethodDef methods[] = {{"leaf_blocks", leaf_blocks, METH_VARARGS, "list of leaf-block tuples from two arrays"},
                                {NULL, NULL, 0, NULL}};
static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_cc_pylist", NULL, -1, methods};
PyMODINIT_FUNC PyInit__cc_pylist(void) { return PyModule_Create(&module); }
