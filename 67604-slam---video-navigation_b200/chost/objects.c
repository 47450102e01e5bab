/* Host-side helper of the reference-typed drop-in (not on the compute path): builds the tuples of cv2.DMatch that
 * cv2.BFMatcher.match / knnMatch return (final_project/algorithms/matching.py:44, backend/database/database.py:54-55,
 * VAN_ex/code/ex1.py:189-190) from the index / distance arrays the GPU matcher delivers, and reads the indices back
 * out of such a tuple (matching.py:48-69 walks the matches one attribute at a time).
 *
 * cv2's generated constructor spends ~1 us per DMatch(queryIdx, trainIdx, imgIdx, distance) on overload resolution:
 * 3 ms for the 3000 matches of one frame, more than the GPU call that produced them.  The no-argument constructor
 * costs 0.13 us, and the object is PyObject_HEAD followed by cv::DMatch { int queryIdx, trainIdx, imgIdx; float
 * distance; }: the fields are written in place.  The caller (slamfe/_objects.py) proves that layout against the
 * 4-argument constructor and the attribute getters before any of this is used, and keeps the pure-Python path
 * otherwise.  Built with gcc by slamfe.build.build_objects(); no CUDA, no numpy headers (buffer protocol only). */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

typedef struct {
    int32_t query_idx, train_idx, img_idx;
    float distance;
} dmatch_fields;

static int get_buffer(PyObject *obj, Py_buffer *view, Py_ssize_t itemsize, const char *what)
{
    if (PyObject_GetBuffer(obj, view, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) != 0) return -1;
    if (view->itemsize != itemsize) {
        PyErr_Format(PyExc_TypeError, "%s: expected %zd-byte items, got %zd", what, itemsize, view->itemsize);
        PyBuffer_Release(view);
        return -1;
    }
    return 0;
}

/* dmatch_tuple(cls, offset, query_idx[int32], train_idx[int32], distance[float32]) -> tuple of cls, imgIdx = 0 */
static PyObject *dmatch_tuple(PyObject *self, PyObject *args)
{
    PyObject *cls, *qo, *to, *dobj;
    Py_ssize_t offset;
    if (!PyArg_ParseTuple(args, "OnOOO", &cls, &offset, &qo, &to, &dobj)) return NULL;
    Py_buffer q, t, d;
    if (get_buffer(qo, &q, 4, "query_idx") != 0) return NULL;
    if (get_buffer(to, &t, 4, "train_idx") != 0) { PyBuffer_Release(&q); return NULL; }
    if (get_buffer(dobj, &d, 4, "distance") != 0) { PyBuffer_Release(&q); PyBuffer_Release(&t); return NULL; }
    const Py_ssize_t n = q.len / 4;
    PyObject *out = NULL;
    if (t.len / 4 != n || d.len / 4 != n) {
        PyErr_SetString(PyExc_ValueError, "dmatch_tuple: arrays differ in length");
        goto done;
    }
    out = PyTuple_New(n);
    if (!out) goto done;
    const int32_t *qi = (const int32_t *)q.buf, *ti = (const int32_t *)t.buf;
    const float *di = (const float *)d.buf;
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject *m = PyObject_CallNoArgs(cls);
        if (!m) { Py_CLEAR(out); goto done; }
        dmatch_fields f = {qi[i], ti[i], 0, di[i]};
        memcpy((char *)m + offset, &f, sizeof f);
        PyTuple_SET_ITEM(out, i, m);
    }
done:
    PyBuffer_Release(&q); PyBuffer_Release(&t); PyBuffer_Release(&d);
    return out;
}

/* knn_tuples(first, second, has_second[uint8]) -> tuple of (first[i],) or (first[i], next of second) */
static PyObject *knn_tuples(PyObject *self, PyObject *args)
{
    PyObject *first, *second, *ho;
    if (!PyArg_ParseTuple(args, "O!O!O", &PyTuple_Type, &first, &PyTuple_Type, &second, &ho)) return NULL;
    Py_buffer h;
    if (get_buffer(ho, &h, 1, "has_second") != 0) return NULL;
    const Py_ssize_t n = PyTuple_GET_SIZE(first), n2 = PyTuple_GET_SIZE(second);
    PyObject *out = NULL;
    if (h.len != n) {
        PyErr_SetString(PyExc_ValueError, "knn_tuples: has_second has the wrong length");
        goto done;
    }
    out = PyTuple_New(n);
    if (!out) goto done;
    const uint8_t *hs = (const uint8_t *)h.buf;
    Py_ssize_t j = 0;
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject *a = PyTuple_GET_ITEM(first, i), *inner;
        if (hs[i]) {
            if (j >= n2) { PyErr_SetString(PyExc_ValueError, "knn_tuples: too few second matches"); Py_CLEAR(out); goto done; }
            inner = PyTuple_Pack(2, a, PyTuple_GET_ITEM(second, j++));
        } else {
            inner = PyTuple_Pack(1, a);
        }
        if (!inner) { Py_CLEAR(out); goto done; }
        PyTuple_SET_ITEM(out, i, inner);
    }
done:
    PyBuffer_Release(&h);
    return out;
}

/* dmatch_indices(matches: sequence of cls, cls, offset) -> (bytes of int32 queryIdx, bytes of int32 trainIdx) */
static PyObject *dmatch_indices(PyObject *self, PyObject *args)
{
    PyObject *seq, *cls;
    Py_ssize_t offset;
    if (!PyArg_ParseTuple(args, "OOn", &seq, &cls, &offset)) return NULL;
    PyObject *fast = PySequence_Fast(seq, "dmatch_indices: matches must be a sequence");
    if (!fast) return NULL;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
    PyObject *qb = PyBytes_FromStringAndSize(NULL, n * 4), *tb = PyBytes_FromStringAndSize(NULL, n * 4), *out = NULL;
    if (!qb || !tb) goto done;
    int32_t *qi = (int32_t *)PyBytes_AS_STRING(qb), *ti = (int32_t *)PyBytes_AS_STRING(tb);
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject *m = PySequence_Fast_GET_ITEM(fast, i);
        if ((PyObject *)Py_TYPE(m) != cls) {   /* exact type only: anything else goes the attribute way in Python */
            PyErr_SetString(PyExc_TypeError, "dmatch_indices: not a cv2.DMatch");
            goto done;
        }
        dmatch_fields f;
        memcpy(&f, (const char *)m + offset, sizeof f);
        qi[i] = f.query_idx;
        ti[i] = f.train_idx;
    }
    out = PyTuple_Pack(2, qb, tb);
done:
    Py_XDECREF(qb); Py_XDECREF(tb); Py_DECREF(fast);
    return out;
}

static PyMethodDef methods[] = {
    {"dmatch_tuple", dmatch_tuple, METH_VARARGS, "arrays -> tuple of cv2.DMatch (imgIdx 0)"},
    {"knn_tuples", knn_tuples, METH_VARARGS, "first / second matches -> knnMatch's tuple of tuples"},
    {"dmatch_indices", dmatch_indices, METH_VARARGS, "sequence of cv2.DMatch -> (queryIdx bytes, trainIdx bytes)"},
    {NULL, NULL, 0, NULL}};
static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "_slamfe_objects", NULL, -1, methods};
PyMODINIT_FUNC PyInit__slamfe_objects(void) { return PyModule_Create(&moduledef); }
