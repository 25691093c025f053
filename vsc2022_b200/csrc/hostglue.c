/* Host-side glue of VCSLLocalization.localize_all in C (CPython API, no numpy headers: arrays come through the buffer
 * protocol).  The device work of a configs[3] batch takes ~6 ms; turning its 76 000 boxes into Match rows and locating 3 200
 * videos inside their base arrays took ~30 ms of interpreter time.  Both loops are restated here one for one:
 *
 *   vsc_match_rows    [tuple.__new__(Match, row) for row in zip(qid, rid, scores, q_start, q_end, r_start, r_end)]
 *                     (vsc/baseline/localization.py:61-78 builds one Match per box the same way, field by field)
 *   vsc_scan_views    the per-video checks of _DeviceVideos._ensure_views_of_one_array
 *
 * Built by vsc2022_b200/build_ext.py with gcc into csrc/_hostglue.so and loaded with ctypes.PyDLL (the GIL stays held).  The
 * Python implementations remain and are used when the library is absent; tests/test_hostglue_cpu.py compares the two. */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

static int same_format(const char *a, const char *b) { return (!a && !b) || (a && b && strcmp(a, b) == 0); }

static int get_buffer(PyObject *obj, Py_buffer *view, const char *what, Py_ssize_t itemsize) {
    if (PyObject_GetBuffer(obj, view, PyBUF_STRIDED_RO) != 0) return -1;
    if (view->itemsize != itemsize) {
        PyBuffer_Release(view);
        PyErr_Format(PyExc_TypeError, "%s: expected items of %zd bytes", what, itemsize);
        return -1;
    }
    return 0;
}

/* rows[i] = match_type(q_ids[pair_of[i]], r_ids[pair_of[i]], scores[i], q_start[i], q_end[i], r_start[i], r_end[i]);
 * `scores` is any iterable of n objects (a float32 numpy array yields numpy.float32 scalars: the type the reference's
 * `similarity[...].max() - bias` has), the four time arrays are contiguous float64 buffers, pair_of int64. */
PyObject *vsc_match_rows(PyObject *match_type, PyObject *q_ids, PyObject *r_ids, PyObject *pair_of, PyObject *scores,
                         PyObject *q_start, PyObject *q_end, PyObject *r_start, PyObject *r_end) {
    if (!PyType_Check(match_type) || !PyType_IsSubtype((PyTypeObject *)match_type, &PyTuple_Type) || !PyList_Check(q_ids) ||
        !PyList_Check(r_ids)) {
        PyErr_SetString(PyExc_TypeError, "vsc_match_rows: (tuple subclass, list, list, ...) expected");
        return NULL;
    }
    PyTypeObject *type = (PyTypeObject *)match_type;
    Py_buffer bp, b[4];
    PyObject *times[4] = {q_start, q_end, r_start, r_end};
    if (get_buffer(pair_of, &bp, "pair_of", 8) != 0) return NULL;
    int got = 0;
    PyObject *result = NULL, *it = NULL;
    for (; got < 4; ++got)
        if (get_buffer(times[got], &b[got], "time array", 8) != 0) goto done;
    {
        const Py_ssize_t n = bp.len / 8;
        for (int k = 0; k < 4; ++k)
            if (b[k].len / 8 != n || !PyBuffer_IsContiguous(&b[k], 'C')) {
                PyErr_SetString(PyExc_ValueError, "vsc_match_rows: time arrays must be contiguous and as long as pair_of");
                goto done;
            }
        if (!PyBuffer_IsContiguous(&bp, 'C')) { PyErr_SetString(PyExc_ValueError, "vsc_match_rows: pair_of must be contiguous"); goto done; }
        const int64_t *po = (const int64_t *)bp.buf;
        const double *t0 = (const double *)b[0].buf, *t1 = (const double *)b[1].buf, *t2 = (const double *)b[2].buf,
                     *t3 = (const double *)b[3].buf;
        const Py_ssize_t nq = PyList_GET_SIZE(q_ids), nr = PyList_GET_SIZE(r_ids);
        it = PyObject_GetIter(scores);
        if (!it) goto done;
        result = PyList_New(n);
        if (!result) goto done;
        for (Py_ssize_t i = 0; i < n; ++i) {
            const int64_t p = po[i];
            if (p < 0 || p >= nq || p >= nr) { PyErr_SetString(PyExc_IndexError, "vsc_match_rows: pair index out of range"); Py_CLEAR(result); goto done; }
            PyObject *score = PyIter_Next(it);
            if (!score) {
                if (!PyErr_Occurred()) PyErr_SetString(PyExc_ValueError, "vsc_match_rows: fewer scores than rows");
                Py_CLEAR(result); goto done;
            }
            PyObject *row = type->tp_alloc(type, 7);
            PyObject *f0 = PyFloat_FromDouble(t0[i]), *f1 = PyFloat_FromDouble(t1[i]), *f2 = PyFloat_FromDouble(t2[i]),
                     *f3 = PyFloat_FromDouble(t3[i]);
            if (!row || !f0 || !f1 || !f2 || !f3) {
                Py_XDECREF(row); Py_XDECREF(f0); Py_XDECREF(f1); Py_XDECREF(f2); Py_XDECREF(f3); Py_DECREF(score);
                Py_CLEAR(result); goto done;
            }
            PyObject *qid = PyList_GET_ITEM(q_ids, p), *rid = PyList_GET_ITEM(r_ids, p);
            Py_INCREF(qid); Py_INCREF(rid);
            PyTuple_SET_ITEM(row, 0, qid);      /* Match: query_id, ref_id, score, query_start, query_end, ref_start, ref_end */
            PyTuple_SET_ITEM(row, 1, rid);
            PyTuple_SET_ITEM(row, 2, score);
            PyTuple_SET_ITEM(row, 3, f0);
            PyTuple_SET_ITEM(row, 4, f1);
            PyTuple_SET_ITEM(row, 5, f2);
            PyTuple_SET_ITEM(row, 6, f3);
            PyList_SET_ITEM(result, i, row);
        }
    }
done:
    Py_XDECREF(it);
    for (int k = 0; k < got; ++k) PyBuffer_Release(&b[k]);
    PyBuffer_Release(&bp);
    return result;
}

/* For every id of `ids` (list): v = videos[id]; v.feature must be an ndarray whose .base is `root` with root's trailing shape
 * and strides, v.timestamps an ndarray whose .base is `troot` likewise, both starting at the SAME row of their base arrays.
 * Writes the row and the length per video into rows_out / lens_out (int64 buffers of len(ids)) and returns len(ids); returns
 * -1 (no exception) as soon as one video does not fit -- the caller then takes the general path. */
PyObject *vsc_scan_views(PyObject *videos, PyObject *ids, PyObject *root, PyObject *troot, PyObject *ndarray_type,
                         PyObject *rows_out, PyObject *lens_out) {
    if (!PyDict_Check(videos) || !PyList_Check(ids)) { PyErr_SetString(PyExc_TypeError, "vsc_scan_views: (dict, list, ...) expected"); return NULL; }
    Py_buffer br, bt, bo, bl;
    if (PyObject_GetBuffer(root, &br, PyBUF_STRIDED_RO | PyBUF_FORMAT) != 0) return NULL;
    if (PyObject_GetBuffer(troot, &bt, PyBUF_STRIDED_RO | PyBUF_FORMAT) != 0) { PyBuffer_Release(&br); return NULL; }
    if (PyObject_GetBuffer(rows_out, &bo, PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) != 0) { PyBuffer_Release(&br); PyBuffer_Release(&bt); return NULL; }
    if (PyObject_GetBuffer(lens_out, &bl, PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) != 0) { PyBuffer_Release(&br); PyBuffer_Release(&bt); PyBuffer_Release(&bo); return NULL; }
    const Py_ssize_t n = PyList_GET_SIZE(ids);
    long long answer = -1;
    PyObject *s_feature = PyUnicode_InternFromString("feature"), *s_ts = PyUnicode_InternFromString("timestamps"),
             *s_base = PyUnicode_InternFromString("base");
    if (!s_feature || !s_ts || !s_base) goto out;
    if (br.ndim != 2 || (bt.ndim != 1 && bt.ndim != 2) || bo.len / 8 < n || bl.len / 8 < n || bo.itemsize != 8 || bl.itemsize != 8 ||
        br.strides[0] <= 0 || bt.strides[0] <= 0)
        goto out;
    {
        int64_t *rows = (int64_t *)bo.buf, *lens = (int64_t *)bl.buf;
        Py_ssize_t i = 0;
        for (; i < n; ++i) {
            PyObject *v = PyDict_GetItemWithError(videos, PyList_GET_ITEM(ids, i));     /* borrowed */
            if (!v) { if (PyErr_Occurred()) PyErr_Clear(); break; }
            PyObject *f = PyObject_GetAttr(v, s_feature), *t = f ? PyObject_GetAttr(v, s_ts) : NULL;
            int ok = f && t && (PyObject *)Py_TYPE(f) == ndarray_type && (PyObject *)Py_TYPE(t) == ndarray_type;
            if (ok) {
                PyObject *fb = PyObject_GetAttr(f, s_base), *tb = PyObject_GetAttr(t, s_base);
                ok = fb == root && tb == troot;
                Py_XDECREF(fb); Py_XDECREF(tb);
            }
            Py_buffer vf, vt;
            if (ok && PyObject_GetBuffer(f, &vf, PyBUF_STRIDED_RO | PyBUF_FORMAT) == 0) {
                if (PyObject_GetBuffer(t, &vt, PyBUF_STRIDED_RO | PyBUF_FORMAT) == 0) {
                    ok = same_format(vf.format, br.format) && same_format(vt.format, bt.format) && vf.ndim == 2 && vt.ndim == bt.ndim && vf.shape[0] > 0 && vt.shape[0] == vf.shape[0] &&
                         vf.itemsize == br.itemsize && vt.itemsize == bt.itemsize && vf.shape[1] == br.shape[1] &&
                         vf.strides[0] == br.strides[0] && vf.strides[1] == br.strides[1] && vt.strides[0] == bt.strides[0] &&
                         (bt.ndim == 1 || (vt.shape[1] == bt.shape[1] && vt.strides[1] == bt.strides[1]));
                    if (ok) {
                        const Py_ssize_t off = (const char *)vf.buf - (const char *)br.buf, toff = (const char *)vt.buf - (const char *)bt.buf;
                        const Py_ssize_t row = off / br.strides[0];
                        ok = off >= 0 && off % br.strides[0] == 0 && row + vf.shape[0] <= br.shape[0] && toff == row * bt.strides[0];
                        if (ok) { rows[i] = (int64_t)row; lens[i] = (int64_t)vf.shape[0]; }
                    }
                    PyBuffer_Release(&vt);
                } else { PyErr_Clear(); ok = 0; }
                PyBuffer_Release(&vf);
            } else if (ok) { PyErr_Clear(); ok = 0; }
            Py_XDECREF(f); Py_XDECREF(t);
            if (PyErr_Occurred()) PyErr_Clear();
            if (!ok) break;
        }
        if (i == n) answer = (long long)n;
    }
out:
    Py_XDECREF(s_feature); Py_XDECREF(s_ts); Py_XDECREF(s_base);
    PyBuffer_Release(&br); PyBuffer_Release(&bt); PyBuffer_Release(&bo); PyBuffer_Release(&bl);
    return PyLong_FromLongLong(answer);
}
