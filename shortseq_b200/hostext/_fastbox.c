/* _fastbox.c -- CPython helpers for the host side of the drop-in ShortSeqCounter.
 *
 * The reference builds its objects and fills its dict in C (Cython; counter.pyx:23-54, short_seq.pyx:54-74).  The
 * counting itself runs on the GPU here, but a list of a million `bytes` still has to be gathered into one buffer and
 * the distinct keys have to come back as ShortSeq objects in a dict: done with Python-level loops that costs more than
 * the whole GPU pass.  Three loops in C instead:
 *   gather(list)                      -> (ascii bytes, offsets bytes [int64 * (n+1)])   one pass for sizes, one memcpy pass
 *   box_many(cls, words, lens, W)     -> list of cls instances (slots _packed, _length, _hash filled in place)
 *   fill_counts(dict, objs, counts, order)   dict[objs[order[j]]] = counts[order[j]] in that order, hash known
 * No sequence arithmetic happens here: words, lengths and counts arrive from the CUDA library.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>
#include <structmember.h>

static PyObject *gather(PyObject *self, PyObject *arg)
{
    if (!PyList_CheckExact(arg)) {
        PyErr_SetString(PyExc_TypeError, "gather() needs a list");
        return NULL;
    }
    const Py_ssize_t n = PyList_GET_SIZE(arg);
    Py_ssize_t total = 0;
    PyObject *offs = PyBytes_FromStringAndSize(NULL, (n + 1) * (Py_ssize_t)sizeof(int64_t));
    if (!offs) return NULL;
    int64_t *off = (int64_t *)PyBytes_AS_STRING(offs);
    for (Py_ssize_t i = 0; i < n; i++) {
        PyObject *it = PyList_GET_ITEM(arg, i);
        if (!PyBytes_CheckExact(it)) {          /* the reference's <bytes> cast (counter.pyx:27) */
            Py_DECREF(offs);
            PyErr_Format(PyExc_TypeError, "expected bytes, %s found", Py_TYPE(it)->tp_name);
            return NULL;
        }
        off[i] = total;
        total += PyBytes_GET_SIZE(it);
    }
    off[n] = total;
    PyObject *buf = PyBytes_FromStringAndSize(NULL, total);
    if (!buf) { Py_DECREF(offs); return NULL; }
    char *dst = PyBytes_AS_STRING(buf);
    for (Py_ssize_t i = 0; i < n; i++) {
        PyObject *it = PyList_GET_ITEM(arg, i);
        memcpy(dst + off[i], PyBytes_AS_STRING(it), (size_t)PyBytes_GET_SIZE(it));
    }
    PyObject *res = PyTuple_Pack(2, buf, offs);
    Py_DECREF(buf);
    Py_DECREF(offs);
    return res;
}

static Py_ssize_t slot_offset(PyObject *cls, const char *name)
{
    PyObject *d = PyObject_GetAttrString(cls, name);
    if (!d) return -1;
    if (!Py_IS_TYPE(d, &PyMemberDescr_Type)) {
        Py_DECREF(d);
        PyErr_Format(PyExc_TypeError, "%s is not a slot", name);
        return -1;
    }
    const Py_ssize_t o = ((PyMemberDescrObject *)d)->d_member->offset;
    Py_DECREF(d);
    return o;
}

/* box_many(cls, words: buffer of uint64 [n * W], lens: buffer of int64 [n], W) */
static PyObject *box_many(PyObject *self, PyObject *args)
{
    PyObject *cls;
    Py_buffer wb, lb;
    int W;
    if (!PyArg_ParseTuple(args, "Oy*y*i", &cls, &wb, &lb, &W)) return NULL;
    PyObject *out = NULL;
    const Py_ssize_t n = lb.len / (Py_ssize_t)sizeof(int64_t);
    if (!PyType_Check(cls) || W < 1 || wb.len != n * W * (Py_ssize_t)sizeof(uint64_t)) {
        PyErr_SetString(PyExc_ValueError, "box_many: bad arguments");
        goto done;
    }
    {
        const Py_ssize_t o_packed = slot_offset(cls, "_packed"), o_len = slot_offset(cls, "_length"), o_hash = slot_offset(cls, "_hash");
        if (o_packed < 0 || o_len < 0 || o_hash < 0) goto done;
        PyTypeObject *tp = (PyTypeObject *)cls;
        const uint64_t *w = (const uint64_t *)wb.buf;
        const int64_t *l = (const int64_t *)lb.buf;
        out = PyList_New(n);
        if (!out) goto done;
        /* none of these objects can be part of a reference cycle; without this the collector rescans the growing heap */
        const int gc_was = PyGC_Disable();
        for (Py_ssize_t i = 0; i < n; i++) {
            PyObject *obj = tp->tp_alloc(tp, 0);
            PyObject *packed = obj ? PyTuple_New(W) : NULL;
            if (!packed) { Py_XDECREF(obj); Py_CLEAR(out); if (gc_was) PyGC_Enable(); goto done; }
            for (int j = 0; j < W; j++) PyTuple_SET_ITEM(packed, j, PyLong_FromUnsignedLongLong(w[i * W + j]));
            int64_t h = (int64_t)w[i * W];        /* prehash = first block as Py_hash_t, -1 -> -2 (short_seq_64.pyx:35-36) */
            if (h == -1) h = -2;
            *(PyObject **)((char *)obj + o_packed) = packed;
            *(PyObject **)((char *)obj + o_len) = PyLong_FromLongLong(l[i]);
            if (h >= 0) {                         /* the hash IS the first block: share the int object */
                PyObject *first = PyTuple_GET_ITEM(packed, 0);
                Py_INCREF(first);
                *(PyObject **)((char *)obj + o_hash) = first;
            } else {
                *(PyObject **)((char *)obj + o_hash) = PyLong_FromLongLong(h);
            }
            PyList_SET_ITEM(out, i, obj);
        }
        if (gc_was) PyGC_Enable();
    }
done:
    PyBuffer_Release(&wb);
    PyBuffer_Release(&lb);
    return out;
}

/* fill_counts(dict, objs: list, counts: buffer int64 [n], hashes: buffer int64 [n], order: buffer int64 [m], add: bool) */
/* private but exported CPython API (the reference uses the same calls, counter.pxd:31-37) */
extern int _PyDict_SetItem_KnownHash(PyObject *mp, PyObject *key, PyObject *item, Py_hash_t hash);
extern PyObject *_PyDict_GetItem_KnownHash(PyObject *mp, PyObject *key, Py_hash_t hash);
extern PyObject *_PyDict_NewPresized(Py_ssize_t minused);

static PyObject *fill_counts(PyObject *self, PyObject *args)
{
    PyObject *d, *objs;
    Py_buffer cb, hb, ob;
    int add;
    if (!PyArg_ParseTuple(args, "O!O!y*y*y*p", &PyDict_Type, &d, &PyList_Type, &objs, &cb, &hb, &ob, &add)) return NULL;
    PyObject *res = NULL;
    const Py_ssize_t n = PyList_GET_SIZE(objs), m = ob.len / (Py_ssize_t)sizeof(int64_t);
    const int64_t *cnt = (const int64_t *)cb.buf, *hs = (const int64_t *)hb.buf, *ord = (const int64_t *)ob.buf;
    if (cb.len != n * (Py_ssize_t)sizeof(int64_t) || hb.len != cb.len) { PyErr_SetString(PyExc_ValueError, "fill_counts: bad arguments"); goto done; }
    {
        const int gc_was = PyGC_Disable();
        int failed = 0;
        /* An empty destination is filled through a dict presized for all m keys (no rehash on the way up) that is then
         * merged in: merging a clean, freshly built dict into an empty one clones its table. */
        PyObject *target = d;
        if (!add && PyDict_GET_SIZE(d) == 0 && m > 1024) {
            target = _PyDict_NewPresized(m);
            if (!target) { if (gc_was) PyGC_Enable(); goto done; }
        }
        for (Py_ssize_t j = 0; j < m && !failed; j++) {
            const int64_t i = ord[j];
            if (i < 0 || i >= n) { PyErr_SetString(PyExc_IndexError, "fill_counts: order out of range"); failed = 1; break; }
            PyObject *key = PyList_GET_ITEM(objs, i);
            const Py_hash_t h = (Py_hash_t)hs[i];            /* = hash(key): the first block, -1 -> -2; no Python-level __hash__ call */
            int64_t c = cnt[i];
            if (add) {
                PyObject *old = _PyDict_GetItem_KnownHash(d, key, h);
                if (!old && PyErr_Occurred()) { failed = 1; break; }
                if (old) c += PyLong_AsLongLong(old);
            }
            PyObject *val = PyLong_FromLongLong(c);
            if (!val) { failed = 1; break; }
            if (_PyDict_SetItem_KnownHash(target, key, val, h) < 0) failed = 1;
            Py_DECREF(val);
        }
        if (target != d) {
            if (!failed && PyDict_Update(d, target) < 0) failed = 1;
            Py_DECREF(target);
        }
        if (gc_was) PyGC_Enable();
        if (failed) goto done;
    }
    res = Py_None;
    Py_INCREF(res);
done:
    PyBuffer_Release(&cb);
    PyBuffer_Release(&hb);
    PyBuffer_Release(&ob);
    return res;
}

static PyMethodDef methods[] = {
    {"gather", gather, METH_O, "list of bytes -> (ascii bytes, int64 offsets bytes)"},
    {"box_many", box_many, METH_VARARGS, "box packed keys into ShortSeq objects"},
    {"fill_counts", fill_counts, METH_VARARGS, "insert (object, count) pairs into a dict in a given order"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_fastbox", "host-side helpers of shortseq_b200", -1, methods};

PyMODINIT_FUNC PyInit__fastbox(void) { return PyModule_Create(&moddef); }
