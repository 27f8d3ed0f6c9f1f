/* CPython side door into rebert_recommend_host for the Python mirror (robot_ebert_b200/catalog.py).
 *
 * The request path of a small catalog is latency-bound end to end (2269 x 32, the reference's production shape: 24 us on
 * the device), and calling a 23-argument C function through ctypes costs 5-6 us per request plus ~1 us for every numpy
 * array whose address is looked up (tools/latency_breakdown.py).  This module does nothing but that call: it takes the
 * request's arrays through the buffer protocol, the structs by address, releases the GIL around the call exactly as ctypes
 * does, and returns (status, count).  The arithmetic, the kernels and the C ABI are untouched — the entry point is handed
 * over as an address taken from the already loaded librebert_b200.so, so there is no second copy of the library and no
 * link-time dependency.  When the module is absent the Python layer makes the same call through ctypes.
 * Built by robot_ebert_b200/build.py with the host compiler (plain C, no CUDA). */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

typedef int (*recommend_host_fn)(const void* cat, const float* query, const int32_t* liked_rows, const float* liked_w, int32_t n_liked,
                                 const int32_t* exclude_rows, int32_t n_exclude, const void* device_filter, int32_t k, int32_t kc,
                                 int32_t n_liked_cap, int32_t n_exclude_cap, void* pinned, size_t pinned_bytes, void* device_scratch,
                                 size_t device_bytes, const void* proof, const void* exchange, int64_t* out_rows, double* out_scores,
                                 int32_t* out_count, void* info, void* stream);

static recommend_host_fn g_entry = NULL;

static PyObject* set_entry(PyObject* self, PyObject* arg) {
    (void)self;
    const unsigned long long a = PyLong_AsUnsignedLongLong(arg);
    if (a == (unsigned long long)-1 && PyErr_Occurred()) return NULL;
    g_entry = (recommend_host_fn)(uintptr_t)a;
    Py_RETURN_NONE;
}

/* A request array: None, or a C-contiguous buffer of 4-byte items.  Returns 0 on success. */
static int take(PyObject* o, Py_buffer* view, const void** ptr, Py_ssize_t* count, const char* what) {
    *ptr = NULL;
    *count = 0;
    view->obj = NULL;
    if (o == Py_None) return 0;
    if (PyObject_GetBuffer(o, view, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) != 0) return -1;
    if (view->itemsize != 4) {
        PyErr_Format(PyExc_TypeError, "%s must be a contiguous array of 4-byte items", what);
        PyBuffer_Release(view);
        view->obj = NULL;
        return -1;
    }
    *ptr = view->buf;
    *count = view->len / 4;
    return 0;
}

/* recommend_host(cat, query, liked, weights, exclude, filter, k, kc, n_liked_cap, n_exclude_cap, pinned, pinned_bytes,
 *                device, device_bytes, proof, exchange, out_rows, out_scores, info, stream) -> (status, count)
 * cat / filter / pinned / device / proof / exchange / out_rows / out_scores / info / stream are addresses (0 = NULL);
 * query is float32 [d]; liked / exclude int32; weights float32 (same length as liked). */
static PyObject* recommend_host(PyObject* self, PyObject* const* args, Py_ssize_t nargs) {
    (void)self;
    if (!g_entry) { PyErr_SetString(PyExc_RuntimeError, "pycall: entry point not set"); return NULL; }
    if (nargs != 20) { PyErr_SetString(PyExc_TypeError, "pycall.recommend_host takes 20 arguments"); return NULL; }
    unsigned long long a[20];
    static const int is_addr_or_int[20] = {1, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
    for (int i = 0; i < 20; ++i) {
        a[i] = 0;
        if (!is_addr_or_int[i]) continue;
        a[i] = PyLong_AsUnsignedLongLong(args[i]);
        if (a[i] == (unsigned long long)-1 && PyErr_Occurred()) return NULL;
    }
    Py_buffer vq, vl, vw, ve;
    const void *pq, *pl, *pw, *pe;
    Py_ssize_t nq, nl, nw, ne;
    if (take(args[1], &vq, &pq, &nq, "query") != 0) return NULL;
    if (take(args[2], &vl, &pl, &nl, "liked_rows") != 0) { if (vq.obj) PyBuffer_Release(&vq); return NULL; }
    if (take(args[3], &vw, &pw, &nw, "weights") != 0) { if (vq.obj) PyBuffer_Release(&vq); if (vl.obj) PyBuffer_Release(&vl); return NULL; }
    if (take(args[4], &ve, &pe, &ne, "exclude_rows") != 0) {
        if (vq.obj) PyBuffer_Release(&vq);
        if (vl.obj) PyBuffer_Release(&vl);
        if (vw.obj) PyBuffer_Release(&vw);
        return NULL;
    }
    int rc = 0;
    int32_t count = 0;
    if (pw && nw != nl) {
        PyErr_SetString(PyExc_ValueError, "weights must match liked_rows");
        rc = -1;
    } else {
        Py_BEGIN_ALLOW_THREADS
        rc = g_entry((const void*)(uintptr_t)a[0], (const float*)pq, (const int32_t*)pl, (const float*)pw, (int32_t)nl, (const int32_t*)pe,
                     (int32_t)ne, (const void*)(uintptr_t)a[5], (int32_t)a[6], (int32_t)a[7], (int32_t)a[8], (int32_t)a[9],
                     (void*)(uintptr_t)a[10], (size_t)a[11], (void*)(uintptr_t)a[12], (size_t)a[13], (const void*)(uintptr_t)a[14],
                     (const void*)(uintptr_t)a[15], (int64_t*)(uintptr_t)a[16], (double*)(uintptr_t)a[17], &count, (void*)(uintptr_t)a[18],
                     (void*)(uintptr_t)a[19]);
        Py_END_ALLOW_THREADS
    }
    if (vq.obj) PyBuffer_Release(&vq);
    if (vl.obj) PyBuffer_Release(&vl);
    if (vw.obj) PyBuffer_Release(&vw);
    if (ve.obj) PyBuffer_Release(&ve);
    if (rc == -1 && PyErr_Occurred()) return NULL;
    return Py_BuildValue("ii", rc, (int)count);
}

static PyMethodDef methods[] = {
    {"set_entry", (PyCFunction)set_entry, METH_O, "address of rebert_recommend_host in the loaded library"},
    {"recommend_host", (PyCFunction)(void (*)(void))recommend_host, METH_FASTCALL, "one request through rebert_recommend_host"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "_pycall", "fast call into rebert_recommend_host", -1, methods, NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__pycall(void) { return PyModule_Create(&moduledef); }
