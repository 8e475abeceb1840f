/* CPython glue of batch_obs (reference: ss_baselines/common/utils.py:129-156): hands the per-env observation arrays of one
 * sensor to the C-ABI library's avl_host_gather without a Python-level loop.  Extracting 128 buffer addresses per step in
 * Python (ndarray.ctypes.data: ~1.8 us each) cost as much as copying the 7.3 MB they point to.
 *
 *   gather(fn_addr, arrays, dst_addr, piece_bytes, format) -> bool
 *     fn_addr      address of avl_host_gather (taken from the ctypes handle of libavlen_b200.so)
 *     arrays       sequence of n objects exporting C-contiguous buffers of exactly piece_bytes bytes and item format `format`
 *     dst_addr     staging buffer; piece i lands at dst_addr + i * piece_bytes
 *   Returns False (nothing copied) when some piece does not fit that description — the caller takes its numpy path. */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <string.h>

typedef int (*gather_fn)(const void* const*, void* const*, const long long*, int);

static const char* skip_order(const char* f) {
  return (f && (*f == '<' || *f == '=' || *f == '@' || *f == '|')) ? f + 1 : f;
}

static PyObject* py_gather(PyObject* self, PyObject* args) {
  unsigned long long fn_addr, dst_addr;
  PyObject* seq;
  long long piece;
  const char* format;
  if (!PyArg_ParseTuple(args, "KOKLs", &fn_addr, &seq, &dst_addr, &piece, &format)) return NULL;
  PyObject* fast = PySequence_Fast(seq, "arrays must be a sequence");
  if (!fast) return NULL;
  const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
  if (n == 0 || piece <= 0) {
    Py_DECREF(fast);
    Py_RETURN_FALSE;
  }
  Py_buffer* views = (Py_buffer*)PyMem_Malloc(sizeof(Py_buffer) * (size_t)n);
  const void** src = (const void**)PyMem_Malloc(sizeof(void*) * (size_t)n);
  void** dst = (void**)PyMem_Malloc(sizeof(void*) * (size_t)n);
  long long* bytes = (long long*)PyMem_Malloc(sizeof(long long) * (size_t)n);
  if (!views || !src || !dst || !bytes) {
    PyMem_Free(views); PyMem_Free(src); PyMem_Free(dst); PyMem_Free(bytes);
    Py_DECREF(fast);
    return PyErr_NoMemory();
  }
  const char* want = skip_order(format);
  Py_ssize_t got = 0;
  int ok = 1;
  for (; got < n; ++got) {
    PyObject* item = PySequence_Fast_GET_ITEM(fast, got);
    if (PyObject_GetBuffer(item, &views[got], PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) != 0) {
      PyErr_Clear();
      ok = 0;
      break;
    }
    const char* have = skip_order(views[got].format ? views[got].format : "B");
    if (views[got].len != (Py_ssize_t)piece || strcmp(have, want) != 0) {
      ++got;  /* this view is held and must be released */
      ok = 0;
      break;
    }
    src[got] = views[got].buf;
    dst[got] = (void*)(uintptr_t)(dst_addr + (unsigned long long)got * (unsigned long long)piece);
    bytes[got] = piece;
  }
  int rc = 0;
  if (ok) {
    gather_fn fn = (gather_fn)(uintptr_t)fn_addr;
    Py_BEGIN_ALLOW_THREADS
    rc = fn(src, dst, bytes, (int)n);
    Py_END_ALLOW_THREADS
  }
  for (Py_ssize_t i = 0; i < got; ++i) PyBuffer_Release(&views[i]);
  PyMem_Free(views); PyMem_Free(src); PyMem_Free(dst); PyMem_Free(bytes);
  Py_DECREF(fast);
  if (ok && rc != 0) {
    PyErr_Format(PyExc_RuntimeError, "avl_host_gather failed with status %d", rc);
    return NULL;
  }
  if (ok) Py_RETURN_TRUE;
  Py_RETURN_FALSE;
}

static PyMethodDef methods[] = {{"gather", py_gather, METH_VARARGS, "gather(fn_addr, arrays, dst_addr, piece_bytes, format) -> bool"},
                                {NULL, NULL, 0, NULL}};
static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "_pyhost", "batch_obs glue over avl_host_gather", -1, methods};
PyMODINIT_FUNC PyInit__pyhost(void) { return PyModule_Create(&moduledef); }
