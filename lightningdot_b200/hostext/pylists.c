/* Host-side helper of DenseFlatIndexer.search_knn (dvl/indexer/faiss_indexers.py:85-87):
 *     db_ids = [[self.index_id_to_db_id[i] for i in query_top_idxs] for query_top_idxs in indexes]
 * nq * k Python object references (1 M at the bench's 10 000 x 100) built by two nested list comprehensions in the
 * reference.  The search itself takes a few milliseconds on the GPU, so this materialisation is what the caller of
 * search_knn waits for; here it is one C loop over the int64 label matrix: PyList_New + PyList_SET_ITEM, no intermediate
 * object array, no per-element interpreter dispatch.  Negative labels index from the end, exactly like the list indexing of
 * the reference (faiss label -1 for a short index -> the LAST id).
 *
 * Not part of the C ABI of include/ldot.h (that library has no Python dependency): a separate small shared object,
 * loaded with ctypes.PyDLL (the GIL stays held).  Built by lightningdot_b200/build.py with the host compiler. */
#include <Python.h>

PyObject* ldot_py_gather_lists(PyObject* ids, const long long* idx, long long nq, long long k) {
  if (!PyList_Check(ids)) {
    PyErr_SetString(PyExc_TypeError, "index_id_to_db_id must be a list");
    return NULL;
  }
  const Py_ssize_t n = PyList_GET_SIZE(ids);
  PyObject* out = PyList_New((Py_ssize_t)nq);
  if (!out) return NULL;
  for (long long q = 0; q < nq; ++q) {
    PyObject* row = PyList_New((Py_ssize_t)k);
    if (!row) {
      Py_DECREF(out);
      return NULL;
    }
    PyList_SET_ITEM(out, (Py_ssize_t)q, row);   /* (owned by `out` from here: one DECREF releases everything) */
    const long long* r = idx + q * k;
    for (long long j = 0; j < k; ++j) {
      long long i = r[j];
      if (i < 0) i += n;
      if (i < 0 || i >= n) {
        /* the unfilled slots of `row` are NULL, which list deallocation accepts */
        Py_DECREF(out);
        PyErr_SetString(PyExc_IndexError, "list index out of range");
        return NULL;
      }
      PyObject* o = PyList_GET_ITEM(ids, (Py_ssize_t)i);
      Py_INCREF(o);
      PyList_SET_ITEM(row, (Py_ssize_t)j, o);
    }
  }
  return out;
}
