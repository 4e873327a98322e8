/* One-pass tensorizer of the reference's list-of-dict feature batches — host side, plain C on the
 * CPython API, loaded with ctypes.PyDLL (the GIL is held; no numpy C API: outputs are raw buffers of numpy arrays).
 *
 * The reference walks the B x L token dicts once PER FEATURE in Python (feat2tensor, model/BaseLine/model.py:186-224,
 * 22 / 14 times per call, + the mm fill of model.py:283-296): 84 % of its CPU forward (SURVEY probe7). packed.py's
 * vectorised-numpy version still costs 3-5 us per token; this walk costs one dict lookup per (token, feature).
 *
 * Semantics are exactly packed.pack_from_dicts' (kept as the Python restatement the tests compare against):
 *   - sparse feature:  ids[t, col] = int(tok[key])                    (KeyError if the dict lacks the key, model.py:222)
 *   - array feature:   the token's list with padding ids (0) dropped   (row 0 is the all-zero padding row)
 *   - mm feature:      tok[key] if present else zeros                  (model.py:288-293)
 * A token is a dict (fast path) or any mapping; a row is a list or any sequence (np.array(dicts, dtype=object)).
 * Every function returns 0 / a count on success and -1 with a Python exception set.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

static PyObject* tok_get(PyObject* tok, PyObject* key, int required) {
  /* borrowed-or-new reference normalised to NEW (caller decrefs); NULL + exception, or NULL without exception when
   * the key is absent and not required */
  PyObject* v;
  if (PyDict_Check(tok)) {
    v = PyDict_GetItemWithError(tok, key);
    if (v) { Py_INCREF(v); return v; }
    if (PyErr_Occurred()) return NULL;
    if (required) PyErr_SetObject(PyExc_KeyError, key);
    return NULL;
  }
  if (!required) {
    int has = PySequence_Contains(tok, key);   /* `k in tok` */
    if (has < 0) return NULL;
    if (!has) return NULL;
  }
  return PyObject_GetItem(tok, key);
}

static int as_i32(PyObject* v, int32_t* out) {
  long x;
  if (PyLong_Check(v)) {
    x = PyLong_AsLong(v);
  } else {
    PyObject* i = PyNumber_Index(v);   /* numpy integer scalars */
    if (!i) {
      PyErr_Clear();
      i = PyNumber_Long(v);            /* what np.array(..., dtype=int64) would accept (e.g. a float-typed id) */
      if (!i) return -1;
    }
    x = PyLong_AsLong(i);
    Py_DECREF(i);
  }
  if (x == -1 && PyErr_Occurred()) return -1;
  if (x < INT32_MIN || x > INT32_MAX) {
    PyErr_SetString(PyExc_OverflowError, "feature id does not fit int32");
    return -1;
  }
  *out = (int32_t)x;
  return 0;
}

/* rows of the batch as fast sequences; fills row_items[b] (borrowed arrays valid while seqs[b] lives) */
typedef struct { PyObject* seq; PyObject** items; Py_ssize_t n; } row_t;

static int open_row(PyObject* feature_array, Py_ssize_t b, Py_ssize_t L, row_t* r) {
  PyObject* row = PySequence_GetItem(feature_array, b);
  if (!row) return -1;
  r->seq = PySequence_Fast(row, "a feature sequence must be a sequence of token dicts");
  Py_DECREF(row);
  if (!r->seq) return -1;
  r->n = PySequence_Fast_GET_SIZE(r->seq);
  r->items = PySequence_Fast_ITEMS(r->seq);
  if (r->n != L) {
    Py_DECREF(r->seq);
    r->seq = NULL;
    PyErr_SetString(PyExc_ValueError, "setting an array element with a sequence: ragged feature sequences");
    return -1;
  }
  return 0;
}

/* ids[(b*L + l) * n_single + cols[j]] = int(tok[keys[j]]) for every token and every key */
int tgr_pack_single(PyObject* feature_array, long B, long L, PyObject* keys, const int32_t* cols, long n_keys, long n_single,
                    int32_t* ids) {
  if (!PyTuple_Check(keys) || PyTuple_GET_SIZE(keys) != n_keys) {
    PyErr_SetString(PyExc_TypeError, "keys must be a tuple of n_keys strings");
    return -1;
  }
  for (long b = 0; b < B; ++b) {
    row_t r;
    if (open_row(feature_array, b, L, &r)) return -1;
    for (long l = 0; l < L; ++l) {
      PyObject* tok = r.items[l];
      int32_t* dst = ids + ((size_t)b * L + l) * n_single;
      for (long j = 0; j < n_keys; ++j) {
        PyObject* v = tok_get(tok, PyTuple_GET_ITEM(keys, j), 1);
        if (!v) { Py_DECREF(r.seq); return -1; }
        const int rc = as_i32(v, dst + cols[j]);
        Py_DECREF(v);
        if (rc) { Py_DECREF(r.seq); return -1; }
      }
    }
    Py_DECREF(r.seq);
  }
  return 0;
}

/* One array feature in one walk: cnt[t] = number of non-zero values of token t, the values themselves back to back
 * (token order) into out[0..cap). Returns the total; if it exceeds cap the tail was counted but not written and the
 * caller repeats the call with a buffer of that size. */
long long tgr_pack_array(PyObject* feature_array, long B, long L, PyObject* key, int32_t* cnt, int32_t* out, long long cap) {
  long long total = 0;
  for (long b = 0; b < B; ++b) {
    row_t r;
    if (open_row(feature_array, b, L, &r)) return -1;
    for (long l = 0; l < L; ++l) {
      PyObject* v = tok_get(r.items[l], key, 1);
      if (!v) { Py_DECREF(r.seq); return -1; }
      PyObject* lst = PySequence_Fast(v, "an array feature must be a sequence of ids");
      Py_DECREF(v);
      if (!lst) { Py_DECREF(r.seq); return -1; }
      const Py_ssize_t m = PySequence_Fast_GET_SIZE(lst);
      PyObject** it = PySequence_Fast_ITEMS(lst);
      int32_t c = 0;
      for (Py_ssize_t i = 0; i < m; ++i) {
        int32_t x;
        if (as_i32(it[i], &x)) { Py_DECREF(lst); Py_DECREF(r.seq); return -1; }
        if (x != 0) {
          if (out && total + c < cap) out[total + c] = x;
          ++c;
        }
      }
      Py_DECREF(lst);
      if (cnt) cnt[(size_t)b * L + l] = c;
      total += c;
    }
    Py_DECREF(r.seq);
  }
  return total;
}

/* out[(b*L + l), 0:dim] = tok[key] when present (out is pre-zeroed). Contiguous float32 buffers are copied directly. */
int tgr_pack_mm(PyObject* feature_array, long B, long L, PyObject* key, long dim, float* out) {
  for (long b = 0; b < B; ++b) {
    row_t r;
    if (open_row(feature_array, b, L, &r)) return -1;
    for (long l = 0; l < L; ++l) {
      PyObject* v = tok_get(r.items[l], key, 0);
      if (!v) {
        if (PyErr_Occurred()) { Py_DECREF(r.seq); return -1; }
        continue;   /* absent: zeros */
      }
      if (v == Py_None) { Py_DECREF(v); continue; }
      float* dst = out + ((size_t)b * L + l) * dim;
      int done = 0;
      if (PyObject_CheckBuffer(v)) {
        Py_buffer view;
        if (PyObject_GetBuffer(v, &view, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) == 0) {
          if (view.format && strcmp(view.format, "f") == 0 && view.itemsize == 4) {
            if (view.len != (Py_ssize_t)dim * 4) {
              const Py_ssize_t got = view.len / 4;
              PyBuffer_Release(&view);
              Py_DECREF(v);
              Py_DECREF(r.seq);
              PyErr_Format(PyExc_ValueError, "mm feature vector has %zd elements, expected %ld", got, dim);
              return -1;
            }
            memcpy(dst, view.buf, (size_t)dim * 4);
            done = 1;
          }
          PyBuffer_Release(&view);
        } else {
          PyErr_Clear();
        }
      }
      if (!done) {   /* any other sequence of numbers */
        PyObject* lst = PySequence_Fast(v, "an mm feature must be a vector");
        if (!lst) { Py_DECREF(v); Py_DECREF(r.seq); return -1; }
        if (PySequence_Fast_GET_SIZE(lst) != dim) {
          Py_DECREF(lst); Py_DECREF(v); Py_DECREF(r.seq);
          PyErr_Format(PyExc_ValueError, "mm feature vector has the wrong length, expected %ld", dim);
          return -1;
        }
        PyObject** it = PySequence_Fast_ITEMS(lst);
        for (long i = 0; i < dim; ++i) {
          const double x = PyFloat_AsDouble(it[i]);
          if (x == -1.0 && PyErr_Occurred()) { Py_DECREF(lst); Py_DECREF(v); Py_DECREF(r.seq); return -1; }
          dst[i] = (float)x;
        }
        Py_DECREF(lst);
      }
      Py_DECREF(v);
    }
    Py_DECREF(r.seq);
  }
  return 0;
}

int tgr_pack_abi_version(void) { return 1; }
