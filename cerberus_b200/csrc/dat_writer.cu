// Host-side serialiser of WSI instance tables into the reference's `.dat` format.
//
// infer/wsi.py:853 writes  joblib.dump({"Nuclei": {uuid: {"box", "centroid", "contour", "prob",
// "type"}}, ...})  - for a 20000^2 slide half a million instances, i.e. 1.6 million small numpy
// arrays; pickling them object by object costs 8 s (joblib: over a minute). The instances live
// here as a few big arrays (the device tables of csrc/instinfo.cu), so the pickle STREAM of the
// dict items is produced directly from those arrays: one pass, no Python objects. The stream uses
// protocol-2 opcodes only and refers to the objects every array shares (numpy's _reconstruct,
// ndarray, the two dtypes, the key strings) through memo slots the caller defines in a preamble
// (cerberus_b200/infer/dat_writer.py); pickle.load / joblib.load read it back as the same dict a
// plain pickle.dump of the Python objects gives (tests/test_dat_writer.py).
#include <cstdint>
#include <cstring>

#include "capi_internal.cuh"

namespace {

struct Out {
  uint8_t* p;
  int64_t cap, n;
  void put(const void* src, int64_t len) {
    if (p && n + len <= cap) std::memcpy(p + n, src, static_cast<size_t>(len));
    n += len;
  }
  void op(uint8_t c) { put(&c, 1); }
  void get(uint8_t memo) {  // BINGET
    const uint8_t b[2] = {'h', memo};
    put(b, 2);
  }
  void u32(uint32_t v) { put(&v, 4); }  // little endian hosts only (x86-64 / aarch64)
  void small_int(int64_t v) {          // BININT1 / BININT / LONG1
    if (v >= 0 && v < 256) {
      const uint8_t b[2] = {'K', static_cast<uint8_t>(v)};
      put(b, 2);
    } else if (v >= INT32_MIN && v <= INT32_MAX) {
      op('J');
      const int32_t x = static_cast<int32_t>(v);
      put(&x, 4);
    } else {
      const uint8_t b[2] = {0x8a, 8};  // LONG1, 8 bytes two's complement little endian
      put(b, 2);
      put(&v, 8);
    }
  }
  // numpy array: reconstruct(ndarray, (0,), b"b") + state (1, shape, dtype, False, bytes)
  void array_head(const uint8_t* memo) {
    get(memo[0]);  // _reconstruct
    get(memo[1]);  // ndarray
    get(memo[2]);  // (0,)
    get(memo[3]);  // b"b"
    op(0x87);      // TUPLE3
    op('R');       // REDUCE
    op('(');       // MARK (state tuple)
    small_int(1);
  }
  void array_tail(uint8_t dtype_memo, const void* data, int64_t bytes) {
    get(dtype_memo);
    op(0x89);  // NEWFALSE (is_fortran)
    op('B');   // BINBYTES
    u32(static_cast<uint32_t>(bytes));
    put(data, bytes);
    op('t');  // TUPLE
    op('b');  // BUILD
  }
};

}  // namespace

// memo[0..3] = _reconstruct, ndarray, (0,), b"b"; memo[4] = dtype int64, memo[5] = dtype float64;
// memo[6..10] = the key strings "box", "centroid", "contour", "prob", "type".
// prob / type may be NULL (the value None is written, as tiatoolbox does without a type map).
// Returns the number of bytes of the stream (written only if out != NULL and out_cap suffices).
extern "C" int64_t cerb_pickle_instances(const char* uid_hex, const int64_t* box, const double* centroid,
                                         const int64_t* contour_off, const int64_t* contour_xy,
                                         const double* prob, const int64_t* type, int64_t n,
                                         const uint8_t* memo, uint8_t* out, int64_t out_cap) {
  if (!uid_hex || !box || !centroid || !contour_off || !contour_xy || !memo || n < 0) return -1;
  Out o{out, out_cap, 0};
  for (int64_t i = 0; i < n; ++i) {
    if (i % 1000 == 0 && i > 0) {  // batches of 1000 items per SETITEMS, like pickle's batch_setitems
      o.op('u');
      o.op('(');
    }
    o.op('X');  // BINUNICODE
    o.u32(32);
    o.put(uid_hex + 32 * i, 32);
    o.op('}');  // EMPTY_DICT
    o.op('(');
    // box: int64 [4]
    o.get(memo[6]);
    o.array_head(memo);
    o.small_int(4);
    o.op(0x85);  // TUPLE1
    o.array_tail(memo[4], box + 4 * i, 32);
    // centroid: float64 [2]
    o.get(memo[7]);
    o.array_head(memo);
    o.small_int(2);
    o.op(0x85);
    o.array_tail(memo[5], centroid + 2 * i, 16);
    // contour: int64 [k, 2]
    const int64_t k = contour_off[i + 1] - contour_off[i];
    if (k < 0 || k * 16 > 0x7fffffffLL) return -2;
    o.get(memo[8]);
    o.array_head(memo);
    o.small_int(k);
    o.small_int(2);
    o.op(0x86);  // TUPLE2
    o.array_tail(memo[4], contour_xy + 2 * contour_off[i], k * 16);
    // prob: float or None
    o.get(memo[9]);
    if (prob) {
      o.op('G');  // BINFLOAT: 8 bytes big endian
      uint64_t bits;
      std::memcpy(&bits, prob + i, 8);
      uint8_t be[8];
      for (int b = 0; b < 8; ++b) be[b] = static_cast<uint8_t>(bits >> (56 - 8 * b));
      o.put(be, 8);
    } else {
      o.op('N');
    }
    // type: int or None
    o.get(memo[10]);
    if (type) o.small_int(type[i]);
    else o.op('N');
    o.op('u');  // SETITEMS of the instance dict
  }
  return o.n;
}
