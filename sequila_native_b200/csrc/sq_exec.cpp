// Host-side exec node of the `Cuda` interval join (declared in include/sequila_exec.h): the C++
// counterpart of IntervalJoinExec / IntervalJoinStream (reference interval_join.rs) above the C ABI
// of include/sequila_cuda.h.  It only marshals Arrow buffers: hashing of the `on` keys, the i32
// view of the interval columns, and one sq_* call per step; searching, emitting and `take` run on
// the GPU.
#include "sequila_exec.h"

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <system_error>
#include <thread>
#include <vector>

#include "sequila_cuda.h"
#include "sq_keyhash.h"

#define SQ_API extern "C" __attribute__((visibility("default")))

namespace {

using Clock = std::chrono::steady_clock;

// option cuda_exec_trace=1: per-phase wall times of every probe batch on stderr (development aid)
struct Trace {
  bool on;
  Clock::time_point t;
  explicit Trace(sq_ctx* ctx) : on(false), t(Clock::now()) {
    char v[8] = {0};
    on = sq_ctx_get_option(ctx, "cuda_exec_trace", v, sizeof v) == SQ_OK && v[0] == '1';
  }
  void lap(const char* what) {
    if (!on) return;
    const auto n = Clock::now();
    fprintf(stderr, "[sq_exec] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
    t = n;
  }
};

// std::vector storage in pinned host memory from the library's pool.  The probe columns of a tile (key hashes, starts,
// ends, the columns the probe side contributes to the output) go to the device with cudaMemcpyAsync: from pageable memory
// that copy is staged through the driver's bounce buffer at a fraction of the link rate and is not asynchronous at all.
template <class T>
struct PinnedAlloc {
  using value_type = T;
  sq_ctx* ctx = nullptr;
  explicit PinnedAlloc(sq_ctx* c) : ctx(c) {}
  template <class U> PinnedAlloc(const PinnedAlloc<U>& o) : ctx(o.ctx) {}
  T* allocate(size_t n) {
    void* p = nullptr;
    if (sq_host_alloc(ctx, n * sizeof(T), &p) != SQ_OK) throw std::bad_alloc();
    return static_cast<T*>(p);
  }
  void deallocate(T* p, size_t) { sq_host_free(ctx, p); }
  template <class U> bool operator==(const PinnedAlloc<U>& o) const { return ctx == o.ctx; }
  template <class U> bool operator!=(const PinnedAlloc<U>& o) const { return ctx != o.ctx; }
};
template <class T> using PinnedVec = std::vector<T, PinnedAlloc<T>>;

using sqkey::mix64;  // the key hash is defined once, in sq_keyhash.h (shared with the device-side scanner)

enum class Kind { Fixed, Utf8, LargeUtf8 };

struct ColType {
  std::string format, name;
  int64_t flags = 0;
  Kind kind = Kind::Fixed;
  uint32_t width = 0;  // bytes per value for Kind::Fixed
  // dictionary-encoded strings (DictionaryArray<intN, Utf8 | LargeUtf8>): kind = Fixed over the index type (the
  // indices are gathered like any fixed-width column), the values travel once per batch as the dictionary
  bool dict = false;
  bool dict_large = false;  // dictionary values are LargeUtf8
};

// fixed-width Arrow primitive formats -> byte width (Columnar format spec, "Data type description")
bool parse_format(const char* f, ColType* t) {
  t->format = f;
  if (!strcmp(f, "u")) { t->kind = Kind::Utf8; return true; }
  if (!strcmp(f, "U")) { t->kind = Kind::LargeUtf8; return true; }
  t->kind = Kind::Fixed;
  if (!strcmp(f, "c") || !strcmp(f, "C")) t->width = 1;
  else if (!strcmp(f, "s") || !strcmp(f, "S") || !strcmp(f, "e")) t->width = 2;
  else if (!strcmp(f, "i") || !strcmp(f, "I") || !strcmp(f, "f") || !strncmp(f, "tdD", 3) || !strncmp(f, "tts", 3) ||
           !strncmp(f, "ttm", 3))
    t->width = 4;
  else if (!strcmp(f, "l") || !strcmp(f, "L") || !strcmp(f, "g") || !strncmp(f, "tdm", 3) || !strncmp(f, "ttu", 3) ||
           !strncmp(f, "ttn", 3) || !strncmp(f, "ts", 2) || !strncmp(f, "tD", 2))
    t->width = 8;
  else return false;
  return true;
}

struct Side {
  std::vector<ColType> cols;
};

struct Owned {  // release bookkeeping of an exported array / schema
  std::vector<void*> pinned;
  std::vector<void*> heap;
  std::vector<ArrowArray*> children;
  std::vector<ArrowSchema*> schema_children;
  std::vector<const void*> buffers;
  std::vector<ArrowArray*> child_ptrs;
  std::vector<ArrowSchema*> schema_child_ptrs;
  std::string s1, s2;
  sq_ctx* ctx = nullptr;
};

void release_array(ArrowArray* a) {
  if (!a || !a->release) return;
  auto* o = static_cast<Owned*>(a->private_data);
  for (ArrowArray* c : o->children) {
    if (c->release) c->release(c);
    delete c;
  }
  if (a->dictionary) {  // dictionary values of a dictionary-encoded output column (created by make_dictionary)
    if (a->dictionary->release) a->dictionary->release(a->dictionary);
    delete a->dictionary;
    a->dictionary = nullptr;
  }
  for (void* p : o->pinned) sq_host_free(o->ctx, p);
  for (void* p : o->heap) free(p);
  delete o;
  a->release = nullptr;
}

void release_schema(ArrowSchema* s) {
  if (!s || !s->release) return;
  auto* o = static_cast<Owned*>(s->private_data);
  for (ArrowSchema* c : o->schema_children) {
    if (c->release) c->release(c);
    delete c;
  }
  if (s->dictionary) {
    if (s->dictionary->release) s->dictionary->release(s->dictionary);
    delete s->dictionary;
    s->dictionary = nullptr;
  }
  delete o;
  s->release = nullptr;
}

}  // namespace

// per-partition state of a probe batch that is being emitted in windows (low-memory mode)
struct PartState {
  const ArrowArray* batch = nullptr;
  uint64_t n = 0;
  std::vector<std::pair<uint64_t, uint64_t>> windows;  // (pair offset, pair count), cut at probe-row boundaries
  size_t next = 0;
};

// probe batches of a partition waiting to be joined as ONE tile (sq_exec_probe_push / _pop)
struct Pending {
  std::vector<ArrowArray> batches;  // owned (moved in)
  uint64_t rows = 0;
};

// One coalesced tile on its way through the GPU.  The caller's thread concatenates, hashes and casts the tile (prep),
// a worker thread runs the probe and assembles the output batch; while it does, the caller pushes and preps the next tile.
struct TileJob {
  std::thread th;
  Pending take;             // the batches the tile was made of (owned)
  ArrowArray joined{};      // their concatenation when there is more than one
  const ArrowArray* tile = nullptr;
  PinnedVec<uint64_t> keys;
  PinnedVec<int32_t> start, end;
  ArrowArray out{};
  uint64_t n_pairs = 0;
  uint64_t worker_ns = 0;
  int rc = 0;
  explicit TileJob(sq_ctx* ctx) : keys(PinnedAlloc<uint64_t>(ctx)), start(PinnedAlloc<int32_t>(ctx)), end(PinnedAlloc<int32_t>(ctx)) {}
  ~TileJob() {
    if (th.joinable()) th.join();
    if (joined.release) joined.release(&joined);
    for (ArrowArray& b : take.batches) if (b.release) b.release(&b);
    if (out.release) out.release(&out);  // an output nobody collected
  }
};

struct sq_exec {
  sq_exec_config cfg{};
  std::vector<int32_t> on_left, on_right, projection;
  Side left, right;
  sq_ctx* ctx = nullptr;
  sq_index* index = nullptr;
  std::vector<ArrowArray> build_batches;  // owned (moved in)
  std::vector<int32_t> build_col_id;      // left column -> sq_index column id (or -1 when not projected)
  std::map<int32_t, std::vector<std::string>> build_dicts;  // left dictionary column -> values unified over the build batches
  std::mutex mu;
  std::map<int32_t, sq_stream*> streams;
  std::map<int32_t, PartState> parts;
  std::map<int32_t, Pending> pending;
  std::map<int32_t, std::unique_ptr<TileJob>> jobs;  // partition -> the tile its worker thread is joining
  std::string err;
  uint64_t m[16] = {};
  bool built = false;

  int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    std::lock_guard<std::mutex> g(mu);
    err = buf;
    return code;
  }
};

namespace {

int read_side(sq_exec* e, const ArrowSchema* s, Side* out, const char* which) {
  if (!s || !s->format || strcmp(s->format, "+s")) return e->fail(SQ_EINVAL, "%s schema is not a struct (+s)", which);
  out->cols.resize(size_t(s->n_children));
  for (int64_t i = 0; i < s->n_children; ++i) {
    const ArrowSchema* c = s->children[i];
    ColType t;
    if (!parse_format(c->format, &t)) {
      t.format = c->format;
      t.width = 0;  // unsupported types are only an error if the column is actually used
      t.kind = Kind::Fixed;
    }
    if (c->dictionary) {
      const char* vf = c->dictionary->format ? c->dictionary->format : "";
      const bool idx_ok = t.kind == Kind::Fixed && strlen(c->format) == 1 && strchr("cCsSiI", c->format[0]) != nullptr;
      if (idx_ok && (!strcmp(vf, "u") || !strcmp(vf, "U"))) {
        t.dict = true;
        t.dict_large = !strcmp(vf, "U");
      } else {
        t.width = 0;  // unsupported dictionary type: only an error if the column is used
        t.kind = Kind::Fixed;
      }
    }
    t.name = c->name ? c->name : "";
    t.flags = c->flags;
    out->cols[size_t(i)] = t;
  }
  return SQ_OK;
}

// value buffers of child `col` of a struct batch, honouring offsets
struct ColView {
  const uint8_t* validity = nullptr;
  const uint8_t* values = nullptr;   // fixed: values ; utf8: offsets
  const uint8_t* data = nullptr;     // utf8 bytes
  int64_t offset = 0, length = 0, null_count = 0;
};

ColView view_of(const ArrowArray* batch, int32_t col) {
  const ArrowArray* a = batch->children[col];
  ColView v;
  v.offset = a->offset + 0;
  v.length = a->length;
  v.null_count = a->null_count;
  v.validity = static_cast<const uint8_t*>(a->buffers[0]);
  v.values = static_cast<const uint8_t*>(a->n_buffers > 1 ? a->buffers[1] : nullptr);
  v.data = static_cast<const uint8_t*>(a->n_buffers > 2 ? a->buffers[2] : nullptr);
  return v;
}

inline bool bit_at(const uint8_t* bm, int64_t i) { return (bm[i >> 3] >> (i & 7)) & 1; }

// ---- dictionary-encoded string columns ----------------------------------------------------------------
typedef std::pair<const uint8_t*, int64_t> StrRef;

// index of row i (indices are non-negative; 1, 2 or 4 bytes wide)
inline uint32_t dict_index(const uint8_t* values, int64_t i, uint32_t width) {
  uint32_t v = 0;
  memcpy(&v, values + i * width, width);
  return v;
}

// the values of a dictionary (Utf8 or LargeUtf8 array), honouring its offset
void dict_strings(const ArrowArray* d, bool large, std::vector<StrRef>* out) {
  out->clear();
  if (!d || d->n_buffers < 3) return;
  const uint8_t* data = static_cast<const uint8_t*>(d->buffers[2]);
  for (int64_t k = 0; k < d->length; ++k) {
    int64_t a, z;
    if (large) { const int64_t* o = static_cast<const int64_t*>(d->buffers[1]) + d->offset; a = o[k]; z = o[k + 1]; }
    else { const int32_t* o = static_cast<const int32_t*>(d->buffers[1]) + d->offset; a = o[k]; z = o[k + 1]; }
    out->emplace_back(data + a, z - a);
  }
}

// per-column hash of one string value (sq_keyhash.h `of_string`)
inline uint64_t string_hash(const uint8_t* p, int64_t len) {
  if (len <= 8) {
    uint64_t raw = 0;
    memcpy(&raw, p, size_t(len));
    return sqkey::of_string(raw, 0, uint64_t(len));
  }
  uint64_t h = sqkey::kFnvOffset;  // FNV-1a over the bytes
  for (int64_t k = 0; k < len; ++k) { h ^= p[k]; h *= sqkey::kFnvPrime; }
  return h;
}

// largest index the dictionary's index type can hold
inline uint64_t dict_index_limit(const std::string& format) {
  switch (format[0]) {
    case 'c': return 127ull;
    case 'C': return 255ull;
    case 's': return 32767ull;
    case 'S': return 65535ull;
    case 'i': return 2147483647ull;
    default: return 4294967295ull;
  }
}

// rows [lo, hi) of a range, on up to four host threads when the range is large (key hashing of a coalesced probe
// tile: 4-5 ms per 1M Utf8 rows on one thread was most of the exec node's time per tile)
template <typename F>
void parallel_rows(int64_t n, F&& fn) {
  const int64_t kMin = 1 << 17;
  int T = int(n / kMin);
  if (T > 4) T = 4;
  if (T <= 1) { fn(int64_t(0), n); return; }
  std::vector<std::thread> pool;
  for (int t = 1; t < T; ++t) pool.emplace_back([&, t] { fn(n * t / T, n * (t + 1) / T); });
  fn(int64_t(0), n / T);
  for (auto& th : pool) th.join();
}

// 64-bit hash of the `on` columns of every row (IJ:1037 / IJ:1211 call create_hashes here).  A NULL key value
// contributes nothing to its row's hash, as in create_hashes (which skips null slots): rows whose key is NULL in
// every `on` column therefore share the seed's hash and only meet each other — never the rows whose value slot
// happens to hold the same bytes.
template <class V64>  // std::vector<uint64_t> or PinnedVec<uint64_t>
int hash_keys(sq_exec* e, const Side& side, const std::vector<int32_t>& on, const ArrowArray* batch, V64* out) {
  const int64_t n = batch->length;
  out->assign(size_t(n), sqkey::seed());  // on=[(1,1)]: every row carries the constant's hash
  for (int32_t col : on) {
    const ColType& t = side.cols[size_t(col)];
    const ColView v = view_of(batch, col);
    const bool has_nulls = v.null_count != 0 && v.validity != nullptr;
    std::vector<uint64_t> before;
    if (has_nulls) before.assign(out->begin(), out->end());  // null slots get their previous value back below
    struct Restore {
      const ColView& v; const std::vector<uint64_t>& before; V64* out; bool on;
      ~Restore() {
        if (!on) return;
        for (int64_t i = 0; i < v.length; ++i)
          if (!bit_at(v.validity, v.offset + i)) (*out)[size_t(i)] = before[size_t(i)];
      }
    } restore{v, before, out, has_nulls};
    if (t.dict) {
      // dictionary-encoded key: hash every dictionary value once, rows look their value's hash up — the same
      // hash a plain Utf8 column of the same strings gets, so a dictionary side and a Utf8 side of one join agree
      std::vector<StrRef> strs;
      dict_strings(batch->children[col]->dictionary, t.dict_large, &strs);
      std::vector<uint64_t> eh(strs.size());
      for (size_t k = 0; k < strs.size(); ++k) eh[k] = string_hash(strs[k].first, strs[k].second);
      parallel_rows(n, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
          const uint32_t ix = dict_index(v.values, v.offset + i, t.width);
          (*out)[size_t(i)] = sqkey::fold((*out)[size_t(i)], ix < eh.size() ? eh[ix] : 0);
        }
      });
    } else if (t.kind == Kind::Fixed) {
      if (t.width != 1 && t.width != 2 && t.width != 4 && t.width != 8)
        return e->fail(SQ_EINVAL, "key column '%s' has unsupported type '%s'", t.name.c_str(), t.format.c_str());
      parallel_rows(n, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
          uint64_t raw = 0;
          memcpy(&raw, v.values + (v.offset + i) * t.width, t.width);
          (*out)[size_t(i)] = sqkey::fold((*out)[size_t(i)], sqkey::of_fixed(raw));
        }
      });
    } else {
      // strings of up to 8 bytes (contig names) hash from one masked 8-byte load; longer ones byte by byte
      const int64_t data_end = n == 0 ? 0
                               : (t.kind == Kind::Utf8 ? int64_t((reinterpret_cast<const int32_t*>(v.values) + v.offset)[n])
                                                       : (reinterpret_cast<const int64_t*>(v.values) + v.offset)[n]);
      parallel_rows(n, [&](int64_t lo_, int64_t hi_) {
      for (int64_t i = lo_; i < hi_; ++i) {
        int64_t a, b;
        if (t.kind == Kind::Utf8) {
          const int32_t* off = reinterpret_cast<const int32_t*>(v.values) + v.offset;
          a = off[i]; b = off[i + 1];
        } else {
          const int64_t* off = reinterpret_cast<const int64_t*>(v.values) + v.offset;
          a = off[i]; b = off[i + 1];
        }
        const int64_t len = b - a;
        uint64_t h;
        if (len <= 8) {
          uint64_t raw = 0;
          if (a + 8 <= data_end) {
            memcpy(&raw, v.data + a, 8);
            raw &= len == 8 ? ~0ull : ((1ull << (8 * len)) - 1ull);
          } else {
            memcpy(&raw, v.data + a, size_t(len));
          }
          h = sqkey::of_string(raw, 0, uint64_t(len));
        } else {
          h = sqkey::kFnvOffset;  // FNV-1a over the bytes
          for (int64_t k = a; k < b; ++k) { h ^= v.data[k]; h *= sqkey::kFnvPrime; }
        }
        (*out)[size_t(i)] = sqkey::fold((*out)[size_t(i)], h);
      }
      });
    }
  }
  return SQ_OK;
}

// evaluate_as_i32 (IJ:1661-1672): the interval column (or `column - 1`) as Int32, overflow is an error
template <class V32>  // std::vector<int32_t> or PinnedVec<int32_t>
int eval_i32(sq_exec* e, sq_stream* st, const Side& side, int32_t col, bool minus_one, const ArrowArray* batch, V32* out) {
  const ColType& t = side.cols[size_t(col)];
  const ColView v = view_of(batch, col);
  const int64_t n = batch->length;
  out->resize(size_t(n));
  if (v.null_count != 0 && v.validity) {
    // the reference reads `.value(i)` of the cast column without looking at validity (IJ:1039-1048), i.e. whatever
    // bytes sit in a null slot; nothing defined can be reproduced, so a NULL coordinate is an error here
    for (int64_t i = 0; i < n; ++i)
      if (!bit_at(v.validity, v.offset + i))
        return e->fail(SQ_EINVAL, "interval column '%s' holds NULL at row %lld: the interval join needs non-null start / end values",
                       t.name.c_str(), (long long)i);
  }
  if (t.format == "i") {
    const int32_t* p = reinterpret_cast<const int32_t*>(v.values) + v.offset;
    if (!minus_one) { out->assign(p, p + n); return SQ_OK; }
    for (int64_t i = 0; i < n; ++i) {
      if (minus_one && p[i] == INT32_MIN) return e->fail(SQ_ECAST, "Arrow error: Arithmetic overflow: Overflow happened on: %d - 1", p[i]);
      (*out)[size_t(i)] = p[i] - (minus_one ? 1 : 0);
    }
    return SQ_OK;
  }
  if (t.format == "l") {  // BIGINT columns (queries/q1-coitrees.sql:6,11): checked cast on the device
    const int64_t* p = reinterpret_cast<const int64_t*>(v.values) + v.offset;
    int rc = sq_cast_i64_to_i32(st, p, uint64_t(n), minus_one ? 1 : 0, out->data());
    if (rc != SQ_OK) return e->fail(rc, "%s", sq_stream_last_error(st));
    return SQ_OK;
  }
  return e->fail(SQ_EINVAL, "interval column '%s' must be Int32 or Int64, got '%s'", t.name.c_str(), t.format.c_str());
}

int stream_for(sq_exec* e, int32_t partition, sq_stream** out) {
  std::lock_guard<std::mutex> g(e->mu);
  auto it = e->streams.find(partition);
  if (it != e->streams.end()) { *out = it->second; return SQ_OK; }
  sq_stream* s = nullptr;
  int rc = sq_stream_create(e->ctx, &s);
  if (rc != SQ_OK) { e->err = sq_last_error(e->ctx); return rc; }
  e->streams[partition] = s;
  *out = s;
  return SQ_OK;
}

// [left columns..., right columns...] index -> (side, column)
inline void split_col(const sq_exec* e, int32_t pc, int* side, int32_t* col) {
  const int32_t nl = int32_t(e->left.cols.size());
  *side = pc < nl ? 0 : 1;
  *col = pc < nl ? pc : pc - nl;
}

ArrowSchema* make_field(const ColType& t) {
  auto* f = new ArrowSchema();
  auto* o = new Owned();
  o->s1 = t.format;
  o->s2 = t.name;
  f->format = o->s1.c_str();
  f->name = o->s2.c_str();
  f->metadata = nullptr;
  f->flags = t.flags;
  f->n_children = 0;
  f->children = nullptr;
  f->dictionary = nullptr;
  f->release = release_schema;
  f->private_data = o;
  if (t.dict) {  // the field of the dictionary values
    auto* d = new ArrowSchema();
    auto* od = new Owned();
    od->s1 = t.dict_large ? "U" : "u";
    od->s2 = "";
    d->format = od->s1.c_str();
    d->name = od->s2.c_str();
    d->metadata = nullptr;
    d->flags = 0;
    d->n_children = 0;
    d->children = nullptr;
    d->dictionary = nullptr;
    d->release = release_schema;
    d->private_data = od;
    f->dictionary = d;
  }
  return f;
}

// a Utf8 / LargeUtf8 array owning copies of `strs`: the dictionary of a dictionary-encoded output column
ArrowArray* make_dictionary(const std::vector<StrRef>& strs, bool large) {
  auto* a = new ArrowArray();
  auto* o = new Owned();
  uint64_t total = 0;
  for (const StrRef& s : strs) total += uint64_t(s.second);
  const size_t ow = large ? 8 : 4;
  void* off = malloc((strs.size() + 1) * ow);
  auto* data = static_cast<uint8_t*>(malloc(total ? total : 1));
  o->heap = {off, data};
  uint64_t at = 0;
  for (size_t k = 0; k <= strs.size(); ++k) {
    if (large) static_cast<int64_t*>(off)[k] = int64_t(at);
    else static_cast<int32_t*>(off)[k] = int32_t(at);
    if (k < strs.size()) {
      if (strs[k].second) memcpy(data + at, strs[k].first, size_t(strs[k].second));
      at += uint64_t(strs[k].second);
    }
  }
  o->buffers = {nullptr, off, data};
  a->length = int64_t(strs.size());
  a->null_count = 0;
  a->offset = 0;
  a->n_buffers = 3;
  a->buffers = o->buffers.data();
  a->n_children = 0;
  a->children = nullptr;
  a->dictionary = nullptr;
  a->release = release_array;
  a->private_data = o;
  return a;
}

}  // namespace

SQ_API int32_t sq_exec_create(const sq_exec_config* cfg, const ArrowSchema* left_schema, const ArrowSchema* right_schema,
                              sq_exec** out) {
  if (!cfg || !out) return SQ_EINVAL;
  *out = nullptr;
  auto e = std::make_unique<sq_exec>();
  e->cfg = *cfg;
  int rc;
  if ((rc = read_side(e.get(), left_schema, &e->left, "left")) || (rc = read_side(e.get(), right_schema, &e->right, "right"))) {
    *out = e.release();  // caller reads the message, then frees
    return rc;
  }
  e->on_left.assign(cfg->on_left, cfg->on_left + (cfg->n_on > 0 ? cfg->n_on : 0));
  e->on_right.assign(cfg->on_right, cfg->on_right + (cfg->n_on > 0 ? cfg->n_on : 0));
  const int32_t nl = int32_t(e->left.cols.size()), nr = int32_t(e->right.cols.size());
  if (cfg->n_projection < 0) {
    for (int32_t i = 0; i < nl + nr; ++i) e->projection.push_back(i);
  } else {
    e->projection.assign(cfg->projection, cfg->projection + cfg->n_projection);
  }
  *out = e.get();
  auto bad = [&](const char* what, int32_t v, int32_t lim) {
    return v < 0 || v >= lim ? e->fail(SQ_EINVAL, "%s index %d out of range [0,%d)", what, v, lim) : SQ_OK;
  };
  for (int32_t c : e->on_left) if ((rc = bad("on_left", c, nl))) { e.release(); return rc; }
  for (int32_t c : e->on_right) if ((rc = bad("on_right", c, nr))) { e.release(); return rc; }
  for (int32_t c : e->projection) if ((rc = bad("projection", c, nl + nr))) { e.release(); return rc; }
  if ((rc = bad("left_start", cfg->left_start, nl)) || (rc = bad("left_end", cfg->left_end, nl)) ||
      (rc = bad("right_start", cfg->right_start, nr)) || (rc = bad("right_end", cfg->right_end, nr))) {
    e.release();
    return rc;
  }
  for (int32_t pc : e->projection) {
    int side; int32_t col;
    split_col(e.get(), pc, &side, &col);
    const ColType& t = (side ? e->right : e->left).cols[size_t(col)];
    if (t.kind == Kind::Fixed && t.width != 4 && t.width != 8 && t.width != 16 && t.width != 1 && t.width != 2) {
      rc = e->fail(SQ_EINVAL, "output column '%s' has unsupported type '%s'", t.name.c_str(), t.format.c_str());
      e.release();
      return rc;
    }
  }
  rc = sq_ctx_create(cfg->device, &e->ctx);
  if (rc != SQ_OK) {  // no usable GPU: an error at execute, never a CPU fallback
    e->err = sq_last_error(nullptr);
    e.release();
    return rc;
  }
  // the build table is the node's own concatenation of the build batches: its payload is kept in the index's sorted order
  // and the pairs carry positions, so every `take` of a build column reads neighbouring rows (a SET of
  // sequila.cuda_build_ids before the build overrides this)
  sq_ctx_set_option(e->ctx, "cuda_build_ids", "positions");
  e.release();
  return SQ_OK;
}

SQ_API int32_t sq_exec_push_build(sq_exec* e, ArrowArray* batch) {
  if (!e || !batch || !batch->release) return SQ_EINVAL;
  if (e->built) return e->fail(SQ_ESTATE, "build side already finished");
  if (batch->n_children != int64_t(e->left.cols.size())) return e->fail(SQ_EINVAL, "build batch has %lld columns, schema has %zu",
                                                                         (long long)batch->n_children, e->left.cols.size());
  e->build_batches.push_back(*batch);  // move (Arrow C data interface: shallow copy + mark source released)
  batch->release = nullptr;
  e->m[0] += 1;
  e->m[1] += uint64_t(batch->length);
  return SQ_OK;
}

SQ_API int32_t sq_exec_finish_build(sq_exec* e) {
  if (!e) return SQ_EINVAL;
  if (e->built) return e->fail(SQ_ESTATE, "build side already finished");
  const auto t0 = Clock::now();
  uint64_t n = 0;
  for (auto& b : e->build_batches) n += uint64_t(b.length);
  std::vector<uint64_t> keys;
  std::vector<int32_t> start, end;
  keys.reserve(n); start.reserve(n); end.reserve(n);
  sq_stream* st = nullptr;
  int rc = stream_for(e, -1, &st);
  if (rc) return rc;
  std::vector<uint64_t> hk;
  std::vector<int32_t> s32, e32;
  for (auto& b : e->build_batches) {  // update_hashmap per batch, rows numbered across batches (IJ:667-681)
    if ((rc = hash_keys(e, e->left, e->on_left, &b, &hk))) return rc;
    if ((rc = eval_i32(e, st, e->left, e->cfg.left_start, false, &b, &s32))) return rc;
    if ((rc = eval_i32(e, st, e->left, e->cfg.left_end, e->cfg.left_end_minus_one != 0, &b, &e32))) return rc;
    keys.insert(keys.end(), hk.begin(), hk.end());
    start.insert(start.end(), s32.begin(), s32.end());
    end.insert(end.end(), e32.begin(), e32.end());
  }
  rc = sq_index_build(e->ctx, keys.data(), start.data(), end.data(), n, &e->index);
  if (rc != SQ_OK) return e->fail(rc, "%s", sq_last_error(e->ctx));

  // build-side payload columns that the projection needs become device-resident (concat_batches, IJ:685)
  e->build_col_id.assign(e->left.cols.size(), -1);
  for (int32_t pc : e->projection) {
    int side; int32_t col;
    split_col(e, pc, &side, &col);
    if (side != 0 || e->build_col_id[size_t(col)] >= 0) continue;
    const ColType& t = e->left.cols[size_t(col)];
    int32_t id = -1;
    bool any_null = false;
    std::vector<uint8_t> validity((n + 7) / 8, 0xFF);
    if (t.dict) {
      // the build batches may each bring their own dictionary: unify the values (order of first occurrence over
      // the batches) and renumber the indices, which then live on the device as a 4-byte payload column
      std::vector<std::string>& gdict = e->build_dicts[col];
      std::map<std::string, uint32_t> gmap;
      std::vector<uint32_t> buf(size_t(n), 0);
      uint64_t r = 0;
      for (auto& b : e->build_batches) {
        const ColView v = view_of(&b, col);
        std::vector<StrRef> strs;
        dict_strings(b.children[col]->dictionary, t.dict_large, &strs);
        std::vector<uint32_t> remap(strs.size());
        for (size_t k = 0; k < strs.size(); ++k) {
          std::string sv(reinterpret_cast<const char*>(strs[k].first), size_t(strs[k].second));
          auto it = gmap.find(sv);
          if (it == gmap.end()) { it = gmap.emplace(sv, uint32_t(gdict.size())).first; gdict.push_back(sv); }
          remap[k] = it->second;
        }
        for (int64_t i = 0; i < v.length; ++i, ++r) {
          const bool null = v.null_count != 0 && v.validity && !bit_at(v.validity, v.offset + i);
          const uint32_t ix = dict_index(v.values, v.offset + i, t.width);
          buf[r] = (!null && ix < remap.size()) ? remap[ix] : 0u;
          if (null) { validity[r >> 3] &= uint8_t(~(1u << (r & 7))); any_null = true; }
        }
      }
      if (!gdict.empty() && gdict.size() - 1 > dict_index_limit(t.format))
        return e->fail(SQ_EINVAL, "column '%s': %zu distinct values over the build batches do not fit index type '%s'",
                       t.name.c_str(), gdict.size(), t.format.c_str());
      rc = sq_index_add_column(e->index, buf.data(), 4, &id);
    } else if (t.kind == Kind::Fixed) {
      // sub-4-byte values are widened into 4-byte slots for the gather and narrowed again on output
      const uint32_t w = t.width < 4 ? 4 : t.width;
      std::vector<uint8_t> buf(size_t(n) * w, 0);
      uint64_t r = 0;
      for (auto& b : e->build_batches) {
        const ColView v = view_of(&b, col);
        const bool nulls = v.null_count != 0 && v.validity;
        if (w == t.width) {  // one copy per batch; only batches that hold NULLs are walked row by row
          if (v.length) memcpy(buf.data() + r * w, v.values + v.offset * t.width, size_t(v.length) * w);
          if (nulls)
            for (int64_t i = 0; i < v.length; ++i)
              if (!bit_at(v.validity, v.offset + i)) { validity[(r + i) >> 3] &= uint8_t(~(1u << ((r + i) & 7))); any_null = true; }
          r += uint64_t(v.length);
          continue;
        }
        for (int64_t i = 0; i < v.length; ++i, ++r) {
          memcpy(buf.data() + r * w, v.values + (v.offset + i) * t.width, t.width);
          if (nulls && !bit_at(v.validity, v.offset + i)) { validity[r >> 3] &= uint8_t(~(1u << (r & 7))); any_null = true; }
        }
      }
      rc = sq_index_add_column(e->index, buf.data(), w, &id);
    } else {
      std::vector<int64_t> off(size_t(n) + 1, 0);
      std::vector<uint8_t> data;
      uint64_t r = 0;
      for (auto& b : e->build_batches) {  // per batch: one copy of the bytes, the offsets shifted onto the running total
        const ColView v = view_of(&b, col);
        if (!v.length) continue;
        const bool nulls = v.null_count != 0 && v.validity;
        const int64_t base = int64_t(data.size());
        int64_t first, last;
        if (t.kind == Kind::Utf8) {
          const int32_t* o = reinterpret_cast<const int32_t*>(v.values) + v.offset;
          first = o[0]; last = o[v.length];
          parallel_rows(v.length, [&](int64_t lo, int64_t hi) { for (int64_t i = lo; i < hi; ++i) off[r + 1 + uint64_t(i)] = base + (int64_t(o[i + 1]) - first); });
        } else {
          const int64_t* o = reinterpret_cast<const int64_t*>(v.values) + v.offset;
          first = o[0]; last = o[v.length];
          parallel_rows(v.length, [&](int64_t lo, int64_t hi) { for (int64_t i = lo; i < hi; ++i) off[r + 1 + uint64_t(i)] = base + (o[i + 1] - first); });
        }
        data.insert(data.end(), v.data + first, v.data + last);
        if (nulls)
          for (int64_t i = 0; i < v.length; ++i)
            if (!bit_at(v.validity, v.offset + i)) { validity[(r + i) >> 3] &= uint8_t(~(1u << ((r + i) & 7))); any_null = true; }
        r += uint64_t(v.length);
      }
      rc = sq_index_add_utf8_column(e->index, off.data(), data.data(), data.size(), &id);
    }
    if (rc != SQ_OK) return e->fail(rc, "%s", sq_last_error(e->ctx));
    if (any_null && (rc = sq_index_set_validity(e->index, id, validity.data())) != SQ_OK) return e->fail(rc, "%s", sq_last_error(e->ctx));
    e->build_col_id[size_t(col)] = id | (any_null ? 0x40000000 : 0);
  }
  for (auto& b : e->build_batches) if (b.release) b.release(&b);  // everything needed now lives on the device
  e->build_batches.clear();
  e->built = true;
  e->m[2] = sq_index_bytes(e->index);
  e->m[9] = sq_index_bytes(e->index);
  e->m[10] = sq_index_keys(e->index);
  e->m[7] = uint64_t(std::chrono::duration_cast<std::chrono::nanoseconds>(Clock::now() - t0).count());
  return SQ_OK;
}

SQ_API int32_t sq_exec_output_schema(const sq_exec* e, ArrowSchema* out) {
  if (!e || !out) return SQ_EINVAL;
  auto* o = new Owned();
  o->s1 = "+s";
  o->s2 = "";
  for (int32_t pc : e->projection) {
    int side; int32_t col;
    split_col(e, pc, &side, &col);
    ColType t = (side ? e->right : e->left).cols[size_t(col)];
    if (t.kind == Kind::LargeUtf8) { t.kind = Kind::Utf8; }  // format is kept; LargeUtf8 output uses int64 offsets
    o->schema_children.push_back(make_field(t));
  }
  o->schema_child_ptrs = o->schema_children;
  out->format = o->s1.c_str();
  out->name = o->s2.c_str();
  out->metadata = nullptr;
  out->flags = 0;
  out->n_children = int64_t(o->schema_children.size());
  out->children = o->schema_child_ptrs.data();
  out->dictionary = nullptr;
  out->release = release_schema;
  out->private_data = o;
  return SQ_OK;
}

namespace {

// hashes the keys, casts the interval columns and runs the probe of one batch on the GPU; the result stays
// on the device.  *n_out = output rows of the whole batch.
// host half: key hashes and the i32 view of the interval columns of one probe tile (`cast` is the stream a BIGINT column is
// cast on: the partition's own, or its second one when a worker thread is using the first)
template <class V64, class V32>
int prep_tile(sq_exec* e, sq_stream* cast, const ArrowArray* batch, V64* keys, V32* start, V32* end) {
  if (uint64_t(batch->length) > 0xFFFFFFFFull) return e->fail(SQ_EINVAL, "probe batch too large");
  int rc;
  Trace tr(e->ctx);
  if ((rc = hash_keys(e, e->right, e->on_right, batch, keys))) return rc;
  tr.lap("hash_keys");
  if ((rc = eval_i32(e, cast, e->right, e->cfg.right_start, false, batch, start))) return rc;
  if ((rc = eval_i32(e, cast, e->right, e->cfg.right_end, e->cfg.right_end_minus_one != 0, batch, end))) return rc;
  tr.lap("eval_i32 x2");
  return SQ_OK;
}

// device half: the probe of one prepared tile; the result stays on the device.  *n_out = output rows of the whole tile.
template <class V64, class V32>
int probe_tile(sq_exec* e, sq_stream* st, const V64& keys, const V32& start, const V32& end, uint64_t* n_out) {
  const uint64_t n = keys.size();
  int rc;
  Trace tr(e->ctx);
  uint64_t n_pairs = 0;
  if (e->cfg.algorithm == SQ_EXEC_NEAREST) {  // one output row per probe row; the left side may be NULL (IJ:1593-1602)
    rc = sq_probe_nearest(st, e->index, keys.data(), start.data(), end.data(), uint32_t(n), nullptr);
    if (rc != SQ_OK) return e->fail(rc, "%s", sq_stream_last_error(st));
    n_pairs = n;
  } else {
    rc = sq_probe_count(st, e->index, keys.data(), start.data(), end.data(), uint32_t(n), &n_pairs);
    if (rc != SQ_OK) return e->fail(rc, "%s", sq_stream_last_error(st));
    // the index pairs stay on the device: only the gathered columns travel back (IJ:1620-1632); with an
    // empty projection (`SELECT count(*) ...`, what the reference's benchmarks run) only the row count of
    // the output batch is needed and the pairs are never written
    if (!e->projection.empty()) {
      rc = sq_probe_emit_pairs(st, nullptr, nullptr, nullptr, n_pairs);
      if (rc != SQ_OK) return e->fail(rc, "%s", sq_stream_last_error(st));
    }
  }
  tr.lap("probe on device");
  *n_out = n_pairs;
  return SQ_OK;
}

// hashes the keys, casts the interval columns and runs the probe of one batch on the GPU
int probe_on_device(sq_exec* e, sq_stream* st, const ArrowArray* batch, uint64_t* n_out) {
  std::vector<uint64_t> keys;
  std::vector<int32_t> start, end;
  int rc = prep_tile(e, st, batch, &keys, &start, &end);
  if (rc) return rc;
  return probe_tile(e, st, keys, start, end, n_out);
}

// one output RecordBatch = `take` of every projected column over the stream's current pair window
int assemble_output(sq_exec* e, sq_stream* st, const ArrowArray* batch, uint64_t n_pairs, ArrowArray* out) {
  const uint64_t n = uint64_t(batch->length);
  const bool nearest = e->cfg.algorithm == SQ_EXEC_NEAREST;
  int rc;
  Trace tr(e->ctx);
  // `out` is assembled in place; on any failure below everything attached so far is released
  auto* own = new Owned();
  own->ctx = e->ctx;
  out->release = release_array;
  out->private_data = own;
  struct Unwind {
    ArrowArray* a; bool armed;
    ~Unwind() { if (armed) release_array(a); }
  } unwind{out, true};
  auto pinned = [&](size_t bytes) -> void* {
    void* p = nullptr;
    if (sq_host_alloc(e->ctx, bytes ? bytes : 8, &p) != SQ_OK) return nullptr;
    return p;
  };
  for (int32_t pc : e->projection) {
    int side; int32_t col;
    split_col(e, pc, &side, &col);
    const ColType& t = (side ? e->right : e->left).cols[size_t(col)];
    auto* child = new ArrowArray();
    auto* co = new Owned();
    co->ctx = e->ctx;
    child->length = int64_t(n_pairs);
    child->offset = 0;
    child->n_children = 0;
    child->children = nullptr;
    child->dictionary = nullptr;
    child->release = release_array;
    child->private_data = co;
    child->null_count = 0;
    own->children.push_back(child);
    const void* validity_out = nullptr;

    // validity of the source column, if it has nulls
    const int32_t bid = side == 0 ? (e->build_col_id[size_t(col)] & 0x3FFFFFFF) : -1;
    const bool build_nulls = side == 0 && (e->build_col_id[size_t(col)] & 0x40000000);
    ColView pv;
    std::vector<uint8_t> probe_bm;
    bool probe_nulls = false;
    if (side == 1) {
      pv = view_of(batch, col);
      if (pv.null_count != 0 && pv.validity) {
        probe_nulls = true;
        probe_bm.assign((n + 7) / 8, 0);
        for (uint64_t i = 0; i < n; ++i) if (bit_at(pv.validity, pv.offset + int64_t(i))) probe_bm[i >> 3] |= uint8_t(1u << (i & 7));
      }
    }
    if ((build_nulls || probe_nulls || (nearest && side == 0)) && n_pairs) {
      void* bm = pinned((n_pairs + 7) / 8);
      if (!bm) return e->fail(SQ_ENOMEM, "pinned allocation failed");
      co->pinned.push_back(bm);
      uint64_t nulls = 0;
      rc = sq_gather_validity(st, side, bid, probe_nulls ? probe_bm.data() : nullptr, static_cast<uint8_t*>(bm), &nulls);
      if (rc != SQ_OK) return e->fail(rc, "%s", sq_stream_last_error(st));
      child->null_count = int64_t(nulls);
      validity_out = nulls ? bm : nullptr;  // Arrow: no bitmap needed when nothing is null
    }

    if (t.kind == Kind::Fixed) {
      const uint32_t w = t.width < 4 ? 4 : t.width;
      void* vals = pinned(size_t(n_pairs) * w);
      if (!vals) return e->fail(SQ_ENOMEM, "pinned allocation failed");
      co->pinned.push_back(vals);
      tr.lap("  alloc");
      if (n_pairs) {
        if (side == 0) {
          rc = sq_gather_column(st, 0, bid, nullptr, w, vals, n_pairs);
        } else {
          std::vector<uint8_t> wide;
          const uint8_t* src = pv.values + pv.offset * t.width;
          if (t.width < 4) {
            wide.assign(size_t(n) * 4, 0);
            for (uint64_t i = 0; i < n; ++i) memcpy(wide.data() + i * 4, src + i * t.width, t.width);
            src = wide.data();
          }
          rc = sq_gather_column(st, 1, -1, src, w, vals, n_pairs);
        }
        if (rc != SQ_OK) return e->fail(rc, "%s", sq_stream_last_error(st));
        if (t.width < 4) {  // narrow back in place
          auto* p = static_cast<uint8_t*>(vals);
          for (uint64_t k = 0; k < n_pairs; ++k) memmove(p + k * t.width, p + k * 4, t.width);
        }
      }
      co->buffers = {validity_out, vals};
      if (t.dict) {  // the values travel once per batch: the build side's unified dictionary, or this batch's own
        std::vector<StrRef> strs;
        if (side == 0) {
          auto it = e->build_dicts.find(col);
          if (it != e->build_dicts.end())
            for (const std::string& sv : it->second) strs.emplace_back(reinterpret_cast<const uint8_t*>(sv.data()), int64_t(sv.size()));
        } else {
          dict_strings(batch->children[col]->dictionary, t.dict_large, &strs);
        }
        child->dictionary = make_dictionary(strs, t.dict_large);
      }
    } else {
      // Utf8 / LargeUtf8: offsets first (gives the byte total), then the bytes
      PinnedVec<int64_t> poff{PinnedAlloc<int64_t>(e->ctx)};
      const uint8_t* pdata = nullptr;
      uint64_t pbytes = 0;
      if (side == 1) {
        poff.resize(size_t(n) + 1);
        if (t.kind == Kind::Utf8) {
          const int32_t* o = reinterpret_cast<const int32_t*>(pv.values) + pv.offset;
          parallel_rows(int64_t(n) + 1, [&](int64_t lo, int64_t hi) { for (int64_t i = lo; i < hi; ++i) poff[size_t(i)] = o[i] - o[0]; });
          pdata = pv.data + o[0];
          pbytes = uint64_t(o[n] - o[0]);
        } else {
          const int64_t* o = reinterpret_cast<const int64_t*>(pv.values) + pv.offset;
          parallel_rows(int64_t(n) + 1, [&](int64_t lo, int64_t hi) { for (int64_t i = lo; i < hi; ++i) poff[size_t(i)] = o[i] - o[0]; });
          pdata = pv.data + o[0];
          pbytes = uint64_t(o[n] - o[0]);
        }
      }
      void* off32 = pinned((size_t(n_pairs) + 1) * 4);
      if (!off32) return e->fail(SQ_ENOMEM, "pinned allocation failed");
      co->pinned.push_back(off32);
      static_cast<int32_t*>(off32)[0] = 0;
      uint64_t total = 0;
      if (n_pairs) {
        rc = sq_gather_utf8(st, side, bid, side ? poff.data() : nullptr, pdata, pbytes, static_cast<int32_t*>(off32), &total);
        if (rc != SQ_OK) return e->fail(rc, "%s", sq_stream_last_error(st));
      }
      void* bytes = pinned(total);
      if (!bytes) return e->fail(SQ_ENOMEM, "pinned allocation failed");
      co->pinned.push_back(bytes);
      if (n_pairs && (rc = sq_gather_utf8_data(st, static_cast<uint8_t*>(bytes), total)) != SQ_OK)
        return e->fail(rc, "%s", sq_stream_last_error(st));
      if (t.kind == Kind::LargeUtf8) {  // widen the offsets for a LargeUtf8 output column
        void* off64 = pinned((size_t(n_pairs) + 1) * 8);
        if (!off64) return e->fail(SQ_ENOMEM, "pinned allocation failed");
        co->pinned.push_back(off64);
        for (uint64_t k = 0; k <= n_pairs; ++k) static_cast<int64_t*>(off64)[k] = static_cast<int32_t*>(off32)[k];
        co->buffers = {validity_out, off64, bytes};
      } else {
        co->buffers = {validity_out, off32, bytes};
      }
    }
    child->n_buffers = int64_t(co->buffers.size());
    child->buffers = co->buffers.data();
    tr.lap(t.kind == Kind::Fixed ? "column (fixed)" : "column (utf8)");
  }
  own->child_ptrs = own->children;
  own->buffers = {nullptr};
  out->length = int64_t(n_pairs);
  out->null_count = 0;
  out->offset = 0;
  out->n_buffers = 1;
  out->buffers = own->buffers.data();
  out->n_children = int64_t(own->children.size());
  out->children = own->child_ptrs.data();
  out->dictionary = nullptr;
  unwind.armed = false;
  return SQ_OK;
}

}  // namespace

namespace {
// the per-batch entry points use the partition's stream themselves: not while a pipelined tile's worker thread does
int no_tile_in_flight(sq_exec* e, int32_t partition) {
  std::lock_guard<std::mutex> g(e->mu);
  if (e->jobs.find(partition) == e->jobs.end()) return SQ_OK;
  e->err = "a pushed tile of this partition is still being joined: collect it with sq_exec_probe_pop(flush) first";
  return SQ_ESTATE;
}
}  // namespace

SQ_API int32_t sq_exec_probe(sq_exec* e, int32_t partition, const ArrowArray* batch, ArrowArray* out) {
  if (!e || !batch || !out) return SQ_EINVAL;
  if (!e->built) return e->fail(SQ_ESTATE, "Expected build side in ready state");  // IJ:1425
  if (batch->n_children != int64_t(e->right.cols.size())) return e->fail(SQ_EINVAL, "probe batch has %lld columns, schema has %zu",
                                                                          (long long)batch->n_children, e->right.cols.size());
  const auto t0 = Clock::now();
  sq_stream* st = nullptr;
  int rc = no_tile_in_flight(e, partition);
  if (rc) return rc;
  if ((rc = stream_for(e, partition, &st))) return rc;
  uint64_t n_pairs = 0;
  if ((rc = probe_on_device(e, st, batch, &n_pairs))) return rc;
  if ((rc = assemble_output(e, st, batch, n_pairs, out))) return rc;
  std::lock_guard<std::mutex> g(e->mu);
  e->m[3] += 1;
  e->m[4] += uint64_t(batch->length);
  e->m[5] += 1;
  e->m[6] += n_pairs;
  e->m[8] += uint64_t(std::chrono::duration_cast<std::chrono::nanoseconds>(Clock::now() - t0).count());
  return SQ_OK;
}

// Low-memory mode (IJ:1433-1530): the probe batch is processed once on the GPU, then handed back as output
// batches of at most cfg.max_output_rows rows, cut at probe-row boundaries; a single probe row with more
// matches than the cap forms a batch of its own (IJ:1467-1471: `&& total_output_rows > 0`).
SQ_API int32_t sq_exec_probe_begin(sq_exec* e, int32_t partition, const ArrowArray* batch) {
  if (!e || !batch) return SQ_EINVAL;
  if (!e->built) return e->fail(SQ_ESTATE, "Expected build side in ready state");
  if (batch->n_children != int64_t(e->right.cols.size())) return e->fail(SQ_EINVAL, "probe batch has %lld columns, schema has %zu",
                                                                          (long long)batch->n_children, e->right.cols.size());
  const auto t0 = Clock::now();
  sq_stream* st = nullptr;
  int rc = no_tile_in_flight(e, partition);
  if (rc) return rc;
  if ((rc = stream_for(e, partition, &st))) return rc;
  uint64_t n_pairs = 0;
  if ((rc = probe_on_device(e, st, batch, &n_pairs))) return rc;
  PartState ps;
  ps.batch = batch;
  ps.n = uint64_t(batch->length);
  const uint64_t cap = e->cfg.max_output_rows > 0 ? uint64_t(e->cfg.max_output_rows) : 1000000ull;  // IJ:1439
  if (e->cfg.algorithm == SQ_EXEC_NEAREST) {
    for (uint64_t o = 0; o < n_pairs; o += cap) ps.windows.emplace_back(o, std::min(cap, n_pairs - o));
  } else {
    std::vector<uint32_t> counts(ps.n);
    if (ps.n && (rc = sq_stream_counts(st, counts.data())) != SQ_OK) return e->fail(rc, "%s", sq_stream_last_error(st));
    uint64_t off = 0, cur = 0;
    for (uint64_t i = 0; i < ps.n; ++i) {
      if (cur + counts[i] > cap && cur > 0) { ps.windows.emplace_back(off, cur); off += cur; cur = 0; }
      cur += counts[i];
    }
    ps.windows.emplace_back(off, cur);  // the last (possibly empty) batch, as in the reference
  }
  if (ps.windows.empty()) ps.windows.emplace_back(0, 0);
  std::lock_guard<std::mutex> g(e->mu);
  e->parts[partition] = std::move(ps);
  e->m[3] += 1;
  e->m[4] += uint64_t(batch->length);
  e->m[8] += uint64_t(std::chrono::duration_cast<std::chrono::nanoseconds>(Clock::now() - t0).count());
  return SQ_OK;
}

SQ_API int32_t sq_exec_probe_next(sq_exec* e, int32_t partition, ArrowArray* out, int32_t* has_more_out) {
  if (!e || !out || !has_more_out) return SQ_EINVAL;
  const auto t0 = Clock::now();
  PartState* ps = nullptr;
  {
    std::lock_guard<std::mutex> g(e->mu);
    auto it = e->parts.find(partition);
    if (it == e->parts.end() || it->second.next >= it->second.windows.size()) {
      e->err = "sq_exec_probe_next without a pending sq_exec_probe_begin";
      return SQ_ESTATE;
    }
    ps = &it->second;
  }
  sq_stream* st = nullptr;
  int rc = stream_for(e, partition, &st);
  if (rc) return rc;
  const auto w = ps->windows[ps->next];
  if (!e->projection.empty() && (rc = sq_stream_set_window(st, w.first, w.second)) != SQ_OK)
    return e->fail(rc, "%s", sq_stream_last_error(st));
  if ((rc = assemble_output(e, st, ps->batch, w.second, out))) return rc;
  ps->next += 1;
  *has_more_out = ps->next < ps->windows.size() ? 1 : 0;
  std::lock_guard<std::mutex> g(e->mu);
  e->m[5] += 1;
  e->m[6] += w.second;
  e->m[8] += uint64_t(std::chrono::duration_cast<std::chrono::nanoseconds>(Clock::now() - t0).count());
  return SQ_OK;
}

// ---- probe-batch coalescing ---------------------------------------------------------------------------------
namespace {

// One struct array holding the rows of `batches` back to back (what concat_batches does for the build side, IJ:685):
// fixed-width values and Utf8 offsets / bytes are copied, validity bitmaps are rebuilt, dictionary columns get ONE
// dictionary (values in order of first occurrence over the batches) with renumbered indices.
// buffer of a concatenated tile: pinned, so that the columns the probe side contributes to the output reach the device at
// link speed (released with the tile through Owned::pinned); nullptr when the pool cannot serve it
void* tile_buffer(sq_exec* e, Owned* co, size_t bytes) {
  void* p = nullptr;
  if (sq_host_alloc(e->ctx, bytes ? bytes : 8, &p) != SQ_OK) return nullptr;
  co->pinned.push_back(p);
  return p;
}

int concat_batches(sq_exec* e, const Side& side, const std::vector<ArrowArray>& batches, ArrowArray* out) {
  uint64_t n = 0;
  for (const ArrowArray& b : batches) n += uint64_t(b.length);
  auto* own = new Owned();
  own->ctx = e->ctx;
  out->release = release_array;
  out->private_data = own;
  out->dictionary = nullptr;
  struct Unwind {
    ArrowArray* a; bool armed;
    ~Unwind() { if (armed) release_array(a); }
  } unwind{out, true};
  for (size_t c = 0; c < side.cols.size(); ++c) {
    const ColType& t = side.cols[c];
    auto* child = new ArrowArray();
    auto* co = new Owned();
    co->ctx = e->ctx;
    child->length = int64_t(n);
    child->offset = 0;
    child->null_count = 0;
    child->n_children = 0;
    child->children = nullptr;
    child->dictionary = nullptr;
    child->release = release_array;
    child->private_data = co;
    own->children.push_back(child);
    if (t.width == 0 && t.kind == Kind::Fixed) {  // a type this node never touches: an all-null placeholder of the right length
      co->buffers = {nullptr, nullptr};
      child->null_count = int64_t(n);
      child->n_buffers = 2;
      child->buffers = co->buffers.data();
      continue;
    }
    uint8_t* validity = nullptr;
    uint64_t nulls = 0;
    auto note_nulls = [&](const ColView& v, uint64_t r0) {
      if (v.null_count == 0 || !v.validity) return;
      if (!validity) {
        validity = static_cast<uint8_t*>(malloc((n + 7) / 8 + 1));
        memset(validity, 0xFF, (n + 7) / 8 + 1);
        co->heap.push_back(validity);
      }
      for (int64_t i = 0; i < v.length; ++i)
        if (!bit_at(v.validity, v.offset + i)) { validity[(r0 + uint64_t(i)) >> 3] &= uint8_t(~(1u << ((r0 + uint64_t(i)) & 7))); ++nulls; }
    };
    if (t.kind == Kind::Fixed) {
      auto* vals = static_cast<uint8_t*>(tile_buffer(e, co, size_t(n) * t.width + 8));
      if (!vals) return e->fail(SQ_ECUDA, "%s", sq_last_error(e->ctx));
      uint64_t r = 0;
      if (t.dict) {
        std::vector<std::string> gdict;
        std::map<std::string, uint32_t> gmap;
        for (const ArrowArray& b : batches) {
          const ColView v = view_of(&b, int32_t(c));
          std::vector<StrRef> strs;
          dict_strings(b.children[c]->dictionary, t.dict_large, &strs);
          std::vector<uint32_t> remap(strs.size());
          for (size_t k = 0; k < strs.size(); ++k) {
            std::string sv(reinterpret_cast<const char*>(strs[k].first), size_t(strs[k].second));
            auto it = gmap.find(sv);
            if (it == gmap.end()) { it = gmap.emplace(sv, uint32_t(gdict.size())).first; gdict.push_back(sv); }
            remap[k] = it->second;
          }
          note_nulls(v, r);
          for (int64_t i = 0; i < v.length; ++i, ++r) {
            const uint32_t ix = dict_index(v.values, v.offset + i, t.width);
            const uint32_t nx = ix < remap.size() ? remap[ix] : 0u;
            memcpy(vals + r * t.width, &nx, t.width);
          }
        }
        if (!gdict.empty() && gdict.size() - 1 > dict_index_limit(t.format))
          return e->fail(SQ_EINVAL, "column '%s': %zu distinct values over the coalesced probe batches do not fit index type '%s'",
                         t.name.c_str(), gdict.size(), t.format.c_str());
        std::vector<StrRef> refs;
        for (const std::string& sv : gdict) refs.emplace_back(reinterpret_cast<const uint8_t*>(sv.data()), int64_t(sv.size()));
        child->dictionary = make_dictionary(refs, t.dict_large);
      } else {
        for (const ArrowArray& b : batches) {
          const ColView v = view_of(&b, int32_t(c));
          note_nulls(v, r);
          if (v.length) memcpy(vals + r * t.width, v.values + v.offset * t.width, size_t(v.length) * t.width);
          r += uint64_t(v.length);
        }
      }
      co->buffers = {validity, vals};
    } else {
      const bool large = t.kind == Kind::LargeUtf8;
      uint64_t total = 0;
      for (const ArrowArray& b : batches) {
        const ColView v = view_of(&b, int32_t(c));
        if (!v.length) continue;
        if (large) { const int64_t* o = reinterpret_cast<const int64_t*>(v.values) + v.offset; total += uint64_t(o[v.length] - o[0]); }
        else { const int32_t* o = reinterpret_cast<const int32_t*>(v.values) + v.offset; total += uint64_t(o[v.length] - o[0]); }
      }
      if (!large && total > 0x7FFFFFFFull)
        return e->fail(SQ_ECAPACITY, "column '%s': the coalesced probe batches hold %llu string bytes; Utf8 offsets are 32-bit",
                       t.name.c_str(), (unsigned long long)total);
      void* off = tile_buffer(e, co, (size_t(n) + 1) * (large ? 8 : 4));
      auto* data = static_cast<uint8_t*>(tile_buffer(e, co, total));
      if (!off || !data) return e->fail(SQ_ECUDA, "%s", sq_last_error(e->ctx));
      uint64_t r = 0, at = 0;
      for (const ArrowArray& b : batches) {  // per batch: ONE copy of its byte range, offsets rebased in a plain loop
        const ColView v = view_of(&b, int32_t(c));
        note_nulls(v, r);
        if (!v.length) continue;
        if (large) {
          const int64_t* o = reinterpret_cast<const int64_t*>(v.values) + v.offset;
          const int64_t base = o[0];
          memcpy(data + at, v.data + base, size_t(o[v.length] - base));
          auto* dst = static_cast<int64_t*>(off) + r;
          for (int64_t i = 0; i < v.length; ++i) dst[i] = int64_t(at) + (o[i] - base);
          at += uint64_t(o[v.length] - base);
        } else {
          const int32_t* o = reinterpret_cast<const int32_t*>(v.values) + v.offset;
          const int32_t base = o[0];
          memcpy(data + at, v.data + base, size_t(o[v.length] - base));
          auto* dst = static_cast<int32_t*>(off) + r;
          for (int64_t i = 0; i < v.length; ++i) dst[i] = int32_t(at) + (o[i] - base);
          at += uint64_t(o[v.length] - base);
        }
        r += uint64_t(v.length);
      }
      if (large) static_cast<int64_t*>(off)[n] = int64_t(at); else static_cast<int32_t*>(off)[n] = int32_t(at);
      co->buffers = {validity, off, data};
    }
    child->null_count = int64_t(nulls);
    child->n_buffers = int64_t(co->buffers.size());
    child->buffers = co->buffers.data();
  }
  own->child_ptrs = own->children;
  own->buffers = {nullptr};
  out->length = int64_t(n);
  out->null_count = 0;
  out->offset = 0;
  out->n_buffers = 1;
  out->buffers = own->buffers.data();
  out->n_children = int64_t(own->children.size());
  out->children = own->child_ptrs.data();
  unwind.armed = false;
  return SQ_OK;
}

uint64_t coalesce_target(sq_exec* e) {
  char v[32] = {0};
  if (sq_ctx_get_option(e->ctx, "cuda_coalesce_rows", v, sizeof v) != SQ_OK) return 1u << 20;
  const long long x = atoll(v);
  return x > 0 ? uint64_t(x) : (1u << 20);
}

}  // namespace

// The reference's probe child yields batches of at most `batch_size` rows (8192 by default, IJ:1192-1233) and joins each on
// its own.  One GPU launch chain + one PCIe round trip per 8192 rows is bound by their latencies (measured: 23 M probe
// rows/s), so the node coalesces: pushed batches wait until `cuda_coalesce_rows` rows (default 1M) are there, then leave as
// ONE tile and ONE output batch (the order of rows is that of the probe batches: maintains_input_order holds).
SQ_API int32_t sq_exec_probe_push(sq_exec* e, int32_t partition, ArrowArray* batch, int32_t* ready_out) {
  if (!e || !batch || !ready_out) return SQ_EINVAL;
  if (!e->built) return e->fail(SQ_ESTATE, "Expected build side in ready state");  // IJ:1425
  if (batch->n_children != int64_t(e->right.cols.size())) return e->fail(SQ_EINVAL, "probe batch has %lld columns, schema has %zu",
                                                                          (long long)batch->n_children, e->right.cols.size());
  const uint64_t target = coalesce_target(e);
  std::lock_guard<std::mutex> g(e->mu);
  Pending& p = e->pending[partition];
  p.batches.push_back(*batch);  // move
  p.rows += uint64_t(batch->length);
  batch->release = nullptr;
  e->m[3] += 1;
  e->m[4] += uint64_t(p.batches.back().length);
  *ready_out = p.rows >= target ? 1 : 0;
  return SQ_OK;
}

// Tiles are pipelined two deep per partition.  A call that finds a tile ready preps it on the calling thread while the
// previous tile's worker thread is still on the GPU, then collects the previous tile's output and starts the worker of the
// new one; the output of a tile therefore leaves one call later than the tile went in (at the latest with flush, which
// the caller repeats until nothing comes back).  Row order across output batches is the probe order, as before.
SQ_API int32_t sq_exec_probe_pop(sq_exec* e, int32_t partition, int32_t flush, ArrowArray* out, int32_t* has_out) {
  if (!e || !out || !has_out) return SQ_EINVAL;
  *has_out = 0;
  const uint64_t target = coalesce_target(e);
  std::unique_ptr<TileJob> prev, next;
  {
    std::lock_guard<std::mutex> g(e->mu);
    auto jt = e->jobs.find(partition);
    if (jt != e->jobs.end()) { prev = std::move(jt->second); e->jobs.erase(jt); }
    auto it = e->pending.find(partition);
    if (it != e->pending.end() && !it->second.batches.empty() && (it->second.rows >= target || flush)) {
      next.reset(new TileJob(e->ctx));
      next->take = std::move(it->second);
      e->pending.erase(it);
    }
  }
  if (!prev && !next) return SQ_OK;
  const auto t0 = Clock::now();
  auto put_back = [&](std::unique_ptr<TileJob>& j) {
    std::lock_guard<std::mutex> g(e->mu);
    e->jobs[partition] = std::move(j);
  };
  int rc = SQ_OK;
  if (next) {  // host half of the new tile, beside the worker of the previous one
    sq_stream* cast = nullptr;
    if ((rc = stream_for(e, partition + (1 << 20), &cast)) == SQ_OK) {
      next->tile = &next->take.batches[0];
      if (next->take.batches.size() > 1) {
        Trace tr(e->ctx);
        rc = concat_batches(e, e->right, next->take.batches, &next->joined);
        next->tile = &next->joined;
        tr.lap("concat batches");
      }
      if (rc == SQ_OK) rc = prep_tile(e, cast, next->tile, &next->keys, &next->start, &next->end);
    }
    if (rc) {  // the new tile is dropped with its error; the previous one stays collectable
      if (prev) put_back(prev);
      return rc;
    }
  }
  const uint64_t host_ns = uint64_t(std::chrono::duration_cast<std::chrono::nanoseconds>(Clock::now() - t0).count());
  uint64_t worker_ns = 0, n_pairs = 0;
  bool have = false;
  if (prev) {
    if (prev->th.joinable()) prev->th.join();
    rc = prev->rc;
    worker_ns = prev->worker_ns;
    n_pairs = prev->n_pairs;
    if (rc == SQ_OK) {
      *out = prev->out;  // move
      prev->out.release = nullptr;
      have = true;
    }
    prev.reset();
    if (rc) return rc;  // its message is in e->err; `next` is dropped
  }
  if (next) {
    sq_stream* st = nullptr;
    if ((rc = stream_for(e, partition, &st))) { if (have && out->release) out->release(out); return rc; }
    TileJob* j = next.get();
    auto work = [e, st, j]() {
      const auto w0 = Clock::now();
      j->rc = probe_tile(e, st, j->keys, j->start, j->end, &j->n_pairs);
      if (j->rc == SQ_OK) j->rc = assemble_output(e, st, j->tile, j->n_pairs, &j->out);
      j->worker_ns = uint64_t(std::chrono::duration_cast<std::chrono::nanoseconds>(Clock::now() - w0).count());
    };
    try {
      j->th = std::thread(work);
    } catch (const std::system_error&) {
      work();  // no thread to be had: the tile is joined on the calling thread, its output leaves with the next call
    }
    if (!have && flush) {  // nothing older to hand out: this tile's own output
      if (j->th.joinable()) j->th.join();
      rc = j->rc;
      worker_ns += j->worker_ns;
      n_pairs = j->n_pairs;
      if (rc == SQ_OK) {
        *out = j->out;
        j->out.release = nullptr;
        have = true;
      }
      next.reset();
      if (rc) return rc;
    } else {
      put_back(next);
    }
  }
  std::lock_guard<std::mutex> g(e->mu);
  // join_time = the library's own work: this call's host half plus the collected tile's worker (they overlap in wall time)
  e->m[8] += host_ns + worker_ns;
  if (have) {
    *has_out = 1;
    e->m[5] += 1;
    e->m[6] += n_pairs;
    e->m[11] += 1;  // tiles joined
  }
  return SQ_OK;
}

SQ_API int32_t sq_exec_metrics(const sq_exec* e, uint64_t out[16]) {
  if (!e || !out) return SQ_EINVAL;
  memcpy(out, e->m, sizeof e->m);
  return SQ_OK;
}

SQ_API const char* sq_exec_last_error(const sq_exec* e) { return e ? e->err.c_str() : ""; }

// `SET sequila.cuda_* TO value` reaches the exec node's context here (sq_ctx_set_option)
SQ_API int32_t sq_exec_set_option(sq_exec* e, const char* key, const char* value) {
  if (!e) return SQ_EINVAL;
  const int rc = sq_ctx_set_option(e->ctx, key, value);
  if (rc != SQ_OK) return e->fail(rc, "%s", sq_last_error(e->ctx));
  return SQ_OK;
}

SQ_API void sq_exec_free(sq_exec* e) {
  if (!e) return;
  e->jobs.clear();  // joins the workers, releases what they hold
  for (auto& kv : e->pending)
    for (ArrowArray& b : kv.second.batches) if (b.release) b.release(&b);
  for (auto& kv : e->streams) sq_stream_free(kv.second);
  for (auto& b : e->build_batches) if (b.release) b.release(&b);
  if (e->index) sq_index_free(e->index);
  if (e->ctx) sq_ctx_destroy(e->ctx);
  delete e;
}
