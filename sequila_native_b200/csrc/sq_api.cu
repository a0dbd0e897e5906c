// C ABI of libsequila_cuda.so (declared in include/sequila_cuda.h): context / index / stream
// objects, host<->device staging and the call protocol around the kernels in sq_build.cu,
// sq_probe.cu and sq_gather.cu.  No CPU fallback exists anywhere in this library.
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <strings.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "sq_internal.cuh"

#define SQ_API extern "C" __attribute__((visibility("default")))

namespace sq {

int fail(ErrorSlot& e, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  e.set(buf);
  return code;
}

void release(sq_buf& b) {
  if (!b.p) return;
  if (b.pinned) cudaFreeHost(b.p); else cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

int ensure(ErrorSlot& e, sq_buf& b, size_t bytes, bool pinned) {
  if (bytes < 256) bytes = 256;
  if (b.p && b.cap >= bytes) return SQ_OK;
  release(b);
  const size_t want = bytes + bytes / 4;  // grow-only with slack: tiles of similar size reuse it
  cudaError_t err = pinned ? cudaHostAlloc(&b.p, want, cudaHostAllocDefault) : cudaMalloc(&b.p, want);
  if (err != cudaSuccess) {
    b.p = nullptr;
    cudaGetLastError();
    return fail(e, SQ_ENOMEM, "%s of %zu bytes failed: %s", pinned ? "cudaHostAlloc" : "cudaMalloc", want,
                cudaGetErrorString(err));
  }
  b.cap = want;
  b.pinned = pinned;
  return SQ_OK;
}

static int ensure_events(sq_stream* s) {
  if (s->ev_ready) return SQ_OK;
  for (auto& e : s->ev) SQ_CUDA(s->err, cudaEventCreate(&e));
  s->ev_ready = true;
  return SQ_OK;
}

static inline void mark(sq_stream* s, int k) {
  if (s->profiling && s->ev_ready) cudaEventRecord(s->ev[k], s->stream);
}

// Fold the timings of events that are known complete (the stream was just synchronised) into the
// running sums.  pending bit k <=> phase k has a recorded, not yet read, event pair.
static void fold_phases(sq_stream* s) {
  static const int a[5] = {0, 1, 3, 4, 6}, b[5] = {1, 2, 4, 5, 7};
  for (int k = 0; k < 5; ++k) {
    if (!(s->pending & (1u << k))) continue;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s->ev[a[k]], s->ev[b[k]]) == cudaSuccess) {
      s->phase_ms[k] = ms;
      s->phase_sum[k] += ms;
      s->phase_n[k] += 1;
    } else {
      cudaGetLastError();
    }
  }
  s->pending = 0;
}

}  // namespace sq

using namespace sq;

SQ_API int32_t sq_abi_version(void) { return SQ_ABI_VERSION; }

SQ_API int32_t sq_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

static ErrorSlot g_create_err;

SQ_API int32_t sq_ctx_create(int32_t device, sq_ctx** out) {
  if (!out) return SQ_EINVAL;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(g_create_err, SQ_ECUDA, "no usable CUDA device (%s); the cuda interval join has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "0 devices");
  }
  if (device < 0 || device >= n) return fail(g_create_err, SQ_EINVAL, "device %d out of range [0,%d)", device, n);
  sq_ctx* c = new sq_ctx();
  c->device = device;
  cudaDeviceProp prop;
  if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
    delete c;
    return fail(g_create_err, SQ_ECUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  }
  c->sm_count = prop.multiProcessorCount;
  c->l2_persist_max = size_t(prop.persistingL2CacheMaxSize);
  c->l2_window_max = size_t(prop.accessPolicyMaxWindowSize);
  *out = c;
  return SQ_OK;
}

SQ_API void sq_ctx_destroy(sq_ctx* ctx) {
  if (!ctx) return;
  if (ctx->pool) {  // indexes must be freed before their context; whatever the pool still caches goes back to the driver
    cudaSetDevice(ctx->device);
    cudaMemPoolDestroy(static_cast<cudaMemPool_t>(ctx->pool));
    cudaGetLastError();
  }
  delete ctx;
}

// ---- options: the `sequila.cuda_*` keys ---------------------------------------------------------
namespace {
struct OptDesc {
  const char* name;
  std::atomic<int> sq_options::*field;
  int lo, hi;
  const char* const* words;  // optional symbolic values, index = value
};
const char* const kLayoutWords[] = {"auto", "packed", "soa", nullptr};
const char* const kWireWords[] = {"rle", "copy", nullptr};
const char* const kStagedWords[] = {"auto", "on", "off", nullptr};
const char* const kOnOffWords[] = {"off", "on", "force", nullptr};
const char* const kIdsWords[] = {"rows", "positions", nullptr};
const char* const kSortWords[] = {"auto", "wide", nullptr};
const OptDesc kOpts[] = {
    {"cuda_probe_layout", &sq_options::probe_layout, 0, 2, kLayoutWords},
    {"cuda_probe_block", &sq_options::probe_block, 64, 256, nullptr},
    {"cuda_probe_tiles", &sq_options::probe_tiles, 1, 2, nullptr},
    {"cuda_lookback_backoff_ns", &sq_options::lookback_backoff_ns, 0, 1 << 20, nullptr},
    {"cuda_rows_per_bin", &sq_options::rows_per_bin, 0, 1024, nullptr},
    {"cuda_right_idx_wire", &sq_options::right_idx_wire, 0, 1, kWireWords},
    {"cuda_staged_probe", &sq_options::staged_probe, 0, 2, kStagedWords},
    {"cuda_scan_dict_capacity", &sq_options::scan_dict_capacity, 4, 1 << 30, nullptr},
    {"cuda_exec_trace", &sq_options::exec_trace, 0, 1, nullptr},
    {"cuda_pipeline_depth", &sq_options::pipeline_depth, 2, 8, nullptr},
    {"cuda_coalesce_rows", &sq_options::coalesce_rows, 1, 1 << 27, nullptr},
    {"cuda_rank_count", &sq_options::rank_count, 0, 2, kOnOffWords},
    {"cuda_build_ids", &sq_options::build_ids, 0, 1, kIdsWords},
    {"cuda_build_sort", &sq_options::build_sort, 0, 1, kSortWords},
};
const char* strip_prefix(const char* key) { return strncmp(key, "sequila.", 8) == 0 ? key + 8 : key; }
}  // namespace

SQ_API int32_t sq_ctx_set_option(sq_ctx* ctx, const char* key, const char* value) {
  if (!ctx) return SQ_EINVAL;
  if (!key || !value) return fail(ctx->err, SQ_EINVAL, "null option key or value");
  const char* k = strip_prefix(key);
  if (strcmp(k, "cuda_l2_persist_mb") == 0) {
    // L2 set-aside for the probe's directory (device-wide limit; measured on B200: no gain for the join, up to
    // 10 % for count-only launches) -- applied here, once, not per probe call
    char* endp = nullptr;
    const long mb = strtol(value, &endp, 10);
    if (endp == value || *endp || mb < 0) return fail(ctx->err, SQ_EINVAL, "option %s: '%s' is not a size in MiB", k, value);
    size_t want = size_t(mb) << 20;
    if (want > ctx->l2_persist_max) want = ctx->l2_persist_max;
    SQ_CUDA(ctx->err, cudaSetDevice(ctx->device));
    SQ_CUDA(ctx->err, cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
    ctx->l2_persist_bytes = want;
    return SQ_OK;
  }
  for (const OptDesc& d : kOpts) {
    if (strcmp(k, d.name) != 0) continue;
    int v = -1;
    bool ok = false;
    if (d.words)
      for (int i = 0; d.words[i]; ++i)
        if (strcasecmp(value, d.words[i]) == 0) { v = i; ok = true; }
    if (!ok) {
      char* endp = nullptr;
      const long x = strtol(value, &endp, 10);
      ok = endp != value && !*endp && x >= d.lo && x <= d.hi;
      v = int(x);
    }
    if (ok && d.field == &sq_options::probe_block) ok = v == 64 || v == 128 || v == 256;
    if (ok && d.field == &sq_options::scan_dict_capacity) ok = (v & (v - 1)) == 0;
    if (!ok) return fail(ctx->err, SQ_EINVAL, "option %s: invalid value '%s'", k, value);
    (ctx->opt.*(d.field)).store(v, std::memory_order_relaxed);
    return SQ_OK;
  }
  return fail(ctx->err, SQ_EINVAL, "unknown option '%s'", key);
}

SQ_API int32_t sq_ctx_get_option(sq_ctx* ctx, const char* key, char* value_out, size_t capacity) {
  if (!ctx) return SQ_EINVAL;
  if (!key || !value_out || capacity == 0) return fail(ctx->err, SQ_EINVAL, "null option key or output");
  const char* k = strip_prefix(key);
  if (strcmp(k, "cuda_l2_persist_mb") == 0) {
    snprintf(value_out, capacity, "%zu", ctx->l2_persist_bytes >> 20);
    return SQ_OK;
  }
  for (const OptDesc& d : kOpts) {
    if (strcmp(k, d.name) != 0) continue;
    const int v = (ctx->opt.*(d.field)).load(std::memory_order_relaxed);
    if (d.words) snprintf(value_out, capacity, "%s", d.words[v]);
    else snprintf(value_out, capacity, "%d", v);
    return SQ_OK;
  }
  return fail(ctx->err, SQ_EINVAL, "unknown option '%s'", key);
}

SQ_API const char* sq_last_error(const sq_ctx* ctx) {
  // the returned pointer stays valid until the next failing call on the same object
  return ctx ? ctx->err.msg.c_str() : g_create_err.msg.c_str();
}

// Process-wide: buffers handed to Arrow consumers are released whenever the consumer drops them, possibly
// after the context (and the exec node) that allocated them is gone.  Pinned memory is not tied to a device.
static HostPool& host_pool() {
  static HostPool* p = new HostPool();  // never destroyed: finalizers may release buffers during process exit
  return *p;
}

SQ_API int32_t sq_host_alloc(sq_ctx* ctx, size_t bytes, void** out) {
  if (!ctx || !out) return SQ_EINVAL;
  size_t cap = 4096;
  while (cap < bytes) cap <<= 1;
  HostPool& hp = host_pool();
  {
    std::lock_guard<std::mutex> g(hp.mu);
    auto it = hp.idle.find(cap);
    if (it != hp.idle.end()) {
      *out = it->second;
      hp.idle.erase(it);
      hp.idle_bytes -= cap;
      return SQ_OK;
    }
  }
  SQ_CUDA(ctx->err, cudaSetDevice(ctx->device));
  cudaError_t e = cudaHostAlloc(out, cap, cudaHostAllocDefault);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx->err, SQ_ENOMEM, "cudaHostAlloc(%zu): %s", cap, cudaGetErrorString(e)); }
  std::lock_guard<std::mutex> g(hp.mu);
  hp.capacity[*out] = cap;
  return SQ_OK;
}

SQ_API void sq_host_free(sq_ctx*, void* p) {
  if (!p) return;
  {  // the context may be gone already (Arrow arrays outlive the exec node that produced them)
    HostPool& hp = host_pool();
    std::lock_guard<std::mutex> g(hp.mu);
    auto it = hp.capacity.find(p);
    if (it != hp.capacity.end()) {
      constexpr size_t kKeep = size_t(8) << 30;  // pooled pinned memory is capped; beyond it buffers go back to the OS
      if (hp.idle_bytes + it->second <= kKeep) {
        hp.idle.emplace(it->second, p);
        hp.idle_bytes += it->second;
        return;
      }
      hp.capacity.erase(it);
    }
  }
  cudaFreeHost(p);
}

// ---- build ------------------------------------------------------------------------------------
SQ_API int32_t sq_index_build_device(sq_ctx* ctx, const uint64_t* d_key_hash, const int32_t* d_start,
                                     const int32_t* d_end, uint64_t n_rows, void* cuda_stream, sq_index** out) {
  if (!ctx || !out) return SQ_EINVAL;
  *out = nullptr;
  if (n_rows && (!d_key_hash || !d_start || !d_end)) return fail(ctx->err, SQ_EINVAL, "null build column");
  return build_index_device(ctx, d_key_hash, d_start, d_end, n_rows, static_cast<cudaStream_t>(cuda_stream), out);
}

SQ_API int32_t sq_index_build(sq_ctx* ctx, const uint64_t* key_hash, const int32_t* start, const int32_t* end,
                              uint64_t n_rows, sq_index** out) {
  if (!ctx || !out) return SQ_EINVAL;
  *out = nullptr;
  if (n_rows && (!key_hash || !start || !end)) return fail(ctx->err, SQ_EINVAL, "null build column");
  ErrorSlot& E = ctx->err;
  SQ_CUDA(E, cudaSetDevice(ctx->device));
  void* d = nullptr;
  const size_t n = n_rows ? n_rows : 1;
  SQ_CUDA(E, cudaMalloc(&d, n * 16));
  auto* dk = static_cast<uint64_t*>(d);
  auto* ds = reinterpret_cast<int32_t*>(dk + n);
  auto* de = ds + n;
  cudaStream_t st;
  cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  if (e != cudaSuccess) { cudaFree(d); return fail(E, SQ_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
  int rc = SQ_OK;
  if (n_rows) {
    if ((e = cudaMemcpyAsync(dk, key_hash, n_rows * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(ds, start, n_rows * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(de, end, n_rows * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess)
      rc = fail(E, SQ_ECUDA, "H2D copy of build columns: %s", cudaGetErrorString(e));
  }
  if (rc == SQ_OK) rc = build_index_device(ctx, dk, ds, de, n_rows, st, out);
  cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  cudaFree(d);
  return rc;
}

SQ_API uint64_t sq_index_bytes(const sq_index* idx) { return idx ? idx->bytes : 0; }
SQ_API uint64_t sq_index_rows(const sq_index* idx) { return idx ? idx->n_rows : 0; }
SQ_API uint64_t sq_index_keys(const sq_index* idx) { return idx ? idx->n_keys : 0; }
SQ_API int32_t sq_index_uses_packed(const sq_index* idx) { return idx && use_packed(idx) ? 1 : 0; }
SQ_API int32_t sq_index_uses_rank(const sq_index* idx) { return idx && use_rank(idx) ? 1 : 0; }
SQ_API float sq_index_build_ms(const sq_index* idx) { return idx ? idx->build_ms : 0.f; }
SQ_API void sq_index_free(sq_index* idx) { free_index(idx); }

SQ_API int32_t sq_index_sort_key_bits(const sq_index* idx) { return idx ? (idx->narrow_sort ? 32 : 64) : 0; }
SQ_API int32_t sq_index_uses_positions(const sq_index* idx) { return idx && idx->pos_ids ? 1 : 0; }

SQ_API int32_t sq_index_position_rows(const sq_index* idx, uint32_t* rows_out) {
  if (!idx) return SQ_EINVAL;
  ErrorSlot& E = idx->ctx->err;
  if (!idx->pos_ids) return fail(E, SQ_ESTATE, "the index hands out build rows, not positions (option cuda_build_ids)");
  if (idx->n_rows && !rows_out) return fail(E, SQ_EINVAL, "null rows_out");
  SQ_CUDA(E, cudaSetDevice(idx->ctx->device));
  if (idx->n_rows) SQ_CUDA(E, cudaMemcpy(rows_out, idx->d_perm, size_t(idx->n_rows) * 4, cudaMemcpyDeviceToHost));
  return SQ_OK;
}

SQ_API const uint32_t* sq_index_position_rows_device(const sq_index* idx) { return idx && idx->pos_ids ? idx->d_perm : nullptr; }

static int32_t stream_create(sq_ctx* ctx, cudaStream_t ext, bool use_ext, sq_stream** out);

// Position ids (option cuda_build_ids positions): payload is kept in the index's sorted order, value of position j =
// value of build row perm[j], so that gathers by the ids the probe kernels emit read neighbouring rows.  The columns are
// permuted once, when they are registered, with the take kernels themselves (indices = perm).
namespace {
// one scratch stream per index, created at the first permutation and kept (its pinned scalar slot and device scratch
// cost more to allocate than a small column costs to permute)
struct ScratchStream {
  sq_index* idx;
  std::unique_lock<std::mutex> lock;
  sq_stream* s = nullptr;
  explicit ScratchStream(sq_index* i) : idx(i), lock(i->perm_mu) {}
  int32_t open() {
    if (!idx->perm_stream) {
      const int32_t rc = stream_create(idx->ctx, nullptr, false, &idx->perm_stream);  // its message is in the context's slot
      if (rc) return rc;
    }
    s = idx->perm_stream;
    return SQ_OK;
  }
  int32_t done(int32_t rc) {
    if (rc == SQ_OK && cudaStreamSynchronize(s->stream) != cudaSuccess) rc = SQ_ECUDA;
    return rc ? fail(idx->ctx->err, rc, "permuting a payload column: %s", sq_stream_last_error(s)) : SQ_OK;
  }
};

// *d_values (n_rows values in build-row order) -> a new allocation in position order; the old one is freed when owned
int32_t permute_fixed(sq_index* idx, void** d_values, uint32_t width, bool owned) {
  ErrorSlot& E = idx->ctx->err;
  ScratchStream t(idx);
  int32_t rc = t.open();
  if (rc) return rc;
  void* d_new = nullptr;
  SQ_CUDA(E, cudaMalloc(&d_new, size_t(idx->n_rows ? idx->n_rows : 1) * width));
  rc = t.done(launch_gather(t.s, *d_values, idx->d_perm, idx->n_rows, width, d_new));
  if (rc) { cudaFree(d_new); return rc; }
  if (owned) cudaFree(*d_values);
  *d_values = d_new;
  return SQ_OK;
}

int32_t permute_utf8(sq_index* idx, int64_t** d_offsets, void** d_data, uint64_t data_bytes) {
  ErrorSlot& E = idx->ctx->err;
  ScratchStream t(idx);
  int32_t rc = t.open();
  if (rc) return rc;
  const size_t n = size_t(idx->n_rows);
  int64_t* d_off = nullptr;
  uint8_t* d_new = nullptr;
  SQ_CUDA(E, cudaMalloc(&d_off, (n + 1) * 8));
  cudaError_t e = cudaMalloc(&d_new, data_bytes ? data_bytes : 16);
  if (e != cudaSuccess) { cudaFree(d_off); return fail(E, SQ_ECUDA, "cudaMalloc: %s", cudaGetErrorString(e)); }
  uint64_t total = 0;
  rc = launch_str_offsets(t.s, *d_offsets, idx->d_perm, n, d_off, &total);
  if (rc == SQ_OK && total != data_bytes) rc = fail(t.s->err, SQ_EINVAL, "utf8 offsets cover %llu bytes, %llu given",
                                                    (unsigned long long)total, (unsigned long long)data_bytes);
  if (rc == SQ_OK) rc = launch_str_copy(t.s, *d_offsets, static_cast<const uint8_t*>(*d_data), idx->d_perm, n, d_off, d_new);
  rc = t.done(rc);
  if (rc) { cudaFree(d_off); cudaFree(d_new); return rc; }
  cudaFree(*d_offsets);
  cudaFree(*d_data);
  *d_offsets = d_off;
  *d_data = d_new;
  return SQ_OK;
}

int32_t permute_bits(sq_index* idx, uint8_t** d_bitmap) {
  ErrorSlot& E = idx->ctx->err;
  ScratchStream t(idx);
  int32_t rc = t.open();
  if (rc) return rc;
  const size_t bytes = (size_t(idx->n_rows) + 7) / 8;
  uint8_t* d_new = nullptr;
  SQ_CUDA(E, cudaMalloc(&d_new, bytes ? bytes : 16));
  uint64_t nulls = 0;
  rc = t.done(launch_gather_bits(t.s, *d_bitmap, idx->d_perm, idx->n_rows, d_new, &nulls));
  if (rc) { cudaFree(d_new); return rc; }
  cudaFree(*d_bitmap);
  *d_bitmap = d_new;
  return SQ_OK;
}
}  // namespace

SQ_API int32_t sq_index_add_column(sq_index* idx, const void* values, uint32_t width, int32_t* col_id_out) {
  if (!idx || !col_id_out) return SQ_EINVAL;
  ErrorSlot& E = idx->ctx->err;
  if (width != 4 && width != 8 && width != 16) return fail(E, SQ_EINVAL, "column width %u not in {4,8,16}", width);
  if (idx->n_rows && !values) return fail(E, SQ_EINVAL, "null column values");
  SQ_CUDA(E, cudaSetDevice(idx->ctx->device));
  sq_column c;
  c.width = width;
  c.owned = true;
  const size_t bytes = size_t(idx->n_rows ? idx->n_rows : 1) * width;
  SQ_CUDA(E, cudaMalloc(&c.d_values, bytes));
  if (idx->n_rows) {
    cudaError_t e = cudaMemcpy(c.d_values, values, size_t(idx->n_rows) * width, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(c.d_values); return fail(E, SQ_ECUDA, "H2D copy of build column: %s", cudaGetErrorString(e)); }
  }
  if (idx->pos_ids && idx->n_rows) {
    const int32_t rc = permute_fixed(idx, &c.d_values, width, true);
    if (rc) { cudaFree(c.d_values); return rc; }
  }
  std::lock_guard<std::mutex> g(idx->col_mu);
  idx->bytes += bytes;
  idx->columns.push_back(c);
  *col_id_out = int32_t(idx->columns.size() - 1);
  return SQ_OK;
}

SQ_API int32_t sq_index_add_column_device(sq_index* idx, const void* d_values, uint32_t width, int32_t* col_id_out) {
  if (!idx || !col_id_out) return SQ_EINVAL;
  ErrorSlot& E = idx->ctx->err;
  if (width != 4 && width != 8 && width != 16) return fail(E, SQ_EINVAL, "column width %u not in {4,8,16}", width);
  sq_column c;
  c.width = width;
  c.owned = false;
  c.d_values = const_cast<void*>(d_values);
  if (idx->pos_ids && idx->n_rows) {  // the index keeps its own copy in position order
    if (!d_values) return fail(E, SQ_EINVAL, "null column values");
    SQ_CUDA(E, cudaSetDevice(idx->ctx->device));
    const int32_t rc = permute_fixed(idx, &c.d_values, width, false);
    if (rc) return rc;
    c.owned = true;
  }
  std::lock_guard<std::mutex> g(idx->col_mu);
  if (c.owned) idx->bytes += size_t(idx->n_rows) * width;
  idx->columns.push_back(c);
  *col_id_out = int32_t(idx->columns.size() - 1);
  return SQ_OK;
}

// ---- streams ----------------------------------------------------------------------------------
static int32_t stream_create(sq_ctx* ctx, cudaStream_t ext, bool use_ext, sq_stream** out) {
  if (!ctx || !out) return SQ_EINVAL;
  *out = nullptr;
  SQ_CUDA(ctx->err, cudaSetDevice(ctx->device));
  sq_stream* s = new sq_stream();
  s->ctx = ctx;
  if (use_ext) {
    s->stream = ext;
  } else {
    cudaError_t e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete s; return fail(ctx->err, SQ_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    s->own_stream = true;
  }
  *out = s;
  return SQ_OK;
}

SQ_API int32_t sq_stream_create(sq_ctx* ctx, sq_stream** out) { return stream_create(ctx, nullptr, false, out); }
SQ_API int32_t sq_stream_create_on(sq_ctx* ctx, void* cuda_stream, sq_stream** out) {
  return stream_create(ctx, static_cast<cudaStream_t>(cuda_stream), true, out);
}

SQ_API void sq_stream_free(sq_stream* s) {
  if (!s) return;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->stream);
  pipeline_destroy(s);
  for (sq_buf* b : {&s->d_in, &s->d_cnt, &s->d_cnt8, &s->d_state, &s->d_tile, &s->d_scalar, &s->d_left,
                    &s->d_right, &s->d_gather, &s->d_gather2, &s->d_chain, &s->d_strblk, &s->d_strdata, &s->h_in, &s->h_out, &s->h_scalar, &s->h_scan})
    release(*b);
  if (s->ev_ready) for (auto& e : s->ev) cudaEventDestroy(e);
  if (s->ev_rle) cudaEventDestroy(s->ev_rle);
  if (s->own_stream) cudaStreamDestroy(s->stream);
  delete s;
}

SQ_API const char* sq_stream_last_error(const sq_stream* s) { return s ? s->err.msg.c_str() : ""; }

SQ_API uint64_t sq_stream_bytes(const sq_stream* s) {
  if (!s) return 0;
  uint64_t t = 0;
  for (const sq_buf* b : {&s->d_in, &s->d_cnt, &s->d_state, &s->d_tile, &s->d_scalar, &s->d_left,
                          &s->d_right, &s->d_gather, &s->d_gather2, &s->d_chain, &s->d_strblk, &s->d_strdata, &s->h_in, &s->h_out, &s->h_scalar, &s->h_scan})
    t += b->cap;
  return t + pipeline_bytes(s);
}

SQ_API int32_t sq_stream_set_profiling(sq_stream* s, int32_t enabled) {
  if (!s) return SQ_EINVAL;
  s->profiling = enabled != 0;
  s->pending = 0;
  for (int k = 0; k < 5; ++k) { s->phase_ms[k] = 0.f; s->phase_sum[k] = 0.0; s->phase_n[k] = 0; }
  if (s->profiling) { SQ_CUDA(s->err, cudaSetDevice(s->ctx->device)); return ensure_events(s); }
  return SQ_OK;
}

SQ_API int32_t sq_stream_phase_ms(sq_stream* s, float out5[5]) {
  if (!s || !out5) return SQ_EINVAL;
  if (s->profiling) {
    SQ_CUDA(s->err, cudaSetDevice(s->ctx->device));
    SQ_CUDA(s->err, cudaStreamSynchronize(s->stream));
    fold_phases(s);
  }
  for (int k = 0; k < 5; ++k) out5[k] = s->phase_n[k] ? float(s->phase_sum[k] / double(s->phase_n[k])) : 0.f;
  return SQ_OK;
}

SQ_API uint64_t sq_stream_launches(const sq_stream* s) { return s ? s->launches : 0; }
SQ_API const uint32_t* sq_stream_counts_device(const sq_stream* s) {
  return s ? static_cast<const uint32_t*>(s->d_cnt.p) : nullptr;
}

// ---- probe ------------------------------------------------------------------------------------
namespace sq {
// Which packed-line kernel serves a tile.  Measured on B200 (12.5M probe rows per launch, 100M-row index; DESIGN.md
// section 4): on position-sorted tiles as dense as the index (a CTA stages ~13 lines) the staged kernel
// (sq_probe_staged.cu) counts in 0.33 ms against 0.40 ms for k_probe_packed and joins in 0.76 ms against 0.77 ms; on
// sorted tiles 8x sparser than the index (73 lines per CTA, every build row decoded for one probe row in eight) it loses
// (join 0.98 vs 0.84 ms, count 0.50 vs 0.44 ms), and on scattered tiles its CTAs walk global memory thread by thread (3.8 vs
// 0.98 ms).  So: option on / off force it; auto sends only count-only launches of position-local, dense tiles to it.
// "Position-local" is judged from the order of the tile's host columns when the caller has them (1024 adjacent pairs:
// same key, non-decreasing start), from a sampling kernel for device tiles; "dense" from the kernel's own report —
// result[2] = CTAs that could not stage, result[3] = lines staged — backing off exponentially (2, 4, ... 256 tiles).
bool pick_staged(sq_stream* policy, const sq_index* idx, const uint64_t* host_key, const int32_t* host_start,
                 const uint64_t* d_key, const int32_t* d_start, uint32_t n, bool emits, const uint32_t* host_ids) {
  if (!idx->d_lines || n == 0) return false;
  const int opt = idx->ctx->opt.staged_probe.load(std::memory_order_relaxed);
  if (opt == 1) return true;
  if (opt == 2) return false;
  if (n < 64 || emits) return false;
  if (!host_key && !host_ids && !(d_key && d_start)) return false;  // nothing to judge the row order by
  if (!host_key && !host_ids && d_key && d_start) {
    // device tile (the caller synchronises on the result anyway): one 1024-thread sampling kernel, every 32nd tile
    if (policy->order_age == 0) {
      uint32_t r[2] = {0, 1};
      if (sample_order_device(policy, d_key, d_start, n, r) != SQ_OK) return false;
      policy->order_local = r[0] * 8u >= r[1] * 7u;
      policy->order_age = 32;
    }
    policy->order_age -= 1;
    if (!policy->order_local) return false;
  }
  if ((host_key || host_ids) && host_start) {
    const uint32_t m = n - 1 < 1024u ? n - 1 : 1024u;
    const uint32_t stride = (n - 1) / m;
    uint32_t ordered = 0;
    for (uint32_t k = 0; k < m; ++k) {
      const size_t a = size_t(k) * stride;
      const bool same = host_key ? host_key[a] == host_key[a + 1] : host_ids[a] == host_ids[a + 1];
      ordered += (same && host_start[a] <= host_start[a + 1]) ? 1u : 0u;
    }
    if (ordered * 8u < m * 7u) return false;  // not position-ordered: do not even try
  }
  if (policy->staged_skip) { policy->staged_skip -= 1; return false; }
  return true;
}

void staged_feedback(sq_stream* policy, uint32_t n_rows, uint64_t global_ctas, uint64_t staged_lines) {
  const uint64_t ctas = (uint64_t(n_rows) + staged_tile_rows() - 1) / staged_tile_rows();
  policy->staged_tiles += 1;
  policy->staged_global_ctas += global_ctas;
  // more than 1/16 of the CTAs walked global memory (scattered tiles), or the staged CTAs carried more than 32 lines each
  // (probe rows much sparser than the index: decoding every build row costs more than the cooperative walk)
  if (global_ctas * 16 > ctas || staged_lines > 32 * (ctas - global_ctas)) {
    policy->order_local = false;
    policy->staged_backoff = policy->staged_backoff ? (policy->staged_backoff < 128 ? policy->staged_backoff * 2 : 256) : 2;
    policy->staged_skip = policy->staged_backoff;
  } else {
    policy->staged_backoff = 0;
  }
}

int launch_packed_any(sq_stream* s, sq_stream* policy, bool staged, const sq_index* idx, const uint64_t* d_key,
                      const int32_t* d_start, const int32_t* d_end, uint32_t n, uint32_t* d_left, uint32_t* d_right,
                      uint64_t capacity) {
  (void)policy;
  if (staged) return launch_staged(s, idx, d_key, d_start, d_end, n, d_left, d_right, capacity);
  return launch_packed(s, idx, d_key, d_start, d_end, n, d_left, d_right, capacity);
}
}  // namespace sq

// Reads {n_pairs, overflow} written by the probe kernels and closes the count phase.
static int32_t finish_count(sq_stream* s, bool wrote, uint64_t* n_pairs_out, bool staged = false) {
  ErrorSlot& E = s->err;
  int rc;
  if ((rc = ensure(E, s->h_scalar, 256, true))) return rc;
  auto* h = static_cast<unsigned long long*>(s->h_scalar.p);
  SQ_CUDA(E, cudaMemcpyAsync(h, s->d_scalar.p, 32, cudaMemcpyDeviceToHost, s->stream));
  SQ_CUDA(E, cudaStreamSynchronize(s->stream));
  if (staged) staged_feedback(s, s->n_rows, h[2], h[3]);
  s->n_pairs = h[0];
  s->spec_valid = wrote && h[1] == 0;
  s->counted = true;
  s->emitted = false;
  if (s->profiling) { s->pending |= 3u; fold_phases(s); }
  *n_pairs_out = s->n_pairs;
  return SQ_OK;
}

static int32_t check_probe_args(sq_stream* s, const sq_index* idx, const void* k, const void* a, const void* b,
                                uint32_t n, uint64_t* out) {
  if (!s) return SQ_EINVAL;
  if (!idx || !out) return fail(s->err, SQ_EINVAL, "null index or output pointer");
  if (n && (!k || !a || !b)) return fail(s->err, SQ_EINVAL, "null probe column");
  if (idx->ctx->device != s->ctx->device) return fail(s->err, SQ_EINVAL, "index lives on device %d, stream on %d",
                                                       idx->ctx->device, s->ctx->device);
  return SQ_OK;
}

namespace sq {
void tile_begin(sq_stream* s, const sq_index* idx, const uint64_t* dk, const int32_t* ds, const int32_t* de,
                uint32_t n) {
  s->idx = idx;
  s->n_rows = n;
  s->d_q_key = dk;
  s->d_q_start = ds;
  s->d_q_end = de;
  s->counted = false;
  s->emitted = false;
  s->spec_valid = false;
  s->d_spec_left = s->d_spec_right = nullptr;
  s->win_set = false;
  s->win_off = s->win_n = 0;
}
}  // namespace sq
#define begin_tile sq::tile_begin

static int32_t empty_tile(sq_stream* s, uint64_t* n_pairs_out) {
  s->n_pairs = 0;
  s->counted = true;
  s->spec_valid = true;
  *n_pairs_out = 0;
  return SQ_OK;
}

// device tile: one fused pass; writes into (d_left, d_right) when the pairs fit `capacity`
static int32_t join_device(sq_stream* s, const sq_index* idx, const uint64_t* dk, const int32_t* ds,
                           const int32_t* de, uint32_t n, uint32_t* d_left, uint32_t* d_right, uint64_t capacity,
                           uint64_t* n_pairs_out, const uint64_t* host_key = nullptr, const int32_t* host_start = nullptr) {
  begin_tile(s, idx, dk, ds, de, n);
  if (n == 0) return empty_tile(s, n_pairs_out);
  int rc;
  mark(s, 1);
  const bool wrote = d_left != nullptr && capacity > 0;
  bool staged = false;
  if (use_packed(idx)) {
    // narrow index: ONE kernel searches, counts, scans and (when there is room) writes
    staged = pick_staged(s, idx, host_key, host_start, dk, ds, n, wrote);
    if ((rc = launch_packed_any(s, s, staged, idx, dk, ds, de, n, wrote ? d_left : nullptr, d_right, capacity))) return rc;
  } else if (use_rank(idx)) {
    // rank structure over the ends: ONE kernel — rank-difference count, chained scan, one walk that writes
    if ((rc = launch_rank_join(s, idx, dk, ds, de, n, wrote ? d_left : nullptr, d_right, capacity))) return rc;
  } else {
    if ((rc = launch_count(s, idx, dk, ds, de, n))) return rc;
    // speculative emit: the write kernel checks n_pairs <= capacity on the device, so the whole
    // count -> scan -> write chain is enqueued without a host round trip
    if (wrote && (rc = launch_write(s, idx, ds, n, d_left, d_right, capacity))) return rc;
  }
  mark(s, 2);
  s->d_spec_left = d_left;
  s->d_spec_right = d_right;
  return finish_count(s, wrote, n_pairs_out, staged);
}

SQ_API int32_t sq_probe_count_device(sq_stream* s, const sq_index* idx, const uint64_t* d_key_hash,
                                     const int32_t* d_start, const int32_t* d_end, uint32_t n_rows,
                                     uint64_t* n_pairs_out) {
  int rc = check_probe_args(s, idx, d_key_hash, d_start, d_end, n_rows, n_pairs_out);
  if (rc) return rc;
  SQ_CUDA(s->err, cudaSetDevice(s->ctx->device));
  mark(s, 0);
  return join_device(s, idx, d_key_hash, d_start, d_end, n_rows, nullptr, nullptr, 0, n_pairs_out);
}

SQ_API int32_t sq_probe_join_device(sq_stream* s, const sq_index* idx, const uint64_t* d_key_hash,
                                    const int32_t* d_start, const int32_t* d_end, uint32_t n_rows,
                                    uint32_t* d_left_idx_out, uint32_t* d_right_idx_out, uint64_t capacity,
                                    uint64_t* n_pairs_out) {
  int rc = check_probe_args(s, idx, d_key_hash, d_start, d_end, n_rows, n_pairs_out);
  if (rc) return rc;
  SQ_CUDA(s->err, cudaSetDevice(s->ctx->device));
  mark(s, 0);
  if ((rc = join_device(s, idx, d_key_hash, d_start, d_end, n_rows, d_left_idx_out, d_right_idx_out, capacity,
                        n_pairs_out)))
    return rc;
  if (!s->spec_valid)
    return fail(s->err, SQ_ECAPACITY, "output capacity %llu < %llu pairs", (unsigned long long)capacity,
                (unsigned long long)s->n_pairs);
  s->d_last_left = d_left_idx_out;
  s->d_last_right = d_right_idx_out;
  s->emitted = true;
  return SQ_OK;
}

// host tile: H2D, then the fused pass writing speculatively into the stream's own pair buffers
static int32_t count_host(sq_stream* s, const sq_index* idx, const uint64_t* key_hash, const int32_t* start,
                          const int32_t* end, uint32_t n_rows, uint64_t* n_pairs_out) {
  ErrorSlot& E = s->err;
  SQ_CUDA(E, cudaSetDevice(s->ctx->device));
  int rc;
  if (n_rows == 0) { begin_tile(s, idx, nullptr, nullptr, nullptr, 0); return empty_tile(s, n_pairs_out); }
  const size_t n = n_rows;
  if ((rc = ensure(E, s->d_in, n * 16, false))) return rc;
  auto* dk = static_cast<uint64_t*>(s->d_in.p);
  auto* ds = reinterpret_cast<int32_t*>(dk + n);
  auto* de = ds + n;
  if (!s->d_left.p) {  // first tile: guess 4 pairs per probe row; later tiles reuse the grown buffers
    if ((rc = ensure(E, s->d_left, n * 16, false))) return rc;
    if ((rc = ensure(E, s->d_right, n * 16, false))) return rc;
  }
  const uint64_t cap = (s->d_left.cap < s->d_right.cap ? s->d_left.cap : s->d_right.cap) / 4;
  mark(s, 0);
  SQ_CUDA(E, cudaMemcpyAsync(dk, key_hash, n * 8, cudaMemcpyHostToDevice, s->stream));
  SQ_CUDA(E, cudaMemcpyAsync(ds, start, n * 4, cudaMemcpyHostToDevice, s->stream));
  SQ_CUDA(E, cudaMemcpyAsync(de, end, n * 4, cudaMemcpyHostToDevice, s->stream));
  return join_device(s, idx, dk, ds, de, n_rows, static_cast<uint32_t*>(s->d_left.p),
                     static_cast<uint32_t*>(s->d_right.p), cap, n_pairs_out, key_hash, start);
}

SQ_API int32_t sq_probe_count(sq_stream* s, const sq_index* idx, const uint64_t* key_hash, const int32_t* start,
                              const int32_t* end, uint32_t n_rows, uint64_t* n_pairs_out) {
  int rc = check_probe_args(s, idx, key_hash, start, end, n_rows, n_pairs_out);
  if (rc) return rc;
  return count_host(s, idx, key_hash, start, end, n_rows, n_pairs_out);
}

// the write pass of a tile that was only counted (or whose speculative buffers were too small)
namespace sq {
int tile_emit(sq_stream* s, uint32_t* d_left, uint32_t* d_right, uint64_t capacity) {
  if (use_packed(s->idx))
    return launch_packed(s, s->idx, s->d_q_key, s->d_q_start, s->d_q_end, s->n_rows, d_left, d_right, capacity);
  if (use_rank(s->idx))
    return launch_rank_join(s, s->idx, s->d_q_key, s->d_q_start, s->d_q_end, s->n_rows, d_left, d_right, capacity);
  return launch_write(s, s->idx, s->d_q_start, s->n_rows, d_left, d_right, capacity);
}
}  // namespace sq
#define launch_emit sq::tile_emit


// right_idx[k] = the probe row of pair k: row i repeated counts[i] times — sq::expand_counts (sq_rle.cpp, host
// compiler: SSE2 / AVX2 / AVX-512 variants picked at run time)

// On by default (option cuda_right_idx_wire=copy copies right_idx itself).  Measured on a B200 box with 16 host cores, 12.5M
// probe rows / 80.6M pairs per step: one host thread decodes ~1 G pairs/s, so with 4 partition threads the
// decode is the bottleneck (18.4 ms per step vs 13.7 ms with plain copies), with 8 it is not (11.4 ms vs
// 13.7 ms); it pays most where several GPUs share the host link (8 GPUs: 60 ms per step with plain copies).
static bool rle_on_the_wire(const sq_stream* s) {
  return s->ctx->opt.right_idx_wire.load(std::memory_order_relaxed) == 0;
}

static int32_t check_emit(sq_stream* s, const void* left, uint64_t capacity) {
  if (!s) return SQ_EINVAL;
  if (!s->counted) return fail(s->err, SQ_ESTATE, "sq_probe_emit_pairs called without a preceding sq_probe_count");
  if (capacity < s->n_pairs)
    return fail(s->err, SQ_ECAPACITY, "output capacity %llu < %llu pairs", (unsigned long long)capacity,
                (unsigned long long)s->n_pairs);
  (void)left;  // NULL left_idx_out = keep the pairs on the device only (a gather follows)
  return SQ_OK;
}

SQ_API int32_t sq_probe_emit_pairs_device(sq_stream* s, uint32_t* d_left_idx_out, uint32_t* d_right_idx_out,
                                          uint64_t capacity) {
  int rc = check_emit(s, d_left_idx_out, capacity);
  if (rc) return rc;
  if (s->n_pairs && !d_left_idx_out) return fail(s->err, SQ_EINVAL, "null d_left_idx_out");
  SQ_CUDA(s->err, cudaSetDevice(s->ctx->device));
  mark(s, 3);
  if (s->n_pairs && (rc = launch_emit(s, d_left_idx_out, d_right_idx_out, capacity))) return rc;
  mark(s, 4);
  s->d_last_left = d_left_idx_out;
  s->d_last_right = d_right_idx_out;
  s->emitted = true;
  if (s->profiling) s->pending |= 4u;  // read at the next synchronisation point
  return SQ_OK;
}

SQ_API int32_t sq_probe_emit_pairs(sq_stream* s, uint32_t* left_idx_out, uint32_t* right_idx_out,
                                   uint32_t* counts_out, uint64_t capacity) {
  int rc = check_emit(s, left_idx_out, capacity);
  if (rc) return rc;
  ErrorSlot& E = s->err;
  SQ_CUDA(E, cudaSetDevice(s->ctx->device));
  const size_t np = s->n_pairs;
  mark(s, 3);
  const bool reuse = s->spec_valid && s->d_spec_left == s->d_left.p && s->d_spec_right == s->d_right.p;
  if (np && !reuse) {  // the speculative buffers were too small: grow them, re-run only the write kernel
    if ((rc = ensure(E, s->d_left, np * 4, false))) return rc;
    if ((rc = ensure(E, s->d_right, np * 4, false))) return rc;
    if ((rc = launch_emit(s, static_cast<uint32_t*>(s->d_left.p), static_cast<uint32_t*>(s->d_right.p), np))) return rc;
  }
  auto* dl = static_cast<uint32_t*>(s->d_left.p);
  auto* dr = static_cast<uint32_t*>(s->d_right.p);
  mark(s, 4);
  // right_idx is a run-length expansion of the per-row counts (IJ:1611-1618); optionally it travels as the
  // counts (4 B per probe row instead of 4 B per pair) and is decoded into the caller's buffer here while
  // left_idx is still arriving (see rle_on_the_wire).
  const bool rle_wire = right_idx_out && np && rle_on_the_wire(s);
  const uint32_t* h_counts = counts_out;
  if ((counts_out || rle_wire) && s->n_rows) {
    if (!counts_out) {
      if ((rc = ensure(E, s->h_out, size_t(s->n_rows) * 4, true))) return rc;
      h_counts = static_cast<const uint32_t*>(s->h_out.p);
    }
    SQ_CUDA(E, cudaMemcpyAsync(const_cast<uint32_t*>(h_counts), s->d_cnt.p, size_t(s->n_rows) * 4, cudaMemcpyDeviceToHost,
                               s->stream));
    if (rle_wire) {
      if (!s->ev_rle) SQ_CUDA(E, cudaEventCreateWithFlags(&s->ev_rle, cudaEventDisableTiming));
      SQ_CUDA(E, cudaEventRecord(s->ev_rle, s->stream));
    }
  }
  if (np) {
    if (left_idx_out) SQ_CUDA(E, cudaMemcpyAsync(left_idx_out, dl, np * 4, cudaMemcpyDeviceToHost, s->stream));
    if (right_idx_out && !rle_wire) SQ_CUDA(E, cudaMemcpyAsync(right_idx_out, dr, np * 4, cudaMemcpyDeviceToHost, s->stream));
  }
  if (rle_wire) {
    SQ_CUDA(E, cudaEventSynchronize(s->ev_rle));
    expand_counts(h_counts, s->n_rows, right_idx_out, np);
  }
  mark(s, 5);
  SQ_CUDA(E, cudaStreamSynchronize(s->stream));
  s->d_last_left = dl;
  s->d_last_right = dr;
  s->emitted = true;
  if (s->profiling) { s->pending |= 12u; fold_phases(s); }
  return SQ_OK;
}

SQ_API int32_t sq_probe_join(sq_stream* s, const sq_index* idx, const uint64_t* key_hash, const int32_t* start,
                             const int32_t* end, uint32_t n_rows, uint32_t* left_idx_out, uint32_t* right_idx_out,
                             uint32_t* counts_out, uint64_t capacity, uint64_t* n_pairs_out) {
  int rc = check_probe_args(s, idx, key_hash, start, end, n_rows, n_pairs_out);
  if (rc) return rc;
  if ((rc = count_host(s, idx, key_hash, start, end, n_rows, n_pairs_out))) return rc;
  return sq_probe_emit_pairs(s, left_idx_out, right_idx_out, counts_out, capacity);
}

// ---- gather -----------------------------------------------------------------------------------
// Bounded output: gathers and fetches see the window [win_off, win_off + win_n) of the emitted pairs
// (the whole tile unless sq_stream_set_window narrowed it).
static inline uint64_t win_pairs(const sq_stream* s) { return s->win_set ? s->win_n : s->n_pairs; }

static int32_t gather_source(sq_stream* s, int32_t side, int32_t build_col_id, uint32_t* width,
                             const void** d_values, const uint32_t** d_idx) {
  if (!s->emitted) return fail(s->err, SQ_ESTATE, "sq_gather_column called without a preceding emit");
  if (side == 0) {
    sq_index* idx = const_cast<sq_index*>(s->idx);
    std::lock_guard<std::mutex> g(idx->col_mu);
    if (build_col_id < 0 || size_t(build_col_id) >= idx->columns.size())
      return fail(s->err, SQ_EINVAL, "unknown build column id %d", build_col_id);
    *width = idx->columns[build_col_id].width;
    *d_values = idx->columns[build_col_id].d_values;
    *d_idx = s->d_last_left ? s->d_last_left + s->win_off : nullptr;
  } else if (side == 1) {
    if (!s->d_last_right && win_pairs(s)) return fail(s->err, SQ_ESTATE, "right indices were not emitted on the device");
    *d_idx = s->d_last_right ? s->d_last_right + s->win_off : nullptr;
  } else {
    return fail(s->err, SQ_EINVAL, "side must be 0 (build) or 1 (probe)");
  }
  return SQ_OK;
}

SQ_API int32_t sq_gather_column_device(sq_stream* s, int32_t side, int32_t build_col_id, const void* d_probe_values,
                                       uint32_t width, void* d_out, uint64_t capacity) {
  if (!s) return SQ_EINVAL;
  const void* src = d_probe_values;
  const uint32_t* ix = nullptr;
  int rc = gather_source(s, side, build_col_id, &width, &src, &ix);
  if (rc) return rc;
  if (capacity < win_pairs(s)) return fail(s->err, SQ_ECAPACITY, "gather capacity %llu < %llu", (unsigned long long)capacity,
                                         (unsigned long long)win_pairs(s));
  SQ_CUDA(s->err, cudaSetDevice(s->ctx->device));
  mark(s, 6);
  rc = launch_gather(s, src, ix, win_pairs(s), width, d_out);
  mark(s, 7);
  if (rc == SQ_OK && s->profiling) {  // one gather per column: fold each (costs a sync; profiling only)
    SQ_CUDA(s->err, cudaStreamSynchronize(s->stream));
    s->pending |= 16u;
    fold_phases(s);
  }
  return rc;
}

SQ_API int32_t sq_index_pack_columns(sq_index* idx, const int32_t* col_ids, int32_t n_cols, int32_t* pack_id_out) {
  if (!idx || !pack_id_out) return SQ_EINVAL;
  ErrorSlot& E = idx->ctx->err;
  if (!col_ids || n_cols < 1 || n_cols > 4) return fail(E, SQ_EINVAL, "a pack holds 1 to 4 columns, got %d", n_cols);
  const uint32_t* cols[4] = {nullptr, nullptr, nullptr, nullptr};
  {
    std::lock_guard<std::mutex> g(idx->col_mu);
    for (int c = 0; c < n_cols; ++c) {
      if (col_ids[c] < 0 || size_t(col_ids[c]) >= idx->columns.size())
        return fail(E, SQ_EINVAL, "unknown build column id %d", col_ids[c]);
      const sq_column& col = idx->columns[size_t(col_ids[c])];
      if (col.width != 4) return fail(E, SQ_EINVAL, "column %d is %u bytes wide: a pack holds 4-byte columns", col_ids[c], col.width);
      cols[c] = static_cast<const uint32_t*>(col.d_values);
    }
  }
  SQ_CUDA(E, cudaSetDevice(idx->ctx->device));
  sq_pack pk;
  pk.n_cols = n_cols;
  const size_t bytes = size_t(idx->n_rows ? idx->n_rows : 1) * 16;
  SQ_CUDA(E, cudaMalloc(&pk.d_rows, bytes));
  int rc = launch_pack_columns(idx->ctx, cols, n_cols, idx->n_rows, pk.d_rows);
  if (rc) { cudaFree(pk.d_rows); return rc; }
  std::lock_guard<std::mutex> g(idx->col_mu);
  idx->bytes += bytes;
  idx->packs.push_back(pk);
  *pack_id_out = int32_t(idx->packs.size() - 1);
  return SQ_OK;
}

SQ_API int32_t sq_gather_pack_device(sq_stream* s, int32_t pack_id, void* const* d_outs, int32_t n_outs, uint64_t capacity) {
  if (!s) return SQ_EINVAL;
  ErrorSlot& E = s->err;
  if (!s->emitted) return fail(E, SQ_ESTATE, "sq_gather_pack_device called without a preceding emit");
  sq_index* idx = const_cast<sq_index*>(s->idx);
  sq_pack pk;
  {
    std::lock_guard<std::mutex> g(idx->col_mu);
    if (pack_id < 0 || size_t(pack_id) >= idx->packs.size()) return fail(E, SQ_EINVAL, "unknown pack id %d", pack_id);
    pk = idx->packs[size_t(pack_id)];
  }
  if (!d_outs || n_outs < 1 || n_outs > pk.n_cols) return fail(E, SQ_EINVAL, "the pack holds %d columns, %d outputs given", pk.n_cols, n_outs);
  const uint64_t n = win_pairs(s);
  if (capacity < n) return fail(E, SQ_ECAPACITY, "gather capacity %llu < %llu", (unsigned long long)capacity, (unsigned long long)n);
  if (n == 0) return SQ_OK;
  if (!s->d_last_left) return fail(E, SQ_ESTATE, "left indices were not emitted on the device");
  uint32_t* outs[4] = {nullptr, nullptr, nullptr, nullptr};
  for (int c = 0; c < n_outs; ++c) {
    if (!d_outs[c]) return fail(E, SQ_EINVAL, "null gather output %d", c);
    outs[c] = static_cast<uint32_t*>(d_outs[c]);
  }
  SQ_CUDA(E, cudaSetDevice(s->ctx->device));
  return launch_gather_pack(s, pk.d_rows, s->d_last_left + s->win_off, n, outs, n_outs);
}

SQ_API int32_t sq_gather_probe_columns_device(sq_stream* s, const void* const* d_probe_values, void* const* d_outs,
                                              int32_t n_cols, uint64_t capacity) {
  if (!s) return SQ_EINVAL;
  ErrorSlot& E = s->err;
  if (!s->emitted) return fail(E, SQ_ESTATE, "sq_gather_probe_columns_device called without a preceding emit");
  if (!d_probe_values || !d_outs || n_cols < 1 || n_cols > 4) return fail(E, SQ_EINVAL, "1 to 4 probe columns per pass, got %d", n_cols);
  const uint64_t n = win_pairs(s);
  if (capacity < n) return fail(E, SQ_ECAPACITY, "gather capacity %llu < %llu", (unsigned long long)capacity, (unsigned long long)n);
  if (n == 0) return SQ_OK;
  if (!s->d_last_right) return fail(E, SQ_ESTATE, "right indices were not emitted on the device");
  const uint32_t* cols[4] = {nullptr, nullptr, nullptr, nullptr};
  uint32_t* outs[4] = {nullptr, nullptr, nullptr, nullptr};
  for (int c = 0; c < n_cols; ++c) {
    if (!d_probe_values[c] || !d_outs[c]) return fail(E, SQ_EINVAL, "null probe column / output %d", c);
    cols[c] = static_cast<const uint32_t*>(d_probe_values[c]);
    outs[c] = static_cast<uint32_t*>(d_outs[c]);
  }
  SQ_CUDA(E, cudaSetDevice(s->ctx->device));
  return launch_gather_probe_columns(s, cols, s->d_last_right + s->win_off, n, outs, n_cols);
}

SQ_API int32_t sq_gather_column(sq_stream* s, int32_t side, int32_t build_col_id, const void* probe_values,
                                uint32_t width, void* out, uint64_t capacity) {
  if (!s) return SQ_EINVAL;
  ErrorSlot& E = s->err;
  const void* src = nullptr;
  const uint32_t* ix = nullptr;
  int rc = gather_source(s, side, build_col_id, &width, &src, &ix);
  if (rc) return rc;
  if (width != 4 && width != 8 && width != 16) return fail(E, SQ_EINVAL, "gather: unsupported value width %u", width);
  if (capacity < win_pairs(s)) return fail(E, SQ_ECAPACITY, "gather capacity %llu < %llu", (unsigned long long)capacity,
                                         (unsigned long long)win_pairs(s));
  if (win_pairs(s) == 0) return SQ_OK;
  if (!out) return fail(E, SQ_EINVAL, "null gather output");
  SQ_CUDA(E, cudaSetDevice(s->ctx->device));
  const size_t out_bytes = size_t(win_pairs(s)) * width;
  const size_t in_bytes = side == 1 ? size_t(s->n_rows) * width : 0;
  if ((rc = ensure(E, s->d_gather, out_bytes + in_bytes + 32, false))) return rc;
  char* d_out = static_cast<char*>(s->d_gather.p);
  if (side == 1) {
    if (!probe_values) return fail(E, SQ_EINVAL, "null probe column values");
    char* d_src = d_out + ((out_bytes + 15) & ~size_t(15));
    SQ_CUDA(E, cudaMemcpyAsync(d_src, probe_values, in_bytes, cudaMemcpyHostToDevice, s->stream));
    src = d_src;
  }
  mark(s, 6);
  if ((rc = launch_gather(s, src, ix, win_pairs(s), width, d_out))) return rc;
  mark(s, 7);
  SQ_CUDA(E, cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, s->stream));
  SQ_CUDA(E, cudaStreamSynchronize(s->stream));
  if (s->profiling) { s->pending |= 16u; fold_phases(s); }
  return SQ_OK;
}


// ---- nearest (Algorithm::CoitreesNearest on the flat index) ------------------------------------
static int32_t nearest_device(sq_stream* s, const sq_index* idx, const uint64_t* dk, const int32_t* ds,
                              const int32_t* de, uint32_t n, uint32_t* d_left) {
  ErrorSlot& E = s->err;
  begin_tile(s, idx, dk, ds, de, n);
  int rc;
  if ((rc = ensure(E, s->d_right, size_t(n) * 4, false))) return rc;
  if ((rc = launch_nearest(s, idx, dk, ds, de, n, d_left))) return rc;
  if ((rc = launch_iota(s, static_cast<uint32_t*>(s->d_right.p), n))) return rc;
  s->n_pairs = n;  // one output row per probe row (interval_join.rs:1593-1602)
  s->counted = true;
  s->emitted = true;
  s->d_last_left = d_left;
  s->d_last_right = static_cast<const uint32_t*>(s->d_right.p);
  return SQ_OK;
}

SQ_API int32_t sq_probe_nearest_device(sq_stream* s, const sq_index* idx, const uint64_t* d_key_hash,
                                       const int32_t* d_start, const int32_t* d_end, uint32_t n_rows,
                                       uint32_t* d_left_idx_out) {
  uint64_t dummy = 0;
  int rc = check_probe_args(s, idx, d_key_hash, d_start, d_end, n_rows, &dummy);
  if (rc) return rc;
  if (n_rows && !d_left_idx_out) return fail(s->err, SQ_EINVAL, "null d_left_idx_out");
  SQ_CUDA(s->err, cudaSetDevice(s->ctx->device));
  return nearest_device(s, idx, d_key_hash, d_start, d_end, n_rows, d_left_idx_out);
}

SQ_API int32_t sq_probe_nearest(sq_stream* s, const sq_index* idx, const uint64_t* key_hash, const int32_t* start,
                                const int32_t* end, uint32_t n_rows, uint32_t* left_idx_out) {
  uint64_t dummy = 0;
  int rc = check_probe_args(s, idx, key_hash, start, end, n_rows, &dummy);
  if (rc) return rc;
  ErrorSlot& E = s->err;
  SQ_CUDA(E, cudaSetDevice(s->ctx->device));
  const size_t n = n_rows;
  if (n == 0) { begin_tile(s, idx, nullptr, nullptr, nullptr, 0); s->n_pairs = 0; s->counted = s->emitted = true; return SQ_OK; }
  if ((rc = ensure(E, s->d_in, n * 16, false))) return rc;
  if ((rc = ensure(E, s->d_left, n * 4, false))) return rc;
  auto* dk = static_cast<uint64_t*>(s->d_in.p);
  auto* ds = reinterpret_cast<int32_t*>(dk + n);
  auto* de = ds + n;
  SQ_CUDA(E, cudaMemcpyAsync(dk, key_hash, n * 8, cudaMemcpyHostToDevice, s->stream));
  SQ_CUDA(E, cudaMemcpyAsync(ds, start, n * 4, cudaMemcpyHostToDevice, s->stream));
  SQ_CUDA(E, cudaMemcpyAsync(de, end, n * 4, cudaMemcpyHostToDevice, s->stream));
  auto* dl = static_cast<uint32_t*>(s->d_left.p);
  if ((rc = nearest_device(s, idx, dk, ds, de, n_rows, dl))) return rc;
  if (left_idx_out) SQ_CUDA(E, cudaMemcpyAsync(left_idx_out, dl, n * 4, cudaMemcpyDeviceToHost, s->stream));
  SQ_CUDA(E, cudaStreamSynchronize(s->stream));
  return SQ_OK;
}

// ---- Utf8 / validity take ---------------------------------------------------------------------------
SQ_API int32_t sq_index_add_utf8_column(sq_index* idx, const int64_t* offsets, const uint8_t* data, uint64_t data_bytes,
                                        int32_t* col_id_out) {
  if (!idx || !col_id_out) return SQ_EINVAL;
  ErrorSlot& E = idx->ctx->err;
  if (!offsets) return fail(E, SQ_EINVAL, "null utf8 offsets");
  SQ_CUDA(E, cudaSetDevice(idx->ctx->device));
  sq_column c;
  c.width = 0;
  c.owned = true;
  c.data_bytes = data_bytes;
  const size_t n = size_t(idx->n_rows);
  SQ_CUDA(E, cudaMalloc(&c.d_offsets, (n + 1) * 8));
  SQ_CUDA(E, cudaMalloc(&c.d_values, data_bytes ? data_bytes : 16));
  cudaError_t e = cudaMemcpy(c.d_offsets, offsets, (n + 1) * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && data_bytes) e = cudaMemcpy(c.d_values, data, data_bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(c.d_offsets); cudaFree(c.d_values);
    return fail(E, SQ_ECUDA, "H2D copy of utf8 build column: %s", cudaGetErrorString(e));
  }
  if (idx->pos_ids && n) {
    const int32_t rc = permute_utf8(idx, &c.d_offsets, &c.d_values, data_bytes);
    if (rc) { cudaFree(c.d_offsets); cudaFree(c.d_values); return rc; }
  }
  std::lock_guard<std::mutex> g(idx->col_mu);
  idx->bytes += (n + 1) * 8 + data_bytes;
  idx->columns.push_back(c);
  *col_id_out = int32_t(idx->columns.size() - 1);
  return SQ_OK;
}

SQ_API int32_t sq_index_set_validity(sq_index* idx, int32_t col_id, const uint8_t* bitmap) {
  if (!idx) return SQ_EINVAL;
  ErrorSlot& E = idx->ctx->err;
  std::lock_guard<std::mutex> g(idx->col_mu);
  if (col_id < 0 || size_t(col_id) >= idx->columns.size()) return fail(E, SQ_EINVAL, "unknown build column id %d", col_id);
  if (!bitmap) return fail(E, SQ_EINVAL, "null validity bitmap");
  SQ_CUDA(E, cudaSetDevice(idx->ctx->device));
  const size_t bytes = (size_t(idx->n_rows) + 7) / 8;
  uint8_t* d = nullptr;
  SQ_CUDA(E, cudaMalloc(&d, bytes ? bytes : 16));
  cudaError_t e = cudaMemcpy(d, bitmap, bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(d); return fail(E, SQ_ECUDA, "H2D copy of validity: %s", cudaGetErrorString(e)); }
  if (idx->pos_ids && idx->n_rows) {
    const int32_t rc = permute_bits(idx, &d);
    if (rc) { cudaFree(d); return rc; }
  }
  if (idx->columns[col_id].d_validity) cudaFree(idx->columns[col_id].d_validity);
  idx->columns[col_id].d_validity = d;
  idx->bytes += bytes;
  return SQ_OK;
}

static int32_t pick_indices(sq_stream* s, int32_t side, const uint32_t** ix) {
  if (!s->emitted) return fail(s->err, SQ_ESTATE, "gather called without a preceding emit");
  if (side == 0) *ix = s->d_last_left ? s->d_last_left + s->win_off : nullptr;
  else if (side == 1) {
    if (!s->d_last_right && win_pairs(s)) return fail(s->err, SQ_ESTATE, "right indices were not emitted on the device");
    *ix = s->d_last_right ? s->d_last_right + s->win_off : nullptr;
  } else return fail(s->err, SQ_EINVAL, "side must be 0 (build) or 1 (probe)");
  return SQ_OK;
}

SQ_API int32_t sq_gather_utf8(sq_stream* s, int32_t side, int32_t build_col_id, const int64_t* probe_offsets,
                              const uint8_t* probe_data, uint64_t probe_data_bytes, int32_t* out_offsets,
                              uint64_t* total_bytes_out) {
  if (!s || !total_bytes_out) return SQ_EINVAL;
  ErrorSlot& E = s->err;
  const uint32_t* ix = nullptr;
  int rc = pick_indices(s, side, &ix);
  if (rc) return rc;
  SQ_CUDA(E, cudaSetDevice(s->ctx->device));
  const int64_t* d_off = nullptr;
  const uint8_t* d_data = nullptr;
  if (side == 0) {
    sq_index* idx = const_cast<sq_index*>(s->idx);
    std::lock_guard<std::mutex> g(idx->col_mu);
    if (build_col_id < 0 || size_t(build_col_id) >= idx->columns.size() || !idx->columns[build_col_id].d_offsets)
      return fail(E, SQ_EINVAL, "build column %d is not a utf8 column", build_col_id);
    d_off = idx->columns[build_col_id].d_offsets;
    d_data = static_cast<const uint8_t*>(idx->columns[build_col_id].d_values);
  } else {
    if (s->n_rows && !probe_offsets) return fail(E, SQ_EINVAL, "null probe utf8 offsets");
    const size_t off_bytes = (size_t(s->n_rows) + 1) * 8;
    if ((rc = ensure(E, s->d_gather2, off_bytes + probe_data_bytes + 32, false))) return rc;
    auto* p = static_cast<char*>(s->d_gather2.p);
    SQ_CUDA(E, cudaMemcpyAsync(p, probe_offsets, off_bytes, cudaMemcpyHostToDevice, s->stream));
    if (probe_data_bytes)
      SQ_CUDA(E, cudaMemcpyAsync(p + off_bytes, probe_data, probe_data_bytes, cudaMemcpyHostToDevice, s->stream));
    d_off = reinterpret_cast<const int64_t*>(p);
    d_data = reinterpret_cast<const uint8_t*>(p + off_bytes);
  }
  const uint64_t np = win_pairs(s);
  if ((rc = ensure(E, s->d_gather, (np + 1) * 12 + 32, false))) return rc;
  auto* d_out_off = static_cast<int64_t*>(s->d_gather.p);
  uint64_t total = 0;
  if ((rc = launch_str_offsets(s, d_off, ix, np, d_out_off, &total))) return rc;
  if (total > 0x7FFFFFFFull)
    return fail(E, SQ_ECAPACITY, "gathered utf8 column needs %llu bytes; Utf8 offsets are 32-bit", (unsigned long long)total);
  if (out_offsets) {  // Arrow Utf8 offsets are 32-bit: narrow on the device, copy straight into the caller's buffer
    auto* d_off32 = reinterpret_cast<int32_t*>(d_out_off + np + 1);
    if ((rc = launch_narrow_offsets(s, d_out_off, np + 1, d_off32))) return rc;
    SQ_CUDA(E, cudaMemcpyAsync(out_offsets, d_off32, (np + 1) * 4, cudaMemcpyDeviceToHost, s->stream));
    SQ_CUDA(E, cudaStreamSynchronize(s->stream));
  }
  s->str_src_off = d_off;
  s->str_src_data = d_data;
  s->str_idx = ix;
  s->str_out_off = d_out_off;
  s->str_total = total;
  s->str_pending = true;
  *total_bytes_out = total;
  return SQ_OK;
}

SQ_API int32_t sq_gather_utf8_data(sq_stream* s, uint8_t* out_data, uint64_t capacity) {
  if (!s) return SQ_EINVAL;
  ErrorSlot& E = s->err;
  if (!s->str_pending) return fail(E, SQ_ESTATE, "sq_gather_utf8_data without sq_gather_utf8");
  if (capacity < s->str_total) return fail(E, SQ_ECAPACITY, "utf8 data capacity %llu < %llu", (unsigned long long)capacity,
                                           (unsigned long long)s->str_total);
  s->str_pending = false;
  if (s->str_total == 0) return SQ_OK;
  if (!out_data) return fail(E, SQ_EINVAL, "null utf8 output");
  SQ_CUDA(E, cudaSetDevice(s->ctx->device));
  int rc;
  if ((rc = ensure(E, s->d_strdata, s->str_total, false))) return rc;
  auto* d_out = static_cast<uint8_t*>(s->d_strdata.p);
  if ((rc = launch_str_copy(s, s->str_src_off, s->str_src_data, s->str_idx, win_pairs(s), s->str_out_off, d_out))) return rc;
  SQ_CUDA(E, cudaMemcpyAsync(out_data, d_out, s->str_total, cudaMemcpyDeviceToHost, s->stream));
  SQ_CUDA(E, cudaStreamSynchronize(s->stream));
  return SQ_OK;
}

SQ_API int32_t sq_gather_validity(sq_stream* s, int32_t side, int32_t build_col_id, const uint8_t* probe_bitmap,
                                  uint8_t* out_bitmap, uint64_t* null_count_out) {
  if (!s || !null_count_out) return SQ_EINVAL;
  ErrorSlot& E = s->err;
  const uint32_t* ix = nullptr;
  int rc = pick_indices(s, side, &ix);
  if (rc) return rc;
  SQ_CUDA(E, cudaSetDevice(s->ctx->device));
  const uint64_t np = win_pairs(s);
  const size_t out_bytes = (np + 7) / 8;
  const size_t in_bytes = side == 1 ? (size_t(s->n_rows) + 7) / 8 : 0;
  if ((rc = ensure(E, s->d_gather, out_bytes + in_bytes + 64, false))) return rc;
  auto* d_out = static_cast<uint8_t*>(s->d_gather.p);
  const uint8_t* d_in = nullptr;
  if (side == 0) {
    sq_index* idx = const_cast<sq_index*>(s->idx);
    std::lock_guard<std::mutex> g(idx->col_mu);
    if (build_col_id < 0 || size_t(build_col_id) >= idx->columns.size())
      return fail(E, SQ_EINVAL, "unknown build column id %d", build_col_id);
    d_in = idx->columns[build_col_id].d_validity;  // nullptr = every row valid; NULL indices still clear bits
  } else {
    if (!probe_bitmap) return fail(E, SQ_EINVAL, "null probe validity bitmap");
    uint8_t* d = d_out + ((out_bytes + 31) & ~size_t(31));
    SQ_CUDA(E, cudaMemcpyAsync(d, probe_bitmap, in_bytes, cudaMemcpyHostToDevice, s->stream));
    d_in = d;
  }
  if ((rc = launch_gather_bits(s, d_in, ix, np, d_out, null_count_out))) return rc;
  if (np) {
    if (!out_bitmap) return fail(E, SQ_EINVAL, "null validity output");
    SQ_CUDA(E, cudaMemcpyAsync(out_bitmap, d_out, out_bytes, cudaMemcpyDeviceToHost, s->stream));
    SQ_CUDA(E, cudaStreamSynchronize(s->stream));
  }
  return SQ_OK;
}


// ---- bounded output (GPU analogue of interval_join_low_memory, IJ:1433-1530) -----------------------
SQ_API int32_t sq_stream_set_window(sq_stream* s, uint64_t pair_offset, uint64_t n_pairs) {
  if (!s) return SQ_EINVAL;
  if (!s->emitted) return fail(s->err, SQ_ESTATE, "sq_stream_set_window called without a preceding emit");
  if (pair_offset > s->n_pairs || n_pairs > s->n_pairs - pair_offset)
    return fail(s->err, SQ_EINVAL, "window [%llu, +%llu) exceeds the tile's %llu pairs", (unsigned long long)pair_offset,
                (unsigned long long)n_pairs, (unsigned long long)s->n_pairs);
  s->win_set = true;
  s->win_off = pair_offset;
  s->win_n = n_pairs;
  return SQ_OK;
}

SQ_API int32_t sq_fetch_pairs(sq_stream* s, uint32_t* left_idx_out, uint32_t* right_idx_out, uint64_t capacity) {
  if (!s) return SQ_EINVAL;
  ErrorSlot& E = s->err;
  if (!s->emitted) return fail(E, SQ_ESTATE, "sq_fetch_pairs called without a preceding emit");
  const uint64_t n = win_pairs(s);
  if (capacity < n) return fail(E, SQ_ECAPACITY, "output capacity %llu < %llu pairs", (unsigned long long)capacity,
                                (unsigned long long)n);
  if (n == 0) return SQ_OK;
  SQ_CUDA(E, cudaSetDevice(s->ctx->device));
  if (left_idx_out) {
    if (!s->d_last_left) return fail(E, SQ_ESTATE, "left indices are not on the device");
    SQ_CUDA(E, cudaMemcpyAsync(left_idx_out, s->d_last_left + s->win_off, n * 4, cudaMemcpyDeviceToHost, s->stream));
  }
  if (right_idx_out) {
    if (!s->d_last_right) return fail(E, SQ_ESTATE, "right indices are not on the device");
    SQ_CUDA(E, cudaMemcpyAsync(right_idx_out, s->d_last_right + s->win_off, n * 4, cudaMemcpyDeviceToHost, s->stream));
  }
  SQ_CUDA(E, cudaStreamSynchronize(s->stream));
  return SQ_OK;
}

SQ_API int32_t sq_stream_counts(sq_stream* s, uint32_t* counts_out) {
  if (!s) return SQ_EINVAL;
  ErrorSlot& E = s->err;
  if (!s->counted) return fail(E, SQ_ESTATE, "sq_stream_counts called without a counted tile");
  if (s->n_rows == 0) return SQ_OK;
  if (!counts_out) return fail(E, SQ_EINVAL, "null counts_out");
  SQ_CUDA(E, cudaSetDevice(s->ctx->device));
  SQ_CUDA(E, cudaMemcpyAsync(counts_out, s->d_cnt.p, size_t(s->n_rows) * 4, cudaMemcpyDeviceToHost, s->stream));
  SQ_CUDA(E, cudaStreamSynchronize(s->stream));
  return SQ_OK;
}

// ---- boundary helpers -------------------------------------------------------------------------
SQ_API int32_t sq_cast_i64_to_i32(sq_stream* s, const int64_t* values, uint64_t n, int64_t minus, int32_t* out) {
  if (!s) return SQ_EINVAL;
  ErrorSlot& E = s->err;
  if (n && (!values || !out)) return fail(E, SQ_EINVAL, "null cast buffer");
  if (n == 0) return SQ_OK;
  SQ_CUDA(E, cudaSetDevice(s->ctx->device));
  int rc;
  if ((rc = ensure(E, s->d_gather, n * 12 + 32, false))) return rc;
  auto* d_in = static_cast<int64_t*>(s->d_gather.p);
  auto* d_out = reinterpret_cast<int32_t*>(d_in + n);
  SQ_CUDA(E, cudaMemcpyAsync(d_in, values, n * 8, cudaMemcpyHostToDevice, s->stream));
  int64_t bad_value = 0;
  bool bad = false;
  if ((rc = launch_cast_i64(s, d_in, n, minus, d_out, &bad_value, &bad))) return rc;
  if (bad)  // exact text of the reference's failure (interval_join.rs:1959-1965)
    return fail(E, SQ_ECAST, "Arrow error: Cast error: Can't cast value %lld to type Int32", (long long)bad_value);
  SQ_CUDA(E, cudaMemcpyAsync(out, d_out, n * 4, cudaMemcpyDeviceToHost, s->stream));
  SQ_CUDA(E, cudaStreamSynchronize(s->stream));
  return SQ_OK;
}

SQ_API int32_t sq_pairs_digest_device(sq_stream* s, const uint32_t* d_left, const uint32_t* d_right, uint64_t n_pairs,
                                      uint64_t right_offset, uint64_t out3[3]) {
  if (!s || !out3) return SQ_EINVAL;
  if (n_pairs && (!d_left || !d_right)) return fail(s->err, SQ_EINVAL, "null pair buffer");
  SQ_CUDA(s->err, cudaSetDevice(s->ctx->device));
  return launch_digest(s, d_left, d_right, n_pairs, right_offset, out3);
}
