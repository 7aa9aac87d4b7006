#include "host_stage.cuh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace rsvdb {

namespace {
constexpr size_t kChunkBytes = 64u << 20;   // one pinned chunk
constexpr int kRing = 3;

// Minimal fork-join pool: run(n, fn) executes fn(0..n-1) on the workers plus the calling thread.
class Pool {
 public:
  explicit Pool(int nthreads) {
    for (int i = 0; i < nthreads; ++i) workers_.emplace_back([this] { loop(); });
  }
  ~Pool() {
    { std::lock_guard<std::mutex> g(m_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  void run(int n, const std::function<void(int)>& fn) {
    {
      std::lock_guard<std::mutex> g(m_);
      fn_ = &fn; next_ = 0; total_ = n; done_ = 0; ++epoch_;
    }
    cv_.notify_all();
    work();
    std::unique_lock<std::mutex> g(m_);
    cv_done_.wait(g, [this] { return done_ == total_; });
    fn_ = nullptr;
  }

 private:
  void work() {
    for (;;) {
      int i;
      {
        std::lock_guard<std::mutex> g(m_);
        if (!fn_ || next_ >= total_) return;
        i = next_++;
      }
      (*fn_)(i);
      {
        std::lock_guard<std::mutex> g(m_);
        if (++done_ == total_) cv_done_.notify_all();
      }
    }
  }
  void loop() {
    unsigned long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return stop_ || epoch_ != seen; });
        if (stop_) return;
        seen = epoch_;
      }
      work();
    }
  }
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, cv_done_;
  const std::function<void(int)>* fn_ = nullptr;
  int next_ = 0, total_ = 0, done_ = 0;
  unsigned long epoch_ = 0;
  bool stop_ = false;
};
}  // namespace

struct HostStager::Impl {
  double* buf[kRing] = {};
  cudaEvent_t ev[kRing] = {};
  bool busy[kRing] = {};
  Pool* pool = nullptr;
  int nthreads = 1;
  cudaError_t ensure() {
    if (!pool) {
      unsigned hc = std::thread::hardware_concurrency();
      nthreads = (int)std::max(1u, std::min(16u, hc ? hc : 4u));
      pool = new Pool(nthreads - 1);
    }
    for (int i = 0; i < kRing; ++i) {
      if (!buf[i]) { cudaError_t e = cudaHostAlloc((void**)&buf[i], kChunkBytes, cudaHostAllocDefault); if (e != cudaSuccess) return e; }
      if (!ev[i]) { cudaError_t e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming); if (e != cudaSuccess) return e; }
    }
    return cudaSuccess;
  }
};

HostStager::HostStager() : p_(new Impl()) {}
HostStager::~HostStager() {
  for (int i = 0; i < kRing; ++i) {
    if (p_->ev[i]) { cudaEventSynchronize(p_->ev[i]); cudaEventDestroy(p_->ev[i]); }
    if (p_->buf[i]) cudaFreeHost(p_->buf[i]);
  }
  delete p_->pool;
  delete p_;
}

bool HostStager::pageable(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return a.type == cudaMemoryTypeUnregistered;
}

cudaError_t HostStager::upload(cudaStream_t st, double* dst, long long ldd, const double* src, long long lds, long long rows, long long cols) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  cudaError_t e = p_->ensure(); if (e != cudaSuccess) return e;
  const long long chunk_elems = (long long)(kChunkBytes / sizeof(double));
  const long long rt = std::min(rows, chunk_elems);                       // rows per tile
  const long long ct = std::max<long long>(1, chunk_elems / rt);          // columns per tile
  int slot = 0;
  static const bool debug = getenv("RSVDB_STAGE_DEBUG") != nullptr;
  double t_wait = 0, t_pack = 0, t_issue = 0; int ntiles = 0;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  for (long long r0 = 0; r0 < rows; r0 += rt) {
    const long long nr = std::min(rt, rows - r0);
    for (long long c0 = 0; c0 < cols; c0 += ct) {
      const long long nc = std::min(ct, cols - c0);
      double t0 = debug ? now() : 0.0;
      if (p_->busy[slot]) { e = cudaEventSynchronize(p_->ev[slot]); if (e != cudaSuccess) return e; p_->busy[slot] = false; }
      double t1 = debug ? now() : 0.0;
      double* b = p_->buf[slot];
      const double* s0 = src + r0 + c0 * lds;
      // pack the tile (leading dimension nr): split by columns, or by row ranges when the tile is a few long columns
      const int T = p_->nthreads;
      if (nc >= T) {
        p_->pool->run(T, [&](int t) {
          const long long a = nc * t / T, z = nc * (t + 1) / T;
          if (lds == nr) { if (z > a) std::memcpy(b + a * nr, s0 + a * lds, (size_t)(z - a) * nr * sizeof(double)); }
          else for (long long c = a; c < z; ++c) std::memcpy(b + c * nr, s0 + c * lds, (size_t)nr * sizeof(double));
        });
      } else {
        p_->pool->run(T, [&](int t) {
          const long long a = nr * t / T, z = nr * (t + 1) / T;
          if (z > a) for (long long c = 0; c < nc; ++c) std::memcpy(b + c * nr + a, s0 + c * lds + a, (size_t)(z - a) * sizeof(double));
        });
      }
      double t2 = debug ? now() : 0.0;
      e = cudaMemcpy2DAsync(dst + r0 + c0 * ldd, (size_t)ldd * sizeof(double), b, (size_t)nr * sizeof(double), (size_t)nr * sizeof(double), (size_t)nc,
                            cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) return e;
      e = cudaEventRecord(p_->ev[slot], st); if (e != cudaSuccess) return e;
      p_->busy[slot] = true;
      if (debug) { const double t3 = now(); t_wait += t1 - t0; t_pack += t2 - t1; t_issue += t3 - t2; ++ntiles; }
      slot = (slot + 1) % kRing;
    }
  }
  if (debug) std::fprintf(stderr, "[rsvdb stage] %lld x %lld (lds %lld): %d tiles, wait %.2f ms, pack %.2f ms, issue %.2f ms, threads %d\n", rows, cols, lds,
                          ntiles, t_wait * 1e3, t_pack * 1e3, t_issue * 1e3, p_->nthreads);
  return cudaSuccess;
}

}  // namespace rsvdb
